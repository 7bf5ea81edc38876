# usage: bash tools/scale_bench.sh N   (on an N-GPU box): both configs of bench.py at N ranks, launched the way the driver launches them
N=$1
P=$((29500 + N))
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_scale_${N}gpu.json 2> gpurun_out/r02_scale_${N}gpu.err; echo "4k_sad rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+50)) bench.py --gpus $N --config 8k64 --steps 5 --warmup 3 > gpurun_out/r02_scale_${N}gpu_8k64.json 2> gpurun_out/r02_scale_${N}gpu_8k64.err; echo "8k64 rc=$?"
tail -2 gpurun_out/r02_scale_${N}gpu.err | cut -c1-300
cat gpurun_out/r02_scale_${N}gpu.json | cut -c1-600
cat gpurun_out/r02_scale_${N}gpu_8k64.json | cut -c1-900
