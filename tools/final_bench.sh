# the round's measurement runs on one GPU (run on the GPU box): headline + kernel table, reference arm, configs[4], launch list
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$?"
python bench.py --config 8k64 --steps 5 --warmup 3 > gpurun_out/r02_bench_8k64.json 2> gpurun_out/r02_bench_8k64.err; echo "8k64 rc=$?"
python bench.py --config 8k64 --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_8k64_reference.json 2> /dev/null; echo "8k64 reference rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --kernel-launches 1 > gpurun_out/plain_launchlist.json 2> gpurun_out/plain_launchlist.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --kernel-launches 1 > /dev/null 2> gpurun_out/ncu_launchlist.err; echo "launch list rc=$?"
