# ncu captures of the default-dispatch kernels (run on the GPU box).  Each capture only after the same command has exited 0 without ncu; the
# reports are summarised on the box (text + counters JSON) and deleted, since gpurun brings back at most 64 MiB.
#   names: <run_one name>:<kernel regex>:<bench.py kernels entry or "sad_pyramid">:<frames>
mkdir -p gpurun_out/prof
export NCU_COUNTERS_OUT=gpurun_out/prof/r02_kernel_counters.json
W=3840; H=2160
for SPEC in "$@"; do
  IFS=: read -r K R NAME NF <<< "$SPEC"
  export RUN_ONE_NF=$NF
  timeout 120 python tools/run_one.py $K 3 > gpurun_out/prof/$K.time 2>&1 || { echo "$K failed without ncu"; cat gpurun_out/prof/$K.time; continue; }
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$R -s 3 -c 1 -f -o gpurun_out/prof/$K python tools/run_one.py $K 3 > gpurun_out/prof/$K.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof/$K.ncu-rep 2>/dev/null | cut -c1-400 > gpurun_out/prof/r02_ncu_$K.txt
  NCU_COUNTERS_SOURCE="profiles/r02_ncu_$K.txt (ncu --set full, $NF 4K frames)" python tools/ncu_counters.py "$NAME=gpurun_out/prof/$K.ncu-rep:$((NF*W*H)):$NF" | cut -c1-200
  rm -f gpurun_out/prof/$K.ncu-rep
  cat gpurun_out/prof/$K.time
done
