# usage: bash tools/prof_one.sh <tag> <run_one name>:<kernel regex> ...   -> gpurun_out/<tag>_<name>.ncu-rep (+ .time)
TAG=$1; shift
for KV in "$@"; do
  K=${KV%%:*}; R=${KV##*:}
  timeout 120 python tools/run_one.py $K 3 > gpurun_out/${TAG}_$K.time 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:$R -s 3 -c 1 -f -o gpurun_out/${TAG}_$K python tools/run_one.py $K 3 > gpurun_out/${TAG}_$K.log 2>&1
  cat gpurun_out/${TAG}_$K.time
done
