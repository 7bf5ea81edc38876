timeout 120 ./tools/tmem_probe > gpurun_out/tmem_probe.txt 2>&1
for KV in pred_bi:pred_vh inv32:big_inv inv16:big_inv fwd32:fwd_umma pipe8:small_pipeline pred_hv:pred_vh; do
  K=${KV%%:*}; R=${KV##*:}
  timeout 120 python tools/run_one.py $K 3 > gpurun_out/r02_base_$K.time 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:$R -s 3 -c 1 -f -o gpurun_out/r02_base_$K python tools/run_one.py $K 3 > gpurun_out/r02_base_$K.log 2>&1
  cat gpurun_out/r02_base_$K.time
done
tail -14 gpurun_out/tmem_probe.txt
