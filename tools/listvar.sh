export RUN_ONE_EXP=1
for L in 4 8 16; do
  for K in pred_listf8 pred_listf16 pred_listf64 pred_bilistf8 pred_bilistf64; do echo -n "ldb=$L "; HEVCASM_LIST_LDB=$L python tools/run_one.py $K 20; done
done
