export RUN_ONE_EXP=1
for ORDER in raster ctu; do
  export RUN_ONE_ORDER=$ORDER
  for V in "8 3" "8 4" "4 4" "4 5"; do
    set -- $V
    echo -n "order=$ORDER cols=$1 minb=$2: "
    HEVCASM_LIST_COLS=$1 HEVCASM_LIST_MINB=$2 python tools/run_one.py pred_listf8 20
  done
done
for V in "4 3" "4 4"; do set -- $V; echo -n "bi cols=$1 minb=$2: "; HEVCASM_LIST_COLS=$1 HEVCASM_LIST_MINB=$2 python tools/run_one.py pred_bilistf8 20; done
for V in "8 3" "4 4"; do set -- $V; echo -n "64x64 cols=$1 minb=$2: "; HEVCASM_LIST_COLS=$1 HEVCASM_LIST_MINB=$2 python tools/run_one.py pred_listf64 20; done
echo -n "one frame 8x8: "; python tools/run_one.py pred_list8 20; echo -n "one frame 64x64: "; python tools/run_one.py pred_list64 20
