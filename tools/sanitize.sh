# usage (on the GPU box, ONE tool per gpurun call): bash tools/sanitize.sh memcheck|racecheck|synccheck|initcheck
# Runs the parity tests that drive the producer / consumer kernels (mbarriers, TMA boxes, TMEM) and the TMA-staged SAD kernels under
# compute-sanitizer; the log goes to gpurun_out/r02_sanitizer_<tool>.txt (copied to profiles/ afterwards).
TOOL=$1
export PYTORCH_NO_CUDA_MEMORY_CACHING=1     # exact allocations: an out-of-bounds access cannot hide inside torch's pool
SEL='tensor_core or bounded or lists_over_frames or (uni_planes_all_fractions and 200) or bi_planes or pyramid or host or forward_frames or inverse_frames or pipeline'
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 7 --print-limit 20 \
    python -m pytest tests/test_gpu_pred.py tests/test_gpu_transform.py tests/test_gpu_sad.py tests/test_gpu_host_forms.py tests/test_gpu_pipeline.py -m gpu -q -x -k "$SEL" \
    > gpurun_out/r02_sanitizer_$TOOL.txt 2>&1
echo "exit code $?" >> gpurun_out/r02_sanitizer_$TOOL.txt
grep -E "ERROR SUMMARY|passed|failed|exit code|Error" gpurun_out/r02_sanitizer_$TOOL.txt | head -20
