"""ctypes description of the batched C ABI declared in include/hevcasm_batch.h.

One table, used twice: by `hevcasm_b200.lib` to bind libhevcasm_b200.so (the product), and by the test-only
`oracle/binding.py` to bind the CPU twins `oracle_drv_*` / `ref_drv_*`, which take the same arguments except that the
trailing `void *stream` is an `int threads`.
"""
import ctypes as C

P = C.c_void_p          # any data pointer (device pointer for the product, host pointer for the CPU twins)
PD = C.c_ssize_t        # ptrdiff_t
I = C.c_int
U32 = C.c_uint32
LL = C.c_longlong

# name -> argument types WITHOUT the trailing stream/threads argument
BATCH_ABI = {
    "sad_multiref_batch": [P, PD, P, PD, U32, P, I, P, I, P],
    "sad_batch": [P, PD, P, PD, U32, P, P, I, P],
    "sad_sweep_frames": [P, PD, P, PD, I, I, U32, I, I, I, I, I, PD, PD, P],
    "ssd_batch": [P, PD, P, PD, I, P, I, P],
    "ssd_frames": [P, PD, P, PD, I, I, I, I, PD, PD, P],
    "hadamard_satd_batch": [P, PD, P, PD, I, P, I, P],
    "hadamard_satd_frames": [P, PD, P, PD, I, I, I, I, PD, PD, P],
    "ssd_linear_batch": [P, PD, P, PD, I, I, P],
    "pred_uni_frames": [P, PD, P, PD, I, I, I, I, I, I, PD, PD],
    "pred_uni_batch": [P, PD, P, PD, I, P, I],
    "pred_bi_frames": [P, PD, P, P, PD, I, I, I, I, I, I, I, I, PD, PD],
    "pred_bi_batch": [P, PD, P, P, PD, I, P, I],
    "transform_batch": [P, P, PD, I, I, P, I],
    "transform_frames": [P, P, PD, I, I, I, I, I, PD],
    "inverse_transform_add_batch": [P, PD, P, PD, P, I, I, P, I],
    "inverse_transform_add_frames": [P, PD, P, PD, P, I, I, I, I, I, PD, PD],
    "quantize_batch": [P, P, I, I, I, I, I, P],
    "quantize_inverse_batch": [P, P, I, I, LL],
    "quantize_reconstruct_batch": [P, PD, P, PD, P, I, P, I],
    "quantize_reconstruct_frames": [P, PD, P, PD, P, I, I, I, I, PD, PD],
}

# entry points that exist only on the GPU side (no per-call CPU twin; tests compose the twins above)
GPU_ONLY_ABI = {
    "sad_sweep_pyramid_frames": [P, PD, P, PD, I, I, I, I, I, PD, PD, P, P, P, P],
    "sad_sweep_pyramid_best_frames": [P, PD, P, PD, I, I, I, I, I, PD, PD, P, P, P, P],
    "sad_sweep_pyramid_packed_frames": [P, PD, P, PD, I, I, I, I, I, PD, PD, P, P, P, P],
    "residual_pipeline_frames": [P, PD, P, P, P, PD, P, PD, I, I, I, I, I, I, I, I, I, I, PD, PD, PD],
    "transform_from_planes_frames": [P, P, PD, P, PD, I, I, I, I, I, PD, PD],
    "residual_from_planes_pipeline_frames": [P, PD, P, P, P, PD, P, PD, I, I, I, I, I, I, I, I, I, I, PD, PD, PD],
    "sad_list_frames": [P, PD, P, PD, P, I, PD, PD, P],
    "ssd_list_frames": [P, PD, P, PD, P, I, PD, PD, P],
    "hadamard_satd_list_frames": [P, PD, P, PD, P, P, PD, PD, P],
    "transform_list_frames": [P, P, PD, P, P, PD],
    "quantize_reconstruct_list_frames": [P, PD, P, PD, P, P, P, PD, PD],
    "inverse_transform_add_list_frames": [P, PD, P, PD, P, P, P, PD, PD],
    "pred_uni_list_frames": [P, PD, P, PD, I, P, I, PD, PD],
    "pred_bi_list_frames": [P, PD, P, P, PD, I, P, I, PD, PD],
    "pred_uni_frames_bounded": [P, PD, P, PD, I, I, I, I, I, I, PD, PD, PD, PD],
    "pred_bi_frames_bounded": [P, PD, P, P, PD, I, I, I, I, I, I, I, I, PD, PD, PD, PD],
}

# host-memory forms: hevcasm_<name>(hevcasm_cuda_context *ctx, ...host pointers...) - no stream argument
HOST_ABI = {
    "sad_sweep_pyramid_frames_host": [P, P, PD, P, PD, I, I, I, I, I, I, PD, PD, P, P, P, P],
    "sad_sweep_pyramid_best_frames_host": [P, P, PD, P, PD, I, I, I, I, I, I, PD, PD, P, P, P, P],
    "sad_sweep_pyramid_packed_frames_host": [P, P, PD, P, PD, I, I, I, I, I, I, PD, PD, P, P, P, P],
    "pred_uni_frames_host": [P, P, PD, P, PD, I, I, I, I, I, I, I, PD, PD],
    "residual_pipeline_frames_host": [P, P, PD, P, P, P, PD, P, PD, I, I, I, I, I, I, I, I, I, I, PD, PD, PD],
}


def HEVCASM_RECT(w, h):
    """reference hevcasm.h:156"""
    return (w << 8) | h
