// Entry points declared in hevcasm_batch.h whose kernels are not written yet: they fail loudly (never compute).
#include "common.cuh"
#define NOT_YET return HEVCASM_ERR_ARGUMENT
extern "C" {
int hevcasm_pred_uni_frames_host(hevcasm_cuda_context *, uint8_t *, ptrdiff_t, const uint8_t *, ptrdiff_t, int, int, int, int, int, int, int, ptrdiff_t, ptrdiff_t) { NOT_YET; }
int hevcasm_residual_pipeline_frames_host(hevcasm_cuda_context *, uint8_t *, ptrdiff_t, int16_t *, int32_t *, const int16_t *, ptrdiff_t, const uint8_t *, ptrdiff_t, int, int, int, int, int, int, int, int, int, int, ptrdiff_t, ptrdiff_t, ptrdiff_t) { NOT_YET; }
}
