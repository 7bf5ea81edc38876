/*
 * Linear sum of squared differences over a flat run of samples - function-select API.
 * Declaration-compatible with the reference's src/lib/diff.h (:48-54); written afresh.
 */
#ifndef INCLUDED_diff_h
#define INCLUDED_diff_h

#include "hevcasm.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef int hevcasm_ssd_linear(const uint8_t *src0, const uint8_t *src1, int size);

/* returns the implementation for `size` under `mask`, or 0 (reference diff.c:54-63); only HEVCASM_CUDA is implemented here */
hevcasm_ssd_linear *HEVCASM_API hevcasm_get_ssd_linear(int size, hevcasm_instruction_set mask);

hevcasm_test_function hevcasm_test_ssd_linear;

#ifdef __cplusplus
}
#endif

#endif
