/*
 * HEVC transforms: forward DCT/DST and inverse transform + add-to-predictor - function-select API.
 * Declaration-compatible with the reference's src/lib/residual_decode.h (:54, :56-74, :82, :84-102);
 * written afresh.
 */
#ifndef INCLUDED_hevcasm_residual_decode_h
#define INCLUDED_hevcasm_residual_decode_h

#include "hevcasm.h"
#include <assert.h>

#ifdef __cplusplus
extern "C" {
#endif

/* dst = clip8(pred + inverse_transform(coeffs)); coeffs is N*N contiguous, coeffs[v*N+u] */
typedef void hevcasm_inverse_transform_add(uint8_t *dst, ptrdiff_t stride_dst, const uint8_t *pred, ptrdiff_t stride_pred, const int16_t *coeffs);

typedef struct {
    hevcasm_inverse_transform_add *dst;    /* 4x4 DST-VII */
    hevcasm_inverse_transform_add *dct[4]; /* DCT 4, 8, 16, 32 */
} hevcasm_table_inverse_transform_add;

static inline hevcasm_inverse_transform_add **hevcasm_get_inverse_transform_add(hevcasm_table_inverse_transform_add *table, int trType, int log2TrafoSize)
{
    if (!trType) return &table->dct[log2TrafoSize - 2];
    assert(log2TrafoSize == 2);
    return &table->dst;
}

/* "encoder" selected a faster, non-conforming 32x32 in the reference (residual_decode.c:489-495); the GPU
 * path is exact for both values. */
void HEVCASM_API hevcasm_populate_inverse_transform_add(hevcasm_table_inverse_transform_add *table, hevcasm_instruction_set mask, int encoder);
void HEVCASM_API hevcasm_test_inverse_transform_add(int *error_count, hevcasm_instruction_set mask);

/* coeffs (N*N contiguous) = forward transform of the residual block at src (stride in int16 elements) */
typedef void hevcasm_transform(int16_t *coeffs, const int16_t *src, ptrdiff_t src_stride);

typedef struct {
    hevcasm_transform *dst;
    hevcasm_transform *dct[4];
} hevcasm_table_transform;

static inline hevcasm_transform **hevcasm_get_transform(hevcasm_table_transform *table, int trType, int log2TrafoSize)
{
    if (!trType) return &table->dct[log2TrafoSize - 2];
    assert(log2TrafoSize == 2);
    return &table->dst;
}

void HEVCASM_API hevcasm_populate_transform(hevcasm_table_transform *table, hevcasm_instruction_set mask);
void HEVCASM_API hevcasm_test_transform(int *error_count, hevcasm_instruction_set mask);

#ifdef __cplusplus
}
#endif

#endif
