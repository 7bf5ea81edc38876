/*
 * Sum of absolute differences - function-select API.
 * Declaration-compatible with the reference's src/lib/sad.h (types :50, :52-67, :95, :97-102; getters
 * :69-87, :104-108); written afresh.  Batched GPU entry points: hevcasm_batch.h.
 */
#ifndef INCLUDED_sad_h
#define INCLUDED_sad_h

#include "hevcasm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* SAD of one w x h block against one reference block; rect = HEVCASM_RECT(w, h). */
typedef int hevcasm_sad(const uint8_t *src, ptrdiff_t stride_src, const uint8_t *ref, ptrdiff_t stride_ref, uint32_t rect);

/* Eleven dedicated slots plus a catch-all, in the reference's member order (ABI of the table). */
typedef struct {
    hevcasm_sad *sad64x64;
    hevcasm_sad *sad64x32;
    hevcasm_sad *sad32x64;
    hevcasm_sad *sad32x32;
    hevcasm_sad *sad32x16;
    hevcasm_sad *sad16x32;
    hevcasm_sad *sad16x16;
    hevcasm_sad *sad16x8;
    hevcasm_sad *sad8x16;
    hevcasm_sad *sad8x8;
    hevcasm_sad *sad8x4;
    hevcasm_sad *sadGeneric;
} hevcasm_table_sad;

static inline hevcasm_sad **hevcasm_get_sad(hevcasm_table_sad *table, int width, int height)
{
    const int rect = HEVCASM_RECT(width, height);
    if (rect == HEVCASM_RECT(64, 64)) return &table->sad64x64;
    if (rect == HEVCASM_RECT(64, 32)) return &table->sad64x32;
    if (rect == HEVCASM_RECT(32, 64)) return &table->sad32x64;
    if (rect == HEVCASM_RECT(32, 32)) return &table->sad32x32;
    if (rect == HEVCASM_RECT(32, 16)) return &table->sad32x16;
    if (rect == HEVCASM_RECT(16, 32)) return &table->sad16x32;
    if (rect == HEVCASM_RECT(16, 16)) return &table->sad16x16;
    if (rect == HEVCASM_RECT(16, 8)) return &table->sad16x8;
    if (rect == HEVCASM_RECT(8, 16)) return &table->sad8x16;
    if (rect == HEVCASM_RECT(8, 8)) return &table->sad8x8;
    if (rect == HEVCASM_RECT(8, 4)) return &table->sad8x4;
    return &table->sadGeneric;
}

void HEVCASM_API hevcasm_populate_sad(hevcasm_table_sad *table, hevcasm_instruction_set mask);
hevcasm_test_function hevcasm_test_sad;

/* SAD of one block against several reference blocks sharing one stride; only ways == 4 exists. */
typedef void hevcasm_sad_multiref(const uint8_t *src, ptrdiff_t stride_src, const uint8_t *ref[], ptrdiff_t stride_ref, int sad[], uint32_t rect);

typedef struct {
    hevcasm_sad_multiref *lookup[16][16]; /* [(w>>2)-1][(h>>2)-1] */
    hevcasm_sad_multiref *sadGeneric_4;
} hevcasm_table_sad_multiref;

static inline hevcasm_sad_multiref **hevcasm_get_sad_multiref(hevcasm_table_sad_multiref *table, int ways, int width, int height)
{
    return ways == 4 ? &table->lookup[(width >> 2) - 1][(height >> 2) - 1] : 0;
}

void HEVCASM_API hevcasm_populate_sad_multiref(hevcasm_table_sad_multiref *table, hevcasm_instruction_set mask);
hevcasm_test_function hevcasm_test_sad_multiref;

#ifdef __cplusplus
}
#endif

#endif
