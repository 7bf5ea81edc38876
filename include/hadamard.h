/*
 * Hadamard-transformed SAD (SATD) - function-select API.
 * Declaration-compatible with the reference's src/lib/hadamard.h (:55-66); written afresh.
 * Batched GPU entry points: hevcasm_batch.h.
 */
#ifndef INCLUDED_hadamard_h
#define INCLUDED_hadamard_h

#include "hevcasm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* SATD of one N x N block pair, N = 2, 4 or 8 chosen by the table slot */
typedef int hevcasm_hadamard_satd(const uint8_t *srcA, ptrdiff_t stride_srcA, const uint8_t *srcB, ptrdiff_t stride_srcB);

typedef struct {
    hevcasm_hadamard_satd *satd[3]; /* 2x2, 4x4, 8x8 */
} hevcasm_table_hadamard_satd;

static inline hevcasm_hadamard_satd **hevcasm_get_hadamard_satd(hevcasm_table_hadamard_satd *table, int log2TrafoSize)
{
    return &table->satd[log2TrafoSize - 1];
}

void HEVCASM_API hevcasm_populate_hadamard_satd(hevcasm_table_hadamard_satd *table, hevcasm_instruction_set mask);
void HEVCASM_API hevcasm_test_hadamard_satd(int *error_count, hevcasm_instruction_set mask);

#ifdef __cplusplus
}
#endif

#endif
