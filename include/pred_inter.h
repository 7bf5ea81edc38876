/*
 * HEVC non-weighted inter prediction (luma 8-tap, chroma 4-tap; uni and bi) - function-select API.
 * Declaration-compatible with the reference's src/lib/pred_inter.h (:56, :58-67, :76, :78-88); written afresh.
 * As in the reference (:42) a function may write to the right of the destination block, up to the next
 * multiple of 16 bytes.
 */
#ifndef INCLUDED_hevcasm_prediction_inter_h
#define INCLUDED_hevcasm_prediction_inter_h

#include "hevcasm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ref points at the integer-MV sample of the block origin; xFrac/yFrac in 1/4 (luma) or 1/8 (chroma) units */
typedef void HEVCASM_API hevcasm_pred_uni_8to8(uint8_t *dst, ptrdiff_t stride_dst, const uint8_t *ref, ptrdiff_t stride_ref, int nPbW, int nPbH, int xFrac, int yFrac);

typedef struct {
    hevcasm_pred_uni_8to8 *p[2][9][2][2]; /* [taps/4-1][ceil(w/taps)][xFrac!=0][yFrac!=0] */
} hevcasm_table_pred_uni_8to8;

static inline hevcasm_pred_uni_8to8 **hevcasm_get_pred_uni_8to8(hevcasm_table_pred_uni_8to8 *table, int taps, int w, int h, int xFrac, int yFrac)
{
    (void)h;
    return &table->p[taps / 4 - 1][(w + taps - 1) / taps][xFrac != 0][yFrac != 0];
}

void HEVCASM_API hevcasm_populate_pred_uni_8to8(hevcasm_table_pred_uni_8to8 *table, hevcasm_instruction_set mask);
hevcasm_test_function hevcasm_test_pred_uni;

/* both references share one stride */
typedef void hevcasm_pred_bi_8to8(uint8_t *dst0, ptrdiff_t stride_dst, const uint8_t *ref0, const uint8_t *ref1, ptrdiff_t stride_ref, int nPbW, int nPbH, int xFrac0, int yFrac0, int xFrac1, int yFrac1);

typedef struct {
    hevcasm_pred_bi_8to8 *p[2][5][2]; /* [taps/4-1][ceil(w/(2*taps))][any fraction != 0] */
} hevcasm_table_pred_bi_8to8;

static inline hevcasm_pred_bi_8to8 **hevcasm_get_pred_bi_8to8(hevcasm_table_pred_bi_8to8 *table, int taps, int w, int h, int xFracA, int yFracA, int xFracB, int yFracB)
{
    (void)h;
    return &table->p[taps / 4 - 1][(w + 2 * taps - 1) / (2 * taps)][(xFracA | yFracA | xFracB | yFracB) != 0];
}

void HEVCASM_API hevcasm_populate_pred_bi_8to8(hevcasm_table_pred_bi_8to8 *table, hevcasm_instruction_set mask);
hevcasm_test_function hevcasm_test_pred_bi;

#ifdef __cplusplus
}
#endif

#endif
