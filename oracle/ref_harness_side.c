/*
 * TEST INFRASTRUCTURE ONLY.  One side of oracle/ref_harness.c: the library self-test's cases (hevcasm_b200/csrc/selftest_cases.inc)
 * compiled against the function-select API of ONE library, reached through function pointers resolved with dlsym.  Built twice by
 * oracle/Makefile: -DSIDE=ref (oracle/_ref/libhevcasm_cref.so, the reference's own compiled C path) and -DSIDE=cuda
 * (hevcasm_b200/libhevcasm_b200.so).  The headers are the REFERENCE's (-I /root/reference/src/lib): both libraries implement them.
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sad.h"
#include "ssd.h"
#include "pred_inter.h"
#include "residual_decode.h"
#include "quantize.h"
#include "hadamard.h"
#include "diff.h"

#define SIDE_CAT2(a, b) a##_##b
#define SIDE_CAT(a, b) SIDE_CAT2(a, b)
#define S(name) SIDE_CAT(SIDE, name)

static void (*p_sad)(hevcasm_table_sad *, hevcasm_instruction_set);
static void (*p_sad_multiref)(hevcasm_table_sad_multiref *, hevcasm_instruction_set);
static void (*p_ssd)(hevcasm_table_ssd *, hevcasm_instruction_set);
static void (*p_quantize_inverse)(hevcasm_table_quantize_inverse *, hevcasm_instruction_set);
static void (*p_quantize)(hevcasm_table_quantize *, hevcasm_instruction_set);
static void (*p_quantize_reconstruct)(hevcasm_table_quantize_reconstruct *, hevcasm_instruction_set);
static void (*p_pred_uni)(hevcasm_table_pred_uni_8to8 *, hevcasm_instruction_set);
static void (*p_pred_bi)(hevcasm_table_pred_bi_8to8 *, hevcasm_instruction_set);
static void (*p_inverse_transform_add)(hevcasm_table_inverse_transform_add *, hevcasm_instruction_set, int);
static void (*p_transform)(hevcasm_table_transform *, hevcasm_instruction_set);
static void (*p_hadamard_satd)(hevcasm_table_hadamard_satd *, hevcasm_instruction_set);
static hevcasm_ssd_linear *(*p_get_ssd_linear)(int, hevcasm_instruction_set);

#define hevcasm_populate_sad (*p_sad)
#define hevcasm_populate_sad_multiref (*p_sad_multiref)
#define hevcasm_populate_ssd (*p_ssd)
#define hevcasm_populate_quantize_inverse (*p_quantize_inverse)
#define hevcasm_populate_quantize (*p_quantize)
#define hevcasm_populate_quantize_reconstruct (*p_quantize_reconstruct)
#define hevcasm_populate_pred_uni_8to8 (*p_pred_uni)
#define hevcasm_populate_pred_bi_8to8 (*p_pred_bi)
#define hevcasm_populate_inverse_transform_add (*p_inverse_transform_add)
#define hevcasm_populate_transform (*p_transform)
#define hevcasm_populate_hadamard_satd (*p_hadamard_satd)
#define hevcasm_get_ssd_linear (*p_get_ssd_linear)

#include "selftest_cases.inc"

int S(init)(void *dl)
{
    int missing = 0;
#define BIND(ptr, sym) do { *(void **)&ptr = dlsym(dl, sym); if (!ptr) { fprintf(stderr, "ref_harness: %s not found\n", sym); ++missing; } } while (0)
    BIND(p_sad, "hevcasm_populate_sad");
    BIND(p_sad_multiref, "hevcasm_populate_sad_multiref");
    BIND(p_ssd, "hevcasm_populate_ssd");
    BIND(p_quantize_inverse, "hevcasm_populate_quantize_inverse");
    BIND(p_quantize, "hevcasm_populate_quantize");
    BIND(p_quantize_reconstruct, "hevcasm_populate_quantize_reconstruct");
    BIND(p_pred_uni, "hevcasm_populate_pred_uni_8to8");
    BIND(p_pred_bi, "hevcasm_populate_pred_bi_8to8");
    BIND(p_inverse_transform_add, "hevcasm_populate_inverse_transform_add");
    BIND(p_transform, "hevcasm_populate_transform");
    BIND(p_hadamard_satd, "hevcasm_populate_hadamard_satd");
    BIND(p_get_ssd_linear, "hevcasm_get_ssd_linear");
#undef BIND
    return missing;
}

int S(total)(void) { return (int)KAT_TOTAL; }

/* every case of the self-test on the inputs of `salt`; digests into out[KAT_TOTAL]; returns the number of empty slots met */
int S(run)(uint64_t salt, int mask, uint64_t *out)
{
    const hevcasm_instruction_set m = (hevcasm_instruction_set)mask;
    int missing = 0;
    kat_seed_salt = salt;
    memset(out, 0, sizeof(uint64_t) * KAT_TOTAL);
    missing += kat_run_sad(m, out);
    missing += kat_run_sad_multiref(m, out);
    missing += kat_run_ssd(m, out);
    missing += kat_run_quantize_inverse(m, out);
    missing += kat_run_quantize(m, out);
    missing += kat_run_quantize_reconstruct(m, out);
    missing += kat_run_pred_uni(m, out);
    missing += kat_run_pred_bi(m, out);
    missing += kat_run_inverse_transform_add(m, out);
    missing += kat_run_transform(m, out);
    missing += kat_run_hadamard_satd(m, out);
    missing += kat_run_ssd_linear(m, out);
    return missing;
}
