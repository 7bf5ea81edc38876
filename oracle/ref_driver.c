/*
 * TEST INFRASTRUCTURE ONLY.  Compiled by oracle/build_ref.sh together with the reference's own
 * src/lib C files (in place, from /root/reference) into oracle/_ref/libhevcasm_cref.so.
 *
 * Binds oracle/batch_driver.inc to the reference's own function-select tables, populated exactly the way a
 * codec would (hevcasm_populate_* + hevcasm_get_*) with the mask HEVCASM_C_REF|HEVCASM_C_OPT - the only ISA
 * variants that can be built in an image without yasm/nasm.  Also exports flat per-block wrappers (ref_*) with
 * the oracle's argument order so tests can compare one block at a time.
 *
 * Optional: the 4-way 32x32 / 64x64 SAD can be routed through libvpx's AVX2 intrinsics file (the one piece of
 * hand-written SIMD in the tree that needs no assembler) for the CPU baseline: ref_drv_set_avx2_sad(1).
 */
#include "sad.h"
#include "ssd.h"
#include "pred_inter.h"
#include "residual_decode.h"
#include "quantize.h"
#include "hadamard.h"
#include "diff.h"

static const hevcasm_instruction_set k_mask = (hevcasm_instruction_set)(HEVCASM_C_REF | HEVCASM_C_OPT);

static hevcasm_table_sad t_sad;
static hevcasm_table_sad_multiref t_sad4;
static hevcasm_table_ssd t_ssd;
static hevcasm_table_pred_uni_8to8 t_uni;
static hevcasm_table_pred_bi_8to8 t_bi;
static hevcasm_table_transform t_fwd;
static hevcasm_table_inverse_transform_add t_inv;
static hevcasm_table_quantize t_q;
static hevcasm_table_quantize_inverse t_iq;
static hevcasm_table_quantize_reconstruct t_rec;
static hevcasm_table_hadamard_satd t_satd;
static int g_ready = 0;
static int g_avx2_sad = 0;

/* libvpx/vp9/encoder/x86/vp9_sad4d_intrin_avx2.c:13 and :83 */
void vp9_sad32x32x4d_avx2(uint8_t *src, int src_stride, uint8_t *ref[4], int ref_stride, uint32_t res[4]);
void vp9_sad64x64x4d_avx2(uint8_t *src, int src_stride, uint8_t *ref[4], int ref_stride, uint32_t res[4]);

static int ref_tables_init(void)
{
    if (g_ready) return 0;
    hevcasm_populate_sad(&t_sad, k_mask);
    hevcasm_populate_sad_multiref(&t_sad4, k_mask);
    hevcasm_populate_ssd(&t_ssd, k_mask);
    hevcasm_populate_pred_uni_8to8(&t_uni, k_mask);
    hevcasm_populate_pred_bi_8to8(&t_bi, k_mask);
    hevcasm_populate_transform(&t_fwd, k_mask);
    hevcasm_populate_inverse_transform_add(&t_inv, k_mask, 1);
    hevcasm_populate_quantize(&t_q, k_mask);
    hevcasm_populate_quantize_inverse(&t_iq, k_mask);
    hevcasm_populate_quantize_reconstruct(&t_rec, k_mask);
    hevcasm_populate_hadamard_satd(&t_satd, k_mask);
    g_ready = 1;
    return 0;
}

void ref_drv_set_avx2_sad(int on) { g_avx2_sad = on; }

int ref_sad(const uint8_t *src, ptrdiff_t ss, const uint8_t *ref, ptrdiff_t sr, uint32_t rect)
{
    ref_tables_init();
    return (*hevcasm_get_sad(&t_sad, (int)(rect >> 8), (int)(rect & 0xff)))(src, ss, ref, sr, rect);
}

void ref_sad_multiref_4(const uint8_t *src, ptrdiff_t ss, const uint8_t *ref[], ptrdiff_t sr, int sad[], uint32_t rect)
{
    ref_tables_init();
    if (g_avx2_sad && rect == HEVCASM_RECT(32, 32)) {
        vp9_sad32x32x4d_avx2((uint8_t *)src, (int)ss, (uint8_t **)ref, (int)sr, (uint32_t *)sad);
        return;
    }
    if (g_avx2_sad && rect == HEVCASM_RECT(64, 64)) {
        vp9_sad64x64x4d_avx2((uint8_t *)src, (int)ss, (uint8_t **)ref, (int)sr, (uint32_t *)sad);
        return;
    }
    (*hevcasm_get_sad_multiref(&t_sad4, 4, (int)(rect >> 8), (int)(rect & 0xff)))(src, ss, ref, sr, sad, rect);
}

int ref_ssd(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb, int log2)
{
    ref_tables_init();
    return (*hevcasm_get_ssd(&t_ssd, log2))(a, sa, b, sb, 1 << log2, 1 << log2);
}

void ref_pred_uni(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref, ptrdiff_t sr, int taps, int w, int h, int xFrac, int yFrac)
{
    ref_tables_init();
    (*hevcasm_get_pred_uni_8to8(&t_uni, taps, w, h, xFrac, yFrac))(dst, sd, ref, sr, w, h, xFrac, yFrac);
}

void ref_pred_bi(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref0, const uint8_t *ref1, ptrdiff_t sr, int taps, int w, int h,
                 int xFrac0, int yFrac0, int xFrac1, int yFrac1)
{
    ref_tables_init();
    (*hevcasm_get_pred_bi_8to8(&t_bi, taps, w, h, xFrac0, yFrac0, xFrac1, yFrac1))(dst, sd, ref0, ref1, sr, w, h, xFrac0, yFrac0,
                                                                                xFrac1, yFrac1);
}

void ref_transform(int16_t *coeffs, const int16_t *src, ptrdiff_t stride, int trType, int log2)
{
    ref_tables_init();
    (*hevcasm_get_transform(&t_fwd, trType, log2))(coeffs, src, stride);
}

void ref_inverse_transform_add(uint8_t *dst, ptrdiff_t sd, const uint8_t *pred, ptrdiff_t sp, const int16_t *coeffs, int trType,
                               int log2)
{
    ref_tables_init();
    (*hevcasm_get_inverse_transform_add(&t_inv, trType, log2))(dst, sd, pred, sp, coeffs);
}

int ref_quantize(int16_t *dst, const int16_t *src, int scale, int shift, int offset, int n)
{
    ref_tables_init();
    return (*hevcasm_get_quantize(&t_q))(dst, src, scale, shift, offset, n);
}

void ref_quantize_inverse(int16_t *dst, const int16_t *src, int scale, int shift, int n)
{
    ref_tables_init();
    (*hevcasm_get_quantize_inverse(&t_iq))(dst, src, scale, shift, n);
}

void ref_quantize_reconstruct(uint8_t *rec, ptrdiff_t sr, const uint8_t *pred, ptrdiff_t sp, const int16_t *res, int log2)
{
    ref_tables_init();
    (*hevcasm_get_quantize_reconstruct(&t_rec, log2))(rec, sr, pred, sp, res, 1 << log2);
}

int ref_hadamard_satd(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb, int log2)
{
    ref_tables_init();
    return (*hevcasm_get_hadamard_satd(&t_satd, log2))(a, sa, b, sb);
}

int ref_ssd_linear(const uint8_t *p0, const uint8_t *p1, int size)
{
    return hevcasm_get_ssd_linear(size, k_mask)(p0, p1, size);
}

int ref_pred_coefficient(int taps, int frac, int k)
{
    extern int hevcasm_pred_coefficient(int n, int fractionalPosition, int k);
    return hevcasm_pred_coefficient(taps, frac, k);
}

#define DRV(name) ref_drv_##name
#define BLK_INIT() ref_tables_init()
#define BLK_SAD(src, ss, ref, sr, rect) ref_sad(src, ss, ref, sr, rect)
#define BLK_SAD4(src, ss, refs, sr, sad, rect) ref_sad_multiref_4(src, ss, refs, sr, sad, rect)
#define BLK_SSD(a, sa, b, sb, log2) ref_ssd(a, sa, b, sb, log2)
#define BLK_SATD(a, sa, b, sb, log2) ref_hadamard_satd(a, sa, b, sb, log2)
#define BLK_SSD_LINEAR(p0, p1, size) ref_ssd_linear(p0, p1, size)
#define BLK_PRED_UNI(dst, sd, ref, sr, taps, w, h, xf, yf) ref_pred_uni(dst, sd, ref, sr, taps, w, h, xf, yf)
#define BLK_PRED_BI(dst, sd, r0, r1, sr, taps, w, h, xf0, yf0, xf1, yf1) \
    ref_pred_bi(dst, sd, r0, r1, sr, taps, w, h, xf0, yf0, xf1, yf1)
#define BLK_TRANSFORM(coeffs, src, stride, trType, log2) ref_transform(coeffs, src, stride, trType, log2)
#define BLK_INV_TRANSFORM_ADD(dst, sd, pred, sp, coeffs, trType, log2) ref_inverse_transform_add(dst, sd, pred, sp, coeffs, trType, log2)
#define BLK_QUANTIZE(dst, src, scale, shift, offset, n) ref_quantize(dst, src, scale, shift, offset, n)
#define BLK_QUANTIZE_INVERSE(dst, src, scale, shift, n) ref_quantize_inverse(dst, src, scale, shift, n)
#define BLK_QUANTIZE_RECONSTRUCT(rec, sr, pred, sp, res, log2) ref_quantize_reconstruct(rec, sr, pred, sp, res, log2)
#include "batch_driver.inc"
