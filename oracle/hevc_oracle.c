/*
 * TEST INFRASTRUCTURE ONLY - see hevc_oracle.h.  Parity status: PINNED against the reference's
 * compiled C path (tests/test_oracle_vs_reference.py, tests/golden/).
 *
 * All arithmetic is 32-bit two's-complement int with arithmetic right shift, exactly like the
 * reference's C (conventions).  Stores to int16_t truncate (wrap) unless a clip
 * is written explicitly.
 */
#include "hevc_oracle.h"

#include <stdlib.h>
#include <string.h>

static inline int clip3(int lo, int hi, int x) { return x < lo ? lo : (x > hi ? hi : x); }

/* ------------------------------------------------------------------ SAD / SSD */

/* sad.c:47-60: sum over the w x h rectangle of |src - ref| */
int oracle_sad(const uint8_t *src, ptrdiff_t ss, const uint8_t *ref, ptrdiff_t sr, uint32_t rect)
{
    const int w = (int)(rect >> 8), h = (int)(rect & 0xff);
    int s = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            s += abs((int)src[y * ss + x] - (int)ref[y * sr + x]);
    return s;
}

/* sad.c:101-121: the same against four reference pointers sharing one stride */
void oracle_sad_multiref_4(const uint8_t *src, ptrdiff_t ss, const uint8_t *ref[], ptrdiff_t sr, int sad[], uint32_t rect)
{
    for (int i = 0; i < 4; ++i) sad[i] = oracle_sad(src, ss, ref[i], sr, rect);
}

/* ssd.c:43-55 */
int oracle_ssd(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb, int w, int h)
{
    int s = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int d = (int)a[y * sa + x] - (int)b[y * sb + x];
            s += d * d;
        }
    return s;
}

/* hadamard.c:75-131: D = A - B; T = H D H^T with the +-1 Sylvester matrix H (the order of the outputs is irrelevant to the sum);
 * result = (N/4 + sum |T|) / (N/2) in integer arithmetic */
int oracle_hadamard_satd(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb, int log2)
{
    const int n = 1 << log2;
    int d[8][8], t[8][8];
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x) d[y][x] = (int)a[y * sa + x] - (int)b[y * sb + x];
    /* H[k][x] = (-1)^popcount(k & x) */
    for (int y = 0; y < n; ++y)
        for (int k = 0; k < n; ++k) {
            int s = 0;
            for (int x = 0; x < n; ++x) s += (__builtin_popcount(k & x) & 1) ? -d[y][x] : d[y][x];
            t[y][k] = s;
        }
    int sum = n / 4;
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j) {
            int s = 0;
            for (int y = 0; y < n; ++y) s += (__builtin_popcount(j & y) & 1) ? -t[y][k] : t[y][k];
            sum += abs(s);
        }
    return sum / (n / 2);
}

/* diff.c:45-54 */
int oracle_ssd_linear(const uint8_t *p0, const uint8_t *p1, int size)
{
    int s = 0;
    for (int i = 0; i < size; ++i) {
        const int d = (int)p0[i] - (int)p1[i];
        s += d * d;
    }
    return s;
}

/* ------------------------------------------------------------------ interpolation */

/* HEVC (H.265 8.5.3.3.3) interpolation filters; equal to pred_inter.c:57-63 and :69-79 */
static const int8_t k_luma[4][8] = {
    {0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1}};
static const int8_t k_chroma[8][4] = {{0, 64, 0, 0},  {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4},
                                      {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2}};

int oracle_pred_coefficient(int taps, int frac, int k) { return taps == 8 ? k_luma[frac][k] : k_chroma[frac][k]; }

/*
 * One separable FIR pass, pred_inter.c:90-138 (hevcasm_pred_uni_generic):
 *   a = ((add << shift) >> 1) + sum_k c[k] * src[x + (k - taps/2 + 1) * tap_stride];  a >>= shift;
 *   16-bit destination: truncating store;  8-bit destination: Clip3(0,255).
 */
static void fir_pass(void *dst, int dst16, ptrdiff_t sd, const void *src, int src16, ptrdiff_t ss, int w, int h,
                     ptrdiff_t tap_stride, int taps, int frac, int shift, int add)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int a = (add << shift) >> 1;
            for (int k = 0; k < taps; ++k) {
                const ptrdiff_t o = y * ss + x + (k - taps / 2 + 1) * tap_stride;
                const int s = src16 ? ((const int16_t *)src)[o] : ((const uint8_t *)src)[o];
                a += oracle_pred_coefficient(taps, frac, k) * s;
            }
            a >>= shift;
            if (dst16) ((int16_t *)dst)[y * sd + x] = (int16_t)a;
            else ((uint8_t *)dst)[y * sd + x] = (uint8_t)clip3(0, 255, a);
        }
}

/* pred_inter.c:141-228 and the C branch of the selector :231-292 */
void oracle_pred_uni(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref, ptrdiff_t sr, int taps, int w, int h, int xFrac, int yFrac)
{
    if (!xFrac && !yFrac) { /* full-pel: row copies, :141-151 */
        for (int y = 0; y < h; ++y) memcpy(dst + y * sd, ref + y * sr, (size_t)w);
    } else if (xFrac && !yFrac) { /* :154-159 / :203-208: (sum+32)>>6, clip */
        fir_pass(dst, 0, sd, ref, 0, sr, w, h, 1, taps, xFrac, 6, 1);
    } else if (!xFrac && yFrac) { /* :162-167 / :211-216 */
        fir_pass(dst, 0, sd, ref, 0, sr, w, h, sr, taps, yFrac, 6, 1);
    } else { /* :170-179 / :219-228: exact int16 H pass over h+taps-1 rows, then (sum+2048)>>12 */
        const int above = taps / 2 - 1;
        int16_t *mid = (int16_t *)malloc(sizeof(int16_t) * (size_t)(h + taps - 1) * (size_t)w);
        fir_pass(mid, 1, w, ref - above * sr, 0, sr, w, h + taps - 1, 1, taps, xFrac, 0, 0);
        fir_pass(dst, 0, sd, mid + above * w, 1, w, w, h, w, taps, yFrac, 12, 1);
        free(mid);
    }
}

/* pred_inter.c:490-530: per reference H (shift 0) then V (>>6, truncating int16), then (A+B+64)>>7 clipped.
 * Both passes always run, a zero fraction being the {64} filter. */
void oracle_pred_bi(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref0, const uint8_t *ref1, ptrdiff_t sr, int taps, int w, int h,
                    int xFrac0, int yFrac0, int xFrac1, int yFrac1)
{
    const int above = taps / 2 - 1;
    const size_t plane = (size_t)(h + taps - 1) * (size_t)w;
    int16_t *mid = (int16_t *)malloc(sizeof(int16_t) * plane * 3);
    int16_t *v[2] = {mid + plane, mid + 2 * plane};
    const uint8_t *ref[2] = {ref0, ref1};
    const int xf[2] = {xFrac0, xFrac1}, yf[2] = {yFrac0, yFrac1};
    for (int i = 0; i < 2; ++i) {
        fir_pass(mid, 1, w, ref[i] - above * sr, 0, sr, w, h + taps - 1, 1, taps, xf[i], 0, 0);
        fir_pass(v[i], 1, w, mid + above * w, 1, w, w, h, w, taps, yf[i], 6, 0);
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            dst[y * sd + x] = (uint8_t)clip3(0, 255, ((int)v[0][y * w + x] + (int)v[1][y * w + x] + 64) >> 7);
    free(mid);
}

/* ------------------------------------------------------------------ transforms */

/*
 * T_N[k][x] = round-to-HEVC-integer of 64*sqrt(2)*cos(k*(2x+1)*pi/(2N)) (64 for k = 0), generated from the first
 * column of the 32-point matrix by cosine symmetry; equal to the literal tables at residual_decode.c:623-629 (4),
 * :662-672 (8), :719-735 (16), :795-826 (32).  The 4x4 DST matrix is H.265 (8-xx), equal to the expressions at
 * residual_decode.c:592-611.
 */
int oracle_transform_matrix(int trType, int N, int k, int x)
{
    static const int8_t col[33] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                                   61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9,  4,  0};
    static const int8_t dst4[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};
    if (trType) return dst4[k][x];
    int m = (k * (32 / N) * (2 * x + 1)) % 128, sign = 1;
    if (m > 64) m = 128 - m;
    if (m > 32) { m = 64 - m; sign = -1; }
    return sign * col[m];
}

static const int k_fwd_shift[4][2] = {{1, 8}, {2, 9}, {3, 10}, {4, 11}}; /* residual_decode.c:855-892 */

/* Forward: stage(src, stride, s1) then stage(temp, N, s2), each stage = transpose((T * row + 2^(s-1)) >> s), truncating
 * int16 store (no clip) - residual_decode.c:592-852. */
void oracle_transform(int16_t *coeffs, const int16_t *src, ptrdiff_t stride, int trType, int log2)
{
    const int N = 1 << log2;
    int16_t tmp[32 * 32];
    const int s1 = k_fwd_shift[log2 - 2][0], s2 = k_fwd_shift[log2 - 2][1];
    for (int y = 0; y < N; ++y)     /* stage 1: rows of the residual -> tmp[u][y] */
        for (int u = 0; u < N; ++u) {
            int a = 1 << (s1 - 1);
            for (int x = 0; x < N; ++x) a += oracle_transform_matrix(trType, N, u, x) * src[y * stride + x];
            tmp[u * N + y] = (int16_t)(a >> s1);
        }
    for (int u = 0; u < N; ++u)     /* stage 2: rows of tmp -> coeffs[v][u] */
        for (int v = 0; v < N; ++v) {
            int a = 1 << (s2 - 1);
            for (int y = 0; y < N; ++y) a += oracle_transform_matrix(trType, N, v, y) * tmp[u * N + y];
            coeffs[v * N + u] = (int16_t)(a >> s2);
        }
}

/* Inverse: two transposing stages with shifts 7 and 12, each clipping to int16 (residual_decode.c:69-347), then
 * dst = clip8((int16_t)(pred + res)) (hevcasm_add_residual / hevcasm_clip, :350-368). */
void oracle_inverse_transform_add(uint8_t *dst, ptrdiff_t sd, const uint8_t *pred, ptrdiff_t sp, const int16_t *coeffs,
                                  int trType, int log2)
{
    const int N = 1 << log2;
    int16_t b[32 * 32], r[32 * 32];
    for (int u = 0; u < N; ++u)     /* stage 1: column u of coeffs -> row u of b */
        for (int y = 0; y < N; ++y) {
            int a = 64;
            for (int v = 0; v < N; ++v) a += oracle_transform_matrix(trType, N, v, y) * coeffs[v * N + u];
            b[u * N + y] = (int16_t)clip3(-32768, 32767, a >> 7);
        }
    for (int y = 0; y < N; ++y)     /* stage 2: column y of b -> row y of r */
        for (int x = 0; x < N; ++x) {
            int a = 2048;
            for (int u = 0; u < N; ++u) a += oracle_transform_matrix(trType, N, u, x) * b[u * N + y];
            r[y * N + x] = (int16_t)clip3(-32768, 32767, a >> 12);
        }
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) {
            const int16_t s = (int16_t)((int)pred[y * sp + x] + (int)r[y * N + x]);
            dst[y * sd + x] = (uint8_t)clip3(0, 255, s);
        }
}

/* ------------------------------------------------------------------ quantisation */

/* quantize.c:160-186: returns the OR of all outputs */
int oracle_quantize(int16_t *dst, const int16_t *src, int scale, int shift, int offset, int n)
{
    const int off = offset << (shift - 16);
    int cbf = 0;
    for (int i = 0; i < n; ++i) {
        const int x = src[i];
        int q = ((x < 0 ? -x : x) * scale + off) >> shift;
        if (x < 0) q = -q;
        q = clip3(-32768, 32767, q);
        cbf |= q;
        dst[i] = (int16_t)q;
    }
    return cbf;
}

/* quantize.c:53-62 */
void oracle_quantize_inverse(int16_t *dst, const int16_t *src, int scale, int shift, int n)
{
    for (int i = 0; i < n; ++i) dst[i] = (int16_t)clip3(-32768, 32767, (src[i] * scale + (1 << (shift - 1))) >> shift);
}

/* quantize.c:292-302 */
void oracle_quantize_reconstruct(uint8_t *rec, ptrdiff_t sr, const uint8_t *pred, ptrdiff_t sp, const int16_t *res, int n)
{
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x) rec[y * sr + x] = (uint8_t)clip3(0, 255, (int)pred[y * sp + x] + (int)res[y * n + x]);
}

/* ------------------------------------------------------------------ batch drivers bound to the oracle */

#define DRV(name) oracle_drv_##name
#define BLK_INIT() 0
#define BLK_SAD(src, ss, ref, sr, rect) oracle_sad(src, ss, ref, sr, rect)
#define BLK_SAD4(src, ss, refs, sr, sad, rect) oracle_sad_multiref_4(src, ss, refs, sr, sad, rect)
#define BLK_SSD(a, sa, b, sb, log2) oracle_ssd(a, sa, b, sb, 1 << (log2), 1 << (log2))
#define BLK_SATD(a, sa, b, sb, log2) oracle_hadamard_satd(a, sa, b, sb, log2)
#define BLK_SSD_LINEAR(p0, p1, size) oracle_ssd_linear(p0, p1, size)
#define BLK_PRED_UNI(dst, sd, ref, sr, taps, w, h, xf, yf) oracle_pred_uni(dst, sd, ref, sr, taps, w, h, xf, yf)
#define BLK_PRED_BI(dst, sd, r0, r1, sr, taps, w, h, xf0, yf0, xf1, yf1) \
    oracle_pred_bi(dst, sd, r0, r1, sr, taps, w, h, xf0, yf0, xf1, yf1)
#define BLK_TRANSFORM(coeffs, src, stride, trType, log2) oracle_transform(coeffs, src, stride, trType, log2)
#define BLK_INV_TRANSFORM_ADD(dst, sd, pred, sp, coeffs, trType, log2) \
    oracle_inverse_transform_add(dst, sd, pred, sp, coeffs, trType, log2)
#define BLK_QUANTIZE(dst, src, scale, shift, offset, n) oracle_quantize(dst, src, scale, shift, offset, n)
#define BLK_QUANTIZE_INVERSE(dst, src, scale, shift, n) oracle_quantize_inverse(dst, src, scale, shift, n)
#define BLK_QUANTIZE_RECONSTRUCT(rec, sr, pred, sp, res, log2) oracle_quantize_reconstruct(rec, sr, pred, sp, res, 1 << (log2))
#include "batch_driver.inc"
