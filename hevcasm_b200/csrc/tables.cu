// hevcasm_b200 - the reference's function-select tables (hevcasm_populate_* / hevcasm_get_*) for the GPU build.
//
// Mirrors the selection mechanism of the reference (sad.c:63-98, :128-180; ssd.c:63-86; pred_inter.c:231-367, :558-612;
// residual_decode.c:461-510, :906-938; quantize.c:65-79, :189-204, :324-330): populate(table, mask) writes a function
// pointer or 0 ("no implementation") into every slot.  This library implements exactly one instruction set,
// HEVCASM_CUDA (bit 9): with that bit in the mask every slot the reference's C path serves is populated, without it
// every slot is 0 - there is no CPU implementation to fall back to.
//
// A per-block slot is a batch-of-one call of the batched entry points of hevcasm_batch.h: the block (plus the filter
// halo the reference itself reads) is copied to a small device scratch area, the SAME kernel the batched path uses is
// launched for one element, and the result is copied back before the slot returns.  That keeps the reference's
// synchronous, host-pointer contract - and is therefore only meant for parity testing through the reference's own
// harness and for bring-up; throughput comes from the batched forms.  Slots are serialised by a mutex (they share the
// scratch area) and abort() with a message if CUDA fails: the reference's kernels cannot report errors either.
#include "common.cuh"

#include "sad.h"
#include "ssd.h"
#include "pred_inter.h"
#include "residual_decode.h"
#include "quantize.h"
#include "hadamard.h"
#include "diff.h"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace {

constexpr size_t kRegion = 64 * 1024;  // bytes per scratch region
constexpr int kRegions = 6;

struct Scratch {
    std::mutex mu;
    cudaStream_t s = nullptr;
    uint8_t *dev = nullptr;
    bool ready = false;
    int device = 0;   // the device the stream and the staging area belong to: whichever was current at the first slot call

    void init()
    {
        if (ready) return;
        check(cudaGetDevice(&device), "cudaGetDevice");
        check(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
        check(cudaMalloc((void **)&dev, kRegion * kRegions), "cudaMalloc");
        check(cudaMemsetAsync(dev, 0, kRegion * kRegions, s), "cudaMemset");
        ready = true;
    }
    static void check(cudaError_t e, const char *what)
    {
        if (e != cudaSuccess) {
            fprintf(stderr, "hevcasm_b200: per-block CUDA slot failed in %s: %s (no CPU fallback exists)\n", what, cudaGetErrorString(e));
            abort();
        }
    }
    static void check_hv(int e, const char *what)
    {
        if (e != 0) {
            fprintf(stderr, "hevcasm_b200: per-block CUDA slot failed in %s: %s (no CPU fallback exists)\n", what, hevcasm_cuda_error_string(e));
            abort();
        }
    }
    uint8_t *region(int i) { return dev + (size_t)i * kRegion; }
    void up2d(void *d, size_t dpitch, const void *h, size_t hpitch, size_t width_bytes, size_t rows)
    {
        check(cudaMemcpy2DAsync(d, dpitch, h, hpitch, width_bytes, rows, cudaMemcpyHostToDevice, s), "H2D");
    }
    void down2d(void *h, size_t hpitch, const void *d, size_t dpitch, size_t width_bytes, size_t rows)
    {
        check(cudaMemcpy2DAsync(h, hpitch, d, dpitch, width_bytes, rows, cudaMemcpyDeviceToHost, s), "D2H");
    }
    void up(void *d, const void *h, size_t bytes) { check(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s), "H2D"); }
    void down(void *h, const void *d, size_t bytes) { check(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s), "D2H"); }
    void sync() { check(cudaStreamSynchronize(s), "synchronize"); }
};

Scratch g_scratch;

// Holds the scratch lock and makes the scratch's device current for the duration of one slot call (a caller thread may have another
// device current; launching on a foreign stream would fail), restoring the caller's device afterwards.
struct Locked {
    std::lock_guard<std::mutex> lock;
    Scratch &sc;
    int caller_device = -1;
    Locked() : lock(g_scratch.mu), sc(g_scratch)
    {
        sc.init();
        int cur = 0;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != sc.device) {
            caller_device = cur;
            Scratch::check(cudaSetDevice(sc.device), "cudaSetDevice");
        }
    }
    ~Locked()
    {
        if (caller_device >= 0) cudaSetDevice(caller_device);
    }
};

// region roles
enum { R_A = 0, R_B = 1, R_C = 2, R_OUT = 3, R_DESC = 4, R_ZERO = 5 };  // R_ZERO stays all-zero: the {0, 0} block position
constexpr int kPitch = 128;                                            // row pitch of staged planes (bytes or int16 elements)

// ------------------------------------------------------------------------------------------------ SAD / SSD

int sad_cuda(const uint8_t *src, ptrdiff_t ss, const uint8_t *ref, ptrdiff_t sr, uint32_t rect)
{
    const int w = (int)(rect >> 8), h = (int)(rect & 0xff);
    Locked l;
    Scratch &sc = l.sc;
    sc.up2d(sc.region(R_A), kPitch, src, (size_t)ss, w, h);
    sc.up2d(sc.region(R_B), kPitch, ref, (size_t)sr, w, h);
    Scratch::check_hv(hevcasm_sad_batch(sc.region(R_A), kPitch, sc.region(R_B), kPitch, rect, (const int16_t *)sc.region(R_ZERO), nullptr, 1,
                                        (int32_t *)sc.region(R_OUT), sc.s),
                      "hevcasm_sad_batch");
    int32_t out = 0;
    sc.down(&out, sc.region(R_OUT), sizeof out);
    sc.sync();
    return out;
}

void sad_multiref_4_cuda(const uint8_t *src, ptrdiff_t ss, const uint8_t *ref[], ptrdiff_t sr, int sad[], uint32_t rect)
{
    const int w = (int)(rect >> 8), h = (int)(rect & 0xff);
    Locked l;
    Scratch &sc = l.sc;
    sc.up2d(sc.region(R_A), kPitch, src, (size_t)ss, w, h);
    int16_t cand[8];
    for (int i = 0; i < 4; ++i) {  // the four reference blocks are stacked 64 rows apart in one staged plane
        sc.up2d(sc.region(R_B) + (size_t)i * 64 * kPitch, kPitch, ref[i], (size_t)sr, w, h);
        cand[2 * i] = 0, cand[2 * i + 1] = (int16_t)(64 * i);
    }
    sc.up(sc.region(R_DESC), cand, sizeof cand);
    Scratch::check_hv(hevcasm_sad_multiref_batch(sc.region(R_A), kPitch, sc.region(R_B), kPitch, rect, (const int16_t *)sc.region(R_ZERO), 1,
                                                 (const int16_t *)sc.region(R_DESC), 4, (int32_t *)sc.region(R_OUT), sc.s),
                      "hevcasm_sad_multiref_batch");
    int32_t out[4];
    sc.down(out, sc.region(R_OUT), sizeof out);
    sc.sync();
    for (int i = 0; i < 4; ++i) sad[i] = out[i];
}

int ssd_cuda(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb, int w, int h)
{
    int log2 = 2;
    while ((1 << log2) < w) ++log2;
    Locked l;
    Scratch &sc = l.sc;
    sc.up2d(sc.region(R_A), kPitch, a, (size_t)sa, w, h);
    sc.up2d(sc.region(R_B), kPitch, b, (size_t)sb, w, h);
    Scratch::check_hv(hevcasm_ssd_batch(sc.region(R_A), kPitch, sc.region(R_B), kPitch, log2, (const int16_t *)sc.region(R_ZERO), 1,
                                        (int32_t *)sc.region(R_OUT), sc.s),
                      "hevcasm_ssd_batch");
    int32_t out = 0;
    sc.down(&out, sc.region(R_OUT), sizeof out);
    sc.sync();
    return out;
}

template <int LOG2>
int hadamard_satd_cuda(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb)
{
    constexpr int N = 1 << LOG2;
    Locked l;
    Scratch &sc = l.sc;
    sc.up2d(sc.region(R_A), kPitch, a, (size_t)sa, N, N);
    sc.up2d(sc.region(R_B), kPitch, b, (size_t)sb, N, N);
    Scratch::check_hv(hevcasm_hadamard_satd_batch(sc.region(R_A), kPitch, sc.region(R_B), kPitch, LOG2, (const int16_t *)sc.region(R_ZERO), 1,
                                                  (int32_t *)sc.region(R_OUT), sc.s),
                      "hevcasm_hadamard_satd_batch");
    int32_t out = 0;
    sc.down(&out, sc.region(R_OUT), sizeof out);
    sc.sync();
    return out;
}

int ssd_linear_cuda(const uint8_t *p0, const uint8_t *p1, int size)
{
    if (size <= 0) return 0;
    if (size > 33025) {   // the limit of hevcasm_ssd_linear_batch: 255^2 * size must fit the reference's int accumulator (and the staging region)
        fprintf(stderr, "hevcasm_b200: ssd_linear slot serves size <= 33025 (got %d)\n", size);
        abort();
    }
    Locked l;
    Scratch &sc = l.sc;
    sc.up(sc.region(R_A), p0, (size_t)size);
    sc.up(sc.region(R_B), p1, (size_t)size);
    Scratch::check_hv(hevcasm_ssd_linear_batch(sc.region(R_A), 0, sc.region(R_B), 0, size, 1, (int32_t *)sc.region(R_OUT), sc.s), "hevcasm_ssd_linear_batch");
    int32_t out = 0;
    sc.down(&out, sc.region(R_OUT), sizeof out);
    sc.sync();
    return out;
}

// ------------------------------------------------------------------------------------------------ inter prediction

// stages exactly the footprint the reference's C reads for this (xFrac, yFrac): the block plus taps/2-1 samples
// left/above and taps/2 right/below in the filtered directions (pred_inter.c:90-228); returns the device pointer
// that corresponds to `ref`
const uint8_t *stage_ref(Scratch &sc, int region, const uint8_t *ref, ptrdiff_t sr, int taps, int w, int h, bool fx, bool fy)
{
    const int left = fx ? taps / 2 - 1 : 0, right = fx ? taps / 2 : 0, top = fy ? taps / 2 - 1 : 0, bottom = fy ? taps / 2 : 0;
    uint8_t *org = sc.region(region) + 8 * kPitch + 16;  // sample (0,0) of the staged block; rows/cols around it stay zero
    sc.up2d(org - top * kPitch - left, kPitch, ref - (ptrdiff_t)top * sr - left, (size_t)sr, w + left + right, h + top + bottom);
    return org;
}

template <int TAPS>
void pred_uni_cuda(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref, ptrdiff_t sr, int w, int h, int xFrac, int yFrac)
{
    Locked l;
    Scratch &sc = l.sc;
    const uint8_t *dref = stage_ref(sc, R_A, ref, sr, TAPS, w, h, xFrac != 0, yFrac != 0);
    const int16_t pu[6] = {0, 0, (int16_t)w, (int16_t)h, (int16_t)xFrac, (int16_t)yFrac};
    sc.up(sc.region(R_DESC), pu, sizeof pu);
    Scratch::check_hv(hevcasm_pred_uni_batch(sc.region(R_OUT), kPitch, dref, kPitch, TAPS, (const int16_t *)sc.region(R_DESC), 1, sc.s), "hevcasm_pred_uni_batch");
    sc.down2d(dst, (size_t)sd, sc.region(R_OUT), kPitch, w, h);
    sc.sync();
}

template <int TAPS>
void pred_bi_cuda(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref0, const uint8_t *ref1, ptrdiff_t sr, int w, int h, int xFrac0, int yFrac0, int xFrac1,
                  int yFrac1)
{
    Locked l;
    Scratch &sc = l.sc;
    // the reference's bi path always runs both passes (pred_inter.c:504-527), so it always reads the full halo
    const uint8_t *d0 = stage_ref(sc, R_A, ref0, sr, TAPS, w, h, true, true);
    const uint8_t *d1 = stage_ref(sc, R_B, ref1, sr, TAPS, w, h, true, true);
    // both staged planes must be addressed from one base with one stride: ref1 = ref0 + (R_B - R_A) regions
    const int16_t pu[8] = {0, 0, (int16_t)w, (int16_t)h, (int16_t)xFrac0, (int16_t)yFrac0, (int16_t)xFrac1, (int16_t)yFrac1};
    sc.up(sc.region(R_DESC), pu, sizeof pu);
    Scratch::check_hv(hevcasm_pred_bi_batch(sc.region(R_OUT), kPitch, d0, d1, kPitch, TAPS, (const int16_t *)sc.region(R_DESC), 1, sc.s), "hevcasm_pred_bi_batch");
    sc.down2d(dst, (size_t)sd, sc.region(R_OUT), kPitch, w, h);
    sc.sync();
}

// ------------------------------------------------------------------------------------------------ transforms

template <int LOG2, int TRTYPE>
void transform_cuda(int16_t *coeffs, const int16_t *src, ptrdiff_t stride)
{
    constexpr int N = 1 << LOG2;
    Locked l;
    Scratch &sc = l.sc;
    sc.up2d(sc.region(R_A), kPitch * 2, src, (size_t)stride * 2, N * 2, N);
    Scratch::check_hv(hevcasm_transform_batch((int16_t *)sc.region(R_OUT), (const int16_t *)sc.region(R_A), kPitch, LOG2, TRTYPE,
                                              (const int16_t *)sc.region(R_ZERO), 1, sc.s),
                      "hevcasm_transform_batch");
    sc.down(coeffs, sc.region(R_OUT), N * N * 2);
    sc.sync();
}

template <int LOG2, int TRTYPE>
void inverse_transform_add_cuda(uint8_t *dst, ptrdiff_t sd, const uint8_t *pred, ptrdiff_t sp, const int16_t *coeffs)
{
    constexpr int N = 1 << LOG2;
    Locked l;
    Scratch &sc = l.sc;
    sc.up(sc.region(R_A), coeffs, N * N * 2);
    sc.up2d(sc.region(R_B), kPitch, pred, (size_t)sp, N, N);
    Scratch::check_hv(hevcasm_inverse_transform_add_batch(sc.region(R_OUT), kPitch, sc.region(R_B), kPitch, (const int16_t *)sc.region(R_A), LOG2, TRTYPE,
                                                          (const int16_t *)sc.region(R_ZERO), 1, sc.s),
                      "hevcasm_inverse_transform_add_batch");
    sc.down2d(dst, (size_t)sd, sc.region(R_OUT), kPitch, N, N);
    sc.sync();
}

// ------------------------------------------------------------------------------------------------ quantisation

int quantize_cuda(int16_t *dst, const int16_t *src, int scale, int shift, int offset, int n)
{
    if (n <= 0) return 0;
    if (n % 16 != 0 || (size_t)n * 2 > kRegion) {
        fprintf(stderr, "hevcasm_b200: quantize slot needs n %% 16 == 0 and n <= %zu (got %d), like the reference's SIMD path\n", kRegion / 2, n);
        abort();
    }
    Locked l;
    Scratch &sc = l.sc;
    sc.up(sc.region(R_A), src, (size_t)n * 2);
    // one coded-block flag per power-of-two run; a run length that is not a power of two is served as runs of 16
    int npb = n;
    if (npb & (npb - 1) || npb > 1024) npb = 16;
    const int runs = n / npb;
    Scratch::check_hv(hevcasm_quantize_batch((int16_t *)sc.region(R_OUT), (const int16_t *)sc.region(R_A), scale, shift, offset, npb, runs,
                                             (int32_t *)sc.region(R_C), sc.s),
                      "hevcasm_quantize_batch");
    sc.down(dst, sc.region(R_OUT), (size_t)n * 2);
    static int32_t flags[kRegion / 2 / 16];
    sc.down(flags, sc.region(R_C), (size_t)runs * 4);
    sc.sync();
    int cbf = 0;
    for (int i = 0; i < runs; ++i) cbf |= flags[i];  // the reference returns the OR over all n outputs
    return cbf;
}

void quantize_inverse_cuda(int16_t *dst, const int16_t *src, int scale, int shift, int n)
{
    if (n <= 0) return;
    if ((size_t)n * 2 > kRegion) {
        fprintf(stderr, "hevcasm_b200: quantize_inverse slot serves n <= %zu (got %d)\n", kRegion / 2, n);
        abort();
    }
    Locked l;
    Scratch &sc = l.sc;
    sc.up(sc.region(R_A), src, (size_t)n * 2);
    Scratch::check_hv(hevcasm_quantize_inverse_batch((int16_t *)sc.region(R_OUT), (const int16_t *)sc.region(R_A), scale, shift, n, sc.s),
                      "hevcasm_quantize_inverse_batch");
    sc.down(dst, sc.region(R_OUT), (size_t)n * 2);
    sc.sync();
}

void quantize_reconstruct_cuda(uint8_t *rec, ptrdiff_t sr, const uint8_t *pred, ptrdiff_t sp, const int16_t *res, int n)
{
    int log2 = 2;
    while ((1 << log2) < n) ++log2;
    Locked l;
    Scratch &sc = l.sc;
    sc.up(sc.region(R_A), res, (size_t)n * n * 2);
    sc.up2d(sc.region(R_B), kPitch, pred, (size_t)sp, n, n);
    Scratch::check_hv(hevcasm_quantize_reconstruct_batch(sc.region(R_OUT), kPitch, sc.region(R_B), kPitch, (const int16_t *)sc.region(R_A), log2,
                                                         (const int16_t *)sc.region(R_ZERO), 1, sc.s),
                      "hevcasm_quantize_reconstruct_batch");
    sc.down2d(rec, (size_t)sr, sc.region(R_OUT), kPitch, n, n);
    sc.sync();
}

inline bool cuda_bit(hevcasm_instruction_set mask) { return ((int)mask & (int)HEVCASM_CUDA) != 0; }

}  // namespace

// ================================================================================================ populate

extern "C" void hevcasm_populate_sad(hevcasm_table_sad *table, hevcasm_instruction_set mask)
{
    for (int height = 4; height <= 64; height += 4)  // reference sad.c:89-98: every multiple of 4 up to 64
        for (int width = 4; width <= 64; width += 4) *hevcasm_get_sad(table, width, height) = cuda_bit(mask) ? &sad_cuda : 0;
}

extern "C" void hevcasm_populate_sad_multiref(hevcasm_table_sad_multiref *table, hevcasm_instruction_set mask)
{
    table->sadGeneric_4 = 0;  // never written by the reference either (sad.h:97-102)
    for (int height = 4; height <= 64; height += 4)
        for (int width = 4; width <= 64; width += 4) *hevcasm_get_sad_multiref(table, 4, width, height) = cuda_bit(mask) ? &sad_multiref_4_cuda : 0;
}

extern "C" void hevcasm_populate_ssd(hevcasm_table_ssd *table, hevcasm_instruction_set mask)
{
    for (int log2 = 2; log2 <= 6; ++log2) *hevcasm_get_ssd(table, log2) = cuda_bit(mask) ? &ssd_cuda : 0;
}

extern "C" void hevcasm_populate_hadamard_satd(hevcasm_table_hadamard_satd *table, hevcasm_instruction_set mask)
{
    const bool on = cuda_bit(mask);  // reference hadamard.c:137-160
    *hevcasm_get_hadamard_satd(table, 1) = on ? &hadamard_satd_cuda<1> : 0;
    *hevcasm_get_hadamard_satd(table, 2) = on ? &hadamard_satd_cuda<2> : 0;
    *hevcasm_get_hadamard_satd(table, 3) = on ? &hadamard_satd_cuda<3> : 0;
}

extern "C" hevcasm_ssd_linear *hevcasm_get_ssd_linear(int size, hevcasm_instruction_set mask)
{
    (void)size;  // reference diff.c:54-63: one implementation for every size
    return cuda_bit(mask) ? &ssd_linear_cuda : 0;
}

extern "C" void hevcasm_populate_pred_uni_8to8(hevcasm_table_pred_uni_8to8 *table, hevcasm_instruction_set mask)
{
    memset(table, 0, sizeof *table);
    if (!cuda_bit(mask)) return;
    // reference pred_inter.c:351-367: luma widths up to 64, chroma up to 32... the slot index is ceil(w / taps), 1..8
    for (int taps = 4; taps <= 8; taps += 4)
        for (int w = 1; w <= 8 * taps; ++w)
            for (int xf = 0; xf < 2; ++xf)
                for (int yf = 0; yf < 2; ++yf)
                    *hevcasm_get_pred_uni_8to8(table, taps, w, 0, xf, yf) = taps == 8 ? &pred_uni_cuda<8> : &pred_uni_cuda<4>;
}

extern "C" void hevcasm_populate_pred_bi_8to8(hevcasm_table_pred_bi_8to8 *table, hevcasm_instruction_set mask)
{
    memset(table, 0, sizeof *table);
    if (!cuda_bit(mask)) return;
    for (int taps = 4; taps <= 8; taps += 4)
        for (int w = 1; w <= 8 * taps; ++w)
            for (int f = 0; f < 2; ++f) *hevcasm_get_pred_bi_8to8(table, taps, w, 0, f, 0, 0, 0) = taps == 8 ? &pred_bi_cuda<8> : &pred_bi_cuda<4>;
}

extern "C" void hevcasm_populate_transform(hevcasm_table_transform *table, hevcasm_instruction_set mask)
{
    const bool on = cuda_bit(mask);
    *hevcasm_get_transform(table, 1, 2) = on ? &transform_cuda<2, 1> : 0;
    *hevcasm_get_transform(table, 0, 2) = on ? &transform_cuda<2, 0> : 0;
    *hevcasm_get_transform(table, 0, 3) = on ? &transform_cuda<3, 0> : 0;
    *hevcasm_get_transform(table, 0, 4) = on ? &transform_cuda<4, 0> : 0;
    *hevcasm_get_transform(table, 0, 5) = on ? &transform_cuda<5, 0> : 0;
}

extern "C" void hevcasm_populate_inverse_transform_add(hevcasm_table_inverse_transform_add *table, hevcasm_instruction_set mask, int encoder)
{
    (void)encoder;  // the reference trades conformance for speed at 32x32 when set (residual_decode.c:489-495); the GPU path is exact either way
    const bool on = cuda_bit(mask);
    *hevcasm_get_inverse_transform_add(table, 1, 2) = on ? &inverse_transform_add_cuda<2, 1> : 0;
    *hevcasm_get_inverse_transform_add(table, 0, 2) = on ? &inverse_transform_add_cuda<2, 0> : 0;
    *hevcasm_get_inverse_transform_add(table, 0, 3) = on ? &inverse_transform_add_cuda<3, 0> : 0;
    *hevcasm_get_inverse_transform_add(table, 0, 4) = on ? &inverse_transform_add_cuda<4, 0> : 0;
    *hevcasm_get_inverse_transform_add(table, 0, 5) = on ? &inverse_transform_add_cuda<5, 0> : 0;
}

extern "C" void hevcasm_populate_quantize(hevcasm_table_quantize *table, hevcasm_instruction_set mask)
{
    *hevcasm_get_quantize(table) = cuda_bit(mask) ? &quantize_cuda : 0;
}

extern "C" void hevcasm_populate_quantize_inverse(hevcasm_table_quantize_inverse *table, hevcasm_instruction_set mask)
{
    *hevcasm_get_quantize_inverse(table) = cuda_bit(mask) ? &quantize_inverse_cuda : 0;
}

extern "C" void hevcasm_populate_quantize_reconstruct(hevcasm_table_quantize_reconstruct *table, hevcasm_instruction_set mask)
{
    for (int log2 = 2; log2 <= 5; ++log2) *hevcasm_get_quantize_reconstruct(table, log2) = cuda_bit(mask) ? &quantize_reconstruct_cuda : 0;
}
