// hevcasm_b200 - quantise / inverse quantise / reconstruct kernels for sm_100a.
//
// Reference semantics (kupix/hevcasm): quantize.c:160-186 (hevcasm_quantize_c_ref), quantize.c:53-62
// (hevcasm_quantize_inverse_c_ref), quantize.c:292-302 (hevcasm_quantize_reconstruct_c_ref).
//
// These are pure streaming kernels (4 B of HBM traffic per coefficient): 128-bit loads/stores, two independent
// vectors in flight per thread, arithmetic in 32-bit int exactly as the C, saturating pack (cvt.pack.sat) for the
// final clip.  The per-block coded-block flag of quantize is an OR reduction across the lanes that share a block
// (warp shuffles; shared memory only when a block spans more than one warp).
#include "common.cuh"

namespace hv {

__device__ __forceinline__ int lo16(uint32_t w) { return (int)(short)(w & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t w) { return (int)w >> 16; }

// ---- quantize ------------------------------------------------------------------------------------

// sign(x) * ((|x| * scale + off) >> shift); the clip to int16 is done by the saturating pack of the caller
__device__ __forceinline__ int quant1(int x, int scale, int off, int shift)
{
    const int q = (abs(x) * scale + off) >> shift;
    return x < 0 ? -q : q;
}

// returns the quantised pair packed, ORs the two clipped outputs (sign-extended, like the C `cbf |= x`) into cbf
__device__ __forceinline__ uint32_t quant_word(uint32_t w, int scale, int off, int shift, int &cbf)
{
    const uint32_t r = pack_sat_s16(quant1(lo16(w), scale, off, shift), quant1(hi16(w), scale, off, shift));
    cbf |= lo16(r) | hi16(r);
    return r;
}

constexpr int Q_NT = 256;

// Each CTA covers 2 x Q_NT vectors (of 8 coefficients); thread t owns vectors t and t + Q_NT of the CTA's chunk, so
// both loads of a warp are fully coalesced and the lanes of one transform block stay adjacent.
__global__ void __launch_bounds__(Q_NT) quantize_kernel(int4 *__restrict__ dst, const int4 *__restrict__ src, int scale, int shift,
                                                        int off, int log2_vpb /* log2(vectors per block) */, long long n_vec,
                                                        int32_t *__restrict__ cbf)
{
    __shared__ int s_or[2][Q_NT / 32];
    const int tid = threadIdx.x;
    const long long base = (long long)blockIdx.x * (2 * Q_NT);
    int v_or[2] = {0, 0};
    int4 v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const long long i = base + h * Q_NT + tid;
        if (i < n_vec) v[h] = ldg_stream(src + i);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const long long i = base + h * Q_NT + tid;
        if (i < n_vec) {
            int4 r;
            r.x = (int)quant_word((uint32_t)v[h].x, scale, off, shift, v_or[h]);
            r.y = (int)quant_word((uint32_t)v[h].y, scale, off, shift, v_or[h]);
            r.z = (int)quant_word((uint32_t)v[h].z, scale, off, shift, v_or[h]);
            r.w = (int)quant_word((uint32_t)v[h].w, scale, off, shift, v_or[h]);
            stg_stream(dst + i, r);
        }
    }
    if (!cbf) return;
    const int vpb = 1 << log2_vpb;
    // OR across the lanes of a block (vpb is a power of two, so groups never straddle a warp unless vpb > 32)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
        if (o < vpb) {
            v_or[0] |= __shfl_xor_sync(0xffffffffu, v_or[0], o);
            v_or[1] |= __shfl_xor_sync(0xffffffffu, v_or[1], o);
        }
    if (vpb <= 32) {
        if ((tid & (vpb - 1)) == 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long long i = base + h * Q_NT + tid;
                if (i < n_vec) cbf[i >> log2_vpb] = v_or[h];
            }
        }
    } else {
        if ((tid & 31) == 0) s_or[0][tid >> 5] = v_or[0], s_or[1][tid >> 5] = v_or[1];
        __syncthreads();
        const int wpb = vpb >> 5;  // warps per block: 2 or 4
        if (tid < 2 * (Q_NT / vpb)) {
            const int h = tid / (Q_NT / vpb), b = tid % (Q_NT / vpb);
            int r = 0;
            for (int k = 0; k < wpb; ++k) r |= s_or[h][b * wpb + k];
            const long long i = base + h * Q_NT + (long long)b * vpb;
            if (i < n_vec) cbf[i >> log2_vpb] = r;
        }
    }
}

// ---- inverse quantize ----------------------------------------------------------------------------

__device__ __forceinline__ uint32_t dequant_word(uint32_t w, int scale, int add, int shift)
{
    return pack_sat_s16((lo16(w) * scale + add) >> shift, (hi16(w) * scale + add) >> shift);
}

__global__ void __launch_bounds__(Q_NT) quantize_inverse_kernel(int4 *__restrict__ dst, const int4 *__restrict__ src, int scale, int shift,
                                                                long long n_vec)
{
    const int add = 1 << (shift - 1);
    const long long base = (long long)blockIdx.x * (2 * Q_NT) + threadIdx.x;
    int4 v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
        if (base + h * Q_NT < n_vec) v[h] = ldg_stream(src + base + h * Q_NT);
#pragma unroll
    for (int h = 0; h < 2; ++h)
        if (base + h * Q_NT < n_vec) {
            int4 r;
            r.x = (int)dequant_word((uint32_t)v[h].x, scale, add, shift);
            r.y = (int)dequant_word((uint32_t)v[h].y, scale, add, shift);
            r.z = (int)dequant_word((uint32_t)v[h].z, scale, add, shift);
            r.w = (int)dequant_word((uint32_t)v[h].w, scale, add, shift);
            stg_stream(dst + base + h * Q_NT, r);
        }
}

// element-granular forms for unaligned pointers / odd tails (still on the GPU: there is no CPU path)
__global__ void quantize_inverse_scalar_kernel(int16_t *__restrict__ dst, const int16_t *__restrict__ src, int scale, int shift, long long first,
                                               long long n)
{
    const long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (int16_t)clip16((src[i] * scale + (1 << (shift - 1))) >> shift);
}

// ---- reconstruct ---------------------------------------------------------------------------------

// rec = clip8(pred + res) for 8 samples: res as one int4, pred / rec as two words
__device__ __forceinline__ uint2 recon8(uint2 pred, int4 res)
{
    const uint32_t p0 = pred.x, p1 = pred.y;
    uint2 r;
    r.x = pack_sat_u8((int)(p0 & 0xff) + lo16((uint32_t)res.x), (int)((p0 >> 8) & 0xff) + hi16((uint32_t)res.x),
                      (int)((p0 >> 16) & 0xff) + lo16((uint32_t)res.y), (int)(p0 >> 24) + hi16((uint32_t)res.y));
    r.y = pack_sat_u8((int)(p1 & 0xff) + lo16((uint32_t)res.z), (int)((p1 >> 8) & 0xff) + hi16((uint32_t)res.z),
                      (int)((p1 >> 16) & 0xff) + lo16((uint32_t)res.w), (int)(p1 >> 24) + hi16((uint32_t)res.w));
    return r;
}

// byte-granular accessors so any pointer alignment / stride works
__device__ __forceinline__ uint32_t ld4(const uint8_t *p)
{
    if (((uintptr_t)p & 3) == 0) return *reinterpret_cast<const uint32_t *>(p);
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ void st4(uint8_t *p, uint32_t v)
{
    if (((uintptr_t)p & 3) == 0) {
        *reinterpret_cast<uint32_t *>(p) = v;
    } else {
        p[0] = (uint8_t)v, p[1] = (uint8_t)(v >> 8), p[2] = (uint8_t)(v >> 16), p[3] = (uint8_t)(v >> 24);
    }
}

struct ReconParams {
    uint8_t *rec;
    const uint8_t *pred;
    const int16_t *res;
    ptrdiff_t sr, sp, fs_rec, fs_pred;
    int nbx, nby;                // frames form: block grid per frame (blockIdx.y = block row, blockIdx.z = frame)
    const int16_t *blk_xy;       // list form: nbx = number of blocks, nby = 1
    int desc_w;                  // list form: int16 per entry - (x, y), or (x, y, frame) for the *_list_frames form
};

// The residual of a block row is one flat run of int16 (blocks are N*N contiguous), so the kernel streams it: unit u is
// the u-th group of 8 residuals of the block row (for 4x4: 8 residuals = two block rows of 4), decoded with compile-time
// divisors into (block, row, 8-sample segment).  A warp's residual loads are 512 contiguous bytes, its predictor /
// reconstruction accesses whole 32-byte sectors; every thread keeps UNR independent units in flight.
// PA: plane pointers and strides are 16-byte multiples (regular grid only) -> natural-width vector accesses, no checks.
template <int LOG2, bool PA>
__global__ void __launch_bounds__(256) reconstruct_kernel(ReconParams p)
{
    constexpr int N = 1 << LOG2, UNR = 4;
    constexpr int UPB = N * N / 8;  // units per block
    const int f = blockIdx.z, by = blockIdx.y;
    // 16x16: a row of one block is only half a 32-byte sector of the predictor, so units are dealt to lanes over PAIRS of blocks - a warp takes
    // eight rows of two neighbouring blocks (whole sectors on the plane side, two 256-byte runs on the residual side); 92 -> 81 us per 16 4K frames
    constexpr bool PAIRS = N == 16;
    const int units = PAIRS ? ((p.nbx + 1) / 2) * (2 * UPB) : p.nbx * UPB;
    const int u0 = blockIdx.x * (256 * UNR) + threadIdx.x;
    int4 rv[UNR];
    uint2 pv[UNR];
    const uint8_t *pp[UNR];
    ptrdiff_t ro[UNR];
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
        const int u = u0 + k * 256;
        if (u >= units) continue;
        int bx = u / UPB, w = u % UPB;  // block, unit within the block
        if (PAIRS) {
            const int v = u % (2 * UPB), l = v & 31;
            bx = 2 * (u / (2 * UPB)) + ((l >> 1) & 1), w = (8 * (v >> 5) + (l >> 2)) * 2 + (l & 1);
            if (bx >= p.nbx) continue;
        }
        int x, y, r, xs;
        size_t blk;
        if (N == 4) r = 2 * w, xs = 0;
        else r = w / (N / 8), xs = (w % (N / 8)) * 8;
        int fk = f;
        if (!PA && p.blk_xy) {   // (PA is only ever chosen for regular grids: the list branch folds away there)
            const int16_t *e = p.blk_xy + (size_t)p.desc_w * bx;
            x = e[0], y = e[1], blk = (size_t)bx;
            if (p.desc_w == 3) fk = e[2];
        } else {
            x = bx * N, y = by * N, blk = ((size_t)f * p.nby + by) * p.nbx + bx;
        }
        rv[k] = ldg_stream(reinterpret_cast<const int4 *>(p.res + blk * (N * N)) + w);
        pp[k] = p.pred + fk * p.fs_pred + (ptrdiff_t)(y + r) * p.sp + x + xs;
        ro[k] = fk * p.fs_rec + (ptrdiff_t)(y + r) * p.sr + x + xs;
        if (N == 4) pv[k] = make_uint2(PA ? __ldg(reinterpret_cast<const uint32_t *>(pp[k])) : ld4(pp[k]),
                                       PA ? __ldg(reinterpret_cast<const uint32_t *>(pp[k] + p.sp)) : ld4(pp[k] + p.sp));
        else pv[k] = PA ? __ldg(reinterpret_cast<const uint2 *>(pp[k])) : make_uint2(ld4(pp[k]), ld4(pp[k] + 4));
    }
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
        const int u = u0 + k * 256;
        if (u >= units) continue;
        if (PAIRS && 2 * (u / (2 * UPB)) + (((u % (2 * UPB)) >> 1) & 1) >= p.nbx) continue;
        const uint2 o = recon8(pv[k], rv[k]);
        uint8_t *rp = p.rec + ro[k];
        if (N == 4) {
            if (PA) *reinterpret_cast<uint32_t *>(rp) = o.x, *reinterpret_cast<uint32_t *>(rp + p.sr) = o.y;
            else st4(rp, o.x), st4(rp + p.sr, o.y);
        } else {
            if (PA) *reinterpret_cast<uint2 *>(rp) = o;
            else st4(rp, o.x), st4(rp + 4, o.y);
        }
    }
}

}  // namespace hv

using namespace hv;

static int ilog2_exact(int v)
{
    int l = 0;
    while ((1 << l) < v) ++l;
    return (1 << l) == v ? l : -1;
}

extern "C" int hevcasm_quantize_batch(int16_t *dst, const int16_t *src, int scale, int shift, int offset, int n_per_block, int n_blocks,
                                      int32_t *cbf, void *stream)
{
    // reference quantize.c:162-168 asserts: scale, offset < 0x8000, 16 <= shift <= 27
    const int l = ilog2_exact(n_per_block);
    if (l < 4 || l > 10 || n_blocks < 0 || shift < 16 || shift > 27 || scale < 0 || scale >= 0x8000 || offset < 0 || offset >= 0x8000 ||
        (((uintptr_t)dst | (uintptr_t)src) & 15))
        return HEVCASM_ERR_ARGUMENT;
    if (n_blocks == 0) return 0;
    const long long n_vec = (long long)n_blocks * (n_per_block / 8);
    const unsigned grid = (unsigned)((n_vec + 2 * Q_NT - 1) / (2 * Q_NT));
    HV_LAUNCH(quantize_kernel, grid, Q_NT, 0, stream, reinterpret_cast<int4 *>(dst), reinterpret_cast<const int4 *>(src), scale, shift,
              offset << (shift - 16), l - 3, n_vec, cbf);
    return 0;
}

extern "C" int hevcasm_quantize_inverse_batch(int16_t *dst, const int16_t *src, int scale, int shift, long long n_total, void *stream)
{
    if (n_total < 0 || shift < 1 || shift > 30) return HEVCASM_ERR_ARGUMENT;
    if (n_total == 0) return 0;
    long long done = 0;
    if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0 && n_total >= 8) {
        const long long n_vec = n_total / 8;
        const unsigned grid = (unsigned)((n_vec + 2 * Q_NT - 1) / (2 * Q_NT));
        HV_LAUNCH(quantize_inverse_kernel, grid, Q_NT, 0, stream, reinterpret_cast<int4 *>(dst), reinterpret_cast<const int4 *>(src), scale,
                  shift, n_vec);
        done = n_vec * 8;
    }
    if (done < n_total) {
        const long long rem = n_total - done;
        HV_LAUNCH(quantize_inverse_scalar_kernel, (unsigned)((rem + 255) / 256), 256, 0, stream, dst, src, scale, shift, done, n_total);
    }
    return 0;
}

template <int LOG2>
static int launch_reconstruct_t(const ReconParams &p, int n_frames, bool pa, void *stream)
{
    constexpr int N = 1 << LOG2, UPB = N * N / 8;
    const long long units = N == 16 ? (long long)((p.nbx + 1) / 2) * (2 * UPB) : (long long)p.nbx * UPB;   // 16x16: dealt over pairs of blocks
    const dim3 grid((unsigned)((units + 1023) / 1024), p.nby, n_frames);
    if (pa) HV_LAUNCH((reconstruct_kernel<LOG2, true>), grid, 256, 0, stream, p);
    else HV_LAUNCH((reconstruct_kernel<LOG2, false>), grid, 256, 0, stream, p);
    return 0;
}

static int launch_reconstruct(const ReconParams &p, int log2, int n_frames, void *stream)
{
    const bool pa = !p.blk_xy && ((((uintptr_t)p.rec | (uintptr_t)p.pred | (uintptr_t)p.sr | (uintptr_t)p.sp | (uintptr_t)p.fs_rec | (uintptr_t)p.fs_pred) & 15) == 0);
    switch (log2) {
        case 2: return launch_reconstruct_t<2>(p, n_frames, pa, stream);
        case 3: return launch_reconstruct_t<3>(p, n_frames, pa, stream);
        case 4: return launch_reconstruct_t<4>(p, n_frames, pa, stream);
        default: return launch_reconstruct_t<5>(p, n_frames, pa, stream);
    }
}

extern "C" int hevcasm_quantize_reconstruct_batch(uint8_t *rec, ptrdiff_t sr, const uint8_t *pred, ptrdiff_t sp, const int16_t *res,
                                                  int log2size, const int16_t *blk_xy, int n, void *stream)
{
    if (log2size < 2 || log2size > 5 || n < 0 || !blk_xy || ((uintptr_t)res & 15)) return HEVCASM_ERR_ARGUMENT;
    if (n == 0) return 0;
    ReconParams p;
    p.rec = rec, p.pred = pred, p.res = res, p.sr = sr, p.sp = sp, p.fs_rec = 0, p.fs_pred = 0;
    p.nbx = n, p.nby = 1, p.blk_xy = blk_xy, p.desc_w = 2;
    return launch_reconstruct(p, log2size, 1, stream);
}

// Transform-unit lists of a batch of frames, bucketed by size: entries (x, y, frame), first the n_by_size[0] 4x4 blocks, then the 8x8, 16x16 and
// 32x32 ones; the residual blocks lie contiguous in the same order.  One launch per size present.
extern "C" int hevcasm_quantize_reconstruct_list_frames(uint8_t *rec, ptrdiff_t sr, const uint8_t *pred, ptrdiff_t sp, const int16_t *res, const int16_t *tus,
                                                        const int *n_by_size, ptrdiff_t fs_rec, ptrdiff_t fs_pred, void *stream)
{
    if (!n_by_size || ((uintptr_t)res & 15)) return HEVCASM_ERR_ARGUMENT;
    long long total = 0;
    for (int c = 0; c < 4; ++c) {
        if (n_by_size[c] < 0) return HEVCASM_ERR_ARGUMENT;
        total += n_by_size[c];
    }
    if (total && !tus) return HEVCASM_ERR_ARGUMENT;
    long long first = 0, off = 0;
    for (int c = 0; c < 4; ++c) {
        const int n = n_by_size[c], log2 = 2 + c;
        if (n) {
            ReconParams p;
            p.rec = rec, p.pred = pred, p.res = res + off, p.sr = sr, p.sp = sp, p.fs_rec = fs_rec, p.fs_pred = fs_pred;
            p.nbx = n, p.nby = 1, p.blk_xy = tus + 3 * first, p.desc_w = 3;
            const int e = launch_reconstruct(p, log2, 1, stream);
            if (e) return e;
        }
        first += n, off += (long long)n << (2 * log2);
    }
    return 0;
}

extern "C" int hevcasm_quantize_reconstruct_frames(uint8_t *rec, ptrdiff_t sr, const uint8_t *pred, ptrdiff_t sp, const int16_t *res, int width,
                                                   int height, int log2size, int n_frames, ptrdiff_t fs_rec, ptrdiff_t fs_pred, void *stream)
{
    if (log2size < 2 || log2size > 5 || n_frames < 0 || width < 0 || height < 0 || ((uintptr_t)res & 15)) return HEVCASM_ERR_ARGUMENT;
    ReconParams p;
    p.rec = rec, p.pred = pred, p.res = res, p.sr = sr, p.sp = sp, p.fs_rec = fs_rec, p.fs_pred = fs_pred;
    p.nbx = width >> log2size, p.nby = height >> log2size, p.blk_xy = nullptr, p.desc_w = 2;
    if (p.nbx == 0 || p.nby == 0 || n_frames == 0) return 0;
    return launch_reconstruct(p, log2size, n_frames, stream);
}
