// hevcasm_b200 - Hadamard SATD (2x2, 4x4, 8x8) and linear SSD for sm_100a.
//
// Reference semantics (kupix/hevcasm): hadamard.c:75-131 (compute_satd: D = A - B, T = H D H^T with the +-1 Hadamard matrix,
// result (N/4 + sum |T|) / (N/2) in integer arithmetic) and diff.c:45-54 (hevcasm_ssd_linear_c_ref).  These are the cost
// functions an encoder evaluates right after the integer-pel SAD search.
//
// SATD: one thread per block, the whole transform in registers (the order of the Hadamard outputs does not matter for the
// sum of absolute values, so the butterflies run in place); consecutive lanes take consecutive blocks of a block row, so
// the plane accesses of a warp are contiguous.  Linear SSD: one warp per run, |a-b| per byte (VABSDIFF4) squared and
// summed with IDP.4A.
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace hv {
#ifdef HEVCASM_EXPERIMENTS
#include "satd_umma.cuh"   // 4x4 / 8x8: horizontal Hadamard pass on tcgen05 (measured, not adopted: profiles/r01_satd.md)
#endif

// unsigned 32-bit division by a run-time constant as multiply-high + shifts (exact for every 32-bit numerator)
struct SatdDiv {
    uint32_t d, m, s;
    static SatdDiv make(uint32_t d)
    {
        SatdDiv f{d ? d : 1u, 0, 0};
        while ((1ull << f.s) < f.d) ++f.s;
        f.m = (uint32_t)((((1ull << f.s) - f.d) << 32) / f.d + 1);
        return f;
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const
    {
        if (s == 0) return n;
        const uint32_t t = __umulhi(n, m);
        return (((n - t) >> 1) + t) >> (s - 1);
    }
};

struct SatdGrid {
    const int16_t *blk_xy;
    int nbx, nby;
    long long n;
    SatdDiv per_frame, per_row;
    bool fast;   // regular grid, n < 2^32: locate by multiply-high
    int desc_w = 2;   // int16 per list entry: (x, y), or (x, y, frame) for the *_list_frames form
    void finish()
    {
        fast = !blk_xy && n > 0 && n < (1ll << 32) && nbx > 0 && nby > 0;
        if (fast) per_frame = SatdDiv::make((uint32_t)(nbx * nby)), per_row = SatdDiv::make((uint32_t)nbx);
    }
    // REGULAR: the calling kernel is only ever launched on regular grids, so the list branch folds away
    template <bool REGULAR = false>
    __device__ __forceinline__ void locate(long long i, int log2, int &x, int &y, int &f) const
    {
        f = 0;
        if (!REGULAR && blk_xy) {
            const int16_t *e = blk_xy + i * desc_w;
            x = e[0], y = e[1];
            if (desc_w == 3) f = e[2];
        } else if (fast) {
            const uint32_t u = (uint32_t)i, fr = per_frame.div(u), r = u - fr * per_frame.d, row = per_row.div(r);
            f = (int)fr, y = (int)(row << log2), x = (int)((r - row * per_row.d) << log2);
        } else {
            const long long per = (long long)nbx * nby;
            f = (int)(i / per);
            const int r = (int)(i - f * per);
            y = (r / nbx) << log2, x = (r % nbx) << log2;
        }
    }
};

template <int N>
__device__ __forceinline__ void load_row_bytes(const uint8_t *p, int (&v)[N], bool aligned)
{
    if (N >= 4 && aligned) {
#pragma unroll
        for (int k = 0; k < N / 4; ++k) {
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(p) + k);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[4 * k + j] = (int)((w >> (8 * j)) & 0xff);
        }
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = (int)__ldg(p + j);
    }
}

// in-place Hadamard butterflies over N values with the given register stride pattern
template <int N>
__device__ __forceinline__ void hadamard(int (&v)[N])
{
#pragma unroll
    for (int h = 1; h < N; h <<= 1)
#pragma unroll
        for (int i = 0; i < N; i += 2 * h)
#pragma unroll
            for (int j = i; j < i + h; ++j) {
                const int a = v[j], b = v[j + h];
                v[j] = a + b, v[j + h] = a - b;
            }
}

// 4-point Hadamard rows as IDP.4A byte patterns (+1 = 0x01, -1 = 0xff), and their negatives for the subtrahend
__device__ __forceinline__ int had4_row(uint32_t a, uint32_t b, int u)
{
    constexpr uint32_t P[4] = {0x01010101u, 0xff01ff01u, 0xffff0101u, 0x01ffff01u}, M[4] = {0xffffffffu, 0x01ff01ffu, 0x0101ffffu, 0xff0101ffu};
    return dp4a_us(b, (int)M[u], dp4a_us(a, (int)P[u], 0));
}

// One thread per block.  The horizontal pass never unpacks a byte: output u of a 4-sample group is IDP.4A(a, H_u) +
// IDP.4A(b, -H_u), on the FMA pipe; for 8x8 the two groups of a row are chained with equal / opposite signs, so the whole
// horizontal pass is 32 IDP.4A per row and the ALU pipe is left to the vertical butterflies (192 add/sub) and the
// |.| accumulation (one VABSDIFF each): 256 instructions per pipe per block, 8 per sample, where the byte-unpacking
// version needed ~17 on the ALU pipe alone (1.15 -> see profiles/r01_satd.md).  Needs 4-byte aligned rows; others take
// satd_generic_kernel.
template <int LOG2>
__global__ void __launch_bounds__(128) satd_kernel(const uint8_t *__restrict__ a, ptrdiff_t sa, const uint8_t *__restrict__ b, ptrdiff_t sb, ptrdiff_t fs_a,
                                                   ptrdiff_t fs_b, SatdGrid g, int32_t *__restrict__ out)
{
    constexpr int N = 1 << LOG2;
    static_assert(N == 4 || N == 8, "byte-wise horizontal pass: 4x4 and 8x8");
    const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
    if (i >= g.n) return;
    int x, y, f;
    g.template locate<true>(i, LOG2, x, y, f);   // launch_satd picks this kernel for regular grids only
    const uint8_t *pa = a + f * fs_a + (ptrdiff_t)y * sa + x, *pb = b + f * fs_b + (ptrdiff_t)y * sb + x;
    int d[N][N];
#pragma unroll
    for (int r = 0; r < N; ++r) {
        const uint32_t *ra = reinterpret_cast<const uint32_t *>(pa + (ptrdiff_t)r * sa), *rb = reinterpret_cast<const uint32_t *>(pb + (ptrdiff_t)r * sb);
        if (N == 4) {
            const uint32_t wa = __ldg(ra), wb = __ldg(rb);
#pragma unroll
            for (int u = 0; u < 4; ++u) d[r][u] = had4_row(wa, wb, u);
        } else {
            const uint32_t a0 = __ldg(ra), a1 = __ldg(ra + 1), b0 = __ldg(rb), b1 = __ldg(rb + 1);
            constexpr uint32_t P[4] = {0x01010101u, 0xff01ff01u, 0xffff0101u, 0x01ffff01u}, M[4] = {0xffffffffu, 0x01ff01ffu, 0x0101ffffu, 0xff0101ffu};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int lo = dp4a_us(b0, (int)M[u], dp4a_us(a0, (int)P[u], 0));
                d[r][u] = dp4a_us(b1, (int)M[u], dp4a_us(a1, (int)P[u], lo));       // group 0 + group 1
                d[r][u + 4] = dp4a_us(b1, (int)P[u], dp4a_us(a1, (int)M[u], lo));   // group 0 - group 1
            }
        }
    }
    int sum = N / 4;
#pragma unroll
    for (int c = 0; c < N; ++c) {
        int col[N];
#pragma unroll
        for (int r = 0; r < N; ++r) col[r] = d[r][c];
        hadamard<N>(col);   // along y
#pragma unroll
        for (int r = 0; r < N; ++r) sum = (int)__sad(col[r], 0, (unsigned)sum);
    }
    out[i] = sum / (N / 2);
}

template <int LOG2>
__global__ void __launch_bounds__(128) satd_generic_kernel(const uint8_t *__restrict__ a, ptrdiff_t sa, const uint8_t *__restrict__ b, ptrdiff_t sb,
                                                           ptrdiff_t fs_a, ptrdiff_t fs_b, SatdGrid g, int32_t *__restrict__ out)
{
    constexpr int N = 1 << LOG2;
    const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
    if (i >= g.n) return;
    int x, y, f;
    g.locate(i, LOG2, x, y, f);
    const uint8_t *pa = a + f * fs_a + (ptrdiff_t)y * sa + x, *pb = b + f * fs_b + (ptrdiff_t)y * sb + x;
    const bool al = ((((uintptr_t)pa | (uintptr_t)pb | (uintptr_t)sa | (uintptr_t)sb) & 3) == 0);
    int d[N][N];
#pragma unroll
    for (int r = 0; r < N; ++r) {
        int va[N], vb[N];
        load_row_bytes<N>(pa + (ptrdiff_t)r * sa, va, al);
        load_row_bytes<N>(pb + (ptrdiff_t)r * sb, vb, al);
#pragma unroll
        for (int c = 0; c < N; ++c) d[r][c] = va[c] - vb[c];
        hadamard<N>(d[r]);  // along x
    }
    int sum = N / 4;
#pragma unroll
    for (int c = 0; c < N; ++c) {
        int col[N];
#pragma unroll
        for (int r = 0; r < N; ++r) col[r] = d[r][c];
        hadamard<N>(col);   // along y
#pragma unroll
        for (int r = 0; r < N; ++r) sum += abs(col[r]);
    }
    out[i] = sum / (N / 2);
}

// 2x2 blocks of a regular grid, four adjacent blocks (8 samples x 2 rows) per thread: the rows come as 8-byte words, a block's four bytes
// are gathered into one word per plane (PRMT) and each of its four Hadamard outputs is two IDP.4A - the +-1 pattern on plane A, the negated
// pattern on plane B (H(A - B) = H A - H B) - so the 9-bit differences are never formed.  The results leave as one 16-byte store.
// (The generic kernel, one thread per 2x2 block with 2-byte loads: 142 us per 16 4K plane pairs, 0.43 of the HBM roofline.)
__global__ void __launch_bounds__(128) satd2_rows_kernel(const uint8_t *__restrict__ a, ptrdiff_t sa, const uint8_t *__restrict__ b, ptrdiff_t sb, ptrdiff_t fs_a,
                                                         ptrdiff_t fs_b, int groups_per_row, int nby, int32_t *__restrict__ out)
{
    const int gx = blockIdx.x * 128 + threadIdx.x, by = blockIdx.y, f = blockIdx.z;
    if (gx >= groups_per_row) return;
    const uint8_t *pa = a + f * fs_a + (ptrdiff_t)(2 * by) * sa + 8 * gx, *pb = b + f * fs_b + (ptrdiff_t)(2 * by) * sb + 8 * gx;
    const uint2 a0 = __ldg(reinterpret_cast<const uint2 *>(pa)), a1 = __ldg(reinterpret_cast<const uint2 *>(pa + sa));
    const uint2 b0 = __ldg(reinterpret_cast<const uint2 *>(pb)), b1 = __ldg(reinterpret_cast<const uint2 *>(pb + sb));
    // block k: bytes (r0c0, r0c1, r1c0, r1c1)
    const uint32_t A[4] = {__byte_perm(a0.x, a1.x, 0x5410), __byte_perm(a0.x, a1.x, 0x7632), __byte_perm(a0.y, a1.y, 0x5410), __byte_perm(a0.y, a1.y, 0x7632)};
    const uint32_t B[4] = {__byte_perm(b0.x, b1.x, 0x5410), __byte_perm(b0.x, b1.x, 0x7632), __byte_perm(b0.y, b1.y, 0x5410), __byte_perm(b0.y, b1.y, 0x7632)};
    constexpr int P0 = 0x01010101, P1 = (int)0xff01ff01, P2 = (int)0xffff0101, P3 = (int)0x01ffff01;   // (+,+,+,+) (+,-,+,-) (+,+,-,-) (+,-,-,+)
    constexpr int M0 = (int)0xffffffff, M1 = (int)0x01ff01ff, M2 = (int)0x0101ffff, M3 = (int)0xff0101ff;   // the same, negated
    int r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int t0 = dp4a_us(B[k], M0, dp4a_us(A[k], P0, 0)), t1 = dp4a_us(B[k], M1, dp4a_us(A[k], P1, 0));
        const int t2 = dp4a_us(B[k], M2, dp4a_us(A[k], P2, 0)), t3 = dp4a_us(B[k], M3, dp4a_us(A[k], P3, 0));
        r[k] = abs(t0) + abs(t1) + abs(t2) + abs(t3);
    }
    *reinterpret_cast<int4 *>(out + ((size_t)f * nby + by) * (4 * (size_t)groups_per_row) + 4 * (size_t)gx) = make_int4(r[0], r[1], r[2], r[3]);
}

// one warp per run; any alignment, any size
__global__ void __launch_bounds__(256) ssd_linear_kernel(const uint8_t *__restrict__ p0, ptrdiff_t rs0, const uint8_t *__restrict__ p1, ptrdiff_t rs1, int size,
                                                         int n_runs, int32_t *__restrict__ out)
{
    const int run = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (run >= n_runs) return;
    const uint8_t *a = p0 + (ptrdiff_t)run * rs0, *b = p1 + (ptrdiff_t)run * rs1;
    uint32_t acc = 0;
    const int words = size >> 2;
    for (int w = lane; w < words; w += 32) {
        const uint8_t *qa = a + 4 * w, *qb = b + 4 * w;
        uint32_t wa, wb;
        if ((((uintptr_t)qa | (uintptr_t)qb) & 3) == 0) {
            wa = __ldg(reinterpret_cast<const uint32_t *>(qa)), wb = __ldg(reinterpret_cast<const uint32_t *>(qb));
        } else {  // byte-granular: never touches a byte outside the run
            wa = (uint32_t)qa[0] | ((uint32_t)qa[1] << 8) | ((uint32_t)qa[2] << 16) | ((uint32_t)qa[3] << 24);
            wb = (uint32_t)qb[0] | ((uint32_t)qb[1] << 8) | ((uint32_t)qb[2] << 16) | ((uint32_t)qb[3] << 24);
        }
        const uint32_t dd = __vabsdiffu4(wa, wb);
        acc = dp4a_uu(dd, dd, acc);
    }
    for (int k = 4 * words + lane; k < size; k += 32) {
        const int dd = (int)a[k] - (int)b[k];
        acc += (uint32_t)(dd * dd);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[run] = (int32_t)acc;
}

}  // namespace hv

using namespace hv;

// 4x4 / 8x8 on the tensor cores (satd_umma.cuh): regular grids over 16-byte aligned planes
#ifdef HEVCASM_EXPERIMENTS
template <int LOG2>
static int launch_satd_umma(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb, ptrdiff_t fs_a, ptrdiff_t fs_b, const SatdGrid &g, int32_t *out,
                            void *stream, bool *taken)
{
    *taken = false;
    constexpr int N = 1 << LOG2;
    const long long per = (long long)g.nbx * g.nby;
    if (g.blk_xy || per <= 0 || g.n % per != 0) return 0;
    const int n_frames = (int)(g.n / per);
    if (!tma::describable(sa, fs_a, n_frames) || !tma::describable(sb, fs_b, n_frames)) return 0;
    su::Params P{};
    if (tma::describe_u32_swizzled128(&P.tm[0], a, sa, fs_a, (long long)N * g.nbx, (long long)N * g.nby, n_frames, su::TROWS) ||
        tma::describe_u32_swizzled128(&P.tm[1], b, sb, fs_b, (long long)N * g.nbx, (long long)N * g.nby, n_frames, su::TROWS))
        return 0;
    P.out = out, P.nbx = g.nbx, P.nby = g.nby;
    P.tiles_x = (g.nbx * N + su::TCOLS - 1) / su::TCOLS, P.tiles_y = (g.nby * N + su::TROWS - 1) / su::TROWS;
    const long long tiles = (long long)P.tiles_x * P.tiles_y * n_frames;
    if (tiles >= (1ll << 30)) return 0;
    P.n_tiles = (int)tiles;
    if (set_max_smem(su::satd_umma_kernel<LOG2>, su::SMEM_BYTES)) return (int)cudaErrorInvalidValue;
    *taken = true;
    long long grid = std::min<long long>(tiles, (long long)sm_count());   // one persistent CTA per SM
    if (const char *e = tune::knob("HEVCASM_SATD_UMMA_GRID")) grid = std::max(1ll, std::min<long long>(grid, atoll(e)));   // test knob: more tiles per CTA
    return launch(su::satd_umma_kernel<LOG2>, dim3((unsigned)grid), dim3(su::THREADS), (size_t)su::SMEM_BYTES, stream, P);
}
#endif

static int launch_satd(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb, ptrdiff_t fs_a, ptrdiff_t fs_b, int log2size, const SatdGrid &g,
                       int32_t *out, void *stream)
{
    if (g.n == 0) return 0;
#ifdef HEVCASM_EXPERIMENTS
    // HEVCASM_SATD_PATH=umma: 4x4 / 8x8 on the tensor-core kernel when the planes allow it; =umma_only: fail instead of falling back (tests)
    const char *pin = tune::knob("HEVCASM_SATD_PATH");
    if (log2size >= 2 && pin && !strncmp(pin, "umma", 4)) {
        bool taken = false;
        const int e = log2size == 3 ? launch_satd_umma<3>(a, sa, b, sb, fs_a, fs_b, g, out, stream, &taken) : launch_satd_umma<2>(a, sa, b, sb, fs_a, fs_b, g, out, stream, &taken);
        if (taken) return e;
        if (!strcmp(pin, "umma_only")) return HEVCASM_ERR_ARGUMENT;
    }
#endif
    const unsigned grid = (unsigned)((g.n + 127) / 128);
    // byte-wise kernels: every row of every block starts on a 4-byte boundary (regular grids of 4x4 / 8x8 blocks on 4-byte aligned planes)
    uintptr_t m = (uintptr_t)a | (uintptr_t)b | (uintptr_t)sa | (uintptr_t)sb;
    if (g.n > g.nbx * (long long)g.nby) m |= (uintptr_t)fs_a | (uintptr_t)fs_b;
    const bool bytewise = !g.blk_xy && (m & 3) == 0 && !tune::knob("HEVCASM_SATD_GENERIC");
    if (log2size == 1) {
        // regular grid, rows of whole groups of four blocks, everything 8- / 16-byte aligned: the row kernel
        const long long per = (long long)g.nbx * g.nby;
        const uintptr_t m8 = (uintptr_t)a | (uintptr_t)b | (uintptr_t)sa | (uintptr_t)sb | (g.n > per ? (uintptr_t)fs_a | (uintptr_t)fs_b : 0);
        if (!g.blk_xy && per > 0 && g.n % per == 0 && (g.nbx & 3) == 0 && (m8 & 7) == 0 && ((uintptr_t)out & 15) == 0 && g.nby <= 65535 && g.n / per <= 65535 &&
            !tune::knob("HEVCASM_SATD_GENERIC"))
            return launch(satd2_rows_kernel, dim3((unsigned)((g.nbx / 4 + 127) / 128), (unsigned)g.nby, (unsigned)(g.n / per)), 128, 0, stream, a, sa, b, sb, fs_a, fs_b,
                          g.nbx / 4, g.nby, out);
        return launch(satd_generic_kernel<1>, grid, 128, 0, stream, a, sa, b, sb, fs_a, fs_b, g, out);
    }
    if (log2size == 2)
        return bytewise ? launch(satd_kernel<2>, grid, 128, 0, stream, a, sa, b, sb, fs_a, fs_b, g, out)
                        : launch(satd_generic_kernel<2>, grid, 128, 0, stream, a, sa, b, sb, fs_a, fs_b, g, out);
    return bytewise ? launch(satd_kernel<3>, grid, 128, 0, stream, a, sa, b, sb, fs_a, fs_b, g, out)
                    : launch(satd_generic_kernel<3>, grid, 128, 0, stream, a, sa, b, sb, fs_a, fs_b, g, out);
}

extern "C" int hevcasm_hadamard_satd_batch(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb, int log2size, const int16_t *blk_xy, int n,
                                           int32_t *satd, void *stream)
{
    if (log2size < 1 || log2size > 3 || n < 0 || (n > 0 && !blk_xy)) return HEVCASM_ERR_ARGUMENT;
    SatdGrid g{blk_xy, 0, 0, n};
    g.finish();
    return launch_satd(a, sa, b, sb, 0, 0, log2size, g, satd, stream);
}

// Block lists of a batch of frames, bucketed by size: entries (x, y, frame), first the n_by_size[0] 2x2 blocks, then the 4x4 and the 8x8 ones;
// satd[i] in list order.  One launch per size present.
extern "C" int hevcasm_hadamard_satd_list_frames(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb, const int16_t *blks, const int *n_by_size,
                                                 ptrdiff_t fs_a, ptrdiff_t fs_b, int32_t *satd, void *stream)
{
    if (!n_by_size) return HEVCASM_ERR_ARGUMENT;
    long long total = 0;
    for (int c = 0; c < 3; ++c) {
        if (n_by_size[c] < 0) return HEVCASM_ERR_ARGUMENT;
        total += n_by_size[c];
    }
    if (total && (!blks || !satd)) return HEVCASM_ERR_ARGUMENT;
    long long first = 0;
    for (int c = 0; c < 3; ++c) {
        const int n = n_by_size[c];
        if (n) {
            SatdGrid g{blks + 3 * first, 0, 0, n};
            g.desc_w = 3;
            g.finish();
            const int e = launch_satd(a, sa, b, sb, fs_a, fs_b, 1 + c, g, satd + first, stream);
            if (e) return e;
        }
        first += n;
    }
    return 0;
}

extern "C" int hevcasm_hadamard_satd_frames(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb, int width, int height, int log2size, int n_frames,
                                            ptrdiff_t fs_a, ptrdiff_t fs_b, int32_t *satd, void *stream)
{
    if (log2size < 1 || log2size > 3 || n_frames < 0 || width < 0 || height < 0) return HEVCASM_ERR_ARGUMENT;
    SatdGrid g{nullptr, width >> log2size, height >> log2size, 0};
    g.n = (long long)g.nbx * g.nby * n_frames;
    g.finish();
    return launch_satd(a, sa, b, sb, fs_a, fs_b, log2size, g, satd, stream);
}

extern "C" int hevcasm_ssd_linear_batch(const uint8_t *p0, ptrdiff_t rs0, const uint8_t *p1, ptrdiff_t rs1, int size, int n_runs, int32_t *ssd, void *stream)
{
    // 255^2 * size must fit the reference's int accumulator
    if (size < 0 || size > 33025 || n_runs < 0) return HEVCASM_ERR_ARGUMENT;
    if (n_runs == 0) return 0;
    HV_LAUNCH(ssd_linear_kernel, (unsigned)(((long long)n_runs * 32 + 255) / 256), 256, 0, stream, p0, rs0, p1, rs1, size, n_runs, ssd);
    return 0;
}
