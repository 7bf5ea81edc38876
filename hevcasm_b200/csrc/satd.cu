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

namespace hv {

struct SatdGrid {
    const int16_t *blk_xy;
    int nbx, nby;
    long long n;
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

template <int LOG2>
__global__ void __launch_bounds__(128) satd_kernel(const uint8_t *__restrict__ a, ptrdiff_t sa, const uint8_t *__restrict__ b, ptrdiff_t sb, ptrdiff_t fs_a,
                                                   ptrdiff_t fs_b, SatdGrid g, int32_t *__restrict__ out)
{
    constexpr int N = 1 << LOG2;
    const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
    if (i >= g.n) return;
    int x, y, f = 0;
    if (g.blk_xy) {
        x = g.blk_xy[2 * i], y = g.blk_xy[2 * i + 1];
    } else {
        const long long per = (long long)g.nbx * g.nby;
        f = (int)(i / per);
        const int r = (int)(i - f * per);
        y = (r / g.nbx) << LOG2, x = (r % g.nbx) << LOG2;
    }
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

static int launch_satd(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb, ptrdiff_t fs_a, ptrdiff_t fs_b, int log2size, const SatdGrid &g,
                       int32_t *out, void *stream)
{
    if (g.n == 0) return 0;
    const unsigned grid = (unsigned)((g.n + 127) / 128);
    if (log2size == 1) return launch(satd_kernel<1>, grid, 128, 0, stream, a, sa, b, sb, fs_a, fs_b, g, out);
    if (log2size == 2) return launch(satd_kernel<2>, grid, 128, 0, stream, a, sa, b, sb, fs_a, fs_b, g, out);
    return launch(satd_kernel<3>, grid, 128, 0, stream, a, sa, b, sb, fs_a, fs_b, g, out);
}

extern "C" int hevcasm_hadamard_satd_batch(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb, int log2size, const int16_t *blk_xy, int n,
                                           int32_t *satd, void *stream)
{
    if (log2size < 1 || log2size > 3 || n < 0 || (n > 0 && !blk_xy)) return HEVCASM_ERR_ARGUMENT;
    SatdGrid g{blk_xy, 0, 0, n};
    return launch_satd(a, sa, b, sb, 0, 0, log2size, g, satd, stream);
}

extern "C" int hevcasm_hadamard_satd_frames(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb, int width, int height, int log2size, int n_frames,
                                            ptrdiff_t fs_a, ptrdiff_t fs_b, int32_t *satd, void *stream)
{
    if (log2size < 1 || log2size > 3 || n_frames < 0 || width < 0 || height < 0) return HEVCASM_ERR_ARGUMENT;
    SatdGrid g{nullptr, width >> log2size, height >> log2size, 0};
    g.n = (long long)g.nbx * g.nby * n_frames;
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
