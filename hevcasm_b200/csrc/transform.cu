// hevcasm_b200 - forward transforms and inverse transform + add, 4x4 (DCT, DST) .. 32x32, for sm_100a.
//
// Reference semantics (kupix/hevcasm): hevcasm_transform = residual_decode.c:855-892 over the stages :592-852;
// hevcasm_inverse_transform_add = residual_decode.c:371-413 over the stages :69-347 and hevcasm_add_residual :359-368.
// Arithmetic building blocks and the exactness argument are in transform.cuh.
//
// Kernel shapes:
//   * 4x4 and 8x8: one thread owns a whole block in registers (both stages and the transpose between them are
//     register renaming); consecutive lanes take consecutive blocks of a block row, so plane accesses coalesce.
//   * 16x16 and 32x32: a warp owns 64 block columns (4 resp. 2 blocks).  Each lane runs two 1-D transforms per stage
//     on int16 pairs; the stage-to-stage transpose goes through a conflict-free padded shared-memory tile private to
//     the warp (only __syncwarp, no CTA barrier).
// The second inverse stage skips the reference's clip to int16: |value| <= 23040 there, so the clip can never bind,
// and for the same reason the int16 truncation inside hevcasm_clip (:350-356) is the identity.
#include "transform.cuh"
#include "tma.cuh"
#include "umma.cuh"
#ifdef HEVCASM_EXPERIMENTS
#include "transform_imma.cuh"
#endif

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace hv {
using namespace tr;

// where block i of a batch lives: explicit list, or raster order over a regular grid of n_frames planes
// unsigned 32-bit division by a run-time constant as multiply-high + shifts (Granlund-Montgomery, round-up variant:
// exact for every 32-bit numerator): 4 integer instructions instead of the ~20 (32-bit) / ~100 (64-bit) of a division
struct FastDiv {
    uint32_t d, m, s;   // divisor, magic multiplier, shift;  s == 0 <=> d == 1
    __host__ static FastDiv make(uint32_t d)
    {
        FastDiv f{d ? d : 1u, 0, 0};
        while ((1ull << f.s) < f.d) ++f.s;   // ceil(log2 d)
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

// (lo16(a >> S), lo16(b >> S)) in one word, the truncating store of the forward transforms.  Written as LEFT shifts by 16 - S and a PRMT of the
// high halves: ptxas is free to put left shifts on the FMA pipe (IMAD.SHL) - the fused pipelines run against the ALU pipe, which the right
// shifts (SHF) share with PRMT, IADD3 and the saturating packs.
template <int S>
__device__ __forceinline__ uint32_t shr_lolo(int a, int b)
{
    static_assert(S >= 0 && S <= 16, "shift");
    return __byte_perm((uint32_t)a << (16 - S), (uint32_t)b << (16 - S), 0x7632);
}

struct BlockGrid {
    const int16_t *blk_xy;
    int nbx, nby;
    long long n;
    FastDiv per_frame, per_row;   // blocks per frame / per block row (regular grids with n < 2^32; finish() fills them)
    bool fast;
    int desc_w = 2;               // int16 per list entry: (x, y), or (x, y, frame) for the *_list_frames forms
    void finish()
    {
        fast = !blk_xy && n > 0 && n < (1ll << 32) && nbx > 0 && nby > 0;
        if (fast) per_frame = FastDiv::make((uint32_t)(nbx * nby)), per_row = FastDiv::make((uint32_t)nbx);
    }
    // REGULAR: the caller's kernel variant is only ever launched on regular grids (the plane-aligned variants), so the list branch folds away
    template <bool REGULAR = false>
    __device__ __forceinline__ void locate(long long i, int log2, int &x, int &y, int &f) const
    {
        if (!REGULAR && blk_xy) {
            const int16_t *e = blk_xy + i * desc_w;
            x = e[0], y = e[1], f = desc_w == 3 ? e[2] : 0;
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

// ---- alignment-agnostic row access (NW 32-bit words) ---------------------------------------------------
// PA ("plane aligned", decided on the host): base pointers, row strides and frame strides are multiples of 16 bytes and the
// block grid is regular, so a row segment of NW words is naturally aligned to its own size - no run-time alignment dispatch.
template <int NW, bool PA = false>
__device__ __forceinline__ void load_words(const void *ptr, uint32_t *w)
{
    if (PA) {
        if (NW % 4 == 0) {
#pragma unroll
            for (int i = 0; i < NW / 4; ++i) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(ptr) + i);
                w[4 * i] = v.x, w[4 * i + 1] = v.y, w[4 * i + 2] = v.z, w[4 * i + 3] = v.w;
            }
        } else if (NW == 2) {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(ptr));
            w[0] = v.x, w[1] = v.y;
        } else {
#pragma unroll
            for (int i = 0; i < NW; ++i) w[i] = __ldg(reinterpret_cast<const uint32_t *>(ptr) + i);
        }
        return;
    }
    const uintptr_t a = (uintptr_t)ptr;
    if (NW % 4 == 0 && (a & 15) == 0) {
#pragma unroll
        for (int i = 0; i < NW / 4; ++i) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(ptr) + i);
            w[4 * i] = v.x, w[4 * i + 1] = v.y, w[4 * i + 2] = v.z, w[4 * i + 3] = v.w;
        }
    } else if (NW % 2 == 0 && (a & 7) == 0) {
#pragma unroll
        for (int i = 0; i < NW / 2; ++i) {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(ptr) + i);
            w[2 * i] = v.x, w[2 * i + 1] = v.y;
        }
    } else if ((a & 3) == 0) {
#pragma unroll
        for (int i = 0; i < NW; ++i) w[i] = __ldg(reinterpret_cast<const uint32_t *>(ptr) + i);
    } else if ((a & 1) == 0) {
        const uint16_t *p = reinterpret_cast<const uint16_t *>(ptr);
#pragma unroll
        for (int i = 0; i < NW; ++i) w[i] = (uint32_t)__ldg(p + 2 * i) | ((uint32_t)__ldg(p + 2 * i + 1) << 16);
    } else {
        const uint8_t *p = reinterpret_cast<const uint8_t *>(ptr);
#pragma unroll
        for (int i = 0; i < NW; ++i)
            w[i] = (uint32_t)__ldg(p + 4 * i) | ((uint32_t)__ldg(p + 4 * i + 1) << 8) | ((uint32_t)__ldg(p + 4 * i + 2) << 16) |
                   ((uint32_t)__ldg(p + 4 * i + 3) << 24);
    }
}

template <int NW, bool PA = false>
__device__ __forceinline__ void store_words(void *ptr, const uint32_t *w)
{
    if (PA) {
        if (NW % 4 == 0) {
#pragma unroll
            for (int i = 0; i < NW / 4; ++i) reinterpret_cast<uint4 *>(ptr)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        } else if (NW == 2) {
            *reinterpret_cast<uint2 *>(ptr) = make_uint2(w[0], w[1]);
        } else {
#pragma unroll
            for (int i = 0; i < NW; ++i) reinterpret_cast<uint32_t *>(ptr)[i] = w[i];
        }
        return;
    }
    const uintptr_t a = (uintptr_t)ptr;
    if (NW % 4 == 0 && (a & 15) == 0) {
#pragma unroll
        for (int i = 0; i < NW / 4; ++i) reinterpret_cast<uint4 *>(ptr)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    } else if (NW % 2 == 0 && (a & 7) == 0) {
#pragma unroll
        for (int i = 0; i < NW / 2; ++i) reinterpret_cast<uint2 *>(ptr)[i] = make_uint2(w[2 * i], w[2 * i + 1]);
    } else if ((a & 3) == 0) {
#pragma unroll
        for (int i = 0; i < NW; ++i) reinterpret_cast<uint32_t *>(ptr)[i] = w[i];
    } else if ((a & 1) == 0) {
        uint16_t *p = reinterpret_cast<uint16_t *>(ptr);
#pragma unroll
        for (int i = 0; i < NW; ++i) p[2 * i] = (uint16_t)w[i], p[2 * i + 1] = (uint16_t)(w[i] >> 16);
    } else {
        uint8_t *p = reinterpret_cast<uint8_t *>(ptr);
#pragma unroll
        for (int i = 0; i < NW; ++i)
            p[4 * i] = (uint8_t)w[i], p[4 * i + 1] = (uint8_t)(w[i] >> 8), p[4 * i + 2] = (uint8_t)(w[i] >> 16), p[4 * i + 3] = (uint8_t)(w[i] >> 24);
    }
}

// four reconstructed samples: clip8(pred byte + (r >> 12)), packed
__device__ __forceinline__ uint32_t recon_word(uint32_t pw, const int *r)
{
    // the predictor byte is picked out of its word and added by one IDP.4A (pattern 1 in byte j) on the FMA pipe
    return pack_sat_u8((int)dp4a_uu(pw, 0x00000001u, (uint32_t)(r[0] >> 12)), (int)dp4a_uu(pw, 0x00000100u, (uint32_t)(r[1] >> 12)),
                       (int)dp4a_uu(pw, 0x00010000u, (uint32_t)(r[2] >> 12)), (int)dp4a_uu(pw, 0x01000000u, (uint32_t)(r[3] >> 12)));
}

// ================================================================================================ 4x4 / 8x8

constexpr int SMALL_NT = 128;

// Coefficient arrays are N*N contiguous int16 per block, so a thread-per-block kernel would touch them with 16-byte
// accesses 32 or 128 bytes apart - half-used sectors on every request (ncu: 1.6x L2 over-fetch for the 8x8 inverse,
// profiles/r01_transforms.md).  Instead a warp moves the 32 consecutive blocks it owns (1 KB / 4 KB contiguous) with fully
// coalesced 128-bit accesses and transposes them through a swizzled, conflict-free shared-memory tile private to the warp.
template <int LOG2>
struct SmallIo {
    static constexpr int N = 1 << LOG2, HW = N / 2, CH = N * N * 2 / 16;  // 16-byte chunks per block: 2 (4x4) or 8 (8x8)
    static constexpr int WARP_CHUNKS = 32 * CH;
    __device__ static __forceinline__ int pos(int b, int c) { return b * CH + (CH == 8 ? (c ^ (b & 7)) : (c ^ ((b >> 2) & 1))); }

    // global (blocks first .. first+31, clipped to n) -> Cw of this lane's block
    __device__ static __forceinline__ void load(const int16_t *coeffs, long long first, long long n, int lane, int4 *sm, uint32_t (&Cw)[N][HW])
    {
        const int4 *g = reinterpret_cast<const int4 *>(coeffs + first * (N * N));
        const long long valid_chunks = (n - first < 32 ? n - first : 32) * CH;
        int4 v[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const int j = lane + 32 * k;
            v[k] = j < valid_chunks ? ldg_stream(g + j) : make_int4(0, 0, 0, 0);
        }
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const int j = lane + 32 * k;
            sm[pos(j / CH, j % CH)] = v[k];
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int4 t = sm[pos(lane, c)];
            if (CH == 8) {
                Cw[c][0] = t.x, Cw[c][1] = t.y, Cw[c][HW - 2] = t.z, Cw[c][HW - 1] = t.w;
            } else {
                Cw[2 * c][0] = t.x, Cw[2 * c][1] = t.y, Cw[(2 * c + 1) % N][0] = t.z, Cw[(2 * c + 1) % N][1] = t.w;
            }
        }
        __syncwarp();
    }

    // Yw of this lane's block -> global
    __device__ static __forceinline__ void store(int16_t *coeffs, long long first, long long n, int lane, int4 *sm, const uint32_t (&Yw)[N][HW])
    {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            int4 t;
            if (CH == 8) t = make_int4((int)Yw[c][0], (int)Yw[c][1], (int)Yw[c][HW - 2], (int)Yw[c][HW - 1]);
            else t = make_int4((int)Yw[2 * c][0], (int)Yw[2 * c][1], (int)Yw[(2 * c + 1) % N][0], (int)Yw[(2 * c + 1) % N][1]);
            sm[pos(lane, c)] = t;
        }
        __syncwarp();
        int4 *g = reinterpret_cast<int4 *>(coeffs + first * (N * N));
        const long long valid_chunks = (n - first < 32 ? n - first : 32) * CH;
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const int j = lane + 32 * k;
            if (j < valid_chunks) stg_stream(g + j, sm[pos(j / CH, j % CH)]);
        }
        __syncwarp();
    }
};

// forward transform of the block at src (row stride `stride`), result in coefficient memory order: Yw[v][u/2] = (Y[v][u], Y[v][u+1])
template <int LOG2, bool DST, bool PA>
__device__ __forceinline__ void small_fwd_core(const int16_t *src, ptrdiff_t stride, uint32_t (&Yw)[1 << LOG2][(1 << LOG2) / 2])
{
    constexpr int N = 1 << LOG2, HW = N / 2, S1 = fwd_shift1(LOG2), S2 = fwd_shift2(LOG2);
    uint32_t X[N][HW];
#pragma unroll
    for (int r = 0; r < N; ++r) load_words<HW, PA>(src + (ptrdiff_t)r * stride, X[r]);

    // stage 1 (along x): A[u][y]; kept as vertical pairs Aw[u][y/2] = (A[u][y], A[u][y+1]), the operand order stage 2 needs
    uint32_t Aw[N][HW];
#pragma unroll
    for (int r = 0; r < N; r += 2) {
        int a0[N], a1[N];
        fwd_matrix<N, DST>(X[r], a0, 1 << (S1 - 1));
        fwd_matrix<N, DST>(X[r + 1], a1, 1 << (S1 - 1));
#pragma unroll
        for (int u = 0; u < N; ++u) Aw[u][r / 2] = shr_lolo<S1>(a0[u], a1[u]);  // truncating, residual_decode.c:674-682
    }
    // stage 2 (along y): Y[v][u], emitted as horizontal pairs (Y[v][u], Y[v][u+1]) = the memory order of coeffs[v*N+u]
#pragma unroll
    for (int u = 0; u < N; u += 2) {
        int b0[N], b1[N];
        fwd_matrix<N, DST>(Aw[u], b0, 1 << (S2 - 1));
        fwd_matrix<N, DST>(Aw[u + 1], b1, 1 << (S2 - 1));
#pragma unroll
        for (int v = 0; v < N; ++v) Yw[v][u / 2] = shr_lolo<S2>(b0[v], b1[v]);
    }
}

// The same with the residual formed on the fly from two 8-bit blocks (src - pred): stage 1 runs on the bytes themselves (IDP.4A with T
// for src and -T for pred), so the 9-bit difference never exists and the instruction count equals the int16 form's.  This is the batched
// counterpart of f265_lbd_dct_8_avx2 (reference f265/dct.asm:561), which also takes two 8-bit blocks.  pw = the predictor block's rows.
template <int LOG2, bool DST, bool PA>
__device__ __forceinline__ void small_fwd_core_planes(const uint8_t *src, ptrdiff_t ss, const uint32_t (&pw)[1 << LOG2][(1 << LOG2) / 4],
                                                      uint32_t (&Yw)[1 << LOG2][(1 << LOG2) / 2])
{
    constexpr int N = 1 << LOG2, HW = N / 2, S1 = fwd_shift1(LOG2), S2 = fwd_shift2(LOG2);
    uint32_t S[N][N / 4];
#pragma unroll
    for (int r = 0; r < N; ++r) load_words<N / 4, PA>(src + (ptrdiff_t)r * ss, S[r]);
    uint32_t Aw[N][HW];
#pragma unroll
    for (int r = 0; r < N; r += 2) {
        int a0[N], a1[N];
        fwd_matrix_bytes<N, DST>(S[r], pw[r], a0, 1 << (S1 - 1));
        fwd_matrix_bytes<N, DST>(S[r + 1], pw[r + 1], a1, 1 << (S1 - 1));
#pragma unroll
        for (int u = 0; u < N; ++u) Aw[u][r / 2] = shr_lolo<S1>(a0[u], a1[u]);
    }
#pragma unroll
    for (int u = 0; u < N; u += 2) {
        int b0[N], b1[N];
        fwd_matrix<N, DST>(Aw[u], b0, 1 << (S2 - 1));
        fwd_matrix<N, DST>(Aw[u + 1], b1, 1 << (S2 - 1));
#pragma unroll
        for (int v = 0; v < N; ++v) Yw[v][u / 2] = shr_lolo<S2>(b0[v], b1[v]);
    }
}

template <int LOG2, bool DST, bool PA>
__global__ void __launch_bounds__(SMALL_NT) small_fwd_kernel(int16_t *__restrict__ coeffs, const int16_t *__restrict__ res, ptrdiff_t stride,
                                                             ptrdiff_t fs, BlockGrid g)
{
    using Io = SmallIo<LOG2>;
    constexpr int N = 1 << LOG2, HW = N / 2;
    __shared__ __align__(16) int4 io[SMALL_NT / 32][Io::WARP_CHUNKS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long i = (long long)blockIdx.x * SMALL_NT + threadIdx.x, first = i - lane;
    if (first >= g.n) return;  // whole warp out of range
    uint32_t Yw[N][HW];
    if (i < g.n) {
        int x, y, f;
        g.template locate<PA>(i, LOG2, x, y, f);
        small_fwd_core<LOG2, DST, PA>(res + f * fs + (ptrdiff_t)y * stride + x, stride, Yw);
    }
    Io::store(coeffs, first, g.n, lane, io[warp], Yw);
}

// dst = clip8(pred + inverse transform of Cw); Cw in coefficient memory order
// the predictor block, fetched before the arithmetic starts so that its latency hides behind the transform
template <int N, bool PA>
__device__ __forceinline__ void load_pred(const uint8_t *pp, ptrdiff_t sp, uint32_t (&pw)[N][N / 4])
{
#pragma unroll
    for (int r = 0; r < N; ++r) load_words<N / 4, PA>(pp + (ptrdiff_t)r * sp, pw[r]);
}

template <int LOG2, bool DST, bool PA>
__global__ void __launch_bounds__(SMALL_NT) small_fwd_planes_kernel(int16_t *__restrict__ coeffs, const uint8_t *__restrict__ src, ptrdiff_t ss, ptrdiff_t fs_src,
                                                                    const uint8_t *__restrict__ pred, ptrdiff_t sp, ptrdiff_t fs_pred, BlockGrid g)
{
    using Io = SmallIo<LOG2>;
    constexpr int N = 1 << LOG2, HW = N / 2;
    __shared__ __align__(16) int4 io[SMALL_NT / 32][Io::WARP_CHUNKS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long i = (long long)blockIdx.x * SMALL_NT + threadIdx.x, first = i - lane;
    if (first >= g.n) return;  // whole warp out of range
    uint32_t Yw[N][HW];
    if (i < g.n) {
        int x, y, f;
        g.template locate<PA>(i, LOG2, x, y, f);
        uint32_t pw[N][N / 4];
        load_pred<N, PA>(pred + f * fs_pred + (ptrdiff_t)y * sp + x, sp, pw);
        small_fwd_core_planes<LOG2, DST, PA>(src + f * fs_src + (ptrdiff_t)y * ss + x, ss, pw, Yw);
    }
    Io::store(coeffs, first, g.n, lane, io[warp], Yw);
}

template <int LOG2, bool DST, bool PA>
__device__ __forceinline__ void small_inv_core(const uint32_t (&Cw)[1 << LOG2][(1 << LOG2) / 2], uint8_t *dp, ptrdiff_t sd,
                                               const uint32_t (&pw)[1 << LOG2][(1 << LOG2) / 4])
{
    constexpr int N = 1 << LOG2, HW = N / 2;
    // stage 1 (along v, shift 7, clip16): B[u][y]
    int B[N][N];
#pragma unroll
    for (int j = 0; j < HW; ++j) {
        uint32_t pl[HW], ph[HW];
        static_for<0, HW>([&](auto kk) {
            constexpr int k = HV_V(kk), r0 = DST ? 2 * k : pair_row(N, k, 0), r1 = DST ? 2 * k + 1 : pair_row(N, k, 1);
            pl[k] = lolo(Cw[r0][j], Cw[r1][j]);
            ph[k] = hihi(Cw[r0][j], Cw[r1][j]);
        });
        if (DST) {
            inv_dst4(pl, B[2 * j], 64);
            inv_dst4(ph, B[2 * j + 1], 64);
        } else {
            InvBfly<N>::run(pl, B[2 * j], 64);
            InvBfly<N>::run(ph, B[2 * j + 1], 64);
        }
    }
    // stage 2 (along u, shift 12) + add to the predictor
#pragma unroll
    for (int r = 0; r < N; ++r) {
        uint32_t p[HW];
        static_for<0, HW>([&](auto kk) {
            constexpr int k = HV_V(kk), r0 = DST ? 2 * k : pair_row(N, k, 0), r1 = DST ? 2 * k + 1 : pair_row(N, k, 1);
            p[k] = pack_sat_s16(B[r0][r] >> 7, B[r1][r] >> 7);  // the clip to int16 of residual_decode.c stage 1
        });
        int o[N];
        if (DST) inv_dst4(p, o, 2048);
        else InvBfly<N>::run(p, o, 2048);
        uint32_t ow[N / 4];
#pragma unroll
        for (int k = 0; k < N / 4; ++k) ow[k] = recon_word(pw[r][k], o + 4 * k);
        store_words<N / 4, PA>(dp + (ptrdiff_t)r * sd, ow);
    }
}

template <int LOG2, bool DST, bool PA>
__global__ void __launch_bounds__(SMALL_NT) small_inv_kernel(uint8_t *__restrict__ dst, ptrdiff_t sd, const uint8_t *__restrict__ pred, ptrdiff_t sp,
                                                             ptrdiff_t fs_dst, ptrdiff_t fs_pred, const int16_t *__restrict__ coeffs, BlockGrid g)
{
    using Io = SmallIo<LOG2>;
    constexpr int N = 1 << LOG2, HW = N / 2;
    __shared__ __align__(16) int4 io[SMALL_NT / 32][Io::WARP_CHUNKS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long i = (long long)blockIdx.x * SMALL_NT + threadIdx.x, first = i - lane;
    if (first >= g.n) return;
    const bool valid = i < g.n;
    int x = 0, y = 0, f = 0;
    uint32_t pw[N][N / 4];
    if (valid) {
        g.template locate<PA>(i, LOG2, x, y, f);
        load_pred<N, PA>(pred + f * fs_pred + (ptrdiff_t)y * sp + x, sp, pw);
    }
    uint32_t Cw[N][HW];
    Io::load(coeffs, first, g.n, lane, io[warp], Cw);
    if (valid) small_inv_core<LOG2, DST, PA>(Cw, dst + f * fs_dst + (ptrdiff_t)y * sd + x, sd, pw);
}

// ================================================================================================ 16x16 / 32x32

constexpr int BIG_NT = 128;

template <int LOG2>
struct BigGeom {
    static constexpr int N = 1 << LOG2, HW = N / 2, WB = 64 / N;   // words per row, blocks per warp
    static constexpr int PITCH = HW + 4;                           // 20 / 12 words: conflict-free 128-bit row access
    static constexpr int BLK_STRIDE = N * PITCH + (N == 32 ? 16 : 8);
    static constexpr int WARP_WORDS = WB * BLK_STRIDE;
};

// Inverse of one block by the HW lanes that own it.  W[v] = (C[v][2uw], C[v][2uw+1]) - this lane's two coefficient columns.
// Must be called by all 32 lanes of the warp (it contains a __syncwarp); lanes with !valid only take part in the barrier.
template <int LOG2, bool PA>
__device__ __forceinline__ void big_inv_core(uint32_t *tmp, int b, int uw, bool valid, const uint32_t (&W)[1 << LOG2], uint8_t *dp, ptrdiff_t sd,
                                             const uint8_t *pp, ptrdiff_t sp)
{
    using G = BigGeom<LOG2>;
    constexpr int N = G::N, HW = G::HW;
    if (N == 32 && valid) {  // 32x32: the stage-2 predictor rows are pulled into L1 now (no registers held across stage 1)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pp + (ptrdiff_t)uw * sp));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pp + (ptrdiff_t)(uw + HW) * sp));
    }
    if (valid) {  // stage 1: columns 2*uw and 2*uw+1
        uint32_t p0[HW], p1[HW];
        int o0[N], o1[N];
        static_for<0, HW>([&](auto kk) {
            constexpr int k = HV_V(kk);
            p0[k] = lolo(W[pair_row(N, k, 0)], W[pair_row(N, k, 1)]);
            p1[k] = hihi(W[pair_row(N, k, 0)], W[pair_row(N, k, 1)]);
        });
        InvBfly2<N>::run(p0, p1, o0, o1, 64);   // both columns against each coefficient pair
#pragma unroll
        for (int r = 0; r < N; ++r) tmp[b * G::BLK_STRIDE + r * G::PITCH + uw] = pack_sat_s16(o0[r] >> 7, o1[r] >> 7);
    }
    // the two predictor rows of stage 2 are requested here, between the stages: the coefficient words are dead by now, and
    // the transpose + the first row's butterfly hide the latency (ncu: 24 % of all stall samples sat on the first use of a
    // predictor word loaded right before it, profiles/r01_transforms.md)
    // (16x16 only: at 32x32 the 16 extra live registers and the unrolled second stage cost more occupancy than the latency they hide)
    constexpr bool EARLY = N == 16;
    uint32_t pw2[2][N / 4];
    if (EARLY && valid) {
#pragma unroll
        for (int h = 0; h < 2; ++h) load_words<N / 4, PA>(pp + (ptrdiff_t)(uw + h * HW) * sp, pw2[h]);
    }
    __syncwarp();
    if (valid) {  // stage 2: this lane owns rows uw and uw + N/2 of block b
#pragma unroll(EARLY ? 2 : 1)
        for (int h = 0; h < 2; ++h) {
            const int r = uw + h * HW;
            uint32_t Bw[HW];
            const uint4 *row = reinterpret_cast<const uint4 *>(tmp + b * G::BLK_STRIDE + r * G::PITCH);
#pragma unroll
            for (int k = 0; k < HW / 4; ++k) {
                const uint4 v = row[k];
                Bw[4 * k] = v.x, Bw[4 * k + 1] = v.y, Bw[4 * k + 2] = v.z, Bw[4 * k + 3] = v.w;
            }
            uint32_t p[HW];
            static_for<0, HW>([&](auto kk) {
                constexpr int k = HV_V(kk), r0 = pair_row(N, k, 0), r1 = pair_row(N, k, 1);
                p[k] = (r0 & 1) ? hihi(Bw[r0 / 2], Bw[r1 / 2]) : lolo(Bw[r0 / 2], Bw[r1 / 2]);
            });
            int o[N];
            InvBfly<N>::run(p, o, 2048);
            uint32_t ow[N / 4];
            if (!EARLY) load_words<N / 4, PA>(pp + (ptrdiff_t)r * sp, pw2[0]);
#pragma unroll
            for (int k = 0; k < N / 4; ++k) ow[k] = recon_word(pw2[EARLY ? h : 0][k], o + 4 * k);
            store_words<N / 4, PA>(dp + (ptrdiff_t)r * sd, ow);
        }
    }
}

// (A persistent variant with the coefficient blocks prefetched into shared memory by cp.async.bulk, two stages per warp, was measured in
//  round 2: 165 vs 130 us (32x32) and 136 vs 118 us (16x16) - 128 registers and 16 warps per SM lose more than the prefetch gains; capping
//  this kernel at 80 registers for 6 CTAs per SM changes nothing either (130.0 us, 20 bytes of spills): it is bound by its pipes.)
template <int LOG2, bool PA, int GROUPS = 1>
__global__ void __launch_bounds__(BIG_NT) big_inv_kernel(uint8_t *__restrict__ dst, ptrdiff_t sd, const uint8_t *__restrict__ pred, ptrdiff_t sp,
                                                         ptrdiff_t fs_dst, ptrdiff_t fs_pred, const int16_t *__restrict__ coeffs, BlockGrid g)
{
    using G = BigGeom<LOG2>;
    constexpr int N = G::N, HW = G::HW;
    __shared__ __align__(16) uint32_t tmp_all[BIG_NT / 32][G::WARP_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = lane / HW, uw = lane % HW;
    // GROUPS block groups per warp: the coefficients of all of them are requested before the first is transformed
    uint32_t W[GROUPS][N];
    long long gb[GROUPS];
#pragma unroll
    for (int k = 0; k < GROUPS; ++k) {
        gb[k] = (((long long)blockIdx.x * GROUPS + k) * (BIG_NT / 32) + warp) * G::WB + b;
        if (gb[k] < g.n) {
            const uint32_t *cw = reinterpret_cast<const uint32_t *>(coeffs + gb[k] * (N * N)) + uw;
#pragma unroll
            for (int v = 0; v < N; ++v) W[k][v] = __ldg(cw + v * HW);
        }
    }
#pragma unroll
    for (int k = 0; k < GROUPS; ++k) {
        const bool valid = gb[k] < g.n;
        int x = 0, y = 0, f = 0;
        if (valid) g.template locate<PA>(gb[k], LOG2, x, y, f);
        big_inv_core<LOG2, PA>(tmp_all[warp], b, uw, valid, W[k], dst + f * fs_dst + (ptrdiff_t)y * sd + x, sd, pred + f * fs_pred + (ptrdiff_t)y * sp + x, sp);
        if (k + 1 < GROUPS) __syncwarp();
    }
}

// Forward transform of one block by the HW lanes that own it; on return W[v] = (Y[v][2uw], Y[v][2uw+1]).
template <int LOG2, bool PA>
__device__ __forceinline__ void big_fwd_core(uint32_t *tmp, int b, int uw, bool valid, const int16_t *src, ptrdiff_t stride, uint32_t (&W)[1 << LOG2])
{
    using G = BigGeom<LOG2>;
    constexpr int N = G::N, HW = G::HW, S1 = fwd_shift1(LOG2), S2 = fwd_shift2(LOG2);
    if (valid) {  // stage 1: rows uw and uw + N/2 of block b, along x
        // 16x16: both rows are requested before the first is transformed (the loop is unrolled); 32x32 keeps one copy of the butterfly code
        uint32_t Xboth[N == 16 ? 2 : 1][HW];
        if (N == 16) {
#pragma unroll
            for (int h = 0; h < 2; ++h) load_words<HW, PA>(src + (ptrdiff_t)(uw + h * HW) * stride, Xboth[h % (N == 16 ? 2 : 1)]);
        }
#pragma unroll(N == 16 ? 2 : 1)
        for (int h = 0; h < 2; ++h) {
            const int r = uw + h * HW;
            uint32_t Xw[HW];
            if (N == 16) {
#pragma unroll
                for (int k = 0; k < HW; ++k) Xw[k] = Xboth[h % (N == 16 ? 2 : 1)][k];
            } else {
                load_words<HW, PA>(src + (ptrdiff_t)r * stride, Xw);
            }
            int xv[N], a[N];
            uint32_t range = 0;
#pragma unroll
            for (int k = 0; k < HW; ++k) xv[2 * k] = s16lo(Xw[k]), xv[2 * k + 1] = s16hi(Xw[k]), range = fits15_acc(range, Xw[k]);
            if (N == 32 && fits15(range)) FwdBflyPacked<N>::run(xv, a, 1 << (S1 - 1));   // top-level odd part on IDP.2A (transform.cuh); 16x16: measured slower (141 vs 130 us)
            else FwdBfly<N>::run(xv, a, 1 << (S1 - 1));
            uint4 *row = reinterpret_cast<uint4 *>(tmp + b * G::BLK_STRIDE + r * G::PITCH);
#pragma unroll
            for (int k = 0; k < HW / 4; ++k)
                row[k] = make_uint4(shr_lolo<S1>(a[8 * k], a[8 * k + 1]), shr_lolo<S1>(a[8 * k + 2], a[8 * k + 3]),
                                    shr_lolo<S1>(a[8 * k + 4], a[8 * k + 5]), shr_lolo<S1>(a[8 * k + 6], a[8 * k + 7]));
        }
    }
    __syncwarp();
    if (valid) {  // stage 2: columns 2*uw and 2*uw+1, along y
        int x0[N], c0[N];
        uint32_t range = 0;
#pragma unroll
        for (int r = 0; r < N; ++r) {
            const uint32_t w = tmp[b * G::BLK_STRIDE + r * G::PITCH + uw];
            x0[r] = s16lo(w), range = fits15_acc(range, w);
        }
        const bool packed = N == 32 && fits15(range);   // both columns of this lane
        if (packed) FwdBflyPacked<N>::run(x0, c0, 1 << (S2 - 1));
        else FwdBfly<N>::run(x0, c0, 1 << (S2 - 1));
#pragma unroll
        for (int r = 0; r < N; ++r) x0[r] = s16hi(tmp[b * G::BLK_STRIDE + r * G::PITCH + uw]);
        int c1[N];
        if (packed) FwdBflyPacked<N>::run(x0, c1, 1 << (S2 - 1));
        else FwdBfly<N>::run(x0, c1, 1 << (S2 - 1));
#pragma unroll
        for (int v = 0; v < N; ++v) W[v] = shr_lolo<S2>(c0[v], c1[v]);
    }
    __syncwarp();  // tmp may be reused by the caller
}

template <int LOG2, bool PA>
__global__ void __launch_bounds__(BIG_NT) big_fwd_kernel(int16_t *__restrict__ coeffs, const int16_t *__restrict__ res, ptrdiff_t stride, ptrdiff_t fs,
                                                         BlockGrid g)
{
    using G = BigGeom<LOG2>;
    constexpr int N = G::N, HW = G::HW;
    __shared__ __align__(16) uint32_t tmp_all[BIG_NT / 32][G::WARP_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = lane / HW, uw = lane % HW;
    const long long gb = ((long long)blockIdx.x * (BIG_NT / 32) + warp) * G::WB + b;
    const bool valid = gb < g.n;
    int x = 0, y = 0, f = 0;
    if (valid) g.template locate<PA>(gb, LOG2, x, y, f);
    uint32_t W[N];
    big_fwd_core<LOG2, PA>(tmp_all[warp], b, uw, valid, res + f * fs + (ptrdiff_t)y * stride + x, stride, W);
    if (valid) {
        uint32_t *out = reinterpret_cast<uint32_t *>(coeffs + gb * (N * N)) + uw;
#pragma unroll
        for (int v = 0; v < N; ++v) out[v * HW] = W[v];
    }
}

#include "transform_fwd_umma.cuh"  // forward 16x16 / 32x32: first stage on tcgen05, second in registers (the default for batches that fill the chip)
#ifdef HEVCASM_EXPERIMENTS
// measured, not adopted (profiles/r01_transforms.md); compiled into libhevcasm_b200_exp.so only
#include "transform_umma.cuh"      // 16x16 / 32x32 inverse, both stages on tcgen05 (needs BlockGrid, load_words, store_words)
#include "transform_inv_umma.cuh"  // inverse 16x16 / 32x32: first stage in registers, second on tcgen05

// ================================================================================================ 16x16 / 32x32 inverse on IMMA

constexpr int IMMA_NT = 128;

template <int LOG2, bool PA>
__global__ void __launch_bounds__(IMMA_NT) imma_inv_kernel(uint8_t *__restrict__ dst, ptrdiff_t sd, const uint8_t *__restrict__ pred, ptrdiff_t sp,
                                                           ptrdiff_t fs_dst, ptrdiff_t fs_pred, const int16_t *__restrict__ coeffs, BlockGrid grid)
{
    constexpr int N = 1 << LOG2, MT = N / 16, NTL = N / 8;  // m-tiles, n-tiles
    constexpr int PITCH = 2 * N + 16;                       // bytes per tile row: N int16 + 16 bytes of padding (conflict-free ldmatrix / row access)
    using MM = Imma<N>;
    __shared__ __align__(16) uint8_t tiles[IMMA_NT / 32][N * PITCH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    uint8_t *tile = tiles[warp];

    // constant fragments with the slot permutation baked in
    uint32_t A1[MT][MM::AR];   // stage 1: A[m = y][slot] = T[kperm(slot)][y]
    uint32_t B2[NTL][MM::BR];  // stage 2: B[slot][n = x] = T[kperm(slot)][x]
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int r = 0; r < MM::AR; ++r) {
            const int y = 16 * mt + g + 8 * (r & 1), s0 = 16 * (r >> 1) + 4 * t;
            A1[mt][r] = pack4(tcoef<N>(kperm(s0), y), tcoef<N>(kperm(s0 + 1), y), tcoef<N>(kperm(s0 + 2), y), tcoef<N>(kperm(s0 + 3), y));
        }
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
        for (int r = 0; r < MM::BR; ++r) {
            const int x = 8 * nt + g, s0 = 16 * r + 4 * t;
            B2[nt][r] = pack4(tcoef<N>(kperm(s0), x), tcoef<N>(kperm(s0 + 1), x), tcoef<N>(kperm(s0 + 2), x), tcoef<N>(kperm(s0 + 3), x));
        }

    const long long n_warps = (long long)gridDim.x * (IMMA_NT / 32);
    for (long long blk = (long long)blockIdx.x * (IMMA_NT / 32) + warp; blk < grid.n; blk += n_warps) {
        // ---- block -> padded shared tile (coalesced 128-bit loads)
        {
            const int4 *src = reinterpret_cast<const int4 *>(coeffs + blk * (N * N));
            constexpr int CPR = N / 8, TOTAL = N * CPR;  // 16-byte chunks per row / per block
#pragma unroll
            for (int k = 0; k < TOTAL / 32; ++k) {
                const int i = lane + 32 * k;
                *reinterpret_cast<int4 *>(tile + (i / CPR) * PITCH + (i % CPR) * 16) = ldg_stream(src + i);
            }
        }
        int x0, y0, f;
        grid.locate(blk, LOG2, x0, y0, f);
        // predictor segment of this lane for the epilogue: 16 samples of row (lane*16) / N ...
        constexpr int SEGW = N == 32 ? 16 : 8, SEGS = N / SEGW;  // samples per epilogue unit, units per row
        constexpr int EITER = N * SEGS / 32;                     // epilogue iterations per lane (2 for 32x32, 1 for 16x16)
        uint32_t pw[EITER][SEGW / 4];
#pragma unroll
        for (int e = 0; e < EITER; ++e) {
            const int u = lane + 32 * e, row = u / SEGS, seg = u % SEGS;
            load_words<SEGW / 4, PA>(pred + f * fs_pred + (ptrdiff_t)(y0 + row) * sp + x0 + SEGW * seg, pw[e]);
        }
        __syncwarp();

        // ---- stage 1: Bt[y][u], kept as packed int16 pairs w[mt][nt][h] = (Bt[y][8nt+2t], Bt[y][8nt+2t+1]), y = 16mt + g + 8h
        uint32_t w[MT][NTL][2];
#pragma unroll
        for (int nt = 0; nt < NTL; nt += (N == 32 ? 1 : 2)) {
            uint32_t r[4];
            uint32_t bh[2][MM::BR], bl[2][MM::BR];
            if (N == 32) {  // four 8x8 matrices down the 32 rows of one 8-column slab
                ldmatrix_x4_trans(r, tile + (8 * (lane >> 3) + (lane & 7)) * PITCH + nt * 16);
                bh[0][0] = __byte_perm(r[0], r[1], 0x7531), bl[0][0] = __byte_perm(r[0], r[1], 0x6420);
                bh[0][MM::BR - 1] = __byte_perm(r[2], r[3], 0x7531), bl[0][MM::BR - 1] = __byte_perm(r[2], r[3], 0x6420);
            } else {        // 16 rows: matrices (v 0-7, 8-15) x (slab nt, slab nt+1)
                ldmatrix_x4_trans(r, tile + (8 * ((lane >> 3) & 1) + (lane & 7)) * PITCH + (nt + (lane >> 4)) * 16);
                bh[0][0] = __byte_perm(r[0], r[1], 0x7531), bl[0][0] = __byte_perm(r[0], r[1], 0x6420);
                bh[1][0] = __byte_perm(r[2], r[3], 0x7531), bl[1][0] = __byte_perm(r[2], r[3], 0x6420);
            }
#pragma unroll
            for (int q = 0; q < (N == 32 ? 1 : 2); ++q)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    int dh[4], d[4];
                    const int zero[4] = {0, 0, 0, 0};
                    MM::s8s8(dh, A1[mt], bh[q], zero);
                    const int c[4] = {dh[0] * 256 + 64, dh[1] * 256 + 64, dh[2] * 256 + 64, dh[3] * 256 + 64};
                    MM::s8u8(d, A1[mt], bl[q], c);
                    w[mt][nt + q][0] = pack_sat_s16(d[0] >> 7, d[1] >> 7);
                    w[mt][nt + q][1] = pack_sat_s16(d[2] >> 7, d[3] >> 7);
                }
        }
        __syncwarp();  // every lane is done reading the coefficient tile: it now receives the residual

        // ---- stage 2: R[y][x] = sum_u Bt[y][u] T[u][x]; A fragments straight from the stage-1 registers
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            uint32_t ah[MM::AR], al[MM::AR];
#pragma unroll
            for (int r = 0; r < MM::AR; ++r) {
                const int h = r & 1, n0 = 2 * (r >> 1);  // row g + 8h; slots 16*(r>>1) + 4t.. <- n-tiles n0, n0+1
                ah[r] = __byte_perm(w[mt][n0][h], w[mt][n0 + 1][h], 0x7531);
                al[r] = __byte_perm(w[mt][n0][h], w[mt][n0 + 1][h], 0x6420);
            }
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
                int eh[4], e[4];
                const int zero[4] = {0, 0, 0, 0};
                MM::s8s8(eh, ah, B2[nt], zero);
                const int c[4] = {eh[0] * 256 + 2048, eh[1] * 256 + 2048, eh[2] * 256 + 2048, eh[3] * 256 + 2048};
                MM::u8s8(e, al, B2[nt], c);
                // |R| <= 23040: the reference's clip to int16 can never bind here
                *reinterpret_cast<uint32_t *>(tile + (16 * mt + g) * PITCH + (8 * nt + 2 * t) * 2) = pack16(e[0] >> 12, e[1] >> 12);
                *reinterpret_cast<uint32_t *>(tile + (16 * mt + g + 8) * PITCH + (8 * nt + 2 * t) * 2) = pack16(e[2] >> 12, e[3] >> 12);
            }
        }
        __syncwarp();

        // ---- epilogue: 16 residuals + 16 predictor samples -> 16 reconstructed bytes per lane and iteration
#pragma unroll
        for (int e = 0; e < EITER; ++e) {
            const int u = lane + 32 * e, row = u / SEGS, seg = u % SEGS;
            uint32_t rw[SEGW / 2];
#pragma unroll
            for (int k = 0; k < SEGW / 8; ++k) {
                const uint4 v = *reinterpret_cast<const uint4 *>(tile + row * PITCH + seg * (SEGW * 2) + 16 * k);
                rw[4 * k] = v.x, rw[4 * k + 1] = v.y, rw[4 * k + 2] = v.z, rw[4 * k + 3] = v.w;
            }
            uint32_t ow[SEGW / 4];
#pragma unroll
            for (int k = 0; k < SEGW / 4; ++k) {
                const uint32_t p = pw[e][k];
                ow[k] = pack_sat_u8((int)(p & 0xff) + s16lo(rw[2 * k]), (int)((p >> 8) & 0xff) + s16hi(rw[2 * k]), (int)((p >> 16) & 0xff) + s16lo(rw[2 * k + 1]),
                                    (int)(p >> 24) + s16hi(rw[2 * k + 1]));
            }
            store_words<SEGW / 4, PA>(dst + f * fs_dst + (ptrdiff_t)(y0 + row) * sd + x0 + SEGW * seg, ow);
        }
        __syncwarp();  // the tile is overwritten by the next block
    }
}

#endif  // HEVCASM_EXPERIMENTS

// ================================================================================================ fused residual pipeline
//
// forward transform -> quantize (levels + cbf written) -> inverse quantize -> inverse transform -> add to predictor, per block,
// without the coefficients ever leaving the chip.  Element semantics: the composition of the reference's hevcasm_transform,
// hevcasm_quantize (quantize.c:160-186), hevcasm_quantize_inverse (quantize.c:53-62) and hevcasm_inverse_transform_add.

struct QuantParams {
    int q_scale, q_shift, q_off /* offset << (shift-16) */, q_offn /* 2^shift - 1 - q_off */, iq_scale, iq_shift;
};
inline QuantParams make_quant_params(int q_scale, int q_shift, int q_offset, int iq_scale, int iq_shift)
{
    const int off = q_offset << (q_shift - 16);
    return QuantParams{q_scale, q_shift, off, (1 << q_shift) - 1 - off, iq_scale, iq_shift};
}

// Quantise then dequantise one word (two coefficients); the level word goes to `lv`, the OR of the level WORDS into cbf2 (cbf_fold turns
// it into the reference's return value).  Same integers as quantize.c:160-186 / :53-62 with fewer instructions:
//   * sign(x) * ((|x| * s + off) >> sh)  =  (x * s + (x < 0 ? 2^sh - 1 - off : off)) >> sh      (-floor(a / b) = floor((-a + b - 1) / b)),
//     which replaces abs, compare and negate by one select;
//   * in the quantiser's domain (|x| <= 2^15, s < 2^15, sh >= 16, off < 2^(sh-1), checked by the entry point) |level| <= 2^14, so the
//     reference's clip to int16 cannot bind and the levels feed the dequantiser without being re-extracted from the packed word;
//   * the OR of sign-extended levels is the sign extension of the OR of their 16-bit patterns, so one LOP3 per word does for cbf.
__device__ __forceinline__ uint32_t quant_dequant_word(uint32_t w, const QuantParams &q, uint32_t &lv, uint32_t &cbf2)
{
    const int x0 = s16lo(w), x1 = s16hi(w);
    const int l0 = (x0 * q.q_scale + (x0 < 0 ? q.q_offn : q.q_off)) >> q.q_shift, l1 = (x1 * q.q_scale + (x1 < 0 ? q.q_offn : q.q_off)) >> q.q_shift;
    lv = pack16(l0, l1);
    cbf2 |= lv;
    const int add = 1 << (q.iq_shift - 1);
    return pack_sat_s16((l0 * q.iq_scale + add) >> q.iq_shift, (l1 * q.iq_scale + add) >> q.iq_shift);
}
__device__ __forceinline__ int cbf_fold(uint32_t cbf2) { return (int)(short)((cbf2 & 0xffffu) | (cbf2 >> 16)); }

struct PipelineParams {
    uint8_t *rec;
    const uint8_t *pred;
    const uint8_t *src;          // "from planes" form: residual = src - pred, formed on the fly (res is null then)
    ptrdiff_t s_src, fs_src;
    const int16_t *res;
    int16_t *levels;
    int32_t *cbf;
    ptrdiff_t s_rec, s_pred, s_res, fs_rec, fs_pred, fs_res;
    QuantParams q;
};

template <int LOG2, bool DST, bool PA, bool PLANES = false>
__global__ void __launch_bounds__(SMALL_NT) small_pipeline_kernel(PipelineParams p, BlockGrid g)
{
    using Io = SmallIo<LOG2>;
    constexpr int N = 1 << LOG2, HW = N / 2;
    __shared__ __align__(16) int4 io[SMALL_NT / 32][Io::WARP_CHUNKS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long i = (long long)blockIdx.x * SMALL_NT + threadIdx.x, first = i - lane;
    if (first >= g.n) return;
    const bool valid = i < g.n;
    int x = 0, y = 0, f = 0;
    uint32_t Yw[N][HW], L[N][HW], pw[N][N / 4];
    uint32_t cbf = 0;
    if (valid) {
        g.template locate<PA>(i, LOG2, x, y, f);
        load_pred<N, PA>(p.pred + f * p.fs_pred + (ptrdiff_t)y * p.s_pred + x, p.s_pred, pw);
        if (PLANES) small_fwd_core_planes<LOG2, DST, PA>(p.src + f * p.fs_src + (ptrdiff_t)y * p.s_src + x, p.s_src, pw, Yw);
        else small_fwd_core<LOG2, DST, PA>(p.res + f * p.fs_res + (ptrdiff_t)y * p.s_res + x, p.s_res, Yw);
#pragma unroll
        for (int v = 0; v < N; ++v)
#pragma unroll
            for (int k = 0; k < HW; ++k) Yw[v][k] = quant_dequant_word(Yw[v][k], p.q, L[v][k], cbf);
        if (p.cbf) p.cbf[i] = cbf_fold(cbf);
    }
    Io::store(p.levels, first, g.n, lane, io[warp], L);
    if (valid) small_inv_core<LOG2, DST, PA>(Yw, p.rec + f * p.fs_rec + (ptrdiff_t)y * p.s_rec + x, p.s_rec, pw);
}

template <int LOG2, bool PA>
__global__ void __launch_bounds__(BIG_NT) big_pipeline_kernel(PipelineParams p, BlockGrid g)
{
    using G = BigGeom<LOG2>;
    constexpr int N = G::N, HW = G::HW;
    __shared__ __align__(16) uint32_t tmp_all[BIG_NT / 32][G::WARP_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = lane / HW, uw = lane % HW;
    const long long gb = ((long long)blockIdx.x * (BIG_NT / 32) + warp) * G::WB + b;
    const bool valid = gb < g.n;
    int x = 0, y = 0, f = 0;
    if (valid) g.template locate<PA>(gb, LOG2, x, y, f);
    uint32_t W[N];
    big_fwd_core<LOG2, PA>(tmp_all[warp], b, uw, valid, p.res + f * p.fs_res + (ptrdiff_t)y * p.s_res + x, p.s_res, W);
    uint32_t cbf = 0;
    if (valid) {
        uint32_t *lv = reinterpret_cast<uint32_t *>(p.levels + gb * (N * N)) + uw;
#pragma unroll
        for (int v = 0; v < N; ++v) {
            uint32_t L;
            W[v] = quant_dequant_word(W[v], p.q, L, cbf);
            lv[v * HW] = L;
        }
    }
    // OR across the HW lanes that share the block (HW = 8 or 16: groups are aligned, xor stays inside the group)
#pragma unroll
    for (int o = 1; o < HW; o <<= 1) cbf |= __shfl_xor_sync(0xffffffffu, cbf, o);
    if (valid && uw == 0 && p.cbf) p.cbf[gb] = cbf_fold(cbf);
    big_inv_core<LOG2, PA>(tmp_all[warp], b, uw, valid, W, p.rec + f * p.fs_rec + (ptrdiff_t)y * p.s_rec + x, p.s_rec,
                       p.pred + f * p.fs_pred + (ptrdiff_t)y * p.s_pred + x, p.s_pred);
}

// Second half of the 32x32 pipeline when the first runs on the tensor cores (ft::fwd_umma_kernel<5, true> has written the LEVELS): dequantise,
// inverse transform, add to the predictor; the coded-block flag is the OR of the level words the kernel reads anyway.
template <int LOG2, bool PA>
__global__ void __launch_bounds__(BIG_NT) big_dequant_inv_kernel(PipelineParams p, BlockGrid g)
{
    using G = BigGeom<LOG2>;
    constexpr int N = G::N, HW = G::HW;
    __shared__ __align__(16) uint32_t tmp_all[BIG_NT / 32][G::WARP_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = lane / HW, uw = lane % HW;
    const long long gb = ((long long)blockIdx.x * (BIG_NT / 32) + warp) * G::WB + b;
    const bool valid = gb < g.n;
    uint32_t W[N];
    uint32_t cbf = 0;
    int x = 0, y = 0, f = 0;
    if (valid) {
        const uint32_t *lw = reinterpret_cast<const uint32_t *>(p.levels + gb * (N * N)) + uw;
#pragma unroll
        for (int v = 0; v < N; ++v) W[v] = __ldg(lw + v * HW);
        g.template locate<PA>(gb, LOG2, x, y, f);
        const int add = 1 << (p.q.iq_shift - 1);
#pragma unroll
        for (int v = 0; v < N; ++v) {
            cbf |= W[v];
            W[v] = pack_sat_s16((s16lo(W[v]) * p.q.iq_scale + add) >> p.q.iq_shift, (s16hi(W[v]) * p.q.iq_scale + add) >> p.q.iq_shift);
        }
    }
#pragma unroll
    for (int o = 1; o < HW; o <<= 1) cbf |= __shfl_xor_sync(0xffffffffu, cbf, o);
    if (valid && uw == 0 && p.cbf) p.cbf[gb] = cbf_fold(cbf);
    big_inv_core<LOG2, PA>(tmp_all[warp], b, uw, valid, W, p.rec + f * p.fs_rec + (ptrdiff_t)y * p.s_rec + x, p.s_rec,
                           p.pred + f * p.fs_pred + (ptrdiff_t)y * p.s_pred + x, p.s_pred);
}

}  // namespace hv

// ================================================================================================ C ABI
using namespace hv;

static bool aligned16(const void *a, ptrdiff_t s1_bytes, ptrdiff_t s2_bytes = 0, const void *b = nullptr, ptrdiff_t s3_bytes = 0, ptrdiff_t s4_bytes = 0)
{
    return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)s1_bytes | (uintptr_t)s2_bytes | (uintptr_t)s3_bytes | (uintptr_t)s4_bytes) & 15) == 0;
}

template <bool PA>
static int launch_fwd_t(int16_t *coeffs, const int16_t *res, ptrdiff_t stride, ptrdiff_t fs, int log2, int trType, const BlockGrid &g, void *stream)
{
    const unsigned small_grid = (unsigned)((g.n + SMALL_NT - 1) / SMALL_NT);
    if (log2 == 2 && trType) return launch(small_fwd_kernel<2, true, PA>, small_grid, SMALL_NT, 0, stream, coeffs, res, stride, fs, g);
    if (log2 == 2) return launch(small_fwd_kernel<2, false, PA>, small_grid, SMALL_NT, 0, stream, coeffs, res, stride, fs, g);
    if (log2 == 3) return launch(small_fwd_kernel<3, false, PA>, small_grid, SMALL_NT, 0, stream, coeffs, res, stride, fs, g);
    if (log2 == 4) return launch(big_fwd_kernel<4, PA>, (unsigned)((g.n + 15) / 16), BIG_NT, 0, stream, coeffs, res, stride, fs, g);
    return launch(big_fwd_kernel<5, PA>, (unsigned)((g.n + 7) / 8), BIG_NT, 0, stream, coeffs, res, stride, fs, g);
}

// forward 16x16 / 32x32 with the first stage on the tensor cores (transform_fwd_umma.cuh): regular grids over 16-byte aligned planes.
// Default: batches of at least 6 tiles (128 x 128 samples) per SM - measured crossover on 4K planes, us per launch, tensor / butterfly:
// 17.1 / 15.0 (1 plane), 24.3 / 25.2 (2), 43.7 / 48.3 (4), 70.5 / 88.8 (8).  HEVCASM_FWD_PATH=umma: whenever the planes allow it;
// =umma_only: fail instead of falling back (tests); =butterfly: never.
template <int LOG2>
static int launch_fwd_umma(int16_t *coeffs, const int16_t *res, ptrdiff_t stride, ptrdiff_t fs, const BlockGrid &g, void *stream, bool forced, bool *taken,
                           const int *quant = nullptr /* {scale, shift, off, offn}: write levels instead of coefficients */)
{
    *taken = false;
    constexpr int BS = 1 << LOG2, TB = 128 / BS;
    const long long per = (long long)g.nbx * g.nby;
    if (per <= 0 || g.n % per != 0) return 0;
    const int n_frames = (int)(g.n / per);
    if (!tma::describable(stride * 2, fs * 2, n_frames)) return 0;
    ft::Params P{};
    P.coeffs = coeffs, P.nbx = g.nbx, P.nby = g.nby;
    P.tiles_x = (g.nbx + TB - 1) / TB, P.tiles_y = (g.nby + TB - 1) / TB;
    const long long tiles = (long long)P.tiles_x * P.tiles_y * n_frames;
    if (tiles >= (1ll << 30) || (!forced && tiles < 6ll * sm_count())) return 0;
    P.n_tiles = (int)tiles;
    if (tma::describe_u8_swizzled(&P.tmres, reinterpret_cast<const uint8_t *>(res), stride * 2, fs * 2, 2ll * BS * g.nbx, (long long)BS * g.nby, n_frames, 128, ft::TROWS))
        return 0;
    if (ft::ft_tables_init() || set_max_smem(ft::fwd_umma_kernel<LOG2, false>, ft::SMEM_BYTES) || set_max_smem(ft::fwd_umma_kernel<LOG2, true>, ft::SMEM_BYTES))
        return (int)cudaErrorInvalidValue;
    *taken = true;
    long long grid = std::min<long long>(tiles, (long long)sm_count());   // one persistent CTA per SM
    if (const char *e = tune::knob("HEVCASM_FWD_UMMA_GRID")) grid = std::max(1ll, std::min<long long>(grid, atoll(e)));   // test knob: more tiles per CTA
    if (quant) {
        P.q_scale = quant[0], P.q_shift = quant[1], P.q_off = quant[2], P.q_offn = quant[3];
        return launch(ft::fwd_umma_kernel<LOG2, true>, dim3((unsigned)grid), dim3(ft::THREADS), (size_t)ft::SMEM_BYTES, stream, P);
    }
    return launch(ft::fwd_umma_kernel<LOG2, false>, dim3((unsigned)grid), dim3(ft::THREADS), (size_t)ft::SMEM_BYTES, stream, P);
}

static int launch_fwd(int16_t *coeffs, const int16_t *res, ptrdiff_t stride, ptrdiff_t fs, int log2, int trType, const BlockGrid &g, void *stream)
{
    if (g.n == 0) return 0;
    if (((uintptr_t)coeffs & 15) != 0) return HEVCASM_ERR_ARGUMENT;
    const bool pa = !g.blk_xy && aligned16(res, stride * 2, fs * 2);
    const char *pin = tune::knob("HEVCASM_FWD_PATH");
    const bool forced = pin && !strncmp(pin, "umma", 4);
    if ((log2 == 5 || log2 == 4) && !(pin && !strcmp(pin, "butterfly"))) {
        bool taken = false;
        const int e = !pa ? 0 : log2 == 5 ? launch_fwd_umma<5>(coeffs, res, stride, fs, g, stream, forced, &taken) : launch_fwd_umma<4>(coeffs, res, stride, fs, g, stream, forced, &taken);
        if (taken) return e;
        if (pin && !strcmp(pin, "umma_only")) return HEVCASM_ERR_ARGUMENT;
    }
    return pa ? launch_fwd_t<true>(coeffs, res, stride, fs, log2, trType, g, stream) : launch_fwd_t<false>(coeffs, res, stride, fs, log2, trType, g, stream);
}

template <bool PA>
static int launch_inv_t(uint8_t *dst, ptrdiff_t sd, const uint8_t *pred, ptrdiff_t sp, ptrdiff_t fs_dst, ptrdiff_t fs_pred, const int16_t *coeffs, int log2,
                        int trType, const BlockGrid &g, void *stream)
{
    const unsigned small_grid = (unsigned)((g.n + SMALL_NT - 1) / SMALL_NT);
    if (log2 == 2 && trType) return launch(small_inv_kernel<2, true, PA>, small_grid, SMALL_NT, 0, stream, dst, sd, pred, sp, fs_dst, fs_pred, coeffs, g);
    if (log2 == 2) return launch(small_inv_kernel<2, false, PA>, small_grid, SMALL_NT, 0, stream, dst, sd, pred, sp, fs_dst, fs_pred, coeffs, g);
    if (log2 == 3) return launch(small_inv_kernel<3, false, PA>, small_grid, SMALL_NT, 0, stream, dst, sd, pred, sp, fs_dst, fs_pred, coeffs, g);
#ifdef HEVCASM_EXPERIMENTS
    // The tensor-core formulation is exact and tested, but NOT adopted: on B200 legacy mma.sync s8 issues at 512 MAC/clk/SM and
    // the kernel runs 178 us vs 145 us (32x32) / 200 vs 119 us (16x16) for the butterfly (profiles/r01_transforms.md).
    // HEVCASM_INV_PATH=imma selects it for A/B profiling.
    const char *pin = tune::knob("HEVCASM_INV_PATH");
    if (pin && !strcmp(pin, "imma") && imma_tables_init() == 0) {
        const unsigned blocks = (unsigned)std::min<long long>((g.n + IMMA_NT / 32 - 1) / (IMMA_NT / 32), 148 * 8);  // persistent warps, 8 CTAs per SM
        if (log2 == 4) return launch(imma_inv_kernel<4, PA>, blocks, IMMA_NT, 0, stream, dst, sd, pred, sp, fs_dst, fs_pred, coeffs, g);
        return launch(imma_inv_kernel<5, PA>, blocks, IMMA_NT, 0, stream, dst, sd, pred, sp, fs_dst, fs_pred, coeffs, g);
    }
    // tcgen05 (kind::i8, TMEM accumulators) formulation: HEVCASM_INV_PATH=umma (transform_umma.cuh)
    if (pin && !strcmp(pin, "umma")) {
        const long long groups = log2 == 4 ? (g.n + 7) / 8 : (g.n + 3) / 4;
        const unsigned blocks = (unsigned)std::min<long long>(groups, (long long)sm_count() * 6);   // persistent CTAs, 6 per SM
        if (log2 == 4) return launch(umma_inv_kernel<4, PA>, blocks, UMMA_NT, 0, stream, dst, sd, pred, sp, fs_dst, fs_pred, coeffs, g);
        return launch(umma_inv_kernel<5, PA>, blocks, UMMA_NT, 0, stream, dst, sd, pred, sp, fs_dst, fs_pred, coeffs, g);
    }
#endif
    if (log2 == 4) {
        // two block groups per warp, both groups' coefficients requested up front: 115 us per 16 4K frames against 119.5 with one (three: 119, four:
        // 127; 32x32 with two: 146 against 129 - 128 registers)
#ifdef HEVCASM_EXPERIMENTS
        if (const char *k = tune::knob("HEVCASM_INV16_GROUPS")) {
            const int groups = atoi(k);
            if (groups == 1) return launch(big_inv_kernel<4, PA, 1>, (unsigned)((g.n + 15) / 16), BIG_NT, 0, stream, dst, sd, pred, sp, fs_dst, fs_pred, coeffs, g);
            if (groups == 3) return launch(big_inv_kernel<4, PA, 3>, (unsigned)((g.n + 47) / 48), BIG_NT, 0, stream, dst, sd, pred, sp, fs_dst, fs_pred, coeffs, g);
            if (groups == 4) return launch(big_inv_kernel<4, PA, 4>, (unsigned)((g.n + 63) / 64), BIG_NT, 0, stream, dst, sd, pred, sp, fs_dst, fs_pred, coeffs, g);
        }
#endif
        return launch(big_inv_kernel<4, PA, 2>, (unsigned)((g.n + 31) / 32), BIG_NT, 0, stream, dst, sd, pred, sp, fs_dst, fs_pred, coeffs, g);
    }
#ifdef HEVCASM_EXPERIMENTS
    if (tune::knob("HEVCASM_INV32_GROUPS")) return launch(big_inv_kernel<5, PA, 2>, (unsigned)((g.n + 15) / 16), BIG_NT, 0, stream, dst, sd, pred, sp, fs_dst, fs_pred, coeffs, g);
#endif
    return launch(big_inv_kernel<5, PA>, (unsigned)((g.n + 7) / 8), BIG_NT, 0, stream, dst, sd, pred, sp, fs_dst, fs_pred, coeffs, g);
}

#ifdef HEVCASM_EXPERIMENTS
// inverse 16x16 / 32x32 with the second stage on the tensor cores (transform_inv_umma.cuh): regular grids over 16-byte aligned planes
template <int LOG2>
static int launch_inv_umma(uint8_t *dst, ptrdiff_t sd, const uint8_t *pred, ptrdiff_t sp, ptrdiff_t fs_dst, ptrdiff_t fs_pred, const int16_t *coeffs,
                           const BlockGrid &g, void *stream, bool forced, bool *taken)
{
    *taken = false;
    constexpr int BS = 1 << LOG2, TB = 128 / BS;
    const long long per = (long long)g.nbx * g.nby;
    if (per <= 0 || g.n % per != 0) return 0;
    const int n_frames = (int)(g.n / per);
    fi::Params P{};
    P.dst = dst, P.pred = pred, P.coeffs = coeffs, P.sd = sd, P.sp = sp, P.fs_dst = fs_dst, P.fs_pred = fs_pred, P.nbx = g.nbx, P.nby = g.nby;
    P.tiles_x = (g.nbx + TB - 1) / TB, P.tiles_y = (g.nby + TB - 1) / TB;
    const long long tiles = (long long)P.tiles_x * P.tiles_y * n_frames;
    if (tiles >= (1ll << 30) || (!forced && tiles < 6ll * sm_count())) return 0;
    P.n_tiles = (int)tiles;
    if (!tma::describable(sp, fs_pred, n_frames) ||
        tma::describe_u8_swizzled(&P.tmpred, pred, sp, fs_pred, (long long)BS * g.nbx, (long long)BS * g.nby, n_frames, 128, fi::TROWS))
        return 0;
    if (fi::fi_tables_init() || set_max_smem(fi::inv_umma_kernel<LOG2>, fi::SMEM_BYTES)) return (int)cudaErrorInvalidValue;
    *taken = true;
    long long grid = std::min<long long>(tiles, (long long)sm_count());   // one persistent CTA per SM
    if (const char *e = tune::knob("HEVCASM_INV_UMMA_GRID")) grid = std::max(1ll, std::min<long long>(grid, atoll(e)));   // test knob: more tiles per CTA
    return launch(fi::inv_umma_kernel<LOG2>, dim3((unsigned)grid), dim3(fi::THREADS), (size_t)fi::SMEM_BYTES, stream, P);
}
#endif

static int launch_inv(uint8_t *dst, ptrdiff_t sd, const uint8_t *pred, ptrdiff_t sp, ptrdiff_t fs_dst, ptrdiff_t fs_pred, const int16_t *coeffs, int log2,
                      int trType, const BlockGrid &g, void *stream)
{
    if (g.n == 0) return 0;
    if (((uintptr_t)coeffs & 15) != 0) return HEVCASM_ERR_ARGUMENT;
    const bool pa = !g.blk_xy && aligned16(dst, sd, fs_dst, pred, sp, fs_pred);
#ifdef HEVCASM_EXPERIMENTS
    // HEVCASM_INV_PATH=hybrid: 16x16 / 32x32 with the second stage on tcgen05 whenever the planes allow it; =hybrid_only: fail instead of
    // falling back (tests)
    const char *ipin = tune::knob("HEVCASM_INV_PATH");
    if ((log2 == 5 || log2 == 4) && ipin && !strncmp(ipin, "hybrid", 6)) {
        bool taken = false;
        const int e = !pa ? 0 : log2 == 5 ? launch_inv_umma<5>(dst, sd, pred, sp, fs_dst, fs_pred, coeffs, g, stream, true, &taken)
                                          : launch_inv_umma<4>(dst, sd, pred, sp, fs_dst, fs_pred, coeffs, g, stream, true, &taken);
        if (taken) return e;
        if (!strcmp(ipin, "hybrid_only")) return HEVCASM_ERR_ARGUMENT;
    }
#endif
    return pa ? launch_inv_t<true>(dst, sd, pred, sp, fs_dst, fs_pred, coeffs, log2, trType, g, stream)
              : launch_inv_t<false>(dst, sd, pred, sp, fs_dst, fs_pred, coeffs, log2, trType, g, stream);
}

static bool tr_args_ok(int log2, int trType) { return log2 >= 2 && log2 <= 5 && (trType == 0 || (trType == 1 && log2 == 2)); }

extern "C" int hevcasm_transform_batch(int16_t *coeffs, const int16_t *residual, ptrdiff_t stride, int log2size, int trType, const int16_t *blk_xy, int n,
                                       void *stream)
{
    if (!tr_args_ok(log2size, trType) || n < 0 || (n > 0 && !blk_xy)) return HEVCASM_ERR_ARGUMENT;
    BlockGrid g{blk_xy, 0, 0, n};
    g.finish();
    return launch_fwd(coeffs, residual, stride, 0, log2size, trType, g, stream);
}

extern "C" int hevcasm_transform_frames(int16_t *coeffs, const int16_t *residual, ptrdiff_t stride, int width, int height, int log2size, int trType,
                                        int n_frames, ptrdiff_t fs, void *stream)
{
    if (!tr_args_ok(log2size, trType) || n_frames < 0 || width < 0 || height < 0) return HEVCASM_ERR_ARGUMENT;
    BlockGrid g{nullptr, width >> log2size, height >> log2size, 0};
    g.n = (long long)g.nbx * g.nby * n_frames;
    g.finish();
    return launch_fwd(coeffs, residual, stride, fs, log2size, trType, g, stream);
}

extern "C" int hevcasm_inverse_transform_add_batch(uint8_t *dst, ptrdiff_t sd, const uint8_t *pred, ptrdiff_t sp, const int16_t *coeffs, int log2size,
                                                   int trType, const int16_t *blk_xy, int n, void *stream)
{
    if (!tr_args_ok(log2size, trType) || n < 0 || (n > 0 && !blk_xy)) return HEVCASM_ERR_ARGUMENT;
    BlockGrid g{blk_xy, 0, 0, n};
    g.finish();
    return launch_inv(dst, sd, pred, sp, 0, 0, coeffs, log2size, trType, g, stream);
}

// Transform-unit lists of a batch of frames, bucketed by size class: entries (x, y, frame), first n_by_class[0] 4x4 DST blocks, then the 4x4,
// 8x8, 16x16 and 32x32 DCT blocks; the coefficient blocks lie contiguous in the same order.  One launch per class present.
static const int kClassLog2[5] = {2, 2, 3, 4, 5}, kClassType[5] = {1, 0, 0, 0, 0};
static bool tu_counts_ok(const int *n_by_class, const int16_t *tus)
{
    if (!n_by_class) return false;
    long long total = 0;
    for (int c = 0; c < 5; ++c) {
        if (n_by_class[c] < 0) return false;
        total += n_by_class[c];
    }
    return total == 0 || tus;
}

extern "C" int hevcasm_transform_list_frames(int16_t *coeffs, const int16_t *residual, ptrdiff_t stride, const int16_t *tus, const int *n_by_class, ptrdiff_t fs,
                                             void *stream)
{
    if (!tu_counts_ok(n_by_class, tus)) return HEVCASM_ERR_ARGUMENT;
    long long first = 0, coef = 0;
    for (int c = 0; c < 5; ++c) {
        const int n = n_by_class[c], log2 = kClassLog2[c];
        if (n) {
            BlockGrid g{tus + 3 * first, 0, 0, n};
            g.desc_w = 3;
            g.finish();
            const int e = launch_fwd(coeffs + coef, residual, stride, fs, log2, kClassType[c], g, stream);
            if (e) return e;
        }
        first += n, coef += (long long)n << (2 * log2);
    }
    return 0;
}

extern "C" int hevcasm_inverse_transform_add_list_frames(uint8_t *dst, ptrdiff_t sd, const uint8_t *pred, ptrdiff_t sp, const int16_t *coeffs, const int16_t *tus,
                                                         const int *n_by_class, ptrdiff_t fs_dst, ptrdiff_t fs_pred, void *stream)
{
    if (!tu_counts_ok(n_by_class, tus)) return HEVCASM_ERR_ARGUMENT;
    long long first = 0, coef = 0;
    for (int c = 0; c < 5; ++c) {
        const int n = n_by_class[c], log2 = kClassLog2[c];
        if (n) {
            BlockGrid g{tus + 3 * first, 0, 0, n};
            g.desc_w = 3;
            g.finish();
            const int e = launch_inv(dst, sd, pred, sp, fs_dst, fs_pred, coeffs + coef, log2, kClassType[c], g, stream);
            if (e) return e;
        }
        first += n, coef += (long long)n << (2 * log2);
    }
    return 0;
}

extern "C" int hevcasm_inverse_transform_add_frames(uint8_t *dst, ptrdiff_t sd, const uint8_t *pred, ptrdiff_t sp, const int16_t *coeffs, int width,
                                                    int height, int log2size, int trType, int n_frames, ptrdiff_t fs_dst, ptrdiff_t fs_pred, void *stream)
{
    if (!tr_args_ok(log2size, trType) || n_frames < 0 || width < 0 || height < 0) return HEVCASM_ERR_ARGUMENT;
    BlockGrid g{nullptr, width >> log2size, height >> log2size, 0};
    g.n = (long long)g.nbx * g.nby * n_frames;
    g.finish();
    return launch_inv(dst, sd, pred, sp, fs_dst, fs_pred, coeffs, log2size, trType, g, stream);
}

template <bool PA>
static int launch_pipeline_t(const PipelineParams &p, const BlockGrid &g, int log2size, int trType, void *stream)
{
    const unsigned small_grid = (unsigned)((g.n + SMALL_NT - 1) / SMALL_NT);
    if (log2size == 2 && trType) return launch(small_pipeline_kernel<2, true, PA>, small_grid, SMALL_NT, 0, stream, p, g);
    if (log2size == 2) return launch(small_pipeline_kernel<2, false, PA>, small_grid, SMALL_NT, 0, stream, p, g);
    if (log2size == 3) return launch(small_pipeline_kernel<3, false, PA>, small_grid, SMALL_NT, 0, stream, p, g);
    if (log2size == 4) return launch(big_pipeline_kernel<4, PA>, (unsigned)((g.n + 15) / 16), BIG_NT, 0, stream, p, g);
    return launch(big_pipeline_kernel<5, PA>, (unsigned)((g.n + 7) / 8), BIG_NT, 0, stream, p, g);
}

extern "C" int hevcasm_residual_pipeline_frames(uint8_t *rec, ptrdiff_t s_rec, int16_t *levels, int32_t *cbf, const int16_t *residual, ptrdiff_t s_res,
                                                const uint8_t *pred, ptrdiff_t s_pred, int width, int height, int log2size, int trType, int q_scale,
                                                int q_shift, int q_offset, int iq_scale, int iq_shift, int n_frames, ptrdiff_t fs_rec, ptrdiff_t fs_res,
                                                ptrdiff_t fs_pred, void *stream)
{
    // quantiser domain of the reference: quantize.c:162-168 asserts; inverse shift must leave a rounding bit
    if (!tr_args_ok(log2size, trType) || n_frames < 0 || width < 0 || height < 0 || q_shift < 16 || q_shift > 27 || q_scale < 0 || q_scale >= 0x8000 ||
        q_offset < 0 || q_offset >= 0x8000 || iq_shift < 1 || iq_shift > 30 || ((uintptr_t)levels & 15))
        return HEVCASM_ERR_ARGUMENT;
    BlockGrid g{nullptr, width >> log2size, height >> log2size, 0};
    g.n = (long long)g.nbx * g.nby * n_frames;
    g.finish();
    if (g.n == 0) return 0;
    PipelineParams p{};
    p.rec = rec, p.pred = pred, p.res = residual, p.levels = levels, p.cbf = cbf;
    p.s_rec = s_rec, p.s_pred = s_pred, p.s_res = s_res, p.fs_rec = fs_rec, p.fs_pred = fs_pred, p.fs_res = fs_res;
    p.q = make_quant_params(q_scale, q_shift, q_offset, iq_scale, iq_shift);
    const bool pa = aligned16(rec, s_rec, fs_rec, pred, s_pred, fs_pred) && aligned16(residual, s_res * 2, fs_res * 2);
    // 32x32 on batches that fill the chip: two kernels - the tensor-core forward transform with the quantiser in its epilogue writes the levels,
    // the butterfly inverse dequantises them on the way in - instead of the one fused CUDA-core kernel (361 us per 16 4K frames).
    // HEVCASM_PIPE32=fused pins the fused kernel (A/B).
    if (log2size == 5 && pa && !tune::knob("HEVCASM_PIPE32")) {
        const int quant[4] = {p.q.q_scale, p.q.q_shift, p.q.q_off, p.q.q_offn};
        bool taken = false;
        const int e = launch_fwd_umma<5>(levels, residual, s_res, fs_res, g, stream, false, &taken, quant);
        if (taken) return e ? e : launch(big_dequant_inv_kernel<5, true>, (unsigned)((g.n + 7) / 8), BIG_NT, 0, stream, p, g);
    }
#ifdef HEVCASM_EXPERIMENTS
    if (log2size == 4 && pa && tune::knob("HEVCASM_PIPE16_SPLIT")) {   // the same split for 16x16 (A/B)
        const int quant[4] = {p.q.q_scale, p.q.q_shift, p.q.q_off, p.q.q_offn};
        bool taken = false;
        const int e = launch_fwd_umma<4>(levels, residual, s_res, fs_res, g, stream, false, &taken, quant);
        if (taken) return e ? e : launch(big_dequant_inv_kernel<4, true>, (unsigned)((g.n + 15) / 16), BIG_NT, 0, stream, p, g);
    }
#endif
    return pa ? launch_pipeline_t<true>(p, g, log2size, trType, stream) : launch_pipeline_t<false>(p, g, log2size, trType, stream);
}

// ------------------------------------------------------------------------------------------------ residual formed on the fly (src - pred)

extern "C" int hevcasm_transform_from_planes_frames(int16_t *coeffs, const uint8_t *src, ptrdiff_t s_src, const uint8_t *pred, ptrdiff_t s_pred, int width,
                                                    int height, int log2size, int trType, int n_frames, ptrdiff_t fs_src, ptrdiff_t fs_pred, void *stream)
{
    // 4x4 (DCT, DST) and 8x8: the block sizes whose first stage runs on bytes.  (f265_lbd_dct_8_avx2 is the 8x8 case.)
    if (!tr_args_ok(log2size, trType) || log2size > 3 || n_frames < 0 || width < 0 || height < 0 || ((uintptr_t)coeffs & 15)) return HEVCASM_ERR_ARGUMENT;
    BlockGrid g{nullptr, width >> log2size, height >> log2size, 0};
    g.n = (long long)g.nbx * g.nby * n_frames;
    g.finish();
    if (g.n == 0) return 0;
    const unsigned grid = (unsigned)((g.n + SMALL_NT - 1) / SMALL_NT);
    const bool pa = aligned16(src, s_src, fs_src, pred, s_pred, fs_pred);
#define HV_FP(L_, D_)                                                                                                                              \
    return pa ? launch(small_fwd_planes_kernel<L_, D_, true>, grid, SMALL_NT, 0, stream, coeffs, src, s_src, fs_src, pred, s_pred, fs_pred, g)      \
              : launch(small_fwd_planes_kernel<L_, D_, false>, grid, SMALL_NT, 0, stream, coeffs, src, s_src, fs_src, pred, s_pred, fs_pred, g)
    if (log2size == 2 && trType) HV_FP(2, true);
    if (log2size == 2) HV_FP(2, false);
    HV_FP(3, false);
#undef HV_FP
}

extern "C" int hevcasm_residual_from_planes_pipeline_frames(uint8_t *rec, ptrdiff_t s_rec, int16_t *levels, int32_t *cbf, const uint8_t *src, ptrdiff_t s_src,
                                                            const uint8_t *pred, ptrdiff_t s_pred, int width, int height, int log2size, int trType, int q_scale,
                                                            int q_shift, int q_offset, int iq_scale, int iq_shift, int n_frames, ptrdiff_t fs_rec, ptrdiff_t fs_src,
                                                            ptrdiff_t fs_pred, void *stream)
{
    if (!tr_args_ok(log2size, trType) || log2size > 3 || n_frames < 0 || width < 0 || height < 0 || q_shift < 16 || q_shift > 27 || q_scale < 0 ||
        q_scale >= 0x8000 || q_offset < 0 || q_offset >= 0x8000 || iq_shift < 1 || iq_shift > 30 || ((uintptr_t)levels & 15))
        return HEVCASM_ERR_ARGUMENT;
    BlockGrid g{nullptr, width >> log2size, height >> log2size, 0};
    g.n = (long long)g.nbx * g.nby * n_frames;
    g.finish();
    if (g.n == 0) return 0;
    PipelineParams p{};
    p.rec = rec, p.pred = pred, p.src = src, p.res = nullptr, p.levels = levels, p.cbf = cbf;
    p.s_rec = s_rec, p.s_pred = s_pred, p.s_src = s_src, p.s_res = 0, p.fs_rec = fs_rec, p.fs_pred = fs_pred, p.fs_src = fs_src, p.fs_res = 0;
    p.q = make_quant_params(q_scale, q_shift, q_offset, iq_scale, iq_shift);
    const bool pa = aligned16(rec, s_rec, fs_rec, pred, s_pred, fs_pred) && aligned16(src, s_src, fs_src);
    const unsigned grid = (unsigned)((g.n + SMALL_NT - 1) / SMALL_NT);
#define HV_PP(L_, D_)                                                                                       \
    return pa ? launch(small_pipeline_kernel<L_, D_, true, true>, grid, SMALL_NT, 0, stream, p, g)           \
              : launch(small_pipeline_kernel<L_, D_, false, true>, grid, SMALL_NT, 0, stream, p, g)
    if (log2size == 2 && trType) HV_PP(2, true);
    if (log2size == 2) HV_PP(2, false);
    HV_PP(3, false);
#undef HV_PP
}
