// hevcasm_b200 - Hadamard SATD (4x4, 8x8) with the HORIZONTAL Hadamard pass on the 5th-generation tensor cores.
// (included by satd.cu inside namespace hv; tcgen05 wrappers in umma.cuh)
//
// SATD = (N/4 + sum |H (A - B) H^T|) / (N/2)  (hadamard.c:75-131).  The horizontal pass of the difference is linear in the two
// planes, H (A - B) = H A + (-H) B, and both planes are bytes, so it is two int8 matrix products accumulating into ONE int32
// accumulator - exact, one TMEM read-back per sample, and no thread ever touches an input byte:
//     D[m = (block column bc, k)][n = plane row r] = sum_kk Hc[m][kk] * A[r][kk] + sum_kk (-Hc)[m][kk] * B[r][kk],   kk = byte of the tile row
//     Hc[(bc, k)][N bc + x] = H[k][x] = (-1)^popcount(k & x)   (+-1, s8, K-major, built once per CTA)
// with the two tiles (128 bytes x 256 rows each) delivered by TMA boxes with the 128-byte swizzle = the swizzled K-major operand.
// 8 MMAs (4 K-steps x 2 planes, M = 128, N = 256) per tile.  TMEM lane = (bc, k), column = plane row: a thread reads the N rows of
// a block as N consecutive columns, runs the vertical Hadamard butterfly in registers, sums the absolute values, and the N lanes of
// a block add up by shuffles.  (The order of the Hadamard outputs does not matter for the sum of absolute values.)
// Producer / consumer structure as in pred_umma.cuh (uv): one producer warp, two accumulators, two stages, mbarrier hand-offs.
#pragma once

namespace su {

constexpr int TCOLS = 128, TROWS = 256;        // tile: 128 samples x 256 rows = MMA K x N
constexpr int H_BYTES = 128 * TCOLS;           // one constant operand (+H or -H): [chunk (8)][m (128)][16]
constexpr int BOX_BYTES = 128 * TROWS;         // one plane's tile
constexpr int STAGE_BYTES = 2 * BOX_BYTES;     // both planes
constexpr int B_OFF = 2 * H_BYTES, BAR_OFF = B_OFF + 2 * STAGE_BYTES;
constexpr int SMEM_BYTES = 1024 + BAR_OFF + 64;
constexpr int CONSUMERS = 512, THREADS = CONSUMERS + 32;

struct alignas(64) Params {
    CUtensorMap tm[2];        // the two planes as bytes: (N nbx, N nby, frames); boxes of 128 bytes x 256 rows, 128-byte swizzle
    int32_t *out;             // satd[frame][by][bx]
    int nbx, nby;
    int tiles_x, tiles_y, n_tiles;
};

template <int LOG2>
__global__ void __launch_bounds__(THREADS, 1) satd_umma_kernel(const __grid_constant__ Params P)
{
    constexpr int N = 1 << LOG2, BPR = TCOLS / N, RPW = TROWS / 4, BPT = RPW / N;   // block size; blocks per tile row; rows per warpgroup; blocks per thread and tile
    extern __shared__ __align__(128) uint8_t su_raw[];
    uint8_t *const smem = su_raw + ((1024 - (tma::smem_u32(su_raw) & 1023)) & 1023);
    uint8_t *const sH = smem;                    // [+H / -H][chunk][m][16]
    uint8_t *const sB = smem + B_OFF;            // [stage][plane][row][128]
    uint64_t *const full = reinterpret_cast<uint64_t *>(smem + BAR_OFF);   // [2] both tiles of the stage have landed
    uint64_t *const done = full + 2;                                       // [2] the MMAs into the accumulator have completed
    uint64_t *const consumed = full + 4;                                   // [2] every consumer has read the accumulator
    uint32_t *const tmem_slot = reinterpret_cast<uint32_t *>(full + 6);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) tma::mbar_init(full + i, 1), tma::mbar_init(done + i, 1), tma::mbar_init(consumed + i, CONSUMERS / 32);
    }
    if (threadIdx.x < 32) umma::tmem_alloc<512>(tmem_slot);
    __syncthreads();

    const int t0 = blockIdx.x, tstep = gridDim.x;
    const int n_mine = t0 < P.n_tiles ? (P.n_tiles - t0 + tstep - 1) / tstep : 0;
    const int per = P.tiles_x * P.tiles_y;
    auto tile_xyf = [&](int it, int &tx, int &ty, int &tf) {
        const int t = t0 + it * tstep;
        tf = t / per;
        const int r = t - tf * per;
        ty = r / P.tiles_x, tx = r - ty * P.tiles_x;
    };
    auto request = [&](int q) {   // producer: both planes' tiles of tile q into stage q & 1
        int tx, ty, tf;
        tile_xyf(q, tx, ty, tf);
        const int s = q & 1;
        tma::mbar_expect_tx(full + s, STAGE_BYTES);
#pragma unroll
        for (int p = 0; p < 2; ++p) tma::load_box_3d(sB + s * STAGE_BYTES + p * BOX_BYTES, &P.tm[p], tx * (TCOLS / 4), ty * TROWS, tf, full + s);   // x in 32-bit words
    };
    if (threadIdx.x == CONSUMERS) {
        if (n_mine > 0) request(0);
        if (n_mine > 1) request(1);
    }

    // constant operands, one 16-byte chunk per step: Hc[(bc, k)][N bc + x] = (-1)^popcount(k & x), and its negative
    for (int idx = threadIdx.x; idx < 2 * (TCOLS / 16) * 128; idx += THREADS) {
        const int m = idx & 127, kc = (idx >> 7) & 7, p = idx >> 10, bc = m / N, k = m % N;
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const int kk = 16 * kc + b;
            if (kk / N == bc) {
                const int neg = (__popc(k & (kk % N)) & 1) ^ p;
                w[b >> 2] |= (neg ? 0xffu : 0x01u) << (8 * (b & 3));
            }
        }
        *reinterpret_cast<uint4 *>(sH + p * H_BYTES + kc * (128 * 16) + m * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tm = *tmem_slot;

    if (threadIdx.x >= CONSUMERS) {
        // ------------------------------------------------------------------------------------------------ producer
        if (threadIdx.x == CONSUMERS) {
            constexpr uint32_t IDESC = umma::idesc_i8(true, false, false, TROWS);   // A = +-1 (s8), B = plane bytes (u8), both K-major
#pragma unroll 1
            for (int q = 0; q < n_mine; ++q) {
                const int s = q & 1;
                const uint32_t ph = (q >> 1) & 1;
                if (q >= 2) tma::mbar_wait(consumed + s, ph ^ 1);   // tile q-2 has left this accumulator
                tma::mbar_wait(full + s, ph);
                umma::fence_after();
#pragma unroll
                for (int p = 0; p < 2; ++p)
#pragma unroll
                    for (int ks = 0; ks < TCOLS / 32; ++ks) {
                        // A: K-major, no swizzle (LBO = distance between 16-byte k chunks, SBO = between groups of 8 rows).  B: swizzled K-major,
                        // groups of 8 rows 1024 bytes apart; a K-step advances the start address by 32 bytes inside the swizzle row
                        const uint64_t da = umma::smem_desc(tma::smem_u32(sH + p * H_BYTES + ks * 2 * (128 * 16)), 128 * 16, 128);
                        const uint64_t db = umma::smem_desc(tma::smem_u32(sB + s * STAGE_BYTES + p * BOX_BYTES) + ks * 32, 16, 1024, 2);
                        umma::mma_i8(tm + s * TROWS, da, db, IDESC, p | ks);
                    }
                umma::commit(done + s);
                if (q + 2 < n_mine) {   // the tile after next takes this stage as soon as these MMAs have read it
                    tma::mbar_wait(done + s, ph);
                    request(q + 2);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------------------------------------ consumers
        const int wg = threadIdx.x >> 7, m = threadIdx.x & 127, warp = m >> 5;   // RPW plane rows of the tile; TMEM lane = (block column, k)
        const int bc = m / N, k = m % N;
        const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16) + RPW * wg;
#pragma unroll 1
        for (int it = 0; it < n_mine; ++it) {
            const int a = it & 1;
            int tx, ty, tf;
            tile_xyf(it, tx, ty, tf);
            tma::mbar_wait(done + a, (it >> 1) & 1);
            umma::fence_after();
            int sum[BPT];   // this lane's share of sum |T| per block
            {
                const uint32_t t = tl + a * TROWS;
                int v[2][8];
                umma::tmem_ld8(t, v[0]);
                umma::tmem_ld_wait(v[0]);
#pragma unroll
                for (int c = 0; c < RPW / 8; ++c) {
                    if (c + 1 < RPW / 8) umma::tmem_ld8(t + 8 * (c + 1), v[(c + 1) & 1]);
                    int x[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[j] = v[c & 1][j];
                    // vertical Hadamard butterflies in place: within groups of N rows
#pragma unroll
                    for (int h = 1; h < N; h <<= 1)
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (!(j & h)) {
                                const int p0 = x[j], p1 = x[j | h];
                                x[j] = p0 + p1, x[j | h] = p0 - p1;
                            }
#pragma unroll
                    for (int g = 0; g < 8 / N; ++g) {
                        int s = 0;
#pragma unroll
                        for (int j = 0; j < N; ++j) s += abs(x[g * N + j]);
                        sum[c * (8 / N) + g] = s;
                    }
                    if (c + 1 < RPW / 8) umma::tmem_ld_wait(v[(c + 1) & 1]);
                }
            }
            umma::fence_before();   // this thread's TMEM reads are complete
            tma::mbar_arrive_warp(consumed + a);
            // the N lanes (k) of a block add up; lane k = 0 writes (N/4 + sum) / (N/2)
            const int bx = tx * BPR + bc;
#pragma unroll
            for (int i = 0; i < BPT; ++i) {
                int s = sum[i];
#pragma unroll
                for (int o = 1; o < N; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const int by = ty * (TROWS / N) + wg * BPT + i;
                if (k == 0 && bx < P.nbx && by < P.nby) P.out[((long long)tf * P.nby + by) * P.nbx + bx] = (N / 4 + s) / (N / 2);
            }
        }
    }
    umma::fence_before();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc<512>(*tmem_slot);
}

}  // namespace su
