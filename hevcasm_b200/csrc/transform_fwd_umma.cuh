// hevcasm_b200 - forward 16x16 / 32x32 DCT with its FIRST stage on the 5th-generation tensor cores and the second in the threads.
// (written out for 32x32; for 16x16 a tile is 8 x 8 blocks and a thread runs two 16-point second stages per tile)
// (included by transform.cu inside namespace hv, after FwdBfly / BlockGrid; tcgen05 wrappers in umma.cuh)
//
// Stage 1 of the reference (residual_decode.c:795-852, shift 4) is tmp[k][j] = (sum_i T[k][i] * X[j][i] + 8) >> 4 per block
// row j: a matrix product on int16 data.  An int16 is lo + 256 * hi, and the tensor cores take the raw bytes of the residual
// tile unsplit: the tile row (256 bytes = 4 blocks x 32 int16) is the K dimension, and the constant operand carries T[k][i]
// on the EVEN byte positions (2i) of its own block for the low-byte product and on the ODD positions (2i+1) for the
// high-byte product, zeros everywhere else.  The same shared-memory bytes are read as a u8-typed operand for the first and
// as an s8-typed operand for the second, so the bytes of the "other" half only ever meet zero coefficients:
//     D_p[m = (block column bc, frequency k)][n = plane row r] = sum_kk A_p[m][kk] * B[kk][r],   kk = byte of the tile row
//     A_p[(bc, k)][64 bc + 2 i + p] = T[k][i]  (s8, K-major, host-built image, copied per CTA by cp.async.bulk);   B = the residual tile as TMA delivers it
// (two boxes of 128 bytes x 128 rows with the 128-byte swizzle = the swizzled K-major operand, no thread touches an input
// sample).  16 MMAs (8 K-steps x {lo, hi}) of M = 128, N = 128 per tile of 4 x 4 blocks.  TMEM lane = (bc, k), column = plane
// row, so the thread (block row = warpgroup, block column = warp, k = lane) finds tmp[k][0..31] of ITS block in 2 x 32 TMEM
// columns: it recombines lo + 256 hi, rounds, truncates to int16 exactly like the reference's store, runs the second stage
// (32-point partial butterfly over j, transform.cuh) in registers and stores coeffs[v * 32 + k] for v = 0..31 - 64
// contiguous bytes per warp and v.  Producer / consumer structure as in pred_umma.cuh (uv): one producer warp, two
// accumulators (2 x 256 TMEM columns), two residual stages, mbarrier hand-offs, no barrier among the consumers.
#pragma once

namespace ft {

constexpr int TROWS = 128;                     // plane rows per tile = MMA N (128 / BS blocks)
constexpr int TBYTES = 256;                    // bytes per tile row = MMA K (8 steps of 32) = 128 samples
constexpr int A_BYTES = 128 * TBYTES;          // one constant operand: [chunk (16)][m (128)][16]
constexpr int BOX_BYTES = 128 * TROWS;         // one box: 128 bytes x 128 rows
constexpr int STAGE_BYTES = 2 * BOX_BYTES;
constexpr int NST = 4;                         // residual stages in flight
constexpr int B_OFF = 2 * A_BYTES, BAR_OFF = B_OFF + NST * STAGE_BYTES;
constexpr int SMEM_BYTES = 1024 + BAR_OFF + 128;
constexpr int CONSUMERS = 512, THREADS = CONSUMERS + 64;   // four consumer warpgroups (32 plane rows each) + the loader warp + the MMA warp (one thread of each works)

// the two constant operands in their shared-memory layout [size (16, 32)][lo / hi][chunk (16)][m (128)][16], built on the host once
__device__ uint4 g_ft_A[2][2 * A_BYTES / 16];
static int ft_tables_init()   // internal linkage: the statics of an INLINE function are process-unique (STB_GNU_UNIQUE) and would be shared with the experiments library, whose copy of the table would then never be filled
{
    // one copy per device of this process (the symbol lives in each device's module image)
    static std::mutex mu;
    static bool ready[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(mu);
    if (ready[dev]) return 0;
    const int done = [] {
        static uint8_t a[2][2 * A_BYTES];
        memset(a, 0, sizeof a);
        for (int sz = 0; sz < 2; ++sz) {
            const int BS = 16 << sz;
            for (int p = 0; p < 2; ++p)
                for (int m = 0; m < 128; ++m) {
                    const int bc = m / BS, k = m % BS;
                    for (int i = 0; i < BS; ++i) {
                        const int kk = 2 * BS * bc + 2 * i + p;   // byte of the tile row that holds the low (p = 0) / high (p = 1) half of sample i of block column bc
                        a[sz][p * A_BYTES + (kk >> 4) * (128 * 16) + m * 16 + (kk & 15)] = (uint8_t)(int8_t)dct(BS, k, i);
                    }
                }
        }
        return (int)cudaMemcpyToSymbol(g_ft_A, a, sizeof a);
    }();
    ready[dev] = done == 0;
    return done;
}

struct alignas(64) Params {
    CUtensorMap tmres;        // residual planes as bytes: (2 BS nbx bytes, BS nby rows, frames); boxes of 128 bytes x 128 rows, 128-byte swizzle
    int16_t *coeffs;
    int nbx, nby;             // blocks per plane row / column
    int tiles_x, tiles_y, n_tiles;
    int q_scale, q_shift, q_off, q_offn;   // QUANT: the quantiser of hevcasm_quantize in the form of transform.cu's quant_dequant_word
};

// QUANT: the coefficients leave quantised (levels) - the first half of the fused residual pipeline for 32x32 blocks
template <int LOG2, bool QUANT = false>
__global__ void __launch_bounds__(THREADS, 1) fwd_umma_kernel(const __grid_constant__ Params P)
{
    constexpr int BS = 1 << LOG2, TB = 128 / BS, BPT = 32 / BS;   // block size; blocks per tile side; blocks per thread and tile (one below the other)
    constexpr int S1 = fwd_shift1(LOG2), S2 = fwd_shift2(LOG2);
    extern __shared__ __align__(128) uint8_t ft_raw[];
    uint8_t *const smem = ft_raw + ((1024 - (tma::smem_u32(ft_raw) & 1023)) & 1023);
    uint8_t *const sA = smem;                    // [lo / hi][chunk][m][16]
    uint8_t *const sB = smem + B_OFF;            // [stage][box][row][128]
    uint64_t *const full = reinterpret_cast<uint64_t *>(smem + BAR_OFF);   // [NST] the residual boxes of the stage have landed
    uint64_t *const empty = full + NST;                                    // [NST] the MMAs reading the stage have completed
    uint64_t *const done = empty + NST;                                    // [2] the MMAs into the accumulator have completed
    uint64_t *const consumed = done + 2;                                   // [2] every consumer warp has read the accumulator
    uint64_t *const cready = consumed + 2;                                 // the constant operands have landed
    uint32_t *const tmem_slot = reinterpret_cast<uint32_t *>(cready + 1);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NST; ++i) tma::mbar_init(full + i, 1), tma::mbar_init(empty + i, 1);
#pragma unroll
        for (int i = 0; i < 2; ++i) tma::mbar_init(done + i, 1), tma::mbar_init(consumed + i, CONSUMERS / 32);
        tma::mbar_init(cready, 1);
    }
    if (threadIdx.x < 32) umma::tmem_alloc<512>(tmem_slot);
    __syncthreads();

    const int t0 = blockIdx.x, tstep = gridDim.x;
    const int n_mine = t0 < P.n_tiles ? (P.n_tiles - t0 + tstep - 1) / tstep : 0;
    const int per = P.tiles_x * P.tiles_y;
    const int sf = tstep / per, sby = (tstep - sf * per) / P.tiles_x, sbx = tstep - sf * per - sby * P.tiles_x;
    int cf = t0 / per, cy = (t0 - cf * per) / P.tiles_x, cx = t0 - cf * per - cy * P.tiles_x;
    auto advance = [&](int &x, int &y, int &f) {
        x += sbx;
        if (x >= P.tiles_x) x -= P.tiles_x, ++y;
        y += sby;
        if (y >= P.tiles_y) y -= P.tiles_y, ++f;
        f += sf;
    };
    // Service work is split over two single-thread warps (as in pred_umma.cuh: one thread doing requests AND MMA issue was the critical
    // path): the loader waits `empty[stage]` and requests the residual boxes, the MMA thread waits `full[stage]` / `consumed[acc]`,
    // issues the 16 MMAs and commits `done[acc]` + `empty[stage]`.
    int lq = 0, lst = 0;   // loader: next tile to request, its stage, the parity to wait for on `empty` (a fresh barrier passes a wait for parity 1)
    uint32_t eph = 1;
    auto request = [&]() {
        tma::mbar_expect_tx(full + lst, 2 * BOX_BYTES);
        uint8_t *b = sB + lst * STAGE_BYTES;
        tma::load_box_3d(b, &P.tmres, cx * TBYTES, cy * TROWS, cf, full + lst);
        tma::load_box_3d(b + BOX_BYTES, &P.tmres, cx * TBYTES + 128, cy * TROWS, cf, full + lst);
        advance(cx, cy, cf);
        ++lq;
        if (++lst == NST) lst = 0, eph ^= 1;
    };
    if (threadIdx.x == CONSUMERS) {
        // ------------------------------------------------------------------------------------------------ loader (prologue)
        // constant operands A_p[(bc, k)][64 bc + 2 i + p] = T[k][i]: the host-built image arrives by bulk copies while the first tiles load
        tma::mbar_expect_tx(cready, 2 * A_BYTES);
#pragma unroll
        for (int i = 0; i < 4; ++i) tma::bulk_load_1d(sA + i * (A_BYTES / 2), reinterpret_cast<const uint8_t *>(g_ft_A[LOG2 - 4]) + i * (A_BYTES / 2), A_BYTES / 2, cready);
#pragma unroll 1
        while (lq < NST && lq < n_mine) request();   // the first NST tiles travel during the rest of the prologue
    }
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tm = *tmem_slot;

    if (threadIdx.x >= CONSUMERS) {
        if (threadIdx.x == CONSUMERS) {
            // ------------------------------------------------------------------------------------------------ loader
#pragma unroll 1
            while (lq < n_mine) {
                tma::mbar_wait(empty + lst, eph);
                request();
            }
        } else if (threadIdx.x == CONSUMERS + 32) {
            // ------------------------------------------------------------------------------------------------ MMA issue
            constexpr uint32_t ID_LO = umma::idesc_i8(true, false, false, TROWS), ID_HI = umma::idesc_i8(true, true, false, TROWS);
            // A: K-major, no swizzle (LBO = distance between 16-byte k chunks, SBO = between groups of 8 rows).  B: swizzled K-major,
            // groups of 8 rows 1024 bytes apart; a K-step advances the start address by 32 bytes inside the swizzle row
            const uint64_t da0 = umma::smem_desc(tma::smem_u32(sA), 128 * 16, 128), db0 = umma::smem_desc(tma::smem_u32(sB), 16, 1024, 2);
            tma::mbar_wait(cready, 0);
            int st = 0;
            uint32_t fph = 0;
#pragma unroll 1
            for (int q = 0; q < n_mine; ++q) {
                const int s = q & 1;
                tma::mbar_wait(full + st, fph);
                if (q >= 2) tma::mbar_wait(consumed + s, ((q >> 1) & 1) ^ 1);   // tile q-2 has left this accumulator
                umma::fence_after();
#pragma unroll
                for (int p = 0; p < 2; ++p)
#pragma unroll
                    for (int ks = 0; ks < TBYTES / 32; ++ks)   // (descriptor arithmetic on the 14-bit address field, 16-byte units)
                        umma::mma_i8(tm + s * 256 + p * TROWS, da0 + (uint64_t)((p * A_BYTES + ks * 2 * (128 * 16)) >> 4),
                                     db0 + (uint64_t)((st * STAGE_BYTES + (ks >> 2) * BOX_BYTES + (ks & 3) * 32) >> 4), p ? ID_HI : ID_LO, ks);
                umma::commit(done + s);
                umma::commit(empty + st);
                if (++st == NST) st = 0, fph ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------------------------------------------ consumers
        const int wg = threadIdx.x >> 7, m = threadIdx.x & 127, warp = m >> 5;   // 32 plane rows of the tile; TMEM lane = (block column, frequency k)
        const int bc = m / BS, k = m % BS;
        const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16) + 32 * wg;
#pragma unroll 1
        for (int it = 0; it < n_mine; ++it) {
            const int a = it & 1;
            tma::mbar_wait(done + a, (it >> 1) & 1);
            umma::fence_after();
            // stage 1 out of TMEM: tmp[k][j] = (lo + 256 hi + round) >> S1, truncated to int16 (residual_decode.c:846)
            int x[32];
            {
                const uint32_t t = tl + a * 256;
                int lo[2][8], hi[2][8];
                umma::tmem_ld8(t, lo[0]);
                umma::tmem_ld8(t + TROWS, hi[0]);
                umma::tmem_ld_wait(lo[0]);
                umma::tmem_ld_wait(hi[0]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < 3) {
                        umma::tmem_ld8(t + 8 * (c + 1), lo[(c + 1) & 1]);
                        umma::tmem_ld8(t + TROWS + 8 * (c + 1), hi[(c + 1) & 1]);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        x[8 * c + j] = (int)(short)((lo[c & 1][j] + (hi[c & 1][j] << 8) + (1 << (S1 - 1))) >> S1);
                    }
                    if (c < 3) {
                        umma::tmem_ld_wait(lo[(c + 1) & 1]);
                        umma::tmem_ld_wait(hi[(c + 1) & 1]);
                    }
                }
            }
            umma::fence_before();   // this thread's TMEM reads are complete
            tma::mbar_arrive_warp(consumed + a);
            // stage 2 in registers: out[v] = (sum_j T[v][j] tmp[k][j] + round) >> S2 -> coeffs[v * BS + k]  (residual_decode.c:855-892)
            const int bcg = cx * TB + bc;
#pragma unroll
            for (int h = 0; h < BPT; ++h) {
                const int rb = cy * TB + wg * BPT + h;
                if (rb < P.nby && bcg < P.nbx) {
                    int o[BS];
                    // 32-bit butterfly for any int16 input.  (The packed IDP.2A odd part needs a range test per value: measured 128 vs 116 us.)
                    FwdBfly<BS>::run(x + h * BS, o, 1 << (S2 - 1));
                    int16_t *out = P.coeffs + (((long long)cf * P.nby + rb) * P.nbx + bcg) * (BS * BS) + k;
#pragma unroll
                    for (int v = 0; v < BS; ++v) {
                        int c = (int)(int16_t)(o[v] >> S2);   // the coefficient as the reference stores it
                        if (QUANT) c = (c * P.q_scale + (c < 0 ? P.q_offn : P.q_off)) >> P.q_shift;
                        out[v * BS] = (int16_t)c;
                    }
                }
            }
            advance(cx, cy, cf);
        }
    }
    umma::fence_before();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc<512>(*tmem_slot);
}

}  // namespace ft
