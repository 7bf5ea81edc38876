// hevcasm_b200 - inverse 16x16 / 32x32 transform + add with its SECOND stage on the 5th-generation tensor cores.
// (included by transform.cu inside namespace hv, after InvBfly / recon_word / BlockGrid; tcgen05 wrappers in umma.cuh)
//
// The first stage of the inverse (residual_decode.c:69-347, shift 7, clip to int16) contracts over block ROWS, so the bytes of a
// coefficient lie along a non-contracted dimension and the raw-int16-tile trick of the forward transform
// (transform_fwd_umma.cuh) does not apply to it.  The second stage contracts over u:
//     res[y][x] = (sum_u T[u][x] * tmp[u][y] + 2048) >> 12
// so stage 1 stays in the threads (one coefficient column per thread, IDP.2A butterfly in registers) and each thread writes its
// clipped int16 results as tmp[y][u] into shared memory, row by row, in exactly the 128-byte-swizzled K-major pattern a TMA box
// would have produced.  That tile (128 rows x 256 bytes = 4 x 4 blocks of 32 or 8 x 8 of 16) is the A operand of
//     D_p[m = tile row][n = tile column (block column bc, x)] = sum_kk A[m][kk] * B_p[n][kk],   kk = byte of the tile row
//     B_p[(bc, x)][2 BS bc + 2 u + p] = T[u][x]   (s8, K-major, host-built)
// read once as u8 (low bytes, p = 0) and once as s8 (high bytes, p = 1): 16 MMAs of M = N = 128 per tile.  TMEM lane = tile
// row, column = tile column, so a thread ends up with 32 ADJACENT output samples of one row: lo + 256 hi, rounding shift,
// predictor add (IDP.4A), cvt.pack.sat, two 16-byte stores.  The consumers fill the operand of tile i+1 before they read the
// accumulator of tile i, so the MMAs of one tile run during the CUDA-core work of its neighbours.
#pragma once

namespace fi {

constexpr int TROWS = 128, TBYTES = 256;       // tile: 128 rows x 128 samples (256 bytes of int16 per row) = MMA M x K
constexpr int C_BYTES = 128 * TBYTES;          // one constant operand (lo or hi): [chunk (16)][n (128)][16]
constexpr int HALF_BYTES = 128 * TROWS;        // one 128-byte-wide half of the data operand: 16 groups of 8 rows x 1024 bytes
constexpr int STAGE_BYTES = 2 * HALF_BYTES;
constexpr int CO_BYTES = TROWS * TBYTES;       // the tile's coefficient blocks (block rows one after the other, as in memory)
constexpr int PR_BYTES = TROWS * 128;          // the tile's predictor rows (128 bytes x 128 rows, 128-byte swizzle)
constexpr int A_OFF = 2 * C_BYTES, CO_OFF = A_OFF + 2 * STAGE_BYTES, PR_OFF = CO_OFF + 2 * CO_BYTES, BAR_OFF = PR_OFF + 2 * PR_BYTES;
constexpr int SMEM_BYTES = 1024 + BAR_OFF + 96;
constexpr int CONSUMERS = 512, THREADS = CONSUMERS + 64;   // + the MMA warp and the loader warp

// the constant operands in their shared-memory layout [size (16, 32)][lo / hi][chunk (16)][n (128)][16], built on the host once
__device__ uint4 g_fi_B[2][2 * C_BYTES / 16];
static int fi_tables_init()
{
    // one copy per device of this process (the symbol lives in each device's module image)
    static std::mutex mu;
    static bool ready[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(mu);
    if (ready[dev]) return 0;
    const int done = [] {
        static uint8_t a[2][2 * C_BYTES];
        memset(a, 0, sizeof a);
        for (int sz = 0; sz < 2; ++sz) {
            const int BS = 16 << sz;
            for (int p = 0; p < 2; ++p)
                for (int n = 0; n < 128; ++n) {
                    const int bc = n / BS, x = n % BS;
                    for (int u = 0; u < BS; ++u) {
                        const int kk = 2 * BS * bc + 2 * u + p;   // byte of the tile row that holds the low (p = 0) / high (p = 1) half of tmp[y][u] of block column bc
                        a[sz][p * C_BYTES + (kk >> 4) * (128 * 16) + n * 16 + (kk & 15)] = (uint8_t)(int8_t)dct(BS, u, x);
                    }
                }
        }
        return (int)cudaMemcpyToSymbol(g_fi_B, a, sizeof a);
    }();
    ready[dev] = done == 0;
    return done;
}

struct alignas(64) Params {
    CUtensorMap tmpred;       // predictor planes: (BS nbx bytes, BS nby rows, frames); boxes of 128 bytes x 128 rows, 128-byte swizzle
    uint8_t *dst;
    const uint8_t *pred;
    const int16_t *coeffs;
    ptrdiff_t sd, sp, fs_dst, fs_pred;
    int nbx, nby;             // blocks per plane row / column
    int tiles_x, tiles_y, n_tiles;
};

template <int LOG2>
__global__ void __launch_bounds__(THREADS, 1) inv_umma_kernel(const __grid_constant__ Params P)
{
    constexpr int BS = 1 << LOG2, TB = 128 / BS;   // block size; blocks per tile side
    extern __shared__ __align__(128) uint8_t fi_raw[];
    uint8_t *const smem = fi_raw + ((1024 - (tma::smem_u32(fi_raw) & 1023)) & 1023);
    uint8_t *const sC = smem;                    // [lo / hi][chunk][n][16]
    uint8_t *const sA = smem + A_OFF;            // [stage][half][row][128], 128-byte swizzle
    uint64_t *const filled = reinterpret_cast<uint64_t *>(smem + BAR_OFF);   // [2] the warpgroup pair in charge has written the operand stage
    uint64_t *const done = filled + 2;                                       // [2] the MMAs into the accumulator have completed
    uint64_t *const consumed = filled + 4;                                   // [2] every consumer has read the accumulator (and its predictor rows)
    uint64_t *const loaded_c = filled + 6;                                   // [2] the tile's coefficient blocks have landed
    uint64_t *const loaded_p = filled + 8;                                   // [2] the tile's predictor rows have landed
    uint32_t *const tmem_slot = reinterpret_cast<uint32_t *>(filled + 10);
    uint8_t *const sCo = smem + CO_OFF, *const sPr = smem + PR_OFF;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) tma::mbar_init(filled + i, CONSUMERS / 2 / 32), tma::mbar_init(done + i, 1), tma::mbar_init(consumed + i, CONSUMERS / 32), tma::mbar_init(loaded_c + i, 1), tma::mbar_init(loaded_p + i, 1);
    }
    if (threadIdx.x < 32) umma::tmem_alloc<512>(tmem_slot);
    for (int idx = threadIdx.x; idx < 2 * C_BYTES / 16; idx += THREADS) reinterpret_cast<uint4 *>(sC)[idx] = g_fi_B[LOG2 - 4][idx];
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tm = *tmem_slot;

    const int t0 = blockIdx.x, tstep = gridDim.x;
    const int n_mine = t0 < P.n_tiles ? (P.n_tiles - t0 + tstep - 1) / tstep : 0;
    const int per = P.tiles_x * P.tiles_y;
    auto tile_xyf = [&](int it, int &tx, int &ty, int &tf) {
        const int t = t0 + it * tstep;
        tf = t / per;
        const int r = t - tf * per;
        ty = r / P.tiles_x, tx = r - ty * P.tiles_x;
    };

    if (threadIdx.x >= CONSUMERS) {
        // ------------------------------------------------------------------------------------------------ producer (MMA issue only)
        if (threadIdx.x == CONSUMERS) {
            constexpr uint32_t ID_LO = umma::idesc_i8(false, true, false, 128), ID_HI = umma::idesc_i8(true, true, false, 128);   // A = data (u8 / s8), B = coefficients (s8)
#pragma unroll 1
            for (int q = 0; q < n_mine; ++q) {
                const int s = q & 1;
                const uint32_t ph = (q >> 1) & 1;
                if (q >= 2) tma::mbar_wait(consumed + s, ph ^ 1);   // tile q-2 has left this accumulator
                tma::mbar_wait(filled + s, ph);
                umma::fence_after();
#pragma unroll
                for (int p = 0; p < 2; ++p)
#pragma unroll
                    for (int ks = 0; ks < TBYTES / 32; ++ks) {
                        // A: swizzled K-major, groups of 8 rows 1024 bytes apart; a K-step advances the start address by 32 bytes inside the swizzle
                        // row.  B: K-major, no swizzle (LBO = distance between 16-byte k chunks, SBO = between groups of 8 rows)
                        const uint64_t da = umma::smem_desc(tma::smem_u32(sA + s * STAGE_BYTES + (ks >> 2) * HALF_BYTES) + (ks & 3) * 32, 16, 1024, 2);
                        const uint64_t db = umma::smem_desc(tma::smem_u32(sC + p * C_BYTES + ks * 2 * (128 * 16)), 128 * 16, 128);
                        umma::mma_i8(tm + s * 256 + p * 128, da, db, p ? ID_HI : ID_LO, ks);
                    }
                umma::commit(done + s);
            }
        } else if (threadIdx.x == CONSUMERS + 32) {
            // ---------------------------------------------------------------------------------------------- loader: coefficient blocks (one bulk
            // copy per block row: the tile's blocks of a row are contiguous) and predictor rows (one TMA box) of tile q into stage q & 1
            // The coefficient stage of tile q is free once the first stage of tile q-2 has run (an iteration before that tile's epilogue
            // frees the predictor stage), so the coefficients run one tile ahead of the predictor rows: neither wait holds the other up.
#pragma unroll 1
            for (int q = 0; q <= n_mine; ++q) {
                if (q < n_mine) {
                    const int s = q & 1;
                    int tx, ty, tf;
                    tile_xyf(q, tx, ty, tf);
                    const int nb = min(TB, P.nbx - tx * TB), rows = min(TB, P.nby - ty * TB);
                    if (q >= 2) tma::mbar_wait(filled + s, ((q >> 1) & 1) ^ 1);
                    tma::mbar_expect_tx(loaded_c + s, (uint32_t)(rows * nb * BS * BS * 2));
                    for (int br = 0; br < rows; ++br)
                        tma::bulk_load_1d(sCo + s * CO_BYTES + br * (TB * BS * BS * 2),
                                          P.coeffs + (((long long)tf * P.nby + ty * TB + br) * P.nbx + tx * TB) * (BS * BS), (uint32_t)(nb * BS * BS * 2), loaded_c + s);
                }
                if (q >= 1) {
                    const int t = q - 1, s = t & 1;
                    int tx, ty, tf;
                    tile_xyf(t, tx, ty, tf);
                    if (t >= 2) tma::mbar_wait(consumed + s, ((t >> 1) & 1) ^ 1);
                    tma::mbar_expect_tx(loaded_p + s, PR_BYTES);
                    tma::load_box_3d(sPr + s * PR_BYTES, &P.tmpred, tx * 128, ty * TROWS, tf, loaded_p + s);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------------------------------------ consumers
        const int wg = threadIdx.x >> 7, m = threadIdx.x & 127, warp = m >> 5, lane = m & 31;
        // stage 1 of tile `it` into operand stage it & 1.  A work item = two adjacent coefficient columns (2 uw, 2 uw + 1) of one block,
        // loaded as 32-bit words (coalesced rows), two IDP.2A butterflies, one clipped int16 pair stored per row.  A tile has
        // (128 / BS)^2 * BS / 2 = 256 (32x32) or 512 (16x16) items; the two warpgroup pairs take the tiles' first stages alternately
        // (pair it & 1 fills tile it), so each thread runs one or two items every other tile.
        auto stage1 = [&](int it) {
            if ((wg >> 1) != (it & 1)) return;
            int tx, ty, tf;
            tile_xyf(it, tx, ty, tf);
            uint8_t *const stage = sA + (it & 1) * STAGE_BYTES;
            tma::mbar_wait(loaded_c + (it & 1), (it >> 1) & 1);
            constexpr int HW = BS / 2, ITEMS = TB * TB * HW;
#pragma unroll 1
            for (int j = (wg & 1) * 128 + m; j < ITEMS; j += 256) {
                const int b = j / HW, uw = j % HW, br = b / TB, bc = b % TB;
                const int rb = ty * TB + br, bcg = tx * TB + bc;
                if (rb < P.nby && bcg < P.nbx) {
                    const uint32_t *cw = reinterpret_cast<const uint32_t *>(sCo + (it & 1) * CO_BYTES) + b * (BS * HW) + uw;   // block b of the staged tile
                    uint32_t W[BS];
#pragma unroll
                    for (int v = 0; v < BS; ++v) W[v] = cw[v * HW];
                    uint32_t p[HW];
                    int o0[BS], o1[BS];
                    static_for<0, HW>([&](auto kq) {
                        constexpr int k = HV_V(kq);
                        p[k] = lolo(W[pair_row(BS, k, 0)], W[pair_row(BS, k, 1)]);
                    });
                    InvBfly<BS>::run(p, o0, 64);
                    static_for<0, HW>([&](auto kq) {
                        constexpr int k = HV_V(kq);
                        p[k] = hihi(W[pair_row(BS, k, 0)], W[pair_row(BS, k, 1)]);
                    });
                    InvBfly<BS>::run(p, o1, 64);
                    // tile row r = BS br + y: r & 7 = y & 7, so the swizzled position of this item's word depends on y & 7 only
                    const int kk = 2 * BS * bc + 4 * uw, chunk = (kk & 127) >> 4;   // byte of the tile row
                    uint8_t *const base = stage + (kk >> 7) * HALF_BYTES + (kk & 15) + ((BS * br) >> 3) * 1024;
#pragma unroll
                    for (int y = 0; y < BS; ++y)   // the clip to int16 of residual_decode.c stage 1 = cvt.pack.sat
                        *reinterpret_cast<uint32_t *>(base + (y >> 3) * 1024 + (y & 7) * 128 + ((chunk ^ (y & 7)) << 4)) = pack_sat_s16(o0[y] >> 7, o1[y] >> 7);
                }
            }
            umma::fence_async_smem();   // the operand bytes -> visible to the tensor cores
            tma::mbar_arrive_warp(filled + (it & 1));
        };
        if (n_mine > 0) stage1(0);
#pragma unroll 1
        for (int it = 0; it < n_mine; ++it) {
            if (it + 1 < n_mine) stage1(it + 1);   // its stage was last read by the MMAs of tile it-1, whose completion this thread has observed
            const int a = it & 1;
            // stage 2 out of TMEM: tile row r = 32 warp + lane, tile columns 32 wg .. 32 wg + 31 (the predictor row is requested before the wait)
            int tx, ty, tf;
            tile_xyf(it, tx, ty, tf);
            const int r = 32 * warp + lane, yg = ty * TROWS + r, xg = tx * 128 + 32 * wg;
            const bool row_ok = yg < P.nby * BS;
            uint8_t *dp = P.dst + tf * P.fs_dst + (ptrdiff_t)yg * P.sd + xg;
            bool ok[2];
#pragma unroll
            for (int g = 0; g < 2; ++g) ok[g] = row_ok && (xg + 16 * g) / BS < P.nbx;
            tma::mbar_wait(loaded_p + a, (it >> 1) & 1);
            uint4 pw[2];   // this row's 32 predictor bytes: 16-byte chunks 2 wg, 2 wg + 1 of the swizzled row
#pragma unroll
            for (int g = 0; g < 2; ++g) pw[g] = *reinterpret_cast<const uint4 *>(sPr + a * PR_BYTES + (r >> 3) * 1024 + (r & 7) * 128 + (((2 * wg + g) ^ (r & 7)) << 4));
            tma::mbar_wait(done + a, (it >> 1) & 1);
            umma::fence_after();
            const uint32_t t = tm + ((uint32_t)(warp * 32) << 16) + a * 256 + 32 * wg;
            uint32_t ow[8];
            {
                int lo[2][8], hi[2][8];
                umma::tmem_ld8(t, lo[0]);
                umma::tmem_ld8(t + 128, hi[0]);
                umma::tmem_ld_wait(lo[0]);
                umma::tmem_ld_wait(hi[0]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < 3) {
                        umma::tmem_ld8(t + 8 * (c + 1), lo[(c + 1) & 1]);
                        umma::tmem_ld8(t + 128 + 8 * (c + 1), hi[(c + 1) & 1]);
                    }
                    int rr[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) rr[j] = lo[c & 1][j] + (hi[c & 1][j] << 8) + 2048;
                    const uint4 w = pw[c >> 1];
                    ow[2 * c] = recon_word((c & 1) ? w.z : w.x, rr);
                    ow[2 * c + 1] = recon_word((c & 1) ? w.w : w.y, rr + 4);
                    if (c < 3) {
                        umma::tmem_ld_wait(lo[(c + 1) & 1]);
                        umma::tmem_ld_wait(hi[(c + 1) & 1]);
                    }
                }
            }
            umma::fence_before();       // this thread's TMEM reads are complete
            umma::fence_async_smem();   // .. and its reads of the predictor stage precede the loader's next box
            tma::mbar_arrive_warp(consumed + a);
#pragma unroll
            for (int g = 0; g < 2; ++g)
                if (ok[g]) reinterpret_cast<uint4 *>(dp)[g] = make_uint4(ow[4 * g], ow[4 * g + 1], ow[4 * g + 2], ow[4 * g + 3]);
        }
    }
    umma::fence_before();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc<512>(*tmem_slot);
}

}  // namespace fi
