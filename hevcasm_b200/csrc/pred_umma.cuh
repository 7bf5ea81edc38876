// hevcasm_b200 - two-pass interpolation positions with one pass of the separable filter on the 5th-generation tensor cores
// (tcgen05.mma kind::i8, operands staged by TMA, accumulators in TMEM).  Two kernels:
//   namespace um: HORIZONTAL pass on the tensor cores, thread = output column walking down the rows  (used for two references)
//   namespace uv: VERTICAL pass on the tensor cores, thread = output row walking along the columns     (used for one reference)
// (included by pred.cu inside namespace hv::ip, after PredParams / PackedCoefs / pack16 / dp2a_lo; measurements and the
// experiment series in profiles/r01_pred_umma.md)
//
// ---- um ----
// The streaming kernels spend 6.1 IDP + 8 other instructions per two-pass sample and are bound by the FMA pipe and by
// issue slots (profiles/r01_pred.md).  The horizontal pass is a banded matrix product, and int8 x int8 -> int32 is exactly
// what tcgen05.mma kind::i8 computes:
//     D[m][n] = sum_k A[m][k] * B[k][n],   m = output column of the tile (128), n = row of the tile (80), k = input column (160)
//     A[m][k] = tap[k - m - (16 - LEFT)]   (a Toeplitz band, s8, built once per CTA in shared memory, K-major)
//     B[k][n] = ref[y0 - TOP + n][x0 - 16 + k]   (u8: the image rows themselves, K-major as they lie in memory)
// The image operand arrives by two TMA boxes per tile - {128 bytes x 80 rows} with the 128-byte swizzle (K-steps 0..3) and
// {32 bytes x 80 rows} with the 32-byte swizzle (K-step 4) - which land as the tensor cores' swizzled K-major operand
// form, so no thread ever touches an input byte.  (A single 4-D box (16 bytes, rows, 16-byte chunks) gives the no-swizzle
// form and is just as exact, but its 800 16-byte pieces per tile throttle the TMA unit: 166 us per 16 4K planes.)  Five
// K-steps of 32 accumulate the 160 input columns.  The exact horizontal sums come back from TMEM with thread = output
// column, 8 rows per tcgen05.ld - precisely the "thread walks down its column" form the vertical pass wants: pairs of
// consecutive rows in a register ring, IDP.2A with the vertical taps, rounding shift, clip.  What remains on the CUDA
// cores per sample is 4 IDP.2A + 1 PRMT + ~4.  tools/umma_fir_probe.cu pinned the TMA stride order and the descriptors.
#pragma once

#ifdef HEVCASM_EXPERIMENTS   // the horizontal-pass-first kernel is measured-but-not-adopted (126 vs 110 us): experiments build only
namespace um {

constexpr int TCOLS = 128;                 // output columns per tile = MMA M = TMEM lanes = threads
constexpr int TROWS = 72;                  // output rows per tile
constexpr int NROWS = 80;                  // MMA N: staged rows (TROWS + TAPS - 1 <= 80, a multiple of 16)
constexpr int KBYTES = 160, KCH = KBYTES / 16;   // staged bytes per row: 16 left of the tile + 128 + 16
constexpr int A_BYTES = TCOLS * KBYTES;    // 20480: Toeplitz operand, [chunk][m][16]
constexpr int B1_BYTES = NROWS * 128, B2_BYTES = NROWS * 32;   // image operand: 128-byte-swizzled part (k < 128), 32-byte-swizzled part (k >= 128)
constexpr int B_BYTES = 13 * 1024;         // one staged operand (12800 bytes), padded so that every one starts on a 1024-byte boundary
constexpr int O_BYTES = TROWS * TCOLS;     // one output tile, row-major

struct alignas(64) Params {
    CUtensorMap tm128[2], tm32[2];   // per reference: bytes from x = -16, rows from -(TAPS/2-1), frames; boxes of 128 / 32 bytes x NROWS rows
    CUtensorMap tmdst;               // destination planes, boxes of TCOLS bytes x TROWS rows (valid when dst16)
    uint8_t *dst;
    ptrdiff_t sd, fs_dst;
    int width, height;
    int dst16;                // destination rows are 16-byte aligned (pointer and strides): tiles leave by TMA store
    int tiles_x, tiles_y, n_tiles;
    int8_t xtap[2][8];        // horizontal taps per reference
    int y2[2][4];             // vertical tap pairs per reference (PackedCoefs::y2)
};

// One CTA per SM holds NWG independent WARPGROUPS of 128 threads (6 for one reference, 3 for two; only the two-reference form is still instantiated - HEVCASM_PRED_BI=hfirst).  Each warpgroup walks over its own
// tiles with its own image stage, output buffers, barriers and accumulator columns, so the tensor-core and TMA latency of one
// overlaps the vertical passes of the others; they share the Toeplitz operand and one 512-column TMEM allocation.
template <int TAPS, bool BI>
struct Geom {
    static constexpr int NREF = BI ? 2 : 1, NWG = BI ? 3 : 6;
    static constexpr int WG_BYTES = NREF * B_BYTES + 2 * O_BYTES;  // image stage + two output buffers of one warpgroup
    static constexpr int BAR_OFF = NREF * A_BYTES + NWG * WG_BYTES;
    static constexpr int SMEM_BYTES = 1024 + BAR_OFF + NWG * 16 + 16;   // alignment slack + operands + per-warpgroup barriers + the TMEM slot
    static constexpr int ACC_COLS = BI ? 160 : 80;                 // accumulator columns per warpgroup (reference r at + 80 r)
    static constexpr int THREADS = NWG * TCOLS;
};

// The vertical pass of one tile for one thread (= one output column).  Everything about the row index is compile-time (the
// 80 staged rows are fully unrolled), the output byte goes to shared memory at an immediate offset.  The horizontal sums
// arrive 8 rows per tcgen05.ld; the load of the next 8 is in flight while these 8 are consumed.
template <int TAPS, bool BI>
__device__ __forceinline__ void vertical_pass(const Params &P, uint32_t tlane, uint8_t *ocol /* obuf + column */)
{
    constexpr int NREF = BI ? 2 : 1, CH = 8, NCH = NROWS / CH;
    static_assert(CH % TAPS == 0, "ring slots must be compile-time");
    uint32_t ring[NREF][TAPS];
    int prev[NREF];
#pragma unroll
    for (int rf = 0; rf < NREF; ++rf) {
        prev[rf] = 0;
#pragma unroll
        for (int k = 0; k < TAPS; ++k) ring[rf][k] = 0;
    }
    int y2[NREF][TAPS / 2];
#pragma unroll
    for (int rf = 0; rf < NREF; ++rf)
#pragma unroll
        for (int g = 0; g < TAPS / 2; ++g) y2[rf][g] = P.y2[rf][g];
    int v[2][NREF][CH];
#pragma unroll
    for (int rf = 0; rf < NREF; ++rf) umma::tmem_ld8(tlane + NROWS * rf, v[0][rf]);
#pragma unroll
    for (int rf = 0; rf < NREF; ++rf) umma::tmem_ld_wait(v[0][rf]);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        if (c + 1 < NCH) {
#pragma unroll
            for (int rf = 0; rf < NREF; ++rf) umma::tmem_ld8(tlane + NROWS * rf + CH * (c + 1), v[(c + 1) & 1][rf]);
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int r = CH * c + j, y = r - (TAPS - 1);   // staged row, output row
#pragma unroll
            for (int rf = 0; rf < NREF; ++rf) {
                ring[rf][r % TAPS] = pack16(prev[rf], v[c & 1][rf][j]);   // pair (r-1, r)
                prev[rf] = v[c & 1][rf][j];
            }
            if (y < 0 || y >= TROWS) continue;
            // output row y takes the pairs ending at rows y+1, y+3, .. = slots (r + 2 + 2g) mod TAPS
            int acc[NREF];
#pragma unroll
            for (int rf = 0; rf < NREF; ++rf) {
                int a = BI ? 0 : 2048;
#pragma unroll
                for (int g = 0; g < TAPS / 2; ++g) a = dp2a_lo(ring[rf][(r + 2 + 2 * g) % TAPS], y2[rf][g], a);
                acc[rf] = a;
            }
            int o;
            if (BI) o = ((int)(short)(acc[0] >> 6) + (int)(short)(acc[NREF - 1] >> 6) + 64) >> 7;   // int16 wrap as in the reference's C
            else o = acc[0] >> 12;
            ocol[y * TCOLS] = (uint8_t)__vimin_s32_relu(o, 255);   // clip to [0, 255]: one VIMNMX
        }
        if (c + 1 < NCH) {
#pragma unroll
            for (int rf = 0; rf < NREF; ++rf) umma::tmem_ld_wait(v[(c + 1) & 1][rf]);
        }
    }
}

template <int TAPS, bool BI>
__global__ void __launch_bounds__(Geom<TAPS, BI>::THREADS, 1) pred_umma_kernel(const __grid_constant__ Params P)
{
    using G = Geom<TAPS, BI>;
    constexpr int NREF = G::NREF, NWG = G::NWG, LEFT = TAPS / 2 - 1;
    static_assert(TROWS + TAPS - 1 <= NROWS && NWG * G::ACC_COLS <= 512, "tile rows / TMEM columns");
    extern __shared__ __align__(128) uint8_t us_raw[];
    uint8_t *const us_smem = us_raw + ((1024 - (tma::smem_u32(us_raw) & 1023)) & 1023);   // the 128-byte swizzle atoms sit on 1024-byte boundaries
    const int wg = threadIdx.x / TCOLS, tid = threadIdx.x - wg * TCOLS, warp = tid >> 5;   // warpgroup, thread and warp inside it
    uint8_t *const sA = us_smem;
    uint8_t *const sB = us_smem + NREF * A_BYTES + wg * G::WG_BYTES;    // image stage (NREF operands), then the two output buffers
    uint8_t *const sO = sB + NREF * B_BYTES;
    uint64_t *const full = reinterpret_cast<uint64_t *>(us_smem + G::BAR_OFF + wg * 16);   // image boxes of this warpgroup have landed
    uint64_t *const done = full + 1;                                                       // MMA completion of this warpgroup
    uint32_t *const tmem_slot = reinterpret_cast<uint32_t *>(us_smem + G::BAR_OFF + NWG * 16);
    auto wg_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + wg), "n"(TCOLS) : "memory"); };

    // Toeplitz bands, one 16-byte chunk per step: output column m reads staged bytes m + (16 - LEFT) .. + TAPS - 1
    for (int i = threadIdx.x; i < NREF * KCH * TCOLS; i += G::THREADS) {
        const int m = i % TCOLS, kc = (i / TCOLS) % KCH, rf = i / (TCOLS * KCH);
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const int t = 16 * kc + b - m - (16 - LEFT);
            if (t >= 0 && t < TAPS) w[b >> 2] |= (uint32_t)(uint8_t)P.xtap[rf][t] << (8 * (b & 3));
        }
        *reinterpret_cast<uint4 *>(sA + rf * A_BYTES + kc * (TCOLS * 16) + m * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (tid == 0) {
        tma::mbar_init(full, 1);
        tma::mbar_init(done, 1);
    }
    if (threadIdx.x < 32) umma::tmem_alloc<512>(tmem_slot);
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tm = *tmem_slot + wg * G::ACC_COLS, tlane = tm + ((uint32_t)(warp * 32) << 16);
    constexpr uint32_t IDESC = umma::idesc_i8(true, false, false, NROWS);   // A = taps (s8), B = image (u8), both K-major

    // tile -> (column, row, frame) of tiles: one division at the start, additions with carries afterwards
    const int t0 = blockIdx.x * NWG + wg, tstep = gridDim.x * NWG;
    const int per = P.tiles_x * P.tiles_y;
    const int sf = tstep / per, sby = (tstep - sf * per) / P.tiles_x, sbx = tstep - sf * per - sby * P.tiles_x;
    int cf = t0 / per, cy = (t0 - cf * per) / P.tiles_x, cx = t0 - cf * per - cy * P.tiles_x;   // this tile
    auto advance = [&](int &x, int &y, int &f) {
        x += sbx;
        if (x >= P.tiles_x) x -= P.tiles_x, ++y;
        y += sby;
        if (y >= P.tiles_y) y -= P.tiles_y, ++f;
        f += sf;
    };
    auto request = [&](int x, int y, int f) {   // thread 0 of the warpgroup: the image boxes of tile (x, y, f) into the stage
        tma::mbar_expect_tx(full, NREF * (B1_BYTES + B2_BYTES));
#pragma unroll
        for (int rf = 0; rf < NREF; ++rf) {
            uint8_t *b = sB + rf * B_BYTES;
            tma::load_box_3d(b, &P.tm128[rf], x * TCOLS, y * TROWS, f, full);
            tma::load_box_3d(b + B1_BYTES, &P.tm32[rf], x * TCOLS + 128, y * TROWS, f, full);
        }
    };
    if (tid == 0 && t0 < P.n_tiles) request(cx, cy, cf);
    // a TMA store clips rows exactly but columns only at 16-byte granularity (measured: a 200-byte-wide plane was written up to
    // byte 207), so a partial right-hand tile of a plane whose width is not a multiple of 16 leaves by byte stores instead
    const bool tma_all = P.dst16 && (P.width & 15) == 0;

    int it = 0;
#pragma unroll 1
    for (int t = t0; t < P.n_tiles; t += tstep, ++it) {
        uint8_t *const obuf = sO + (it & 1) * O_BYTES;
        if (tid == 0) {
            tma::mbar_wait(full, it & 1);
            umma::fence_after();
#pragma unroll
            for (int rf = 0; rf < NREF; ++rf)
#pragma unroll
                for (int ks = 0; ks < KBYTES / 32; ++ks) {
                    // A: K-major, no swizzle (LBO = distance between 16-byte k chunks, SBO = between groups of 8 rows).  B: swizzled K-major,
                    // groups of 8 rows 1024 (256) bytes apart; a K-step advances the start address by 32 bytes inside the swizzle row
                    const uint64_t da = umma::smem_desc(tma::smem_u32(sA + rf * A_BYTES + ks * 2 * (TCOLS * 16)), TCOLS * 16, 128);
                    const uint32_t bb = tma::smem_u32(sB + rf * B_BYTES);
                    const uint64_t db = ks < 4 ? umma::smem_desc(bb + ks * 32, 16, 1024, 2) : umma::smem_desc(bb + B1_BYTES, 16, 256, 6);
                    umma::mma_i8(tm + NROWS * rf, da, db, IDESC, ks);
                }
            // the store of the tile before last has finished reading the output buffer this tile is about to fill; everybody
            // learns that through `done`
            tma::store_wait_read<1>();
            umma::commit(done);
        }
        tma::mbar_wait(done, it & 1);
        umma::fence_after();
        // the MMAs have consumed the image stage: the next tile's rows travel during this tile's vertical pass
        if (tid == 0 && t + tstep < P.n_tiles) {
            int x = cx, y = cy, f = cf;
            advance(x, y, f);
            request(x, y, f);
        }

        // ---- vertical pass: this thread owns output column x = cx * 128 + tid.  Output bytes go to a row-major 128-byte-pitch
        // buffer, which leaves as one TMA store (clipped to the plane by the hardware)
        vertical_pass<TAPS, BI>(P, tlane, obuf + tid);
        umma::fence_async_smem();   // the output bytes -> visible to the TMA store
        umma::fence_before();       // this tile's TMEM reads are complete before the next tile's MMAs overwrite the accumulator
        wg_sync();
        const int x0 = cx * TCOLS, y0 = cy * TROWS;
        if (tma_all || (P.dst16 && x0 + TCOLS <= P.width)) {
            if (tid == 0) {
                tma::store_box_3d(&P.tmdst, x0, y0, cf, obuf);
                tma::store_commit();
            }
        } else {
            // byte stores, a thread per column.  (The buffer is next written two tiles on, after the barrier of the tile in between.)
            const int rows = min(TROWS, P.height - y0), cols = min(TCOLS, P.width - x0);
            uint8_t *d = P.dst + cf * P.fs_dst + (ptrdiff_t)y0 * P.sd + x0;
            if (tid < cols) {
#pragma unroll 1
                for (int r = 0; r < rows; ++r) d[(ptrdiff_t)r * P.sd + tid] = obuf[r * TCOLS + tid];
            }
        }
        advance(cx, cy, cf);
    }
    if (tid == 0) tma::store_wait_read<0>();   // shared memory stays valid until the last stores have read it
    umma::fence_before();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc<512>(*tmem_slot);
}

}  // namespace um
#endif  // HEVCASM_EXPERIMENTS

// ================================================================================================================================
// The same idea with the passes swapped, for one reference: the VERTICAL pass on the tensor cores, the horizontal pass in the
// threads.  With 8-bit input the first pass of the reference has shift 0 (pred_inter.c:182-200), so the two passes are one exact
// separable integer convolution and their order is free (the intermediate has the same range [-6120, 22440] either way).
//     D[m][n] = sum_k A[m][k] * B[k][n],   m = output row of the tile (128), n = staged column (256), k = staged row (160)
//     A[m][k] = ytap[k - m]                (Toeplitz band of the vertical taps, s8, K-major, built once per CTA)
//     B[k][n] = ref[y0 - TOP + k][x0 - 16 + n]   (u8, MN-major: image rows exactly as they lie in memory)
// Two TMA boxes {128 bytes x 136 rows} with the 128-byte swizzle ARE the swizzled MN-major operand (descriptor pinned by
// tools/umma_vfirst_probe.cu; a swizzled box must start on a 16-byte boundary, hence the 16-byte left margin).  Now a thread
// = one output ROW, and what it reads from TMEM are consecutive COLUMNS: it slides the horizontal taps along its row (pairs in
// a register ring, IDP.2A), and its output bytes are horizontally adjacent - four are clipped and packed by two cvt.pack.sat,
// sixteen leave in one STS.128 into a swizzled buffer; no byte stores, no barrier among the consumers.  Per output sample on the CUDA
// cores (8-tap, round 2): 3 IDP.2A + 1 IADD3 + 1 SHF + 0.5 I2IP + 0.06 STS, pairs straight from tcgen05.ld.pack::16b.
// Tile = 128 rows x 192 columns; four consumer warpgroups take 48 columns each.
namespace uv {

constexpr int TROWS = 128;                 // output rows per tile = MMA M = TMEM lanes
constexpr int NWG = 4, CPW = 48;           // consumer warpgroups; output columns per warpgroup (three 16-byte chunks)
constexpr int TCOLS = NWG * CPW;           // 192 output columns per tile: 3840, 1920 and 7680 are whole numbers of tiles
constexpr int N = 256;                     // MMA N = staged columns (16 left of the tile + 192 + halo; two 128-byte boxes)
constexpr int KROWS = 160;                 // MMA K (5 steps of 32): staged rows, of which BOXR are loaded
constexpr int BOXR = TROWS + 8;            // rows per box (>= TROWS + TAPS - 1)
constexpr int A_BYTES = TROWS * KROWS;     // 20480: Toeplitz operand, [chunk][m][16]
constexpr int BOX_BYTES = BOXR * 128;      // one 128-column block of a stage: the BOXR rows that arrive.  The MMAs read KROWS rows; rows BOXR ..
                                           // KROWS-1 are whatever bytes follow in shared memory and only ever meet zero taps (integers: 0 * x = 0)
constexpr int STAGE_BYTES = 2 * BOX_BYTES;
constexpr int NST = 3;                     // image stages in flight.  A set's boxes are requested as soon as the MMAs NST sets earlier have read
                                           // the stage; with two stages the kernel was bound by that round trip (TMA latency + MMA time) / 2 per
                                           // set - 61 / 109 us per 16 4K planes whatever the consumers cost (profiles/r02_pred.md)
// one output tile in shared memory = the two boxes its TMA stores read: columns 0..127 as rows of 128 bytes with the 128-byte swizzle,
// columns 128..191 as rows of 64 bytes with the 64-byte swizzle.  A thread owns one row; with the swizzle the eight threads of a quarter
// warp hit eight different 16-byte bank groups, so every STS.128 is conflict-free (the row-major buffer of round 1 had 76 % of its
// shared-memory wavefronts in conflict, profiles/r01_ncu_pred_hv_umma6.txt).
constexpr int OA_BYTES = TROWS * 128, OB_BYTES = TROWS * 64, O_BYTES = OA_BYTES + OB_BYTES;
constexpr int NOBUF = 3;                   // output buffers: the store of tile i-2 may still be reading while tile i is written
constexpr int CONSUMERS = NWG * 128, THREADS = CONSUMERS + 96;   // + loader, MMA and store warps (one thread of each works)
constexpr uint32_t TX_BYTES = 2 * BOXR * 128;
template <bool BI>
struct Geom {
    static constexpr int NREF = BI ? 2 : 1;
    // output buffers before the stages: the stray rows the last stage's MMAs read beyond its boxes must lie inside the allocation
    static constexpr int O_OFF = NREF * A_BYTES, B_OFF = O_OFF + NOBUF * O_BYTES, BAR_OFF = B_OFF + NST * STAGE_BYTES + (KROWS - BOXR) * 128;
    static constexpr int SMEM_BYTES = 1024 + BAR_OFF + 256;
};

struct alignas(64) Params {
    CUtensorMap tmref[2];     // per reference: 32-bit words from x = -16, rows from -(TAPS/2-1), frames; boxes of 32 words x BOXR rows, 128-byte swizzle
    CUtensorMap tmdst[2];     // destination planes: boxes of 128 bytes (128-byte swizzle) and of 64 bytes (64-byte swizzle) x TROWS rows (valid when dst16)
    uint8_t *dst;
    ptrdiff_t sd, fs_dst;
    int width, height;
    int dst16;                // destination rows are 16-byte aligned (pointer and strides): tiles leave by TMA store
    int tiles_x, tiles_y, n_tiles;
    long long *prof;          // experiments build: per-phase clock64 totals of CTA 0 ([role][phase]: loader, mma, storer, consumer thread 0), else null
    int xfrac[2];             // horizontal fraction per reference (8-tap: selects the instantiation of the horizontal pass)
    int8_t ytap[2][8];        // vertical taps per reference (the MMA's Toeplitz bands)
    int x2[2][4];             // horizontal tap pairs (t0,t1) (t2,t3) .. per reference (PackedCoefs::x2e)
    int x2o[2][4];            // horizontal tap pairs (t1,t2) (t3,t4) (t5,t6) per reference: the pairing of a window that starts on an odd column
};

// byte offset of 16-byte chunk g (0..11) of row `row` in an output buffer
__device__ __forceinline__ int out_chunk_offset(int row, int g)
{
    return g < 8 ? row * 128 + ((g ^ (row & 7)) << 4) : OA_BYTES + row * 64 + (((g - 8) ^ ((row >> 1) & 3)) << 4);
}

//   MODE 0 (one reference)      : out = clip((sum + 2048) >> 12), four neighbours per word
//   MODE 1 (first of two)       : mid = (int16)(sum >> 6), two neighbours per word (the reference's C truncates to int16, pred_inter.c:124)
//   MODE 2 (second of two)      : out = clip((mid + (int16)(sum >> 6) + 64) >> 7)  (pred_inter.c:490-501)
template <int MODE>
__device__ __forceinline__ void emit(int a, int x, int (&o)[4], uint32_t (&out)[CPW / 4], uint32_t (&mid)[CPW / 2])
{
    if (MODE == 0) {
        o[x & 3] = a >> 12;
    } else if (MODE == 1) {
        o[x & 1] = a >> 6;
        if (x & 1) mid[x >> 1] = pack16(o[0], o[1]);   // the low halves: int16 wrap
    } else {
        // (mid, this reference) as one int16 pair; IDP.2A with taps (1, 1) sign-extends and adds both halves
        const uint32_t pr = __byte_perm(mid[x >> 1], (uint32_t)(a >> 6), (x & 1) ? 0x5432 : 0x5410);
        o[x & 3] = dp2a_lo(pr, 0x0101, 64) >> 7;
    }
    if (MODE != 1 && (x & 3) == 3) out[x >> 2] = pack_sat_u8(o[0], o[1], o[2], o[3]);   // clip to [0, 255] and pack: two cvt.pack.sat
}

// The horizontal pass of one thread, any filter (used for the 4-tap filter): output row = its TMEM lane, CPW output columns from
// CPW + TAPS - 1 staged ones; pairs of neighbours packed by PRMT into a register ring, TAPS / 2 IDP.2A per sample.  Everything about the
// column index is compile-time; 8 columns arrive per tcgen05.ld, the next 8 are in flight while these are consumed.
template <int TAPS, int MODE>
__device__ __forceinline__ void horizontal_pass(const int (&xtap2)[4], uint32_t tcol /* lane, accumulator, first staged column */, uint32_t (&out)[CPW / 4],
                                                uint32_t (&mid)[CPW / 2])
{
    constexpr int CH = 8, NCH = (CPW + TAPS - 1 + CH - 1) / CH;
    static_assert(CH % TAPS == 0, "ring slots must be compile-time");
    uint32_t ring[TAPS];
    int prev = 0;
#pragma unroll
    for (int k = 0; k < TAPS; ++k) ring[k] = 0;
    int x2[TAPS / 2];
#pragma unroll
    for (int g = 0; g < TAPS / 2; ++g) x2[g] = xtap2[g];
    int v[2][CH], o[4];
    umma::tmem_ld8(tcol, v[0]);
    umma::tmem_ld_wait(v[0]);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        if (c + 1 < NCH) umma::tmem_ld8(tcol + CH * (c + 1), v[(c + 1) & 1]);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int r = CH * c + j, x = r - (TAPS - 1);   // staged column, output column
            if (x >= CPW) continue;
            ring[r % TAPS] = pack16(prev, v[c & 1][j]);   // pair (r-1, r)
            prev = v[c & 1][j];
            if (x < 0) continue;
            // output column x takes the pairs ending at columns x+1, x+3, .. = slots (r + 2 + 2g) mod TAPS
            int a = MODE == 0 ? 2048 : 0;
#pragma unroll
            for (int g = 0; g < TAPS / 2; ++g) a = dp2a_lo(ring[(r + 2 + 2 * g) % TAPS], x2[g], a);
            emit<MODE>(a, x, o, out, mid);
        }
        if (c + 1 < NCH) umma::tmem_ld_wait(v[(c + 1) & 1]);
    }
}

// The horizontal pass for the 8-tap luma filters, written around two facts (tools/tmem_probe.cu, profiles/r02_pred.md):
//  * tcgen05.ld ... .pack::16b hands back the accumulator columns as int16 pairs (2i, 2i+1) - the operand form of IDP.2A - with no
//    PRMT at all, but only for EVEN column addresses.  A window that starts on an even column q uses the pairs with taps
//    (t0,t1) (t2,t3) (t4,t5) (t6,t7); one that starts on an odd column uses (.,t0) (t1,t2) (t3,t4) (t5,t6) (t7,.).
//  * every luma filter has unit taps at its ends - (-1,4,-10,58,17,-5,1,0), (-1,4,-11,40,40,-11,4,-1), (0,1,-5,17,58,-10,4,-1) - and a
//    unit tap is an integer add of the plain 32-bit column value (IADD3, ALU pipe) instead of a lane of a dot product (FMA pipe, half
//    rate).  With the right split every output sample costs 3 IDP.2A + 1 IADD3 (half-sample positions: 4 + 0 on even windows), where
//    the round-1 kernel spent 4 IDP.2A + 1.1 PRMT and was bound by the IDP issue rate (ncu: fma pipe 36.4 % = 73 % of the IDP rate).
// q counts staged columns from one LEFT of the first tap of output column 0 (so that q = 0 is an even TMEM column): output column x reads
// q = x+1 .. x+8.  Columns arrive in chunks of 8 (one packed + one plain load), chunk c+1 in flight while c-1 and c are consumed.
//   FR = horizontal fraction 1, 2, 3; FR = 0 (second reference of a pair at a full-sample column) is the single tap 64.
template <int FR, int MODE>
__device__ __forceinline__ void horizontal_pass8(const int (&xe)[4], const int (&xo)[4], uint32_t tq0 /* lane, accumulator, column q = 0 */,
                                                 uint32_t (&out)[CPW / 4], uint32_t (&mid)[CPW / 2])
{
    constexpr int NCH = CPW / 8 + 1;
    uint32_t pk[3][4];
    int pv[3][8], o[4];
    int ce[4], co[3];
#pragma unroll
    for (int g = 0; g < 4; ++g) ce[g] = xe[g];
#pragma unroll
    for (int g = 0; g < 3; ++g) co[g] = xo[g];
#define HV_PK(q) pk[((q) >> 3) % 3][((q) & 7) >> 1]
#define HV_PV(q) pv[((q) >> 3) % 3][(q) & 7]
    umma::tmem_ld8_pack16(tq0, pk[0]);
    umma::tmem_ld8(tq0, pv[0]);
    umma::tmem_ld_wait(pk[0], pv[0]);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        if (c + 1 < NCH) {
            umma::tmem_ld8_pack16(tq0 + 8 * (c + 1), pk[(c + 1) % 3]);
            umma::tmem_ld8(tq0 + 8 * (c + 1), pv[(c + 1) % 3]);
        }
        if (c >= 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int x = 8 * (c - 1) + i;
                int a = MODE == 0 ? 2048 : 0;
                if (FR == 0) {
                    a = HV_PV(x + 4) << 6;
                } else if (x & 1) {   // window starts on the even column q = x+1
                    if (FR == 1) {
                        a += HV_PV(x + 7);
                        a = dp2a_lo(HV_PK(x + 5), ce[2], dp2a_lo(HV_PK(x + 3), ce[1], dp2a_lo(HV_PK(x + 1), ce[0], a)));
                    } else if (FR == 2) {
                        a = dp2a_lo(HV_PK(x + 7), ce[3], dp2a_lo(HV_PK(x + 5), ce[2], dp2a_lo(HV_PK(x + 3), ce[1], dp2a_lo(HV_PK(x + 1), ce[0], a))));
                    } else {
                        a += HV_PV(x + 2);
                        a = dp2a_lo(HV_PK(x + 7), ce[3], dp2a_lo(HV_PK(x + 5), ce[2], dp2a_lo(HV_PK(x + 3), ce[1], a)));
                    }
                    // (the pairs (-1, 4) / (4, -1) as shift-and-add on the plain values instead of an IDP.2A: ptxas turns the adds into IMAD on
                    //  the same FMA pipe - measured slower, 1765 vs 1422 cycles per tile for the first reference's pass)
                } else {              // window starts on the odd column q = x+1: pairs (x+2,x+3) (x+4,x+5) (x+6,x+7) carry taps 1..6
                    if (FR == 1) a -= HV_PV(x + 1);
                    else if (FR == 2) a = a - HV_PV(x + 1) - HV_PV(x + 8);
                    else a -= HV_PV(x + 8);
                    a = dp2a_lo(HV_PK(x + 6), co[2], dp2a_lo(HV_PK(x + 4), co[1], dp2a_lo(HV_PK(x + 2), co[0], a)));
                }
                emit<MODE>(a, x, o, out, mid);
            }
        }
        if (c + 1 < NCH) umma::tmem_ld_wait(pk[(c + 1) % 3], pv[(c + 1) % 3]);
    }
#undef HV_PK
#undef HV_PV
}

template <int TAPS, int MODE>
__device__ __forceinline__ void horizontal_any(const Params &P, int rf, uint32_t tcol /* lane, accumulator, first staged column of this warpgroup */,
                                               uint32_t (&out)[CPW / 4], uint32_t (&mid)[CPW / 2])
{
    if (TAPS == 8) {
        switch (P.xfrac[rf]) {   // uniform over the launch
            case 1: horizontal_pass8<1, MODE>(P.x2[rf], P.x2o[rf], tcol - 1, out, mid); break;
            case 2: horizontal_pass8<2, MODE>(P.x2[rf], P.x2o[rf], tcol - 1, out, mid); break;
            case 3: horizontal_pass8<3, MODE>(P.x2[rf], P.x2o[rf], tcol - 1, out, mid); break;
            default: horizontal_pass8<0, MODE>(P.x2[rf], P.x2o[rf], tcol - 1, out, mid); break;
        }
    } else {
        horizontal_pass<TAPS, MODE>(P.x2[rf], tcol, out, mid);
    }
}

// Work is a sequence of MMA sets q = 0, 1, 2, ..: one per tile (one reference) or two per tile (q even: reference 0, q odd:
// reference 1).  Set q uses image stage q % NST and accumulator q & 1, so with two references each has its own accumulator and
// the MMAs of one overlap the horizontal pass over the other.
//
// Warp roles (profiles/r02_pred.md): 16 consumer warps and THREE single-thread service warps.  Round 1 ran all service work - TMA
// requests, MMA issue, TMA stores - in one thread, and that thread was the critical path: 2700 cycles of serial issue latency per set
// (650 for the two stores, 740 for five MMAs, 670 for the request) against 1900 cycles of consumer arithmetic, measured with clock64
// around every phase.  Split over three warps the phases overlap:
//   loader : waits `empty[stage]` (the MMAs that read the stage have completed), requests the set's two boxes -> `full[stage]`
//   mma    : waits `full[stage]` and `consumed[acc]`, issues the five MMAs, commits `empty[stage]`, waits `ofree[buffer]`, commits `done[acc]`
//   storer : waits `ready[buffer]` (all consumers have written the tile), issues the TMA stores, waits until they have read the buffer,
//            arrives on `ofree[buffer]`
// Consumers arrive once per WARP (tma::mbar_arrive_warp).
#ifdef HEVCASM_EXPERIMENTS
#define HV_LAP(k) lap(k)
#else
#define HV_LAP(k) ((void)0)
#endif
template <int TAPS, bool BI>
__global__ void __launch_bounds__(THREADS, 1) pred_vh_kernel(const __grid_constant__ Params P)
{
    using G = Geom<BI>;
    constexpr int NREF = G::NREF, SPT = BI ? 2 : 1;   // MMA sets per tile
    static_assert(TROWS + TAPS - 1 <= BOXR && BOXR <= KROWS && 16 + TCOLS + TAPS / 2 <= N && CPW % 16 == 0 && ((16 - (TAPS / 2 - 1)) & 1) == 1, "tile geometry");
    extern __shared__ __align__(128) uint8_t us_raw[];
    uint8_t *const us_smem = us_raw + ((1024 - (tma::smem_u32(us_raw) & 1023)) & 1023);   // the 128-byte swizzle atoms sit on 1024-byte boundaries
    uint8_t *const sA = us_smem;                    // [reference][chunk][m][16]
    uint8_t *const sB = us_smem + G::B_OFF;         // [stage][128-column block][row][128]
    uint8_t *const sO = us_smem + G::O_OFF;         // [buffer]{[row][128], [row][64]}
    uint64_t *const full = reinterpret_cast<uint64_t *>(us_smem + G::BAR_OFF);   // [NST] image boxes of the stage have landed
    uint64_t *const empty = full + NST;                                          // [NST] the MMAs reading the stage have completed
    uint64_t *const done = empty + NST;                                          // [2] the MMAs into the accumulator have completed (and the output buffer is free)
    uint64_t *const consumed = done + 2;                                         // [2] every consumer warp has read the accumulator
    uint64_t *const ready = consumed + 2;                                        // [NOBUF] every consumer warp has written its part of the output tile
    uint64_t *const ofree = ready + NOBUF;                                       // [NOBUF] the TMA stores have read the output buffer
    uint32_t *const tmem_slot = reinterpret_cast<uint32_t *>(ofree + NOBUF);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NST; ++i) tma::mbar_init(full + i, 1), tma::mbar_init(empty + i, 1);
#pragma unroll
        for (int i = 0; i < 2; ++i) tma::mbar_init(done + i, 1), tma::mbar_init(consumed + i, CONSUMERS / 32);
#pragma unroll
        for (int i = 0; i < NOBUF; ++i) tma::mbar_init(ready + i, CONSUMERS / 32), tma::mbar_init(ofree + i, 1);
    }
    if (threadIdx.x < 32) umma::tmem_alloc<512>(tmem_slot);
    __syncthreads();   // barriers exist: the loader's first requests go out before the Toeplitz operands are built

    // tile -> (column, row, frame) of tiles: one division at the start, additions with carries afterwards (every role walks the same sequence)
    const int t0 = blockIdx.x, tstep = gridDim.x;
    const int n_mine = t0 < P.n_tiles ? (P.n_tiles - t0 + tstep - 1) / tstep : 0, nq = n_mine * SPT;   // this CTA's tiles, MMA sets
    const int per = P.tiles_x * P.tiles_y;
    const int sf = tstep / per, sby = (tstep - sf * per) / P.tiles_x, sbx = tstep - sf * per - sby * P.tiles_x;
    int cf = t0 / per, cy = (t0 - cf * per) / P.tiles_x, cx = t0 - cf * per - cy * P.tiles_x;
    auto advance = [&]() {
        cx += sbx;
        if (cx >= P.tiles_x) cx -= P.tiles_x, ++cy;
        cy += sby;
        if (cy >= P.tiles_y) cy -= P.tiles_y, ++cf;
        cf += sf;
    };
    const int role = (int)threadIdx.x - CONSUMERS;   // 0: loader, 32: mma, 64: storer (first lane of each service warp)
#ifdef HEVCASM_EXPERIMENTS
    long long pt[6] = {0, 0, 0, 0, 0, 0}, tk = clock64();
    auto lap = [&](int k) { const long long t = clock64(); pt[k] += t - tk; tk = t; };
#endif

    // loader state: set index, its stage, the parity to wait for on `empty` (a fresh barrier passes a wait for parity 1)
    int lq = 0, lst = 0;
    uint32_t eph = 1;
    auto request = [&]() {
        tma::mbar_expect_tx(full + lst, TX_BYTES);
        uint8_t *b = sB + lst * STAGE_BYTES;
        const CUtensorMap *map = &P.tmref[BI ? (lq & 1) : 0];
        tma::load_box_3d(b, map, cx * (TCOLS / 4), cy * TROWS, cf, full + lst);   // x in 32-bit words
        tma::load_box_3d(b + BOX_BYTES, map, cx * (TCOLS / 4) + 32, cy * TROWS, cf, full + lst);
        if (!BI || (lq & 1)) advance();
        ++lq;
        if (++lst == NST) lst = 0, eph ^= 1;
    };
    if (role == 0) {   // the first NST sets travel while the Toeplitz operands are built
#pragma unroll 1
        while (lq < NST && lq < nq) request();
    }

    // Toeplitz bands, one 16-byte chunk per step: output row m reads staged rows m .. m + TAPS - 1
    for (int i = threadIdx.x; i < NREF * (KROWS / 16) * TROWS; i += THREADS) {
        const int m = i % TROWS, kc = (i / TROWS) % (KROWS / 16), rf = i / (TROWS * (KROWS / 16));
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const int t = 16 * kc + b - m;
            if (t >= 0 && t < TAPS) w[b >> 2] |= (uint32_t)(uint8_t)P.ytap[rf][t] << (8 * (b & 3));
        }
        *reinterpret_cast<uint4 *>(sA + rf * A_BYTES + kc * (TROWS * 16) + m * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    // (the staged rows no box writes, BOXR .. KROWS-1, only ever meet zero taps)
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tm = *tmem_slot;

    // a TMA store clips rows exactly but columns only at 16-byte granularity (measured: a 200-byte-wide plane was written up to
    // byte 207), so a partial right-hand tile of a plane whose width is not a multiple of 16 leaves by byte stores instead
    const bool tma_all = P.dst16 && (P.width & 15) == 0;
    auto by_tma = [&](int x) { return tma_all || (P.dst16 && (x + 1) * TCOLS <= P.width); };

    if (role == 0) {
        // ------------------------------------------------------------------------------------------------ loader
#pragma unroll 1
        while (lq < nq) {
            tma::mbar_wait(empty + lst, eph);
            HV_LAP(0);
            request();
            HV_LAP(1);
        }
    } else if (role == 32) {
        // ------------------------------------------------------------------------------------------------ MMA issue
        constexpr uint32_t IDESC = umma::idesc_i8(true, false, false, N, true);   // A = taps (s8, K-major), B = image (u8, MN-major)
        // A: K-major, no swizzle (LBO = distance between 16-byte k chunks, SBO = between groups of 8 rows).  B: 128-byte-swizzled
        // MN-major (SBO = groups of 8 k, 1024 bytes; LBO = the second 128-column block); a K-step of 32 rows is 4096 bytes
        const uint64_t da0 = umma::smem_desc(tma::smem_u32(sA), TROWS * 16, 128), db0 = umma::smem_desc(tma::smem_u32(sB), BOX_BYTES, 1024, 2);
        int st = 0, ob = 0;
        uint32_t fph = 0, oph = 1;
#pragma unroll 1
        for (int q = 0; q < nq; ++q) {
            const int s = q & 1;   // accumulator (and, with two references, the reference)
            // descriptors first, waits afterwards: whatever arithmetic follows a wait is serial latency between `consumed` and `done`.
            // (descriptor arithmetic on the 14-bit address field, 16-byte units; the operands end below 256 KB, so no carry leaves the field)
            uint64_t da[KROWS / 32], db[KROWS / 32];
#pragma unroll
            for (int ks = 0; ks < KROWS / 32; ++ks) {
                da[ks] = da0 + (uint64_t)(((BI ? s : 0) * A_BYTES + ks * 2 * (TROWS * 16)) >> 4);
                db[ks] = db0 + (uint64_t)((st * STAGE_BYTES + ks * 4096) >> 4);
            }
            if (!BI || s == 1) {
                // the consumers fill output buffer `ob` once they see `done`: the TMA stores of the tile NOBUF back must have read it (long ago)
                tma::mbar_wait(ofree + ob, oph);
                if (++ob == NOBUF) ob = 0, oph ^= 1;
            }
            tma::mbar_wait(full + st, fph);
            HV_LAP(0);
            if (q >= 2) tma::mbar_wait(consumed + s, ((q >> 1) & 1) ^ 1);   // set q-2 has left this accumulator
            umma::fence_after();
            HV_LAP(1);
#pragma unroll
            for (int ks = 0; ks < KROWS / 32; ++ks) umma::mma_i8(tm + s * N, da[ks], db[ks], IDESC, ks);
            umma::commit(done + s);
            umma::commit(empty + st);   // the stage may be refilled as soon as these MMAs have read it
            if (++st == NST) st = 0, fph ^= 1;
            HV_LAP(2);
        }
    } else if (role == 64) {
        // ------------------------------------------------------------------------------------------------ TMA stores
        int ob = 0;
        uint32_t rph = 0;
#pragma unroll 1
        for (int t = 0; t < n_mine; ++t) {
            tma::mbar_wait(ready + ob, rph);
            HV_LAP(0);
            if (by_tma(cx)) {   // one box per swizzle span
                tma::store_box_3d(&P.tmdst[0], cx * TCOLS, cy * TROWS, cf, sO + ob * O_BYTES);
                if (cx * TCOLS + 128 < P.width) tma::store_box_3d(&P.tmdst[1], cx * TCOLS + 128, cy * TROWS, cf, sO + ob * O_BYTES + OA_BYTES);
                tma::store_commit();
                HV_LAP(1);
                tma::store_wait_read<0>();
            }
            tma::mbar_arrive(ofree + ob);
            advance();
            if (++ob == NOBUF) ob = 0, rph ^= 1;
            HV_LAP(2);
        }
    } else if (role < 0) {
        // ------------------------------------------------------------------------------------------------ consumers
        const int wg = threadIdx.x >> 7, row = threadIdx.x & 127, warp = row >> 5;   // warpgroup; output row of the tile = TMEM lane; warp inside the warpgroup
        const uint32_t tlane = tm + ((uint32_t)(warp * 32) << 16) + (16 - (TAPS / 2 - 1)) + wg * CPW;
        int ob = 0;
#pragma unroll 1
        for (int it = 0; it < n_mine; ++it) {
            uint8_t *const obuf = sO + ob * O_BYTES;
            uint32_t out[CPW / 4], mid[CPW / 2];
            int a;   // accumulator of the set that produces the output bytes
            if (BI) {
                a = 1;
                const uint32_t ph = it & 1;
                tma::mbar_wait(done + 0, ph);
                umma::fence_after();
                HV_LAP(0);
                horizontal_any<TAPS, 1>(P, 0, tlane, out, mid);
                umma::fence_before();
                tma::mbar_arrive_warp(consumed + 0);   // reference 0 of the next tile may take accumulator 0
                HV_LAP(1);
                tma::mbar_wait(done + 1, ph);
                umma::fence_after();
                HV_LAP(2);
                horizontal_any<TAPS, 2>(P, 1, tlane + N, out, mid);
            } else {
                a = it & 1;
                tma::mbar_wait(done + a, (it >> 1) & 1);
                umma::fence_after();
                HV_LAP(0);
                horizontal_any<TAPS, 0>(P, 0, tlane + a * N, out, mid);
            }
            HV_LAP(3);
#pragma unroll
            for (int i = 0; i < CPW / 16; ++i)
                *reinterpret_cast<uint4 *>(obuf + out_chunk_offset(row, wg * (CPW / 16) + i)) = make_uint4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
            umma::fence_before();       // this thread's TMEM reads are complete before its warp reports the accumulator consumed
            tma::mbar_arrive_warp(consumed + a);
            if (by_tma(cx)) {
                umma::fence_async_smem();   // the output bytes -> visible to the TMA store
            } else {
                // byte stores (partial right-hand tile of an odd width, or a destination TMA cannot describe): all consumers, after all have written
                asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory");
                const int x0 = cx * TCOLS, y0 = cy * TROWS, rows = min(TROWS, P.height - y0), cols = min(TCOLS, P.width - x0);
                uint8_t *d = P.dst + cf * P.fs_dst + (ptrdiff_t)y0 * P.sd + x0;
                for (int i = threadIdx.x; i < rows * cols; i += CONSUMERS) {
                    const int r = i / cols, c = i - r * cols;
                    d[(ptrdiff_t)r * P.sd + c] = obuf[out_chunk_offset(r, c >> 4) + (c & 15)];
                }
            }
            tma::mbar_arrive_warp(ready + ob);
            advance();
            if (++ob == NOBUF) ob = 0;
            HV_LAP(4);
        }
    }
#ifdef HEVCASM_EXPERIMENTS
    if (P.prof && blockIdx.x == 0 && (role == 0 || role == 32 || role == 64 || threadIdx.x == 0))
        for (int k = 0; k < 6; ++k) P.prof[(role < 0 ? 3 : role / 32) * 8 + k] = pt[k];
#endif
    umma::fence_before();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc<512>(*tmem_slot);
}
#undef HV_LAP

}  // namespace uv
