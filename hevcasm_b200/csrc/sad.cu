// hevcasm_b200 - SAD / SSD kernels for sm_100a.
//
// Reference semantics: sad.c:47-60 (hevcasm_sad_c_ref), sad.c:101-121 (hevcasm_sad_multiref_4_c_ref),
// ssd.c:43-55 (hevcasm_ssd_c_ref) of kupix/hevcasm.  All outputs are exact 32-bit integer sums.
//
// Design (not a translation of the x86 psadbw loops):
//   * motion-estimation sweep: a CTA stages the source tile and the reference search window of that tile in shared
//     memory once; each thread owns one 8x8 (or 8x4 / 4x8 / 4x4) source CELL in registers and walks the window rows,
//     feeding every window row to all candidates it belongs to.  64 accumulators (8 dy x 8 dx) live in registers; one
//     VABSDIFF4.U8.ACC retires 4 absolute differences + the add.  Byte-misaligned candidates cost one funnel shift
//     per 4-byte window word per misalignment, shared by the two candidates (dx, dx+4) that use it.
//   * cell sums are composed on chip into the requested PU size(s) (8x8 -> 16x16 -> 32x32 -> 64x64 for the pyramid
//     entry point), so the frames are read from HBM once for all PU sizes; results leave through a swizzled shared
//     staging buffer as fully coalesced 128-bit stores.
//   * SSD: |a-b| per byte (VABSDIFF4) then IDP.4A.U8.U8 of the difference with itself, 4x4 partials composed in
//     shared memory.
#include "common.cuh"
#include "tma.cuh"

#include <cstdlib>
#include <cstring>

namespace hv {

// ------------------------------------------------------------------------------------------------ cell core

template <int CW, int CH>
struct SadCell {
    static constexpr int KW = CW / 4;  // words per cell row
    uint32_t src[CH][KW];
    uint32_t acc[8][8];  // [dy][dx]

    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[j][i] = 0;
    }

    // s -> word (0, 0) of this cell in the staged source tile
    __device__ __forceinline__ void load_src(const uint32_t *s, int pitch_words)
    {
#pragma unroll
        for (int r = 0; r < CH; ++r) {
            if (KW == 2) {
                const uint2 v = *reinterpret_cast<const uint2 *>(s + r * pitch_words);
                src[r][0] = v.x;
                src[r][KW - 1] = v.y;
            } else {
                src[r][0] = s[r * pitch_words];
            }
        }
    }

    // one window row r (0 .. CH+6) given the KW+2 words that start at the cell's x: feeds every candidate the row belongs to.
    // shc = nullptr: realignment by funnel shift (ALU pipe); else shc[s-1] = 2^(32-8s) and the shifts run on the FMA pipe.
    template <bool FS = false>
    __device__ __forceinline__ void consume_row(int r, const uint32_t (&W)[KW + 2], const uint32_t *shc = nullptr)
    {
        uint32_t S[4][KW + 1];
#pragma unroll
        for (int j = 0; j <= KW; ++j) {
            S[0][j] = W[j];
            if (FS) {
                S[1][j] = shr_bytes_fma(W[j], W[j + 1], shc[0]);
                S[2][j] = shr_bytes_fma(W[j], W[j + 1], shc[1]);
                S[3][j] = shr_bytes_fma(W[j], W[j + 1], shc[2]);
            } else {
                S[1][j] = shr_bytes(W[j], W[j + 1], 1);
                S[2][j] = shr_bytes(W[j], W[j + 1], 2);
                S[3][j] = shr_bytes(W[j], W[j + 1], 3);
            }
        }
#pragma unroll
        for (int dy = 0; dy < 8; ++dy) {
            const int sr = r - dy;
            if (sr < 0 || sr >= CH) continue;
#pragma unroll
            for (int dx = 0; dx < 8; ++dx)
#pragma unroll
                for (int k = 0; k < KW; ++k) acc[dy][dx] = sad4(src[sr][k], S[dx & 3][(dx >> 2) + k], acc[dy][dx]);
        }
    }

    // win -> the staged window word holding candidate (dx index 0, dy index 0) of this cell, i.e. window row 0 at the
    // cell's x.  Consumes CH+7 window rows of CW+8 bytes.
    __device__ __forceinline__ void run(const uint32_t *win, int pitch_words)
    {
#pragma unroll
        for (int r = 0; r < CH + 7; ++r) {
            uint32_t W[KW + 2];
            if (KW == 2) {
                const uint2 a = *reinterpret_cast<const uint2 *>(win + r * pitch_words);
                const uint2 b = *reinterpret_cast<const uint2 *>(win + r * pitch_words + 2);
                W[0] = a.x, W[1] = a.y, W[KW] = b.x, W[KW + 1] = b.y;
            } else {
#pragma unroll
                for (int j = 0; j < KW + 2; ++j) W[j] = win[r * pitch_words + j];
            }
            consume_row(r, W);
        }
    }

    // The same for a window whose rows sit in shared memory with the byte alignment they have in global memory (TMA boxes
    // start on 16-byte boundaries): win8 -> the 8-byte aligned word pair that contains the cell's first window byte;
    // WO1 = 1 if that byte lies in the odd word of the pair, bs = its byte offset inside the word (0 when BS is false).
    template <int WO1, bool BS, bool FS>
    __device__ __forceinline__ void run_aligned(const uint32_t *win8, int pitch_words, int bs, const uint32_t *shc)
    {
        static_assert(KW == 2, "8-wide cells only");
#pragma unroll
        for (int r = 0; r < CH + 7; ++r) {
            uint32_t L[6];
            const uint2 a = *reinterpret_cast<const uint2 *>(win8 + r * pitch_words);
            const uint2 b = *reinterpret_cast<const uint2 *>(win8 + r * pitch_words + 2);
            L[0] = a.x, L[1] = a.y, L[2] = b.x, L[3] = b.y, L[4] = 0, L[5] = 0;
            if (WO1 || BS) {
                const uint2 c = *reinterpret_cast<const uint2 *>(win8 + r * pitch_words + 4);
                L[4] = c.x, L[5] = c.y;
            }
            uint32_t W[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) W[j] = BS ? __funnelshift_r(L[WO1 + j], L[WO1 + j + 1], 8 * bs) : L[WO1 + j];
            consume_row<FS>(r, W, shc);
        }
    }
};

// swizzled int4 staging of 64 candidates per cell: [cell][16 groups of 4 candidates]
__device__ __forceinline__ int cb_index(int cell, int g) { return cell * 16 + (g ^ (cell & 15)); }

template <int CW, int CH>
__device__ __forceinline__ void store_cell(int4 *cb, int cell, const SadCell<CW, CH> &c)
{
#pragma unroll
    for (int g = 0; g < 16; ++g) {
        const int dy = g >> 1, dx = (g & 1) * 4;
        cb[cb_index(cell, g)] = make_int4((int)c.acc[dy][dx], (int)c.acc[dy][dx + 1], (int)c.acc[dy][dx + 2], (int)c.acc[dy][dx + 3]);
    }
}

__device__ __forceinline__ int4 add4(int4 a, int4 b) { return make_int4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// ------------------------------------------------------------------------------------------------ pyramid sweep

namespace pyr {
constexpr int TW = 128, TH = 64, CX = TW / 8, CY = TH / 8, NT = CX * CY;  // 128 threads, one 8x8 cell each
constexpr int WIN_PITCH = (TW + 16) / 4;                                   // words; keeps rows 16-byte aligned
constexpr int WIN_ROWS = TH + 8;
constexpr int SRC_PITCH = TW / 4;
constexpr int STAGE_BYTES = (WIN_PITCH * WIN_ROWS + SRC_PITCH * TH) * 4;  // 18560
constexpr int CB_BYTES = NT * 64 * 4;                                      // 32768 (aliases the staging area)
constexpr int L16_BYTES = (CX / 2) * (CY / 2) * 64 * 4;                    // 8192
constexpr int L32_BYTES = (CX / 4) * (CY / 4) * 64 * 4;                    // 2048
constexpr int SMEM_BYTES = (CB_BYTES > STAGE_BYTES ? CB_BYTES : STAGE_BYTES) + L16_BYTES + L32_BYTES;
}  // namespace pyr

struct PyramidParams {
    const uint8_t *src, *ref;
    ptrdiff_t ss, sr, fs_src, fs_ref;
    int width, height, dx0, dy0;
    int tile_x_base;  // first tile column this launch covers (the generic kernel finishes what the fast one leaves)
    int32_t *out[4];  // 8, 16, 32, 64
};

__global__ void __launch_bounds__(pyr::NT) sad_sweep_pyramid_kernel(PyramidParams p)
{
    using namespace pyr;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *win = reinterpret_cast<uint32_t *>(smem);
    uint32_t *srct = win + WIN_PITCH * WIN_ROWS;
    int4 *cb = reinterpret_cast<int4 *>(smem);
    int4 *l16 = reinterpret_cast<int4 *>(smem + (CB_BYTES > STAGE_BYTES ? CB_BYTES : STAGE_BYTES));
    int4 *l32 = l16 + L16_BYTES / 16;

    const int tid = threadIdx.x;
    const int f = blockIdx.z;
    const int npx8 = p.width >> 3, npy8 = p.height >> 3;
    const int cx0 = (blockIdx.x + p.tile_x_base) * CX, cy0 = blockIdx.y * CY;  // first cell of the tile
    const int vcx = min(CX, npx8 - cx0), vcy = min(CY, npy8 - cy0);
    const int x0 = cx0 * 8, y0 = cy0 * 8;

    const uint8_t *src = p.src + f * p.fs_src + (ptrdiff_t)y0 * p.ss + x0;
    const uint8_t *ref = p.ref + f * p.fs_ref + (ptrdiff_t)(y0 + p.dy0) * p.sr + (x0 + p.dx0);
    stage_tile_u8(srct, SRC_PITCH, src, p.ss, vcx * 2, vcy * 8, tid, NT);
    stage_tile_u8(win, WIN_PITCH, ref, p.sr, vcx * 2 + 2, vcy * 8 + 7, tid, NT);
    __syncthreads();

    const int cx = tid % CX, cy = tid / CX;
    SadCell<8, 8> cell;
    cell.clear();
    if (cx < vcx && cy < vcy) {
        cell.load_src(srct + cy * 8 * SRC_PITCH + cx * 2, SRC_PITCH);
        cell.run(win + cy * 8 * WIN_PITCH + cx * 2, WIN_PITCH);
    }
    __syncthreads();  // staging area is dead from here on; cb aliases it
    store_cell(cb, tid, cell);
    __syncthreads();

    // level 0 (8x8): linear, coalesced copy-out
    if (p.out[0]) {
        int4 *o = reinterpret_cast<int4 *>(p.out[0]) + (size_t)f * npy8 * npx8 * 16;
        for (int i = tid; i < NT * 16; i += NT) {
            const int c = i >> 4, g = i & 15, ccx = c % CX, ccy = c / CX;
            if (ccx < vcx && ccy < vcy) o[((size_t)(cy0 + ccy) * npx8 + (cx0 + ccx)) * 16 + g] = cb[cb_index(c, g)];
        }
    }
    // level 1 (16x16) from four cells
    {
        const int npx = p.width >> 4, npy = p.height >> 4, px0 = cx0 >> 1, py0 = cy0 >> 1;
        int4 *o = p.out[1] ? reinterpret_cast<int4 *>(p.out[1]) + (size_t)f * npy * npx * 16 : nullptr;
        for (int i = tid; i < (CX / 2) * (CY / 2) * 16; i += NT) {
            const int pu = i >> 4, g = i & 15, px = pu % (CX / 2), py = pu / (CX / 2);
            const int c00 = (2 * py) * CX + 2 * px;
            const int4 v = add4(add4(cb[cb_index(c00, g)], cb[cb_index(c00 + 1, g)]), add4(cb[cb_index(c00 + CX, g)], cb[cb_index(c00 + CX + 1, g)]));
            l16[i] = v;
            if (o && px0 + px < npx && py0 + py < npy) o[((size_t)(py0 + py) * npx + (px0 + px)) * 16 + g] = v;
        }
    }
    __syncthreads();
    // level 2 (32x32) from four 16x16
    {
        const int npx = p.width >> 5, npy = p.height >> 5, px0 = cx0 >> 2, py0 = cy0 >> 2;
        int4 *o = p.out[2] ? reinterpret_cast<int4 *>(p.out[2]) + (size_t)f * npy * npx * 16 : nullptr;
        for (int i = tid; i < (CX / 4) * (CY / 4) * 16; i += NT) {
            const int pu = i >> 4, g = i & 15, px = pu % (CX / 4), py = pu / (CX / 4);
            const int q00 = ((2 * py) * (CX / 2) + 2 * px) * 16 + g;
            const int4 v = add4(add4(l16[q00], l16[q00 + 16]), add4(l16[q00 + (CX / 2) * 16], l16[q00 + (CX / 2) * 16 + 16]));
            l32[i] = v;
            if (o && px0 + px < npx && py0 + py < npy) o[((size_t)(py0 + py) * npx + (px0 + px)) * 16 + g] = v;
        }
    }
    __syncthreads();
    // level 3 (64x64) from four 32x32
    if (p.out[3]) {
        const int npx = p.width >> 6, npy = p.height >> 6, px0 = cx0 >> 3, py0 = cy0 >> 3;
        int4 *o = reinterpret_cast<int4 *>(p.out[3]) + (size_t)f * npy * npx * 16;
        for (int i = tid; i < (CX / 8) * (CY / 8) * 16; i += NT) {
            const int pu = i >> 4, g = i & 15, px = pu % (CX / 8), py = pu / (CX / 8);
            const int q00 = ((2 * py) * (CX / 4) + 2 * px) * 16 + g;
            const int4 v = add4(add4(l32[q00], l32[q00 + 16]), add4(l32[q00 + (CX / 4) * 16], l32[q00 + (CX / 4) * 16 + 16]));
            if (px0 + px < npx && py0 + py < npy) o[((size_t)(py0 + py) * npx + (px0 + px)) * 16 + g] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------ pyramid sweep, fast path
//
// Same tile, same arithmetic, but every loop has a compile-time trip count and compile-time divisors: the generic kernel
// above spends two thirds of its issue slots on staging index arithmetic (a division by a run-time row length per word)
// and on bounds logic in the copy-out loops (profiles/r01_sad_pyramid_v1.md).  Preconditions, checked by the host:
// tiles are full in x (the right-hand strip of a frame whose width is not a multiple of 128 goes to the generic kernel),
// source rows are 16-byte aligned and window rows 4-byte aligned.  Tiles may be partial in y: row indices are clamped
// for the loads (never reading below the last needed row) and the stores are masked.
template <int LEVEL_MASK>
__global__ void __launch_bounds__(pyr::NT) sad_pyramid_fast_kernel(PyramidParams p)
{
    using namespace pyr;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *win = reinterpret_cast<uint32_t *>(smem);
    uint32_t *srct = win + WIN_PITCH * WIN_ROWS;
    int4 *cb = reinterpret_cast<int4 *>(smem);
    int4 *l16 = reinterpret_cast<int4 *>(smem + (CB_BYTES > STAGE_BYTES ? CB_BYTES : STAGE_BYTES));
    int4 *l32 = l16 + L16_BYTES / 16;

    const int tid = threadIdx.x, f = blockIdx.z;
    const int npx8 = p.width >> 3, npy8 = p.height >> 3;
    const int cx0 = blockIdx.x * CX, cy0 = blockIdx.y * CY;
    const int vcy = min(CY, npy8 - cy0);
    const int x0 = cx0 * 8, y0 = cy0 * 8;

    // ---- stage: source tile with 128-bit loads, window with 32-bit loads (its origin is only 4-byte aligned)
    {
        const uint8_t *src = p.src + f * p.fs_src + (ptrdiff_t)y0 * p.ss + x0;
        const int last = vcy * 8 - 1;
        int4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = tid + k * NT, row = idx >> 3, ch = idx & 7;
            v[k] = ldg_stream(reinterpret_cast<const int4 *>(src + (ptrdiff_t)min(row, last) * p.ss) + ch);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = tid + k * NT, row = idx >> 3, ch = idx & 7;
            *reinterpret_cast<int4 *>(srct + row * SRC_PITCH + ch * 4) = v[k];
        }
        const uint8_t *ref = p.ref + f * p.fs_ref + (ptrdiff_t)(y0 + p.dy0) * p.sr + (x0 + p.dx0);
        constexpr int WW = TW / 4 + 2, TOTAL = WW * WIN_ROWS, ITERS = (TOTAL + NT - 1) / NT;  // 34 words x 71 rows
        const int lastw = vcy * 8 + 6;
        uint32_t w[ITERS];
#pragma unroll
        for (int k = 0; k < ITERS; ++k) {
            const int idx = min(tid + k * NT, TOTAL - 1), row = idx / WW, col = idx - row * WW;
            w[k] = __ldg(reinterpret_cast<const uint32_t *>(ref + (ptrdiff_t)min(row, lastw) * p.sr) + col);
        }
#pragma unroll
        for (int k = 0; k < ITERS; ++k) {
            const int idx = tid + k * NT, row = idx / WW, col = idx - row * WW;
            if (idx < TOTAL) win[row * WIN_PITCH + col] = w[k];
        }
    }
    __syncthreads();

    const int cx = tid % CX, cy = tid / CX;
    SadCell<8, 8> cell;
    cell.clear();
    cell.load_src(srct + cy * 8 * SRC_PITCH + cx * 2, SRC_PITCH);
    cell.run(win + cy * 8 * WIN_PITCH + cx * 2, WIN_PITCH);
    __syncthreads();  // staging area is dead from here on; cb aliases it
    store_cell(cb, tid, cell);
    __syncthreads();

    // ---- copy-out.  i = tid + 128 k : int4 slot g = tid & 15 of cell / PU (tid >> 4) + 8 k
    const int g = tid & 15, hi = tid >> 4;  // hi = 0..7
    if (LEVEL_MASK & 1) {
        // cell c = hi + 8k -> ccx = hi + 8 (k & 1), ccy = k >> 1
        int4 *o = reinterpret_cast<int4 *>(p.out[0]) + ((size_t)f * npy8 + cy0) * npx8 * 16 + (size_t)(cx0 + hi) * 16 + g;
        const int s0 = g ^ hi, s1 = g ^ (hi + 8);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int ccy = k >> 1, c = hi + 8 * k;
            const int4 v = cb[c * 16 + ((k & 1) ? s1 : s0)];
            if (ccy < vcy) o[(size_t)ccy * npx8 * 16 + (k & 1) * 128] = v;
        }
    }
    // level 1 (16x16): PU = hi + 8k -> px = hi, py = k;  cells (2py, 2px), (2py, 2px+1), (2py+1, 2px), (2py+1, 2px+1)
    {
        const int npx = p.width >> 4, npy = p.height >> 4, py0 = cy0 >> 1;
        int4 *o = (LEVEL_MASK & 2) ? reinterpret_cast<int4 *>(p.out[1]) + ((size_t)f * npy + py0) * npx * 16 + (size_t)((cx0 >> 1) + hi) * 16 + g : nullptr;
        const int sa = g ^ (2 * hi), sb = g ^ (2 * hi + 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c00 = (2 * k) * CX + 2 * hi;
            const int4 v = add4(add4(cb[c00 * 16 + sa], cb[(c00 + 1) * 16 + sb]), add4(cb[(c00 + CX) * 16 + sa], cb[(c00 + CX + 1) * 16 + sb]));
            l16[tid + k * NT] = v;
            if ((LEVEL_MASK & 2) && py0 + k < npy && 2 * k < vcy) o[(size_t)k * npx * 16] = v;
        }
    }
    __syncthreads();
    // level 2 (32x32): PU = hi (0..7) -> px = hi & 3, py = hi >> 2
    {
        const int npx = p.width >> 5, npy = p.height >> 5, px = hi & 3, py = hi >> 2;
        const int q00 = ((2 * py) * (CX / 2) + 2 * px) * 16 + g;
        const int4 v = add4(add4(l16[q00], l16[q00 + 16]), add4(l16[q00 + (CX / 2) * 16], l16[q00 + (CX / 2) * 16 + 16]));
        l32[tid] = v;
        if ((LEVEL_MASK & 4) && (cy0 >> 2) + py < npy && (cx0 >> 2) + px < npx)
            reinterpret_cast<int4 *>(p.out[2])[(((size_t)f * npy + (cy0 >> 2) + py) * npx + (cx0 >> 2) + px) * 16 + g] = v;
    }
    if (LEVEL_MASK & 8) {
        __syncthreads();
        // level 3 (64x64): PU = hi (0..1), first 32 threads
        if (tid < 32) {
            const int npx = p.width >> 6, npy = p.height >> 6, px = hi;
            const int q00 = (2 * px) * 16 + g;
            const int4 v = add4(add4(l32[q00], l32[q00 + 16]), add4(l32[q00 + (CX / 4) * 16], l32[q00 + (CX / 4) * 16 + 16]));
            if ((cy0 >> 3) < npy && (cx0 >> 3) + px < npx) reinterpret_cast<int4 *>(p.out[3])[(((size_t)f * npy + (cy0 >> 3)) * npx + (cx0 >> 3) + px) * 16 + g] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------ pyramid sweep, TMA staged
//
// The preferred path.  One elected thread asks the TMA unit for the two boxes of the tile - the 128 x 64 source tile and
// the 144 x 71 search window, whose origin (x0 + dx0, y0 + dy0) may have any byte alignment - and every thread then waits
// on the transaction barrier.  No thread computes a load address, nothing is clamped at frame edges (the hardware
// zero-fills what lies outside the declared extent; those cells are never stored), and the copy-out is the unrolled one
// of the fast kernel.  What remains per 8x8 cell is 1024 VABSDIFF4 + 135 funnel shifts + 30 64-bit shared loads + ~50
// instructions of copy-out / composition.
struct PyramidTmaParams {
    CUtensorMap tm_src, tm_ref;
    int width, height;
    int win_shift;  // byte offset (0..15) of the window origin inside its 16-byte aligned box
    uint32_t shc[3];  // 2^24, 2^16, 2^8 as run-time values (see shr_bytes_fma); used by the FS = true instantiations
    int32_t *out[4];
};

// TMA boxes must start on a 16-byte boundary in global memory (anything else faults - measured with tools/tma_probe.cu), so
// the window box starts at the 16-byte boundary below the window origin and is 16 bytes wider; the cell indexes it with the
// residual byte offset (run_aligned).
namespace pyr {
constexpr int TMA_WIN_BYTES = TW + 32, TMA_WIN_PITCH = TMA_WIN_BYTES / 4, TMA_WIN_ROWS = TH + 7;  // 160 x 71
constexpr int TMA_SRC_OFF = (TMA_WIN_BYTES * TMA_WIN_ROWS + 127) / 128 * 128;
constexpr uint32_t TMA_TX_BYTES = TMA_WIN_BYTES * TMA_WIN_ROWS + TW * TH;
static_assert(TMA_SRC_OFF + TW * TH <= CB_BYTES, "TMA staging must fit under the copy-out buffer");
}  // namespace pyr

// key of a candidate for the fused argmin: SAD in the high bits, candidate index in the low 6 -> the minimum key is the
// smallest SAD and, among equals, the first candidate in raster order (64x64: 1044480 * 64 < 2^31)
__device__ __forceinline__ uint32_t best_key4(int4 v, int g)
{
    const uint32_t c = 4 * g;
    return min(min((uint32_t)v.x * 64 + c, (uint32_t)v.y * 64 + c + 1), min((uint32_t)v.z * 64 + c + 2, (uint32_t)v.w * 64 + c + 3));
}
// minimum over the 16 lanes that share a PU (lanes differ in their low 4 bits)
__device__ __forceinline__ uint32_t best_reduce16(uint32_t k)
{
#pragma unroll
    for (int o = 8; o; o >>= 1) k = min(k, __shfl_xor_sync(0xffffffffu, k, o));
    return k;
}
__device__ __forceinline__ void best_store(int32_t *out, size_t pu, uint32_t key) { reinterpret_cast<int2 *>(out)[pu] = make_int2((int)(key >> 6), (int)(key & 63)); }

// the four SADs of an int4 as uint16 (8x8 and 16x16 PUs: at most 8*8*255 = 16 320 resp. 16*16*255 = 65 280, so nothing is lost)
__device__ __forceinline__ uint2 pack_u16x4(int4 v) { return make_uint2(__byte_perm((uint32_t)v.x, (uint32_t)v.y, 0x5410), __byte_perm((uint32_t)v.z, (uint32_t)v.w, 0x5410)); }

// OUT = 0: out[l] receive the 64 SADs of every PU as int32.  OUT = 1: out[l] receive {min SAD, candidate index} per PU.
// OUT = 2: as 0, but the 8x8 and 16x16 levels are written as uint16 (half the bytes that have to cross PCIe in the host forms).
constexpr int OUT_FULL = 0, OUT_BEST = 1, OUT_PACKED = 2;
template <int LEVEL_MASK, int WO1, bool BS, int OUT, bool FS>
__global__ void __launch_bounds__(pyr::NT, 5) sad_pyramid_tma_kernel(const __grid_constant__ PyramidTmaParams p)
{
    using namespace pyr;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *win = reinterpret_cast<uint32_t *>(smem);                 // 71 rows x 160 bytes
    uint32_t *srct = reinterpret_cast<uint32_t *>(smem + TMA_SRC_OFF);  // 64 rows x 128 bytes, 128-byte aligned
    int4 *cb = reinterpret_cast<int4 *>(smem);
    int4 *l16 = reinterpret_cast<int4 *>(smem + CB_BYTES);
    int4 *l32 = l16 + L16_BYTES / 16;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + SMEM_BYTES);  // behind everything else; never aliased

    const int tid = threadIdx.x, f = blockIdx.z;
    const int npx8 = p.width >> 3, npy8 = p.height >> 3;
    const int cx0 = blockIdx.x * CX, cy0 = blockIdx.y * CY;
    const int vcx = min(CX, npx8 - cx0), vcy = min(CY, npy8 - cy0);
    const int x0 = cx0 * 8, y0 = cy0 * 8;

    if (tid == 0) tma::mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        tma::mbar_expect_tx(bar, TMA_TX_BYTES);  // the two boxes, zero-filled parts included
        tma::load_box_3d(win, &p.tm_ref, x0, y0, f, bar);
        tma::load_box_3d(srct, &p.tm_src, x0, y0, f, bar);
    }
    tma::mbar_wait(bar, 0);

    const int cx = tid % CX, cy = tid / CX;
    SadCell<8, 8> cell;
    cell.clear();
    cell.load_src(srct + cy * 8 * SRC_PITCH + cx * 2, SRC_PITCH);
    cell.template run_aligned<WO1, BS, FS>(win + cy * 8 * TMA_WIN_PITCH + cx * 2 + ((p.win_shift >> 3) << 1), TMA_WIN_PITCH, p.win_shift & 3, p.shc);
    __syncthreads();  // staging area is dead from here on; cb aliases it
    store_cell(cb, tid, cell);
    __syncthreads();

    const int g = tid & 15, hi = tid >> 4;
    if (OUT == OUT_BEST) {
        // level 0: every thread reduces its own cell straight from the accumulators
        if ((LEVEL_MASK & 1) && cx < vcx && cy < vcy) {
            uint32_t k = 0xffffffffu;
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) k = min(k, cell.acc[j][i] * 64 + (uint32_t)(8 * j + i));
            best_store(p.out[0], ((size_t)f * npy8 + cy0 + cy) * npx8 + cx0 + cx, k);
        }
        {
            const int npx = p.width >> 4, npy = p.height >> 4, py0 = cy0 >> 1;
            const int sa = g ^ (2 * hi), sb = g ^ (2 * hi + 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c00 = (2 * k) * CX + 2 * hi;
                const int4 v = add4(add4(cb[c00 * 16 + sa], cb[(c00 + 1) * 16 + sb]), add4(cb[(c00 + CX) * 16 + sa], cb[(c00 + CX + 1) * 16 + sb]));
                l16[tid + k * NT] = v;
                const uint32_t key = best_reduce16(best_key4(v, g));
                if ((LEVEL_MASK & 2) && g == 0 && (cx0 >> 1) + hi < npx && py0 + k < npy) best_store(p.out[1], ((size_t)f * npy + py0 + k) * npx + (cx0 >> 1) + hi, key);
            }
        }
        __syncthreads();
        {
            const int npx = p.width >> 5, npy = p.height >> 5, px = hi & 3, py = hi >> 2;
            const int q00 = ((2 * py) * (CX / 2) + 2 * px) * 16 + g;
            const int4 v = add4(add4(l16[q00], l16[q00 + 16]), add4(l16[q00 + (CX / 2) * 16], l16[q00 + (CX / 2) * 16 + 16]));
            l32[tid] = v;
            const uint32_t key = best_reduce16(best_key4(v, g));
            if ((LEVEL_MASK & 4) && g == 0 && (cy0 >> 2) + py < npy && (cx0 >> 2) + px < npx) best_store(p.out[2], ((size_t)f * npy + (cy0 >> 2) + py) * npx + (cx0 >> 2) + px, key);
        }
        if (LEVEL_MASK & 8) {
            __syncthreads();
            if (tid < 32) {
                const int npx = p.width >> 6, npy = p.height >> 6, px = hi;
                const int q00 = (2 * px) * 16 + g;
                const int4 v = add4(add4(l32[q00], l32[q00 + 16]), add4(l32[q00 + (CX / 4) * 16], l32[q00 + (CX / 4) * 16 + 16]));
                const uint32_t key = best_reduce16(best_key4(v, g));
                if (g == 0 && (cy0 >> 3) < npy && (cx0 >> 3) + px < npx) best_store(p.out[3], ((size_t)f * npy + (cy0 >> 3)) * npx + (cx0 >> 3) + px, key);
            }
        }
        return;
    }
    // (the element index of a PU's 4-candidate group is the same for both output widths: int4 index = uint2 index)
    auto put = [](int32_t *base, size_t idx, int4 v, bool packed) {
        if (packed) reinterpret_cast<uint2 *>(base)[idx] = pack_u16x4(v);
        else reinterpret_cast<int4 *>(base)[idx] = v;
    };
    constexpr bool PK = OUT == OUT_PACKED;
    if (LEVEL_MASK & 1) {
        const size_t o = ((size_t)f * npy8 + cy0) * npx8 * 16 + (size_t)(cx0 + hi) * 16 + g;
        const int s0 = g ^ hi, s1 = g ^ (hi + 8);
        const bool ok0 = hi < vcx, ok1 = hi + 8 < vcx;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int ccy = k >> 1, c = hi + 8 * k;
            const int4 v = cb[c * 16 + ((k & 1) ? s1 : s0)];
            if (ccy < vcy && ((k & 1) ? ok1 : ok0)) put(p.out[0], o + (size_t)ccy * npx8 * 16 + (k & 1) * 128, v, PK);
        }
    }
    {
        const int npx = p.width >> 4, npy = p.height >> 4, py0 = cy0 >> 1;
        const size_t o = ((size_t)f * npy + py0) * npx * 16 + (size_t)((cx0 >> 1) + hi) * 16 + g;
        const int sa = g ^ (2 * hi), sb = g ^ (2 * hi + 1);
        const bool okx = (cx0 >> 1) + hi < npx;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c00 = (2 * k) * CX + 2 * hi;
            const int4 v = add4(add4(cb[c00 * 16 + sa], cb[(c00 + 1) * 16 + sb]), add4(cb[(c00 + CX) * 16 + sa], cb[(c00 + CX + 1) * 16 + sb]));
            l16[tid + k * NT] = v;
            if ((LEVEL_MASK & 2) && okx && py0 + k < npy) put(p.out[1], o + (size_t)k * npx * 16, v, PK);
        }
    }
    __syncthreads();
    {
        const int npx = p.width >> 5, npy = p.height >> 5, px = hi & 3, py = hi >> 2;
        const int q00 = ((2 * py) * (CX / 2) + 2 * px) * 16 + g;
        const int4 v = add4(add4(l16[q00], l16[q00 + 16]), add4(l16[q00 + (CX / 2) * 16], l16[q00 + (CX / 2) * 16 + 16]));
        l32[tid] = v;
        if ((LEVEL_MASK & 4) && (cy0 >> 2) + py < npy && (cx0 >> 2) + px < npx)
            reinterpret_cast<int4 *>(p.out[2])[(((size_t)f * npy + (cy0 >> 2) + py) * npx + (cx0 >> 2) + px) * 16 + g] = v;
    }
    if (LEVEL_MASK & 8) {
        __syncthreads();
        if (tid < 32) {
            const int npx = p.width >> 6, npy = p.height >> 6, px = hi;
            const int q00 = (2 * px) * 16 + g;
            const int4 v = add4(add4(l32[q00], l32[q00 + 16]), add4(l32[q00 + (CX / 4) * 16], l32[q00 + (CX / 4) * 16 + 16]));
            if ((cy0 >> 3) < npy && (cx0 >> 3) + px < npx) reinterpret_cast<int4 *>(p.out[3])[(((size_t)f * npy + (cy0 >> 3)) * npx + (cx0 >> 3) + px) * 16 + g] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------ one PU size, TMA staged
//
// hevcasm_sad_sweep_frames for PUs whose sides are 8, 16, 32 or 64 (14 of the reference's 23 partitions, sad.c:231-240): the
// same tile, staging and cell arithmetic as the pyramid kernel; the copy-out composes mx x my cells into the one requested
// PU size and places the 8 x 8 candidate tile (dx0+wx .., dy0+wy ..) inside the caller's ncx x ncy window.
struct RectTmaParams {
    CUtensorMap tm_src, tm_ref;
    int width, height, win_shift;
    int lmx, lmy;             // log2 of cells per PU in x / y
    int npx, npy;             // PU grid per frame
    int ncx, nc, wx, wy, nvx, nvy;  // window row length, candidates per PU, position and valid size of this 8 x 8 candidate tile
    uint32_t shc[3];
    int32_t *out;
};

template <int WO1, bool BS, bool FS>
__global__ void __launch_bounds__(pyr::NT, 5) sad_rect_tma_kernel(const __grid_constant__ RectTmaParams p)
{
    using namespace pyr;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *win = reinterpret_cast<uint32_t *>(smem);
    uint32_t *srct = reinterpret_cast<uint32_t *>(smem + TMA_SRC_OFF);
    int4 *cb = reinterpret_cast<int4 *>(smem);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + SMEM_BYTES);

    const int tid = threadIdx.x, f = blockIdx.z;
    const int cx0 = blockIdx.x * CX, cy0 = blockIdx.y * CY;
    const int x0 = cx0 * 8, y0 = cy0 * 8;

    if (tid == 0) tma::mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        tma::mbar_expect_tx(bar, TMA_TX_BYTES);
        tma::load_box_3d(win, &p.tm_ref, x0, y0, f, bar);
        tma::load_box_3d(srct, &p.tm_src, x0, y0, f, bar);
    }
    tma::mbar_wait(bar, 0);

    const int cx = tid % CX, cy = tid / CX;
    SadCell<8, 8> cell;
    cell.clear();
    cell.load_src(srct + cy * 8 * SRC_PITCH + cx * 2, SRC_PITCH);
    cell.template run_aligned<WO1, BS, FS>(win + cy * 8 * TMA_WIN_PITCH + cx * 2 + ((p.win_shift >> 3) << 1), TMA_WIN_PITCH, p.win_shift & 3, p.shc);
    __syncthreads();
    store_cell(cb, tid, cell);
    __syncthreads();

    // PUs of this tile: (CX >> lmx) x (CY >> lmy); item = (PU, slot g of 4 candidates)
    const int lpx = 4 - p.lmx, npu_tile = (CX >> p.lmx) * (CY >> p.lmy), mx = 1 << p.lmx, my = 1 << p.lmy;
    const int pu_x0 = cx0 >> p.lmx, pu_y0 = cy0 >> p.lmy;
    const bool vec = ((p.ncx | p.wx) & 3) == 0 && p.nvx == 8;
    for (int i = tid; i < npu_tile * 16; i += NT) {
        const int pu = i >> 4, g = i & 15, ppx = pu & ((1 << lpx) - 1), ppy = pu >> lpx;
        const int gx = pu_x0 + ppx, gy = pu_y0 + ppy;
        if (gx >= p.npx || gy >= p.npy) continue;
        int4 v;
        if (mx == 1 && my == 1) {   // 8x8 PUs: a PU is one cell (uniform branch: no composition loops)
            v = cb[cb_index(ppy * CX + ppx, g)];
        } else {
            v = make_int4(0, 0, 0, 0);
            for (int b = 0; b < my; ++b)
                for (int a = 0; a < mx; ++a) {
                    const int c = ((ppy << p.lmy) + b) * CX + (ppx << p.lmx) + a;
                    v = add4(v, cb[cb_index(c, g)]);
                }
        }
        const int dyi = g >> 1, dxi = (g & 1) * 4;
        if (dyi >= p.nvy) continue;
        int32_t *o = p.out + (((size_t)f * p.npy + gy) * p.npx + gx) * p.nc + (p.wy + dyi) * p.ncx + p.wx + dxi;
        if (vec) {
            *reinterpret_cast<int4 *>(o) = v;
        } else {
            if (dxi + 0 < p.nvx) o[0] = v.x;
            if (dxi + 1 < p.nvx) o[1] = v.y;
            if (dxi + 2 < p.nvx) o[2] = v.z;
            if (dxi + 3 < p.nvx) o[3] = v.w;
        }
    }
}

// ------------------------------------------------------------------------------------------------ single-size sweep
//
// Any PU size w x h (multiples of 4, <= 64), any dense candidate window, processed in tiles of 8 x 8 candidates.
// A CTA owns TPX x TPY whole PUs (<= 16 x 8 cells) and loops over the window tiles.

struct SweepParams {
    const uint8_t *src, *ref;
    ptrdiff_t ss, sr, fs_src, fs_ref;
    int w, h, npx, npy;       // PU size, PU grid per frame
    int tpx, tpy;             // PUs per CTA tile
    int dx0, dy0, ncx, ncy;   // candidate window
    int32_t *out;
};

constexpr int SWEEP_NT = 128;
constexpr int SWEEP_WIN_PITCH = 36, SWEEP_WIN_ROWS = 72, SWEEP_SRC_PITCH = 32, SWEEP_SRC_ROWS = 64;
constexpr int SWEEP_STAGE_BYTES = (SWEEP_WIN_PITCH * SWEEP_WIN_ROWS + SWEEP_SRC_PITCH * SWEEP_SRC_ROWS) * 4;
constexpr int SWEEP_CB_BYTES = SWEEP_NT * 64 * 4;
constexpr int SWEEP_SMEM_BYTES = SWEEP_STAGE_BYTES + SWEEP_CB_BYTES;

template <int CW, int CH>
__global__ void __launch_bounds__(SWEEP_NT) sad_sweep_kernel(SweepParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *win = reinterpret_cast<uint32_t *>(smem);
    uint32_t *srct = win + SWEEP_WIN_PITCH * SWEEP_WIN_ROWS;
    int4 *cb = reinterpret_cast<int4 *>(smem + SWEEP_STAGE_BYTES);

    const int tid = threadIdx.x, f = blockIdx.z;
    const int mx = p.w / CW, my = p.h / CH;                   // cells per PU
    const int pu_x0 = blockIdx.x * p.tpx, pu_y0 = blockIdx.y * p.tpy;
    const int vpx = min(p.tpx, p.npx - pu_x0), vpy = min(p.tpy, p.npy - pu_y0);
    const int ncellx = vpx * mx, ncelly = vpy * my;           // valid cells of this tile (<= 16 x 8)
    const int x0 = pu_x0 * p.w, y0 = pu_y0 * p.h;
    const int tile_w = vpx * p.w, tile_h = vpy * p.h;
    const int nc = p.ncx * p.ncy;

    const uint8_t *src = p.src + f * p.fs_src + (ptrdiff_t)y0 * p.ss + x0;
    stage_tile_u8(srct, SWEEP_SRC_PITCH, src, p.ss, tile_w / 4, tile_h, tid, SWEEP_NT);

    const int cx = tid % 16, cy = tid / 16;
    const bool active = cx < ncellx && cy < ncelly;
    SadCell<CW, CH> cell;
    bool src_loaded = false;

    for (int wy = 0; wy < p.ncy; wy += 8)
        for (int wx = 0; wx < p.ncx; wx += 8) {
            const int nvx = min(8, p.ncx - wx), nvy = min(8, p.ncy - wy);  // valid candidates of this window tile
            const uint8_t *ref = p.ref + f * p.fs_ref + (ptrdiff_t)(y0 + p.dy0 + wy) * p.sr + (x0 + p.dx0 + wx);
            __syncthreads();  // previous iteration's readers of win / cb are done
            stage_tile_u8(win, SWEEP_WIN_PITCH, ref, p.sr, (tile_w + nvx - 1 + 3) / 4, tile_h + nvy - 1, tid, SWEEP_NT);
            __syncthreads();
            cell.clear();
            if (active) {
                if (!src_loaded) {
                    cell.load_src(srct + cy * CH * SWEEP_SRC_PITCH + cx * (CW / 4), SWEEP_SRC_PITCH);
                    src_loaded = true;
                }
                cell.run(win + cy * CH * SWEEP_WIN_PITCH + cx * (CW / 4), SWEEP_WIN_PITCH);
            }
            store_cell(cb, tid, cell);
            __syncthreads();
            // compose cells into PUs and write the valid candidates
            for (int i = tid; i < vpx * vpy * 16; i += SWEEP_NT) {
                const int pu = i >> 4, g = i & 15, px = pu % vpx, py = pu / vpx;
                int4 v = make_int4(0, 0, 0, 0);
                for (int b = 0; b < my; ++b)
                    for (int a = 0; a < mx; ++a) v = add4(v, cb[cb_index((py * my + b) * 16 + px * mx + a, g)]);
                const int dyi = g >> 1, dxi = (g & 1) * 4;
                if (dyi < nvy) {
                    int32_t *o = p.out + (((size_t)f * p.npy + (pu_y0 + py)) * p.npx + (pu_x0 + px)) * nc + (wy + dyi) * p.ncx + wx + dxi;
                    if (dxi + 0 < nvx) o[0] = v.x;
                    if (dxi + 1 < nvx) o[1] = v.y;
                    if (dxi + 2 < nvx) o[2] = v.z;
                    if (dxi + 3 < nvx) o[3] = v.w;
                }
            }
        }
}

// ------------------------------------------------------------------------------------------------ list forms

__device__ __forceinline__ uint32_t ldg_word_unaligned(const uint8_t *p)
{
    const int a = (int)((uintptr_t)p & 3);
    const uint32_t *pa = reinterpret_cast<const uint32_t *>(p - a);
    uint32_t lo = __ldg(pa);
    if (a) lo = shr_bytes(lo, __ldg(pa + 1), a);
    return lo;
}

// one warp per (PU, candidate).  per_pu_mv: cand holds one vector per PU (n_cand == 1) or may be null.
__global__ void __launch_bounds__(256) sad_list_kernel(const uint8_t *__restrict__ src, ptrdiff_t ss, const uint8_t *__restrict__ ref,
                                                       ptrdiff_t sr, int w, int h, const int16_t *__restrict__ pu_xy, int n_pu,
                                                       const int16_t *__restrict__ cand, int n_cand, int per_pu_mv,
                                                       int32_t *__restrict__ out)
{
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (long long)n_pu * n_cand) return;
    const int i = (int)(gw / n_cand), c = (int)(gw - (long long)i * n_cand);
    const int x = pu_xy[2 * i], y = pu_xy[2 * i + 1];
    int dx = 0, dy = 0;
    if (cand) {
        const int ci = per_pu_mv ? i : c;
        dx = cand[2 * ci], dy = cand[2 * ci + 1];
    }
    const uint8_t *s = src + (ptrdiff_t)y * ss + x;
    const uint8_t *r = ref + (ptrdiff_t)(y + dy) * sr + (x + dx);
    const int wpr = w >> 2, total = wpr * h;
    uint32_t acc = 0;
    for (int idx = lane; idx < total; idx += 32) {
        const int row = idx / wpr, k = idx - row * wpr;
        acc = sad4(ldg_word_unaligned(s + (ptrdiff_t)row * ss + 4 * k), ldg_word_unaligned(r + (ptrdiff_t)row * sr + 4 * k), acc);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[gw] = (int32_t)acc;
}

// ------------------------------------------------------------------------------------------------ SSD

// one warp per listed block
__global__ void __launch_bounds__(256) ssd_list_kernel(const uint8_t *__restrict__ a, ptrdiff_t sa, const uint8_t *__restrict__ b,
                                                       ptrdiff_t sb, int n_size, const int16_t *__restrict__ blk_xy, int n,
                                                       int32_t *__restrict__ out)
{
    const int gw = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (gw >= n) return;
    const int x = blk_xy[2 * gw], y = blk_xy[2 * gw + 1];
    const uint8_t *pa = a + (ptrdiff_t)y * sa + x, *pb = b + (ptrdiff_t)y * sb + x;
    const int wpr = n_size >> 2, total = wpr * n_size;
    uint32_t acc = 0;
    for (int idx = lane; idx < total; idx += 32) {
        const int row = idx / wpr, k = idx - row * wpr;
        const uint32_t d = __vabsdiffu4(ldg_word_unaligned(pa + (ptrdiff_t)row * sa + 4 * k), ldg_word_unaligned(pb + (ptrdiff_t)row * sb + 4 * k));
        acc = dp4a_uu(d, d, acc);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[gw] = (int32_t)acc;
}

// Regular grid.  CTA tile = 128 x 32 samples, 64 threads... each thread owns a 16-byte x 4-row patch, i.e. four 4x4
// partial sums; partials are composed into N x N blocks through shared memory.
namespace ssdk {
constexpr int TW = 256, TH = 32, NT = (TW / 16) * (TH / 4);  // 128 threads
constexpr int PX = TW / 4, PY = TH / 4;                       // 64 x 8 partials
}  // namespace ssdk

struct SsdParams {
    const uint8_t *a, *b;
    ptrdiff_t sa, sb, fs_a, fs_b;
    int log2, nbx, nby;  // block size, block grid per frame
    int32_t *out;
};

template <bool ALIGNED>
__global__ void __launch_bounds__(ssdk::NT) ssd_frames_kernel(SsdParams p)
{
    using namespace ssdk;
    __shared__ __align__(16) uint32_t part[PY][PX + 4];
    const int tid = threadIdx.x, f = blockIdx.z;
    const int N = 1 << p.log2;
    const int TH_EFF = N > TH ? N : TH;              // a 64x64 block needs two 32-row passes
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH_EFF;
    const int vw = min(TW, p.nbx * N - x0), vh = min(TH_EFF, p.nby * N - y0);  // valid samples of this tile
    const int tx = tid % (TW / 16), ty = tid / (TW / 16);
    const uint8_t *a = p.a + f * p.fs_a + (ptrdiff_t)y0 * p.sa + x0;
    const uint8_t *b = p.b + f * p.fs_b + (ptrdiff_t)y0 * p.sb + x0;

    for (int pass = 0; pass < TH_EFF / TH; ++pass) {
        uint32_t acc[4] = {0, 0, 0, 0};
        const int yy = pass * TH + ty * 4;
        if (tx * 16 < vw && yy < vh) {
            uint32_t wa[4][4], wb[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint8_t *ra = a + (ptrdiff_t)(yy + r) * p.sa + tx * 16, *rb = b + (ptrdiff_t)(yy + r) * p.sb + tx * 16;
                if (ALIGNED) {
                    const int4 va = ldg_stream(reinterpret_cast<const int4 *>(ra)), vb = ldg_stream(reinterpret_cast<const int4 *>(rb));
                    wa[r][0] = va.x, wa[r][1] = va.y, wa[r][2] = va.z, wa[r][3] = va.w;
                    wb[r][0] = vb.x, wb[r][1] = vb.y, wb[r][2] = vb.z, wb[r][3] = vb.w;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const bool in = tx * 16 + 4 * k < vw;  // never touch words right of the last valid block
                        wa[r][k] = in ? ldg_word_unaligned(ra + 4 * k) : 0;
                        wb[r][k] = in ? ldg_word_unaligned(rb + 4 * k) : 0;
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t d = __vabsdiffu4(wa[r][k], wb[r][k]);
                    acc[k] = dp4a_uu(d, d, acc[k]);
                }
        }
        __syncthreads();
        *reinterpret_cast<uint4 *>(&part[ty][tx * 4]) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
        __syncthreads();
        // compose (N/4)^2 partials per block.  Blocks taller than the tile accumulate over passes in `carry`.
        const int m = N >> 2;                           // partials per block side
        const int mrows = min(m, PY);                   // partial rows available per pass
        const int bpr = PX / m, bpc = PY / mrows;       // blocks per tile row / column (per pass)
        for (int i = tid; i < bpr * bpc; i += NT) {
            const int bx = i % bpr, by = i / bpr;
            uint32_t s = 0;
            for (int r = 0; r < mrows; ++r)
                for (int c = 0; c < m; ++c) s += part[by * mrows + r][bx * m + c];
            const int gx = blockIdx.x * bpr + bx;
            const int gy = N > TH ? blockIdx.y : blockIdx.y * bpc + by;
            if (gx < p.nbx && gy < p.nby) {
                int32_t *o = p.out + ((size_t)f * p.nby + gy) * p.nbx + gx;
                if (N > TH && pass > 0) *o += (int32_t)s;  // same thread wrote pass 0
                else *o = (int32_t)s;
            }
        }
    }
}

}  // namespace hv

// ================================================================================================ C ABI

using namespace hv;

// realignment shifts on the FMA pipe unless HEVCASM_SAD_SHIFT=alu (A/B switch)
static void fill_shc(uint32_t (&shc)[3])
{
    const char *e = tune::knob("HEVCASM_SAD_SHIFT");
    const bool alu = e && !strcmp(e, "alu");
    shc[0] = alu ? 0u : 1u << 24, shc[1] = 1u << 16, shc[2] = 1u << 8;
}

static bool rect_ok(uint32_t rect, int &w, int &h)
{
    w = (int)(rect >> 8), h = (int)(rect & 0xff);
    return w >= 4 && h >= 4 && w <= 64 && h <= 64 && (w & 3) == 0 && (h & 3) == 0;
}

// ---- PU lists over a batch of frames: every PU has its own size, integer vector and frame ------------------------------------
// pus[i] = {x, y, w, h, dx, dy, frame}: out[i] = SAD (or SSD) of the w x h block of `src` at (x, y) and the block of `ref` at (x + dx, y + dy),
// both in plane `frame`.  Eight lanes per PU, four PUs per warp: the lanes walk the block's words in row-major order (32 contiguous bytes per
// PU and step), so an 8x8 PU keeps all eight busy for two steps where a whole warp per PU left half of it idle (mixed 8..64 list of 16 4K
// frames: 236 us with a warp per PU).  A descriptor whose w or h is not a multiple of 4 in 4..64 gives -1.
template <bool SSD>
__global__ void __launch_bounds__(256) pu_list_cost_kernel(const uint8_t *__restrict__ src, ptrdiff_t ss, ptrdiff_t fs_src, const uint8_t *__restrict__ ref,
                                                           ptrdiff_t sr, ptrdiff_t fs_ref, const int16_t *__restrict__ pus, int n_pu, int32_t *__restrict__ out)
{
    const long long gi = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    const bool live = gi < n_pu;
    int w = 0, h = 0;
    const uint8_t *s = src, *r = ref;
    if (live) {
        const int16_t *d = pus + gi * 7;
        const int x = d[0], y = d[1], dx = d[4], dy = d[5], f = d[6];
        w = d[2], h = d[3];
        s = src + f * fs_src + (ptrdiff_t)y * ss + x;
        r = ref + f * fs_ref + (ptrdiff_t)(y + dy) * sr + (x + dx);
    }
    const bool legal = w >= 4 && h >= 4 && w <= 64 && h <= 64 && !((w | h) & 3);
    const int wpr = w >> 2, total = legal ? wpr * h : 0;
    uint32_t acc = 0;
    for (int idx = sub; idx < total; idx += 8) {
        const int row = idx / wpr, k = idx - row * wpr;
        const uint32_t a = ldg_word_unaligned(s + (ptrdiff_t)row * ss + 4 * k), b = ldg_word_unaligned(r + (ptrdiff_t)row * sr + 4 * k);
        if (SSD) {
            const uint32_t df = __vabsdiffu4(a, b);
            acc = dp4a_uu(df, df, acc);
        } else {
            acc = sad4(a, b, acc);
        }
    }
#pragma unroll
    for (int o = 4; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (live && sub == 0) out[gi] = legal ? (int32_t)acc : -1;
}

extern "C" int hevcasm_sad_list_frames(const uint8_t *src, ptrdiff_t ss, const uint8_t *ref, ptrdiff_t sr, const int16_t *pus, int n_pu, ptrdiff_t fs_src,
                                       ptrdiff_t fs_ref, int32_t *sad, void *stream)
{
    if (n_pu < 0 || (n_pu > 0 && (!pus || !sad))) return HEVCASM_ERR_ARGUMENT;
    if (n_pu == 0) return 0;
    HV_LAUNCH(pu_list_cost_kernel<false>, (unsigned)((n_pu + 31) / 32), 256, 0, stream, src, ss, fs_src, ref, sr, fs_ref, pus, n_pu, sad);
    return 0;
}

extern "C" int hevcasm_ssd_list_frames(const uint8_t *srcA, ptrdiff_t sa, const uint8_t *srcB, ptrdiff_t sb, const int16_t *pus, int n_pu, ptrdiff_t fs_a,
                                       ptrdiff_t fs_b, int32_t *ssd, void *stream)
{
    if (n_pu < 0 || (n_pu > 0 && (!pus || !ssd))) return HEVCASM_ERR_ARGUMENT;
    if (n_pu == 0) return 0;
    HV_LAUNCH(pu_list_cost_kernel<true>, (unsigned)((n_pu + 31) / 32), 256, 0, stream, srcA, sa, fs_a, srcB, sb, fs_b, pus, n_pu, ssd);
    return 0;
}

extern "C" int hevcasm_sad_multiref_batch(const uint8_t *src, ptrdiff_t ss, const uint8_t *ref, ptrdiff_t sr, uint32_t rect,
                                          const int16_t *pu_xy, int n_pu, const int16_t *cand, int n_cand, int32_t *sad, void *stream)
{
    int w, h;
    if (!rect_ok(rect, w, h) || n_pu < 0 || n_cand < 1 || !cand) return HEVCASM_ERR_ARGUMENT;
    if (n_pu == 0) return 0;
    const long long warps = (long long)n_pu * n_cand;
    const unsigned grid = (unsigned)((warps + 7) / 8);
    HV_LAUNCH(sad_list_kernel, grid, 256, 0, stream, src, ss, ref, sr, w, h, pu_xy, n_pu, cand, n_cand, 0, sad);
    return 0;
}

extern "C" int hevcasm_sad_batch(const uint8_t *src, ptrdiff_t ss, const uint8_t *ref, ptrdiff_t sr, uint32_t rect, const int16_t *pu_xy,
                                 const int16_t *mv_xy, int n_pu, int32_t *sad, void *stream)
{
    int w, h;
    if (!rect_ok(rect, w, h) || n_pu < 0) return HEVCASM_ERR_ARGUMENT;
    if (n_pu == 0) return 0;
    const unsigned grid = (unsigned)((n_pu + 7) / 8);
    HV_LAUNCH(sad_list_kernel, grid, 256, 0, stream, src, ss, ref, sr, w, h, pu_xy, n_pu, mv_xy, 1, 1, sad);
    return 0;
}

extern "C" int hevcasm_sad_sweep_frames(const uint8_t *src, ptrdiff_t ss, const uint8_t *ref, ptrdiff_t sr, int width, int height,
                                        uint32_t rect, int dx0, int dy0, int ncx, int ncy, int n_frames, ptrdiff_t fs_src,
                                        ptrdiff_t fs_ref, int32_t *sad, void *stream)
{
    int w, h;
    if (!rect_ok(rect, w, h) || ncx < 1 || ncy < 1 || (long long)ncx * ncy > 256 || n_frames < 0 || width < 0 || height < 0) return HEVCASM_ERR_ARGUMENT;
    SweepParams p;
    p.src = src, p.ref = ref, p.ss = ss, p.sr = sr, p.fs_src = fs_src, p.fs_ref = fs_ref;
    p.w = w, p.h = h, p.npx = width / w, p.npy = height / h;
    p.dx0 = dx0, p.dy0 = dy0, p.ncx = ncx, p.ncy = ncy, p.out = sad;
    if (p.npx == 0 || p.npy == 0 || n_frames == 0) return 0;
    // sides of 8, 16, 32, 64 with TMA-describable planes: the TMA-staged kernel, one launch per 8 x 8 tile of the candidate window
    {
        const char *pin = tune::knob("HEVCASM_SAD_PATH");
        auto pow2_8_64 = [](int v) { return v == 8 || v == 16 || v == 32 || v == 64; };
        if ((!pin || !strcmp(pin, "tma")) && pow2_8_64(w) && pow2_8_64(h) && ((uintptr_t)src & 15) == 0 && ((uintptr_t)sad & 15) == 0 &&
            tma::describable(ss, fs_src, n_frames) && tma::describable(sr, fs_ref, n_frames)) {
            RectTmaParams t;
            fill_shc(t.shc);
            const int ext_x = p.npx * w, ext_y = p.npy * h;  // area covered by whole PUs
            t.width = ext_x, t.height = ext_y;
            t.lmx = w == 8 ? 0 : w == 16 ? 1 : w == 32 ? 2 : 3, t.lmy = h == 8 ? 0 : h == 16 ? 1 : h == 32 ? 2 : 3;
            t.npx = p.npx, t.npy = p.npy, t.ncx = ncx, t.nc = ncx * ncy, t.out = sad;
            int xs = 0;
            int e = tma::describe_u8(&t.tm_src, src, ss, fs_src, ext_x, ext_y, n_frames, pyr::TW, pyr::TH, &xs);
            const dim3 grid((ext_x / 8 + pyr::CX - 1) / pyr::CX, (ext_y / 8 + pyr::CY - 1) / pyr::CY, n_frames);
            const size_t smem_bytes = pyr::SMEM_BYTES + 16;
            for (int wy = 0; wy < ncy && !e; wy += 8)
                for (int wx = 0; wx < ncx && !e; wx += 8) {
                    t.wx = wx, t.wy = wy, t.nvx = ncx - wx < 8 ? ncx - wx : 8, t.nvy = ncy - wy < 8 ? ncy - wy : 8;
                    e = tma::describe_u8(&t.tm_ref, ref + (ptrdiff_t)(dy0 + wy) * sr + (dx0 + wx), sr, fs_ref, (long long)ext_x + 7, (long long)ext_y + 7, n_frames,
                                         pyr::TMA_WIN_BYTES, pyr::TMA_WIN_ROWS, &t.win_shift);
                    if (e) break;
                    const int wo1 = (t.win_shift >> 2) & 1, bs = t.win_shift & 3;
#define HV_RECT(WO1_, BS_)                                                \
    do {                                                                  \
        if (t.shc[0]) {                                                   \
            auto kern = sad_rect_tma_kernel<WO1_, BS_, true>;             \
            HV_CUDA((cudaError_t)set_max_smem(kern, smem_bytes));         \
            HV_LAUNCH(kern, grid, pyr::NT, smem_bytes, stream, t);        \
        } else {                                                          \
            auto kern = sad_rect_tma_kernel<WO1_, BS_, false>;            \
            HV_CUDA((cudaError_t)set_max_smem(kern, smem_bytes));         \
            HV_LAUNCH(kern, grid, pyr::NT, smem_bytes, stream, t);        \
        }                                                                 \
    } while (0)
                    if (wo1 && bs) HV_RECT(1, true);
                    else if (wo1) HV_RECT(1, false);
                    else if (bs) HV_RECT(0, true);
                    else HV_RECT(0, false);
#undef HV_RECT
                }
            if (!e) return 0;
            return e;
        }
    }
    const int cw = (w & 7) ? 4 : 8, ch = (h & 7) ? 4 : 8;
    p.tpx = (16 * cw) / w > 0 ? (16 * cw) / w : 1;
    p.tpy = (8 * ch) / h > 0 ? (8 * ch) / h : 1;
    const dim3 grid((p.npx + p.tpx - 1) / p.tpx, (p.npy + p.tpy - 1) / p.tpy, n_frames);
#define HV_SWEEP(CW_, CH_)                                                        \
    do {                                                                          \
        auto kern = sad_sweep_kernel<CW_, CH_>;                                   \
        HV_CUDA((cudaError_t)set_max_smem(kern, SWEEP_SMEM_BYTES));               \
        HV_LAUNCH(kern, grid, SWEEP_NT, SWEEP_SMEM_BYTES, stream, p);             \
    } while (0)
    if (cw == 8 && ch == 8) HV_SWEEP(8, 8);
    else if (cw == 8) HV_SWEEP(8, 4);
    else if (ch == 8) HV_SWEEP(4, 8);
    else HV_SWEEP(4, 4);
#undef HV_SWEEP
    return 0;
}

extern "C" int hevcasm_sad_sweep_pyramid_frames(const uint8_t *src, ptrdiff_t ss, const uint8_t *ref, ptrdiff_t sr, int width, int height,
                                                int dx0, int dy0, int n_frames, ptrdiff_t fs_src, ptrdiff_t fs_ref, int32_t *sad8,
                                                int32_t *sad16, int32_t *sad32, int32_t *sad64, void *stream)
{
    if (width < 8 || height < 8 || n_frames < 0) return HEVCASM_ERR_ARGUMENT;
    if (n_frames == 0) return 0;
    PyramidParams p;
    p.src = src, p.ref = ref, p.ss = ss, p.sr = sr, p.fs_src = fs_src, p.fs_ref = fs_ref;
    p.width = width, p.height = height, p.dx0 = dx0, p.dy0 = dy0, p.tile_x_base = 0;
    p.out[0] = sad8, p.out[1] = sad16, p.out[2] = sad32, p.out[3] = sad64;
    const int npx8 = width >> 3, npy8 = height >> 3;
    const int tiles_x = (npx8 + pyr::CX - 1) / pyr::CX, tiles_y = (npy8 + pyr::CY - 1) / pyr::CY;
    int full_x = 0;
    // path selection: TMA staged (any byte alignment, needs 16-byte strides) > LDG fast (4-byte aligned window) > generic.
    // HEVCASM_SAD_PATH=tma|fast|generic pins one for A/B profiling.
    const char *pin = tune::knob("HEVCASM_SAD_PATH");
    const bool all_levels = sad8 && sad16 && sad32 && sad64;
    const bool want_tma = !pin || !strcmp(pin, "tma"), want_fast = !pin || !strcmp(pin, "fast");
    if (want_tma && all_levels && ((uintptr_t)src & 15) == 0 && tma::describable(ss, fs_src, n_frames) && tma::describable(sr, fs_ref, n_frames)) {
        PyramidTmaParams t;
        t.width = width, t.height = height;
        fill_shc(t.shc);
        for (int l = 0; l < 4; ++l) t.out[l] = p.out[l];
        // source: bytes [0, npx8*8) x [0, npy8*8);  window: origin (dx0, dy0), 7 more bytes / rows than the source extent
        int xs_src = 0;
        int e = tma::describe_u8(&t.tm_src, src, ss, fs_src, (long long)npx8 * 8, (long long)npy8 * 8, n_frames, pyr::TW, pyr::TH, &xs_src);
        if (!e)
            e = tma::describe_u8(&t.tm_ref, ref + (ptrdiff_t)dy0 * sr + dx0, sr, fs_ref, (long long)npx8 * 8 + 7, (long long)npy8 * 8 + 7, n_frames,
                                 pyr::TMA_WIN_BYTES, pyr::TMA_WIN_ROWS, &t.win_shift);
        if (!e) {
            const dim3 grid(tiles_x, tiles_y, n_frames);
            const size_t smem_bytes = pyr::SMEM_BYTES + 16;
            const int wo1 = (t.win_shift >> 2) & 1, bs = t.win_shift & 3;
#define HV_TMA(WO1_, BS_)                                                 \
    do {                                                                  \
        if (t.shc[0]) {                                                   \
            auto kern = sad_pyramid_tma_kernel<15, WO1_, BS_, OUT_FULL, true>;  \
            HV_CUDA((cudaError_t)set_max_smem(kern, smem_bytes));         \
            HV_LAUNCH(kern, grid, pyr::NT, smem_bytes, stream, t);        \
        } else {                                                          \
            auto kern = sad_pyramid_tma_kernel<15, WO1_, BS_, OUT_FULL, false>; \
            HV_CUDA((cudaError_t)set_max_smem(kern, smem_bytes));         \
            HV_LAUNCH(kern, grid, pyr::NT, smem_bytes, stream, t);        \
        }                                                                 \
    } while (0)
            if (wo1 && bs) HV_TMA(1, true);
            else if (wo1) HV_TMA(1, false);
            else if (bs) HV_TMA(0, true);
            else HV_TMA(0, false);
#undef HV_TMA
            return 0;
        }
    }
    // fast path: all four outputs wanted, 16-byte aligned source rows, 4-byte aligned window rows
    const bool fast = want_fast && all_levels && (((uintptr_t)src | (uintptr_t)ss | (uintptr_t)fs_src) & 15) == 0 &&
                      ((((uintptr_t)ref + (uintptr_t)(ptrdiff_t)dx0) | (uintptr_t)sr | (uintptr_t)fs_ref) & 3) == 0;
    if (fast) {
        full_x = npx8 / pyr::CX;
        if (full_x > 0) {
            auto kern = sad_pyramid_fast_kernel<15>;
            HV_CUDA((cudaError_t)set_max_smem(kern, pyr::SMEM_BYTES));
            HV_LAUNCH(kern, dim3(full_x, tiles_y, n_frames), pyr::NT, pyr::SMEM_BYTES, stream, p);
        }
    }
    if (full_x < tiles_x) {
        p.tile_x_base = full_x;
        HV_CUDA((cudaError_t)set_max_smem(sad_sweep_pyramid_kernel, pyr::SMEM_BYTES));
        HV_LAUNCH(sad_sweep_pyramid_kernel, dim3(tiles_x - full_x, tiles_y, n_frames), pyr::NT, pyr::SMEM_BYTES, stream, p);
    }
    return 0;
}

// the TMA-staged pyramid kernel in one of its compact output modes (all four outputs required, TMA-describable planes)
template <int OUT>
static int launch_pyramid_compact(const uint8_t *src, ptrdiff_t ss, const uint8_t *ref, ptrdiff_t sr, int width, int height, int dx0, int dy0, int n_frames,
                                  ptrdiff_t fs_src, ptrdiff_t fs_ref, void *o8, void *o16, void *o32, void *o64, void *stream)
{
    if (width < 8 || height < 8 || n_frames < 0 || !o8 || !o16 || !o32 || !o64) return HEVCASM_ERR_ARGUMENT;
    if (n_frames == 0) return 0;
    // TMA staged only: needs a 16-byte aligned source origin and 16-byte multiples for all strides
    if (((uintptr_t)src & 15) != 0 || !tma::describable(ss, fs_src, n_frames) || !tma::describable(sr, fs_ref, n_frames)) return HEVCASM_ERR_ARGUMENT;
    if (OUT == OUT_PACKED && ((((uintptr_t)o8 | (uintptr_t)o16) & 7) != 0 || (((uintptr_t)o32 | (uintptr_t)o64) & 15) != 0)) return HEVCASM_ERR_ARGUMENT;
    const int npx8 = width >> 3, npy8 = height >> 3;
    PyramidTmaParams t;
    t.width = width, t.height = height;
    fill_shc(t.shc);
    t.out[0] = (int32_t *)o8, t.out[1] = (int32_t *)o16, t.out[2] = (int32_t *)o32, t.out[3] = (int32_t *)o64;
    int xs_src = 0;
    int e = tma::describe_u8(&t.tm_src, src, ss, fs_src, (long long)npx8 * 8, (long long)npy8 * 8, n_frames, pyr::TW, pyr::TH, &xs_src);
    if (!e)
        e = tma::describe_u8(&t.tm_ref, ref + (ptrdiff_t)dy0 * sr + dx0, sr, fs_ref, (long long)npx8 * 8 + 7, (long long)npy8 * 8 + 7, n_frames, pyr::TMA_WIN_BYTES,
                             pyr::TMA_WIN_ROWS, &t.win_shift);
    if (e) return e;
    const dim3 grid((npx8 + pyr::CX - 1) / pyr::CX, (npy8 + pyr::CY - 1) / pyr::CY, n_frames);
    const size_t smem_bytes = pyr::SMEM_BYTES + 16;
    const int wo1 = (t.win_shift >> 2) & 1, bs = t.win_shift & 3;
#define HV_TMA_COMPACT(WO1_, BS_)                                         \
    do {                                                                  \
        if (t.shc[0]) {                                                   \
            auto kern = sad_pyramid_tma_kernel<15, WO1_, BS_, OUT, true>;    \
            HV_CUDA((cudaError_t)set_max_smem(kern, smem_bytes));         \
            HV_LAUNCH(kern, grid, pyr::NT, smem_bytes, stream, t);        \
        } else {                                                          \
            auto kern = sad_pyramid_tma_kernel<15, WO1_, BS_, OUT, false>;   \
            HV_CUDA((cudaError_t)set_max_smem(kern, smem_bytes));         \
            HV_LAUNCH(kern, grid, pyr::NT, smem_bytes, stream, t);        \
        }                                                                 \
    } while (0)
    if (wo1 && bs) HV_TMA_COMPACT(1, true);
    else if (wo1) HV_TMA_COMPACT(1, false);
    else if (bs) HV_TMA_COMPACT(0, true);
    else HV_TMA_COMPACT(0, false);
#undef HV_TMA_COMPACT
    return 0;
}

extern "C" int hevcasm_sad_sweep_pyramid_best_frames(const uint8_t *src, ptrdiff_t ss, const uint8_t *ref, ptrdiff_t sr, int width, int height, int dx0, int dy0,
                                                     int n_frames, ptrdiff_t fs_src, ptrdiff_t fs_ref, int32_t *best8, int32_t *best16, int32_t *best32,
                                                     int32_t *best64, void *stream)
{
    return launch_pyramid_compact<OUT_BEST>(src, ss, ref, sr, width, height, dx0, dy0, n_frames, fs_src, fs_ref, best8, best16, best32, best64, stream);
}

extern "C" int hevcasm_sad_sweep_pyramid_packed_frames(const uint8_t *src, ptrdiff_t ss, const uint8_t *ref, ptrdiff_t sr, int width, int height, int dx0, int dy0,
                                                       int n_frames, ptrdiff_t fs_src, ptrdiff_t fs_ref, uint16_t *sad8, uint16_t *sad16, int32_t *sad32,
                                                       int32_t *sad64, void *stream)
{
    return launch_pyramid_compact<OUT_PACKED>(src, ss, ref, sr, width, height, dx0, dy0, n_frames, fs_src, fs_ref, sad8, sad16, sad32, sad64, stream);
}

extern "C" int hevcasm_ssd_batch(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb, int log2size, const int16_t *blk_xy, int n,
                                 int32_t *ssd, void *stream)
{
    if (log2size < 2 || log2size > 6 || n < 0) return HEVCASM_ERR_ARGUMENT;
    if (n == 0) return 0;
    HV_LAUNCH(ssd_list_kernel, (unsigned)((n + 7) / 8), 256, 0, stream, a, sa, b, sb, 1 << log2size, blk_xy, n, ssd);
    return 0;
}

extern "C" int hevcasm_ssd_frames(const uint8_t *a, ptrdiff_t sa, const uint8_t *b, ptrdiff_t sb, int width, int height, int log2size,
                                  int n_frames, ptrdiff_t fs_a, ptrdiff_t fs_b, int32_t *ssd, void *stream)
{
    if (log2size < 2 || log2size > 6 || n_frames < 0 || width < 0 || height < 0) return HEVCASM_ERR_ARGUMENT;
    SsdParams p;
    p.a = a, p.b = b, p.sa = sa, p.sb = sb, p.fs_a = fs_a, p.fs_b = fs_b, p.log2 = log2size;
    p.nbx = width >> log2size, p.nby = height >> log2size, p.out = ssd;
    if (p.nbx == 0 || p.nby == 0 || n_frames == 0) return 0;
    const int N = 1 << log2size, th = N > ssdk::TH ? N : ssdk::TH;
    const dim3 grid((p.nbx * N + ssdk::TW - 1) / ssdk::TW, (p.nby * N + th - 1) / th, n_frames);
    // the 128-bit path needs 16-byte aligned rows and a block grid that ends on a 16-sample boundary (no over-read)
    const bool aligned = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)sa | (uintptr_t)sb | (uintptr_t)fs_a | (uintptr_t)fs_b |
                           (uintptr_t)(p.nbx * (1 << log2size))) & 15) == 0;
    if (aligned) HV_LAUNCH(ssd_frames_kernel<true>, grid, ssdk::NT, 0, stream, p);
    else HV_LAUNCH(ssd_frames_kernel<false>, grid, ssdk::NT, 0, stream, p);
    return 0;
}
