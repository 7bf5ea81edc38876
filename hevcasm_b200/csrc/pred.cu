// hevcasm_b200 - HEVC non-weighted inter prediction (luma 8-tap, chroma 4-tap; uni and bi) for sm_100a.
//
// Reference semantics (kupix/hevcasm pred_inter.c): coefficient tables :53-83, the generic FIR :90-138, the uni
// recipes :141-228 (copy / H / V / HV) and the bi recipe :490-530 (H exact -> V >>6 stored int16 -> (A+B+64)>>7).
//
// Design.  The first pass of the reference's separable filter is EXACT (shift 0, result fits int16), so the 2-D sum
// sum_y sum_x cy[y]*cx[x]*p is the same integer in either order; only the final rounding shift is applied to it.  This
// kernel therefore filters VERTICALLY FIRST on the 8-bit samples and horizontally second on the int16 intermediate -
// the order in which both passes map onto the byte/halfword dot-product instructions without data reshuffling in
// shared memory:
//   * vertical pass on bytes: a thread owns 4 columns x 8 rows.  The 4 bytes a dp4a consumes must run along the tap
//     direction, so the rows of a column are gathered with PRMT (6 per row offset for 4 columns) and each gathered word
//     feeds IDP.4A.U8.S8 against 4 packed coefficients: 2 IDP per luma sample, 1 per chroma sample;
//   * the exact int16 intermediate goes to shared memory row-major (the reference's stack `intermediate`);
//   * horizontal pass on int16: a thread owns 8 adjacent outputs of a row, loads 16 intermediates with two 128-bit
//     shared loads and uses IDP.2A.LO.S16.S8 on adjacent pairs (odd-aligned pairs are one PRMT, shared by 4 outputs);
//   * H-only positions use IDP.4A on byte windows cut with a funnel shift; V-only positions the vertical pass alone;
//     full-pel positions are a copy.
// Everything is staged once per CTA tile (source tile + filter halo) in shared memory; HBM traffic is the algorithmic
// 1 B in + 1 B out per sample (2 + 1 for bi).
#include "common.cuh"

#include <cstdlib>

namespace hv {
namespace ip {

// H.265 8.5.3.3.3 interpolation filters (equal to reference pred_inter.c:57-63, :69-79)
__constant__ int8_t c_luma[4][8] = {{0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1}};
__constant__ int8_t c_chroma[8][4] = {{0, 64, 0, 0},   {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4},
                                      {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2}};

enum Mode { COPY = 0, H_ONLY = 1, V_ONLY = 2, HV = 3, RUNTIME = -1 };

template <int TAPS>
struct Coefs {
    int p4[TAPS / 4];  // 4 x s8 per word (dp4a operand)
    int p2[TAPS / 2];  // 2 x s8 in the low half (dp2a.lo operand)
    __device__ __forceinline__ void load(int frac)
    {
        int c[TAPS];
#pragma unroll
        for (int k = 0; k < TAPS; ++k) c[k] = TAPS == 8 ? (int)c_luma[frac & 3][k] : (int)c_chroma[frac & 7][k];
#pragma unroll
        for (int g = 0; g < TAPS / 4; ++g)
            p4[g] = (c[4 * g] & 0xff) | ((c[4 * g + 1] & 0xff) << 8) | ((c[4 * g + 2] & 0xff) << 16) | ((c[4 * g + 3] & 0xff) << 24);
#pragma unroll
        for (int g = 0; g < TAPS / 2; ++g) p2[g] = (c[2 * g] & 0xff) | ((c[2 * g + 1] & 0xff) << 8);
    }
};

// ---- vertical pass on bytes: 4 columns x R rows of exact tap sums -------------------------------------------
// s -> staged word holding the 4 columns at the first needed row (output row 0 needs rows 0 .. TAPS-1)
template <int TAPS, int R>
__device__ __forceinline__ void vpass_bytes(const uint32_t *s, int pitch_words, const int (&cy4)[TAPS / 4], int (&out)[R][4])
{
    constexpr int NS = R + TAPS - 1;
    uint32_t S[NS];
#pragma unroll
    for (int r = 0; r < NS; ++r) S[r] = s[r * pitch_words];
    uint32_t U01[NS - 1], U23[NS - 1];  // (col0 row r, col0 row r+1, col1 row r, col1 row r+1) and the same for columns 2, 3
#pragma unroll
    for (int r = 0; r < NS - 1; ++r) {
        U01[r] = __byte_perm(S[r], S[r + 1], 0x5140);
        U23[r] = __byte_perm(S[r], S[r + 1], 0x7362);
    }
    uint32_t T[NS - 3][4];  // T[j][c] = column c, rows j .. j+3
#pragma unroll
    for (int j = 0; j < NS - 3; ++j) {
        T[j][0] = __byte_perm(U01[j], U01[j + 2], 0x5410);
        T[j][1] = __byte_perm(U01[j], U01[j + 2], 0x7632);
        T[j][2] = __byte_perm(U23[j], U23[j + 2], 0x5410);
        T[j][3] = __byte_perm(U23[j], U23[j + 2], 0x7632);
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int a = 0;
#pragma unroll
            for (int g = 0; g < TAPS / 4; ++g) a = dp4a_us(T[r + 4 * g][c], cy4[g], a);
            out[r][c] = a;
        }
}

// ---- horizontal pass on the int16 intermediate: 8 adjacent outputs -----------------------------------------
// m -> 16 intermediates (8 words, 16-byte aligned) starting 4 columns left of the first output
template <int TAPS>
__device__ __forceinline__ void hpass_mid(const uint32_t *m, const int (&cx2)[TAPS / 2], int round, int (&out)[8])
{
    uint32_t M[8];
    const uint4 a = *reinterpret_cast<const uint4 *>(m), b = *reinterpret_cast<const uint4 *>(m + 4);
    M[0] = a.x, M[1] = a.y, M[2] = a.z, M[3] = a.w, M[4] = b.x, M[5] = b.y, M[6] = b.z, M[7] = b.w;
    constexpr int OFF = 4 - (TAPS / 2 - 1);  // first tap of output i sits at intermediate i + OFF
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int acc = round;
#pragma unroll
        for (int g = 0; g < TAPS / 2; ++g) {
            const int s = i + OFF + 2 * g;
            const uint32_t pair = (s & 1) ? __byte_perm(M[s >> 1], M[(s >> 1) + 1], 0x5432) : M[s >> 1];
            acc = dp2a_lo(pair, cx2[g], acc);
        }
        out[i] = acc;
    }
}

// ---- horizontal pass on bytes (H-only positions): 8 adjacent outputs ------------------------------------------
// s -> 4 staged words starting 4 bytes left of the first output
template <int TAPS>
__device__ __forceinline__ void hpass_bytes(const uint32_t *s, const int (&cx4)[TAPS / 4], int round, int (&out)[8])
{
    uint32_t W[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) W[k] = s[k];
    constexpr int LEFT = TAPS / 2 - 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int acc = round;
#pragma unroll
        for (int g = 0; g < TAPS / 4; ++g) {
            const int o = i - LEFT + 4 * g + 4;
            const uint32_t win = (o & 3) ? shr_bytes(W[o >> 2], W[(o >> 2) + 1], o & 3) : W[o >> 2];
            acc = dp4a_us(win, cx4[g], acc);
        }
        out[i] = acc;
    }
}

// ---- stores that never touch a byte outside [p, p + nvalid) ---------------------------------------------------
__device__ __forceinline__ void store4(uint8_t *p, uint32_t v, int nvalid)
{
    if (nvalid >= 4 && ((uintptr_t)p & 3) == 0) {
        *reinterpret_cast<uint32_t *>(p) = v;
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (i < nvalid) p[i] = (uint8_t)(v >> (8 * i));
}
__device__ __forceinline__ void store8(uint8_t *p, uint32_t lo, uint32_t hi, int nvalid)
{
    if (nvalid >= 8 && ((uintptr_t)p & 7) == 0) {
        *reinterpret_cast<uint2 *>(p) = make_uint2(lo, hi);
        return;
    }
    store4(p, lo, nvalid);
    store4(p + 4, hi, nvalid - 4);
}

struct PredParams {
    uint8_t *dst;
    const uint8_t *ref0, *ref1;
    ptrdiff_t sd, sr, fs_dst, fs_ref;
    int width, height;           // plane form: tile grid over width x height, blockIdx.z = frame
    int xf0, yf0, xf1, yf1;
    const int16_t *pus;          // list form: one CTA per descriptor
    int n_pu;
};

constexpr int NT = 128;

template <int TAPS, int TW, int TH>
struct Geom {
    static constexpr int LEFT = TAPS / 2 - 1, RIGHT = TAPS / 2;
    static constexpr int SP = TW / 4 + 4;             // staged source pitch (words): 4 left + TW + up to 8 right, + slack for partial items
    static constexpr int SROWS = TH + TAPS - 1;
    static constexpr int MP = (TW + 16) / 2;          // intermediate pitch (words): (TW + 8) int16 + 8 slack, keeps rows 16-byte aligned
    static constexpr int SRC_WORDS = SP * SROWS, MID_WORDS = MP * TH;
    static constexpr int SMEM_BYTES = (SRC_WORDS + MID_WORDS) * 4;
    static constexpr int R = 8;                        // rows per vertical-pass item
};

// One reference of one tile through the separable filter.  On return:
//   HV / bi : `mid` holds the exact vertically filtered intermediate (columns x = -4 .. w+3 at index x+4)
//   others  : results have been written to dst
// Returns nothing; the caller runs the horizontal pass (it differs between uni and bi).
template <int TAPS, int TW, int TH>
__device__ __forceinline__ void stage_source(uint32_t *src_s, const uint8_t *ref, ptrdiff_t sr, int w, int h, bool need_h, bool need_v, int tid)
{
    using G = Geom<TAPS, TW, TH>;
    const int xoff = need_h ? 4 : 0, top = need_v ? G::LEFT : 0;
    const int bytes = w + xoff + (need_h ? G::RIGHT : 0);
    const int rows = h + (need_v ? TAPS - 1 : 0);
    stage_tile_u8(src_s, G::SP, ref - (ptrdiff_t)top * sr - xoff, sr, (bytes + 3) >> 2, rows, tid, NT);
}

// vertical pass of a whole tile into `mid` (exact sums as int16); quads cover staged columns 0 .. ncols-1
template <int TAPS, int TW, int TH>
__device__ __forceinline__ void vertical_to_mid(uint32_t *mid, const uint32_t *src_s, const Coefs<TAPS> &cy, int ncols, int h, int tid)
{
    using G = Geom<TAPS, TW, TH>;
    const int nq = (ncols + 3) >> 2, nrg = (h + G::R - 1) / G::R;
    for (int id = tid; id < nq * nrg; id += NT) {
        const int q = id % nq, rg = id / nq;
        int v[G::R][4];
        vpass_bytes<TAPS, G::R>(src_s + rg * G::R * G::SP + q, G::SP, cy.p4, v);
#pragma unroll
        for (int r = 0; r < G::R; ++r)
            *reinterpret_cast<uint2 *>(mid + (rg * G::R + r) * G::MP + 2 * q) = make_uint2(pack16(v[r][0], v[r][1]), pack16(v[r][2], v[r][3]));
    }
}

template <int TAPS, int TW, int TH, bool BI, int FIXED_MODE>
__global__ void __launch_bounds__(NT) pred_kernel(PredParams p)
{
    using G = Geom<TAPS, TW, TH>;
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *src_s = smem, *mid = smem + G::SRC_WORDS;
    const int tid = threadIdx.x;
    constexpr int FB = TAPS == 8 ? 2 : 3, FM = (1 << FB) - 1;

    // ---- locate the tile
    uint8_t *dst;
    const uint8_t *ref[2];
    int w, h, xf[2], yf[2];
    if (p.pus) {
        const int16_t *d = p.pus + (size_t)blockIdx.x * (BI ? 8 : 6);
        const int x = d[0], y = d[1];
        w = d[2], h = d[3];
        if (w <= 0 || h <= 0 || w > TW || h > TH) return;
        dst = p.dst + (ptrdiff_t)y * p.sd + x;
        ref[0] = p.ref0 + (ptrdiff_t)(y + (d[5] >> FB)) * p.sr + (x + (d[4] >> FB));
        xf[0] = d[4] & FM, yf[0] = d[5] & FM;
        if (BI) {
            ref[1] = p.ref1 + (ptrdiff_t)(y + (d[7] >> FB)) * p.sr + (x + (d[6] >> FB));
            xf[1] = d[6] & FM, yf[1] = d[7] & FM;
        }
    } else {
        const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, f = blockIdx.z;
        w = min(TW, p.width - x0), h = min(TH, p.height - y0);
        dst = p.dst + f * p.fs_dst + (ptrdiff_t)y0 * p.sd + x0;
        const ptrdiff_t o = f * p.fs_ref + (ptrdiff_t)y0 * p.sr + x0;
        ref[0] = p.ref0 + o, xf[0] = p.xf0, yf[0] = p.yf0;
        if (BI) ref[1] = p.ref1 + o, xf[1] = p.xf1, yf[1] = p.yf1;
    }

    if (!BI) {
        const int mode = FIXED_MODE != RUNTIME ? FIXED_MODE : ((xf[0] ? 1 : 0) | (yf[0] ? 2 : 0));
        const bool need_h = mode & 1, need_v = mode & 2;
        stage_source<TAPS, TW, TH>(src_s, ref[0], p.sr, w, h, need_h, need_v, tid);
        __syncthreads();
        if (mode == COPY) {
            const int nw = (w + 3) >> 2;
            for (int id = tid; id < nw * h; id += NT) {
                const int q = id % nw, y = id / nw;
                store4(dst + (ptrdiff_t)y * p.sd + 4 * q, src_s[y * G::SP + q], w - 4 * q);
            }
        } else if (mode == H_ONLY) {
            Coefs<TAPS> cx;
            cx.load(xf[0]);
            const int nj = (w + 7) >> 3;
            for (int id = tid; id < nj * h; id += NT) {
                const int jj = id % nj, y = id / nj;
                int o[8];
                hpass_bytes<TAPS>(src_s + y * G::SP + 2 * jj, cx.p4, 32, o);
                store8(dst + (ptrdiff_t)y * p.sd + 8 * jj, pack_sat_u8(o[0] >> 6, o[1] >> 6, o[2] >> 6, o[3] >> 6),
                       pack_sat_u8(o[4] >> 6, o[5] >> 6, o[6] >> 6, o[7] >> 6), w - 8 * jj);
            }
        } else if (mode == V_ONLY) {
            Coefs<TAPS> cy;
            cy.load(yf[0]);
            const int nq = (w + 3) >> 2, nrg = (h + G::R - 1) / G::R;
            for (int id = tid; id < nq * nrg; id += NT) {
                const int q = id % nq, rg = id / nq;
                int v[G::R][4];
                vpass_bytes<TAPS, G::R>(src_s + rg * G::R * G::SP + q, G::SP, cy.p4, v);
#pragma unroll
                for (int r = 0; r < G::R; ++r) {
                    const int y = rg * G::R + r;
                    if (y < h)
                        store4(dst + (ptrdiff_t)y * p.sd + 4 * q, pack_sat_u8((v[r][0] + 32) >> 6, (v[r][1] + 32) >> 6, (v[r][2] + 32) >> 6, (v[r][3] + 32) >> 6),
                               w - 4 * q);
                }
            }
        } else {
            Coefs<TAPS> cx, cy;
            cx.load(xf[0]);
            cy.load(yf[0]);
            vertical_to_mid<TAPS, TW, TH>(mid, src_s, cy, w + 4 + G::RIGHT, h, tid);
            __syncthreads();
            const int nj = (w + 7) >> 3;
            for (int id = tid; id < nj * h; id += NT) {
                const int jj = id % nj, y = id / nj;
                int o[8];
                hpass_mid<TAPS>(mid + y * G::MP + 4 * jj, cx.p2, 2048, o);
                store8(dst + (ptrdiff_t)y * p.sd + 8 * jj, pack_sat_u8(o[0] >> 12, o[1] >> 12, o[2] >> 12, o[3] >> 12),
                       pack_sat_u8(o[4] >> 12, o[5] >> 12, o[6] >> 12, o[7] >> 12), w - 8 * jj);
            }
        }
    } else {
        // bi: per reference the exact 2-D sum >> 6, truncated to int16 (reference pred_inter.c:504-527 always runs both passes,
        // a zero fraction being the {64} filter), then (A + B + 64) >> 7 clipped
        constexpr int ITEMS = (TW / 8) * TH / NT;
        uint32_t va[ITEMS][4];  // first reference's values, packed int16 pairs
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            Coefs<TAPS> cx, cy;
            cx.load(xf[r]);
            cy.load(yf[r]);
            if (r) __syncthreads();  // everyone is done reading the first reference's intermediate
            stage_source<TAPS, TW, TH>(src_s, ref[r], p.sr, w, h, true, true, tid);
            __syncthreads();
            vertical_to_mid<TAPS, TW, TH>(mid, src_s, cy, w + 4 + G::RIGHT, h, tid);
            __syncthreads();
            const int nj = (w + 7) >> 3;
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const int id = tid + k * NT;
                const int jj = id % nj, y = id / nj;
                if (id < nj * h) {
                    int o[8];
                    hpass_mid<TAPS>(mid + y * G::MP + 4 * jj, cx.p2, 0, o);
                    if (r == 0) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) va[k][i] = pack16(o[2 * i] >> 6, o[2 * i + 1] >> 6);  // truncating int16 store
                    } else {
                        int s[8];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            s[2 * i] = ((int)(short)(va[k][i] & 0xffff) + (int)(short)(o[2 * i] >> 6) + 64) >> 7;
                            s[2 * i + 1] = (((int)va[k][i] >> 16) + (int)(short)(o[2 * i + 1] >> 6) + 64) >> 7;
                        }
                        store8(dst + (ptrdiff_t)y * p.sd + 8 * jj, pack_sat_u8(s[0], s[1], s[2], s[3]), pack_sat_u8(s[4], s[5], s[6], s[7]), w - 8 * jj);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ plane form, fast path
//
// Whole planes with 16-byte aligned reference rows (the normal frame store).  Same arithmetic building blocks; what
// changes is everything around them: the tile and its halo are staged with clamped 128-bit loads (no per-word address
// arithmetic, no run-time divisions), every loop has a compile-time trip count (edge tiles compute the full tile and
// mask the stores), and stores are 64-bit when the destination rows are 8-byte aligned.
//   tile = 128 x 32 outputs, 128 threads.  Staged row: 16-byte chunks from 16 bytes left of the tile when a horizontal
//   pass needs the left halo (block column 0 at byte XPAD = 16), from the tile's own column 0 otherwise.
template <int TAPS, int MODE, bool BI>
struct FastGeom {
    static constexpr int TW = 128, TH = 32, R = 8;
    static constexpr bool NEED_H = BI || (MODE & 1), NEED_V = BI || (MODE & 2);
    static constexpr int LEFT = TAPS / 2 - 1, RIGHT = TAPS / 2;
    static constexpr int XPAD = NEED_H ? 16 : 0;
    static constexpr int CHUNKS = NEED_H ? 10 : 8;                 // 16-byte chunks per staged row
    static constexpr int SPW = CHUNKS * 4;                         // staged pitch in words
    static constexpr int SROWS = TH + (NEED_V ? TAPS - 1 : 0);
    static constexpr int TOP = NEED_V ? LEFT : 0;
    static constexpr int MP = (TW + 16) / 2;                       // intermediate pitch in words (TW + 8 int16, + slack)
    static constexpr int SRC_WORDS = SPW * SROWS, MID_WORDS = (NEED_H && NEED_V) ? MP * TH : 0;
    static constexpr int SMEM_BYTES = (SRC_WORDS + MID_WORDS) * 4;
    static constexpr int Q0 = NEED_H ? 3 : 0;                      // first staged quad the vertical pass visits (x = -4 or 0)
    static constexpr int NQ = NEED_H ? (TW + 8) / 4 : TW / 4;      // quads per row group
};

template <class G>
__device__ __forceinline__ void stage_fast(uint32_t *src_s, const uint8_t *tile /* ref at the tile's (0,0) */, ptrdiff_t sr, int w, int h, int tid)
{
    const uint8_t *base = tile - (ptrdiff_t)G::TOP * sr - G::XPAD;
    const int last_row = h + (G::NEED_V ? G::LEFT + G::RIGHT : 0) - 1;                       // last row anything reads
    const int last_ch = (G::XPAD + w + (G::NEED_H ? G::RIGHT : 0) - 1) >> 4;                  // last chunk anything reads
    constexpr int TOTAL = G::CHUNKS * G::SROWS, ITERS = (TOTAL + NT - 1) / NT;
    int4 v[ITERS];
#pragma unroll
    for (int k = 0; k < ITERS; ++k) {
        const int idx = min(tid + k * NT, TOTAL - 1), row = idx / G::CHUNKS, ch = idx - row * G::CHUNKS;
        v[k] = __ldg(reinterpret_cast<const int4 *>(base + (ptrdiff_t)min(row, last_row) * sr) + min(ch, last_ch));
    }
#pragma unroll
    for (int k = 0; k < ITERS; ++k) {
        const int idx = tid + k * NT;
        if (idx < TOTAL) reinterpret_cast<int4 *>(src_s)[idx] = v[k];   // rows are CHUNKS int4 long: the staged tile is dense
    }
}

template <bool DST8>
__device__ __forceinline__ void put8(uint8_t *p, uint32_t lo, uint32_t hi, int nvalid)
{
    if (DST8) {
        if (nvalid >= 8) *reinterpret_cast<uint2 *>(p) = make_uint2(lo, hi);
        else if (nvalid > 0) store8(p, lo, hi, nvalid);
    } else if (nvalid > 0) {
        store8(p, lo, hi, nvalid);
    }
}

// vertical pass of the whole staged tile into `mid` (exact int16 sums), fast-path geometry
template <int TAPS, class G>
__device__ __forceinline__ void vertical_to_mid_fast(uint32_t *mid, const uint32_t *src_s, const Coefs<TAPS> &cy, int tid)
{
    constexpr int ITEMS = G::NQ * (G::TH / G::R), ITERS = (ITEMS + NT - 1) / NT;
#pragma unroll 1
    for (int k = 0; k < ITERS; ++k) {
        const int id = tid + k * NT;
        if (id >= ITEMS) break;
        const int q = id % G::NQ, rg = id / G::NQ;
        int v[G::R][4];
        vpass_bytes<TAPS, G::R>(src_s + rg * G::R * G::SPW + G::Q0 + q, G::SPW, cy.p4, v);
#pragma unroll
        for (int r = 0; r < G::R; ++r)
            *reinterpret_cast<uint2 *>(mid + (rg * G::R + r) * G::MP + 2 * q) = make_uint2(pack16(v[r][0], v[r][1]), pack16(v[r][2], v[r][3]));
    }
}

template <int TAPS, int MODE, bool BI, bool DST8>
__global__ void __launch_bounds__(NT) pred_plane_fast_kernel(PredParams p)
{
    using G = FastGeom<TAPS, MODE, BI>;
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *src_s = smem, *mid = smem + G::SRC_WORDS;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * G::TW, y0 = blockIdx.y * G::TH, f = blockIdx.z;
    const int w = min(G::TW, p.width - x0), h = min(G::TH, p.height - y0);
    uint8_t *dst = p.dst + f * p.fs_dst + (ptrdiff_t)y0 * p.sd + x0;
    const ptrdiff_t ro = f * p.fs_ref + (ptrdiff_t)y0 * p.sr + x0;
    constexpr int NJ = G::TW / 8, HITEMS = NJ * G::TH / NT;  // 8-wide output groups per row; groups per thread

    if (!BI && MODE == COPY) {
        // 2 x (LDG.128 -> two 64-bit stores) per thread, straight through registers
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int idx = tid + k * NT, row = idx >> 3, ch = idx & 7;
            if (row < h && ch * 16 < w) {
                const int4 v = __ldg(reinterpret_cast<const int4 *>(p.ref0 + ro + (ptrdiff_t)row * p.sr) + ch);  // may read up to 15 bytes right of w (aligned chunk)
                uint8_t *d = dst + (ptrdiff_t)row * p.sd + ch * 16;
                put8<DST8>(d, (uint32_t)v.x, (uint32_t)v.y, w - ch * 16);
                put8<DST8>(d + 8, (uint32_t)v.z, (uint32_t)v.w, w - ch * 16 - 8);
            }
        }
        return;
    }

    if (!BI) {
        stage_fast<G>(src_s, p.ref0 + ro, p.sr, w, h, tid);
        __syncthreads();
        if (MODE == H_ONLY) {
            Coefs<TAPS> cx;
            cx.load(p.xf0);
#pragma unroll
            for (int k = 0; k < HITEMS; ++k) {
                const int id = tid + k * NT, jj = id % NJ, y = id / NJ;
                int o[8];
                hpass_bytes<TAPS>(src_s + y * G::SPW + 3 + 2 * jj, cx.p4, 32, o);
                if (y < h)
                    put8<DST8>(dst + (ptrdiff_t)y * p.sd + 8 * jj, pack_sat_u8(o[0] >> 6, o[1] >> 6, o[2] >> 6, o[3] >> 6),
                               pack_sat_u8(o[4] >> 6, o[5] >> 6, o[6] >> 6, o[7] >> 6), w - 8 * jj);
            }
        } else if (MODE == V_ONLY) {
            Coefs<TAPS> cy;
            cy.load(p.yf0);
            const int q = tid % G::NQ, rg = tid / G::NQ;  // exactly one item per thread
            int v[G::R][4];
            vpass_bytes<TAPS, G::R>(src_s + rg * G::R * G::SPW + q, G::SPW, cy.p4, v);
#pragma unroll
            for (int r = 0; r < G::R; ++r) {
                const int y = rg * G::R + r;
                if (y < h && 4 * q < w)
                    store4(dst + (ptrdiff_t)y * p.sd + 4 * q, pack_sat_u8((v[r][0] + 32) >> 6, (v[r][1] + 32) >> 6, (v[r][2] + 32) >> 6, (v[r][3] + 32) >> 6),
                           w - 4 * q);
            }
        } else {
            Coefs<TAPS> cx, cy;
            cx.load(p.xf0);
            cy.load(p.yf0);
            vertical_to_mid_fast<TAPS, G>(mid, src_s, cy, tid);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < HITEMS; ++k) {
                const int id = tid + k * NT, jj = id % NJ, y = id / NJ;
                int o[8];
                hpass_mid<TAPS>(mid + y * G::MP + 4 * jj, cx.p2, 2048, o);
                if (y < h)
                    put8<DST8>(dst + (ptrdiff_t)y * p.sd + 8 * jj, pack_sat_u8(o[0] >> 12, o[1] >> 12, o[2] >> 12, o[3] >> 12),
                               pack_sat_u8(o[4] >> 12, o[5] >> 12, o[6] >> 12, o[7] >> 12), w - 8 * jj);
            }
        }
    } else {
        uint32_t va[HITEMS][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            Coefs<TAPS> cx, cy;
            cx.load(r ? p.xf1 : p.xf0);
            cy.load(r ? p.yf1 : p.yf0);
            if (r) __syncthreads();
            stage_fast<G>(src_s, (r ? p.ref1 : p.ref0) + ro, p.sr, w, h, tid);
            __syncthreads();
            vertical_to_mid_fast<TAPS, G>(mid, src_s, cy, tid);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < HITEMS; ++k) {
                const int id = tid + k * NT, jj = id % NJ, y = id / NJ;
                int o[8];
                hpass_mid<TAPS>(mid + y * G::MP + 4 * jj, cx.p2, 0, o);
                if (r == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) va[k][i] = pack16(o[2 * i] >> 6, o[2 * i + 1] >> 6);
                } else {
                    int s[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        s[2 * i] = ((int)(short)(va[k][i] & 0xffff) + (int)(short)(o[2 * i] >> 6) + 64) >> 7;
                        s[2 * i + 1] = (((int)va[k][i] >> 16) + (int)(short)(o[2 * i + 1] >> 6) + 64) >> 7;
                    }
                    if (y < h) put8<DST8>(dst + (ptrdiff_t)y * p.sd + 8 * jj, pack_sat_u8(s[0], s[1], s[2], s[3]), pack_sat_u8(s[4], s[5], s[6], s[7]), w - 8 * jj);
                }
            }
        }
    }
}

template <int TAPS, int MODE, bool BI>
int launch_plane_fast(const PredParams &p, dim3 grid, bool dst8, void *stream)
{
    using G = FastGeom<TAPS, MODE, BI>;
    return dst8 ? launch(pred_plane_fast_kernel<TAPS, MODE, BI, true>, grid, dim3(NT), (size_t)G::SMEM_BYTES, stream, p)
                : launch(pred_plane_fast_kernel<TAPS, MODE, BI, false>, grid, dim3(NT), (size_t)G::SMEM_BYTES, stream, p);
}

template <int TAPS>
int launch_uni_planes_fast(const PredParams &p, dim3 grid, int mode, bool dst8, void *stream)
{
    switch (mode) {
        case COPY: return launch_plane_fast<TAPS, COPY, false>(p, grid, dst8, stream);
        case H_ONLY: return launch_plane_fast<TAPS, H_ONLY, false>(p, grid, dst8, stream);
        case V_ONLY: return launch_plane_fast<TAPS, V_ONLY, false>(p, grid, dst8, stream);
        default: return launch_plane_fast<TAPS, HV, false>(p, grid, dst8, stream);
    }
}

template <int TAPS, int TW, int TH, bool BI, int FIXED_MODE>
int launch_pred(const PredParams &p, dim3 grid, void *stream)
{
    using G = Geom<TAPS, TW, TH>;
    auto kern = pred_kernel<TAPS, TW, TH, BI, FIXED_MODE>;
    if (G::SMEM_BYTES > 48 * 1024) {
        const int e = set_max_smem(kern, G::SMEM_BYTES);
        if (e) return e;
    }
    return launch(kern, grid, dim3(NT), (size_t)G::SMEM_BYTES, stream, p);
}

// plane form: tile 128 x 32;  list form: tile 64 x 64 (the largest prediction unit)
constexpr int PTW = 128, PTH = 32, LTW = 64, LTH = 64;

template <int TAPS>
int launch_uni_planes(const PredParams &p, dim3 grid, int mode, void *stream)
{
    switch (mode) {
        case COPY: return launch_pred<TAPS, PTW, PTH, false, COPY>(p, grid, stream);
        case H_ONLY: return launch_pred<TAPS, PTW, PTH, false, H_ONLY>(p, grid, stream);
        case V_ONLY: return launch_pred<TAPS, PTW, PTH, false, V_ONLY>(p, grid, stream);
        default: return launch_pred<TAPS, PTW, PTH, false, HV>(p, grid, stream);
    }
}

}  // namespace ip
}  // namespace hv

using namespace hv;
using namespace hv::ip;

static bool frac_ok(int taps, int f) { return f >= 0 && f < (taps == 8 ? 4 : 8); }

// the fast plane kernels issue aligned 128-bit loads on the reference rows (and may therefore touch up to 16 bytes left and
// 15 bytes right of the reference's own footprint - hevcasm_batch.h documents the padding this needs)
static bool planes_fast_ok(const uint8_t *ref0, const uint8_t *ref1, ptrdiff_t sr, ptrdiff_t fs_ref, int n_frames)
{
    if (getenv("HEVCASM_PRED_GENERIC")) return false;
    uintptr_t m = (uintptr_t)ref0 | (uintptr_t)sr | (ref1 ? (uintptr_t)ref1 : 0);
    if (n_frames > 1) m |= (uintptr_t)fs_ref;
    return (m & 15) == 0;
}
static bool aligned8(const uint8_t *dst, ptrdiff_t sd, ptrdiff_t fs_dst, int n_frames)
{
    uintptr_t m = (uintptr_t)dst | (uintptr_t)sd;
    if (n_frames > 1) m |= (uintptr_t)fs_dst;
    return (m & 7) == 0;
}

extern "C" int hevcasm_pred_uni_frames(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref, ptrdiff_t sr, int width, int height, int taps, int xFrac, int yFrac,
                                       int n_frames, ptrdiff_t fs_dst, ptrdiff_t fs_ref, void *stream)
{
    if ((taps != 8 && taps != 4) || !frac_ok(taps, xFrac) || !frac_ok(taps, yFrac) || width < 0 || height < 0 || n_frames < 0) return HEVCASM_ERR_ARGUMENT;
    if (width == 0 || height == 0 || n_frames == 0) return 0;
    PredParams p{};
    p.dst = dst, p.ref0 = ref, p.sd = sd, p.sr = sr, p.fs_dst = fs_dst, p.fs_ref = fs_ref, p.width = width, p.height = height;
    p.xf0 = xFrac, p.yf0 = yFrac;
    const dim3 grid((width + PTW - 1) / PTW, (height + PTH - 1) / PTH, n_frames);
    const int mode = (xFrac ? 1 : 0) | (yFrac ? 2 : 0);
    if (planes_fast_ok(ref, nullptr, sr, fs_ref, n_frames)) {
        const bool dst8 = aligned8(dst, sd, fs_dst, n_frames);
        return taps == 8 ? launch_uni_planes_fast<8>(p, grid, mode, dst8, stream) : launch_uni_planes_fast<4>(p, grid, mode, dst8, stream);
    }
    return taps == 8 ? launch_uni_planes<8>(p, grid, mode, stream) : launch_uni_planes<4>(p, grid, mode, stream);
}

extern "C" int hevcasm_pred_bi_frames(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref0, const uint8_t *ref1, ptrdiff_t sr, int width, int height, int taps,
                                      int xFrac0, int yFrac0, int xFrac1, int yFrac1, int n_frames, ptrdiff_t fs_dst, ptrdiff_t fs_ref, void *stream)
{
    if ((taps != 8 && taps != 4) || !frac_ok(taps, xFrac0) || !frac_ok(taps, yFrac0) || !frac_ok(taps, xFrac1) || !frac_ok(taps, yFrac1) || width < 0 ||
        height < 0 || n_frames < 0)
        return HEVCASM_ERR_ARGUMENT;
    if (width == 0 || height == 0 || n_frames == 0) return 0;
    PredParams p{};
    p.dst = dst, p.ref0 = ref0, p.ref1 = ref1, p.sd = sd, p.sr = sr, p.fs_dst = fs_dst, p.fs_ref = fs_ref, p.width = width, p.height = height;
    p.xf0 = xFrac0, p.yf0 = yFrac0, p.xf1 = xFrac1, p.yf1 = yFrac1;
    const dim3 grid((width + PTW - 1) / PTW, (height + PTH - 1) / PTH, n_frames);
    if (planes_fast_ok(ref0, ref1, sr, fs_ref, n_frames)) {
        const bool dst8 = aligned8(dst, sd, fs_dst, n_frames);
        return taps == 8 ? launch_plane_fast<8, HV, true>(p, grid, dst8, stream) : launch_plane_fast<4, HV, true>(p, grid, dst8, stream);
    }
    return taps == 8 ? launch_pred<8, PTW, PTH, true, RUNTIME>(p, grid, stream) : launch_pred<4, PTW, PTH, true, RUNTIME>(p, grid, stream);
}

extern "C" int hevcasm_pred_uni_batch(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref, ptrdiff_t sr, int taps, const int16_t *pus, int n_pu, void *stream)
{
    if ((taps != 8 && taps != 4) || n_pu < 0 || (n_pu > 0 && !pus)) return HEVCASM_ERR_ARGUMENT;
    if (n_pu == 0) return 0;
    PredParams p{};
    p.dst = dst, p.ref0 = ref, p.sd = sd, p.sr = sr, p.pus = pus, p.n_pu = n_pu;
    return taps == 8 ? launch_pred<8, LTW, LTH, false, RUNTIME>(p, dim3(n_pu), stream) : launch_pred<4, LTW, LTH, false, RUNTIME>(p, dim3(n_pu), stream);
}

extern "C" int hevcasm_pred_bi_batch(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref0, const uint8_t *ref1, ptrdiff_t sr, int taps, const int16_t *pus, int n_pu,
                                     void *stream)
{
    if ((taps != 8 && taps != 4) || n_pu < 0 || (n_pu > 0 && !pus)) return HEVCASM_ERR_ARGUMENT;
    if (n_pu == 0) return 0;
    PredParams p{};
    p.dst = dst, p.ref0 = ref0, p.ref1 = ref1, p.sd = sd, p.sr = sr, p.pus = pus, p.n_pu = n_pu;
    return taps == 8 ? launch_pred<8, LTW, LTH, true, RUNTIME>(p, dim3(n_pu), stream) : launch_pred<4, LTW, LTH, true, RUNTIME>(p, dim3(n_pu), stream);
}
