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
#include "tma.cuh"
#include "umma.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cstdio>

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
template <int TAPS, int NC>
__device__ __forceinline__ void hpass_mid(const uint32_t *m, const int (&cx2)[NC], int round, int (&out)[8])
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
    int desc_frame;              // list form: 1 = every descriptor ends with a frame index (planes fs_dst / fs_ref apart), 0 = one plane
    const uint8_t *lo[2], *hi[2];   // *_bounded plane forms: the readable bytes of each reference, [lo, hi) = its own footprint; null = unbounded
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
__device__ __forceinline__ void stage_source(uint32_t *src_s, const uint8_t *ref, ptrdiff_t sr, int w, int h, bool need_h, bool need_v, int tid,
                                             const uint8_t *lo = nullptr, const uint8_t *hi = nullptr)
{
    using G = Geom<TAPS, TW, TH>;
    const int xoff = need_h ? 4 : 0, top = need_v ? G::LEFT : 0;
    const int bytes = w + xoff + (need_h ? G::RIGHT : 0);
    const int rows = h + (need_v ? TAPS - 1 : 0);
    if (hi) stage_tile_u8_bounded(src_s, G::SP, ref - (ptrdiff_t)top * sr - xoff, sr, (bytes + 3) >> 2, rows, tid, NT, lo, hi);
    else stage_tile_u8(src_s, G::SP, ref - (ptrdiff_t)top * sr - xoff, sr, (bytes + 3) >> 2, rows, tid, NT);
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
        stage_source<TAPS, TW, TH>(src_s, ref[0], p.sr, w, h, need_h, need_v, tid, p.lo[0], p.hi[0]);
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
            stage_source<TAPS, TW, TH>(src_s, ref[r], p.sr, w, h, true, true, tid, p.lo[r], p.hi[r]);
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
// changes is everything around them (profiles/r01_pred.md: the first version spent 40 % of its issue slots outside the
// filters): the packed coefficients come from the host in the kernel parameters; the tile and its halo are staged with
// clamped 128-bit loads, three rows per warp instruction, pointers advanced incrementally; every loop has a compile-time
// trip count (edge tiles compute the full tile and mask the stores); stores are 64-bit when the destination rows are
// 8-byte aligned; and the tile is 128 x 64 so that the vertical pass (34 column quads x 8 row groups) fills its warps.
//   Staged row: 16-byte chunks from 16 bytes left of the tile when a horizontal pass needs the left halo (block column 0
//   at byte XPAD = 16), from the tile's own column 0 otherwise.
// Coefficients packed on the host the way the dot-product instructions consume them.  Besides the natural packings the
// fast path uses ROTATED ones: instead of re-aligning the data to the taps (a PRMT / funnel shift per operand) the taps
// are slid along aligned data words, padded with zeros - one more IDP for a misaligned output, no realignment at all.
struct PackedCoefs {
    int x4[2];      // horizontal taps, 4 per word, for dp4a on byte windows (H-only positions)
    int x2e[4];     // horizontal taps as pairs (c0,c1) (c2,c3) ..           : outputs whose first tap sits on an even intermediate
    int x2o[5];     // the same slid by one: (0,c0) (c1,c2) (c3,c4) (c5,c6) (c7,0) : outputs whose first tap sits on an odd one
    int y4s[4][3];  // vertical taps slid by s = 0..3 rows over row groups of four: byte j of word g holds tap 4g + j - s (or 0)
    int y2[4];      // vertical taps as pairs (c0,c1) (c2,c3) .. for dp2a on vertically packed int16 pairs (streaming kernel)
};
// vertical pass on bytes with rotated taps: 4 columns x 8 rows.  Rows are gathered once per aligned group of four
// (32 PRMT per item instead of 76); output row r = 4 k0 + s takes groups k0, k0+1 (and k0+2 when s != 0).
// s -> staged word of the 4 columns at the first needed row; 4*NG rows are read (the last ones may only meet zero taps).
template <int TAPS>
__device__ __forceinline__ void vpass_rot(const uint32_t *s, int pitch_words, const int (&y4s)[4][3], int (&out)[8][4])
{
    constexpr int NG = TAPS == 8 ? 4 : 3;  // row groups: rows 0..15 (8-tap) or 0..11 (4-tap)
    uint32_t T[NG][4];
#pragma unroll
    for (int k = 0; k < NG; ++k) {
        const uint32_t r0 = s[(4 * k) * pitch_words], r1 = s[(4 * k + 1) * pitch_words], r2 = s[(4 * k + 2) * pitch_words], r3 = s[(4 * k + 3) * pitch_words];
        const uint32_t a01 = __byte_perm(r0, r1, 0x5140), a23 = __byte_perm(r0, r1, 0x7362);
        const uint32_t b01 = __byte_perm(r2, r3, 0x5140), b23 = __byte_perm(r2, r3, 0x7362);
        T[k][0] = __byte_perm(a01, b01, 0x5410);
        T[k][1] = __byte_perm(a01, b01, 0x7632);
        T[k][2] = __byte_perm(a23, b23, 0x5410);
        T[k][3] = __byte_perm(a23, b23, 0x7632);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int k0 = r >> 2, sft = r & 3;
        constexpr int FULL = TAPS / 4;  // groups an aligned output needs
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int a = 0;
#pragma unroll
            for (int g = 0; g < FULL; ++g) a = dp4a_us(T[k0 + g][c], y4s[sft][g], a);
            if (sft) a = dp4a_us(T[k0 + FULL][c], y4s[sft][FULL], a);
            out[r][c] = a;
        }
    }
}

// horizontal pass on the int16 intermediate with rotated taps: 8 adjacent outputs, no realignment of the pairs
template <int TAPS>
__device__ __forceinline__ void hpass_rot(const uint32_t *m, const int (&x2e)[4], const int (&x2o)[5], int round, int (&out)[8])
{
    uint32_t M[8];
    const uint4 a = *reinterpret_cast<const uint4 *>(m), b = *reinterpret_cast<const uint4 *>(m + 4);
    M[0] = a.x, M[1] = a.y, M[2] = a.z, M[3] = a.w, M[4] = b.x, M[5] = b.y, M[6] = b.z, M[7] = b.w;
    constexpr int OFF = 4 - (TAPS / 2 - 1);  // odd for both tap counts
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int acc = round;
        if ((i + OFF) & 1) {  // first tap on an odd intermediate: slide the taps by one
#pragma unroll
            for (int g = 0; g <= TAPS / 2; ++g) acc = dp2a_lo(M[((i + OFF - 1) >> 1) + g], x2o[g], acc);
        } else {
#pragma unroll
            for (int g = 0; g < TAPS / 2; ++g) acc = dp2a_lo(M[((i + OFF) >> 1) + g], x2e[g], acc);
        }
        out[i] = acc;
    }
}

struct FastParams {
    PredParams p;
    PackedCoefs c[2];  // per reference
};

template <int TAPS, int MODE, bool BI>
struct FastGeom {
    static constexpr int TW = 128, TH = 64, R = 8;
    static constexpr bool NEED_H = BI || (MODE & 1), NEED_V = BI || (MODE & 2);
    static constexpr int LEFT = TAPS / 2 - 1, RIGHT = TAPS / 2;
    static constexpr int XPAD = NEED_H ? 16 : 0;
    static constexpr int CHUNKS = NEED_H ? 10 : 8;                 // 16-byte chunks per staged row
    static constexpr int SPW = CHUNKS * 4;                         // staged pitch in words
    static constexpr int SROWS = TH + (NEED_V ? TAPS - 1 : 0);
    static constexpr int TOP = NEED_V ? LEFT : 0;
    static constexpr int MP = (TW + 16) / 2;                       // intermediate pitch in words (TW + 8 int16, + slack)
    static constexpr int SRC_ROWS_ALLOC = NEED_V ? TH + (TAPS == 8 ? 8 : 4) : TH;   // vpass_rot reads whole groups of 4 rows (the extra ones meet zero taps)
    static constexpr int SRC_WORDS = SPW * SRC_ROWS_ALLOC, MID_WORDS = (NEED_H && NEED_V) ? MP * TH : 0;
    static constexpr int SMEM_BYTES = (SRC_WORDS + MID_WORDS) * 4;
    static constexpr int Q0 = NEED_H ? 3 : 0;                      // first staged quad the vertical pass visits (x = -4 or 0)
    static constexpr int NQ = NEED_H ? (TW + 8) / 4 : TW / 4;      // quads per row group
    static constexpr int ROWS_PER_WARP_LOAD = 32 / CHUNKS;         // 3 (10 chunks) or 4 (8 chunks) rows per warp instruction
    static constexpr int ROWS_PER_STEP = ROWS_PER_WARP_LOAD * (NT / 32);
    static constexpr int STEPS = (SROWS + ROWS_PER_STEP - 1) / ROWS_PER_STEP;
};

template <class G>
__device__ __forceinline__ void stage_fast(uint32_t *src_s, const uint8_t *tile /* ref at the tile's (0,0) */, ptrdiff_t sr, int w, int h, int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
    const int rl = lane / G::CHUNKS, ch = lane - rl * G::CHUNKS;   // row within the warp's group, chunk
    const bool active = rl < G::ROWS_PER_WARP_LOAD;
    const int last_row = h + (G::NEED_V ? G::LEFT + G::RIGHT : 0) - 1;                       // last row anything reads
    const int last_ch = (G::XPAD + w + (G::NEED_H ? G::RIGHT : 0) - 1) >> 4;                  // last chunk anything reads
    const uint8_t *base = tile - (ptrdiff_t)G::TOP * sr - G::XPAD + 16 * min(ch, last_ch);
    const int row0 = warp * G::ROWS_PER_WARP_LOAD + rl;
    int4 v[G::STEPS];
#pragma unroll
    for (int k = 0; k < G::STEPS; ++k) {
        const int row = row0 + k * G::ROWS_PER_STEP;
        if (active) v[k] = __ldg(reinterpret_cast<const int4 *>(base + (ptrdiff_t)min(row, last_row) * sr));
    }
    int4 *d = reinterpret_cast<int4 *>(src_s) + row0 * G::CHUNKS + ch;
#pragma unroll
    for (int k = 0; k < G::STEPS; ++k) {
        const int row = row0 + k * G::ROWS_PER_STEP;
        if (active && row < G::SROWS) d[k * G::ROWS_PER_STEP * G::CHUNKS] = v[k];
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
// (natural taps here: in the two-pass positions the FMA pipe (IDP) is the busier one, so the PRMT-heavy gather of
// vpass_bytes balances the two integer pipes better than vpass_rot - measured 1.36 vs 1.24 Tsamples/s, profiles/r01_pred.md)
template <int TAPS, class G>
__device__ __forceinline__ void vertical_to_mid_fast(uint32_t *mid, const uint32_t *src_s, const int (&y4s)[4][3], int tid)
{
    constexpr int ITEMS = G::NQ * (G::TH / G::R), ITERS = (ITEMS + NT - 1) / NT;
#pragma unroll 1
    for (int k = 0; k < ITERS; ++k) {
        const int id = tid + k * NT;
        if (id >= ITEMS) break;
        const int q = id % G::NQ, rg = id / G::NQ;
        int v[G::R][4];
        int cy4[TAPS / 4];
#pragma unroll
        for (int g = 0; g < TAPS / 4; ++g) cy4[g] = y4s[0][g];
        vpass_bytes<TAPS, G::R>(src_s + rg * G::R * G::SPW + G::Q0 + q, G::SPW, cy4, v);
        uint2 *m = reinterpret_cast<uint2 *>(mid + rg * G::R * G::MP + 2 * q);
#pragma unroll
        for (int r = 0; r < G::R; ++r) m[r * (G::MP / 2)] = make_uint2(pack16(v[r][0], v[r][1]), pack16(v[r][2], v[r][3]));
    }
}

template <int TAPS>
__device__ __forceinline__ void take(int (&d4)[TAPS / 4], const int (&s4)[2])
{
#pragma unroll
    for (int i = 0; i < TAPS / 4; ++i) d4[i] = s4[i];
}
template <int TAPS, int MODE, bool BI, bool DST8>
__global__ void __launch_bounds__(NT) pred_plane_fast_kernel(const __grid_constant__ FastParams fp)
{
    using G = FastGeom<TAPS, MODE, BI>;
    const PredParams &p = fp.p;
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *src_s = smem, *mid = smem + G::SRC_WORDS;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * G::TW, y0 = blockIdx.y * G::TH, f = blockIdx.z;
    const int w = min(G::TW, p.width - x0), h = min(G::TH, p.height - y0);
    uint8_t *dst = p.dst + f * p.fs_dst + (ptrdiff_t)y0 * p.sd + x0;
    const ptrdiff_t ro = f * p.fs_ref + (ptrdiff_t)y0 * p.sr + x0;
    constexpr int NJ = G::TW / 8, HITEMS = NJ * G::TH / NT;  // 8-wide output groups per row (16); groups per thread (8)
    // horizontal-pass ownership: group jj = tid % 16 of rows y = tid / 16 + 8 k
    const int jj = tid % NJ, yb = tid / NJ;
    uint8_t *drow = dst + (ptrdiff_t)yb * p.sd + 8 * jj;
    const ptrdiff_t dstep = (ptrdiff_t)(NT / NJ) * p.sd;
    const int nvalid = w - 8 * jj;

    if (!BI && MODE == COPY) {
        // 128-bit loads -> two 64-bit stores, straight through registers: chunk ch = tid % 8 of rows tid / 8 + 16 k
        const int ch = tid & 7, r0 = tid >> 3;
        const uint8_t *s = p.ref0 + ro + (ptrdiff_t)r0 * p.sr + ch * 16;
        uint8_t *d = dst + (ptrdiff_t)r0 * p.sd + ch * 16;
        const int nv = w - ch * 16;
        int4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (r0 + 16 * k < h && nv > 0) v[k] = __ldg(reinterpret_cast<const int4 *>(s + (ptrdiff_t)(16 * k) * p.sr));  // may read <= 15 bytes right of w
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (r0 + 16 * k < h && nv > 0) {
                uint8_t *dk = d + (ptrdiff_t)(16 * k) * p.sd;
                put8<DST8>(dk, (uint32_t)v[k].x, (uint32_t)v[k].y, nv);
                put8<DST8>(dk + 8, (uint32_t)v[k].z, (uint32_t)v[k].w, nv - 8);
            }
        return;
    }

    if (!BI) {
        stage_fast<G>(src_s, p.ref0 + ro, p.sr, w, h, tid);
        __syncthreads();
        if (MODE == H_ONLY) {
            int cx4[TAPS / 4];
            take<TAPS>(cx4, fp.c[0].x4);
            const uint32_t *s = src_s + yb * G::SPW + 3 + 2 * jj;
#pragma unroll
            for (int k = 0; k < HITEMS; ++k) {
                int o[8];
                hpass_bytes<TAPS>(s + k * (NT / NJ) * G::SPW, cx4, 32, o);
                if (yb + k * (NT / NJ) < h)
                    put8<DST8>(drow + k * dstep, pack_sat_u8(o[0] >> 6, o[1] >> 6, o[2] >> 6, o[3] >> 6), pack_sat_u8(o[4] >> 6, o[5] >> 6, o[6] >> 6, o[7] >> 6),
                               nvalid);
            }
        } else if (MODE == V_ONLY) {
            constexpr int ITERS = G::NQ * (G::TH / G::R) / NT;  // 2
#pragma unroll 1
            for (int k = 0; k < ITERS; ++k) {
                const int id = tid + k * NT, q = id % G::NQ, rg = id / G::NQ;
                int v[G::R][4];
                vpass_rot<TAPS>(src_s + rg * G::R * G::SPW + q, G::SPW, fp.c[0].y4s, v);
                uint8_t *d = dst + (ptrdiff_t)(rg * G::R) * p.sd + 4 * q;
#pragma unroll
                for (int r = 0; r < G::R; ++r)
                    if (rg * G::R + r < h && 4 * q < w)
                        store4(d + (ptrdiff_t)r * p.sd, pack_sat_u8((v[r][0] + 32) >> 6, (v[r][1] + 32) >> 6, (v[r][2] + 32) >> 6, (v[r][3] + 32) >> 6), w - 4 * q);
            }
        } else {
            vertical_to_mid_fast<TAPS, G>(mid, src_s, fp.c[0].y4s, tid);
            __syncthreads();
            const uint32_t *m = mid + yb * G::MP + 4 * jj;
#pragma unroll
            for (int k = 0; k < HITEMS; ++k) {
                int o[8];
#ifdef HV_HROT  // experiment: slid taps in the horizontal pass too (slower: FMA-pipe bound)
                hpass_rot<TAPS>(m + k * (NT / NJ) * G::MP, fp.c[0].x2e, fp.c[0].x2o, 2048, o);
#else
                hpass_mid<TAPS>(m + k * (NT / NJ) * G::MP, fp.c[0].x2e, 2048, o);
#endif
                if (yb + k * (NT / NJ) < h)
                    put8<DST8>(drow + k * dstep, pack_sat_u8(o[0] >> 12, o[1] >> 12, o[2] >> 12, o[3] >> 12),
                               pack_sat_u8(o[4] >> 12, o[5] >> 12, o[6] >> 12, o[7] >> 12), nvalid);
            }
        }
    } else {
        uint32_t va[HITEMS][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (r) __syncthreads();
            stage_fast<G>(src_s, (r ? p.ref1 : p.ref0) + ro, p.sr, w, h, tid);
            __syncthreads();
            vertical_to_mid_fast<TAPS, G>(mid, src_s, fp.c[r].y4s, tid);
            __syncthreads();
            const uint32_t *m = mid + yb * G::MP + 4 * jj;
#pragma unroll
            for (int k = 0; k < HITEMS; ++k) {
                int o[8];
#ifdef HV_HROT  // experiment: slid taps in the horizontal pass too (slower: FMA-pipe bound)
                hpass_rot<TAPS>(m + k * (NT / NJ) * G::MP, fp.c[r].x2e, fp.c[r].x2o, 0, o);
#else
                hpass_mid<TAPS>(m + k * (NT / NJ) * G::MP, fp.c[r].x2e, 0, o);
#endif
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
                    if (yb + k * (NT / NJ) < h) put8<DST8>(drow + k * dstep, pack_sat_u8(s[0], s[1], s[2], s[3]), pack_sat_u8(s[4], s[5], s[6], s[7]), nvalid);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ plane form, streaming path
//
// The tile kernels above pay ~18 issue slots per two-pass sample, a third of them for moving the intermediate through shared
// memory and re-pairing it (profiles/r01_pred.md).  This kernel has no shared memory and no barrier at all: a thread owns
// FOUR ADJACENT OUTPUT COLUMNS and walks down a strip of rows.  Per source row it loads the 16 bytes around its columns
// (three aligned words, neighbours overlap in L1), runs the HORIZONTAL filter on bytes (IDP.4A on funnel-shifted windows -
// here the reference's own order, horizontal first, is the cheap one), packs the exact int16 result with the previous
// row's into a vertical pair (one PRMT per column) and keeps the last eight pairs in a register ring; every pair feeds the
// four output rows whose tap pairs it completes (IDP.2A), so the VERTICAL filter never re-reads or re-aligns anything.
//   per two-pass sample: 2.2 IDP.4A + 1.7 SHF (H) + 1 PRMT (pair) + 4 IDP.2A (V) + 1 SHF + 0.5 I2IP + 1 LDG + 0.25 STG = ~12
// Needs 4-byte aligned reference rows and destination rows; anything else goes to the tile kernels.
constexpr int STRIP = 64;  // output rows per thread (32 and 128 measured: no better)

template <int TAPS>
__device__ __forceinline__ void hrow4(const uint32_t (&W)[3], const int (&cx4)[TAPS / 4], int (&t)[4])
{
    constexpr int LEFT = TAPS / 2 - 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int acc = 0;
#pragma unroll
        for (int g = 0; g < TAPS / 4; ++g) {
            const int o = i - LEFT + 4 * g + 4;  // byte offset of the window inside W (W[0] starts 4 bytes left of column 0)
            const uint32_t win = (o & 3) ? shr_bytes(W[o >> 2], W[((o >> 2) + 1) % 3], o & 3) : W[o >> 2];
            acc = dp4a_us(win, cx4[g], acc);
        }
        t[i] = acc;
    }
}

// one reference of the streaming filter: ring of vertical pairs + the running previous row
template <int TAPS>
struct StreamRef {
    static constexpr int RING = TAPS;          // pairs ending at the last TAPS rows
    uint32_t ring[RING][4];
    int prev[4];
};

// One trip of TAPS input rows: all loads first, then the arithmetic row by row.  CHECK = false: every row exists and every
// output row of the trip is inside the strip (the common case: whole strips, all trips but the first and the tail), so
// nothing is clamped or tested; FIRST = the trip that only primes the ring (its single output comes from its last row).
// Measured alternatives that lost (profiles/r01_pred.md): a register double buffer for the next trip's loads (121-192
// registers, half the occupancy) and L1 prefetches of the next trip.
//   SM = true: the rows come from a TMA-staged box in shared memory (pred_stream_tma_kernel): src[] then points at the
// thread's first word of the trip's first row inside the box, rows are SM_PITCH bytes apart, nothing needs clamping
// (rows outside the tensor arrive zero-filled and only feed output rows that are skipped) and src[] is not advanced.
constexpr int SM_PITCH = 544;  // bytes per staged row: 16-byte aligned start + up to 12 bytes of shift + 4 + 512 + 8, rounded to 16
//   NR < TAPS: the unchecked tail of a strip (its last TAPS-1 input rows).
template <int TAPS, int MODE, bool BI, bool FIRST, bool CHECK, bool SM = false, int NR = TAPS>
__device__ __forceinline__ void stream_trip(const FastParams &fp, StreamRef<TAPS> (&st)[BI ? 2 : 1], const uint8_t *(&src)[BI ? 2 : 1], uint8_t *&d, int r0,
                                            int rows_in, int h, int nvalid)
{
    constexpr bool NEED_H = (BI && MODE != COPY) || (MODE & 1), NEED_V = (BI && MODE != COPY) || (MODE & 2);   // bi: both passes, except the full-sample average (MODE = COPY)
    constexpr int NREF = BI ? 2 : 1;
    const PredParams &p = fp.p;
    uint32_t W[NREF][TAPS][3];
#pragma unroll
    for (int k = 0; k < NR; ++k)
#pragma unroll
        for (int rf = 0; rf < NREF; ++rf) {
            if (SM) {
                const uint32_t *row = reinterpret_cast<const uint32_t *>(src[rf] + k * SM_PITCH);
                if (NEED_H) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) W[rf][k][j] = row[j];
                } else {
                    W[rf][k][0] = row[0];
                }
                continue;
            }
            const ptrdiff_t off = CHECK ? (ptrdiff_t)(min(r0 + k, rows_in - 1) - r0) * p.sr : (ptrdiff_t)k * p.sr;
            const uint32_t *row = reinterpret_cast<const uint32_t *>(src[rf] + off);
            if (NEED_H) {
#pragma unroll
                for (int j = 0; j < 3; ++j) W[rf][k][j] = __ldg(row + j);   // bytes x-4 .. x+7 cover every tap of the four columns
            } else {
                W[rf][k][0] = __ldg(row);
            }
        }
    if (!SM) {
#pragma unroll
        for (int rf = 0; rf < NREF; ++rf) src[rf] += (ptrdiff_t)TAPS * p.sr;
    }
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        int vout[NREF][4];
#pragma unroll
        for (int rf = 0; rf < NREF; ++rf) {
            int t[4];
            if (NEED_H) {
                int cx4[TAPS / 4];
                take<TAPS>(cx4, fp.c[rf].x4);
                hrow4<TAPS>(W[rf][k], cx4, t);
            }
            if (NEED_V && !NEED_H) {
                // vertical pass alone: stay on bytes.  Per column a sliding QUAD of the last four rows (one PRMT shifts the new
                // row's byte in); slot k holds the quad ending at row rr, and an output row takes the quads ending at its 4th
                // (and 8th) tap row: one PRMT + TAPS/4 IDP.4A per sample instead of unpack + pair + TAPS/2 IDP.2A.
#pragma unroll
                for (int i = 0; i < 4; ++i) st[rf].ring[k][i] = __byte_perm(st[rf].ring[(k + TAPS - 1) % TAPS][i], W[rf][k][0], 0x4321 + (i << 12));
                if (FIRST && k != TAPS - 1) continue;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    int acc = 32;
#pragma unroll
                    for (int g = 0; g < TAPS / 4; ++g) acc = dp4a_us(st[rf].ring[(k + 4 + 4 * g) % TAPS][i], fp.c[rf].y4s[0][g], acc);
                    vout[rf][i] = acc;
                }
            } else if (NEED_V) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    st[rf].ring[k][i] = pack16(st[rf].prev[i], t[i]);   // pair (rr-1, rr); slot k = rr mod TAPS
                    st[rf].prev[i] = t[i];
                }
                if (FIRST && k != TAPS - 1) continue;                   // priming rows: no output yet
                // output row y = rr - (TAPS-1) uses the pairs ending at rows y+1, y+3, .. = slots (k + 2 + 2g) mod TAPS
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    int acc = BI ? 0 : (NEED_H ? 2048 : 32);
#pragma unroll
                    for (int g = 0; g < TAPS / 2; ++g) acc = dp2a_lo(st[rf].ring[(k + 2 + 2 * g) % TAPS][i], fp.c[rf].y2[g], acc);
                    vout[rf][i] = acc;
                }
            } else if (NEED_H) {
#pragma unroll
                for (int i = 0; i < 4; ++i) vout[rf][i] = t[i] + 32;
            }
        }
        if (NEED_V && FIRST && k != TAPS - 1) continue;
        if (CHECK) {
            const int y = NEED_V ? r0 + k - (TAPS - 1) : r0 + k;
            if (y < 0 || y >= h) continue;
        }
        uint32_t o;
        if (!NEED_H && !NEED_V) {
            o = BI ? __vavgu4(W[0][k][0], W[NREF - 1][k][0]) : W[0][k][0];   // full-sample position(s): the row passes through / (a + b + 1) >> 1 (TMA-fed kernel only)
        } else if (BI) {
            int s4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) s4[i] = ((int)(short)(vout[0][i] >> 6) + (int)(short)(vout[1][i] >> 6) + 64) >> 7;   // int16 wrap as in the reference's C
            o = pack_sat_u8(s4[0], s4[1], s4[2], s4[3]);
        } else {
            constexpr int SH = (NEED_H && NEED_V) ? 12 : 6;
            o = pack_sat_u8(vout[0][0] >> SH, vout[0][1] >> SH, vout[0][2] >> SH, vout[0][3] >> SH);
        }
        // (the TMA-fed kernel runs unchecked trips only in warps without a right-edge lane: no branch per row there.  The LDG
        // kernel keeps the test: without it ptxas schedules its loads across rows and spills at 64 registers)
        if ((SM && !CHECK) || nvalid >= 4) *reinterpret_cast<uint32_t *>(d) = o;
        else store4(d, o, nvalid);
        d += p.sd;   // the destination pointer follows the output rows
    }
}

template <int TAPS, int MODE, bool BI>
__global__ void __launch_bounds__(NT) pred_stream_kernel(const __grid_constant__ FastParams fp)
{
    constexpr bool NEED_H = BI || (MODE & 1), NEED_V = BI || (MODE & 2);
    constexpr int LEFT = TAPS / 2 - 1, NREF = BI ? 2 : 1;
    const PredParams &p = fp.p;
    const int q = blockIdx.x * NT + threadIdx.x, x = 4 * q, f = blockIdx.z;
    if (x >= p.width) return;
    const int y0 = blockIdx.y * STRIP, h = min(STRIP, p.height - y0), nvalid = p.width - x;
    uint8_t *d = p.dst + f * p.fs_dst + (ptrdiff_t)y0 * p.sd + x;
    const uint8_t *src[NREF];
    src[0] = p.ref0 + f * p.fs_ref + (ptrdiff_t)(y0 - (NEED_V ? LEFT : 0)) * p.sr + x - (NEED_H ? 4 : 0);
    if (BI) src[1] = p.ref1 + f * p.fs_ref + (ptrdiff_t)(y0 - LEFT) * p.sr + x - 4;
    const int rows_in = h + (NEED_V ? TAPS - 1 : 0);
    StreamRef<TAPS> st[NREF];

    if (h == STRIP) {
        // whole strip: unchecked trips (+ the priming trip and a checked tail when a vertical pass runs)
        if (NEED_V) {
            stream_trip<TAPS, MODE, BI, true, false>(fp, st, src, d, 0, rows_in, h, nvalid);
#pragma unroll 1
            for (int r0 = TAPS; r0 + TAPS <= STRIP; r0 += TAPS) stream_trip<TAPS, MODE, BI, false, false>(fp, st, src, d, r0, rows_in, h, nvalid);
            stream_trip<TAPS, MODE, BI, false, true>(fp, st, src, d, STRIP, rows_in, h, nvalid);   // rows STRIP .. STRIP+TAPS-2
        } else {
#pragma unroll 1
            for (int r0 = 0; r0 < STRIP; r0 += TAPS) stream_trip<TAPS, MODE, BI, false, false>(fp, st, src, d, r0, rows_in, h, nvalid);
        }
    } else {
#pragma unroll 1
        for (int r0 = 0; r0 < rows_in; r0 += TAPS) stream_trip<TAPS, MODE, BI, false, true>(fp, st, src, d, r0, rows_in, h, nvalid);
    }
}

template <int TAPS, int MODE, bool BI>
int launch_stream(const FastParams &fp, int n_frames, void *stream)
{
    const dim3 grid(((fp.p.width + 3) / 4 + NT - 1) / NT, (fp.p.height + STRIP - 1) / STRIP, n_frames);
    return launch(pred_stream_kernel<TAPS, MODE, BI>, grid, dim3(NT), 0, stream, fp);
}

template <int TAPS>
int launch_uni_stream(const FastParams &fp, int mode, int n_frames, void *stream)
{
    switch (mode) {
        case H_ONLY: return launch_stream<TAPS, H_ONLY, false>(fp, n_frames, stream);
        case V_ONLY: return launch_stream<TAPS, V_ONLY, false>(fp, n_frames, stream);
        default: return launch_stream<TAPS, HV, false>(fp, n_frames, stream);
    }
}

// ------------------------------------------------------------------------------------------------ plane form, streaming path fed by TMA
//
// pred_stream_kernel issues its loads and then waits for them: ncu shows 49 % of its stall samples on the long scoreboard
// and every attempt to prefetch in registers cost more occupancy than it hid (profiles/r01_pred.md).  Here the same
// arithmetic (stream_trip<.., SM = true>) reads its rows from shared memory, and the rows get there by TMA.  A CTA owns
// 512 columns x one strip; the reference is described as a tensor of 32-bit words so that one box can be a whole
// 544-byte row segment (the 512 columns + the filter halo, starting on a 16-byte boundary as the TMA unit requires) x 16
// (or 8) rows.  3 (or 4) such boxes form a ring with a `full` mbarrier (armed with the byte count, completed by the TMA unit)
// and an `empty` mbarrier (one arrival per warp) each; thread 0 re-arms the box the CTA consumed one iteration earlier,
// so all boxes but one are in flight per CTA and no thread waits on DRAM with an instruction slot to fill.  There
// is no CTA-wide barrier in the loop.  Rows outside the declared tensor are zero-filled by the hardware, so strips at the
// bottom edge need no clamping.  (First version: a private ring of 160-byte boxes per warp - the TMA unit then delivers
// only ~2.4 TB/s, 80 us for ANY filter position; wide boxes fix that.)
constexpr int TS_COLS = 4 * NT;                                         // 512 output columns per CTA

struct alignas(64) StreamMaps {
    CUtensorMap tm[2];   // per reference: 32-bit words from (x = -4 or 0, y = -(TAPS/2-1) or 0) of the plane, 16-byte aligned start
    int shift[2];        // bytes between that aligned start and the first byte the kernel wants (0, 4, 8 or 12)
};

// ring geometry: 3 boxes of 16 rows (one reference) or 4 x 2 boxes of 8 rows (two references) - 26 / 35 KB per CTA
template <bool BI>
struct TsRing {
    static constexpr int ROWS = BI ? 8 : 16, STAGES = BI ? 4 : 3, BOX = SM_PITCH * ROWS;   // BOX is a multiple of 128 bytes
};
template <int TAPS, int MODE, bool BI>
struct TsGeom {
    static constexpr int NREF = BI ? 2 : 1;
    static constexpr int RING_BYTES = TsRing<BI>::STAGES * NREF * TsRing<BI>::BOX;
    static constexpr int SMEM_BYTES = RING_BYTES + 2 * TsRing<BI>::STAGES * 8;
};

// resident CTAs per SM the register budget is set for: 8 (64 registers), except the 8-tap two-pass kernel, which spills at 64
// and runs 3 % faster at 7 CTAs x 72 registers; bi: 4 (8-tap) / 6 (4-tap)
template <int TAPS, int MODE, bool BI>
__global__ void __launch_bounds__(NT, (BI && MODE != COPY) ? (TAPS == 8 ? 4 : 6) : ((TAPS == 8 && MODE == HV) ? 7 : 8))
    pred_stream_tma_kernel(const __grid_constant__ FastParams fp, const __grid_constant__ StreamMaps sm, int strip /* output rows per CTA, a multiple of 8 */)
{
    using G = TsGeom<TAPS, MODE, BI>;
    constexpr bool NEED_V = (BI && MODE != COPY) || (MODE & 2);
    constexpr int NREF = G::NREF, TS_ROWS = TsRing<BI>::ROWS, TS_STAGES = TsRing<BI>::STAGES, TS_BOX = TsRing<BI>::BOX;
    constexpr int TRIPS = TS_ROWS / TAPS;   // trips of TAPS rows per staged box
    extern __shared__ __align__(128) uint8_t ts_smem[];
    uint8_t *const box = ts_smem;
    uint64_t *const full = reinterpret_cast<uint64_t *>(ts_smem + G::RING_BYTES), *const empty = full + TS_STAGES;
    const PredParams &p = fp.p;
    const int tid = threadIdx.x, lane = tid & 31;
    const int x0 = blockIdx.x * TS_COLS, x = x0 + 4 * tid, f = blockIdx.z;
    const int y0 = blockIdx.y * strip, h = min(strip, p.height - y0), nvalid = p.width - x;
    const int rows_in = h + (NEED_V ? TAPS - 1 : 0), nbox = (rows_in + TS_ROWS - 1) / TS_ROWS;
    const bool active = x0 + 128 * (tid >> 5) < p.width;          // warp-uniform: the warp owns at least one column
    const bool inner = x0 + 128 * (tid >> 5) + 128 <= p.width;    // warp-uniform: no lane at the right edge
    uint8_t *d = p.dst + f * p.fs_dst + (ptrdiff_t)y0 * p.sd + x;

    auto arm = [&](int b) {   // thread 0: request box b of the strip into its ring slot
        const int s = b % TS_STAGES;
        tma::mbar_expect_tx(full + s, NREF * TS_BOX);
#pragma unroll
        for (int rf = 0; rf < NREF; ++rf) tma::load_box_3d(box + (s * NREF + rf) * TS_BOX, &sm.tm[rf], x0 / 4, y0 + b * TS_ROWS, f, full + s);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TS_STAGES; ++s) tma::mbar_init(full + s, 1), tma::mbar_init(empty + s, NT / 32);
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < TS_STAGES; ++b)
            if (b < nbox) arm(b);
    }

    StreamRef<TAPS> st[NREF];
#pragma unroll 1
    for (int b = 0; b < nbox; ++b) {
        const int s = b % TS_STAGES;
        if (tid == 0 && b >= 1 && b - 1 + TS_STAGES < nbox) {   // the box consumed one iteration ago: every warp has (nearly always) let go of it
            tma::mbar_wait(empty + (b - 1) % TS_STAGES, ((b - 1) / TS_STAGES) & 1);
            arm(b - 1 + TS_STAGES);
        }
        tma::mbar_wait(full + s, (b / TS_STAGES) & 1);
        if (active) {
            const uint8_t *src[NREF];
#pragma unroll
            for (int u = 0; u < TRIPS; ++u) {
                const int r0 = b * TS_ROWS + u * TAPS;
                if (r0 >= rows_in) break;
#pragma unroll
                for (int rf = 0; rf < NREF; ++rf) src[rf] = box + (s * NREF + rf) * TS_BOX + u * TAPS * SM_PITCH + sm.shift[rf] + 4 * tid;
                // trips whose output rows all exist run unchecked: the priming trip (one output row), the trips inside the strip,
                // and - when the strip height is a multiple of TAPS - the tail of TAPS-1 rows
                if (inner && NEED_V && r0 == 0) stream_trip<TAPS, MODE, BI, true, false, true>(fp, st, src, d, r0, rows_in, h, nvalid);
                else if (inner && (!NEED_V || r0 > 0) && r0 + TAPS <= h) stream_trip<TAPS, MODE, BI, false, false, true>(fp, st, src, d, r0, rows_in, h, nvalid);
                else if (inner && NEED_V && r0 == h) stream_trip<TAPS, MODE, BI, false, false, true, TAPS - 1>(fp, st, src, d, r0, rows_in, h, nvalid);
                else stream_trip<TAPS, MODE, BI, false, true, true>(fp, st, src, d, r0, rows_in, h, nvalid);
            }
        }
        __syncwarp();   // every lane of the warp has read the box
        if (lane == 0) tma::mbar_arrive(empty + s);
    }
}

// strip height: about 128 rows (the priming and tail trips of a strip cost ~1.5 trips of overhead), nudged so that the grid
// fills whole waves of resident CTAs - at 4K x 16 frames 120 rows give 1.95 waves where 128 would give 1.84
static int pick_strip(int height, int ctas_per_strip_row, int slots)
{
    int best = 64;
    double best_score = -1;
    for (int strip = 64; strip <= 160; strip += 8) {
        const long long ctas = (long long)ctas_per_strip_row * ((height + strip - 1) / strip);
        const long long waves = (ctas + slots - 1) / slots;
        const double fill = (double)ctas / (double)(waves * slots);
        const double score = fill * (1.0 - 12.0 / (strip + 12.0));   // wave fill x share of rows that are not priming / tail overhead
        if (score > best_score) best_score = score, best = strip;
    }
    return best;
}

template <int TAPS, int MODE, bool BI>
int launch_stream_tma(const FastParams &fp, const StreamMaps &sm, int n_frames, void *stream)
{
    using G = TsGeom<TAPS, MODE, BI>;
    auto kern = pred_stream_tma_kernel<TAPS, MODE, BI>;
    if (G::SMEM_BYTES > 48 * 1024 && set_max_smem(kern, G::SMEM_BYTES)) return (int)cudaErrorInvalidValue;
    static int per_sm = 0;   // resident CTAs per SM of this instantiation (a property of the binary: benign race)
    if (!per_sm) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, NT, (size_t)G::SMEM_BYTES) != cudaSuccess || n < 1) n = 1;
        per_sm = n;
    }
    const int cols = (fp.p.width + TS_COLS - 1) / TS_COLS;
    // one-pass positions and copies are bound by memory: short strips, whose two or three boxes are all requested at once,
    // measured best (32 rows: 45-46 us per 16 4K planes; 64: 47-48; 120: 50-53).  Two-pass positions: see pick_strip.
    int strip = (MODE == HV) ? pick_strip(fp.p.height, cols * n_frames, per_sm * sm_count()) : 32;
    if (const char *e = tune::knob("HEVCASM_PRED_STRIP")) {   // tuning knob: rows per CTA (rounded to a multiple of 8)
        const int v = atoi(e) & ~7;
        if (v >= 8 && v <= 1024) strip = v;
    }
    const dim3 grid(cols, (fp.p.height + strip - 1) / strip, n_frames);
    return launch(kern, grid, dim3(NT), (size_t)G::SMEM_BYTES, stream, fp, sm, strip);
}

template <int TAPS>
int launch_uni_stream_tma(const FastParams &fp, const StreamMaps &sm, int mode, int n_frames, void *stream)
{
    switch (mode) {
        case COPY: return launch_stream_tma<TAPS, COPY, false>(fp, sm, n_frames, stream);
        case H_ONLY: return launch_stream_tma<TAPS, H_ONLY, false>(fp, sm, n_frames, stream);
        case V_ONLY: return launch_stream_tma<TAPS, V_ONLY, false>(fp, sm, n_frames, stream);
        default: return launch_stream_tma<TAPS, HV, false>(fp, sm, n_frames, stream);
    }
}

#include "pred_umma.cuh"  // two-pass positions with the horizontal pass on tcgen05

// ------------------------------------------------------------------------------------------------ PU lists, streaming
//
// Prediction-unit lists (per-PU size and quarter / eighth-sample motion vector).  A CTA takes LIST_G descriptors, a scan turns their widths
// into a prefix of work items - one item = a strip of 8 columns (4 with two references) over the whole height of a PU - and every thread
// picks items off that list, so 8x8 PUs fill the CTA as well as 64x64 ones do (a tile kernel that spent a 128-thread CTA per PU ran 8x8
// lists at 58 Gsamples/s).  A thread streams its strip top to bottom straight from global memory: per input row it loads the strip's
// footprint (aligned 8-byte loads, the start word selected and one funnel shift per word), runs the HORIZONTAL filter on the bytes (IDP.4A),
// packs the exact int16 result with the previous row's into a vertical pair and keeps the last TAPS pairs in a register ring; every output
// row is four (two) IDP.2A per column over that ring.  The first TAPS-1 rows only prime the ring; after that one input row gives one
// output row.  Rows come in batches whose loads are issued one batch ahead (one reference).  Every PU runs the two-pass arithmetic; a zero
// fraction is the {64} filter, for which the two-pass rounding (sum + 2048) >> 12 reduces exactly to the one-pass (sum + 32) >> 6 and to a
// copy, so one code path serves all positions without divergence.  Reference rows and strides may have any alignment.
//   Measured on 16 4K frames of 8x8 / 64x64 PUs with random +-16-sample vectors (profiles/r02_pred.md): 230 / 110 us (round 1's kernel -
// 4-column strips, trips of eight rows with the vertical filter run on all of them, the group's descriptors fetched through an integer
// division - 283 / 171 us).  Every thread-row touches its own line, so a warp's load is ~30 sectors in ~30 lines: the L1 data stage is 60-73 %
// busy and the loads' latency (33 % of the stall samples, 12 warps per SM at 168 registers) is the rest; the list order (raster or CTU
// z-order) and the vector spread (+-16 samples or none: 197 us) change little.  Tried and dropped: a TMA box per PU (a box must start
// 16-byte aligned, and the unit retires one box row per ~1.5 cycles: 24 cycles per 8x8 PU, profiles/r02_tma_box_probe.txt); the whole
// footprint of a small PU requested at once (348 us); 16-byte loads (two per row, but two SELs per word: 275 us); 4-column strips for one
// reference (271 us at 128 registers); a persistent grid, per CTA (251 / 139 us, mixed sizes 10 % slower: fixed stride, uneven groups) and
// per warp with the next descriptors prefetched into registers (230 / 168 us: a warp alone on 32 large PUs).
constexpr int LIST_G = 128;   // descriptors per CTA and trip: 128 PUs of 8x8 are 128 eight-column items, one per thread
// blockIdx.y slices a group's item list; enough slices that a short list of large PUs (16 groups for a 4K frame of 64x64 PUs, 1024 items each)
// still spreads over the chip, one slice when there are groups enough (empty slices of small-PU groups only cost launches)
static dim3 list_grid(int n_pu)
{
    // one CTA per group (the hardware's CTA scheduler balances groups of very different weight - a persistent grid with a fixed stride ran
    // mixed-size lists 10 % slower); a short list is sliced (blockIdx.y): up to eight CTAs share the items of a group
    const int groups = (n_pu + LIST_G - 1) / LIST_G;
    return dim3((unsigned)groups, (unsigned)std::max(1, std::min(8, (4 * sm_count() + groups - 1) / groups)));
}

// horizontal filter of COLS adjacent columns from the aligned footprint words A (A[0] byte 0 = first tap of column 0)
template <int TAPS, int COLS, int NA>
__device__ __forceinline__ void hrow_cols(const uint32_t (&A)[NA], const int (&cx4)[TAPS / 4], int (&t)[COLS])
{
#pragma unroll
    for (int i = 0; i < COLS; ++i) {
        int acc = 0;
#pragma unroll
        for (int g = 0; g < TAPS / 4; ++g) {
            const int o = i + 4 * g;
            const uint32_t win = (o & 3) ? shr_bytes(A[o >> 2], A[((o >> 2) + 1) % NA], o & 3) : A[o >> 2];
            acc = dp4a_us(win, cx4[g], acc);
        }
        t[i] = acc;
    }
}

template <int TAPS, bool BI, int COLS, int MINB, int LDB>
__global__ void __launch_bounds__(NT, MINB) pred_list_stream_kernel(PredParams p)
{
    constexpr int G = LIST_G, NWARP = NT / 32, NREF = BI ? 2 : 1, DW0 = BI ? 8 : 6, LEFT = TAPS / 2 - 1, FB = TAPS == 8 ? 2 : 3, FM = (1 << FB) - 1;
    constexpr int NEED = COLS + TAPS - 1;        // footprint bytes of one strip row
    constexpr int NA = (NEED + 3) / 4;           // aligned words the horizontal filter reads
    // words loaded per row, as aligned loads of LDB bytes: the footprint starts at any byte of the first, and the word it starts in is selected
    // afterwards (one SEL per word and address bit above the word).  Wider loads = fewer requests to the L1 data stage, which 4-byte loads kept
    // 73 % busy on 8x8 PU lists
    constexpr int NLD = (NEED + LDB - 1 + LDB - 1) / LDB, NW = NLD * (LDB / 4);
    static_assert(LDB == 4 || LDB == 8 || LDB == 16, "load width");
    // rows per load batch; PIPE: the loads of the next batch are in flight while this one is filtered (two buffers - one reference only: with
    // two the registers do not fit)
    constexpr int LB = BI ? 2 : TAPS / 2, NB = TAPS / LB;
    constexpr bool PIPE = !BI;
    const int DW = DW0 + p.desc_frame;           // a trailing frame index makes one launch cover the PU lists of a whole batch of frames
    __shared__ int s_prefix[G + 1];
    __shared__ int s_warp_sum[NWARP];
    __shared__ short s_desc[G][10];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // descriptors stream from DRAM and nothing can start before they are here: every CTA asks L2 for the records of the group that the CTA
    // taking this one's place will read (CTAs start in index order, three per SM at a time)
    {
        const size_t ahead = ((size_t)blockIdx.x + 3 * 160) * G * DW * sizeof(int16_t) + (size_t)tid * 128;
        if (blockIdx.y == 0 && (size_t)tid * 128 < (size_t)G * DW * sizeof(int16_t) && ahead < (size_t)p.n_pu * DW * sizeof(int16_t))
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(p.pus) + ahead));
    }
    {
        // thread t owns descriptor t of the group: inclusive scan of the item counts (warp scans, then the four warp totals)
        short dreg[DW0 + 1];
        const long long di = (long long)blockIdx.x * G + tid;
#pragma unroll
        for (int j = 0; j <= DW0; ++j) dreg[j] = 0;
        if (di < p.n_pu) {
            const int16_t *dsc = p.pus + (size_t)di * DW;
#pragma unroll
            for (int j = 0; j < DW0; ++j) dreg[j] = dsc[j];
            if (p.desc_frame) dreg[DW0] = dsc[DW0];
        }
        const int dw = dreg[2], dh = dreg[3];
        const int nq = (dw > 0 && dh > 0 && dw <= 64 && dh <= 64) ? (dw + COLS - 1) / COLS : 0;   // a zero descriptor (past the end) has no items
        int incl = nq;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp_sum[warp] = incl;
#pragma unroll
        for (int j = 0; j <= DW0; ++j) s_desc[tid][j] = dreg[j];
        __syncthreads();
        int base = 0;
#pragma unroll
        for (int k = 0; k < NWARP; ++k)
            if (k < warp) base += s_warp_sum[k];
        s_prefix[tid + 1] = base + incl;
        if (tid == 0) s_prefix[0] = 0;
        __syncthreads();
    }
    const int total = s_prefix[G];
    // blockIdx.y slices the item list, so that a few large PUs spread over several CTAs; for small PUs the extra CTAs find nothing
    // beyond `total` and go on
    for (int item = tid + NT * blockIdx.y; item < total; item += NT * gridDim.y) {
        int pu = 0;
#pragma unroll
        for (int step = G / 2; step; step >>= 1)
            if (s_prefix[pu + step] <= item) pu += step;
        const int q = item - s_prefix[pu];
        const int x = s_desc[pu][0], y = s_desc[pu][1], w = s_desc[pu][2], h = s_desc[pu][3];
        const int nvalid = w - COLS * q, frame = s_desc[pu][DW0];
        uint8_t *d = p.dst + frame * p.fs_dst + (ptrdiff_t)y * p.sd + x + COLS * q;
        const uint8_t *src[NREF];   // first footprint byte of the strip's first input row
        int cx4[NREF][TAPS / 4], cy2[NREF][TAPS / 2];
#pragma unroll
        for (int rf = 0; rf < NREF; ++rf) {
            const int mvx = s_desc[pu][4 + 2 * rf], mvy = s_desc[pu][5 + 2 * rf];
            src[rf] = (rf ? p.ref1 : p.ref0) + frame * p.fs_ref + (ptrdiff_t)(y + (mvy >> FB) - LEFT) * p.sr + x + (mvx >> FB) + COLS * q - LEFT;
            Coefs<TAPS> c;
            c.load(mvx & FM);
#pragma unroll
            for (int g = 0; g < TAPS / 4; ++g) cx4[rf][g] = c.p4[g];
            c.load(mvy & FM);
#pragma unroll
            for (int g = 0; g < TAPS / 2; ++g) cy2[rf][g] = c.p2[g];
        }
        uint32_t ring[NREF][TAPS][COLS];   // ring[.][r % TAPS] = vertical pairs (row r-1, row r) of the horizontal results
        int prev[NREF][COLS];
#pragma unroll
        for (int rf = 0; rf < NREF; ++rf)
#pragma unroll
            for (int i = 0; i < COLS; ++i) prev[rf][i] = 0;
        // one input row: NW aligned words covering the footprint, and the byte the footprint starts at inside the first
        const int rows_in = h + TAPS - 1;
        // one input row: NW aligned words covering the footprint, and the bit the footprint starts at inside the first
        auto fetch = [&](int rf, int r, uint32_t (&wd)[NW], int &sh8) {
            const uint8_t *row = src[rf] + (ptrdiff_t)min(r, rows_in - 1) * p.sr;   // batches are whole: rows past the footprint repeat its last one
            const int a = (int)((uintptr_t)row & (LDB - 1));
            if (LDB == 16) {
                const uint4 *ra = reinterpret_cast<const uint4 *>(row - a);
#pragma unroll
                for (int k = 0; k < NLD; ++k) {
                    const uint4 v = __ldg(ra + k);
                    wd[(4 * k) % NW] = v.x, wd[(4 * k + 1) % NW] = v.y, wd[(4 * k + 2) % NW] = v.z, wd[(4 * k + 3) % NW] = v.w;
                }
            } else if (LDB == 8) {
                const uint2 *ra = reinterpret_cast<const uint2 *>(row - a);
#pragma unroll
                for (int k = 0; k < NLD; ++k) {
                    const uint2 v = __ldg(ra + k);
                    wd[(2 * k) % NW] = v.x, wd[(2 * k + 1) % NW] = v.y;
                }
            } else {
                const uint32_t *ra = reinterpret_cast<const uint32_t *>(row - a);
#pragma unroll
                for (int k = 0; k < NW; ++k) wd[k] = __ldg(ra + k);
            }
            sh8 = a;
        };
        // horizontal filter of one fetched row; its pairs with the previous row go to a ring slot
        auto hpack = [&](int rf, const uint32_t (&wd)[NW], int sh8, uint32_t (&slot)[COLS]) {
            uint32_t A[NA], u[NA + 1];
            if (LDB == 16) {
                uint32_t v[NA + 2];
#pragma unroll
                for (int j = 0; j < NA + 2; ++j) v[j] = (sh8 & 8) ? wd[(j + 2) % NW] : wd[j % NW];
#pragma unroll
                for (int j = 0; j < NA + 1; ++j) u[j] = (sh8 & 4) ? v[j + 1] : v[j];
            } else if (LDB == 8) {
#pragma unroll
                for (int j = 0; j < NA + 1; ++j) u[j] = (sh8 & 4) ? wd[(j + 1) % NW] : wd[j % NW];
            } else {
#pragma unroll
                for (int j = 0; j < NA + 1; ++j) u[j] = wd[j % NW];
            }
#pragma unroll
            for (int j = 0; j < NA; ++j) A[j] = __funnelshift_r(u[j], u[j + 1], 8 * (sh8 & 3));
            int t[COLS];
            hrow_cols<TAPS, COLS, NA>(A, cx4[rf], t);
#pragma unroll
            for (int i = 0; i < COLS; ++i) {
                slot[i] = pack16(prev[rf][i], t[i]);
                prev[rf][i] = t[i];
            }
        };
        // output row j from the ring; J = j mod TAPS picks the slots: the pairs ending at rows j+1, j+3, ..
        auto vstore = [&](int J, int j) {
            int vout[NREF][COLS];
#pragma unroll
            for (int rf = 0; rf < NREF; ++rf)
#pragma unroll
                for (int i = 0; i < COLS; ++i) {
                    int acc = BI ? 0 : 2048;
#pragma unroll
                    for (int g = 0; g < TAPS / 2; ++g) acc = dp2a_lo(ring[rf][(J + 1 + 2 * g) % TAPS][i], cy2[rf][g], acc);
                    vout[rf][i] = acc;
                }
            uint32_t o[COLS / 4];
#pragma unroll
            for (int c = 0; c < COLS / 4; ++c) {
                if (BI) {
                    int s4[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) s4[i] = ((int)(short)(vout[0][4 * c + i] >> 6) + (int)(short)(vout[NREF - 1][4 * c + i] >> 6) + 64) >> 7;
                    o[c] = pack_sat_u8(s4[0], s4[1], s4[2], s4[3]);
                } else {
                    o[c] = pack_sat_u8(vout[0][4 * c] >> 12, vout[0][4 * c + 1] >> 12, vout[0][4 * c + 2] >> 12, vout[0][4 * c + 3] >> 12);
                }
            }
            if (COLS == 8)
                store8(d + (ptrdiff_t)j * p.sd, o[0], o[COLS / 4 - 1], nvalid);
            else
                store4(d + (ptrdiff_t)j * p.sd, o[0], nvalid);
        };
        // input rows in batches of LB, NB batches per ring period: the ring slots and the buffer a batch sits in are compile-time
        uint32_t W[PIPE ? 2 : 1][NREF][LB][NW];
        int sh[PIPE ? 2 : 1][NREF][LB];
        if (PIPE) {
#pragma unroll
            for (int b = 0; b < LB; ++b)
#pragma unroll
                for (int rf = 0; rf < NREF; ++rf) fetch(rf, b, W[0][rf][b], sh[0][rf][b]);   // rows 0 .. LB-1 always exist (h >= 1, LB <= TAPS-1)
        }
#pragma unroll 1
        for (int r0 = 0; r0 < rows_in; r0 += TAPS) {
#pragma unroll
            for (int bt = 0; bt < NB; ++bt) {
                const int cur = PIPE ? (bt & 1) : 0, nxt = PIPE ? (cur ^ 1) : 0, ahead = PIPE ? LB : 0;
                // nothing below is conditional except the output row itself: values defined under a per-row test cost a register move each
                // (262 IMAD.MOV in 1944 instructions when the loads and the filter of a row were skipped past the footprint)
                if (r0 + LB * bt >= rows_in) break;                 // whole batches only: nothing left
                if (r0 + LB * bt + ahead < rows_in) {               // (a batch wholly past the footprint is not requested)
#pragma unroll
                    for (int b = 0; b < LB; ++b)
#pragma unroll
                        for (int rf = 0; rf < NREF; ++rf) fetch(rf, r0 + LB * bt + ahead + b, W[nxt][rf][b], sh[nxt][rf][b]);
                }
#pragma unroll
                for (int b = 0; b < LB; ++b) {
                    const int r = r0 + LB * bt + b;   // r mod TAPS = LB * bt + b
#pragma unroll
                    for (int rf = 0; rf < NREF; ++rf) hpack(rf, W[cur][rf][b], sh[cur][rf][b], ring[rf][LB * bt + b]);
                    if (r >= TAPS - 1 && r < rows_in) vstore((LB * bt + b + 1) % TAPS, r - (TAPS - 1));
                }
            }
        }
    }
}

template <int TAPS, int MODE, bool BI>
int launch_plane_fast(const FastParams &fp, dim3 grid, bool dst8, void *stream)
{
    using G = FastGeom<TAPS, MODE, BI>;
    if (dst8) {
        auto kern = pred_plane_fast_kernel<TAPS, MODE, BI, true>;
        if (G::SMEM_BYTES > 48 * 1024 && set_max_smem(kern, G::SMEM_BYTES)) return (int)cudaErrorInvalidValue;
        return launch(kern, grid, dim3(NT), (size_t)G::SMEM_BYTES, stream, fp);
    }
    auto kern = pred_plane_fast_kernel<TAPS, MODE, BI, false>;
    if (G::SMEM_BYTES > 48 * 1024 && set_max_smem(kern, G::SMEM_BYTES)) return (int)cudaErrorInvalidValue;
    return launch(kern, grid, dim3(NT), (size_t)G::SMEM_BYTES, stream, fp);
}

template <int TAPS>
int launch_uni_planes_fast(const FastParams &fp, dim3 grid, int mode, bool dst8, void *stream)
{
    switch (mode) {
        case COPY: return launch_plane_fast<TAPS, COPY, false>(fp, grid, dst8, stream);
        case H_ONLY: return launch_plane_fast<TAPS, H_ONLY, false>(fp, grid, dst8, stream);
        case V_ONLY: return launch_plane_fast<TAPS, V_ONLY, false>(fp, grid, dst8, stream);
        default: return launch_plane_fast<TAPS, HV, false>(fp, grid, dst8, stream);
    }
}

constexpr int FTW = 128, FTH = 64;  // fast-path tile

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

// H.265 8.5.3.3.3 filters on the host (the device copy is c_luma / c_chroma), packed the way the dot-product instructions take them
static PackedCoefs pack_coefs(int taps, int xFrac, int yFrac)
{
    static const int8_t luma[4][8] = {{0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1}};
    static const int8_t chroma[8][4] = {{0, 64, 0, 0}, {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4}, {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2}};
    int cx[8] = {0}, cy[8] = {0};
    for (int k = 0; k < taps; ++k) cx[k] = taps == 8 ? luma[xFrac][k] : chroma[xFrac][k], cy[k] = taps == 8 ? luma[yFrac][k] : chroma[yFrac][k];
    PackedCoefs c{};
    auto tap = [&](const int *t, int k) { return (k >= 0 && k < taps) ? (t[k] & 0xff) : 0; };
    for (int g = 0; g < 2; ++g) c.x4[g] = tap(cx, 4 * g) | (tap(cx, 4 * g + 1) << 8) | (tap(cx, 4 * g + 2) << 16) | (tap(cx, 4 * g + 3) << 24);
    for (int g = 0; g < 4; ++g) c.x2e[g] = tap(cx, 2 * g) | (tap(cx, 2 * g + 1) << 8);
    for (int g = 0; g < 5; ++g) c.x2o[g] = tap(cx, 2 * g - 1) | (tap(cx, 2 * g) << 8);
    for (int g = 0; g < 4; ++g) c.y2[g] = tap(cy, 2 * g) | (tap(cy, 2 * g + 1) << 8);
    for (int sft = 0; sft < 4; ++sft)
        for (int g = 0; g < 3; ++g) {
            int w = 0;
            for (int j = 0; j < 4; ++j) w |= tap(cy, 4 * g + j - sft) << (8 * j);
            c.y4s[sft][g] = w;
        }
    return c;
}

// the fast plane kernels issue aligned 128-bit loads on the reference rows (and may therefore touch up to 16 bytes left and
// 15 bytes right of the reference's own footprint - hevcasm_batch.h documents the padding this needs)
static bool planes_fast_ok(const uint8_t *ref0, const uint8_t *ref1, ptrdiff_t sr, ptrdiff_t fs_ref, int n_frames)
{
    if (tune::knob("HEVCASM_PRED_GENERIC")) return false;
    uintptr_t m = (uintptr_t)ref0 | (uintptr_t)sr | (ref1 ? (uintptr_t)ref1 : 0);
    if (n_frames > 1) m |= (uintptr_t)fs_ref;
    return (m & 15) == 0;
}
// the streaming kernels issue aligned 32-bit loads / stores: everything 4-byte aligned (HEVCASM_PRED_PATH=tile pins the tile kernels)
static bool stream_ok(const uint8_t *dst, ptrdiff_t sd, ptrdiff_t fs_dst, const uint8_t *ref0, const uint8_t *ref1, ptrdiff_t sr, ptrdiff_t fs_ref, int n_frames)
{
    const char *pin = tune::knob("HEVCASM_PRED_PATH");
    if (tune::knob("HEVCASM_PRED_GENERIC") || (pin && strcmp(pin, "stream"))) return false;
    uintptr_t m = (uintptr_t)dst | (uintptr_t)sd | (uintptr_t)ref0 | (uintptr_t)sr | (ref1 ? (uintptr_t)ref1 : 0);
    if (n_frames > 1) m |= (uintptr_t)fs_dst | (uintptr_t)fs_ref;
    return (m & 3) == 0;
}
// columns per work item of the PU-list kernel and CTAs per SM it is compiled for (HEVCASM_LIST_COLS=4 / 8 and HEVCASM_LIST_MINB pin them in
// the experiments build)
template <bool BI, int COLS, int MINB, int LDB>
static int launch_list_as(int taps, void *stream, const PredParams &p)
{
    const dim3 grid = list_grid(p.n_pu);
    return taps == 8 ? launch(pred_list_stream_kernel<8, BI, COLS, MINB, LDB>, grid, dim3(NT), 0, stream, p)
                     : launch(pred_list_stream_kernel<4, BI, COLS, MINB, LDB>, grid, dim3(NT), 0, stream, p);
}
// one reference: 8-column strips, three CTAs per SM (168 registers), 8-byte loads; two references: 4-column strips, four CTAs per SM.
// HEVCASM_LIST_LDB = 4 / 8 / 16 pins the load width in the experiments build (A/B)
template <bool BI>
static int launch_list(int taps, void *stream, const PredParams &p)
{
#ifdef HEVCASM_EXPERIMENTS
    if (const char *pl = tune::knob("HEVCASM_LIST_LDB")) {
        const int ldb = atoi(pl);
        if constexpr (BI) return ldb == 16 ? launch_list_as<true, 4, 4, 16>(taps, stream, p) : ldb == 8 ? launch_list_as<true, 4, 4, 8>(taps, stream, p) : launch_list_as<true, 4, 4, 4>(taps, stream, p);
        else return ldb == 16 ? launch_list_as<false, 8, 3, 16>(taps, stream, p) : ldb == 8 ? launch_list_as<false, 8, 3, 8>(taps, stream, p) : launch_list_as<false, 8, 3, 4>(taps, stream, p);
    }
#endif
    if constexpr (BI) return launch_list_as<true, 4, 4, 4>(taps, stream, p);
    else return launch_list_as<false, 8, 3, 8>(taps, stream, p);
}

// PU lists: streaming kernel unless HEVCASM_PRED_PATH=tile / HEVCASM_PRED_GENERIC pins the tile kernel (A/B, and the tile
// kernel reads exactly the reference's footprint per position while the streaming one always reads the two-pass footprint)
static bool list_stream_ok()
{
    const char *pin = tune::knob("HEVCASM_PRED_PATH");
    return !tune::knob("HEVCASM_PRED_GENERIC") && !(pin && strcmp(pin, "stream"));
}
// TMA-fed streaming kernel: describes each reference as a (x, y, frame) byte tensor starting at the first byte the filter
// footprint touches (x = -4 with a horizontal pass, y = -(taps/2-1) with a vertical one).  Not possible (-> LDG streaming kernel)
// when the strides are not multiples of 16, or when a row's footprint does not fit its stride (the right halo would then
// lie in the next row, outside the declared row).  HEVCASM_PRED_STREAM=ldg / =tma pin one kernel (A/B runs).
static bool stream_maps(StreamMaps *sm, const PredParams &p, int taps, int mode, bool bi, int n_frames)
{
    const char *pin = tune::knob("HEVCASM_PRED_STREAM");
    if (pin && !strcmp(pin, "ldg")) return false;
    if (!tma::describable(p.sr, p.fs_ref, n_frames)) return false;
    const bool need_h = (bi && mode != COPY) || (mode & 1), need_v = (bi && mode != COPY) || (mode & 2);
    const int left = need_h ? 4 : 0, top = need_v ? taps / 2 - 1 : 0;
    const long long ext_x = (long long)left + p.width + (need_h ? 8 : 0), ext_y = (long long)p.height + (need_v ? taps - 1 : 0);
    const uint8_t *refs[2] = {p.ref0, p.ref1};
    for (int rf = 0; rf < (bi ? 2 : 1); ++rf) {
        const uint8_t *first = refs[rf] - (ptrdiff_t)top * p.sr - left;
        if ((long long)((uintptr_t)first & 15) + ext_x > (long long)p.sr) return false;
        if (tma::describe_u32(&sm->tm[rf], first, p.sr, p.fs_ref, ext_x, ext_y, n_frames, SM_PITCH, bi ? TsRing<true>::ROWS : TsRing<false>::ROWS, &sm->shift[rf])) return false;
    }
    return true;
}
// Two-pass positions on the tensor cores (pred_umma.cuh).  Default: 8-tap positions of batches that fill the chip at least once
// (the kernels are persistent, one CTA per SM, with tiles of 128 x 192 samples); HEVCASM_PRED_HV=umma takes them for
// every size and both filters, HEVCASM_PRED_HV=stream never (A/B runs and the parity tests of either side).
// Measured crossover on 4K planes (us per launch, tensor / streaming): one reference 12.0 / 12.1 (1 plane), 19.3 / 23.3 (4 planes);
// two references 18.1 / 15.1 (1), 24.6 / 23.6 (2), 36.6 / 39.8 (4): the two-reference kernel needs ~5 tiles per SM to pay off.
static bool tensor_path_wanted(int taps, long long n_tiles, bool bi = false)
{
    const char *pin = tune::knob("HEVCASM_PRED_HV");
    if (pin && !strcmp(pin, "umma")) return true;
    if (pin && !strcmp(pin, "stream")) return false;
    // two references: both filters (chroma 84.8 vs 105.3 us per 16 4K planes); one reference: 8-tap only (chroma 54.4 vs 52.9 us on the streaming kernel)
    return (taps == 8 || bi) && n_tiles >= (bi ? 5ll : 1ll) * sm_count();
}
static unsigned tensor_grid(long long n_tiles)
{
    long long g = std::min<long long>(n_tiles, (long long)sm_count());   // one persistent CTA per SM
    if (const char *e = tune::knob("HEVCASM_PRED_UMMA_GRID")) g = std::max(1ll, std::min<long long>(g, atoll(e)));   // test knob: more tiles per CTA
    return (unsigned)g;
}
#ifdef HEVCASM_EXPERIMENTS
// needs 16-byte aligned reference planes / strides
static bool umma_params(um::Params *u, const PredParams &p, int taps, bool bi, int n_frames)
{
    if (!tma::describable(p.sr, p.fs_ref, n_frames)) return false;
    const int top = taps / 2 - 1;
    const long long rows = (long long)p.height + taps - 1;
    const long long ext_x = 16 + (long long)p.width + taps / 2;   // bytes of a row the filter footprints touch, from x = -16
    const long long per = (long long)((p.width + um::TCOLS - 1) / um::TCOLS) * ((p.height + um::TROWS - 1) / um::TROWS);
    if (per * n_frames >= (1ll << 31) || !tensor_path_wanted(taps, per * n_frames)) return false;
    const uint8_t *refs[2] = {p.ref0, p.ref1};
    const int xf[2] = {p.xf0, p.xf1}, yf[2] = {p.yf0, p.yf1};
    for (int rf = 0; rf < (bi ? 2 : 1); ++rf) {
        const uint8_t *base = refs[rf] - (ptrdiff_t)top * p.sr - 16;
        if (tma::describe_u8_swizzled(&u->tm128[rf], base, p.sr, p.fs_ref, ext_x, rows, n_frames, 128, um::NROWS) ||
            tma::describe_u8_swizzled(&u->tm32[rf], base, p.sr, p.fs_ref, ext_x, rows, n_frames, 32, um::NROWS))
            return false;
        const PackedCoefs c = pack_coefs(taps, xf[rf], yf[rf]);
        for (int g = 0; g < 4; ++g) u->y2[rf][g] = c.y2[g];
        for (int k = 0; k < 8; ++k) u->xtap[rf][k] = k < taps ? (int8_t)((c.x4[k >> 2] >> (8 * (k & 3))) & 0xff) : 0;
    }
    u->dst = p.dst, u->sd = p.sd, u->fs_dst = p.fs_dst, u->width = p.width, u->height = p.height;
    u->dst16 = (((uintptr_t)p.dst | (uintptr_t)p.sd | (n_frames > 1 ? (uintptr_t)p.fs_dst : 0)) & 15) == 0 && p.sd > 0 && (n_frames <= 1 || p.fs_dst > 0);
    if (u->dst16) {
        int shift = 0;
        if (tma::describe_u8(&u->tmdst, p.dst, p.sd, p.fs_dst, p.width, p.height, n_frames, um::TCOLS, um::TROWS, &shift) || shift) u->dst16 = 0;
    }
    u->tiles_x = (p.width + um::TCOLS - 1) / um::TCOLS, u->tiles_y = (p.height + um::TROWS - 1) / um::TROWS;
    u->n_tiles = (int)(per * n_frames);
    return true;
}
template <int TAPS, bool BI>
static int launch_umma(const um::Params &u, void *stream)
{
    using G = um::Geom<TAPS, BI>;
    auto kern = um::pred_umma_kernel<TAPS, BI>;
    if (set_max_smem(kern, G::SMEM_BYTES)) return (int)cudaErrorInvalidValue;
    const unsigned grid = tensor_grid((u.n_tiles + G::NWG - 1) / G::NWG);   // one CTA of NWG warpgroups per SM
    return launch(kern, dim3(grid), dim3(G::THREADS), (size_t)G::SMEM_BYTES, stream, u);
}
#endif
// vertical pass on the tensor cores (namespace uv of pred_umma.cuh), one or two references
template <int TAPS, bool BI>
static bool vh_params(uv::Params *u, const PredParams &p, int n_frames)
{
    if (!tma::describable(p.sr, p.fs_ref, n_frames)) return false;
    const int top = TAPS / 2 - 1;
    const long long rows = (long long)p.height + TAPS - 1;
    const long long ext_x = 16 + (long long)p.width + TAPS / 2;   // bytes of a row the filter footprints touch, from x = -16
    u->tiles_x = (p.width + uv::TCOLS - 1) / uv::TCOLS, u->tiles_y = (p.height + uv::TROWS - 1) / uv::TROWS;
    const long long per = (long long)u->tiles_x * u->tiles_y;
    if (per * n_frames >= (1ll << 30) || !tensor_path_wanted(TAPS, per * n_frames, BI)) return false;
    const uint8_t *refs[2] = {p.ref0, p.ref1};
    const int xf[2] = {p.xf0, p.xf1}, yf[2] = {p.yf0, p.yf1};
    for (int rf = 0; rf < (BI ? 2 : 1); ++rf) {
        if (tma::describe_u32_swizzled128(&u->tmref[rf], refs[rf] - (ptrdiff_t)top * p.sr - 16, p.sr, p.fs_ref, ext_x, rows, n_frames, uv::BOXR)) return false;
        const PackedCoefs c = pack_coefs(TAPS, xf[rf], yf[rf]);
        for (int g = 0; g < 4; ++g) u->x2[rf][g] = c.x2e[g], u->x2o[rf][g] = g < 3 ? c.x2o[g + 1] : 0;
        u->xfrac[rf] = xf[rf];
        for (int k = 0; k < 8; ++k) u->ytap[rf][k] = k < TAPS ? (int8_t)((c.y4s[0][k >> 2] >> (8 * (k & 3))) & 0xff) : 0;
    }
    u->dst = p.dst, u->sd = p.sd, u->fs_dst = p.fs_dst, u->width = p.width, u->height = p.height;
    u->dst16 = (((uintptr_t)p.dst | (uintptr_t)p.sd | (n_frames > 1 ? (uintptr_t)p.fs_dst : 0)) & 15) == 0 && p.sd > 0 && (n_frames <= 1 || p.fs_dst > 0);
    if (u->dst16 && (tma::describe_u8_swizzled(&u->tmdst[0], p.dst, p.sd, p.fs_dst, p.width, p.height, n_frames, 128, uv::TROWS) ||
                     tma::describe_u8_swizzled(&u->tmdst[1], p.dst, p.sd, p.fs_dst, p.width, p.height, n_frames, 64, uv::TROWS)))
        u->dst16 = 0;
    u->n_tiles = (int)(per * n_frames);
    return true;
}
template <int TAPS, bool BI>
static int launch_vh(uv::Params &u, void *stream)
{
    auto kern = uv::pred_vh_kernel<TAPS, BI>;
    if (set_max_smem(kern, uv::Geom<BI>::SMEM_BYTES)) return (int)cudaErrorInvalidValue;
    const unsigned grid = tensor_grid(u.n_tiles);
    u.prof = nullptr;
#ifdef HEVCASM_EXPERIMENTS
    if (tune::knob("HEVCASM_PRED_PROF")) {   // per-phase clock totals of CTA 0, printed after a synchronous launch (development aid)
        static long long *buf = nullptr;
        if (!buf) cudaMalloc(&buf, 32 * sizeof(long long));
        cudaMemset(buf, 0, 32 * sizeof(long long));
        u.prof = buf;
        const int e = launch(kern, dim3(grid), dim3(uv::THREADS), (size_t)uv::Geom<BI>::SMEM_BYTES, stream, u);
        long long h[32];
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost);
        const int tiles = (u.n_tiles + (int)grid - 1) / (int)grid;
        fprintf(stderr, "pred_vh prof, cycles per tile (%d tiles): loader wait_empty %lld request %lld | mma wait_ofree+full %lld wait_consumed %lld issue+commit %lld (%lld) | "
                        "storer wait_ready %lld issue %lld wait_read %lld | consumer", tiles, h[0] / tiles, h[1] / tiles, h[8] / tiles, h[9] / tiles, h[10] / tiles, h[11] / tiles,
                h[16] / tiles, h[17] / tiles, h[18] / tiles);
        for (int k = 0; k < 5; ++k) fprintf(stderr, " %lld", h[24 + k] / tiles);
        fprintf(stderr, "\n");
        return e;
    }
#endif
    return launch(kern, dim3(grid), dim3(uv::THREADS), (size_t)uv::Geom<BI>::SMEM_BYTES, stream, u);
}
static bool aligned8(const uint8_t *dst, ptrdiff_t sd, ptrdiff_t fs_dst, int n_frames)
{
    uintptr_t m = (uintptr_t)dst | (uintptr_t)sd;
    if (n_frames > 1) m |= (uintptr_t)fs_dst;
    return (m & 7) == 0;
}

// `exact`: only the kernel that stays inside the reference's own footprint (rounded out to aligned 32-bit words) may run - the caller
// has not promised the 16 readable bytes around it that the aligned, TMA and tensor-core kernels assume (hevcasm_batch.h)
static int pred_uni_frames_impl(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref, ptrdiff_t sr, int width, int height, int taps, int xFrac, int yFrac,
                                int n_frames, ptrdiff_t fs_dst, ptrdiff_t fs_ref, bool exact, void *stream)
{
    if ((taps != 8 && taps != 4) || !frac_ok(taps, xFrac) || !frac_ok(taps, yFrac) || width < 0 || height < 0 || n_frames < 0) return HEVCASM_ERR_ARGUMENT;
    if (width == 0 || height == 0 || n_frames == 0) return 0;
    PredParams p{};
    p.dst = dst, p.ref0 = ref, p.sd = sd, p.sr = sr, p.fs_dst = fs_dst, p.fs_ref = fs_ref, p.width = width, p.height = height;
    p.xf0 = xFrac, p.yf0 = yFrac;
    const dim3 grid((width + PTW - 1) / PTW, (height + PTH - 1) / PTH, n_frames);
    const int mode = (xFrac ? 1 : 0) | (yFrac ? 2 : 0);
    if (exact) {   // the reference's own footprint over the whole batch: rows -(taps/2-1) .. height+taps/2-1, columns likewise, of every frame
        // (a pass that is not run reads no halo: reference pred_inter.c:141-228 picks copy / H / V / HV by the fractions)
        const int bx = xFrac ? taps / 2 - 1 : 0, ax = xFrac ? taps / 2 : 0, by = yFrac ? taps / 2 - 1 : 0, ay = yFrac ? taps / 2 : 0;
        p.lo[0] = ref - (ptrdiff_t)by * sr - bx;
        p.hi[0] = ref + (ptrdiff_t)(n_frames - 1) * fs_ref + (ptrdiff_t)(height - 1 + ay) * sr + width + ax;
        return taps == 8 ? launch_uni_planes<8>(p, grid, mode, stream) : launch_uni_planes<4>(p, grid, mode, stream);
    }
    // every position: the TMA-fed streaming kernel when the planes can be described to the TMA unit (strides multiples
    // of 16, 4-byte aligned rows).  Otherwise: two-pass positions -> LDG streaming kernel; one-pass positions -> tile kernels on
    // 16-byte aligned planes (2.2-2.4 vs 2.0 Tsamples/s), LDG streaming kernel on 4-byte aligned ones; copies -> tile kernels.
    const char *pin = tune::knob("HEVCASM_PRED_PATH");
    const bool tile_ok = planes_fast_ok(ref, nullptr, sr, fs_ref, n_frames) && !(pin && !strcmp(pin, "stream"));
    if (mode == HV) {
        uv::Params u;
        if (taps == 8 ? vh_params<8, false>(&u, p, n_frames) : vh_params<4, false>(&u, p, n_frames))
            return taps == 8 ? launch_vh<8, false>(u, stream) : launch_vh<4, false>(u, stream);
    }
    if (stream_ok(dst, sd, fs_dst, ref, nullptr, sr, fs_ref, n_frames)) {
        FastParams fp{};
        fp.p = p, fp.c[0] = pack_coefs(taps, xFrac, yFrac);
        StreamMaps sm;
        if (stream_maps(&sm, p, taps, mode, false, n_frames))
            return taps == 8 ? launch_uni_stream_tma<8>(fp, sm, mode, n_frames, stream) : launch_uni_stream_tma<4>(fp, sm, mode, n_frames, stream);
        if (mode != COPY && (mode == HV || !tile_ok)) return taps == 8 ? launch_uni_stream<8>(fp, mode, n_frames, stream) : launch_uni_stream<4>(fp, mode, n_frames, stream);
    }
    if (planes_fast_ok(ref, nullptr, sr, fs_ref, n_frames)) {
        const bool dst8 = aligned8(dst, sd, fs_dst, n_frames);
        FastParams fp{};
        fp.p = p, fp.c[0] = pack_coefs(taps, xFrac, yFrac);
        const dim3 fgrid((width + FTW - 1) / FTW, (height + FTH - 1) / FTH, n_frames);
        return taps == 8 ? launch_uni_planes_fast<8>(fp, fgrid, mode, dst8, stream) : launch_uni_planes_fast<4>(fp, fgrid, mode, dst8, stream);
    }
    return taps == 8 ? launch_uni_planes<8>(p, grid, mode, stream) : launch_uni_planes<4>(p, grid, mode, stream);
}

extern "C" int hevcasm_pred_uni_frames(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref, ptrdiff_t sr, int width, int height, int taps, int xFrac, int yFrac,
                                       int n_frames, ptrdiff_t fs_dst, ptrdiff_t fs_ref, void *stream)
{
    return pred_uni_frames_impl(dst, sd, ref, sr, width, height, taps, xFrac, yFrac, n_frames, fs_dst, fs_ref, false, stream);
}

extern "C" int hevcasm_pred_uni_frames_bounded(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref, ptrdiff_t sr, int width, int height, int taps, int xFrac,
                                               int yFrac, int n_frames, ptrdiff_t fs_dst, ptrdiff_t fs_ref, ptrdiff_t slack_before, ptrdiff_t slack_after,
                                               void *stream)
{
    if (slack_before < 0 || slack_after < 0) return HEVCASM_ERR_ARGUMENT;
    return pred_uni_frames_impl(dst, sd, ref, sr, width, height, taps, xFrac, yFrac, n_frames, fs_dst, fs_ref, slack_before < 16 || slack_after < 16, stream);
}

static int pred_bi_frames_impl(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref0, const uint8_t *ref1, ptrdiff_t sr, int width, int height, int taps,
                               int xFrac0, int yFrac0, int xFrac1, int yFrac1, int n_frames, ptrdiff_t fs_dst, ptrdiff_t fs_ref, bool exact, void *stream)
{
    if ((taps != 8 && taps != 4) || !frac_ok(taps, xFrac0) || !frac_ok(taps, yFrac0) || !frac_ok(taps, xFrac1) || !frac_ok(taps, yFrac1) || width < 0 ||
        height < 0 || n_frames < 0)
        return HEVCASM_ERR_ARGUMENT;
    if (width == 0 || height == 0 || n_frames == 0) return 0;
    PredParams p{};
    p.dst = dst, p.ref0 = ref0, p.ref1 = ref1, p.sd = sd, p.sr = sr, p.fs_dst = fs_dst, p.fs_ref = fs_ref, p.width = width, p.height = height;
    p.xf0 = xFrac0, p.yf0 = yFrac0, p.xf1 = xFrac1, p.yf1 = yFrac1;
    const dim3 grid((width + PTW - 1) / PTW, (height + PTH - 1) / PTH, n_frames);
    if (exact) {
        const int b = taps / 2 - 1, a = taps / 2;
        const uint8_t *refs[2] = {ref0, ref1};
        for (int r = 0; r < 2; ++r) {
            p.lo[r] = refs[r] - (ptrdiff_t)b * sr - b;
            p.hi[r] = refs[r] + (ptrdiff_t)(n_frames - 1) * fs_ref + (ptrdiff_t)(height - 1 + a) * sr + width + a;
        }
        return taps == 8 ? launch_pred<8, PTW, PTH, true, RUNTIME>(p, grid, stream) : launch_pred<4, PTW, PTH, true, RUNTIME>(p, grid, stream);
    }
    if (xFrac0 || yFrac0 || xFrac1 || yFrac1) {
#ifdef HEVCASM_EXPERIMENTS
        const char *bk = tune::knob("HEVCASM_PRED_BI");   // A/B: "hfirst" = horizontal pass on the tensor cores (um), default = vertical pass (uv)
        if (bk && !strcmp(bk, "hfirst")) {
            um::Params u;
            if (umma_params(&u, p, taps, true, n_frames)) return taps == 8 ? launch_umma<8, true>(u, stream) : launch_umma<4, true>(u, stream);
        } else
#endif
        {
            uv::Params u;
            if (taps == 8 ? vh_params<8, true>(&u, p, n_frames) : vh_params<4, true>(&u, p, n_frames))
                return taps == 8 ? launch_vh<8, true>(u, stream) : launch_vh<4, true>(u, stream);
        }
    }
    if (stream_ok(dst, sd, fs_dst, ref0, ref1, sr, fs_ref, n_frames)) {
        FastParams fp{};
        fp.p = p, fp.c[0] = pack_coefs(taps, xFrac0, yFrac0), fp.c[1] = pack_coefs(taps, xFrac1, yFrac1);
        StreamMaps sm;
        // both references at a full-sample position: (64a + 64b + 64) >> 7 = (a + b + 1) >> 1, a byte average (the reference's own
        // assembly has the same shortcut, pred_inter_a.asm:580-608); anything else runs both passes on both references as the C does
        const bool avg = !xFrac0 && !yFrac0 && !xFrac1 && !yFrac1;
        if (stream_maps(&sm, p, taps, avg ? COPY : HV, true, n_frames)) {
            if (avg) return taps == 8 ? launch_stream_tma<8, COPY, true>(fp, sm, n_frames, stream) : launch_stream_tma<4, COPY, true>(fp, sm, n_frames, stream);
            return taps == 8 ? launch_stream_tma<8, HV, true>(fp, sm, n_frames, stream) : launch_stream_tma<4, HV, true>(fp, sm, n_frames, stream);
        }
        return taps == 8 ? launch_stream<8, HV, true>(fp, n_frames, stream) : launch_stream<4, HV, true>(fp, n_frames, stream);
    }
    if (planes_fast_ok(ref0, ref1, sr, fs_ref, n_frames)) {
        const bool dst8 = aligned8(dst, sd, fs_dst, n_frames);
        FastParams fp{};
        fp.p = p, fp.c[0] = pack_coefs(taps, xFrac0, yFrac0), fp.c[1] = pack_coefs(taps, xFrac1, yFrac1);
        const dim3 fgrid((width + FTW - 1) / FTW, (height + FTH - 1) / FTH, n_frames);
        return taps == 8 ? launch_plane_fast<8, HV, true>(fp, fgrid, dst8, stream) : launch_plane_fast<4, HV, true>(fp, fgrid, dst8, stream);
    }
    return taps == 8 ? launch_pred<8, PTW, PTH, true, RUNTIME>(p, grid, stream) : launch_pred<4, PTW, PTH, true, RUNTIME>(p, grid, stream);
}

extern "C" int hevcasm_pred_bi_frames(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref0, const uint8_t *ref1, ptrdiff_t sr, int width, int height, int taps,
                                      int xFrac0, int yFrac0, int xFrac1, int yFrac1, int n_frames, ptrdiff_t fs_dst, ptrdiff_t fs_ref, void *stream)
{
    return pred_bi_frames_impl(dst, sd, ref0, ref1, sr, width, height, taps, xFrac0, yFrac0, xFrac1, yFrac1, n_frames, fs_dst, fs_ref, false, stream);
}

extern "C" int hevcasm_pred_bi_frames_bounded(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref0, const uint8_t *ref1, ptrdiff_t sr, int width, int height,
                                              int taps, int xFrac0, int yFrac0, int xFrac1, int yFrac1, int n_frames, ptrdiff_t fs_dst, ptrdiff_t fs_ref,
                                              ptrdiff_t slack_before, ptrdiff_t slack_after, void *stream)
{
    if (slack_before < 0 || slack_after < 0) return HEVCASM_ERR_ARGUMENT;
    return pred_bi_frames_impl(dst, sd, ref0, ref1, sr, width, height, taps, xFrac0, yFrac0, xFrac1, yFrac1, n_frames, fs_dst, fs_ref,
                               slack_before < 16 || slack_after < 16, stream);
}

extern "C" int hevcasm_pred_uni_batch(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref, ptrdiff_t sr, int taps, const int16_t *pus, int n_pu, void *stream)
{
    if ((taps != 8 && taps != 4) || n_pu < 0 || (n_pu > 0 && !pus)) return HEVCASM_ERR_ARGUMENT;
    if (n_pu == 0) return 0;
    PredParams p{};
    p.dst = dst, p.ref0 = ref, p.sd = sd, p.sr = sr, p.pus = pus, p.n_pu = n_pu;
    if (list_stream_ok()) {
        return launch_list<false>(taps, stream, p);
    }
    return taps == 8 ? launch_pred<8, LTW, LTH, false, RUNTIME>(p, dim3(n_pu), stream) : launch_pred<4, LTW, LTH, false, RUNTIME>(p, dim3(n_pu), stream);
}

extern "C" int hevcasm_pred_bi_batch(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref0, const uint8_t *ref1, ptrdiff_t sr, int taps, const int16_t *pus, int n_pu,
                                     void *stream)
{
    if ((taps != 8 && taps != 4) || n_pu < 0 || (n_pu > 0 && !pus)) return HEVCASM_ERR_ARGUMENT;
    if (n_pu == 0) return 0;
    PredParams p{};
    p.dst = dst, p.ref0 = ref0, p.ref1 = ref1, p.sd = sd, p.sr = sr, p.pus = pus, p.n_pu = n_pu;
    if (list_stream_ok()) {
        return launch_list<true>(taps, stream, p);
    }
    return taps == 8 ? launch_pred<8, LTW, LTH, true, RUNTIME>(p, dim3(n_pu), stream) : launch_pred<4, LTW, LTH, true, RUNTIME>(p, dim3(n_pu), stream);
}

// PU lists over a batch of frames in one launch: descriptors carry a trailing frame index
extern "C" int hevcasm_pred_uni_list_frames(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref, ptrdiff_t sr, int taps, const int16_t *pus, int n_pu, ptrdiff_t fs_dst,
                                            ptrdiff_t fs_ref, void *stream)
{
    if ((taps != 8 && taps != 4) || n_pu < 0 || (n_pu > 0 && !pus)) return HEVCASM_ERR_ARGUMENT;
    if (n_pu == 0) return 0;
    PredParams p{};
    p.dst = dst, p.ref0 = ref, p.sd = sd, p.sr = sr, p.fs_dst = fs_dst, p.fs_ref = fs_ref, p.pus = pus, p.n_pu = n_pu, p.desc_frame = 1;
    return launch_list<false>(taps, stream, p);
}

extern "C" int hevcasm_pred_bi_list_frames(uint8_t *dst, ptrdiff_t sd, const uint8_t *ref0, const uint8_t *ref1, ptrdiff_t sr, int taps, const int16_t *pus, int n_pu,
                                           ptrdiff_t fs_dst, ptrdiff_t fs_ref, void *stream)
{
    if ((taps != 8 && taps != 4) || n_pu < 0 || (n_pu > 0 && !pus)) return HEVCASM_ERR_ARGUMENT;
    if (n_pu == 0) return 0;
    PredParams p{};
    p.dst = dst, p.ref0 = ref0, p.ref1 = ref1, p.sd = sd, p.sr = sr, p.fs_dst = fs_dst, p.fs_ref = fs_ref, p.pus = pus, p.n_pu = n_pu, p.desc_frame = 1;
    return launch_list<true>(taps, stream, p);
}
