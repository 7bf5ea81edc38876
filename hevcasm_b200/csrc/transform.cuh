// hevcasm_b200 - exact integer HEVC transform building blocks (device side).
//
// Reference semantics (kupix/hevcasm residual_decode.c): forward stages :592-852 (transpose((T*row + 2^(s-1)) >> s),
// truncating int16 store), inverse stages :69-347 (transposing, clip to int16), matrices :623-629 / :662-672 /
// :719-735 / :795-826 and the 4x4 DST expressions :69-88, :592-611.
//
// A 1-D stage is an exact integer matrix product followed by one rounding shift, and no partial sum can leave 32
// bits (|sum| <= 32*90*32768 < 2^31), so any association order gives identical bits.  That freedom is used here:
//   * inverse (inputs are int16): even/odd partial butterflies evaluated with IDP.2A (two 16-bit x 8-bit MACs per
//     instruction; every HEVC coefficient fits s8); operands are int16 pairs arranged in "butterfly order";
//   * forward N <= 8 and the DST: plain matrix form with IDP.2A straight on the int16 pairs as they sit in memory;
//   * forward N >= 16: even/odd partial butterflies in 32-bit IMAD (the first-level sums need 17+ bits).
// All coefficients are compile-time constants generated from the first column of the 32-point matrix.
#pragma once

#include "common.cuh"

namespace hv {
namespace tr {

// first column of the 32-point HEVC matrix: 64*sqrt(2)*cos(m*pi/64) in HEVC's integer approximation
__host__ __device__ constexpr int col32(int m)
{
    constexpr int c[33] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0};
    return c[m];
}

// T_N[k][x] by cosine symmetry
__host__ __device__ constexpr int dct(int N, int k, int x)
{
    int m = (k * (32 / N) * (2 * x + 1)) % 128;
    if (m > 64) m = 128 - m;
    return m > 32 ? -col32(64 - m) : col32(m);
}

// forward 4x4 DST-VII matrix M[k][x]
__host__ __device__ constexpr int dst4(int k, int x)
{
    constexpr int m[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};
    return m[k][x];
}

// two s8 coefficients in the low half of an IDP.2A "b" operand
__host__ __device__ constexpr int cpair(int a, int b) { return (a & 0xff) | ((b & 0xff) << 8); }

// ---- coefficient pairs in constant memory ------------------------------------------------------------------
// IDP takes its coefficient operand from a (uniform) register, never as an immediate: with the pairs written as
// compile-time immediates ptxas materialises every one with its own UMOV - 20 % of all issued instructions in the 32x32
// inverse (profiles/r01_transforms.md).  Stored in constant memory IN THE ORDER THE BUTTERFLIES CONSUME THEM they arrive
// four at a time (LDCU.128).
struct InvTab {
    // inverse partial butterflies: for N = 32, 16, 8, 4 the odd-part pairs [k][j] = (T_N[4j+1][k], T_N[4j+3][k]), k < N/2, j < N/4
    int v[16 * 8 + 8 * 4 + 4 * 2 + 2 * 1];
    int two[2];     // (64, 64), (64, -64)
    int dst[4][2];  // inverse 4x4 DST: [x][h] = (M[2h][x], M[2h+1][x])
};
__host__ __device__ constexpr int inv_tab_offset(int N) { return N == 32 ? 0 : N == 16 ? 128 : N == 8 ? 160 : 168; }
constexpr InvTab make_inv_tab()
{
    InvTab t{};
    for (int N = 32; N >= 4; N /= 2)
        for (int k = 0; k < N / 2; ++k)
            for (int j = 0; j < N / 4; ++j) t.v[inv_tab_offset(N) + k * (N / 4) + j] = cpair(dct(N, 4 * j + 1, k), dct(N, 4 * j + 3, k));
    t.two[0] = cpair(64, 64), t.two[1] = cpair(64, -64);
    for (int x = 0; x < 4; ++x) t.dst[x][0] = cpair(dst4(0, x), dst4(1, x)), t.dst[x][1] = cpair(dst4(2, x), dst4(3, x));
    return t;
}
__constant__ InvTab c_inv = make_inv_tab();

// forward matrix form (4x4 DCT, 4x4 DST, 8x8 DCT): [u][j] = (T[u][2j], T[u][2j+1])
struct FwdTab {
    int dct4[4][2], dst4[4][2], dct8[8][4];
};
constexpr FwdTab make_fwd_tab()
{
    FwdTab t{};
    for (int u = 0; u < 4; ++u)
        for (int j = 0; j < 2; ++j) t.dct4[u][j] = cpair(dct(4, u, 2 * j), dct(4, u, 2 * j + 1)), t.dst4[u][j] = cpair(dst4(u, 2 * j), dst4(u, 2 * j + 1));
    for (int u = 0; u < 8; ++u)
        for (int j = 0; j < 4; ++j) t.dct8[u][j] = cpair(dct(8, u, 2 * j), dct(8, u, 2 * j + 1));
    return t;
}
__constant__ FwdTab c_fwd = make_fwd_tab();

// the same matrices for BYTE inputs (dp4a: four u8 samples x four s8 coefficients): [u][j] = (T[u][4j] .. T[u][4j+3]); `neg` holds the negated
// coefficients, so that a row of src - pred is transformed as sum T*src + sum (-T)*pred without ever forming the 9-bit difference
struct FwdByteTab {
    int dct4[4], dst4[4], dct8[8][2], ndct4[4], ndst4[4], ndct8[8][2];
};
__host__ __device__ constexpr int cquad(int a, int b, int c, int d) { return (a & 0xff) | ((b & 0xff) << 8) | ((c & 0xff) << 16) | ((d & 0xff) << 24); }
constexpr FwdByteTab make_fwd_byte_tab()
{
    FwdByteTab t{};
    for (int u = 0; u < 4; ++u) {
        t.dct4[u] = cquad(dct(4, u, 0), dct(4, u, 1), dct(4, u, 2), dct(4, u, 3));
        t.ndct4[u] = cquad(-dct(4, u, 0), -dct(4, u, 1), -dct(4, u, 2), -dct(4, u, 3));
        t.dst4[u] = cquad(dst4(u, 0), dst4(u, 1), dst4(u, 2), dst4(u, 3));
        t.ndst4[u] = cquad(-dst4(u, 0), -dst4(u, 1), -dst4(u, 2), -dst4(u, 3));
    }
    for (int u = 0; u < 8; ++u)
        for (int j = 0; j < 2; ++j) {
            t.dct8[u][j] = cquad(dct(8, u, 4 * j), dct(8, u, 4 * j + 1), dct(8, u, 4 * j + 2), dct(8, u, 4 * j + 3));
            t.ndct8[u][j] = cquad(-dct(8, u, 4 * j), -dct(8, u, 4 * j + 1), -dct(8, u, 4 * j + 2), -dct(8, u, 4 * j + 3));
        }
    return t;
}
__constant__ FwdByteTab c_fwd_byte = make_fwd_byte_tab();

#define HV_V(ic) (decltype(ic)::value)

template <int I>
struct IC {
    static constexpr int value = I;
};

// compile-time counted loop: f(IC<I>{}) for I in [B, E)
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (B < E) {
        f(IC<B>{});
        static_for<B + 1, E>(f);
    }
}

// ---- butterfly pair order for the inverse -----------------------------------------------------------
// Pair i (0 <= i < N/2) of an N-point inverse holds input rows (row0, row1):
//   the N/4 pairs of odd rows (1,3) (5,7) ..., then recursively the pairs of the N/2-point transform of the even rows.
__host__ __device__ constexpr int pair_row(int N, int i, int which)
{
    int stride = 1;
    while (N > 2) {
        if (i < N / 4) return stride * (4 * i + 1 + 2 * which);
        i -= N / 4;
        N /= 2;
        stride *= 2;
    }
    return which ? stride : 0;
}

// out[k] = round + sum_v T_N[v][k] * in[v], k < N, with `p` = the int16 input pairs in butterfly order.
template <int N>
struct InvBfly {
    __device__ static __forceinline__ void run(const uint32_t *p, int *out, int round)
    {
        int E[N / 2];
        InvBfly<N / 2>::run(p + N / 4, E, round);
        static_for<0, N / 2>([&](auto k) {
            int o = 0;
            static_for<0, N / 4>([&](auto j) { o = dp2a_lo(p[HV_V(j)], c_inv.v[inv_tab_offset(N) + HV_V(k) * (N / 4) + HV_V(j)], o); });
            out[HV_V(k)] = E[HV_V(k)] + o;
            out[N - 1 - HV_V(k)] = E[HV_V(k)] - o;
        });
    }
};
template <>
struct InvBfly<2> {
    __device__ static __forceinline__ void run(const uint32_t *p, int *out, int round)
    {
        out[0] = dp2a_lo(p[0], c_inv.two[0], round);
        out[1] = dp2a_lo(p[0], c_inv.two[1], round);
    }
};

// Two transforms at once with the SAME coefficients (the two coefficient columns a thread owns): every coefficient pair is fetched once
// (LDCU + R2UR per IDP were 11 % of the 32x32 inverse's instructions, profiles/r02_transforms.md) and feeds two IDP.2A.
template <int N>
struct InvBfly2 {
    __device__ static __forceinline__ void run(const uint32_t *p0, const uint32_t *p1, int *out0, int *out1, int round)
    {
        int E0[N / 2], E1[N / 2];
        InvBfly2<N / 2>::run(p0 + N / 4, p1 + N / 4, E0, E1, round);
        static_for<0, N / 2>([&](auto k) {
            int a = 0, b = 0;
            static_for<0, N / 4>([&](auto j) {
                const int c = c_inv.v[inv_tab_offset(N) + HV_V(k) * (N / 4) + HV_V(j)];
                a = dp2a_lo(p0[HV_V(j)], c, a);
                b = dp2a_lo(p1[HV_V(j)], c, b);
            });
            out0[HV_V(k)] = E0[HV_V(k)] + a, out0[N - 1 - HV_V(k)] = E0[HV_V(k)] - a;
            out1[HV_V(k)] = E1[HV_V(k)] + b, out1[N - 1 - HV_V(k)] = E1[HV_V(k)] - b;
        });
    }
};
template <>
struct InvBfly2<2> {
    __device__ static __forceinline__ void run(const uint32_t *p0, const uint32_t *p1, int *out0, int *out1, int round)
    {
        out0[0] = dp2a_lo(p0[0], c_inv.two[0], round), out0[1] = dp2a_lo(p0[0], c_inv.two[1], round);
        out1[0] = dp2a_lo(p1[0], c_inv.two[0], round), out1[1] = dp2a_lo(p1[0], c_inv.two[1], round);
    }
};

// inverse 4-point DST: out[x] = round + sum_k M[k][x] * in[k]; p = natural pairs (in0,in1), (in2,in3)
__device__ __forceinline__ void inv_dst4(const uint32_t *p, int *out, int round)
{
    static_for<0, 4>([&](auto x) { out[HV_V(x)] = dp2a_lo(p[1], c_inv.dst[HV_V(x)][1], dp2a_lo(p[0], c_inv.dst[HV_V(x)][0], round)); });
}

// ---- forward, matrix form on natural int16 pairs (w[j] = (x[2j], x[2j+1])) ---------------------------
template <int N, bool DST>
__device__ __forceinline__ void fwd_matrix(const uint32_t *w, int *out, int round)
{
    static_for<0, N>([&](auto u) {
        int a = round;
        static_for<0, N / 2>([&](auto j) {
            const int c = DST ? c_fwd.dst4[HV_V(u) & 3][HV_V(j) & 1] : (N == 4 ? c_fwd.dct4[HV_V(u) & 3][HV_V(j) & 1] : c_fwd.dct8[HV_V(u) & 7][HV_V(j) & 3]);
            a = dp2a_lo(w[HV_V(j)], c, a);
        });
        out[HV_V(u)] = a;
    });
}

// ---- forward, matrix form on the BYTES of two 8-bit rows: out[u] = round + sum_x T[u][x] * (s[x] - p[x]) --------------
template <int N, bool DST>
__device__ __forceinline__ void fwd_matrix_bytes(const uint32_t *sw, const uint32_t *pw, int *out, int round)
{
    static_for<0, N>([&](auto u) {
        int a = round;
        static_for<0, N / 4>([&](auto j) {
            const int c = DST ? c_fwd_byte.dst4[HV_V(u) & 3] : (N == 4 ? c_fwd_byte.dct4[HV_V(u) & 3] : c_fwd_byte.dct8[HV_V(u) & 7][HV_V(j) & 1]);
            const int nc = DST ? c_fwd_byte.ndst4[HV_V(u) & 3] : (N == 4 ? c_fwd_byte.ndct4[HV_V(u) & 3] : c_fwd_byte.ndct8[HV_V(u) & 7][HV_V(j) & 1]);
            a = dp4a_us(pw[HV_V(j)], nc, dp4a_us(sw[HV_V(j)], c, a));
        });
        out[HV_V(u)] = a;
    });
}

// ---- forward, even/odd partial butterfly in 32-bit (any int inputs) -----------------------------------
// out[u] = round + sum_x T_N[u][x] * x[x]; OUT_STRIDE spreads the outputs of the recursive even part.
template <int N, int OUT_STRIDE = 1>
struct FwdBfly {
    __device__ static __forceinline__ void run(const int *x, int *out, int round)
    {
        int E[N / 2], O[N / 2];
        static_for<0, N / 2>([&](auto k) {
            E[HV_V(k)] = x[HV_V(k)] + x[N - 1 - HV_V(k)];
            O[HV_V(k)] = x[HV_V(k)] - x[N - 1 - HV_V(k)];
        });
        FwdBfly<N / 2, OUT_STRIDE * 2>::run(E, out, round);
        static_for<0, N / 2>([&](auto w) {
            int a = round;
            static_for<0, N / 2>([&](auto k) { a += dct(N, 2 * HV_V(w) + 1, HV_V(k)) * O[HV_V(k)]; });
            out[(2 * HV_V(w) + 1) * OUT_STRIDE] = a;
        });
    }
};
template <int OUT_STRIDE>
struct FwdBfly<2, OUT_STRIDE> {
    __device__ static __forceinline__ void run(const int *x, int *out, int round)
    {
        out[0] = 64 * (x[0] + x[1]) + round;
        out[OUT_STRIDE] = 64 * (x[0] - x[1]) + round;
    }
};

// ---- forward, the same butterfly with the odd part of its top level on IDP.2A -----------------------------
// The odd part of level M is an (M/2 x M/2) product on the differences O[k] = x[k] - x[M-1-k]: M/2 IMADs per output
// above.  When every input of the transform lies in [-16384, 16383] the differences fit int16, so they can be packed in
// pairs and each output costs M/4 IDP.2A instead: 128 IDP + 8 PRMT instead of 256 IMAD for a 32-point transform.  The
// integer result is the same - the caller tests the range (fits15) and uses FwdBfly for anything larger, so any int16
// input stays exact.  (Packing the next level too needs [-8192, 8191]: random 9-bit residuals leave that range after the
// first 32-point stage often enough that most warps would run both paths - measured 255 vs 184 us.)
struct FwdOddTab {
    int v[16 * 8 + 8 * 4 + 4 * 2];   // M = 32, 16, 8: [w][j] = (T_M[2w+1][2j], T_M[2w+1][2j+1]), w < M/2, j < M/4
};
__host__ __device__ constexpr int fwd_odd_offset(int M) { return M == 32 ? 0 : M == 16 ? 128 : 160; }
constexpr FwdOddTab make_fwd_odd_tab()
{
    FwdOddTab t{};
    for (int M = 32; M >= 8; M /= 2)
        for (int w = 0; w < M / 2; ++w)
            for (int j = 0; j < M / 4; ++j) t.v[fwd_odd_offset(M) + w * (M / 4) + j] = cpair(dct(M, 2 * w + 1, 2 * j), dct(M, 2 * w + 1, 2 * j + 1));
    return t;
}
__constant__ FwdOddTab c_fwd_odd = make_fwd_odd_tab();

template <int N, int OUT_STRIDE = 1, int LEVELS = 1>
struct FwdBflyPacked {
    __device__ static __forceinline__ void run(const int *x, int *out, int round)
    {
        if constexpr (LEVELS == 0 || N < 8) {
            FwdBfly<N, OUT_STRIDE>::run(x, out, round);
        } else {
            int E[N / 2];
            uint32_t Op[N / 4];
            static_for<0, N / 4>([&](auto j) {
                constexpr int k = 2 * HV_V(j);
                E[k] = x[k] + x[N - 1 - k], E[k + 1] = x[k + 1] + x[N - 2 - k];
                Op[HV_V(j)] = __byte_perm((uint32_t)(x[k] - x[N - 1 - k]), (uint32_t)(x[k + 1] - x[N - 2 - k]), 0x5410);
            });
            FwdBflyPacked<N / 2, OUT_STRIDE * 2, LEVELS - 1>::run(E, out, round);
            static_for<0, N / 2>([&](auto w) {
                int a = round;
                static_for<0, N / 4>([&](auto j) { a = dp2a_lo(Op[HV_V(j)], c_fwd_odd.v[fwd_odd_offset(N) + HV_V(w) * (N / 4) + HV_V(j)], a); });
                out[(2 * HV_V(w) + 1) * OUT_STRIDE] = a;
            });
        }
    }
};

// accumulates the range test for packed int16 pairs: after or-ing fits15_acc over all words, both halves of every word lie in
// [-16384, 16383] iff (acc & 0x80008000) == 0 (bits 15 and 14 of each half equal)
__device__ __forceinline__ uint32_t fits15_acc(uint32_t acc, uint32_t w) { return acc | (w ^ (w << 1)); }
__device__ __forceinline__ bool fits15(uint32_t acc) { return (acc & 0x80008000u) == 0; }

__host__ __device__ constexpr int fwd_shift1(int log2) { return log2 - 1; }   // 1, 2, 3, 4   (residual_decode.c:855-892)
__host__ __device__ constexpr int fwd_shift2(int log2) { return log2 + 6; }   // 8, 9, 10, 11

__device__ __forceinline__ int s16lo(uint32_t w) { return (int)(short)(w & 0xffffu); }
__device__ __forceinline__ int s16hi(uint32_t w) { return (int)w >> 16; }
__device__ __forceinline__ uint32_t lolo(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x5410); }  // (a.lo, b.lo)
__device__ __forceinline__ uint32_t hihi(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x7632); }  // (a.hi, b.hi)

}  // namespace tr
}  // namespace hv
