// hevcasm_b200 - inverse 16x16 / 32x32 transform + add on the integer tensor cores (IMMA, mma.sync s8/u8 -> s32).
//
// The north star allows tensor-core transforms only as an exact int32-accumulating formulation that beats the CUDA-core
// butterfly in ncu.  This is that formulation; profiles/r01_transforms.md has the comparison.
//
// Exactness.  A stage is  out = clip16((sum_k T[k][.] * x[k] + round) >> shift)  with |T| <= 90 (fits s8) and x int16.
// Write x = 256*hi + lo with hi = x >> 8 (s8) and lo = x & 255 (u8): sum T*x = 256 * sum T*hi + sum T*lo, two s8 x s8 /
// s8 x u8 matrix products accumulated in int32 (|256 * sum T*hi| <= 256*32*90*128 < 2^31), so the tensor-core result is
// the same integer the reference's C computes; rounding, shift and clip are then applied exactly as in
// residual_decode.c:69-347.
//
// Data flow for one N x N block, one warp (g = lane / 4, t = lane % 4 as in the PTX fragment tables):
//   stage 1 (contracts the vertical frequency v, the SLOW index of coeffs[v][u]):  Bt[y][u] = sum_v T[v][y] * C[v][u]
//       A = T^T (constant, registers), B = C: a B fragment wants 4 consecutive k per thread at one column, i.e. 4 values
//       a row apart - fetched with ldmatrix.trans from a padded shared tile, which yields (v, v+1) pairs; the contraction
//       index is therefore visited in the order  slot 4t+j -> v = {2t, 2t+1, 8+2t, 9+2t}[j] (+16 for the upper half)
//       and the constant A fragments are built with the same permutation.
//   stage 2 (contracts u): R[y][x] = sum_u Bt[y][u] * T[u][x].  The D fragments of stage 1 hold Bt[y = g, g+8][u = 2t, 2t+1
//       (+8 per n-tile)] - exactly the rows and (permuted) k slots this thread's A fragment needs, so the intermediate never
//       leaves registers; B = T (constant) with the same slot permutation.
//   epilogue: R (int16) through the shared tile so that every lane adds a whole predictor row segment and stores 16 bytes.
#pragma once

#include "transform.cuh"

namespace hv {
namespace tr {

__constant__ int8_t c_T32[32][32];   // filled by imma_tables_init(): T_32[k][x]
__constant__ int8_t c_T16[16][16];

static int imma_tables_init()
{
    static int done = [] {
        int8_t t32[32][32], t16[16][16];
        for (int k = 0; k < 32; ++k)
            for (int x = 0; x < 32; ++x) t32[k][x] = (int8_t)dct(32, k, x);
        for (int k = 0; k < 16; ++k)
            for (int x = 0; x < 16; ++x) t16[k][x] = (int8_t)dct(16, k, x);
        cudaError_t e = cudaMemcpyToSymbol(c_T32, t32, sizeof t32);
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_T16, t16, sizeof t16);
        return (int)e;
    }();
    return done;
}

// contraction-slot permutation inside a 16-wide half: slot 4t+j -> {2t, 2t+1, 8+2t, 9+2t}[j]
__device__ __forceinline__ int kperm(int slot)
{
    const int half = slot & 16, s = slot & 15, t = s >> 2, j = s & 3;
    return half + (j < 2 ? 2 * t + j : 8 + 2 * t + (j - 2));
}

template <int N>
__device__ __forceinline__ int tcoef(int k, int x) { return N == 32 ? (int)c_T32[k][x] : (int)c_T16[k][x]; }

__device__ __forceinline__ uint32_t pack4(int b0, int b1, int b2, int b3)
{
    return (uint32_t)(b0 & 0xff) | ((uint32_t)(b1 & 0xff) << 8) | ((uint32_t)(b2 & 0xff) << 16) | ((uint32_t)(b3 & 0xff) << 24);
}

// D = A(16xK, AT) * B(Kx8, BT) + C, K = 32 or 16
#define HV_MMA_K32(AT, BT)                                                                                                              \
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32." AT "." BT ".s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"      \
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])                                                                       \
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]))
#define HV_MMA_K16(AT, BT)                                                                                                   \
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32." AT "." BT ".s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%8,%9,%10};"       \
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])                                                            \
                 : "r"(a[0]), "r"(a[1]), "r"(b[0]), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]))

template <int N>
struct Imma {
    static constexpr int AR = N == 32 ? 4 : 2, BR = N == 32 ? 2 : 1;  // registers per A / B fragment
    __device__ static __forceinline__ void s8s8(int (&d)[4], const uint32_t (&a)[AR], const uint32_t (&b)[BR], const int (&c)[4])
    {
        if constexpr (N == 32) HV_MMA_K32("s8", "s8");
        else HV_MMA_K16("s8", "s8");
    }
    __device__ static __forceinline__ void s8u8(int (&d)[4], const uint32_t (&a)[AR], const uint32_t (&b)[BR], const int (&c)[4])
    {
        if constexpr (N == 32) HV_MMA_K32("s8", "u8");
        else HV_MMA_K16("s8", "u8");
    }
    __device__ static __forceinline__ void u8s8(int (&d)[4], const uint32_t (&a)[AR], const uint32_t (&b)[BR], const int (&c)[4])
    {
        if constexpr (N == 32) HV_MMA_K32("u8", "s8");
        else HV_MMA_K16("u8", "s8");
    }
};

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void *smem_row)
{
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

}  // namespace tr
}  // namespace hv
