// hevcasm_b200 - shared device/host helpers for the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>

#include "hevcasm_batch.h"

namespace hv {

// ---- launch bookkeeping -------------------------------------------------------------------------
void count_launch();  // abi.cu

// Every kernel launch in the library goes through this so that (a) launches are counted for bench.py's
// gpu_launches and (b) launch errors are returned to the C caller instead of being swallowed.
#define HV_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                    \
        kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);           \
        ::hv::count_launch();                                                               \
        cudaError_t hv_e_ = cudaGetLastError();                                             \
        if (hv_e_ != cudaSuccess) return (int)hv_e_;                                        \
    } while (0)

// function form for template kernels (a template-id with commas cannot go through the macro)
template <class K, class... A>
inline int launch(K kernel, dim3 grid, dim3 block, size_t smem, void *stream, A... args)
{
    kernel<<<grid, block, smem, (cudaStream_t)stream>>>(args...);
    count_launch();
    return (int)cudaGetLastError();
}

#define HV_CUDA(call)                                       \
    do {                                                    \
        cudaError_t hv_e_ = (call);                         \
        if (hv_e_ != cudaSuccess) return (int)hv_e_;        \
    } while (0)

template <class K>
inline int set_max_smem(K kernel, size_t bytes)
{
    return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// ---- A/B switches ---------------------------------------------------------------------------------
// The product library chooses every kernel from the call's arguments alone: tune::knob() is a constant there, so the branches
// that test it fold away and no entry point reads the environment.  The experiments build (-DHEVCASM_EXPERIMENTS ->
// libhevcasm_b200_exp.so: the measured-but-not-adopted kernel variants plus these switches, used by tools/ and by the parity
// tests that pin one path) reads HEVCASM_* environment variables here.
namespace tune {
#ifdef HEVCASM_EXPERIMENTS
inline const char *knob(const char *name) { return getenv(name); }
#else
constexpr const char *knob(const char *) { return nullptr; }
#endif
}  // namespace tune

// SMs of the current device (148 on a B200); grids are sized against it
inline int sm_count()
{
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    return n;
}

// ---- small device primitives --------------------------------------------------------------------

// bytes [8*s .. 8*s+32) of the 64-bit value hi:lo  (s in 0..3 => byte-granular realignment of two words)
__device__ __forceinline__ uint32_t shr_bytes(uint32_t lo, uint32_t hi, int s) { return __funnelshift_r(lo, hi, 8 * s); }

// the same value computed on the FMA pipe: (lo >> 8s) + (hi << (32 - 8s)) = hi32(lo * C) + hi * C with C = 2^(32-8s), two IMADs.
// The SAD kernels saturate the ALU pipe (VABSDIFF4 + SHF share it) while the FMA pipe idles, so moving the realignment
// shifts there raises the VABSDIFF4 issue rate (profiles/r01_sad_pyramid.md).  C must reach the kernel as a run-time value
// (kernel parameter): as a literal ptxas strength-reduces the pair back into a shift + LEA.HI on the ALU pipe.
__device__ __forceinline__ uint32_t shr_bytes_fma(uint32_t lo, uint32_t hi, uint32_t c)
{
    uint32_t t, d;
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(hi), "r"(c));
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(c), "r"(t));
    return d;
}

// acc + sum of |a.b[i] - b.b[i]| over the four bytes : one VABSDIFF4.U8.ACC
// (inline PTX on purpose: written as __vsadu4(a,b)+acc the compiler re-associates pairs into 2 x VABSDIFF4 + IADD3)
__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t acc)
{
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}

// acc + sum a.u8[i] * b.s8[i] : IDP.4A.U8.S8
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int acc)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}
// acc + sum a.u8[i] * b.u8[i] : IDP.4A.U8.U8
__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t acc)
{
    uint32_t d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}
// acc + a.s16[0]*b.s8[0] + a.s16[1]*b.s8[1] : IDP.2A.LO.S16.S8
__device__ __forceinline__ int dp2a_lo(uint32_t a, int b, int acc)
{
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}
// acc + a.s16[0]*b.s8[2] + a.s16[1]*b.s8[3] : IDP.2A.HI.S16.S8
__device__ __forceinline__ int dp2a_hi(uint32_t a, int b, int acc)
{
    int d;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}

__device__ __forceinline__ int clip16(int x) { return min(max(x, -32768), 32767); }
__device__ __forceinline__ int clip8(int x) { return min(max(x, 0), 255); }

// pack two int16 (low halves of lo, hi) into one word
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410); }
// saturate four ints to u8 and pack (x0 in byte 0)
__device__ __forceinline__ uint32_t pack_sat_u8(int x0, int x1, int x2, int x3)
{
    // cvt.pack d, a, b, c : d = (c << 16) | (sat(a) << 8) | sat(b)
    uint32_t hi, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(x3), "r"(x2));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x1), "r"(x0), "r"(hi));
    return d;
}
// saturate two ints to s16 and pack (x0 in the low half)
__device__ __forceinline__ uint32_t pack_sat_s16(int x0, int x1)
{
    uint32_t d;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(x1), "r"(x0));
    return d;
}

// streaming (read-once) 128-bit global load / store that do not pollute L1
__device__ __forceinline__ int4 ldg_stream(const int4 *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(int4 *p, const int4 &v)
{
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- cooperative staging of a byte tile with arbitrary global alignment --------------------------
//
// Copies `rows` rows of `row_words`*4 bytes starting at g (any byte alignment, row pitch g_pitch bytes) into shared
// memory at s (4-byte aligned, row pitch s_pitch_words words).  Every thread of the CTA participates.  Only aligned
// 32-bit global loads are issued, and only of words that contain at least one requested byte, so nothing outside
// the requested rows is touched beyond the enclosing aligned word.
__device__ __forceinline__ void stage_tile_u8(uint32_t *s, int s_pitch_words, const uint8_t *g, ptrdiff_t g_pitch, int row_words,
                                              int rows, int tid, int nthreads)
{
    const int total = row_words * rows;
    for (int i = tid; i < total; i += nthreads) {
        const int r = i / row_words, j = i - r * row_words;
        const uint8_t *p = g + (ptrdiff_t)r * g_pitch + 4 * j;
        const int a = (int)((uintptr_t)p & 3);
        const uint32_t *pa = (const uint32_t *)(p - a);
        uint32_t lo = __ldg(pa);
        if (a) lo = shr_bytes(lo, __ldg(pa + 1), a);
        s[r * s_pitch_words + j] = lo;
    }
}

// The same staging with a readable byte range [lo, hi): an aligned word is loaded only if it contains at least one byte of the range
// (such a word cannot fault), anything else is staged as zero - those bytes only ever meet positions no tap reads.  Used by the
// interpolation kernels of the *_bounded entry points, which must stay inside the reference's own footprint.
__device__ __forceinline__ void stage_tile_u8_bounded(uint32_t *s, int s_pitch_words, const uint8_t *g, ptrdiff_t g_pitch, int row_words, int rows, int tid,
                                                      int nthreads, const uint8_t *lo, const uint8_t *hi)
{
    const int total = row_words * rows;
    const uintptr_t l = (uintptr_t)lo, h = (uintptr_t)hi;
    for (int i = tid; i < total; i += nthreads) {
        const int r = i / row_words, j = i - r * row_words;
        const uint8_t *p = g + (ptrdiff_t)r * g_pitch + 4 * j;
        const int a = (int)((uintptr_t)p & 3);
        const uintptr_t pa = (uintptr_t)(p - a);
        uint32_t w0 = (pa + 4 > l && pa < h) ? __ldg((const uint32_t *)pa) : 0u;
        if (a) w0 = shr_bytes(w0, (pa + 8 > l && pa + 4 < h) ? __ldg((const uint32_t *)(pa + 4)) : 0u, a);
        s[r * s_pitch_words + j] = w0;
    }
}

}  // namespace hv
