// hevcasm_b200 - thin wrappers over the 5th-generation tensor-core instructions (tcgen05, kind::i8) and tensor memory:
// shared-memory matrix descriptors, the instruction descriptor, MMA issue / commit, TMEM allocation and loads.
// Field meanings were pinned on the hardware with tools/umma_probe.cu and tools/umma_fir_probe.cu.
#pragma once

#include "tma.cuh"

namespace hv {
namespace umma {

// shared-memory matrix descriptor, no swizzle (LBO / SBO in bytes, multiples of 16)
//   MN-major operand: LBO = distance between groups of 8 k, SBO = distance between chunks of 16 m
//   K-major operand : LBO = distance between chunks of 16 k, SBO = distance between groups of 8 n
//   128-byte-swizzled MN-major operand (rows of 128 bytes along MN as TMA writes them with the 128-byte swizzle, one row per k):
//   SBO = distance between groups of 8 k (1024), LBO = distance between blocks of 128 mn; a K-step of 32 advances `addr` by 4096
//   (tools/umma_vfirst_probe.cu)
//   swizzled K-major operand (layout 2 = 128-byte swizzle, 6 = 32-byte swizzle; atoms of 8 rows x 128 / 32 bytes written by TMA with
//   the matching swizzle mode): SBO = distance between groups of 8 rows (1024 / 256), LBO unused; a K-step advances `addr` by 32 bytes
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0)
{
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)layout << 61);
}
// instruction descriptor: D s32, A / B u8 or s8, A MN-major or K-major, B K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_i8(bool a_signed, bool b_signed, bool a_mn_major, int n, bool b_mn_major = false)
{
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
// D = A * B (accumulate = 0) or D += A * B
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, int accumulate = 0)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all MMAs issued so far by this thread -> one arrival on `bar` when they have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *slot)   // one whole warp
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tma::smem_u32(slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr)  // the same warp
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
// 16 consecutive columns of this thread's TMEM lane (warp w reads lanes 32w .. 32w+31)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
          "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
// 8 consecutive columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
// 8 consecutive columns (the address must be EVEN) as 4 words: the low 16 bits of columns 2i and 2i+1 in the low / high half of word i -
// int16 pairs straight out of TMEM, no PRMT (semantics and alignment pinned by tools/tmem_probe.cu: an odd column address faults)
__device__ __forceinline__ void tmem_ld8_pack16(uint32_t taddr, uint32_t (&v)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.pack::16b.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the wait tied to the registers of a packed + a plain load of the same 8 columns
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&p)[4], int (&v)[8])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(p[0]), "+r"(p[1]), "+r"(p[2]), "+r"(p[3]), "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])::"memory");
}
// the same wait, tied to the registers of an earlier load so that no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(int (&v)[8])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])::"memory");
}

}  // namespace umma

}  // namespace hv
