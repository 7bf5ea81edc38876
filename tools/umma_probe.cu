// tcgen05 kind::i8 probe: D[128 x 32] (s32, TMEM) = A[128 x 32] (u8 or s8, MN-major in smem) * B[32 x 32] (s8, K-major in smem).
// Checks the shared-memory matrix-descriptor fields (LBO / SBO roles for the no-swizzle canonical layouts) and the
// instruction descriptor against a host product before the transform kernels depend on them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe tools/umma_probe.cu && ./umma_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base offset 0, layout type 0 (no swizzle)
}

__global__ void probe(const uint8_t *a_in /*[128][32] (m,k)*/, const int8_t *b_in /*[32][32] (k,n)*/, int32_t *d_out /*[128][32]*/, int a_signed, int swap)
{
    __shared__ __align__(128) uint8_t sa[128 * 32];
    __shared__ __align__(128) uint8_t sb[32 * 32];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // A, MN-major canonical (no swizzle): 16 consecutive m at one k are 16 contiguous bytes; 8 k-rows make a 128-byte core matrix
    for (int i = tid; i < 128 * 32; i += 128) {
        const int m = i / 32, k = i % 32;
        sa[(m / 16) * 512 + (k / 8) * 128 + (k % 8) * 16 + (m % 16)] = a_in[i];
    }
    // B (N x K), K-major canonical: 16 consecutive k at one n are contiguous; 8 n-rows make a core matrix
    for (int i = tid; i < 32 * 32; i += 128) {
        const int k = i / 32, n = i % 32;
        sb[(n / 8) * 256 + (k / 16) * 128 + (n % 8) * 16 + (k % 16)] = (uint8_t)b_in[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint64_t da = swap ? make_desc(smem_u32(sa), 512, 128) : make_desc(smem_u32(sa), 128, 512);
        const uint64_t db = make_desc(smem_u32(sb), 128, 256);
        uint32_t idesc = 0;
        idesc |= 2u << 4;                       // D: s32
        idesc |= (a_signed ? 1u : 0u) << 7;     // A: u8 / s8
        idesc |= 1u << 10;                      // B: s8
        idesc |= 1u << 15;                      // A MN-major
        idesc |= 0u << 16;                      // B K-major
        idesc |= (32u >> 3) << 17;              // N
        idesc |= (128u >> 4) << 24;             // M
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(0)
            : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait for the MMA (bounded)
    {
        uint32_t done = 0;
        for (int spin = 0; spin < (1 << 22) && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done)
                         : "r"(smem_u32(&bar)), "r"(0)
                         : "memory");
        if (!done) __trap();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
          "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
          "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) d_out[(warp * 32 + lane) * 32 + j] = (int32_t)v[j];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm) : "memory");
}

int main()
{
    uint8_t ha[128 * 32];
    int8_t hb[32 * 32];
    uint64_t s = 0x48455643;
    auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); };
    for (auto &x : ha) x = (uint8_t)rnd();
    for (auto &x : hb) x = (int8_t)((int)(rnd() % 181) - 90);
    uint8_t *da; int8_t *db; int32_t *dd;
    cudaMalloc(&da, sizeof ha); cudaMalloc(&db, sizeof hb); cudaMalloc(&dd, 128 * 32 * 4);
    cudaMemcpy(da, ha, sizeof ha, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb, sizeof hb, cudaMemcpyHostToDevice);
    int rc = 1;
    for (int a_signed = 0; a_signed < 2; ++a_signed)
        for (int swap = 0; swap < 2; ++swap) {
            cudaMemset(dd, 0xff, 128 * 32 * 4);
            probe<<<1, 128>>>(da, db, dd, a_signed, swap);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("a_signed %d swap %d: CUDA error %s\n", a_signed, swap, cudaGetErrorString(e)); return 2; }
            static int32_t hd[128 * 32];
            cudaMemcpy(hd, dd, sizeof hd, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 32; ++n) {
                    int acc = 0;
                    for (int k = 0; k < 32; ++k) acc += (a_signed ? (int)(int8_t)ha[m * 32 + k] : (int)ha[m * 32 + k]) * (int)hb[k * 32 + n];
                    bad += acc != hd[m * 32 + n];
                }
            printf("a_signed %d  A desc (LBO,SBO) = %s : %d / 4096 mismatches  (d[0][0..3] = %d %d %d %d)\n", a_signed, swap ? "(512,128)" : "(128,512)", bad, hd[0], hd[1],
                   hd[2], hd[3]);
            if (!bad) rc = 0;
        }
    return rc;
}
