// Probe for the tensor-core horizontal interpolation pass:
//   D[m = output column (128)][n = row (80)] = sum_k A[m][k] * B[k][n]
//   A = banded Toeplitz matrix of the 8 filter taps (s8, K-major, built in shared memory),
//   B = image bytes (u8), rows of 160 bytes starting 16 bytes left of the tile, fetched by ONE 4-D TMA box whose dimension
//       order (16 bytes, rows, 16-byte chunks) lands them in shared memory as the no-swizzle K-major core-matrix layout.
// Checks (a) that cuTensorMapEncodeTiled accepts the (16 B, row stride, 16 B) stride order, (b) the K-major descriptors,
// (c) five accumulating K-steps with N = 80.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_fir_probe tools/umma_fir_probe.cu && ./umma_fir_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0 /* 0 none, 2 = 128B swizzle, 6 = 32B swizzle */)
{
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    if (!done) __trap();
}

constexpr int ROWS = 80, KB = 160, CHUNKS = KB / 16;

__global__ void probe(const __grid_constant__ CUtensorMap tm, const int8_t *coef /*8*/, int32_t *d_out /*[128][80]*/, int x_chunk0, int y0)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *smem = smem_raw;
    uint8_t *sA = smem;                    // [chunk (10)][m (128)][16]  = 20480 bytes
    uint8_t *sB = smem + 128 * KB;         // [chunk (10)][row (80)][16] = 12800 bytes
    __shared__ __align__(8) uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // A[m][k] = coef[k - m - 13] (output column m reads input bytes m + 13 .. m + 20 of the 160-byte row: 16 left pad - 3)
    for (int i = tid; i < 128 * KB; i += 128) {
        const int m = i / KB, k = i % KB, t = k - m - 13;
        sA[(k / 16) * 2048 + m * 16 + (k % 16)] = (t >= 0 && t < 8) ? (uint8_t)coef[t] : 0;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_tma)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm_addr = tmem_base;
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_tma)), "r"(ROWS * KB) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(sB)),
                     "l"(&tm), "r"(0), "r"(y0), "r"(x_chunk0), "r"(0), "r"(smem_u32(&bar_tma))
                     : "memory");
    }
    mbar_wait(&bar_tma, 0);
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t idesc = 0;
        idesc |= 2u << 4;             // D s32
        idesc |= 1u << 7;             // A s8
        idesc |= 0u << 10;            // B u8
        idesc |= (uint32_t)(ROWS >> 3) << 17;
        idesc |= (128u >> 4) << 24;   // both operands K-major
        for (int ks = 0; ks < KB / 32; ++ks) {
            // K-major, no swizzle: LBO = distance between 16-byte k chunks, SBO = distance between groups of 8 rows
            const uint64_t da = make_desc(smem_u32(sA + ks * 2 * 2048), 2048, 128);
            const uint64_t db = make_desc(smem_u32(sB + ks * 2 * ROWS * 16), ROWS * 16, 128);
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm_addr),
                "l"(da), "l"(db), "r"(idesc), "r"(ks)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    mbar_wait(&bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tm_addr + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < ROWS; c0 += 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                       "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) d_out[(warp * 32 + lane) * ROWS + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tm_addr) : "memory");
}

// variant 2: the image operand as two plain 3-D boxes - {128 bytes x 80 rows} with the 128-byte swizzle (K-steps 0..3) and
// {32 bytes x 80 rows} with the 32-byte swizzle (K-step 4): 160 row pieces per tile instead of 800 16-byte pieces
__global__ void probe_sw(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm32, const int8_t *coef, int32_t *d_out, int xb, int y0)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);   // the 128-byte swizzle atom must sit on a 1024-byte boundary
    uint8_t *sB1 = smem;                       // 80 rows x 128 B, 128B-swizzled (10240 bytes, 1024-aligned)
    uint8_t *sB2 = smem + ROWS * 128;          // 80 rows x 32 B, 32B-swizzled (2560 bytes, 256-aligned)
    uint8_t *sA = smem + ROWS * 160;           // Toeplitz, no swizzle, [chunk][m][16]
    __shared__ __align__(8) uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * KB; i += 128) {
        const int m = i / KB, k = i % KB, t = k - m - 13;
        sA[(k / 16) * 2048 + m * 16 + (k % 16)] = (t >= 0 && t < 8) ? (uint8_t)coef[t] : 0;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_tma)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm_addr = tmem_base;
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_tma)), "r"(ROWS * KB) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(sB1)),
                     "l"(&tm128), "r"(xb), "r"(y0), "r"(0), "r"(smem_u32(&bar_tma))
                     : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(sB2)),
                     "l"(&tm32), "r"(xb + 128), "r"(y0), "r"(0), "r"(smem_u32(&bar_tma))
                     : "memory");
    }
    mbar_wait(&bar_tma, 0);
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(ROWS >> 3) << 17) | ((128u >> 4) << 24);
        for (int ks = 0; ks < KB / 32; ++ks) {
            const uint64_t da = make_desc(smem_u32(sA + ks * 2 * 2048), 2048, 128);
            // swizzled K-major: 8-row groups are 1024 (256) bytes apart; a K-step advances the start address by 32 bytes inside the swizzle row
            const uint64_t db = ks < 4 ? make_desc(smem_u32(sB1 + ks * 32), 16, 1024, 2) : make_desc(smem_u32(sB2), 16, 256, 6);
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm_addr),
                "l"(da), "l"(db), "r"(idesc), "r"(ks)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    mbar_wait(&bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tm_addr + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < ROWS; c0 += 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                       "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) d_out[(warp * 32 + lane) * ROWS + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tm_addr) : "memory");
}

int main()
{
    const int W = 1024, H = 256, STRIDE = 1024;
    static uint8_t img[H * STRIDE];
    uint64_t s = 0x48455643;
    auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); };
    for (auto &x : img) x = (uint8_t)rnd();
    const int8_t coef[8] = {-1, 4, -10, 58, 17, -5, 1, 0};
    uint8_t *dimg; int8_t *dcoef; int32_t *dd;
    cudaMalloc(&dimg, sizeof img); cudaMalloc(&dcoef, 8); cudaMalloc(&dd, 128 * ROWS * 4);
    cudaMemcpy(dimg, img, sizeof img, cudaMemcpyHostToDevice);
    cudaMemcpy(dcoef, coef, 8, cudaMemcpyHostToDevice);

    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encoder\n"); return 2; }
    CUtensorMap tm;
    // dims: (16 bytes, rows, 16-byte chunks of a row, frames); strides of dims 1..3 in bytes
    cuuint64_t dim[4] = {16, (cuuint64_t)H, (cuuint64_t)(W / 16), 1};
    cuuint64_t stride[3] = {(cuuint64_t)STRIDE, 16, (cuuint64_t)H * STRIDE};
    cuuint32_t box[4] = {16, ROWS, CHUNKS, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, dimg, dim, stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode (16 B, row stride, 16 B chunk stride): %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 3;

    const int x0 = 256, y0 = 40;   // tile origin: output columns x0 .. x0+127, rows y0 .. y0+79; the box starts 16 bytes left
    const size_t smem = 128 * KB + ROWS * KB;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(tm, dcoef, dd, (x0 - 16) / 16, y0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    static int32_t hd[128 * ROWS];
    cudaMemcpy(hd, dd, sizeof hd, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < ROWS; ++n) {
            int acc = 0;
            for (int t = 0; t < 8; ++t) acc += coef[t] * (int)img[(y0 + n) * STRIDE + x0 + m - 3 + t];
            if (acc != hd[m * ROWS + n] && bad++ < 5) printf("  m %d n %d: want %d got %d\n", m, n, acc, hd[m * ROWS + n]);
        }
    printf("horizontal 8-tap pass on the tensor core: %d / %d mismatches\n", bad, 128 * ROWS);

    // variant 2: swizzled 3-D boxes
    CUtensorMap t128, t32;
    cuuint64_t dim3[3] = {(cuuint64_t)W, (cuuint64_t)H, 1};
    cuuint64_t str3[2] = {(cuuint64_t)STRIDE, (cuuint64_t)H * STRIDE};
    cuuint32_t e3[3] = {1, 1, 1};
    cuuint32_t b128[3] = {128, ROWS, 1}, b32[3] = {32, ROWS, 1};
    CUresult r1 = ((EncodeTiledFn)fn)(&t128, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, dimg, dim3, str3, b128, e3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = ((EncodeTiledFn)fn)(&t32, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, dimg, dim3, str3, b32, e3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode swizzled maps: %d %d\n", (int)r1, (int)r2);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return 3;
    const size_t smem2 = ROWS * 160 + 128 * KB + 1024;
    cudaFuncSetAttribute(probe_sw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    cudaMemset(dd, 0xff, 128 * ROWS * 4);
    probe_sw<<<1, 128, smem2>>>(t128, t32, dcoef, dd, x0 - 16, y0);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    cudaMemcpy(hd, dd, sizeof hd, cudaMemcpyDeviceToHost);
    int bad2 = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < ROWS; ++n) {
            int acc = 0;
            for (int t = 0; t < 8; ++t) acc += coef[t] * (int)img[(y0 + n) * STRIDE + x0 + m - 3 + t];
            if (acc != hd[m * ROWS + n] && bad2++ < 5) printf("  m %d n %d: want %d got %d\n", m, n, acc, hd[m * ROWS + n]);
        }
    printf("swizzled two-box image operand: %d / %d mismatches\n", bad2, 128 * ROWS);
    return (bad || bad2) ? 1 : 0;
}
