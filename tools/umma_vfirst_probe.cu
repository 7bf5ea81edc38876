// Probe for the VERTICAL interpolation pass on the tensor cores with the image as the MN-major B operand:
//   D[m = output row (128)][n = column (256)] = sum_k A[m][k] * B[k][n]
//   A = banded Toeplitz matrix of the 8 vertical taps (s8, K-major, no swizzle, [chunk][m][16]),
//   B = image bytes (u8): two TMA boxes {128 bytes x 160 rows} with the 128-byte swizzle, i.e. rows of the image as they lie in
//       memory = the MN-major operand form (N contiguous).  Which (LBO, SBO) describe it is what this probe finds out.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I hevcasm_b200/csrc -I include -o tools/umma_vfirst_probe tools/umma_vfirst_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "umma.cuh"
using namespace hv;

constexpr int KR = 160, NB = 256, BOX = KR * 128;

__device__ bool soft_wait(uint64_t *bar, uint32_t parity)
{
    for (int spin = 0; spin < (1 << 18); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(tma::smem_u32(bar)), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmap, const int8_t *coef, int32_t *d_out, int xb, int y0, uint32_t lbo, uint32_t sbo,
                                                uint32_t kadv, int n, uint32_t bmajor, uint32_t layout)
{
    extern __shared__ __align__(128) uint8_t raw[];
    uint8_t *smem = raw + ((1024 - (tma::smem_u32(raw) & 1023)) & 1023);
    uint8_t *sB = smem, *sA = smem + 2 * BOX;
    __shared__ __align__(8) uint64_t bar_tma, bar_mma;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * KR; i += 128) {
        const int m = i / KR, k = i % KR, t = k - m;
        sA[(k / 16) * 2048 + m * 16 + (k % 16)] = (t >= 0 && t < 8) ? (uint8_t)coef[t] : 0;
    }
    if (tid == 0) tma::mbar_init(&bar_tma, 1), tma::mbar_init(&bar_mma, 1);
    if (tid < 32) umma::tmem_alloc<256>(&slot);
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tm = slot;
    if (tid == 0) {
        tma::mbar_expect_tx(&bar_tma, 2 * BOX);
        tma::load_box_3d(sB, &tmap, xb, y0, 0, &bar_tma);
        tma::load_box_3d(sB + BOX, &tmap, xb + 128, y0, 0, &bar_tma);
    }
    const bool ok1 = soft_wait(&bar_tma, 0);
    if (!ok1 && tid == 0) printf("TMA wait timed out\n");
    if (tid == 0) {
        umma::fence_after();
        // D s32, A s8 K-major, B u8 MN-major (bit 16), M = 128
        const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | (0u << 15) | (bmajor << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
        for (int ks = 0; ks < KR / 32; ++ks) {
            const uint64_t da = umma::smem_desc(tma::smem_u32(sA + ks * 2 * 2048), 2048, 128);
            const uint64_t db = umma::smem_desc(tma::smem_u32(sB) + ks * kadv, lbo, sbo, layout);
            umma::mma_i8(tm, da, db, idesc, ks);
        }
        umma::commit(&bar_mma);
    }
    const bool ok2 = soft_wait(&bar_mma, 0);
    if (!ok2 && tid == 0) printf("MMA wait timed out\n");
    umma::fence_after();
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < n; c0 += 8) {
        int v[8];
        umma::tmem_ld8(taddr + c0, v);
        umma::tmem_ld_wait(v);
        for (int j = 0; j < 8; ++j) d_out[tid * NB + c0 + j] = v[j];
    }
    // unaligned column starts: columns 5 .. n-4 read again in groups of 8 starting at 5, 13, ..
    int bad = 0;
    for (int c0 = 5; c0 + 8 <= n; c0 += 8) {
        int v[8];
        umma::tmem_ld8(taddr + c0, v);
        umma::tmem_ld_wait(v);
        for (int j = 0; j < 8; ++j) bad += d_out[tid * NB + c0 + j] != v[j];
    }
    if (bad) printf("thread %d: %d mismatches with unaligned TMEM column starts\n", tid, bad);
    umma::fence_before();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc<256>(slot);
}

int main(int argc, char **argv)
{
    const int W = 1024, H = 512, STRIDE = 1024;
    std::vector<uint8_t> img((size_t)H * STRIDE);
    uint64_t s = 0x48455643;
    auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); };
    for (auto &x : img) x = (uint8_t)rnd();
    const int8_t coef[8] = {-1, 4, -10, 58, 17, -5, 1, 0};
    uint8_t *dimg; int8_t *dcoef; int32_t *dd;
    cudaMalloc(&dimg, img.size()); cudaMalloc(&dcoef, 8); cudaMalloc(&dd, 128 * NB * 4);
    cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dcoef, coef, 8, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    if (tma::describe_u8_swizzled(&tm, dimg, STRIDE, 0, W, H, 1, 128, KR)) { printf("encode failed\n"); return 2; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * BOX + 128 * KR + 2048);
    const int xb = argc > 7 ? atoi(argv[7]) : 37, y0 = 11;   // unaligned box start on purpose
    std::vector<int32_t> out(128 * NB);
    if (argc < 5) { printf("usage: probe n kadv lbo sbo\n"); return 2; }
    const int n = atoi(argv[1]);
    const uint32_t kadv = atoi(argv[2]), lbo = atoi(argv[3]), sbo = atoi(argv[4]);
    cudaMemset(dd, 0xff, 128 * NB * 4);
    probe<<<1, 128, 2 * BOX + 128 * KR + 2048>>>(tm, dcoef, dd, xb, y0, lbo, sbo, kadv, n, argc > 5 ? atoi(argv[5]) : 1, argc > 6 ? atoi(argv[6]) : 2);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("n %d kadv %u lbo %u sbo %u: launch failed %s\n", n, kadv, lbo, sbo, cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, bad_lo = 0;
    for (int m = 0; m < 128; ++m)
        for (int x = 0; x < n; ++x) {
            int ref = 0;
            for (int t = 0; t < 8; ++t) ref += coef[t] * img[(size_t)(y0 + m + t) * STRIDE + xb + x];
            if (out[m * NB + x] != ref) ++bad, bad_lo += x < 128;
        }
    printf("n %3d kadv %4u lbo %5u sbo %5u: %d mismatches (%d in columns < 128)\n", n, kadv, lbo, sbo, bad, bad_lo);
    return 0;
}
