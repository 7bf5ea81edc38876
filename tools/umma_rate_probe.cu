// Rate probe for tcgen05.mma.cta_group::1.kind::i8 (M = 128, K = 32 bytes per instruction): cycles per MMA on a loaded chip
// (one CTA per SM, one issuing thread) as a function of N and of the operand layouts.  Operand contents are irrelevant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I hevcasm_b200/csrc -I include -o tools/umma_rate_probe tools/umma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include "umma.cuh"
using namespace hv;

// layoutA / layoutB: 0 = no swizzle (K-major, 16-byte chunks [chunk][row][16]), 2 = 128-byte swizzle, 6 = 32-byte swizzle
__global__ void __launch_bounds__(128, 1) rate(long long *out, int n, int layoutA, int layoutB, int rounds, int ksteps, int b_mn)
{
    extern __shared__ __align__(128) uint8_t raw[];
    uint8_t *smem = raw + ((1024 - (tma::smem_u32(raw) & 1023)) & 1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u * (i & 3);
    if (tid == 0) tma::mbar_init(&bar, 1);
    if (tid < 32) umma::tmem_alloc<512>(&slot);
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tm = slot;
    uint8_t *sA = smem, *sB = smem + 48 * 1024;
    const uint32_t idesc = umma::idesc_i8(true, false, false, n, b_mn != 0);
    if (tid == 0) {
        long long best = 1ll << 60;
        for (int rep = 0; rep < 4; ++rep) {
            const long long t0 = clock64();
            for (int r = 0; r < rounds; ++r)
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint64_t da = layoutA == 0 ? umma::smem_desc(tma::smem_u32(sA) + ks * 2 * 2048, 2048, 128)
                                                     : umma::smem_desc(tma::smem_u32(sA) + (ks & 3) * 32, 16, layoutA == 2 ? 1024 : 256, layoutA);
                    const uint64_t db = b_mn ? umma::smem_desc(tma::smem_u32(sB) + ks * 4096, 20480, 1024, 2) : layoutB == 0 ? umma::smem_desc(tma::smem_u32(sB) + ks * 2 * (n * 16), n * 16, 128)
                                                     : umma::smem_desc(tma::smem_u32(sB) + (ks & 3) * 32, 16, layoutB == 2 ? 1024 : 256, layoutB);
                    umma::mma_i8(tm + (r & 1) * 256, da, db, idesc, ks);
                }
            umma::commit(&bar);
            tma::mbar_wait(&bar, rep & 1);
            const long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        out[blockIdx.x] = best;
    }
    umma::fence_before();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc<512>(slot);
}

int main()
{
    long long *d, h[148];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int rounds = 64, ksteps = 5;
    const int ns[] = {16, 32, 64, 80, 96, 128, 160, 256};
    const int lay[][2] = {{0, 2}, {0, 0}, {2, 2}, {2, 0}, {6, 6}};
    for (auto &l : lay)
        for (int n : ns) {
            rate<<<148, 128, 100 * 1024>>>(d, n, l[0], l[1], rounds, ksteps, 0);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0, mn = 1ll << 60;
            for (long long v : h) mx = v > mx ? v : mx, mn = v < mn ? v : mn;
            const double per = (double)mn / (rounds * ksteps);
            printf("layoutA %d layoutB %d N %3d: %.1f cycles per MMA (min SM; max SM %.1f) = %.0f MAC/clk/SM\n", l[0], l[1], n, per, (double)mx / (rounds * ksteps),
                   128.0 * n * 32 / per);
        }
    // the image operand of the vertical-pass-first interpolation kernel: MN-major, 128-byte swizzle, two blocks of 128 columns
    for (int n : {128, 192, 216, 256}) {
        rate<<<148, 128, 100 * 1024>>>(d, n, 0, 2, rounds, ksteps, 1);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        long long mn = 1ll << 60;
        for (long long v : h) mn = v < mn ? v : mn;
        printf("A K-major no swizzle, B MN-major 128-byte swizzle, N %3d: %.1f cycles per MMA\n", n, (double)mn / (rounds * ksteps));
    }
    return 0;
}
