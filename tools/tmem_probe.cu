// TMEM read probe: (1) what tcgen05.ld ... .pack::16b returns (which columns land in which register halves, at even and odd
// column addresses), (2) the read bandwidth of tensor memory per SM as a function of the number of reading warps and of the
// load width, plain and packed.  One CTA per SM on the whole chip.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I hevcasm_b200/csrc -I include -o tools/tmem_probe tools/tmem_probe.cu
#include <cstdio>
#include <cstdlib>
#include "umma.cuh"
using namespace hv;

__device__ __forceinline__ void st16(uint32_t taddr, const int (&v)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
                 "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
}
__device__ __forceinline__ void ld8p(uint32_t taddr, int (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void ld32(uint32_t taddr, int (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
          "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
          "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

__global__ void __launch_bounds__(128, 1) semantics(int *out, int odd)
{
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid < 32) umma::tmem_alloc<64>(&slot);
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tl = slot + ((uint32_t)(warp * 32) << 16);
    int v[16];
    for (int h = 0; h < 2; ++h) {
        for (int i = 0; i < 16; ++i) v[i] = (tid << 8) + 16 * h + i - 40000 * ((16 * h + i) & 1);   // odd columns negative: the low halves must survive
        st16(tl + 16 * h, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    int a[8], b[8];
    ld8p(tl, a);
    umma::tmem_ld_wait();
    if (odd) ld8p(tl + 1, b); else ld8p(tl + 2, b);
    umma::tmem_ld_wait();
    if (tid == 37)
        for (int i = 0; i < 8; ++i) out[i] = a[i], out[8 + i] = b[i];
    umma::fence_before();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc<64>(slot);
}

// mode 0: x8 plain, 1: x16 plain, 2: x32 plain, 3: x8 packed (16 columns per load)
__global__ void __launch_bounds__(512, 1) bandwidth(long long *out, int mode, int rounds)
{
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid < 32) umma::tmem_alloc<512>(&slot);
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tl = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    int acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        if (mode == 0) {
            int a[8], b[8], c[8], d[8];
            umma::tmem_ld8(tl, a), umma::tmem_ld8(tl + 8, b), umma::tmem_ld8(tl + 16, c), umma::tmem_ld8(tl + 24, d);
            umma::tmem_ld_wait();
            acc += a[0] ^ a[7] ^ b[1] ^ c[2] ^ d[3];
        } else if (mode == 1) {
            int a[16], b[16];
            umma::tmem_ld16(tl, a), umma::tmem_ld16(tl + 16, b);
            umma::tmem_ld_wait();
            acc += a[0] ^ a[15] ^ b[1];
        } else if (mode == 2) {
            int a[32];
            ld32(tl, a);
            umma::tmem_ld_wait();
            acc += a[0] ^ a[31];
        } else {
            int a[8], b[8];
            ld8p(tl, a), ld8p(tl + 16, b);
            umma::tmem_ld_wait();
            acc += a[0] ^ a[7] ^ b[1];
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x7fffffff) out[0] = acc;
    umma::fence_before();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc<512>(slot);
}

int main()
{
    int *ds, hs[16];
    cudaMalloc(&ds, sizeof(hs));
    auto sem = [&](int odd) {
    semantics<<<1, 128>>>(ds, odd);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("semantics(odd=%d) failed: %s\n", odd, cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(hs, ds, sizeof(hs), cudaMemcpyDeviceToHost);
    printf("stored column c of lane 37: (37 << 8) + c - 40000 * (c & 1)\n");
    for (int k = 0; k < 2; ++k)
        for (int i = 0; i < 8; ++i) {
            const int lo = (short)(hs[8 * k + i] & 0xffff), hi = hs[8 * k + i] >> 16;
            printf("packed load at column %d, register %d: lo %d (column %d)  hi %d (column %d)\n", k ? (odd ? 1 : 2) : 0, i, lo, lo < 0 ? lo + 40000 - (37 << 8) : lo - (37 << 8), hi,
                   hi < 0 ? hi + 40000 - (37 << 8) : hi - (37 << 8));
        }
    return 0;
    };
    if (sem(0)) return 1;
    long long *d, h[148];
    cudaMalloc(&d, sizeof(h));
    const int rounds = 4096;
    const char *names[] = {"4 x x8 plain", "2 x x16 plain", "1 x x32 plain", "2 x x8 packed (32 columns)"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps = 4; warps <= 16; warps *= 2) {
            bandwidth<<<148, warps * 32>>>(d, mode, rounds);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("bandwidth failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            long long mn = 1ll << 60;
            for (long long v : h) mn = v < mn ? v : mn;
            const double cols = 32.0 * rounds * warps;   // 32 columns of 32 lanes per warp and round
            printf("%-28s %2d warps: %.1f cycles per round, %.1f TMEM bytes/clk/SM (register bytes %.1f)\n", names[mode], warps, (double)mn / rounds,
                   cols * 32 * 4 / mn, cols * 32 * 4 / mn * (mode == 3 ? 0.5 : 1.0));
        }
    return 0;
}
