// Integer-pipe peak micro-benchmark for the roofline denominators SURVEY.md section 8(d) asks for:
// issue rate (thread-ops / clk / SM) of VABSDIFF4.U8.ACC, IDP.4A, IDP.2A, IMAD, SHF, PRMT, IADD3, LOP3 on this GPU.
// One CTA of 1024 threads per SM, 8 independent dependency chains per thread, cycles from clock64().
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_peak pipe_peak.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

enum Op { VABSDIFF4, IDP4A, IDP2A, IMAD, SHF, PRMT, IADD3, LOP3, MIX_SAD_SHF, MIX_DP4_DP2, MIX_SAD_IMAD, MIX_SAD_PRMT, MIX_IDP_IMAD, MIX_IDP_SHF, MIX_IDP_PRMT, MIX_IMAD_SHF, MIX_IDP_SAD, N_OPS };
static const char *kNames[N_OPS] = {"vabsdiff4_acc", "idp4a", "idp2a", "imad", "shf", "prmt", "iadd3", "lop3", "mix_8sad_1shf", "mix_dp4a_dp2a", "mix_sad_imad", "mix_sad_prmt", "mix_idp_imad", "mix_idp_shf", "mix_idp_prmt", "mix_imad_shf", "mix_idp_sad"};

template <int OP>
__device__ __forceinline__ uint32_t step(uint32_t acc, uint32_t a, uint32_t b)
{
    uint32_t d;
    if (OP == VABSDIFF4) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    else if (OP == IDP4A) asm volatile("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    else if (OP == IDP2A) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    else if (OP == IMAD) asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    else if (OP == SHF) asm volatile("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(acc), "r"(a), "r"(b));
    else if (OP == PRMT) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(acc), "r"(a), "r"(b));
    else if (OP == IADD3) asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(acc), "r"(b + threadIdx.x * (acc & 1)));
    else asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(acc), "r"(a), "r"(b));
    return d;
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) peak_kernel(uint32_t *out, long long *cycles, int iters, uint32_t a, uint32_t b)
{
    uint32_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == MIX_SAD_SHF) {
                    // the SAD inner loop's mix: 8 accumulates per funnel shift
                    acc[i] = step<VABSDIFF4>(acc[i], a, b);
                    if (i == 7) a = step<SHF>(a, b, 8);
                } else if (OP == MIX_DP4_DP2) {
                    acc[i] = (i & 1) ? step<IDP4A>(acc[i], a, b) : step<IDP2A>(acc[i], a, b);
                } else if (OP == MIX_SAD_IMAD) {
                    acc[i] = (i & 1) ? step<VABSDIFF4>(acc[i], a, b) : step<IMAD>(acc[i], a, b);
                } else if (OP == MIX_SAD_PRMT) {
                    acc[i] = (i & 1) ? step<VABSDIFF4>(acc[i], a, b) : step<PRMT>(acc[i], a, b);
                } else if (OP == MIX_IDP_IMAD) {
                    acc[i] = (i & 1) ? step<IDP4A>(acc[i], a, b) : step<IMAD>(acc[i], a, b);
                } else if (OP == MIX_IDP_SHF) {
                    acc[i] = (i & 1) ? step<IDP4A>(acc[i], a, b) : step<SHF>(acc[i], a, b);
                } else if (OP == MIX_IDP_PRMT) {
                    acc[i] = (i & 1) ? step<IDP4A>(acc[i], a, b) : step<PRMT>(acc[i], a, b);
                } else if (OP == MIX_IMAD_SHF) {
                    acc[i] = (i & 1) ? step<IMAD>(acc[i], a, b) : step<SHF>(acc[i], a, b);
                } else if (OP == MIX_IDP_SAD) {
                    acc[i] = (i & 1) ? step<IDP4A>(acc[i], a, b) : step<VABSDIFF4>(acc[i], a, b);
                } else {
                    acc[i] = step<OP>(acc[i], a, b);
                }
            }
    }
    __syncthreads();
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + a;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(int sms, uint32_t *out, long long *cyc, double *ops_clk_sm, double *ms)
{
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    peak_kernel<OP><<<sms, 1024>>>(out, cyc, 64, 0x01020304u, 0x05060708u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    peak_kernel<OP><<<sms, 1024>>>(out, cyc, iters, 0x01020304u, 0x05060708u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float t;
    cudaEventElapsedTime(&t, e0, e1);
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const double ops = 1024.0 * iters * 32.0;  // thread-ops per SM (mix kernels: counted as "accumulate" ops)
    *ops_clk_sm = ops / (double)h[sms / 2];
    *ms = t;
}

int main()
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) {
        printf("{\"error\": \"no device\"}\n");
        return 1;
    }
    const int sms = prop.multiProcessorCount;
    uint32_t *out;
    long long *cyc;
    cudaMalloc(&out, sms * 1024 * sizeof(uint32_t));
    cudaMalloc(&cyc, sms * sizeof(long long));
    double r[N_OPS], ms[N_OPS];
    run<VABSDIFF4>(sms, out, cyc, &r[VABSDIFF4], &ms[VABSDIFF4]);
    run<IDP4A>(sms, out, cyc, &r[IDP4A], &ms[IDP4A]);
    run<IDP2A>(sms, out, cyc, &r[IDP2A], &ms[IDP2A]);
    run<IMAD>(sms, out, cyc, &r[IMAD], &ms[IMAD]);
    run<SHF>(sms, out, cyc, &r[SHF], &ms[SHF]);
    run<PRMT>(sms, out, cyc, &r[PRMT], &ms[PRMT]);
    run<IADD3>(sms, out, cyc, &r[IADD3], &ms[IADD3]);
    run<LOP3>(sms, out, cyc, &r[LOP3], &ms[LOP3]);
    run<MIX_SAD_SHF>(sms, out, cyc, &r[MIX_SAD_SHF], &ms[MIX_SAD_SHF]);
    run<MIX_DP4_DP2>(sms, out, cyc, &r[MIX_DP4_DP2], &ms[MIX_DP4_DP2]);
    run<MIX_SAD_IMAD>(sms, out, cyc, &r[MIX_SAD_IMAD], &ms[MIX_SAD_IMAD]);
    run<MIX_SAD_PRMT>(sms, out, cyc, &r[MIX_SAD_PRMT], &ms[MIX_SAD_PRMT]);
    run<MIX_IDP_IMAD>(sms, out, cyc, &r[MIX_IDP_IMAD], &ms[MIX_IDP_IMAD]);
    run<MIX_IDP_SHF>(sms, out, cyc, &r[MIX_IDP_SHF], &ms[MIX_IDP_SHF]);
    run<MIX_IDP_PRMT>(sms, out, cyc, &r[MIX_IDP_PRMT], &ms[MIX_IDP_PRMT]);
    run<MIX_IMAD_SHF>(sms, out, cyc, &r[MIX_IMAD_SHF], &ms[MIX_IMAD_SHF]);
    run<MIX_IDP_SAD>(sms, out, cyc, &r[MIX_IDP_SAD], &ms[MIX_IDP_SAD]);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"thread_ops_per_clk_per_sm\": {", prop.name, sms, prop.clockRate);
    for (int i = 0; i < N_OPS; ++i) printf("%s\"%s\": %.2f", i ? ", " : "", kNames[i], r[i]);
    printf("}, \"effective_ghz\": {");
    for (int i = 0; i < N_OPS; ++i) printf("%s\"%s\": %.3f", i ? ", " : "", kNames[i], 1024.0 * 4096 * 32 / r[i] / (ms[i] * 1e6));
    printf("}}\n");
    return 0;
}
