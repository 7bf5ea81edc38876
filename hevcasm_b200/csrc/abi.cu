// hevcasm_b200 - library-wide pieces of the C ABI: launch counter, error strings, device probe.
#include "common.cuh"

#include <atomic>
#include <time.h>

namespace hv {
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace hv

extern "C" unsigned long long hevcasm_cuda_launch_count(void) { return hv::g_launches.load(std::memory_order_relaxed); }

extern "C" const char *hevcasm_cuda_error_string(int code)
{
    if (code == 0) return "success";
    if (code == HEVCASM_ERR_ARGUMENT) return "hevcasm: argument outside the shapes this entry point serves";
    return cudaGetErrorString((cudaError_t)code);
}

extern "C" hevcasm_timestamp hevcasm_get_timestamp_ns(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (hevcasm_timestamp)ts.tv_sec * 1000000000ull + (hevcasm_timestamp)ts.tv_nsec;
}

// reference hevcasm.c:99-138 probes cpuid; the GPU build probes for a Blackwell-class device instead.
extern "C" hevcasm_instruction_set hevcasm_instruction_set_support(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return (hevcasm_instruction_set)0;
    }
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return (hevcasm_instruction_set)0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return (hevcasm_instruction_set)0;
    return major == 10 ? HEVCASM_CUDA : (hevcasm_instruction_set)0;
}

// reference hevcasm.c:141-149
extern "C" void hevcasm_print_instruction_set_support(FILE *f, hevcasm_instruction_set mask)
{
    if (!f) f = stdout;
    fprintf(f, "HEVCasm processor instruction set support:\n");
#define X(value, name, description) fprintf(f, "[%c] " #name " (" description ")\n", ((1 << value) & (int)mask) ? 'x' : ' ');
    HEVCASM_INSTRUCTION_SET_XMACRO
#undef X
    fprintf(f, "\n");
}
