// Development probe for hevcasm_b200/csrc/tma.cuh: loads one box through TMA and checks it byte for byte on the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I hevcasm_b200/csrc -I include -o tools/tma_probe tools/tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tma.cuh"
namespace hv { void count_launch() {} }
using namespace hv;

struct P { CUtensorMap tm; int x, y, z, bytes; uint8_t *out; int *status; };

__global__ void probe(const __grid_constant__ P p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 32768);
    if (threadIdx.x == 0) tma::mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        tma::mbar_expect_tx(bar, p.bytes);
        tma::load_box_3d(smem, &p.tm, p.x, p.y, p.z, bar);
    }
    // bounded wait without trap
    const uint32_t addr = tma::smem_u32(bar);
    int ok = 0;
    for (int spin = 0; spin < (1 << 20) && !ok; ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(0) : "memory");
        ok = done;
    }
    if (threadIdx.x == 0) *p.status = ok ? 1 : -1;
    if (ok) for (int i = threadIdx.x; i < p.bytes; i += blockDim.x) p.out[i] = smem[i];
}

int main(int argc, char **argv)
{
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    const int pitch = 512, rows = 200, frames = 2;
    std::vector<uint8_t> h((size_t)pitch * rows * frames);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *out; int *status;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&out, 65536); cudaMalloc(&status, 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 16);
    struct Case { int first_off, ext_x, ext_y, box_x, box_y, x, y, z; } cases[] = {
        {0, 256, 128, 128, 64, 0, 0, 0},        // aligned 128x64 box
        {0, 256, 128, 128, 64, 128, 64, 1},
        {12 * 512 + 12, 263, 135, 144, 71, 0, 0, 0},   // window-like: origin (12,12) -> shift 12
        {12 * 512 + 12, 263, 135, 144, 71, 128, 64, 1},  // partially out of range
        {12 * 512 + 13, 263, 135, 144, 71, 3, 5, 0},   // odd alignment, arbitrary coordinates
        {0, 272, 135, 144, 71, 0, 0, 0},        // 5: 144x71 box, aligned, dim0 multiple of 16
        {0, 263, 135, 144, 71, 0, 0, 0},        // 6: dim0 = 263
        {0, 256, 128, 128, 64, 12, 0, 0},       // 7: 128x64 box at x = 12
        {0, 256, 128, 128, 64, 13, 3, 0},       // 8: 128x64 box at x = 13
        {0, 256, 128, 128, 71, 0, 0, 0},        // 9: 71 rows
        {0, 256, 128, 144, 64, 0, 0, 0},        // 10: 144 wide
        {0, 256, 128, 256, 64, 0, 0, 0},        // 11: 256 wide
        {12, 263, 135, 128, 64, 0, 0, 0},       // 12: first at +12 (shift 12), x = 12
    };
    int fails = 0;
    int ci = -1;
    for (auto &c : cases) {
        ++ci;
        if (only >= 0 && ci != only) continue;
        P p; int shift;
        int e = tma::describe_u8(&p.tm, d + c.first_off, pitch, (ptrdiff_t)pitch * rows, c.ext_x, c.ext_y, frames, c.box_x, c.box_y, &shift);
        p.x = c.x + shift, p.y = c.y, p.z = c.z, p.bytes = c.box_x * c.box_y, p.out = out, p.status = status;
        cudaMemset(out, 0xEE, 65536); cudaMemset(status, 0, 4);
        probe<<<1, 128, 32768 + 16>>>(p);
        cudaError_t err = cudaDeviceSynchronize();
        int st = 0; cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost);
        std::vector<uint8_t> got(p.bytes); cudaMemcpy(got.data(), out, p.bytes, cudaMemcpyDeviceToHost);
        long bad = 0, zero_fill = 0;
        for (int r = 0; r < c.box_y; ++r) for (int b = 0; b < c.box_x; ++b) {
            const int gx = c.x + b, gy = c.y + r;
            const bool inside = gx < c.ext_x && gy < c.ext_y;   // relative to `first`
            const uint8_t want = inside ? h[(size_t)c.z * pitch * rows + c.first_off + (size_t)gy * pitch + gx] : 0;
            if (!inside) ++zero_fill;
            if (got[r * c.box_x + b] != want) ++bad;
        }
        printf("case %d box %dx%d at (%d,%d,%d) shift %d: encode=%d launch=%s status=%d mismatches=%ld (zero-filled expected %ld)\n", ci, c.box_x, c.box_y, c.x, c.y, c.z, shift, e,
               cudaGetErrorString(err), st, bad, zero_fill);
        fails += (e || err != cudaSuccess || st != 1 || bad);
        if (err != cudaSuccess) break;
    }
    return fails;
}
