// How many SMALL TMA boxes per cycle can one SM pull?  (Design input for the prediction-unit list kernels: one box per PU footprint -
// 15 rows of 15 bytes for an 8x8 luma PU - would land row-aligned in shared memory; but a box must START 16-byte aligned, so the probe
// rounds its x coordinates down to 16: an unaligned start faults, as tools/tma_probe.cu cases 7, 8 and 12 show.)
// One CTA per SM, W issuing warps, each keeping two batches of D boxes in flight on two mbarriers; box coordinates walk 8x8 PUs in raster
// order over 16 4K frames with a pseudo-random +-16 sample displacement, like the bench lists.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I hevcasm_b200/csrc -I include -o tools/tma_box_probe tools/tma_box_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "tma.cuh"
namespace hv { void count_launch() {} }
using namespace hv;

struct P { CUtensorMap tm; int box_bytes, slot_bytes, depth, iters, per_lane, step; long long *cycles; unsigned *sink; };

__device__ __forceinline__ unsigned mix(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__device__ __forceinline__ bool wait_bounded(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = tma::smem_u32(bar);
    for (int spin = 0; spin < (1 << 16); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}

__global__ void __launch_bounds__(1024, 1) probe(const __grid_constant__ P p)
{
    extern __shared__ __align__(128) uint8_t smem[];   // everything in dynamic shared memory: ring slots, then the barriers
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint8_t *mine = smem + (size_t)warp * 2 * p.depth * p.slot_bytes;
    uint64_t (*dynbars)[2] = reinterpret_cast<uint64_t (*)[2]>(smem + (size_t)nw * 2 * p.depth * p.slot_bytes);
    uint64_t *const mybar = dynbars[warp];
    if (lane == 0) tma::mbar_init(mybar, 1), tma::mbar_init(mybar + 1, 1);
    __syncthreads();
    const long long t0 = clock64();
    unsigned acc = 0;
    for (int it = 0; it < p.iters + 2; ++it) {
        const int half = it & 1;
        if (it >= 2) {
            if (!wait_bounded(mybar + half, ((it - 2) >> 1) & 1)) { if (lane == 0) atomicMax(p.sink + 1, (unsigned)it + 1); break; }
            acc += mine[(size_t)(half * p.depth + (lane % p.depth)) * p.slot_bytes + lane];
        }
        if (it < p.iters) {
            __syncwarp();
            if (lane == 0) tma::mbar_expect_tx(mybar + half, (uint32_t)(p.depth * p.box_bytes));
            __syncwarp();
            for (int k = p.per_lane ? lane : 0; k < p.depth; k += p.per_lane ? 32 : 1) {
                if (!p.per_lane && lane) break;
                const unsigned pu = ((blockIdx.x * nw + warp) * p.iters + it) * p.depth + k;
                const unsigned h = mix(pu * 2654435761u + 17);
                const int frame = (pu / (480 * 270)) & 15, r = pu % (480 * 270);
                const int x = (64 + (r % 480) * p.step + (int)(h % 33) - 16) & ~15,   // box starts must be 16-byte aligned (tools/tma_probe.cu cases 7, 8, 12 fault)
                          y = 64 + (r / 480) * p.step + (int)((h >> 8) % 33) - 16;
                tma::load_box_3d(mine + (size_t)(half * p.depth + k) * p.slot_bytes, &p.tm, x, y, frame, mybar + half);
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) p.cycles[blockIdx.x] = t1 - t0;
    if (acc == 0xffffffffu) p.sink[0] = acc;
}

int main(int argc, char **argv)
{
    const int H = 2160, PAD = 64, NF = 16, pitch = 4096, rows = H + 2 * PAD;
    uint8_t *img;
    cudaMalloc(&img, (size_t)pitch * rows * NF);
    cudaMemset(img, 7, (size_t)pitch * rows * NF);
    long long *cyc; unsigned *sink;
    cudaMalloc(&cyc, 148 * sizeof(long long));
    cudaMalloc(&sink, 8);
    cudaMemset(sink, 0, 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Box { int bx, by, step; } boxes[] = {{32, 16, 8}, {32, 8, 8}, {48, 24, 16}, {64, 40, 32}, {96, 72, 64}, {128, 16, 8}, {16, 16, 8}};
    printf("box_x box_y  warps depth issue     cycles/box/SM   boxes/us/SM   GB/s(chip, box bytes)\n");
    const int only = argc > 3 ? atoi(argv[3]) : -1;   // argv: [max depth] [max iterations] [box index]
    for (const Box &b : boxes) {
        if (only >= 0 && &b != &boxes[only]) continue;
        P p;
        int shift;
        if (tma::describe_u8(&p.tm, img, pitch, (ptrdiff_t)pitch * rows, pitch, rows, NF, b.bx, b.by, &shift)) { printf("describe failed\n"); return 1; }
        p.box_bytes = b.bx * b.by; p.slot_bytes = (p.box_bytes + 127) & ~127; p.step = b.step;
        p.cycles = cyc; p.sink = sink;
        const int max_depth = argc > 1 ? atoi(argv[1]) : 32, max_iters = argc > 2 ? atoi(argv[2]) : 1 << 30;
        for (int nw : {1, 4, 16})
            for (int per_lane : {0, 1}) {
                int depth = (int)(180 * 1024 / (2 * nw * p.slot_bytes));
                if (depth > max_depth) depth = max_depth;
                if (depth < 1) continue;
                p.depth = depth; p.per_lane = per_lane;
                p.iters = std::min(max_iters, 20000 / (nw * depth) + 4);
                const size_t smem = (size_t)2 * nw * depth * p.slot_bytes + 16 * nw + 1024;
                probe<<<148, 32 * nw, smem>>>(p);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("launch failed: %s (box %dx%d, %d warps, depth %d, per_lane %d, iters %d)\n", cudaGetErrorString(e), b.bx, b.by, nw, depth, per_lane, p.iters); return 1; }
                probe<<<148, 32 * nw, smem>>>(p);
                cudaDeviceSynchronize();
                unsigned hs[2];
                cudaMemcpy(hs, sink, 8, cudaMemcpyDeviceToHost);
                if (hs[1]) { printf("wait timed out at batch %u (box %dx%d, %d warps, depth %d, per_lane %d)\n", hs[1] - 1, b.bx, b.by, nw, depth, per_lane); cudaMemset(sink, 0, 8); }
                std::vector<long long> h(148);
                cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
                double mean = 0;
                for (long long c : h) mean += (double)c / 148;
                const double nbox = (double)p.iters * depth * nw, cpb = mean / nbox;
                printf("%5d %5d %6d %5d %-9s %12.1f %13.1f %12.0f\n", b.bx, b.by, nw, depth, per_lane ? "per-lane" : "lane 0", cpb, 1965.0 / cpb,
                       148.0 * p.box_bytes * 1.965 / cpb);
            }
    }
    return 0;
}
