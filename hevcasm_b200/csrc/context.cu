// hevcasm_b200 - host-memory forms of the batched entry points (hevcasm_batch.h, "host-memory forms").
//
// A hevcasm_cuda_context owns three streams (copy-in, compute, copy-out), an event pool and one device arena.  A *_host
// call cuts the batch into chunks of whole frames, and runs them through a ring of arena slots as a three-stage
// pipeline: chunk f+1 travels host->device while chunk f is computed and the results of chunk f-1 travel device->host.
// The data path is the same kernels the device-pointer entry points launch - there is no CPU arithmetic here.
//
// Host buffers should be page-locked (hevcasm_cuda_host_alloc) for the copies to overlap; pageable memory works but
// serialises.
#include "common.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <vector>

#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

struct hevcasm_cuda_context {
    int device = 0;
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    uint8_t *arena = nullptr;
    size_t arena_bytes = 0;
    std::vector<cudaEvent_t> events;
    size_t next_event = 0;

    cudaEvent_t event()
    {
        if (next_event == events.size()) {
            cudaEvent_t e;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
            events.push_back(e);
        }
        return events[next_event++];
    }
};

namespace {

constexpr size_t kAlign = 256;
inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// carves 256-byte aligned pieces out of one arena slot
struct Carver {
    uint8_t *base;
    size_t used = 0;
    template <class T>
    T *take(size_t bytes)
    {
        T *p = reinterpret_cast<T *>(base + used);
        used += round_up(bytes, kAlign);
        return p;
    }
};

// a padded plane batch on the device: `frames` planes of (height + 2*pad) rows x pitch bytes, elem bytes per sample
struct DevPlanes {
    size_t pitch_elems, rows, frame_elems, elem;
    int pad;
    size_t bytes(int frames) const { return (size_t)frames * frame_elems * elem; }
    size_t origin() const { return (size_t)pad * pitch_elems + pad; }  // element offset of sample (0,0)
};

DevPlanes plan_planes(int width, int height, int pad, size_t elem)
{
    DevPlanes d;
    d.elem = elem, d.pad = pad;
    d.pitch_elems = round_up(((size_t)width + 2 * pad) * elem, kAlign) / elem;
    d.rows = (size_t)height + 2 * pad;
    d.frame_elems = d.pitch_elems * d.rows;
    return d;
}

// host (sample (0,0) of frame f0 at h, row stride hs, frame stride hfs; all in elements) -> device planes, `frames` frames
int copy_in(const DevPlanes &d, void *dev, const void *h, ptrdiff_t hs, ptrdiff_t hfs, int width, int frames, cudaStream_t s)
{
    for (int f = 0; f < frames; ++f) {
        const uint8_t *src = (const uint8_t *)h + ((ptrdiff_t)f * hfs - (ptrdiff_t)d.pad * hs - d.pad) * (ptrdiff_t)d.elem;
        uint8_t *dst = (uint8_t *)dev + (size_t)f * d.frame_elems * d.elem;
        HV_CUDA(cudaMemcpy2DAsync(dst, d.pitch_elems * d.elem, src, (size_t)hs * d.elem, ((size_t)width + 2 * d.pad) * d.elem, d.rows,
                                  cudaMemcpyHostToDevice, s));
    }
    return 0;
}

// device planes -> host, interior (width x height) only
int copy_out(const DevPlanes &d, const void *dev, void *h, ptrdiff_t hs, ptrdiff_t hfs, int width, int height, int frames, cudaStream_t s)
{
    for (int f = 0; f < frames; ++f) {
        uint8_t *dst = (uint8_t *)h + (ptrdiff_t)f * hfs * (ptrdiff_t)d.elem;
        const uint8_t *src = (const uint8_t *)dev + ((size_t)f * d.frame_elems + d.origin()) * d.elem;
        HV_CUDA(cudaMemcpy2DAsync(dst, (size_t)hs * d.elem, src, d.pitch_elems * d.elem, (size_t)width * d.elem, height, cudaMemcpyDeviceToHost, s));
    }
    return 0;
}

// Three-stage pipeline over chunks of whole frames; stage callbacks enqueue on the stream they are given.  The arena is a byte ring: a
// chunk takes the next `frames * per_frame` bytes (wrapping to the start when the tail is too short), and its copy-in first waits for the
// copy-out of every earlier chunk whose bytes it overlaps.
//   Chunk sizes: `out_in_ratio` = result bytes / input bytes per frame.  When the results are the larger side (the SAD sweeps), the
// copy-out stream is the bottleneck and its rate depends on the size of the copies (one B200 of this pool, with a copy-in running beside it:
// 46.7 GB/s in 128 copies of 6 MB, 50.0 in 32, 51.8 in one; tools/pcie_probe.py) - so chunks GROW slowly: the first is one frame (the
// pipeline fills in one frame's copy-in) and chunk c has 1 + 0.15 * (frames before c) frames, 14 chunks for 32 frames.  Measured on the 32-frame
// 4K sweep (packed results), same box: one frame per chunk 15.5-16.0 Gsamples/s, two or four 16.0, growth 0.05 / 0.10 / 0.15 / 0.20 / 0.37:
// 16.2 / 16.6 / 16.6-16.8 / 15.8 / 15.0 (larger chunks make the copy-out wait for the next chunk's copy-in).  Otherwise chunks are ~32 MB as
// before (>= 3 in flight).
template <class In, class Run, class Out>
int run_pipeline(hevcasm_cuda_context *ctx, int n_frames, size_t bytes_per_frame, double out_in_ratio, In in, Run run, Out out)
{
    if (n_frames == 0) return 0;
    HV_CUDA(cudaSetDevice(ctx->device));
    const size_t per_frame = round_up(bytes_per_frame, kAlign) + 16 * kAlign;  // slack for per-piece alignment
    if (per_frame > ctx->arena_bytes) return HEVCASM_ERR_ARGUMENT;               // arena cannot hold even one frame
    const int fit = (int)std::min<size_t>(ctx->arena_bytes / per_frame, (size_t)n_frames);   // frames the arena holds at once
    const int cap = std::max(1, fit / 3);                                                    // >= 3 chunks in flight so the three stages overlap
    const int flat = std::max(1, std::min(cap, std::max(1, (int)((size_t)(32u << 20) / per_frame))));
    const bool grow = out_in_ratio > 1.05;
    constexpr double kGrowth = 0.15;   // frames(c) = 1 + kGrowth * frames before c
    ctx->next_event = 0;
    struct Busy {
        size_t begin, end;
        cudaEvent_t done;
    };
    std::vector<Busy> busy;   // chunks whose copy-out may still be reading their bytes, oldest first
    size_t head = 0;
    auto enqueue = [&]() -> int {
        for (int f0 = 0; f0 < n_frames;) {
            int nf = grow ? std::min(cap, 1 + (int)(kGrowth * f0)) : flat;
            nf = std::min(nf, n_frames - f0);
            const size_t bytes = (size_t)nf * per_frame;
            if (head + bytes > ctx->arena_bytes) head = 0;
            const size_t begin = head, end = head + bytes;
            for (size_t i = 0; i < busy.size();) {
                if (busy[i].begin < end && begin < busy[i].end) {
                    HV_CUDA(cudaStreamWaitEvent(ctx->s_in, busy[i].done, 0));   // those bytes are free again once that chunk's results have left
                    busy.erase(busy.begin() + i);
                } else {
                    ++i;
                }
            }
            uint8_t *slot = ctx->arena + begin;
            int e = in(slot, f0, nf, ctx->s_in);
            if (e) return e;
            cudaEvent_t ev_in = ctx->event(), ev_run = ctx->event(), ev_out = ctx->event();
            if (!ev_in || !ev_run || !ev_out) return (int)cudaErrorMemoryAllocation;
            HV_CUDA(cudaEventRecord(ev_in, ctx->s_in));
            HV_CUDA(cudaStreamWaitEvent(ctx->s_run, ev_in, 0));
            e = run(slot, f0, nf, ctx->s_run);
            if (e) return e;
            HV_CUDA(cudaEventRecord(ev_run, ctx->s_run));
            HV_CUDA(cudaStreamWaitEvent(ctx->s_out, ev_run, 0));
            e = out(slot, f0, nf, ctx->s_out);
            if (e) return e;
            HV_CUDA(cudaEventRecord(ev_out, ctx->s_out));
            busy.push_back({begin, end, ev_out});
            head = end;
            f0 += nf;
        }
        return 0;
    };
    const int e = enqueue();
    // success or not, nothing may still be reading or writing the caller's buffers (or be waiting on this call's events) when we return
    const cudaError_t s1 = cudaStreamSynchronize(ctx->s_in), s2 = cudaStreamSynchronize(ctx->s_run), s3 = cudaStreamSynchronize(ctx->s_out);
    if (e) return e;
    return s1 != cudaSuccess ? (int)s1 : s2 != cudaSuccess ? (int)s2 : (int)s3;
}

}  // namespace

extern "C" hevcasm_cuda_context *hevcasm_cuda_context_create(int device, size_t arena_bytes)
{
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    hevcasm_cuda_context *ctx = new (std::nothrow) hevcasm_cuda_context;
    if (!ctx) return nullptr;
    ctx->device = device;
    bool ok = cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_run, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking) == cudaSuccess;
    if (ok && arena_bytes) ok = cudaMalloc((void **)&ctx->arena, arena_bytes) == cudaSuccess;
    if (!ok) {
        hevcasm_cuda_context_destroy(ctx);
        return nullptr;
    }
    ctx->arena_bytes = arena_bytes;
    return ctx;
}

extern "C" void hevcasm_cuda_context_destroy(hevcasm_cuda_context *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_run) cudaStreamDestroy(ctx->s_run);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    delete ctx;
}

extern "C" void *hevcasm_cuda_context_stream(hevcasm_cuda_context *ctx) { return ctx ? (void *)ctx->s_run : nullptr; }

// ---- page-locked host memory --------------------------------------------------------------------------------
// hevcasm_cuda_host_alloc: cudaHostAlloc.  hevcasm_cuda_host_alloc_near(bytes, device): pages bound to the NUMA node the GPU hangs off
// (sysfs numa_node of its PCI function; mbind) and then registered with the driver - on a two-socket 8-GPU box the D2H result streams
// of the ranks otherwise all land on whichever node the allocating thread happened to run on, and the end-to-end figure stops scaling
// at that node's memory bandwidth (SCALE_r01: 0.20 efficiency at 8 GPUs).  Falls back to cudaHostAlloc when the node is unknown.
namespace {
std::mutex g_host_mu;
std::map<void *, size_t> g_mapped;   // regions from mmap + cudaHostRegister: pointer -> mapped bytes

int gpu_numa_node(int device)
{
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return -1;
    for (char *c = bus; *c; ++c)
        if (*c >= 'A' && *c <= 'F') *c = (char)(*c - 'A' + 'a');   // sysfs spells the address in lower case
    char path[96];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}
}  // namespace

extern "C" void *hevcasm_cuda_host_alloc(size_t bytes)
{
    void *p = nullptr;
    return cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}

// the NUMA node hevcasm_cuda_host_alloc_near binds to for `device`, or -1 (unknown: plain page-locked memory is used)
extern "C" int hevcasm_cuda_device_numa_node(int device) { return gpu_numa_node(device); }

extern "C" void *hevcasm_cuda_host_alloc_near(size_t bytes, int device)
{
    const int node = gpu_numa_node(device);
    if (node < 0 || node >= 64 || bytes == 0) return hevcasm_cuda_host_alloc(bytes);
    const size_t page = (size_t)sysconf(_SC_PAGESIZE), len = round_up(bytes, page);
    void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return hevcasm_cuda_host_alloc(bytes);
    const unsigned long mask = 1ul << node;
#ifdef SYS_mbind
    const long rc = syscall(SYS_mbind, p, len, 2 /* MPOL_BIND */, &mask, 64ul + 1, 0u);
#else
    const long rc = -1;
#endif
    if (rc != 0) {   // no NUMA policy available (container without CAP_SYS_NICE, single node): plain page-locked memory
        munmap(p, len);
        return hevcasm_cuda_host_alloc(bytes);
    }
    memset(p, 0, len);   // fault the pages in on the bound node before the driver pins them
    if (cudaHostRegister(p, len, cudaHostRegisterPortable) != cudaSuccess) {
        (void)cudaGetLastError();
        munmap(p, len);
        return hevcasm_cuda_host_alloc(bytes);
    }
    std::lock_guard<std::mutex> lock(g_host_mu);
    g_mapped[p] = len;
    return p;
}

extern "C" void hevcasm_cuda_host_free(void *p)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> lock(g_host_mu);
        auto it = g_mapped.find(p);
        if (it != g_mapped.end()) {
            cudaHostUnregister(p);
            munmap(p, it->second);
            g_mapped.erase(it);
            return;
        }
    }
    cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------------------ SAD pyramid

// mode 0: 64 int32 SADs per PU (hevcasm_sad_sweep_pyramid_frames); 1: {min SAD, candidate} per PU (..._best_frames);
// 2: 64 SADs per PU, uint16 for the 8x8 and 16x16 levels (..._packed_frames)
static int sad_pyramid_host(int mode, hevcasm_cuda_context *ctx, const uint8_t *src, ptrdiff_t ss, const uint8_t *ref, ptrdiff_t sr, int width, int height, int pad,
                            int dx0, int dy0, int n_frames, ptrdiff_t fs_src, ptrdiff_t fs_ref, void *const host_out[4])
{
    if (!ctx || width < 8 || height < 8 || n_frames < 0 || pad < 0) return HEVCASM_ERR_ARGUMENT;
    // the window [dx0, dx0+8) x [dy0, dy0+8) must lie inside the padding that travels with the frames
    if (dx0 < -pad || dy0 < -pad || dx0 + 7 > pad || dy0 + 7 > pad) return HEVCASM_ERR_ARGUMENT;
    if (mode != 0 && !(host_out[0] && host_out[1] && host_out[2] && host_out[3])) return HEVCASM_ERR_ARGUMENT;
    const DevPlanes d = plan_planes(width, height, pad, 1);
    const size_t per_pu = mode == 1 ? 2 : 64;
    size_t out_bytes[4];  // per frame
    for (int l = 0; l < 4; ++l)
        out_bytes[l] = host_out[l] ? (size_t)(width >> (3 + l)) * (height >> (3 + l)) * per_pu * ((mode == 2 && l < 2) ? 2 : 4) : 0;
    size_t per_frame = 2 * (d.frame_elems + kAlign);
    for (int l = 0; l < 4; ++l) per_frame += out_bytes[l] + kAlign;

    struct Slot {
        uint8_t *src, *ref;
        uint8_t *out[4];
    };
    auto carve = [&](uint8_t *slot, int nf) {
        Carver c{slot};
        Slot s;
        s.src = c.take<uint8_t>(d.bytes(nf));
        s.ref = c.take<uint8_t>(d.bytes(nf));
        for (int l = 0; l < 4; ++l) s.out[l] = out_bytes[l] ? c.take<uint8_t>(out_bytes[l] * nf) : nullptr;
        return s;
    };
    const double out_in_ratio = (double)(out_bytes[0] + out_bytes[1] + out_bytes[2] + out_bytes[3]) / (double)(2 * d.frame_elems);
    return run_pipeline(
        ctx, n_frames, per_frame, out_in_ratio,
        [&](uint8_t *slot, int f0, int nf, cudaStream_t s) {
            const Slot k = carve(slot, nf);
            int e = copy_in(d, k.src, src + (ptrdiff_t)f0 * fs_src, ss, fs_src, width, nf, s);
            return e ? e : copy_in(d, k.ref, ref + (ptrdiff_t)f0 * fs_ref, sr, fs_ref, width, nf, s);
        },
        [&](uint8_t *slot, int, int nf, cudaStream_t s) {
            const Slot k = carve(slot, nf);
            const uint8_t *a = k.src + d.origin(), *b = k.ref + d.origin();
            const ptrdiff_t pitch = (ptrdiff_t)d.pitch_elems, fs = (ptrdiff_t)d.frame_elems;
            if (mode == 0)
                return hevcasm_sad_sweep_pyramid_frames(a, pitch, b, pitch, width, height, dx0, dy0, nf, fs, fs, (int32_t *)k.out[0], (int32_t *)k.out[1],
                                                        (int32_t *)k.out[2], (int32_t *)k.out[3], s);
            if (mode == 1)
                return hevcasm_sad_sweep_pyramid_best_frames(a, pitch, b, pitch, width, height, dx0, dy0, nf, fs, fs, (int32_t *)k.out[0], (int32_t *)k.out[1],
                                                             (int32_t *)k.out[2], (int32_t *)k.out[3], s);
            return hevcasm_sad_sweep_pyramid_packed_frames(a, pitch, b, pitch, width, height, dx0, dy0, nf, fs, fs, (uint16_t *)k.out[0], (uint16_t *)k.out[1],
                                                           (int32_t *)k.out[2], (int32_t *)k.out[3], s);
        },
        [&](uint8_t *slot, int f0, int nf, cudaStream_t s) {
            const Slot k = carve(slot, nf);
            for (int l = 0; l < 4; ++l)
                if (host_out[l] && out_bytes[l])
                    HV_CUDA(cudaMemcpyAsync((uint8_t *)host_out[l] + (size_t)f0 * out_bytes[l], k.out[l], out_bytes[l] * nf, cudaMemcpyDeviceToHost, s));
            return 0;
        });
}

extern "C" int hevcasm_sad_sweep_pyramid_frames_host(hevcasm_cuda_context *ctx, const uint8_t *src, ptrdiff_t ss, const uint8_t *ref, ptrdiff_t sr,
                                                     int width, int height, int pad, int dx0, int dy0, int n_frames, ptrdiff_t fs_src,
                                                     ptrdiff_t fs_ref, int32_t *sad8, int32_t *sad16, int32_t *sad32, int32_t *sad64)
{
    void *const out[4] = {sad8, sad16, sad32, sad64};
    return sad_pyramid_host(0, ctx, src, ss, ref, sr, width, height, pad, dx0, dy0, n_frames, fs_src, fs_ref, out);
}

extern "C" int hevcasm_sad_sweep_pyramid_best_frames_host(hevcasm_cuda_context *ctx, const uint8_t *src, ptrdiff_t ss, const uint8_t *ref, ptrdiff_t sr,
                                                          int width, int height, int pad, int dx0, int dy0, int n_frames, ptrdiff_t fs_src,
                                                          ptrdiff_t fs_ref, int32_t *best8, int32_t *best16, int32_t *best32, int32_t *best64)
{
    void *const out[4] = {best8, best16, best32, best64};
    return sad_pyramid_host(1, ctx, src, ss, ref, sr, width, height, pad, dx0, dy0, n_frames, fs_src, fs_ref, out);
}

extern "C" int hevcasm_sad_sweep_pyramid_packed_frames_host(hevcasm_cuda_context *ctx, const uint8_t *src, ptrdiff_t ss, const uint8_t *ref, ptrdiff_t sr,
                                                            int width, int height, int pad, int dx0, int dy0, int n_frames, ptrdiff_t fs_src,
                                                            ptrdiff_t fs_ref, uint16_t *sad8, uint16_t *sad16, int32_t *sad32, int32_t *sad64)
{
    void *const out[4] = {sad8, sad16, sad32, sad64};
    return sad_pyramid_host(2, ctx, src, ss, ref, sr, width, height, pad, dx0, dy0, n_frames, fs_src, fs_ref, out);
}

// ------------------------------------------------------------------------------------------------ inter prediction planes

extern "C" int hevcasm_pred_uni_frames_host(hevcasm_cuda_context *ctx, uint8_t *dst, ptrdiff_t sd, const uint8_t *ref, ptrdiff_t sr, int width, int height,
                                            int pad, int taps, int xFrac, int yFrac, int n_frames, ptrdiff_t fs_dst, ptrdiff_t fs_ref)
{
    if (!ctx || width <= 0 || height <= 0 || n_frames < 0 || (taps != 8 && taps != 4)) return HEVCASM_ERR_ARGUMENT;
    // the filter reads taps/2-1 samples before and taps/2 after the plane; the aligned kernels touch whole 16-byte chunks
    if (pad < 16) return HEVCASM_ERR_ARGUMENT;
    const DevPlanes din = plan_planes(width, height, pad, 1), dout = plan_planes(width, height, 0, 1);
    const size_t per_frame = din.frame_elems + dout.frame_elems + 2 * kAlign;
    struct Slot {
        uint8_t *ref, *dst;
    };
    auto carve = [&](uint8_t *slot, int nf) {
        Carver c{slot};
        Slot s;
        s.ref = c.take<uint8_t>(din.bytes(nf));
        s.dst = c.take<uint8_t>(dout.bytes(nf));
        return s;
    };
    return run_pipeline(
        ctx, n_frames, per_frame, 0.0,
        [&](uint8_t *slot, int f0, int nf, cudaStream_t s) { return copy_in(din, carve(slot, nf).ref, ref + (ptrdiff_t)f0 * fs_ref, sr, fs_ref, width, nf, s); },
        [&](uint8_t *slot, int, int nf, cudaStream_t s) {
            const Slot k = carve(slot, nf);
            return hevcasm_pred_uni_frames(k.dst, dout.pitch_elems, k.ref + din.origin(), din.pitch_elems, width, height, taps, xFrac, yFrac, nf, dout.frame_elems,
                                           din.frame_elems, s);
        },
        [&](uint8_t *slot, int f0, int nf, cudaStream_t s) {
            return copy_out(dout, carve(slot, nf).dst, dst + (ptrdiff_t)f0 * fs_dst, sd, fs_dst, width, height, nf, s);
        });
}

// ------------------------------------------------------------------------------------------------ fused residual pipeline

extern "C" int hevcasm_residual_pipeline_frames_host(hevcasm_cuda_context *ctx, uint8_t *rec, ptrdiff_t s_rec, int16_t *levels, int32_t *cbf,
                                                     const int16_t *residual, ptrdiff_t s_res, const uint8_t *pred, ptrdiff_t s_pred, int width, int height,
                                                     int log2size, int trType, int q_scale, int q_shift, int q_offset, int iq_scale, int iq_shift,
                                                     int n_frames, ptrdiff_t fs_rec, ptrdiff_t fs_res, ptrdiff_t fs_pred)
{
    if (!ctx || width <= 0 || height <= 0 || n_frames < 0 || log2size < 2 || log2size > 5 || !levels) return HEVCASM_ERR_ARGUMENT;
    const int n = 1 << log2size;
    const size_t nb = (size_t)(width >> log2size) * (height >> log2size);  // blocks per frame
    const DevPlanes d8 = plan_planes(width, height, 0, 1), d16 = plan_planes(width, height, 0, 2);
    const size_t per_frame = 2 * d8.frame_elems + d16.frame_elems * 2 + nb * n * n * 2 + nb * 4 + 5 * kAlign;
    struct Slot {
        uint8_t *pred, *rec;
        int16_t *res, *levels;
        int32_t *cbf;
    };
    auto carve = [&](uint8_t *slot, int nf) {
        Carver c{slot};
        Slot s;
        s.pred = c.take<uint8_t>(d8.bytes(nf));
        s.rec = c.take<uint8_t>(d8.bytes(nf));
        s.res = c.take<int16_t>(d16.bytes(nf));
        s.levels = c.take<int16_t>(nb * n * n * 2 * nf);
        s.cbf = c.take<int32_t>(nb * 4 * nf);
        return s;
    };
    return run_pipeline(
        ctx, n_frames, per_frame, 0.0,
        [&](uint8_t *slot, int f0, int nf, cudaStream_t s) {
            const Slot k = carve(slot, nf);
            int e = copy_in(d8, k.pred, pred + (ptrdiff_t)f0 * fs_pred, s_pred, fs_pred, width, nf, s);
            return e ? e : copy_in(d16, k.res, residual + (ptrdiff_t)f0 * fs_res, s_res, fs_res, width, nf, s);
        },
        [&](uint8_t *slot, int, int nf, cudaStream_t s) {
            const Slot k = carve(slot, nf);
            return hevcasm_residual_pipeline_frames(k.rec, d8.pitch_elems, k.levels, k.cbf, k.res, d16.pitch_elems, k.pred, d8.pitch_elems, width, height, log2size,
                                                    trType, q_scale, q_shift, q_offset, iq_scale, iq_shift, nf, d8.frame_elems, d16.frame_elems, d8.frame_elems, s);
        },
        [&](uint8_t *slot, int f0, int nf, cudaStream_t s) {
            const Slot k = carve(slot, nf);
            // only the area the block grid covers is defined (and may be written)
            int e = copy_out(d8, k.rec, rec + (ptrdiff_t)f0 * fs_rec, s_rec, fs_rec, (width >> log2size) << log2size, (height >> log2size) << log2size, nf, s);
            if (e) return e;
            HV_CUDA(cudaMemcpyAsync(levels + (size_t)f0 * nb * n * n, k.levels, nb * n * n * 2 * nf, cudaMemcpyDeviceToHost, s));
            if (cbf) HV_CUDA(cudaMemcpyAsync(cbf + (size_t)f0 * nb, k.cbf, nb * 4 * nf, cudaMemcpyDeviceToHost, s));
            return 0;
        });
}
