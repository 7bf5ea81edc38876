// hevcasm_b200 - TMA (cp.async.bulk.tensor) staging of byte-plane tiles: host-side descriptor, device-side issue + wait.
//
// A frame batch is described to the TMA unit as a 3-D uint8 tensor (x, y, frame).  Tile coordinates are ELEMENT
// granular, so a search window or a filter footprint that starts at any byte offset lands in shared memory as a dense
// box with no per-thread address arithmetic at all, and everything outside the declared extent is zero-filled by the
// hardware instead of being read - that replaces both the alignment handling and the edge clamping of a load loop.
#pragma once

#include <cuda.h>  // CUtensorMap and enums only; the encoder is fetched from the driver at run time (no -lcuda)
#include <string.h>

#include "common.cuh"

namespace hv {
namespace tma {

// ---- host ------------------------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encoder()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// cuTensorMapEncodeTiled behind a small per-thread cache: a codec calls the same entry point on the same frame stores over and over,
// and re-encoding two or three descriptors on every call was measurable host overhead per launch.  Direct-mapped, keyed by every
// argument of the encoder; a hit copies 128 bytes.
struct EncodeKey {
    uint64_t base, dim[4], stride[3];
    uint32_t box[4], dtype, rank, swizzle, pad;
    bool operator==(const EncodeKey &o) const { return memcmp(this, &o, sizeof *this) == 0; }
};
inline CUresult encode_cached(CUtensorMap *map, CUtensorMapDataType dtype, cuuint32_t rank, void *base, const cuuint64_t *dim, const cuuint64_t *stride,
                              const cuuint32_t *box, const cuuint32_t *estr, CUtensorMapSwizzle swizzle)
{
    struct Slot {
        EncodeKey key;
        CUtensorMap map;
        bool valid;
    };
    constexpr int SLOTS = 64;
    static thread_local Slot cache[SLOTS];
    EncodeKey k;
    memset(&k, 0, sizeof k);
    k.base = (uint64_t)(uintptr_t)base, k.dtype = (uint32_t)dtype, k.rank = rank, k.swizzle = (uint32_t)swizzle;
    for (cuuint32_t i = 0; i < rank; ++i) k.dim[i] = dim[i], k.box[i] = box[i];
    for (cuuint32_t i = 0; i + 1 < rank; ++i) k.stride[i] = stride[i];
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof k / 8; ++i) h = (h ^ reinterpret_cast<const uint64_t *>(&k)[i]) * 1099511628211ull;
    Slot &s = cache[(h >> 20) % SLOTS];
    if (s.valid && s.key == k) {
        *map = s.map;
        return CUDA_SUCCESS;
    }
    const CUresult r = encoder()(map, dtype, rank, base, dim, stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) s.key = k, s.map = *map, s.valid = true;
    return r;
}

// can a plane batch with these byte strides be described at all?  (row / frame strides must be multiples of 16 bytes)
inline bool describable(ptrdiff_t row_stride, ptrdiff_t frame_stride, int n_frames)
{
    return encoder() && row_stride > 0 && (row_stride & 15) == 0 && (n_frames <= 1 || (frame_stride > 0 && (frame_stride & 15) == 0));
}

// Describes the bytes [first, first + extent_x) x extent_y rows x n_frames, where `first` may have any alignment: the
// tensor starts at the enclosing 16-byte boundary and *x_shift receives the offset to add to every x coordinate.
// box = {box_x bytes (multiple of 16, <= 256), box_y rows (<= 256), 1 frame}.
inline int describe_bytes(CUtensorMap *map, const uint8_t *first, ptrdiff_t row_stride, ptrdiff_t frame_stride, long long extent_x, long long extent_y,
                          int n_frames, int box_x, int box_y, int *x_shift, int elem /* 1: uint8 elements, 4: 32-bit words */)
{
    const uintptr_t a = (uintptr_t)first;
    const int shift = (int)(a & 15);
    *x_shift = shift;
    if (n_frames <= 1) frame_stride = row_stride * (ptrdiff_t)extent_y;  // unused, but must be a legal stride
    long long row_bytes = extent_x + shift;
    if (row_bytes > (long long)row_stride) row_bytes = (long long)row_stride;  // never describe more than one row's worth per row
    cuuint64_t dim[3] = {(cuuint64_t)((row_bytes + elem - 1) / elem), (cuuint64_t)extent_y, (cuuint64_t)(n_frames < 1 ? 1 : n_frames)};
    cuuint64_t stride[2] = {(cuuint64_t)row_stride, (cuuint64_t)frame_stride};
    cuuint32_t box[3] = {(cuuint32_t)(box_x / elem), (cuuint32_t)box_y, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode_cached(map, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)(a - shift), dim, stride, box, estr,
                                     CU_TENSOR_MAP_SWIZZLE_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}
inline int describe_u8(CUtensorMap *map, const uint8_t *first, ptrdiff_t row_stride, ptrdiff_t frame_stride, long long extent_x, long long extent_y,
                       int n_frames, int box_x, int box_y, int *x_shift)
{
    return describe_bytes(map, first, row_stride, frame_stride, extent_x, extent_y, n_frames, box_x, box_y, x_shift, 1);
}
// the same bytes described as 32-bit words: boxes may then be up to 1024 bytes wide (box_x still in BYTES, a multiple of 16);
// the x coordinate of a load is in words, and the zero-fill edge is rounded up to a word
inline int describe_u32(CUtensorMap *map, const uint8_t *first, ptrdiff_t row_stride, ptrdiff_t frame_stride, long long extent_x, long long extent_y,
                        int n_frames, int box_x, int box_y, int *x_shift)
{
    return describe_bytes(map, first, row_stride, frame_stride, extent_x, extent_y, n_frames, box_x, box_y, x_shift, 4);
}

// Byte planes whose boxes land in shared memory in the tensor cores' swizzled K-major form: box_x = 128 -> 128-byte swizzle
// (destination 1024-byte aligned), box_x = 64 -> 64-byte swizzle (512-byte aligned), box_x = 32 -> 32-byte swizzle (256-byte aligned).
// `base` must be 16-byte aligned.  (Also used for TMA STORES out of a swizzled staging buffer, which threads can fill without bank conflicts.)
inline int describe_u8_swizzled(CUtensorMap *map, const uint8_t *base, ptrdiff_t row_stride, ptrdiff_t frame_stride, long long extent_x, long long extent_y,
                                int n_frames, int box_x, int box_y)
{
    if (((uintptr_t)base & 15) != 0 || (box_x != 128 && box_x != 64 && box_x != 32)) return (int)cudaErrorInvalidValue;
    if (n_frames <= 1) frame_stride = row_stride * (ptrdiff_t)extent_y;
    if (extent_x > (long long)row_stride) extent_x = (long long)row_stride;
    cuuint64_t dim[3] = {(cuuint64_t)extent_x, (cuuint64_t)extent_y, (cuuint64_t)(n_frames < 1 ? 1 : n_frames)};
    cuuint64_t stride[2] = {(cuuint64_t)row_stride, (cuuint64_t)frame_stride};
    cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode_cached(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)base, dim, stride, box, estr,
                                     box_x == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : box_x == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

// The 128-byte-swizzled form with the bytes described as 32-bit words (box = 32 words x box_y rows; the x coordinate of a load is in
// words): the TMA unit moves word tensors faster than byte tensors.  The zero-fill edge is rounded up to a word.
inline int describe_u32_swizzled128(CUtensorMap *map, const uint8_t *base, ptrdiff_t row_stride, ptrdiff_t frame_stride, long long extent_x, long long extent_y,
                                    int n_frames, int box_y)
{
    if (((uintptr_t)base & 15) != 0) return (int)cudaErrorInvalidValue;
    if (n_frames <= 1) frame_stride = row_stride * (ptrdiff_t)extent_y;
    if (extent_x > (long long)row_stride) extent_x = (long long)row_stride;
    cuuint64_t dim[3] = {(cuuint64_t)((extent_x + 3) / 4), (cuuint64_t)extent_y, (cuuint64_t)(n_frames < 1 ? 1 : n_frames)};
    cuuint64_t stride[2] = {(cuuint64_t)row_stride, (cuuint64_t)frame_stride};
    cuuint32_t box[3] = {32, (cuuint32_t)box_y, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode_cached(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)base, dim, stride, box, estr, CU_TENSOR_MAP_SWIZZLE_128B);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

// The same planes as a 4-D tensor (16 bytes, rows, 16-byte chunks of a row, frames): a box {16, R, C, 1} then lands in shared
// memory as [chunk][row][16 bytes] - the tensor cores' no-swizzle K-major core-matrix layout (8 rows x 16 bytes contiguous,
// chunks R*16 bytes apart).  `base` must be 16-byte aligned; the encoder accepts the (row stride, 16) stride order
// (tools/umma_fir_probe.cu).
inline int describe_u8_chunks(CUtensorMap *map, const uint8_t *base, ptrdiff_t row_stride, ptrdiff_t frame_stride, long long rows, long long chunks, int n_frames,
                              int box_rows, int box_chunks)
{
    if (((uintptr_t)base & 15) != 0) return (int)cudaErrorInvalidValue;
    if (n_frames <= 1) frame_stride = row_stride * (ptrdiff_t)rows;
    cuuint64_t dim[4] = {16, (cuuint64_t)rows, (cuuint64_t)chunks, (cuuint64_t)(n_frames < 1 ? 1 : n_frames)};
    cuuint64_t stride[3] = {(cuuint64_t)row_stride, 16, (cuuint64_t)frame_stride};
    cuuint32_t box[4] = {16, (cuuint32_t)box_rows, (cuuint32_t)box_chunks, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode_cached(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void *)base, dim, stride, box, estr, CU_TENSOR_MAP_SWIZZLE_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

// ---- device ----------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
// One arrival for the whole (converged) warp: the barrier's count is the number of WARPS.  Every arrival is a serialised
// shared-memory atomic - with one per thread, the 512 consumers of a tensor-core kernel spent ~1300 cycles per tile just arriving
// (measured with every other stage of the kernel switched off, profiles/r02_pred.md).  __syncwarp orders the lanes' earlier
// shared-memory writes and fences before lane 0's release-arrive.
__device__ __forceinline__ void mbar_arrive_warp(uint64_t *bar)
{
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
// Waits for the phase with the given parity.  A descriptor / coordinate error would leave the barrier incomplete for ever;
// rather than hang the GPU the wait gives up after ~1 s worth of polls and traps, which surfaces as a launch failure.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    // fast path: the phase has usually completed long before the wait (pipelines run ahead); test_wait never suspends the warp
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
#pragma unroll 1
    for (int spin = 0; spin < (1 << 22); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
// one box of a 3-D tensor -> shared memory (dense rows of box_x bytes); completion is signalled on `bar`
__device__ __forceinline__ void load_box_3d(void *smem_dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem_dst)),
                 "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}
// one dense box in shared memory (rows of box_x bytes) -> a 3-D tensor; the part of the box outside the tensor's extent is not written.
// The writes of the shared-memory bytes must have been made visible to the async proxy (fence.proxy.async) before this is issued.
__device__ __forceinline__ void store_box_3d(const CUtensorMap *map, int x, int y, int z, const void *smem_src)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(x), "r"(y), "r"(z),
                 "r"(smem_u32(smem_src))
                 : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// waits until at most PENDING of this thread's committed store groups are still READING their shared-memory source
template <int PENDING>
__device__ __forceinline__ void store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory");
}
// `bytes` contiguous bytes of global memory -> shared memory (both 16-byte aligned, bytes a multiple of 16); completion on `bar`
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void prefetch_descriptor(const CUtensorMap *map) { asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }

}  // namespace tma
}  // namespace hv
