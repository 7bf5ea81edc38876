/*
 * hevcasm_b200 - batched, stream-aware GPU entry points (the extension SURVEY.md section 8(b) asks for).
 *
 * One block per call cannot feed a GPU, so every per-block function type of the reference's function-select
 * API gets batched siblings here.  The semantics of EACH batch element are exactly those of the per-block
 * reference function cited next to the entry point; outputs are bit-exact with the reference's C path.
 *
 * Conventions (all entry points):
 *   - C linkage, plain pointers and sizes; every data pointer is a DEVICE pointer (cudaMalloc'ed or
 *     otherwise device-accessible); strides are in ELEMENTS of the pointed-to type, like the reference's;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream); work is only enqueued -
 *     the call returns before it runs; the library allocates nothing and keeps no pointer;
 *   - the return value is 0 on success, otherwise a cudaError_t value (hevcasm_cuda_error_string()) or
 *     HEVCASM_ERR_ARGUMENT for a shape the entry point does not serve.  (The reference's kernels cannot fail;
 *     a launch can.)
 *   - "_batch" variants take an explicit list of block positions, "_frames" variants process every block of
 *     a regular grid over `n_frames` planes laid out `frame stride` elements apart (one launch for the lot -
 *     that is how a 4K/8K batch reaches HBM speed);
 *   - frames must be padded like the reference's test buffers (reference pred_inter.c:468-471): the interpolation
 *     FOOTPRINT is taps/2-1 samples left/above and taps/2 right/below the block or plane, the SAD entry points read the
 *     candidate window around each PU.  The interpolation plane entry points (hevcasm_pred_*_frames) and the PU-list entry
 *     points additionally require 16 READABLE bytes before the first and after the last byte of that footprint: their
 *     aligned, TMA and tensor-core kernels move whole 16-byte chunks (up to 16 bytes left and 15 bytes right of a row's
 *     footprint - inside the plane these are neighbouring samples), and the list kernels always read the two-pass footprint,
 *     full-sample motion vectors included.  Any frame store with >= 16 samples and >= 4 rows of padding satisfies this; a
 *     caller who cannot promise it uses the *_bounded plane forms, which then stay inside the footprint;
 *   - position lists are int16_t pairs {x, y} (an 8K plane fits), candidate / motion vectors int16_t {dx, dy}.
 *
 * Host-memory convenience forms (host pointers, H2D/D2H inside) live behind hevcasm_cuda_context at the end.
 */
#ifndef INCLUDED_hevcasm_batch_h
#define INCLUDED_hevcasm_batch_h

#include "hevcasm.h"

#ifdef __cplusplus
extern "C" {
#endif

#define HEVCASM_ERR_ARGUMENT (-1)

const char HEVCASM_API *hevcasm_cuda_error_string(int code);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
unsigned long long HEVCASM_API hevcasm_cuda_launch_count(void);

/* ------------------------------------------------------------------------------------------------ SAD
 * element semantics: reference sad.h:50 / sad.c:47-60 (hevcasm_sad) and sad.h:95 / sad.c:101-121
 * (hevcasm_sad_multiref, 4 references per call).  rect = HEVCASM_RECT(w, h), w and h multiples of 4, <= 64. */

/* sad[i][c] = SAD(src block at pu_xy[i], ref block at pu_xy[i] + cand_dxdy[c]);  i < n_pu, c < n_cand */
int HEVCASM_API hevcasm_sad_multiref_batch(const uint8_t *src, ptrdiff_t stride_src, const uint8_t *ref, ptrdiff_t stride_ref,
                                           uint32_t rect, const int16_t *pu_xy, int n_pu, const int16_t *cand_dxdy, int n_cand,
                                           int32_t *sad, void *stream);

/* sad[i] = SAD(src block at pu_xy[i], ref block at pu_xy[i] + mv_xy[i]); mv_xy may be NULL (co-located) */
int HEVCASM_API hevcasm_sad_batch(const uint8_t *src, ptrdiff_t stride_src, const uint8_t *ref, ptrdiff_t stride_ref, uint32_t rect,
                                  const int16_t *pu_xy, const int16_t *mv_xy, int n_pu, int32_t *sad, void *stream);

/* PU lists over a batch of frames in one launch: pus[i] = {x, y, w, h, dx, dy, frame} (int16 x 7; w, h multiples of 4 in 4..64; (dx, dy) an
 * integer displacement into the reference, 0 0 for a co-located cost).  out[i] = SAD / SSD of src block (x, y) of plane `frame` against
 * ref block (x + dx, y + dy) of plane `frame`; -1 for a descriptor with an illegal size.  Element semantics: sad.h:50, ssd.h:53. */
int HEVCASM_API hevcasm_sad_list_frames(const uint8_t *src, ptrdiff_t stride_src, const uint8_t *ref, ptrdiff_t stride_ref,
                                        const int16_t *pus, int n_pu, ptrdiff_t frame_stride_src, ptrdiff_t frame_stride_ref,
                                        int32_t *sad, void *stream);
int HEVCASM_API hevcasm_ssd_list_frames(const uint8_t *srcA, ptrdiff_t stride_srcA, const uint8_t *srcB, ptrdiff_t stride_srcB,
                                        const int16_t *pus, int n_pu, ptrdiff_t frame_stride_srcA, ptrdiff_t frame_stride_srcB,
                                        int32_t *ssd, void *stream);

/* Motion-estimation sweep: the floor(width/w) x floor(height/h) non-overlapping PUs of every frame, each against
 * the dense candidate window dx in [dx0, dx0+ncx), dy in [dy0, dy0+ncy).
 * sad[frame][py][px][(dy-dy0)*ncx + (dx-dx0)].  ncx*ncy <= 256. */
int HEVCASM_API hevcasm_sad_sweep_frames(const uint8_t *src, ptrdiff_t stride_src, const uint8_t *ref, ptrdiff_t stride_ref,
                                         int width, int height, uint32_t rect, int dx0, int dy0, int ncx, int ncy, int n_frames,
                                         ptrdiff_t frame_stride_src, ptrdiff_t frame_stride_ref, int32_t *sad, void *stream);

/* The same sweep for the four square PU sizes 8, 16, 32, 64 in ONE pass over the frames (the 8x8 partial sums are
 * composed on chip into the larger PUs).  Window fixed at 8 x 8 candidates: dx in [dx0, dx0+8), dy in [dy0, dy0+8).
 * width and height >= 8; a level covers the floor(width/s) x floor(height/s) whole PUs of its size s (so all four levels tile the
 * same area only when both are multiples of 64).  Any of the four outputs may be NULL.  Each output has the layout
 * hevcasm_sad_sweep_frames would produce for that size. */
int HEVCASM_API hevcasm_sad_sweep_pyramid_frames(const uint8_t *src, ptrdiff_t stride_src, const uint8_t *ref, ptrdiff_t stride_ref,
                                                 int width, int height, int dx0, int dy0, int n_frames,
                                                 ptrdiff_t frame_stride_src, ptrdiff_t frame_stride_ref, int32_t *sad8,
                                                 int32_t *sad16, int32_t *sad32, int32_t *sad64, void *stream);

/* The pyramid sweep with the motion-search argmin folded in (the step a caller runs right after the SAD kernel): for every PU
 * the smallest of its 64 SADs and the index c = (dy-dy0)*8 + (dx-dx0) of the candidate that reaches it - the FIRST such
 * candidate in raster order on ties, i.e. what a scan `if (sad[c] < best)` over hevcasm_sad_multiref results returns.
 * bestN[frame][py][px] = {sad, c} as two int32.  All four outputs are required; src must be 16-byte aligned and every stride a
 * multiple of 16 bytes (HEVCASM_ERR_ARGUMENT otherwise).  Writes 1/32 of the bytes of the full sweep. */
int HEVCASM_API hevcasm_sad_sweep_pyramid_best_frames(const uint8_t *src, ptrdiff_t stride_src, const uint8_t *ref, ptrdiff_t stride_ref,
                                                      int width, int height, int dx0, int dy0, int n_frames,
                                                      ptrdiff_t frame_stride_src, ptrdiff_t frame_stride_ref, int32_t *best8,
                                                      int32_t *best16, int32_t *best32, int32_t *best64, void *stream);

/* The full pyramid sweep with the two small levels packed: sad8 and sad16 are uint16_t - exact, an 8x8 SAD is at most 8*8*255 = 16 320 and
 * a 16x16 SAD at most 16*16*255 = 65 280 - while sad32 and sad64 stay int32_t.  Same layout and values as
 * hevcasm_sad_sweep_pyramid_frames, 2.81 instead of 5.31 output bytes per sample (what matters when the results cross PCIe).
 * All four outputs are required; src 16-byte aligned, strides multiples of 16 bytes, sad8 / sad16 8-byte and sad32 / sad64 16-byte
 * aligned (HEVCASM_ERR_ARGUMENT otherwise). */
int HEVCASM_API hevcasm_sad_sweep_pyramid_packed_frames(const uint8_t *src, ptrdiff_t stride_src, const uint8_t *ref, ptrdiff_t stride_ref,
                                                        int width, int height, int dx0, int dy0, int n_frames,
                                                        ptrdiff_t frame_stride_src, ptrdiff_t frame_stride_ref, uint16_t *sad8,
                                                        uint16_t *sad16, int32_t *sad32, int32_t *sad64, void *stream);

/* ------------------------------------------------------------------------------------------------ SSD
 * element semantics: reference ssd.h:53 / ssd.c:43-55, square blocks of size 1<<log2size (2..6) */
int HEVCASM_API hevcasm_ssd_batch(const uint8_t *srcA, ptrdiff_t stride_srcA, const uint8_t *srcB, ptrdiff_t stride_srcB,
                                  int log2size, const int16_t *blk_xy, int n, int32_t *ssd, void *stream);
/* ssd[frame][by][bx] over the regular grid of floor(width/N) x floor(height/N) blocks */
int HEVCASM_API hevcasm_ssd_frames(const uint8_t *srcA, ptrdiff_t stride_srcA, const uint8_t *srcB, ptrdiff_t stride_srcB, int width,
                                   int height, int log2size, int n_frames, ptrdiff_t frame_stride_srcA,
                                   ptrdiff_t frame_stride_srcB, int32_t *ssd, void *stream);

/* ------------------------------------------------------------------------------------------------ SATD, linear SSD
 * element semantics: reference hadamard.h:55 / hadamard.c:75-131 (compute_satd: 2-D Hadamard transform of the difference,
 * (N/4 + sum of absolute values) / (N/2)), blocks of size 1<<log2size with log2size 1..3; and diff.h:48 / diff.c:45-54
 * (hevcasm_ssd_linear over a flat run of `size` samples). */
int HEVCASM_API hevcasm_hadamard_satd_batch(const uint8_t *srcA, ptrdiff_t stride_srcA, const uint8_t *srcB, ptrdiff_t stride_srcB,
                                            int log2size, const int16_t *blk_xy, int n, int32_t *satd, void *stream);
/* the same over the block lists of a batch of frames: blks[i] = {x, y, frame}, bucketed by size - n_by_size[0] 2x2 blocks, then the 4x4 and
 * the 8x8 ones; satd[i] in list order.  One launch per size present. */
int HEVCASM_API hevcasm_hadamard_satd_list_frames(const uint8_t *srcA, ptrdiff_t stride_srcA, const uint8_t *srcB, ptrdiff_t stride_srcB,
                                                  const int16_t *blks, const int *n_by_size, ptrdiff_t frame_stride_srcA,
                                                  ptrdiff_t frame_stride_srcB, int32_t *satd, void *stream);
/* satd[frame][by][bx] over the regular grid of floor(width/N) x floor(height/N) blocks */
int HEVCASM_API hevcasm_hadamard_satd_frames(const uint8_t *srcA, ptrdiff_t stride_srcA, const uint8_t *srcB, ptrdiff_t stride_srcB,
                                             int width, int height, int log2size, int n_frames, ptrdiff_t frame_stride_srcA,
                                             ptrdiff_t frame_stride_srcB, int32_t *satd, void *stream);
/* ssd[i] = sum over `size` samples of (src0[i*run_stride0 + k] - src1[i*run_stride1 + k])^2, i < n_runs; size * 255^2 must fit int32 */
int HEVCASM_API hevcasm_ssd_linear_batch(const uint8_t *src0, ptrdiff_t run_stride0, const uint8_t *src1, ptrdiff_t run_stride1, int size,
                                         int n_runs, int32_t *ssd, void *stream);

/* ------------------------------------------------------------------------------------------------ inter prediction
 * element semantics: reference pred_inter.h:56 (hevcasm_pred_uni_8to8; C path pred_inter.c:90-228) and
 * pred_inter.h:76 (hevcasm_pred_bi_8to8; pred_inter.c:490-530).  taps = 8 (luma, fractions 0..3) or 4 (chroma, 0..7). */

/* whole planes with one fractional position: dst(x, y) = prediction from ref(x, y), 0 <= x < width, 0 <= y < height */
int HEVCASM_API hevcasm_pred_uni_frames(uint8_t *dst, ptrdiff_t stride_dst, const uint8_t *ref, ptrdiff_t stride_ref, int width,
                                        int height, int taps, int xFrac, int yFrac, int n_frames, ptrdiff_t frame_stride_dst,
                                        ptrdiff_t frame_stride_ref, void *stream);
int HEVCASM_API hevcasm_pred_bi_frames(uint8_t *dst, ptrdiff_t stride_dst, const uint8_t *ref0, const uint8_t *ref1,
                                       ptrdiff_t stride_ref, int width, int height, int taps, int xFrac0, int yFrac0, int xFrac1,
                                       int yFrac1, int n_frames, ptrdiff_t frame_stride_dst, ptrdiff_t frame_stride_ref,
                                       void *stream);

/* The same with the readable bytes around the footprint stated: slack_before / slack_after = bytes the caller guarantees to be readable
 * before the first byte (row -(taps/2-1), column -(taps/2-1) of frame 0) and after the last byte of the reference's own footprint.
 * With either below 16 only the kernels that stay inside the footprint run (slower; bit-identical results). */
int HEVCASM_API hevcasm_pred_uni_frames_bounded(uint8_t *dst, ptrdiff_t stride_dst, const uint8_t *ref, ptrdiff_t stride_ref, int width,
                                                int height, int taps, int xFrac, int yFrac, int n_frames, ptrdiff_t frame_stride_dst,
                                                ptrdiff_t frame_stride_ref, ptrdiff_t slack_before, ptrdiff_t slack_after, void *stream);
int HEVCASM_API hevcasm_pred_bi_frames_bounded(uint8_t *dst, ptrdiff_t stride_dst, const uint8_t *ref0, const uint8_t *ref1,
                                               ptrdiff_t stride_ref, int width, int height, int taps, int xFrac0, int yFrac0, int xFrac1,
                                               int yFrac1, int n_frames, ptrdiff_t frame_stride_dst, ptrdiff_t frame_stride_ref,
                                               ptrdiff_t slack_before, ptrdiff_t slack_after, void *stream);

/* prediction-unit lists.  Uni descriptor = 6 x int16 {x, y, w, h, mvx, mvy}; bi descriptor = 8 x int16
 * {x, y, w, h, mvx0, mvy0, mvx1, mvy1}.  Motion vectors are in 1/4 (taps 8) or 1/8 (taps 4) sample units:
 * integer part mv >> 2 (>> 3), fraction mv & 3 (& 7).  w, h <= 64.  Unlike the reference (pred_inter.h:42) nothing is
 * written outside the w x h block, so neighbouring PUs may be processed in one launch. */
int HEVCASM_API hevcasm_pred_uni_batch(uint8_t *dst, ptrdiff_t stride_dst, const uint8_t *ref, ptrdiff_t stride_ref, int taps,
                                       const int16_t *pus, int n_pu, void *stream);
int HEVCASM_API hevcasm_pred_bi_batch(uint8_t *dst, ptrdiff_t stride_dst, const uint8_t *ref0, const uint8_t *ref1,
                                      ptrdiff_t stride_ref, int taps, const int16_t *pus, int n_pu, void *stream);

/* The PU lists of a whole batch of frames (a GOP) in ONE launch: each descriptor carries a trailing frame index - 7 x int16
 * {x, y, w, h, mvx, mvy, frame} resp. 9 x int16 {x, y, w, h, mvx0, mvy0, mvx1, mvy1, frame} - and frame f's planes start
 * frame_stride elements after frame f-1's.  PUs of any sizes and frames may be mixed in any order. */
int HEVCASM_API hevcasm_pred_uni_list_frames(uint8_t *dst, ptrdiff_t stride_dst, const uint8_t *ref, ptrdiff_t stride_ref, int taps,
                                             const int16_t *pus, int n_pu, ptrdiff_t frame_stride_dst, ptrdiff_t frame_stride_ref,
                                             void *stream);
int HEVCASM_API hevcasm_pred_bi_list_frames(uint8_t *dst, ptrdiff_t stride_dst, const uint8_t *ref0, const uint8_t *ref1,
                                            ptrdiff_t stride_ref, int taps, const int16_t *pus, int n_pu, ptrdiff_t frame_stride_dst,
                                            ptrdiff_t frame_stride_ref, void *stream);

/* ------------------------------------------------------------------------------------------------ transforms
 * element semantics: reference residual_decode.h:82 (hevcasm_transform; residual_decode.c:592-893) and
 * residual_decode.h:54 (hevcasm_inverse_transform_add; residual_decode.c:69-413).  log2size 2..5; trType 1 selects
 * the 4x4 DST (log2size must be 2).  Coefficients of block i are N*N contiguous int16 at coeffs + i*N*N. */
int HEVCASM_API hevcasm_transform_batch(int16_t *coeffs, const int16_t *residual, ptrdiff_t stride, int log2size, int trType,
                                        const int16_t *blk_xy, int n, void *stream);
/* coeffs[frame][by][bx][N*N] over the regular grid of floor(width/N) x floor(height/N) blocks */
int HEVCASM_API hevcasm_transform_frames(int16_t *coeffs, const int16_t *residual, ptrdiff_t stride, int width, int height,
                                         int log2size, int trType, int n_frames, ptrdiff_t frame_stride_residual, void *stream);
int HEVCASM_API hevcasm_inverse_transform_add_batch(uint8_t *dst, ptrdiff_t stride_dst, const uint8_t *pred, ptrdiff_t stride_pred,
                                                    const int16_t *coeffs, int log2size, int trType, const int16_t *blk_xy, int n,
                                                    void *stream);
/* Transform-unit lists of a batch of frames in one call (a GOP's TUs, bucketed by size class while the lists are built): tus[i] =
 * {x, y, frame}, x and y multiples of the TU size; the first n_by_class[0] entries are 4x4 DST blocks, then n_by_class[1] 4x4, [2] 8x8, [3] 16x16 and [4] 32x32 DCT
 * blocks; block i's coefficients follow block i-1's (N*N contiguous int16 each), coeffs 16-byte aligned.  One launch per class present.
 * Element semantics as hevcasm_transform_batch / hevcasm_inverse_transform_add_batch. */
int HEVCASM_API hevcasm_transform_list_frames(int16_t *coeffs, const int16_t *residual, ptrdiff_t stride, const int16_t *tus,
                                              const int *n_by_class, ptrdiff_t frame_stride_residual, void *stream);
int HEVCASM_API hevcasm_inverse_transform_add_list_frames(uint8_t *dst, ptrdiff_t stride_dst, const uint8_t *pred, ptrdiff_t stride_pred,
                                                          const int16_t *coeffs, const int16_t *tus, const int *n_by_class,
                                                          ptrdiff_t frame_stride_dst, ptrdiff_t frame_stride_pred, void *stream);
int HEVCASM_API hevcasm_inverse_transform_add_frames(uint8_t *dst, ptrdiff_t stride_dst, const uint8_t *pred, ptrdiff_t stride_pred,
                                                     const int16_t *coeffs, int width, int height, int log2size, int trType,
                                                     int n_frames, ptrdiff_t frame_stride_dst, ptrdiff_t frame_stride_pred,
                                                     void *stream);

/* ------------------------------------------------------------------------------------------------ quantisation
 * element semantics: reference quantize.h:78 / quantize.c:160-186 (quantize), quantize.h:57 / quantize.c:53-62
 * (quantize_inverse), quantize.h:99 / quantize.c:292-302 (quantize_reconstruct). */

/* n_blocks runs of n_per_block coefficients (a power of two, 16..1024); cbf[i] = OR of block i's outputs, the
 * reference's return value; cbf may be NULL.  dst, src 16-byte aligned. */
int HEVCASM_API hevcasm_quantize_batch(int16_t *dst, const int16_t *src, int scale, int shift, int offset, int n_per_block,
                                       int n_blocks, int32_t *cbf, void *stream);
int HEVCASM_API hevcasm_quantize_inverse_batch(int16_t *dst, const int16_t *src, int scale, int shift, long long n_total,
                                               void *stream);
int HEVCASM_API hevcasm_quantize_reconstruct_batch(uint8_t *rec, ptrdiff_t stride_rec, const uint8_t *pred, ptrdiff_t stride_pred,
                                                   const int16_t *res, int log2size, const int16_t *blk_xy, int n, void *stream);
/* the same over the TU lists of a batch of frames: tus[i] = {x, y, frame}, bucketed by size - n_by_size[0] 4x4 blocks, then the 8x8, 16x16 and
 * 32x32 ones; block i's residual follows block i-1's.  One launch per size present. */
int HEVCASM_API hevcasm_quantize_reconstruct_list_frames(uint8_t *rec, ptrdiff_t stride_rec, const uint8_t *pred, ptrdiff_t stride_pred,
                                                         const int16_t *res, const int16_t *tus, const int *n_by_size,
                                                         ptrdiff_t frame_stride_rec, ptrdiff_t frame_stride_pred, void *stream);
int HEVCASM_API hevcasm_quantize_reconstruct_frames(uint8_t *rec, ptrdiff_t stride_rec, const uint8_t *pred, ptrdiff_t stride_pred,
                                                    const int16_t *res, int width, int height, int log2size, int n_frames,
                                                    ptrdiff_t frame_stride_rec, ptrdiff_t frame_stride_pred, void *stream);

/* ------------------------------------------------------------------------------------------------ fused residual pipeline
 * forward transform -> quantize -> (levels written) -> quantize_inverse -> inverse transform -> add to pred,
 * one pass over HBM.  Element semantics = the composition of the four reference
 * functions above, bit for bit.  levels[frame][by][bx][N*N]; cbf[frame][by][bx] (may be NULL). */
int HEVCASM_API hevcasm_residual_pipeline_frames(uint8_t *rec, ptrdiff_t stride_rec, int16_t *levels, int32_t *cbf,
                                                 const int16_t *residual, ptrdiff_t stride_residual, const uint8_t *pred,
                                                 ptrdiff_t stride_pred, int width, int height, int log2size, int trType,
                                                 int q_scale, int q_shift, int q_offset, int iq_scale, int iq_shift, int n_frames,
                                                 ptrdiff_t frame_stride_rec, ptrdiff_t frame_stride_residual,
                                                 ptrdiff_t frame_stride_pred, void *stream);

/* The residual formed on the fly from two 8-bit planes, residual = src - pred, as the encoder-side f265_lbd_dct_8_avx2 (reference
 * f265/dct.asm:561) takes its input: the forward transform alone, and the whole fused pipeline (then the predictor is read once for the
 * residual AND the reconstruction: 5.06 instead of 6.06 bytes per sample, and no int16 residual plane is ever written).  Block sizes
 * 4x4 (DCT or DST) and 8x8, whose first stage runs directly on the bytes.  Element semantics: hevcasm_transform resp. the pipeline above
 * applied to the int16 differences. */
int HEVCASM_API hevcasm_transform_from_planes_frames(int16_t *coeffs, const uint8_t *src, ptrdiff_t stride_src, const uint8_t *pred,
                                                     ptrdiff_t stride_pred, int width, int height, int log2size, int trType, int n_frames,
                                                     ptrdiff_t frame_stride_src, ptrdiff_t frame_stride_pred, void *stream);
int HEVCASM_API hevcasm_residual_from_planes_pipeline_frames(uint8_t *rec, ptrdiff_t stride_rec, int16_t *levels, int32_t *cbf,
                                                             const uint8_t *src, ptrdiff_t stride_src, const uint8_t *pred,
                                                             ptrdiff_t stride_pred, int width, int height, int log2size, int trType,
                                                             int q_scale, int q_shift, int q_offset, int iq_scale, int iq_shift,
                                                             int n_frames, ptrdiff_t frame_stride_rec, ptrdiff_t frame_stride_src,
                                                             ptrdiff_t frame_stride_pred, void *stream);

/* ------------------------------------------------------------------------------------------------ host-memory forms
 * A context owns one CUDA stream and a device staging arena on one device.  The *_host entry points take HOST
 * pointers with the same meaning as the device forms, copy inputs in, run, copy results out and synchronise
 * before returning - this is what the per-block slots of the function-select tables and bench.py's e2e
 * measurement use. */
typedef struct hevcasm_cuda_context hevcasm_cuda_context;

hevcasm_cuda_context HEVCASM_API *hevcasm_cuda_context_create(int device, size_t arena_bytes);
void HEVCASM_API hevcasm_cuda_context_destroy(hevcasm_cuda_context *ctx);
void HEVCASM_API *hevcasm_cuda_context_stream(hevcasm_cuda_context *ctx);
/* page-locked host memory, so that the copies of the *_host forms overlap with compute.  The _near form binds the pages to the NUMA node
 * of `device` (falls back to the plain form when the platform does not say which node that is); both are released by hevcasm_cuda_host_free */
void HEVCASM_API *hevcasm_cuda_host_alloc(size_t bytes);
void HEVCASM_API *hevcasm_cuda_host_alloc_near(size_t bytes, int device);
int HEVCASM_API hevcasm_cuda_device_numa_node(int device);   /* the node the _near form binds to, -1 if the platform does not say */
void HEVCASM_API hevcasm_cuda_host_free(void *p);

/* plane-level host forms: each frame is a tightly described plane {pointer to sample (0,0), stride}; `pad` = the number
 * of valid samples around the width x height area that must travel with it (>= candidate range / filter reach). */
int HEVCASM_API hevcasm_sad_sweep_pyramid_frames_host(hevcasm_cuda_context *ctx, const uint8_t *src, ptrdiff_t stride_src,
                                                      const uint8_t *ref, ptrdiff_t stride_ref, int width, int height, int pad,
                                                      int dx0, int dy0, int n_frames, ptrdiff_t frame_stride_src,
                                                      ptrdiff_t frame_stride_ref, int32_t *sad8, int32_t *sad16, int32_t *sad32,
                                                      int32_t *sad64);
int HEVCASM_API hevcasm_sad_sweep_pyramid_best_frames_host(hevcasm_cuda_context *ctx, const uint8_t *src, ptrdiff_t stride_src,
                                                           const uint8_t *ref, ptrdiff_t stride_ref, int width, int height, int pad,
                                                           int dx0, int dy0, int n_frames, ptrdiff_t frame_stride_src,
                                                           ptrdiff_t frame_stride_ref, int32_t *best8, int32_t *best16, int32_t *best32,
                                                           int32_t *best64);
int HEVCASM_API hevcasm_sad_sweep_pyramid_packed_frames_host(hevcasm_cuda_context *ctx, const uint8_t *src, ptrdiff_t stride_src,
                                                             const uint8_t *ref, ptrdiff_t stride_ref, int width, int height, int pad,
                                                             int dx0, int dy0, int n_frames, ptrdiff_t frame_stride_src,
                                                             ptrdiff_t frame_stride_ref, uint16_t *sad8, uint16_t *sad16, int32_t *sad32,
                                                             int32_t *sad64);
int HEVCASM_API hevcasm_pred_uni_frames_host(hevcasm_cuda_context *ctx, uint8_t *dst, ptrdiff_t stride_dst, const uint8_t *ref,
                                             ptrdiff_t stride_ref, int width, int height, int pad, int taps, int xFrac, int yFrac,
                                             int n_frames, ptrdiff_t frame_stride_dst, ptrdiff_t frame_stride_ref);
int HEVCASM_API hevcasm_residual_pipeline_frames_host(hevcasm_cuda_context *ctx, uint8_t *rec, ptrdiff_t stride_rec, int16_t *levels,
                                                      int32_t *cbf, const int16_t *residual, ptrdiff_t stride_residual,
                                                      const uint8_t *pred, ptrdiff_t stride_pred, int width, int height,
                                                      int log2size, int trType, int q_scale, int q_shift, int q_offset,
                                                      int iq_scale, int iq_shift, int n_frames, ptrdiff_t frame_stride_rec,
                                                      ptrdiff_t frame_stride_residual, ptrdiff_t frame_stride_pred);

#ifdef __cplusplus
}
#endif

#endif
