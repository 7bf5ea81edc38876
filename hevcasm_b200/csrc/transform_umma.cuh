// hevcasm_b200 - inverse 16x16 / 32x32 transform + add on the 5th-generation tensor cores (tcgen05.mma kind::i8, TMEM).
//
// The north star allows tensor-core transforms only as an exact int32-accumulating formulation that beats the CUDA-core
// butterfly in ncu.  The legacy mma.sync route (transform_imma.cuh) is exact but slower than the butterfly; this is the
// tcgen05 route.  profiles/r01_transforms.md has the comparison.
//
// Exactness: as in transform_imma.cuh.  A stage is  out = clip16((sum_k T[k][.] * x[k] + round) >> shift), |T| <= 90 fits
// s8 and x = 256*hi + lo with hi = x >> 8 (s8), lo = x & 255 (u8), so  sum T*x = 256 * (sum T*hi) + (sum T*lo): two
// integer matrix products (s8 x s8 and u8 x s8) accumulated in int32 - the same integer the reference's C computes
// (residual_decode.c:69-347); rounding, shift and clip are then applied exactly as there.
//
// One CTA (128 threads) works on GROUPS of 128/N blocks, so that a stage of the whole group is ONE M = 128 product:
//   stage 1 (contracts the vertical frequency v):  D1[(b,u)][y] = sum_v C_b[v][u] * T[v][y]      A = C^T, B = T
//   stage 2 (contracts the horizontal one u):      D2[(b,y)][x] = sum_u tmp_b[u][y] * T[u][x]    A = tmp^T, B = T
// In both stages the data operand A[m][k] is contiguous along m in its source (a coefficient row C_b[v][.], a thread's
// own row of stage-1 results tmp_b[u][.]), which is exactly the tensor core's MN-major operand form: 16-byte chunks of
// 16 consecutive m at one k, 8 k per 128-byte core matrix.  So the only data rearrangement is the lo/hi byte split
// (2 PRMT per 4 values) on the way into shared memory - no transposes anywhere.  B = T[k][n] sits in shared memory once
// per CTA (K-major, zero rows beyond N).  Accumulators live in TMEM: thread t reads row m = t (tcgen05.ld 32x32b) - in
// stage 1 that is row u of block b, whose N values over y are again 16-byte chunks of the stage-2 operand; in stage 2 it
// is one row of N output samples, to which the thread adds its predictor row and which it stores as N contiguous bytes.
// The matrix-descriptor fields were pinned with tools/umma_probe.cu before this kernel was written.
// (included by transform.cu inside namespace hv, after BlockGrid / load_words / store_words; the tcgen05 wrappers are in umma.cuh)
#pragma once


constexpr int UMMA_NT = 128;

// 16 int16 values (as 8 packed pairs) -> their 16 low bytes and 16 high bytes
__device__ __forceinline__ void split_bytes(const uint32_t (&p)[8], uint4 &lo, uint4 &hi)
{
    lo = make_uint4(__byte_perm(p[0], p[1], 0x6420), __byte_perm(p[2], p[3], 0x6420), __byte_perm(p[4], p[5], 0x6420), __byte_perm(p[6], p[7], 0x6420));
    hi = make_uint4(__byte_perm(p[0], p[1], 0x7531), __byte_perm(p[2], p[3], 0x7531), __byte_perm(p[4], p[5], 0x7531), __byte_perm(p[6], p[7], 0x7531));
}

template <int LOG2, bool PA>
__global__ void __launch_bounds__(UMMA_NT, 6) umma_inv_kernel(uint8_t *__restrict__ dst, ptrdiff_t sd, const uint8_t *__restrict__ pred, ptrdiff_t sp,
                                                              ptrdiff_t fs_dst, ptrdiff_t fs_pred, const int16_t *__restrict__ coeffs, BlockGrid grid)
{
    constexpr int N = 1 << LOG2, BPG = 128 / N;        // blocks per group
    constexpr int KG = 128, SBO = (N / 8) * KG;         // A: 8 k per 128-byte core matrix; chunks of 16 m are SBO apart
    constexpr int A_BYTES = 128 * N;                    // one byte plane of the operand (lo or hi)
    constexpr int CH = N / 16;                          // 16-value chunks per row
    constexpr int TCOLS = 2 * N;                        // TMEM columns: D_lo | D_hi
    // (for N = 16 the MMA still contracts 32 k: the k groups 2, 3 it reads lie in the next chunk / the pad and meet zero rows of B)
    __shared__ __align__(128) uint8_t sA[2 * A_BYTES + 512];
    __shared__ __align__(128) uint8_t sB[N * 32];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < N * 32; i += UMMA_NT) {      // B[n][k] = T[k][n], K-major: chunks of 16 k are 128 bytes apart, groups of 8 n 256
        const int n = i >> 5, k = i & 31;
        sB[(n >> 3) * 256 + (k >> 4) * 128 + (n & 7) * 16 + (k & 15)] = k < N ? (uint8_t)(int8_t)dct(N, k, n) : 0;
    }
    if (tid < 128) *reinterpret_cast<uint32_t *>(sA + 2 * A_BYTES + 4 * tid) = 0;
    if (tid == 0) tma::mbar_init(&bar, 1);
    if (warp == 0) umma::tmem_alloc<TCOLS>(&tmem_slot);
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tm = tmem_slot, tlane = tm + ((uint32_t)(warp * 32) << 16);
    const uint64_t desc_lo = umma::smem_desc(tma::smem_u32(sA), KG, SBO), desc_hi = umma::smem_desc(tma::smem_u32(sA + A_BYTES), KG, SBO);
    const uint64_t desc_b = umma::smem_desc(tma::smem_u32(sB), 128, 256);
    constexpr uint32_t ID_LO = umma::idesc_i8(false, true, true, N), ID_HI = umma::idesc_i8(true, true, true, N);

    const long long n_groups = (grid.n + BPG - 1) / BPG;
    uint32_t phase = 0;
    // coefficient chunks of this thread: 16 values (32 bytes) each; a group has 8 N of them, chunk c = tid (+ 128)
    constexpr int CPT = N / 16;
    int4 cw[CPT][2];
    auto load_coefs = [&](long long g) {
#pragma unroll
        for (int h = 0; h < CPT; ++h) {
            const int c = tid + 128 * h;
            const long long blk = g * BPG + (c * 16) / (N * N);
            if (blk < grid.n) {
                const int4 *src = reinterpret_cast<const int4 *>(coeffs + g * (128 * N) + c * 16);
                cw[h][0] = ldg_stream(src), cw[h][1] = ldg_stream(src + 1);
            } else {
                cw[h][0] = cw[h][1] = make_int4(0, 0, 0, 0);
            }
        }
    };
    if ((long long)blockIdx.x < n_groups) load_coefs(blockIdx.x);

    // this thread's block: index blk = g * BPG + tid / N advances by gridDim.x * BPG per iteration; on a regular grid its
    // (column, row, frame) position is carried along instead of being re-derived with 64-bit divisions every group
    int bx = 0, by = 0, bf = 0, step_x = 0, step_y = 0;
    if (!grid.blk_xy) {
        const long long first = (long long)blockIdx.x * BPG + tid / N, per = (long long)grid.nbx * grid.nby, step = (long long)gridDim.x * BPG;
        bf = (int)(first / per);
        const int r = (int)(first - bf * per);
        by = r / grid.nbx, bx = r - by * grid.nbx;
        step_y = (int)(step / grid.nbx), step_x = (int)(step - (long long)step_y * grid.nbx);
    }

#pragma unroll 1
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        // ---- this thread's output row: block b, row y = tid % N; its predictor row is requested now and used in the epilogue
        const long long blk = g * BPG + tid / N;
        const bool live = blk < grid.n;
        int x0 = 0, y0 = 0, f = 0;
        if (grid.blk_xy) {
            if (live) x0 = grid.blk_xy[grid.desc_w * blk], y0 = grid.blk_xy[grid.desc_w * blk + 1], f = grid.desc_w == 3 ? grid.blk_xy[grid.desc_w * blk + 2] : 0;
        } else {
            x0 = bx << LOG2, y0 = by << LOG2, f = bf;
            bx += step_x, by += step_y;
            if (bx >= grid.nbx) bx -= grid.nbx, ++by;
            while (by >= grid.nby) by -= grid.nby, ++bf;
        }
        uint32_t pw[N / 4];
        const ptrdiff_t row_off = y0 + (tid & (N - 1));
        if (live) load_words<N / 4, PA>(pred + f * fs_pred + row_off * sp + x0, pw);

        // ---- coefficients -> lo / hi byte planes of the stage-1 operand: chunk (v, u0..u0+15) of block b goes to m-chunk (b*N + u0)/16, k = v
#pragma unroll
        for (int h = 0; h < CPT; ++h) {
            const int c = tid + 128 * h, s = c * 16, b = s / (N * N), v = (s / N) & (N - 1), u0 = s & (N - 1);
            const uint32_t p[8] = {(uint32_t)cw[h][0].x, (uint32_t)cw[h][0].y, (uint32_t)cw[h][0].z, (uint32_t)cw[h][0].w,
                                   (uint32_t)cw[h][1].x, (uint32_t)cw[h][1].y, (uint32_t)cw[h][1].z, (uint32_t)cw[h][1].w};
            uint4 lo, hi;
            split_bytes(p, lo, hi);
            const int off = ((b * N + u0) >> 4) * SBO + (v >> 3) * KG + (v & 7) * 16;
            *reinterpret_cast<uint4 *>(sA + off) = lo;
            *reinterpret_cast<uint4 *>(sA + A_BYTES + off) = hi;
        }
        umma::fence_async_smem();
        umma::fence_before();   // (this thread's TMEM reads of the previous group are complete: tmem_ld_wait below)
        __syncthreads();
        if (tid == 0) {
            umma::fence_after();
            umma::mma_i8(tm, desc_lo, desc_b, ID_LO);
            umma::mma_i8(tm + N, desc_hi, desc_b, ID_HI);
            umma::commit(&bar);
        }
        if (g + gridDim.x < n_groups) load_coefs(g + gridDim.x);   // next group's coefficients travel during both stages

        // ---- stage 1 epilogue: tmp[u][y] = clip16((D + 64) >> 7), re-split into the stage-2 operand: chunk (u, y0..y0+15) -> m-chunk (b*N + y0)/16, k = u
        tma::mbar_wait(&bar, phase);
        phase ^= 1;
        umma::fence_after();
        {
            const int b = tid / N, u = tid & (N - 1);
            // (the stage-1 MMAs have completed - their commit was observed - so the operand planes may be overwritten at once;
            //  the accumulators other threads still have to read are in TMEM, not in shared memory)
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
                int dl[16], dh[16];
                umma::tmem_ld16(tlane + 16 * ch, dl);
                umma::tmem_ld16(tlane + N + 16 * ch, dh);
                umma::tmem_ld_wait();
                uint32_t p[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int a0 = ((dl[2 * i] + 64) >> 7) + 2 * dh[2 * i], a1 = ((dl[2 * i + 1] + 64) >> 7) + 2 * dh[2 * i + 1];   // 256*hi is a multiple of 128
                    p[i] = pack_sat_s16(a0, a1);
                }
                uint4 lo4, hi4;
                split_bytes(p, lo4, hi4);
                const int off = ((b * N + 16 * ch) >> 4) * SBO + (u >> 3) * KG + (u & 7) * 16;
                *reinterpret_cast<uint4 *>(sA + off) = lo4;
                *reinterpret_cast<uint4 *>(sA + A_BYTES + off) = hi4;
            }
            umma::fence_before();
        }
        umma::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after();
            umma::mma_i8(tm, desc_lo, desc_b, ID_LO);
            umma::mma_i8(tm + N, desc_hi, desc_b, ID_HI);
            umma::commit(&bar);
        }

        // ---- stage 2 epilogue: row y of block b: clip8(pred + ((D + 2048) >> 12)), N contiguous bytes
        tma::mbar_wait(&bar, phase);
        phase ^= 1;
        umma::fence_after();
        uint32_t ow[N / 4];
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
            int dl[16], dh[16];
            umma::tmem_ld16(tlane + 16 * ch, dl);
            umma::tmem_ld16(tlane + N + 16 * ch, dh);
            umma::tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int r[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = 4 * q + j;
                    // pred + ((D + 2048) >> 12) = (lo + 256 * (hi + 16 * pred) + 2048) >> 12: the predictor byte enters through one IDP.4A
                    const int h2 = (int)dp4a_uu(pw[4 * ch + q], 16u << (8 * j), (uint32_t)dh[i]);
                    r[j] = (dl[i] + (h2 << 8) + 2048) >> 12;
                }
                ow[4 * ch + q] = pack_sat_u8(r[0], r[1], r[2], r[3]);
            }
        }
        if (live) store_words<N / 4, PA>(dst + f * fs_dst + row_off * sd + x0, ow);
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<TCOLS>(tm);
}

