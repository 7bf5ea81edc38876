"""GPU parity at the full sizes of BASELINE.json's configs, every output compared bit for bit with the CPU oracle
(the oracle runs multi-threaded; a 4K plane takes it well under a second per kernel)."""
import os

import numpy as np
import pytest

from hevcasm_b200 import lib, shard, synth
from hevcasm_b200.abi import HEVCASM_RECT
from oracle.binding import ptr
from gpu_util import to_dev, dev_full, dptr, to_host

pytestmark = pytest.mark.gpu
T = os.cpu_count() or 4


def test_config0_1080p_sad_ssd_dct8(oracle):
    """configs[0]: SAD / SSD over every aligned 8..64 block vs the co-located block, 8x8 forward DCT over every block, 1920x1080"""
    W, H = 1920, 1080
    a = synth.smooth_planes(700, 1, W, H, 32)
    b = synth.smooth_planes(700, 1, W, H, 32, shift=(1, 1), noise=6)
    da, db = to_dev(a.buf), to_dev(b.buf)
    for s, log2 in ((8, 3), (16, 4), (32, 5), (64, 6)):
        n = (W // s) * (H // s)
        want = np.zeros(n, np.int32)
        oracle.drv("sad_sweep_frames", ptr(a.buf, a.origin), a.pitch, ptr(b.buf, b.origin), b.pitch, W, H, HEVCASM_RECT(s, s), 0, 0, 1, 1, 1, a.frame_stride,
                   b.frame_stride, ptr(want), threads=T)
        got = dev_full((n,), np.int32, -1)
        lib.call("sad_sweep_frames", dptr(da, a.origin), a.pitch, dptr(db, b.origin), b.pitch, W, H, HEVCASM_RECT(s, s), 0, 0, 1, 1, 1, a.frame_stride, b.frame_stride,
                 dptr(got))
        assert np.array_equal(to_host(got), want), ("sad", s)
        oracle.drv("ssd_frames", ptr(a.buf, a.origin), a.pitch, ptr(b.buf, b.origin), b.pitch, W, H, log2, 1, a.frame_stride, b.frame_stride, ptr(want), threads=T)
        lib.call("ssd_frames", dptr(da, a.origin), a.pitch, dptr(db, b.origin), b.pitch, W, H, log2, 1, a.frame_stride, b.frame_stride, dptr(got))
        assert np.array_equal(to_host(got), want), ("ssd", s)
    res = synth.residual_planes(701, 1, W, H)
    want = np.zeros((W // 8) * (H // 8) * 64, np.int16)
    oracle.drv("transform_frames", ptr(want), ptr(res.buf, res.origin), res.pitch, W, H, 3, 0, 1, res.frame_stride, threads=T)
    got = dev_full(want.shape, np.int16, 0x5a5a)
    dres = to_dev(res.buf)
    lib.call("transform_frames", dptr(got), dptr(dres, res.origin), res.pitch, W, H, 3, 0, 1, res.frame_stride)
    assert np.array_equal(to_host(got), want)


def test_config1_4k_sad_sweep(oracle):
    """configs[1]: 4K, 8x8..64x64 PUs x 64 candidates: the pyramid kernel and its argmin form against four oracle sweeps"""
    W, H = 3840, 2160
    src = synth.smooth_planes(710, 1, W, H, 32)
    ref = synth.smooth_planes(710, 1, W, H, 32, shift=(-2, 3), noise=5)
    ds, dr = to_dev(src.buf), to_dev(ref.buf)
    outs = [dev_full(((W // s) * (H // s) * 64,), np.int32, -1) for s in (8, 16, 32, 64)]
    best = [dev_full(((W // s) * (H // s), 2), np.int32, -1) for s in (8, 16, 32, 64)]
    lib.call("sad_sweep_pyramid_frames", dptr(ds, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, W, H, -4, -4, 1, src.frame_stride, ref.frame_stride,
             *[dptr(o) for o in outs])
    lib.call("sad_sweep_pyramid_best_frames", dptr(ds, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, W, H, -4, -4, 1, src.frame_stride, ref.frame_stride,
             *[dptr(o) for o in best])
    for s, o, bo in zip((8, 16, 32, 64), outs, best):
        want = np.zeros(((W // s) * (H // s), 64), np.int32)
        oracle.drv("sad_sweep_frames", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, W, H, HEVCASM_RECT(s, s), -4, -4, 8, 8, 1,
                   src.frame_stride, ref.frame_stride, ptr(want), threads=T)
        assert np.array_equal(to_host(o).reshape(want.shape), want), s
        b = to_host(bo)
        assert np.array_equal(b[:, 0], want.min(-1)) and np.array_equal(b[:, 1], want.argmin(-1)), s


@pytest.mark.parametrize("taps,W,H", [(8, 3840, 2160), (4, 1920, 1080)])
def test_config2_interpolation_all_positions(oracle, taps, W, H):
    """configs[2]: 4K luma at all 16 and 1080p chroma at all 64 fractional positions; bi at a sweep of position tuples"""
    nfrac = 4 if taps == 8 else 8
    ref = synth.random_planes(720 + taps, 1, W, H, 16)
    ref1 = synth.smooth_planes(721 + taps, 1, W, H, 16)
    dr, dr1 = to_dev(ref.buf), to_dev(ref1.buf)
    want = synth.Planes(np.zeros_like(ref.buf), W, H, 16)
    got = to_dev(want.buf)
    for yf in range(nfrac):
        for xf in range(nfrac):
            oracle.drv("pred_uni_frames", ptr(want.buf, want.origin), want.pitch, ptr(ref.buf, ref.origin), ref.pitch, W, H, taps, xf, yf, 1, want.frame_stride,
                       ref.frame_stride, threads=T)
            lib.call("pred_uni_frames", dptr(got, want.origin), want.pitch, dptr(dr, ref.origin), ref.pitch, W, H, taps, xf, yf, 1, want.frame_stride, ref.frame_stride)
            assert np.array_equal(to_host(got), want.buf), (xf, yf)
    for c in [(f, f, f, f) for f in range(nfrac)] + [(1, 0, 0, nfrac - 1), (0, 0, 0, 0)]:
        oracle.drv("pred_bi_frames", ptr(want.buf, want.origin), want.pitch, ptr(ref.buf, ref.origin), ptr(ref1.buf, ref1.origin), ref.pitch, W, H, taps, *c, 1,
                   want.frame_stride, ref.frame_stride, threads=T)
        lib.call("pred_bi_frames", dptr(got, want.origin), want.pitch, dptr(dr, ref.origin), dptr(dr1, ref1.origin), ref.pitch, W, H, taps, *c, 1,
                 want.frame_stride, ref.frame_stride)
        assert np.array_equal(to_host(got), want.buf), c


def test_config3_4k_residual_pipeline(oracle):
    """configs[3]: 4K residual -> 8x8 DCT -> quant -> dequant -> inverse + add, stage by stage and fused; inverse + add alone at 4..32"""
    from test_gpu_pipeline import oracle_pipeline
    W, H = 3840, 2160
    qp = (26214, 18, 171 << 7, 18432, 6)
    res = synth.residual_planes(730, 1, W, H)
    pred = synth.random_planes(731, 1, W, H, 16)
    lv_w, cbf_w, rec_w = oracle_pipeline(oracle, res, pred, W, H, 3, 0, qp, 1)
    dres, dpred = to_dev(res.buf), to_dev(pred.buf)
    rec = to_dev(synth.random_planes(301, 1, W, H, 16).buf)
    lv, cbf = dev_full(lv_w.shape, np.int16, 1), dev_full(cbf_w.shape, np.int32, 1)
    lib.call("residual_pipeline_frames", dptr(rec, rec_w.origin), rec_w.pitch, dptr(lv), dptr(cbf), dptr(dres, res.origin), res.pitch, dptr(dpred, pred.origin),
             pred.pitch, W, H, 3, 0, *qp, 1, rec_w.frame_stride, res.frame_stride, pred.frame_stride)
    assert np.array_equal(to_host(lv), lv_w) and np.array_equal(to_host(cbf), cbf_w) and np.array_equal(to_host(rec), rec_w.buf)
    for log2 in (2, 3, 4, 5):
        n = 1 << log2
        co = synth.random_int16(732 + log2, (W // n) * (H // n) * n * n, -2000, 2000)
        want = synth.Planes(np.zeros_like(pred.buf), W, H, 16)
        oracle.drv("inverse_transform_add_frames", ptr(want.buf, want.origin), want.pitch, ptr(pred.buf, pred.origin), pred.pitch, ptr(co), W, H, log2, 0, 1,
                   want.frame_stride, pred.frame_stride, threads=T)
        got, dc = to_dev(np.zeros_like(pred.buf)), to_dev(co)
        lib.call("inverse_transform_add_frames", dptr(got, want.origin), want.pitch, dptr(dpred, pred.origin), pred.pitch, dptr(dc), W, H, log2, 0, 1, want.frame_stride,
                 pred.frame_stride)
        assert np.array_equal(to_host(got), want.buf), log2


def test_config4_8k_64_frames_sharded(oracle):
    """configs[4] at its full frame count: 64 8K frames dealt to ranks by hevcasm_b200.shard.  Every frame goes through the SAD sweep (argmin
    form), one two-pass luma interpolation and the 8x8 forward DCT, in batched calls of 8 frames, and EVERY frame's outputs are compared with
    the oracle's by digest; the per-frame digests of a 2-rank deal equal those of the 1-rank run (no result depends on which rank, or
    which position in a batch, a frame had)."""
    W, H, NF, CH = 7680, 4320, 64, 8

    def frames(fs):   # every frame is generated from its own seed, as a rank would
        src = synth.Planes(np.concatenate([synth.random_planes(740 + f, 1, W, H, 16).buf for f in fs]), W, H, 16)
        ref = synth.Planes(np.concatenate([synth.random_planes(860 + f, 1, W, H, 16).buf for f in fs]), W, H, 16)
        res = synth.Planes(np.concatenate([synth.residual_planes(980 + f, 1, W, H).buf for f in fs]), W, H, 0)
        return src, ref, res

    def gpu_digests(src, ref, res):
        n = src.n_frames
        ds, dr, dres = to_dev(src.buf), to_dev(ref.buf), to_dev(res.buf)
        best = [dev_full((n, (W // s) * (H // s), 2), np.int32, -1) for s in (8, 16, 32, 64)]
        lib.call("sad_sweep_pyramid_best_frames", dptr(ds, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, W, H, -4, -4, n, src.frame_stride,
                 ref.frame_stride, *[dptr(o) for o in best])
        pr = to_dev(np.zeros_like(src.buf))
        lib.call("pred_uni_frames", dptr(pr, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, W, H, 8, 2, 1, n, src.frame_stride, ref.frame_stride)
        co = dev_full((n, (W // 8) * (H // 8) * 64), np.int16, 0)
        lib.call("transform_frames", dptr(co), dptr(dres, res.origin), res.pitch, W, H, 3, 0, n, res.frame_stride)
        b, p, c = [to_host(o) for o in best], to_host(pr), to_host(co)
        return [shard.frame_digest(*[x[k] for x in b], p[k], c[k]) for k in range(n)]

    def oracle_digests(src, ref, res):
        n = src.n_frames
        b = []
        for s in (8, 16, 32, 64):
            sad = np.zeros((n, (W // s) * (H // s), 64), np.int32)
            oracle.drv("sad_sweep_frames", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, W, H, HEVCASM_RECT(s, s), -4, -4, 8, 8, n,
                       src.frame_stride, ref.frame_stride, ptr(sad), threads=T)
            b.append(np.stack([sad.min(-1), sad.argmin(-1).astype(np.int32)], -1).astype(np.int32))
        wp = np.zeros_like(src.buf)
        oracle.drv("pred_uni_frames", ptr(wp, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, W, H, 8, 2, 1, n, src.frame_stride, ref.frame_stride, threads=T)
        co = np.zeros((n, (W // 8) * (H // 8) * 64), np.int16)
        oracle.drv("transform_frames", ptr(co), ptr(res.buf, res.origin), res.pitch, W, H, 3, 0, n, res.frame_stride, threads=T)
        return [shard.frame_digest(*[x[k] for x in b], wp[k], co[k]) for k in range(n)]

    want = [None] * NF
    digests = {}
    for world in (1, 2):
        per_frame = [None] * NF
        for rank in range(world):
            f0, f1 = shard.frame_range(NF, rank, world)
            first = f0 + (3 if world == 2 else 0)     # the second deal also cuts its batches at different frames
            cuts = [f0] + list(range(first, f1, CH)) + [f1]
            for a, b_ in zip(cuts[:-1], cuts[1:]):
                if a == b_:
                    continue
                src, ref, res = frames(range(a, b_))
                per_frame[a:b_] = gpu_digests(src, ref, res)
                if world == 1:
                    want[a:b_] = oracle_digests(src, ref, res)
        digests[world] = shard.gather_frame_digests(per_frame, 0, NF)
    assert digests[1] == want, [i for i, (g, w) in enumerate(zip(digests[1], want)) if g != w]
    assert digests[1] == digests[2] and len(set(digests[1])) == NF


def test_default_dispatch_reaches_the_tensor_core_kernels_at_full_size(oracle):
    """The sizes bench.py runs, through the DEFAULT dispatch of the product library (no switches exist there): forward 16x16 / 32x32 over
    16 4K frames (>= 6 tiles per SM -> ft::fwd_umma_kernel), two-reference luma interpolation over 3 4K frames (>= 5 tiles per SM ->
    uv::pred_vh_kernel<8, true>) and two-reference chroma over 16 1080p planes, every output sample compared with the oracle"""
    W, H, NF = 3840, 2160, 16
    res = synth.residual_planes(990, NF, W, H)
    dres = to_dev(res.buf)
    for log2 in (4, 5):
        n = 1 << log2
        nb = (W // n) * (H // n) * NF
        want = np.zeros(nb * n * n, np.int16)
        oracle.drv("transform_frames", ptr(want), ptr(res.buf, res.origin), res.pitch, W, H, log2, 0, NF, res.frame_stride, threads=T)
        got = dev_full(want.shape, np.int16, 0x5a5a)
        lib.call("transform_frames", dptr(got), dptr(dres, res.origin), res.pitch, W, H, log2, 0, NF, res.frame_stride)
        assert np.array_equal(to_host(got), want), log2
    del dres
    for taps, w, h, nf, fr in ((8, 3840, 2160, 3, (1, 2, 3, 1)), (8, 3840, 2160, 3, (2, 0, 0, 3)), (4, 1920, 1080, 16, (3, 5, 6, 1))):
        r0 = synth.random_planes(991 + taps, nf, w, h, 16)
        r1 = synth.smooth_planes(992 + taps, nf, w, h, 16)
        want = synth.Planes(np.zeros_like(r0.buf), w, h, 16)
        oracle.drv("pred_bi_frames", ptr(want.buf, want.origin), want.pitch, ptr(r0.buf, r0.origin), ptr(r1.buf, r1.origin), r0.pitch, w, h, taps, *fr, nf,
                   want.frame_stride, r0.frame_stride, threads=T)
        got, d0, d1 = to_dev(np.zeros_like(r0.buf)), to_dev(r0.buf), to_dev(r1.buf)
        lib.call("pred_bi_frames", dptr(got, want.origin), want.pitch, dptr(d0, r0.origin), dptr(d1, r1.origin), r0.pitch, w, h, taps, *fr, nf, want.frame_stride,
                 r0.frame_stride)
        assert np.array_equal(to_host(got), want.buf), (taps, fr)
