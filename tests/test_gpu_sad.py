"""GPU parity: SAD / SSD entry points of libhevcasm_b200.so against the CPU oracle on the same seeded inputs.
Bit-exact (integer sums).  Everything goes through the C ABI with raw device pointers."""
import numpy as np
import pytest

from hevcasm_b200 import lib, synth
from hevcasm_b200.abi import HEVCASM_RECT
from oracle.binding import ptr
from gpu_util import to_dev, dev_full, dptr, to_host
from test_oracle_vs_reference import PARTITIONS

pytestmark = pytest.mark.gpu


def _frames(n_frames, width, height, pad=24):
    src = synth.smooth_planes(synth.SEED, n_frames, width, height, pad, shift=(0, 0))
    ref = synth.smooth_planes(synth.SEED, n_frames, width, height, pad, shift=(2, -1), noise=5)
    return src, ref


@pytest.mark.parametrize("w,h", PARTITIONS)
def test_sweep_all_partitions(oracle, w, h):
    """the 23 partitions of reference sad.c:231-240, 64 candidates each, on a ragged plane"""
    width, height, nf = 200, 136, 2
    src, ref = _frames(nf, width, height)
    npu = (width // w) * (height // h)
    want = np.zeros((nf, npu, 64), np.int32)
    oracle.drv("sad_sweep_frames", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, width, height,
               HEVCASM_RECT(w, h), -4, -4, 8, 8, nf, src.frame_stride, ref.frame_stride, ptr(want), threads=8)
    ds, dr = to_dev(src.buf), to_dev(ref.buf)
    got = dev_full((nf, npu, 64), np.int32, -1)
    lib.call("sad_sweep_frames", dptr(ds, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, width, height, HEVCASM_RECT(w, h), -4, -4,
             8, 8, nf, src.frame_stride, ref.frame_stride, dptr(got))
    assert np.array_equal(to_host(got), want)


@pytest.mark.parametrize("w,h", [p for p in PARTITIONS if p[0] in (8, 16, 32, 64) and p[1] in (8, 16, 32, 64)])
def test_sweep_tma_partitions(oracle, w, h):
    """the 14 reference partitions with sides 8/16/32/64 on 16-byte aligned planes: the TMA-staged single-size kernel"""
    width, height, nf = 328, 200, 2
    src, ref = _frames(nf, width, height, pad=32)
    npu = (width // w) * (height // h)
    want = np.zeros((nf, npu, 64), np.int32)
    oracle.drv("sad_sweep_frames", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, width, height,
               HEVCASM_RECT(w, h), -4, -4, 8, 8, nf, src.frame_stride, ref.frame_stride, ptr(want), threads=8)
    ds, dr = to_dev(src.buf), to_dev(ref.buf)
    got = dev_full((nf, npu, 64), np.int32, -1)
    lib.call("sad_sweep_frames", dptr(ds, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, width, height, HEVCASM_RECT(w, h), -4, -4,
             8, 8, nf, src.frame_stride, ref.frame_stride, dptr(got))
    assert np.array_equal(to_host(got), want)


@pytest.mark.parametrize("win", [(-3, -2, 5, 3), (-8, -8, 16, 16), (1, 0, 9, 1), (-7, 3, 13, 11)])
def test_sweep_windows(oracle, win):
    dx0, dy0, ncx, ncy = win
    width, height, nf = 144, 80, 1
    src, ref = _frames(nf, width, height, pad=32)
    for w, h in ((16, 16), (8, 4), (24, 32)):
        npu = (width // w) * (height // h)
        want = np.zeros((nf, npu, ncx * ncy), np.int32)
        oracle.drv("sad_sweep_frames", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, width, height,
                   HEVCASM_RECT(w, h), dx0, dy0, ncx, ncy, nf, src.frame_stride, ref.frame_stride, ptr(want), threads=8)
        ds, dr = to_dev(src.buf), to_dev(ref.buf)
        got = dev_full(want.shape, np.int32, -1)
        lib.call("sad_sweep_frames", dptr(ds, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, width, height, HEVCASM_RECT(w, h),
                 dx0, dy0, ncx, ncy, nf, src.frame_stride, ref.frame_stride, dptr(got))
        assert np.array_equal(to_host(got), want), (w, h)


@pytest.mark.parametrize("width,height,dx0,dy0", [(256, 128, -4, -4), (200, 136, -4, -4), (136, 72, -3, 1), (64, 64, 0, 0), (8, 8, -1, -2)])
def test_pyramid(oracle, width, height, dx0, dy0):
    nf = 2
    src, ref = _frames(nf, width, height)
    ds, dr = to_dev(src.buf), to_dev(ref.buf)
    outs, wants = [], []
    for s in (8, 16, 32, 64):
        npu = (width // s) * (height // s)
        want = np.zeros((nf, npu, 64), np.int32)
        if npu:
            oracle.drv("sad_sweep_frames", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, width, height,
                       HEVCASM_RECT(s, s), dx0, dy0, 8, 8, nf, src.frame_stride, ref.frame_stride, ptr(want), threads=8)
        wants.append(want)
        outs.append(dev_full((nf, max(npu, 1), 64), np.int32, -1))
    lib.call("sad_sweep_pyramid_frames", dptr(ds, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, width, height, dx0, dy0, nf,
             src.frame_stride, ref.frame_stride, *[dptr(o) for o in outs])
    for s, o, want in zip((8, 16, 32, 64), outs, wants):
        if want.shape[1]:
            assert np.array_equal(to_host(o)[:, :want.shape[1]], want), s


def test_pyramid_null_outputs_and_uniform_random(oracle):
    width, height, nf = 128, 64, 1
    src = synth.random_planes(11, nf, width, height, 16)
    ref = synth.random_planes(12, nf, width, height, 16)
    ds, dr = to_dev(src.buf), to_dev(ref.buf)
    want = np.zeros((nf, 2, 64), np.int32)
    oracle.drv("sad_sweep_frames", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, width, height,
               HEVCASM_RECT(64, 64), -4, -4, 8, 8, nf, src.frame_stride, ref.frame_stride, ptr(want), threads=4)
    got = dev_full(want.shape, np.int32, -1)
    lib.call("sad_sweep_pyramid_frames", dptr(ds, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, width, height, -4, -4, nf,
             src.frame_stride, ref.frame_stride, None, None, None, dptr(got))
    assert np.array_equal(to_host(got), want)


def test_sad_extremes(oracle):
    """all-0 vs all-255: the maximum SAD 64*64*255 and SSD 64*64*255^2"""
    width = height = 64
    z = synth.Planes(np.zeros((1, 96, 256), np.uint8), width, height, 16)
    o = synth.Planes(np.full((1, 96, 256), 255, np.uint8), width, height, 16)
    dz, do = to_dev(z.buf), to_dev(o.buf)
    got = dev_full((1, 1, 64), np.int32, -1)
    lib.call("sad_sweep_frames", dptr(dz, z.origin), z.pitch, dptr(do, o.origin), o.pitch, width, height, HEVCASM_RECT(64, 64), -4, -4, 8, 8,
             1, z.frame_stride, o.frame_stride, dptr(got))
    assert np.all(to_host(got) == 64 * 64 * 255)
    g2 = dev_full((1,), np.int32, -1)
    lib.call("ssd_frames", dptr(dz, z.origin), z.pitch, dptr(do, o.origin), o.pitch, width, height, 6, 1, z.frame_stride, o.frame_stride, dptr(g2))
    assert int(to_host(g2)[0]) == 64 * 64 * 255 * 255


def test_list_forms(oracle):
    width, height = 192, 128
    src, ref = _frames(1, width, height, pad=40)
    ds, dr = to_dev(src.buf), to_dev(ref.buf)
    rng = synth.splitmix64(77, 4096)
    for (w, h) in ((16, 16), (64, 48), (4, 8), (12, 16)):
        n_pu, n_cand = 37, 19
        xy = np.stack([(rng[:n_pu] % np.uint64(width - w)).astype(np.int16), (rng[100:100 + n_pu] % np.uint64(height - h)).astype(np.int16)], -1)
        cand = np.stack([(rng[200:200 + n_cand] % np.uint64(65)).astype(np.int16) - 32, (rng[300:300 + n_cand] % np.uint64(65)).astype(np.int16) - 32], -1)
        xy, cand = np.ascontiguousarray(xy), np.ascontiguousarray(cand)
        want = np.zeros((n_pu, n_cand), np.int32)
        oracle.drv("sad_multiref_batch", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, HEVCASM_RECT(w, h),
                   ptr(xy), n_pu, ptr(cand), n_cand, ptr(want))
        got = dev_full(want.shape, np.int32, -1)
        dxy, dc = to_dev(xy), to_dev(cand)
        lib.call("sad_multiref_batch", dptr(ds, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, HEVCASM_RECT(w, h), dptr(dxy), n_pu,
                 dptr(dc), n_cand, dptr(got))
        assert np.array_equal(to_host(got), want)
        mv = np.ascontiguousarray(cand[np.arange(n_pu) % n_cand])
        want1 = np.zeros(n_pu, np.int32)
        oracle.drv("sad_batch", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, HEVCASM_RECT(w, h), ptr(xy), ptr(mv),
                   n_pu, ptr(want1))
        got1 = dev_full(want1.shape, np.int32, -1)
        dmv = to_dev(mv)
        lib.call("sad_batch", dptr(ds, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, HEVCASM_RECT(w, h), dptr(dxy), dptr(dmv), n_pu,
                 dptr(got1))
        assert np.array_equal(to_host(got1), want1)
    # empty batch is a no-op
    lib.call("sad_batch", dptr(ds), src.pitch, dptr(dr), ref.pitch, HEVCASM_RECT(8, 8), None, None, 0, None)


@pytest.mark.parametrize("log2", [2, 3, 4, 5, 6])
@pytest.mark.parametrize("shape", [(256, 64, 16), (200, 136, 24), (72, 40, 19)])
def test_ssd(oracle, log2, shape):
    width, height, pad = shape
    nf, n = 2, 1 << log2
    a = synth.random_planes(21, nf, width, height, pad)
    b = synth.smooth_planes(22, nf, width, height, pad)
    nb = (width // n) * (height // n)
    want = np.zeros((nf, max(nb, 1)), np.int32)
    oracle.drv("ssd_frames", ptr(a.buf, a.origin), a.pitch, ptr(b.buf, b.origin), b.pitch, width, height, log2, nf, a.frame_stride,
               b.frame_stride, ptr(want), threads=4)
    da, db = to_dev(a.buf), to_dev(b.buf)
    got = dev_full(want.shape, np.int32, -1)
    lib.call("ssd_frames", dptr(da, a.origin), a.pitch, dptr(db, b.origin), b.pitch, width, height, log2, nf, a.frame_stride, b.frame_stride,
             dptr(got))
    if nb:
        assert np.array_equal(to_host(got)[:, :nb], want[:, :nb])
        xy = synth.grid_xy(width, height, n, n)[::3].copy()
        xy[:, 0] += (width - (width // n) * n)  # shift to unaligned positions where the plane allows
        want2 = np.zeros(len(xy), np.int32)
        oracle.drv("ssd_batch", ptr(a.buf, a.origin), a.pitch, ptr(b.buf, b.origin), b.pitch, log2, ptr(xy), len(xy), ptr(want2))
        got2 = dev_full(want2.shape, np.int32, -1)
        dxy = to_dev(xy)
        lib.call("ssd_batch", dptr(da, a.origin), a.pitch, dptr(db, b.origin), b.pitch, log2, dptr(dxy), len(xy), dptr(got2))
        assert np.array_equal(to_host(got2), want2)


@pytest.mark.parametrize("width,height,dx0,dy0", [(256, 128, -4, -4), (200, 136, -4, -4), (136, 72, -3, 1), (64, 64, 0, 0)])
@pytest.mark.parametrize("kind", ["smooth", "flat"])
def test_pyramid_best(oracle, width, height, dx0, dy0, kind):
    """fused argmin: {min SAD, first candidate index reaching it} per PU == numpy argmin over the oracle's 64 SADs;
    "flat" frames make every candidate tie, so the first-in-raster-order rule is exercised"""
    nf = 2
    if kind == "smooth":
        src, ref = _frames(nf, width, height, pad=32)     # 16-byte aligned source origin, as the entry point requires
    else:
        src = synth.Planes(synth.aligned_copy(np.full((nf, height + 64, synth.pitch_for(width, 32)), 90, np.uint8)), width, height, 32)
        ref = synth.Planes(synth.aligned_copy(np.full((nf, height + 64, synth.pitch_for(width, 32)), 93, np.uint8)), width, height, 32)
        ref.buf[:, 32 + 5:32 + 40, 32 + 7:32 + 90] = 91   # a patch that makes some candidates strictly better for some PUs
    ds, dr = to_dev(src.buf), to_dev(ref.buf)
    outs = []
    for s in (8, 16, 32, 64):
        npu = (width // s) * (height // s)
        outs.append(dev_full((nf, max(npu, 1), 2), np.int32, -1))
    lib.call("sad_sweep_pyramid_best_frames", dptr(ds, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, width, height, dx0, dy0, nf,
             src.frame_stride, ref.frame_stride, *[dptr(o) for o in outs])
    for s, o in zip((8, 16, 32, 64), outs):
        npu = (width // s) * (height // s)
        if not npu:
            continue
        sad = np.zeros((nf, npu, 64), np.int32)
        oracle.drv("sad_sweep_frames", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, width, height, HEVCASM_RECT(s, s), dx0, dy0, 8, 8,
                   nf, src.frame_stride, ref.frame_stride, ptr(sad), threads=8)
        got = to_host(o)[:, :npu]
        assert np.array_equal(got[..., 1], sad.argmin(-1)), s
        assert np.array_equal(got[..., 0], sad.min(-1)), s


def test_pyramid_best_rejects_unaligned():
    src, ref = _frames(1, 64, 64, pad=32)
    ds, dr = to_dev(src.buf), to_dev(ref.buf)
    o = dev_full((64,), np.int32, 0)
    with pytest.raises(lib.HevcasmError):
        lib.call("sad_sweep_pyramid_best_frames", dptr(ds, src.origin + 1), src.pitch, dptr(dr, ref.origin), ref.pitch, 32, 32, -4, -4, 1, src.frame_stride,
                 ref.frame_stride, dptr(o), dptr(o), dptr(o), dptr(o))


@pytest.mark.parametrize("shape", [(256, 128), (200, 136), (3840, 2160)])
def test_pyramid_packed_matches_full(shape):
    """hevcasm_sad_sweep_pyramid_packed_frames = the int32 pyramid with the 8x8 / 16x16 levels narrowed to uint16 (exact), at ragged and
    at full 4K size; the full form is checked against the oracle above"""
    width, height = shape
    nf = 2
    src = synth.smooth_planes(61, nf, width, height, 16)
    ref = synth.smooth_planes(61, nf, width, height, 16, shift=(2, -1), noise=5)
    src.buf[1][:] = 255   # a frame of extreme differences: every 16x16 SAD is 65 280
    ref.buf[1][:] = 0
    ds, dr = to_dev(src.buf), to_dev(ref.buf)
    full = [dev_full((nf * (width // s) * (height // s) * 64,), np.int32, -1) for s in (8, 16, 32, 64)]
    lib.call("sad_sweep_pyramid_frames", dptr(ds, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, width, height, -4, -4, nf, src.frame_stride,
             ref.frame_stride, *[dptr(o) for o in full])
    import torch
    pk = [torch.full((nf * (width // s) * (height // s) * 64,), 3, dtype=torch.uint16 if s < 32 else torch.int32, device="cuda") for s in (8, 16, 32, 64)]
    lib.call("sad_sweep_pyramid_packed_frames", dptr(ds, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, width, height, -4, -4, nf, src.frame_stride,
             ref.frame_stride, *[dptr(o) for o in pk])
    for s, a, b in zip((8, 16, 32, 64), full, pk):
        assert np.array_equal(to_host(a), to_host(b).astype(np.int32)), s
    assert int(to_host(pk[1]).max()) == 65280


def test_pu_lists_over_frames(oracle):
    """hevcasm_sad_list_frames / hevcasm_ssd_list_frames: PUs of every partition size, each with its own integer vector and frame index, in one
    launch = the oracle's per-size list forms on each frame"""
    width, height, nf = 320, 200, 3
    src, ref = _frames(nf, width, height, pad=40)
    ds, dr = to_dev(src.buf), to_dev(ref.buf)
    rng = synth.splitmix64(78, 8192).astype(np.int64)
    rows, k = [], 0
    for rep in range(6):
        for (w, h) in PARTITIONS:
            rows.append((rng[k] % (width - w), rng[k + 1] % (height - h), w, h, rng[k + 2] % 65 - 32, rng[k + 3] % 65 - 32, rng[k + 4] % nf))
            k += 5
    pus = np.array(rows, np.int16)
    want = np.zeros(len(pus), np.int32)
    for (w, h) in PARTITIONS:
        for f in range(nf):
            sel = np.flatnonzero((pus[:, 2] == w) & (pus[:, 3] == h) & (pus[:, 6] == f))
            if not len(sel):
                continue
            xy, mv = np.ascontiguousarray(pus[sel, 0:2]), np.ascontiguousarray(pus[sel, 4:6])
            part = np.zeros(len(sel), np.int32)
            oracle.drv("sad_batch", ptr(src.buf[f], src.origin), src.pitch, ptr(ref.buf[f], ref.origin), ref.pitch, HEVCASM_RECT(w, h), ptr(xy), ptr(mv),
                       len(sel), ptr(part))
            want[sel] = part
    got = dev_full(want.shape, np.int32, -7)
    dp = to_dev(pus)
    lib.call("sad_list_frames", dptr(ds, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, dptr(dp), len(pus), src.frame_stride, ref.frame_stride,
             dptr(got))
    assert np.array_equal(to_host(got), want)
    # SSD: square sizes at a zero vector against the oracle's ssd list form; an illegal size gives -1
    sq = np.array([(rng[k + 5 * i] % (width - 64), rng[k + 5 * i + 1] % (height - 64), 4 << (i % 5), 4 << (i % 5), 0, 0, rng[k + 5 * i + 2] % nf) for i in range(60)]
                  + [(0, 0, 6, 8, 0, 0, 0)], np.int16)
    want2 = np.full(len(sq), -1, np.int32)
    for log2 in range(2, 7):
        for f in range(nf):
            sel = np.flatnonzero((sq[:, 2] == 1 << log2) & (sq[:, 6] == f))
            if not len(sel):
                continue
            xy, part = np.ascontiguousarray(sq[sel, 0:2]), np.zeros(len(sel), np.int32)
            oracle.drv("ssd_batch", ptr(src.buf[f], src.origin), src.pitch, ptr(ref.buf[f], ref.origin), ref.pitch, log2, ptr(xy), len(sel), ptr(part))
            want2[sel] = part
    got2 = dev_full(want2.shape, np.int32, -7)
    dq = to_dev(sq)
    lib.call("ssd_list_frames", dptr(ds, src.origin), src.pitch, dptr(dr, ref.origin), ref.pitch, dptr(dq), len(sq), src.frame_stride, ref.frame_stride,
             dptr(got2))
    assert np.array_equal(to_host(got2), want2)
    lib.call("sad_list_frames", dptr(ds), src.pitch, dptr(dr), ref.pitch, None, 0, 0, 0, None)   # empty list: no-op
