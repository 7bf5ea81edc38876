"""GPU parity: forward transforms and inverse transform + add (4x4 DST/DCT .. 32x32) vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

from hevcasm_b200 import lib, synth
from oracle.binding import ptr
from gpu_util import to_dev, dev_full, dptr, to_host
from test_oracle_vs_reference import TR

pytestmark = pytest.mark.gpu


def _res_planes(seed, nf, width, height, lo, hi, pad=4):
    pitch = synth.pitch_for(width, pad, 64)
    rows = height + 2 * pad
    buf = synth.random_int16(seed, nf * rows * pitch, lo, hi).reshape(nf, rows, pitch)
    return synth.Planes(buf, width, height, pad)


@pytest.mark.parametrize("trType,log2", TR)
@pytest.mark.parametrize("rng", [(-256, 255), (-32768, 32767)])
def test_forward_frames(oracle, trType, log2, rng):
    width, height, nf, n = 200, 104, 3, 1 << log2
    res = _res_planes(60 + log2, nf, width, height, *rng)
    nb = (width // n) * (height // n)
    want = np.zeros(nf * nb * n * n, np.int16)
    oracle.drv("transform_frames", ptr(want), ptr(res.buf, res.origin), res.pitch, width, height, log2, trType, nf, res.frame_stride, threads=8)
    dres = to_dev(res.buf)
    got = dev_full(want.shape, np.int16, 0x5a5a)
    lib.call("transform_frames", dptr(got), dptr(dres, res.origin), res.pitch, width, height, log2, trType, nf, res.frame_stride)
    assert np.array_equal(to_host(got), want)


@pytest.mark.parametrize("grid", [None, "2", "1"])
@pytest.mark.parametrize("rng", [(-256, 255), (-32768, 32767), (-20000, -12000)])
@pytest.mark.parametrize("log2", [5, 4])
def test_forward_tensor_core(oracle, log2, rng, grid, monkeypatch, experiments):
    """forward 16x16 / 32x32 with the first stage on tcgen05 (transform_fwd_umma.cuh): 16-byte aligned planes, block counts that leave
    partial 4x4-block tiles, several frames; "umma_only" makes the call fail rather than fall back, so a pass is the tensor
    path.  Full-range and all-negative inputs exercise the high-byte product and the unpacked second-stage butterfly; with
    1 or 2 CTAs each CTA walks over several tiles (accumulator / stage rotation)."""
    monkeypatch.setenv("HEVCASM_FWD_PATH", "umma_only")
    if grid:
        monkeypatch.setenv("HEVCASM_FWD_UMMA_GRID", grid)
    n = 1 << log2
    for width, height, nf in ((224, 160, 3), (32, 32, 1), (416, 300, 2)):
        res = _res_planes(160 + width, nf, width, height, *rng, pad=8)
        nb = (width // n) * (height // n)
        want = np.zeros(nf * nb * n * n, np.int16)
        oracle.drv("transform_frames", ptr(want), ptr(res.buf, res.origin), res.pitch, width, height, log2, 0, nf, res.frame_stride, threads=8)
        dres = to_dev(res.buf)
        got = dev_full(want.shape, np.int16, 0x5a5a)
        lib.call("transform_frames", dptr(got), dptr(dres, res.origin), res.pitch, width, height, log2, 0, nf, res.frame_stride)
        g = to_host(got)
        bad = np.flatnonzero(g != want)
        assert bad.size == 0, (width, height, nf, int(bad.size), [int(b) // (n * n) for b in bad[:8]], [int(b) % (n * n) for b in bad[:8]], g[bad[:4]], want[bad[:4]])


@pytest.mark.parametrize("trType,log2", TR)
def test_forward_list_unaligned(oracle, trType, log2):
    width, height, n = 160, 96, 1 << log2
    res = _res_planes(70 + log2, 1, width, height, -256, 255, pad=8)
    xy = synth.grid_xy(width - 5, height - 3, n, n)[::2].copy()
    xy += np.array([5, 3], np.int16)       # odd x: rows are only 2-byte aligned
    want = np.zeros(len(xy) * n * n, np.int16)
    oracle.drv("transform_batch", ptr(want), ptr(res.buf, res.origin), res.pitch, log2, trType, ptr(xy), len(xy), threads=4)
    dres, dxy = to_dev(res.buf), to_dev(xy)
    got = dev_full(want.shape, np.int16, 0x5a5a)
    lib.call("transform_batch", dptr(got), dptr(dres, res.origin), res.pitch, log2, trType, dptr(dxy), len(xy))
    assert np.array_equal(to_host(got), want)


def _coeff_sets(seed, n_coef, n):
    full = synth.random_int16(seed, n_coef)                      # full-range, reference residual_decode.c:574
    small = synth.random_int16(seed + 1, n_coef, -600, 600)
    sparse = np.zeros(n_coef, np.int16)
    sparse[::n * n] = synth.random_int16(seed + 2, len(sparse[::n * n]), -4000, 4000)   # DC only
    sparse[1::n * n] = 321
    extreme = np.where(synth.random_bytes(seed + 3, n_coef) & 1, 32767, -32768).astype(np.int16)
    return {"full": full, "small": small, "sparse": sparse, "extreme": extreme}


@pytest.mark.parametrize("trType,log2", TR)
def test_inverse_frames(oracle, trType, log2):
    width, height, nf, n = 200, 104, 2, 1 << log2
    pred = synth.random_planes(80 + log2, nf, width, height, 8)
    nb = (width // n) * (height // n)
    for name, co in _coeff_sets(81 + log2, nf * nb * n * n, n).items():
        want = synth.random_planes(82, nf, width, height, 8)
        got = to_dev(want.buf)
        oracle.drv("inverse_transform_add_frames", ptr(want.buf, want.origin), want.pitch, ptr(pred.buf, pred.origin), pred.pitch, ptr(co), width,
                   height, log2, trType, nf, want.frame_stride, pred.frame_stride, threads=8)
        dp, dc = to_dev(pred.buf), to_dev(co)
        lib.call("inverse_transform_add_frames", dptr(got, want.origin), want.pitch, dptr(dp, pred.origin), pred.pitch, dptr(dc), width, height,
                 log2, trType, nf, want.frame_stride, pred.frame_stride)
        assert np.array_equal(to_host(got), want.buf), name


@pytest.mark.parametrize("path", ["imma", "umma"])
@pytest.mark.parametrize("log2", [4, 5])
def test_inverse_frames_imma_variant(oracle, log2, path, monkeypatch, experiments):
    """the exact tensor-core formulations of the 16x16 / 32x32 inverse: legacy mma.sync s8/u8 -> s32 ("imma", kept for A/B
    profiling) and tcgen05 kind::i8 with TMEM accumulators ("umma")"""
    monkeypatch.setenv("HEVCASM_INV_PATH", path)
    test_inverse_frames(oracle, 0, log2)
    width, height, nf, n = 256, 128, 2, 1 << log2          # 16-byte aligned planes: the plane-aligned instantiation
    pred = synth.random_planes(85, nf, width, height, 16)
    co = synth.random_int16(86 + log2, nf * (width // n) * (height // n) * n * n)
    want = synth.random_planes(87, nf, width, height, 16)
    got = to_dev(want.buf)
    oracle.drv("inverse_transform_add_frames", ptr(want.buf, want.origin), want.pitch, ptr(pred.buf, pred.origin), pred.pitch, ptr(co), width, height, log2, 0, nf,
               want.frame_stride, pred.frame_stride, threads=8)
    dp, dc = to_dev(pred.buf), to_dev(co)
    lib.call("inverse_transform_add_frames", dptr(got, want.origin), want.pitch, dptr(dp, pred.origin), pred.pitch, dptr(dc), width, height, log2, 0, nf,
             want.frame_stride, pred.frame_stride)
    assert np.array_equal(to_host(got), want.buf)


@pytest.mark.parametrize("grid", [None, "2", "1"])
@pytest.mark.parametrize("rng", [(-600, 600), (-32768, 32767)])
@pytest.mark.parametrize("log2", [5, 4])
def test_inverse_tensor_core(oracle, log2, rng, grid, monkeypatch, experiments):
    """inverse 16x16 / 32x32 with the SECOND stage on tcgen05 (transform_inv_umma.cuh): stage 1 in the threads writes the int16
    intermediate into shared memory in the swizzled operand layout.  16-byte aligned planes, block counts that leave partial
    tiles, several frames, full-range coefficients (stage-1 clip, high-byte product); "hybrid_only" fails rather than fall back."""
    monkeypatch.setenv("HEVCASM_INV_PATH", "hybrid_only")
    if grid:
        monkeypatch.setenv("HEVCASM_INV_UMMA_GRID", grid)
    n = 1 << log2
    for width, height, nf in ((224, 160, 3), (32, 32, 1), (416, 300, 2)):
        pred = synth.random_planes(185 + width, nf, width, height, 16)
        co = synth.random_int16(186 + log2, nf * (width // n) * (height // n) * n * n, *rng)
        want = synth.random_planes(187, nf, width, height, 16)
        got = to_dev(want.buf)
        oracle.drv("inverse_transform_add_frames", ptr(want.buf, want.origin), want.pitch, ptr(pred.buf, pred.origin), pred.pitch, ptr(co), width, height, log2, 0,
                   nf, want.frame_stride, pred.frame_stride, threads=8)
        dp, dc = to_dev(pred.buf), to_dev(co)
        lib.call("inverse_transform_add_frames", dptr(got, want.origin), want.pitch, dptr(dp, pred.origin), pred.pitch, dptr(dc), width, height, log2, 0, nf,
                 want.frame_stride, pred.frame_stride)
        assert np.array_equal(to_host(got), want.buf), (width, height, nf)


@pytest.mark.parametrize("trType,log2", TR)
def test_inverse_list_unaligned(oracle, trType, log2):
    width, height, n = 160, 96, 1 << log2
    pred = synth.random_planes(90 + log2, 1, width, height, 8)
    xy = synth.grid_xy(width - 3, height - 1, n, n)[::2].copy()
    xy += np.array([3, 1], np.int16)
    co = synth.random_int16(91 + log2, len(xy) * n * n, -3000, 3000)
    want = synth.random_planes(92, 1, width, height, 8)
    got = to_dev(want.buf)
    oracle.drv("inverse_transform_add_batch", ptr(want.buf, want.origin), want.pitch, ptr(pred.buf, pred.origin), pred.pitch, ptr(co), log2, trType,
               ptr(xy), len(xy), threads=4)
    dp, dc, dxy = to_dev(pred.buf), to_dev(co), to_dev(xy)
    lib.call("inverse_transform_add_batch", dptr(got, want.origin), want.pitch, dptr(dp, pred.origin), pred.pitch, dptr(dc), log2, trType, dptr(dxy),
             len(xy))
    assert np.array_equal(to_host(got), want.buf)


@pytest.mark.parametrize("log2", [2, 3, 4, 5])
def test_roundtrip_property(log2):
    """size-independent property at 1080p: forward DCT then inverse on a zero predictor reproduces the (clipped)
    residual to within the transforms' rounding (max |error| <= 6, mean < 1), and an all-zero residual gives all-zero
    coefficients."""
    width, height, n = 1920, 1080, 1 << log2
    bw, bh = width // n * n, height // n * n
    res = synth.residual_planes(100 + log2, 1, width, height)
    dres = to_dev(res.buf)
    nb = (width // n) * (height // n)
    co = dev_full((nb * n * n,), np.int16, 0)
    lib.call("transform_frames", dptr(co), dptr(dres, res.origin), res.pitch, width, height, log2, 0, 1, res.frame_stride)
    pitch = synth.pitch_for(width, 0)
    pred = dev_full((1, height, pitch), np.uint8, 128)
    out = dev_full((1, height, pitch), np.uint8, 0)
    lib.call("inverse_transform_add_frames", dptr(out), pitch, dptr(pred), pitch, dptr(co), width, height, log2, 0, 1, height * pitch, height * pitch)
    rec = to_host(out)[0, :bh, :bw].astype(np.int32) - 128
    src = np.clip(res.interior(0)[:bh, :bw].astype(np.int32), -128, 127)
    err = np.abs(rec - src)
    assert err.max() <= 6 and err.mean() < 1.0
    zero = dev_full(res.buf.shape, np.int16, 0)
    lib.call("transform_frames", dptr(co), dptr(zero, res.origin), res.pitch, width, height, log2, 0, 1, res.frame_stride)
    assert not to_host(co).any()


def _tu_buckets(seed, nf, width, height):
    """a random quad-tree tiling of every frame into 32x32 / 16x16 / 8x8 / 4x4 transform units (4x4: DST or DCT), bucketed by size class in
    the order the *_list_frames forms take: [4x4 DST, 4x4 DCT, 8x8, 16x16, 32x32], entries (x, y, frame)"""
    r = synth.splitmix64(seed, nf * (width // 32) * (height // 32) * 2).astype(np.int64)
    buckets, k = [[] for _ in range(5)], 0
    for f in range(nf):
        for cy in range(0, height - 31, 32):
            for cx in range(0, width - 31, 32):
                kind, sub = int(r[k] % 5), int(r[k + 1])
                k += 2
                if kind == 4:
                    buckets[4].append((cx, cy, f))
                    continue
                n = 4 << max(kind - 1, 0)          # kinds 0 and 1: 4x4 (DST / DCT mixed), 2: 8x8, 3: 16x16
                for j, y in enumerate(range(cy, cy + 32, n)):
                    for i, x in enumerate(range(cx, cx + 32, n)):
                        c = kind if kind >= 2 else ((sub >> ((i + 8 * j) % 60)) & 1)
                        buckets[c].append((x, y, f))
    return [np.array(b, np.int16).reshape(-1, 3) for b in buckets]


CLASSES = [(1, 2), (0, 2), (0, 3), (0, 4), (0, 5)]   # (trType, log2size) per bucket


def test_tu_lists_over_frames(oracle):
    """hevcasm_inverse_transform_add_list_frames / hevcasm_transform_list_frames: the mixed transform-unit lists of several frames in one call
    = the oracle's per-size list forms on each frame"""
    width, height, nf = 352, 288, 3
    buckets = _tu_buckets(300, nf, width, height)
    assert all(len(b) > 20 for b in buckets)
    counts = np.array([len(b) for b in buckets], np.int32)
    tus = np.ascontiguousarray(np.concatenate(buckets))
    ncoef = [len(b) << (2 * log2) for b, (_, log2) in zip(buckets, CLASSES)]
    start = np.concatenate([[0], np.cumsum(ncoef)])
    # inverse
    pred = synth.random_planes(301, nf, width, height, 8)
    co = synth.random_int16(302, int(start[-1]), -3000, 3000)
    want = synth.random_planes(303, nf, width, height, 8)
    got = to_dev(want.buf)
    for c, ((trType, log2), b) in enumerate(zip(CLASSES, buckets)):
        blk = co[start[c]:start[c + 1]].reshape(len(b), -1)
        for f in range(nf):
            sel = b[:, 2] == f
            xy, cf = np.ascontiguousarray(b[sel, :2]), np.ascontiguousarray(blk[sel])
            oracle.drv("inverse_transform_add_batch", ptr(want.buf[f], want.origin), want.pitch, ptr(pred.buf[f], pred.origin), pred.pitch, ptr(cf), log2,
                       trType, ptr(xy), len(xy), threads=4)
    dp, dc, dt = to_dev(pred.buf), to_dev(co), to_dev(tus)
    lib.call("inverse_transform_add_list_frames", dptr(got, want.origin), want.pitch, dptr(dp, pred.origin), pred.pitch, dptr(dc), dptr(dt), ptr(counts),
             want.frame_stride, pred.frame_stride)
    assert np.array_equal(to_host(got), want.buf)
    # forward
    res = _res_planes(304, nf, width, height, -256, 255)
    want_co = np.zeros(int(start[-1]), np.int16)
    for c, ((trType, log2), b) in enumerate(zip(CLASSES, buckets)):
        out = want_co[start[c]:start[c + 1]].reshape(len(b), -1)
        for f in range(nf):
            sel = np.flatnonzero(b[:, 2] == f)
            xy, cf = np.ascontiguousarray(b[sel, :2]), np.zeros((len(sel), out.shape[1]), np.int16)
            oracle.drv("transform_batch", ptr(cf), ptr(res.buf[f], res.origin), res.pitch, log2, trType, ptr(xy), len(xy), threads=4)
            out[sel] = cf
    dres, dco = to_dev(res.buf), dev_full((int(start[-1]),), np.int16, 0)
    lib.call("transform_list_frames", dptr(dco), dptr(dres, res.origin), res.pitch, dptr(dt), ptr(counts), res.frame_stride)
    assert np.array_equal(to_host(dco), want_co)
    # an empty list is a no-op, a negative count an argument error
    zero = np.zeros(5, np.int32)
    lib.call("inverse_transform_add_list_frames", dptr(got), want.pitch, dptr(dp), pred.pitch, dptr(dc), None, ptr(zero), 0, 0)
    bad = np.array([0, -1, 0, 0, 0], np.int32)
    with pytest.raises(lib.HevcasmError):
        lib.call("transform_list_frames", dptr(dco), dptr(dres), res.pitch, dptr(dt), ptr(bad), 0)
