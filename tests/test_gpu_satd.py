"""GPU parity: Hadamard SATD (2x2, 4x4, 8x8) and linear SSD vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

from hevcasm_b200 import lib, synth
from oracle.binding import ptr
from gpu_util import to_dev, dev_full, dptr, to_host

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log2", [1, 2, 3])
@pytest.mark.parametrize("shape", [(256, 64, 16), (200, 136, 24), (75, 41, 19)])
def test_satd_frames_and_lists(oracle, log2, shape):
    width, height, pad = shape
    nf, n = 2, 1 << log2
    a = synth.random_planes(501, nf, width, height, pad)
    b = synth.smooth_planes(502, nf, width, height, pad)
    nb = (width // n) * (height // n)
    want = np.zeros((nf, nb), np.int32)
    oracle.drv("hadamard_satd_frames", ptr(a.buf, a.origin), a.pitch, ptr(b.buf, b.origin), b.pitch, width, height, log2, nf, a.frame_stride, b.frame_stride,
               ptr(want), threads=4)
    da, db = to_dev(a.buf), to_dev(b.buf)
    got = dev_full(want.shape, np.int32, -1)
    lib.call("hadamard_satd_frames", dptr(da, a.origin), a.pitch, dptr(db, b.origin), b.pitch, width, height, log2, nf, a.frame_stride, b.frame_stride, dptr(got))
    assert np.array_equal(to_host(got), want)
    xy = synth.grid_xy(width - 3, height - 1, n, n)[::3].copy()
    xy += np.array([3, 1], np.int16)           # unaligned block positions
    want2 = np.zeros(len(xy), np.int32)
    oracle.drv("hadamard_satd_batch", ptr(a.buf, a.origin), a.pitch, ptr(b.buf, b.origin), b.pitch, log2, ptr(xy), len(xy), ptr(want2))
    got2 = dev_full(want2.shape, np.int32, -1)
    dxy = to_dev(xy)
    lib.call("hadamard_satd_batch", dptr(da, a.origin), a.pitch, dptr(db, b.origin), b.pitch, log2, dptr(dxy), len(xy), dptr(got2))
    assert np.array_equal(to_host(got2), want2)


@pytest.mark.parametrize("grid", [None, "1"])
@pytest.mark.parametrize("log2", [2, 3])
def test_satd_tensor_core(oracle, log2, grid, monkeypatch, experiments):
    """4x4 / 8x8 SATD with the horizontal Hadamard pass on tcgen05 (satd_umma.cuh): 16-byte aligned planes, block counts that
    leave partial 128 x 256 tiles, several frames; extreme planes (0 vs 255); "umma_only" fails rather than fall back."""
    monkeypatch.setenv("HEVCASM_SATD_PATH", "umma_only")
    if grid:
        monkeypatch.setenv("HEVCASM_SATD_UMMA_GRID", grid)
    n = 1 << log2
    for width, height, nf in ((200, 136, 3), (8, 8, 1), (416, 600, 2)):
        a = synth.random_planes(520 + width, nf, width, height, 16)
        b = synth.smooth_planes(521, nf, width, height, 16)
        nb = (width // n) * (height // n)
        want = np.zeros((nf, nb), np.int32)
        oracle.drv("hadamard_satd_frames", ptr(a.buf, a.origin), a.pitch, ptr(b.buf, b.origin), b.pitch, width, height, log2, nf, a.frame_stride, b.frame_stride,
                   ptr(want), threads=4)
        da, db = to_dev(a.buf), to_dev(b.buf)
        got = dev_full(want.shape, np.int32, -1)
        lib.call("hadamard_satd_frames", dptr(da, a.origin), a.pitch, dptr(db, b.origin), b.pitch, width, height, log2, nf, a.frame_stride, b.frame_stride, dptr(got))
        assert np.array_equal(to_host(got), want), (width, height, nf)
    test_satd_extremes_and_full_size(min_log2=2)


def test_satd_extremes_and_full_size(min_log2=1):
    """all-0 vs all-255 8x8 blocks: (2 + 64*255) / 4; identical planes: 0 for 2x2 and N/4/(N/2) = 0 otherwise; on a 4K frame"""
    width, height = 3840, 2160
    z = synth.Planes(synth.aligned_copy(np.zeros((1, height, 3840), np.uint8)), width, height, 0)
    dz, do = to_dev(z.buf), to_dev(np.full_like(z.buf, 255))
    for log2 in range(min_log2, 4):
        n = 1 << log2
        out = dev_full(((width // n) * (height // n),), np.int32, -1)
        lib.call("hadamard_satd_frames", dptr(dz), 3840, dptr(do), 3840, width, height, log2, 1, 0, 0, dptr(out))
        assert np.all(to_host(out) == (n // 4 + 255 * n * n) // (n // 2))
        lib.call("hadamard_satd_frames", dptr(dz), 3840, dptr(dz), 3840, width, height, log2, 1, 0, 0, dptr(out))
        assert np.all(to_host(out) == (n // 4) // (n // 2))


@pytest.mark.parametrize("size", [0, 1, 15, 16, 64, 333, 512, 4096, 33025])
def test_ssd_linear(oracle, size):
    runs = 37
    stride = max(size, 1) + 5            # odd run stride: every run has a different alignment
    a = synth.random_bytes(510, runs * stride + 64)
    b = synth.random_bytes(511, runs * stride + 64)
    want = np.zeros(runs, np.int32)
    oracle.drv("ssd_linear_batch", ptr(a, 1), stride, ptr(b, 2), stride, size, runs, ptr(want), threads=4)
    da, db = to_dev(a), to_dev(b)
    got = dev_full((runs,), np.int32, -1)
    lib.call("ssd_linear_batch", dptr(da, 1), stride, dptr(db, 2), stride, size, runs, dptr(got))
    assert np.array_equal(to_host(got), want)


def test_satd_lists_over_frames(oracle):
    """hevcasm_hadamard_satd_list_frames: 2x2 / 4x4 / 8x8 blocks at arbitrary positions of several frames in one call = the oracle's per-size list
    form per frame"""
    width, height, nf = 200, 136, 3
    a = synth.random_planes(511, nf, width, height, 16)
    b = synth.smooth_planes(512, nf, width, height, 16)
    rng = synth.splitmix64(513, 4096).astype(np.int64)
    buckets, k = [], 0
    for log2 in (1, 2, 3):
        n = 1 << log2
        cnt = 90 + 7 * log2
        buckets.append(np.stack([rng[k:k + cnt] % (width - n), rng[k + cnt:k + 2 * cnt] % (height - n), rng[k + 2 * cnt:k + 3 * cnt] % nf], -1).astype(np.int16))
        k += 3 * cnt
    counts = np.array([len(x) for x in buckets], np.int32)
    blks = np.ascontiguousarray(np.concatenate(buckets))
    want, first = np.zeros(len(blks), np.int32), 0
    for c, bk in enumerate(buckets):
        for f in range(nf):
            sel = np.flatnonzero(bk[:, 2] == f)
            xy, part = np.ascontiguousarray(bk[sel, :2]), np.zeros(len(sel), np.int32)
            oracle.drv("hadamard_satd_batch", ptr(a.buf[f], a.origin), a.pitch, ptr(b.buf[f], b.origin), b.pitch, 1 + c, ptr(xy), len(xy), ptr(part))
            want[first + sel] = part
        first += len(bk)
    da, db, dk = to_dev(a.buf), to_dev(b.buf), to_dev(blks)
    got = dev_full(want.shape, np.int32, -1)
    lib.call("hadamard_satd_list_frames", dptr(da, a.origin), a.pitch, dptr(db, b.origin), b.pitch, dptr(dk), ptr(counts), a.frame_stride, b.frame_stride, dptr(got))
    assert np.array_equal(to_host(got), want)
