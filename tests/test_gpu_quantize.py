"""GPU parity: quantize / quantize_inverse / quantize_reconstruct vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

from hevcasm_b200 import lib, synth
from oracle.binding import ptr
from gpu_util import to_dev, dev_full, dptr, to_host
from test_oracle_vs_reference import QUANT, DEQUANT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scale,shift,offset", QUANT)
@pytest.mark.parametrize("npb", [16, 64, 256, 1024])
def test_quantize(oracle, scale, shift, offset, npb):
    n_blocks = 203
    src = synth.random_int16(31 + npb, npb * n_blocks)
    src[:4] = [-32768, 32767, 0, -1]
    src[npb * 5:npb * 9] = 0                      # all-zero blocks: cbf must be 0
    src[npb * 9:npb * 10] = 0
    src[npb * 10 - 1] = 20000                     # single trailing coefficient
    want, wcbf = np.zeros_like(src), np.zeros(n_blocks, np.int32)
    oracle.drv("quantize_batch", ptr(want), ptr(src), scale, shift, offset, npb, n_blocks, ptr(wcbf), threads=4)
    d_src = to_dev(src)
    got, gcbf = dev_full(src.shape, np.int16, 77), dev_full((n_blocks,), np.int32, -5)
    lib.call("quantize_batch", dptr(got), dptr(d_src), scale, shift, offset, npb, n_blocks, dptr(gcbf))
    assert np.array_equal(to_host(got), want)
    assert np.array_equal(to_host(gcbf), wcbf)
    got.fill_(77)
    lib.call("quantize_batch", dptr(got), dptr(d_src), scale, shift, offset, npb, n_blocks, None)
    assert np.array_equal(to_host(got), want)


@pytest.mark.parametrize("scale,shift", DEQUANT)
def test_quantize_inverse(oracle, scale, shift):
    for n in (16, 1024 * 37 + 8, 4096 + 5):
        src = synth.random_int16(41 + n, n)
        want = np.zeros_like(src)
        oracle.drv("quantize_inverse_batch", ptr(want), ptr(src), scale, shift, n, threads=2)
        d_src = to_dev(src)
        got = dev_full(src.shape, np.int16, 77)
        lib.call("quantize_inverse_batch", dptr(got), dptr(d_src), scale, shift, n)
        assert np.array_equal(to_host(got), want)


@pytest.mark.parametrize("log2", [2, 3, 4, 5])
def test_quantize_reconstruct(oracle, log2):
    width, height, nf, n = 200, 72, 2, 1 << log2
    pred = synth.random_planes(51, nf, width, height, 8)
    nb = (width // n) * (height // n)
    for lo, hi in ((-256, 255), (-32768, 32767)):
        res = synth.random_int16(52 + log2, nf * nb * n * n, lo, hi)
        want = synth.random_planes(53, nf, width, height, 8)
        got = to_dev(want.buf)
        oracle.drv("quantize_reconstruct_frames", ptr(want.buf, want.origin), want.pitch, ptr(pred.buf, pred.origin), pred.pitch, ptr(res), width,
                   height, log2, nf, want.frame_stride, pred.frame_stride, threads=4)
        dp, dr = to_dev(pred.buf), to_dev(res)
        lib.call("quantize_reconstruct_frames", dptr(got, want.origin), want.pitch, dptr(dp, pred.origin), pred.pitch, dptr(dr), width, height,
                 log2, nf, want.frame_stride, pred.frame_stride)
        assert np.array_equal(to_host(got), want.buf)   # also proves nothing outside the block grid was touched
        # list form at odd (unaligned) positions
        xy = synth.grid_xy(width - 3, height - 1, n, n)[::2].copy()
        xy += np.array([3, 1], np.int16)
        want2 = synth.random_planes(54, 1, width, height, 8)
        got2 = to_dev(want2.buf)
        oracle.drv("quantize_reconstruct_batch", ptr(want2.buf, want2.origin), want2.pitch, ptr(pred.buf, pred.origin), pred.pitch, ptr(res),
                   log2, ptr(xy), len(xy))
        dxy = to_dev(xy)
        lib.call("quantize_reconstruct_batch", dptr(got2, want2.origin), want2.pitch, dptr(dp, pred.origin), pred.pitch, dptr(dr), log2,
                 dptr(dxy), len(xy))
        assert np.array_equal(to_host(got2), want2.buf)


def test_reconstruct_tu_lists_over_frames(oracle):
    """hevcasm_quantize_reconstruct_list_frames: mixed-size TU lists of several frames in one call = the oracle's per-size list form per frame"""
    from test_gpu_transform import _tu_buckets
    width, height, nf = 352, 288, 3
    b5 = _tu_buckets(310, nf, width, height)
    buckets = [np.ascontiguousarray(np.concatenate([b5[0], b5[1]]))] + b5[2:]      # both 4x4 classes are one size here
    counts = np.array([len(b) for b in buckets], np.int32)
    tus = np.ascontiguousarray(np.concatenate(buckets))
    nres = [len(b) << (2 * (2 + c)) for c, b in enumerate(buckets)]
    start = np.concatenate([[0], np.cumsum(nres)])
    pred = synth.random_planes(311, nf, width, height, 8)
    res = synth.random_int16(312, int(start[-1]), -600, 600)
    want = synth.random_planes(313, nf, width, height, 8)
    got = to_dev(want.buf)
    for c, b in enumerate(buckets):
        blk = res[start[c]:start[c + 1]].reshape(len(b), -1)
        for f in range(nf):
            sel = b[:, 2] == f
            xy, rf = np.ascontiguousarray(b[sel, :2]), np.ascontiguousarray(blk[sel])
            oracle.drv("quantize_reconstruct_batch", ptr(want.buf[f], want.origin), want.pitch, ptr(pred.buf[f], pred.origin), pred.pitch, ptr(rf), 2 + c,
                       ptr(xy), len(xy))
    dp, dr, dt = to_dev(pred.buf), to_dev(res), to_dev(tus)
    lib.call("quantize_reconstruct_list_frames", dptr(got, want.origin), want.pitch, dptr(dp, pred.origin), pred.pitch, dptr(dr), dptr(dt), ptr(counts),
             want.frame_stride, pred.frame_stride)
    assert np.array_equal(to_host(got), want.buf)
