"""GPU parity: the fused residual pipeline (forward transform -> quantize -> dequantize -> inverse transform + add) against the
composition of the four CPU-oracle stages, bit-exact: levels, coded-block flags and reconstruction."""
import numpy as np
import pytest

from hevcasm_b200 import lib, synth
from oracle.binding import ptr
from gpu_util import to_dev, dev_full, dptr, to_host
from test_oracle_vs_reference import TR

pytestmark = pytest.mark.gpu

# (q_scale, q_shift, q_offset, iq_scale, iq_shift): the reference test's setting and HM-like ones
QP = [(51, 20, 14, 51, 14), (26214, 18, 171 << 7, 18432, 6), (16384, 21, 85 << 7, 640, 5), (32767, 16, 32767, 32767, 15)]


def oracle_pipeline(oracle, res, pred, width, height, log2, trType, qp, nf):
    n = 1 << log2
    nb = (width // n) * (height // n) * nf
    co, lv, dq, cbf = np.zeros(nb * n * n, np.int16), np.zeros(nb * n * n, np.int16), np.zeros(nb * n * n, np.int16), np.zeros(max(nb, 1), np.int32)
    oracle.drv("transform_frames", ptr(co), ptr(res.buf, res.origin), res.pitch, width, height, log2, trType, nf, res.frame_stride, threads=8)
    oracle.drv("quantize_batch", ptr(lv), ptr(co), qp[0], qp[1], qp[2], n * n, nb, ptr(cbf), threads=8)
    oracle.drv("quantize_inverse_batch", ptr(dq), ptr(lv), qp[3], qp[4], nb * n * n, threads=8)
    rec = synth.random_planes(301, nf, width, height, pred.pad)
    oracle.drv("inverse_transform_add_frames", ptr(rec.buf, rec.origin), rec.pitch, ptr(pred.buf, pred.origin), pred.pitch, ptr(dq), width, height, log2, trType,
               nf, rec.frame_stride, pred.frame_stride, threads=8)
    return lv, cbf[:nb], rec


@pytest.mark.parametrize("trType,log2", TR)
@pytest.mark.parametrize("qp", QP)
def test_pipeline(oracle, trType, log2, qp):
    width, height, nf, n = 200, 104, 2, 1 << log2
    for seed, lo, hi in ((310, -256, 255), (311, -32768, 32767)):
        pitch = synth.pitch_for(width, 4, 64)
        res = synth.Planes(synth.random_int16(seed + log2, nf * (height + 8) * pitch, lo, hi).reshape(nf, height + 8, pitch), width, height, 4)
        pred = synth.random_planes(312 + log2, nf, width, height, 8)
        lv_w, cbf_w, rec_w = oracle_pipeline(oracle, res, pred, width, height, log2, trType, qp, nf)
        d_res, d_pred = to_dev(res.buf), to_dev(pred.buf)
        rec_g = to_dev(synth.random_planes(301, nf, width, height, 8).buf)
        lv_g = dev_full(lv_w.shape, np.int16, 0x5a5a)
        cbf_g = dev_full((max(len(cbf_w), 1),), np.int32, -7)
        lib.call("residual_pipeline_frames", dptr(rec_g, rec_w.origin), rec_w.pitch, dptr(lv_g), dptr(cbf_g), dptr(d_res, res.origin), res.pitch,
                 dptr(d_pred, pred.origin), pred.pitch, width, height, log2, trType, *qp, nf, rec_w.frame_stride, res.frame_stride, pred.frame_stride)
        assert np.array_equal(to_host(lv_g), lv_w), (lo, hi)
        assert np.array_equal(to_host(cbf_g)[:len(cbf_w)], cbf_w)
        assert np.array_equal(to_host(rec_g), rec_w.buf)
        # cbf is optional
        lib.call("residual_pipeline_frames", dptr(rec_g, rec_w.origin), rec_w.pitch, dptr(lv_g), None, dptr(d_res, res.origin), res.pitch,
                 dptr(d_pred, pred.origin), pred.pitch, width, height, log2, trType, *qp, nf, rec_w.frame_stride, res.frame_stride, pred.frame_stride)
        assert np.array_equal(to_host(rec_g), rec_w.buf)


def test_pipeline_full_size_properties():
    """4K, 8x8: an all-zero residual reproduces the predictor exactly with cbf == 0 everywhere; and the fused result equals
    the four separate GPU entry points run one after the other (a checksum of checksums over the whole frame)."""
    width, height, nf, log2 = 3840, 2160, 1, 3
    qp = (26214, 18, 171 << 7, 18432, 6)
    pred = synth.random_planes(320, nf, width, height, 16)
    res = synth.residual_planes(321, nf, width, height)
    nb = (width // 8) * (height // 8)
    d_pred, d_res = to_dev(pred.buf), to_dev(res.buf)
    zero = to_dev(np.zeros_like(res.buf))
    rec = to_dev(np.zeros_like(pred.buf))
    lv, cbf = dev_full((nb * 64,), np.int16, 1), dev_full((nb,), np.int32, 1)
    lib.call("residual_pipeline_frames", dptr(rec, pred.origin), pred.pitch, dptr(lv), dptr(cbf), dptr(zero, res.origin), res.pitch, dptr(d_pred, pred.origin),
             pred.pitch, width, height, log2, 0, *qp, nf, pred.frame_stride, res.frame_stride, pred.frame_stride)
    assert not to_host(cbf).any() and not to_host(lv).any()
    assert np.array_equal(to_host(rec)[0, 16:16 + height, 16:16 + width], pred.interior(0))
    # fused == staged
    lib.call("residual_pipeline_frames", dptr(rec, pred.origin), pred.pitch, dptr(lv), dptr(cbf), dptr(d_res, res.origin), res.pitch, dptr(d_pred, pred.origin),
             pred.pitch, width, height, log2, 0, *qp, nf, pred.frame_stride, res.frame_stride, pred.frame_stride)
    co, lv2, dq = (dev_full((nb * 64,), np.int16, 0) for _ in range(3))
    cbf2 = dev_full((nb,), np.int32, 0)
    rec2 = to_dev(np.zeros_like(pred.buf))
    lib.call("transform_frames", dptr(co), dptr(d_res, res.origin), res.pitch, width, height, log2, 0, nf, res.frame_stride)
    lib.call("quantize_batch", dptr(lv2), dptr(co), qp[0], qp[1], qp[2], 64, nb, dptr(cbf2))
    lib.call("quantize_inverse_batch", dptr(dq), dptr(lv2), qp[3], qp[4], nb * 64)
    lib.call("inverse_transform_add_frames", dptr(rec2, pred.origin), pred.pitch, dptr(d_pred, pred.origin), pred.pitch, dptr(dq), width, height, log2, 0, nf,
             pred.frame_stride, pred.frame_stride)
    assert np.array_equal(to_host(lv), to_host(lv2)) and np.array_equal(to_host(cbf), to_host(cbf2)) and np.array_equal(to_host(rec), to_host(rec2))
    assert to_host(cbf).any()


@pytest.mark.parametrize("log2,tr", [(2, 0), (2, 1), (3, 0)])
@pytest.mark.parametrize("shape", [(200, 104), (256, 128)])
def test_from_planes_forms(oracle, log2, tr, shape):
    """residual = src - pred formed on the fly from two 8-bit planes (the input form of f265_lbd_dct_8_avx2): the forward transform and the fused
    pipeline equal the int16-residual forms applied to the differences, extreme differences (+-255) included"""
    width, height = shape
    nf, n = 3, 1 << log2
    src = synth.random_planes(610 + log2, nf, width, height, 8)
    pred = synth.random_planes(611 + log2, nf, width, height, 8)
    src.buf[0][:] = 255
    pred.buf[0][:] = 0
    src.buf[1, :, ::2] = 0
    pred.buf[1, :, ::2] = 255
    res = synth.Planes((src.buf.astype(np.int16) - pred.buf.astype(np.int16)), width, height, 8)
    nb = (width // n) * (height // n) * nf
    want = np.zeros(nb * n * n, np.int16)
    oracle.drv("transform_frames", ptr(want), ptr(res.buf, res.origin), res.pitch, width, height, log2, tr, nf, res.frame_stride, threads=4)
    ds, dp = to_dev(src.buf), to_dev(pred.buf)
    got = dev_full(want.shape, np.int16, 0x5a5a)
    lib.call("transform_from_planes_frames", dptr(got), dptr(ds, src.origin), src.pitch, dptr(dp, pred.origin), pred.pitch, width, height, log2, tr, nf, src.frame_stride,
             pred.frame_stride)
    assert np.array_equal(to_host(got), want)
    for qp in QP[:2]:
        lv_w, cbf_w, rec_w = oracle_pipeline(oracle, res, pred, width, height, log2, tr, qp, nf)
        rec = to_dev(synth.random_planes(301, nf, width, height, 8).buf)
        lv, cbf = dev_full(lv_w.shape, np.int16, 1), dev_full(cbf_w.shape, np.int32, 1)
        lib.call("residual_from_planes_pipeline_frames", dptr(rec, rec_w.origin), rec_w.pitch, dptr(lv), dptr(cbf), dptr(ds, src.origin), src.pitch, dptr(dp, pred.origin),
                 pred.pitch, width, height, log2, tr, *qp, nf, rec_w.frame_stride, src.frame_stride, pred.frame_stride)
        assert np.array_equal(to_host(lv), lv_w) and np.array_equal(to_host(cbf), cbf_w) and np.array_equal(to_host(rec), rec_w.buf), qp


@pytest.mark.parametrize("qp", QP[:2])
def test_pipeline_32x32_two_kernel_path_at_full_size(oracle, qp):
    """32x32 on a batch that fills the chip (2 x 4K = 1020 tiles >= 6 per SM): the default dispatch runs the tensor-core forward transform with the
    quantiser in its epilogue, then the dequantising inverse - levels, cbf and reconstruction against the oracle's composition, both the
    [-256, 255] and the full int16 residual ranges"""
    width, height, nf, log2 = 3840, 2160, 2, 5
    for seed, lo, hi in ((330, -256, 255), (331, -32768, 32767)):
        pitch = synth.pitch_for(width, 0, 128)
        res = synth.Planes(synth.random_int16(seed, nf * height * pitch, lo, hi).reshape(nf, height, pitch), width, height, 0)
        pred = synth.random_planes(332, nf, width, height, 16)
        lv_w, cbf_w, rec_w = oracle_pipeline(oracle, res, pred, width, height, log2, 0, qp, nf)
        d_res, d_pred = to_dev(res.buf), to_dev(pred.buf)
        rec_g = to_dev(synth.random_planes(333, nf, width, height, 16).buf)
        lv_g = dev_full(lv_w.shape, np.int16, 0x5a5a)
        cbf_g = dev_full((len(cbf_w),), np.int32, -7)
        lib.call("residual_pipeline_frames", dptr(rec_g, rec_w.origin), rec_w.pitch, dptr(lv_g), dptr(cbf_g), dptr(d_res, res.origin), res.pitch,
                 dptr(d_pred, pred.origin), pred.pitch, width, height, log2, 0, *qp, nf, rec_w.frame_stride, res.frame_stride, pred.frame_stride)
        assert np.array_equal(to_host(lv_g), lv_w), (lo, hi)
        assert np.array_equal(to_host(cbf_g), cbf_w)
        # the oracle's reconstruction covers the whole-block area; compare there (2160 = 67.5 blocks: the last half row of blocks is not coded)
        bh = height // 32 * 32
        got = synth.Planes(to_host(rec_g), width, height, 16)
        for f in range(nf):
            assert np.array_equal(got.interior(f)[:bh], rec_w.interior(f)[:bh]), f
