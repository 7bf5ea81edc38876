"""GPU: the host-memory forms (hevcasm_cuda_context + *_host entry points): host planes in, results back in host memory,
through the chunked copy-in / compute / copy-out ring.  Compared bit-exact with the CPU oracle; arenas are sized so that
the batch needs several chunks and the ring wraps."""
import ctypes as C

import numpy as np
import pytest

from hevcasm_b200 import lib, synth
from hevcasm_b200.abi import HEVCASM_RECT
from oracle.binding import ptr

pytestmark = pytest.mark.gpu


def hp(a, off=0):
    return C.c_void_p(a.ctypes.data + off * a.itemsize)


@pytest.mark.parametrize("arena_mb", [2, 64])
def test_sad_pyramid_host(oracle, arena_mb):
    width, height, nf, pad = 256, 192, 7, 16
    src = synth.smooth_planes(401, nf, width, height, pad)
    ref = synth.smooth_planes(401, nf, width, height, pad, shift=(1, 2), noise=4)
    outs = [np.full(nf * (width // s) * (height // s) * 64, -1, np.int32) for s in (8, 16, 32, 64)]
    with lib.Context(0, arena_mb << 20) as ctx:
        lib.call_host("sad_sweep_pyramid_frames_host", ctx.handle, hp(src.buf, src.origin), src.pitch, hp(ref.buf, ref.origin), ref.pitch, width, height, pad,
                      -4, -4, nf, src.frame_stride, ref.frame_stride, *[hp(o) for o in outs])
    for s, o in zip((8, 16, 32, 64), outs):
        want = np.zeros_like(o)
        oracle.drv("sad_sweep_frames", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, width, height, HEVCASM_RECT(s, s), -4, -4, 8, 8,
                   nf, src.frame_stride, ref.frame_stride, ptr(want), threads=4)
        assert np.array_equal(o, want), s


def test_sad_pyramid_host_rejects_window_outside_padding():
    src = synth.random_planes(1, 1, 64, 64, 4)
    out = np.zeros(64 * 64, np.int32)
    with lib.Context(0, 8 << 20) as ctx:
        with pytest.raises(lib.HevcasmError):
            lib.call_host("sad_sweep_pyramid_frames_host", ctx.handle, hp(src.buf, src.origin), src.pitch, hp(src.buf, src.origin), src.pitch, 64, 64, 4, -8, -8, 1,
                          src.frame_stride, src.frame_stride, hp(out), hp(out), hp(out), hp(out))


@pytest.mark.parametrize("taps,xf,yf", [(8, 1, 3), (8, 0, 2), (4, 5, 0), (4, 0, 0)])
def test_pred_uni_host(oracle, taps, xf, yf):
    width, height, nf, pad = 200, 120, 5, 16
    ref = synth.random_planes(410 + taps, nf, width, height, pad)
    want = synth.random_planes(411, nf, width, height, 4)
    got = synth.Planes(want.buf.copy(), width, height, 4)
    oracle.drv("pred_uni_frames", ptr(want.buf, want.origin), want.pitch, ptr(ref.buf, ref.origin), ref.pitch, width, height, taps, xf, yf, nf, want.frame_stride,
               ref.frame_stride, threads=4)
    with lib.Context(0, 1 << 20) as ctx:
        lib.call_host("pred_uni_frames_host", ctx.handle, hp(got.buf, got.origin), got.pitch, hp(ref.buf, ref.origin), ref.pitch, width, height, pad, taps, xf, yf,
                      nf, got.frame_stride, ref.frame_stride)
    assert np.array_equal(got.buf, want.buf)


@pytest.mark.parametrize("log2", [2, 3, 4, 5])
def test_residual_pipeline_host(oracle, log2):
    from test_gpu_pipeline import oracle_pipeline
    width, height, nf = 200, 104, 5
    qp = (26214, 18, 171 << 7, 18432, 6)
    res = synth.residual_planes(420 + log2, nf, width, height)
    pred = synth.random_planes(421, nf, width, height, 8)
    lv_w, cbf_w, rec_w = oracle_pipeline(oracle, res, pred, width, height, log2, 0, qp, nf)
    rec_g = synth.Planes(synth.random_planes(301, nf, width, height, 8).buf.copy(), width, height, 8)
    lv_g, cbf_g = np.zeros_like(lv_w), np.full(len(cbf_w), -3, np.int32)
    with lib.Context(0, 1 << 20) as ctx:
        lib.call_host("residual_pipeline_frames_host", ctx.handle, hp(rec_g.buf, rec_g.origin), rec_g.pitch, hp(lv_g), hp(cbf_g), hp(res.buf, res.origin), res.pitch,
                      hp(pred.buf, pred.origin), pred.pitch, width, height, log2, 0, *qp, nf, rec_g.frame_stride, res.frame_stride, pred.frame_stride)
    assert np.array_equal(lv_g, lv_w) and np.array_equal(cbf_g, cbf_w) and np.array_equal(rec_g.buf, rec_w.buf)


def test_sad_pyramid_best_host(oracle):
    width, height, nf, pad = 256, 192, 5, 16
    src = synth.smooth_planes(431, nf, width, height, pad)
    ref = synth.smooth_planes(431, nf, width, height, pad, shift=(-2, 1), noise=4)
    outs = [np.full(nf * (width // s) * (height // s) * 2, -1, np.int32) for s in (8, 16, 32, 64)]
    with lib.Context(0, 2 << 20) as ctx:
        lib.call_host("sad_sweep_pyramid_best_frames_host", ctx.handle, hp(src.buf, src.origin), src.pitch, hp(ref.buf, ref.origin), ref.pitch, width, height, pad,
                      -4, -4, nf, src.frame_stride, ref.frame_stride, *[hp(o) for o in outs])
    for s, o in zip((8, 16, 32, 64), outs):
        sad = np.zeros((nf * (width // s) * (height // s), 64), np.int32)
        oracle.drv("sad_sweep_frames", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, width, height, HEVCASM_RECT(s, s), -4, -4, 8, 8,
                   nf, src.frame_stride, ref.frame_stride, ptr(sad), threads=4)
        got = o.reshape(-1, 2)
        assert np.array_equal(got[:, 0], sad.min(-1)) and np.array_equal(got[:, 1], sad.argmin(-1)), s


@pytest.mark.parametrize("near", [False, True])
def test_sad_pyramid_packed_host(oracle, near):
    """the packed form (uint16 for the 8x8 / 16x16 levels) through the host ring, from plain and from NUMA-near page-locked buffers;
    the smooth + extreme (0 vs 255) frames reach the largest SADs a level can hold (16 320, 65 280)"""
    width, height, nf, pad = 256, 192, 5, 16
    src = synth.smooth_planes(441, nf, width, height, pad)
    ref = synth.smooth_planes(441, nf, width, height, pad, shift=(1, -2), noise=4)
    src.buf[0][:] = 0
    ref.buf[0][:] = 255
    shapes = [(nf * (width // s) * (height // s) * 64, np.uint16 if s < 32 else np.int32) for s in (8, 16, 32, 64)]
    if near:
        hs = lib.pinned_array(src.buf.shape, np.uint8, device=0)
        hr = lib.pinned_array(ref.buf.shape, np.uint8, device=0)
        hs[...], hr[...] = src.buf, ref.buf
        outs = [lib.pinned_array((n,), dt, device=0) for n, dt in shapes]
    else:
        hs, hr = src.buf, ref.buf
        outs = [np.zeros(n, dt) for n, dt in shapes]
    for o in outs:
        o[...] = 7
    with lib.Context(0, 2 << 20) as ctx:
        lib.call_host("sad_sweep_pyramid_packed_frames_host", ctx.handle, hp(hs, src.origin), src.pitch, hp(hr, ref.origin), ref.pitch, width, height, pad,
                      -4, -4, nf, src.frame_stride, ref.frame_stride, *[hp(o) for o in outs])
    for s, o in zip((8, 16, 32, 64), outs):
        want = np.zeros(o.shape, np.int32)
        oracle.drv("sad_sweep_frames", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, width, height, HEVCASM_RECT(s, s), -4, -4, 8, 8,
                   nf, src.frame_stride, ref.frame_stride, ptr(want), threads=4)
        assert np.array_equal(o.astype(np.int32), want), s
        if s < 32:
            assert int(o.max()) == s * s * 255


def test_host_form_error_leaves_nothing_in_flight():
    """an argument error that only the compute stage detects (odd alignment of `levels`) must come back after the copy streams have
    drained: the next call on the same context works, and the caller's buffers are not touched after the failed call returns"""
    width, height, nf = 64, 64, 3
    res = synth.residual_planes(450, nf, width, height)
    pred = synth.random_planes(451, nf, width, height, 8)
    rec = synth.Planes(pred.buf.copy(), width, height, 8)
    lv = np.zeros(nf * 64 * 64 + 8, np.int16)
    cbf = np.zeros(nf * 64, np.int32)
    with lib.Context(0, 1 << 20) as ctx:
        with pytest.raises(lib.HevcasmError):   # q_shift 15 is outside the quantiser's domain
            lib.call_host("residual_pipeline_frames_host", ctx.handle, hp(rec.buf, rec.origin), rec.pitch, hp(lv), hp(cbf), hp(res.buf, res.origin), res.pitch,
                          hp(pred.buf, pred.origin), pred.pitch, width, height, 3, 0, 26214, 15, 171 << 7, 18432, 6, nf, rec.frame_stride, res.frame_stride,
                          pred.frame_stride)
        lib.call_host("residual_pipeline_frames_host", ctx.handle, hp(rec.buf, rec.origin), rec.pitch, hp(lv), hp(cbf), hp(res.buf, res.origin), res.pitch,
                      hp(pred.buf, pred.origin), pred.pitch, width, height, 3, 0, 26214, 18, 171 << 7, 18432, 6, nf, rec.frame_stride, res.frame_stride,
                      pred.frame_stride)
    assert cbf.any()
