"""GPU parity: inter-prediction interpolation (luma 8-tap / chroma 4-tap, uni and bi; plane and PU-list forms) vs the CPU
oracle, bit-exact, at every fractional position (the reference's own test only visits fractions 0 and 1)."""
import numpy as np
import pytest

from hevcasm_b200 import lib, synth
from oracle.binding import ptr
from gpu_util import to_dev, dev_full, dptr, to_host

pytestmark = pytest.mark.gpu

# reference pred_inter.c:438-444: luma partitions; chroma uses the same list scaled by 1/2
LUMA_PARTITIONS = [(64, 64), (64, 48), (64, 32), (64, 16), (48, 64), (32, 64), (32, 32), (32, 24), (32, 16), (32, 8), (24, 32), (16, 64),
                   (16, 32), (16, 16), (16, 12), (16, 8), (16, 4), (12, 16), (8, 32), (8, 16), (8, 8), (8, 4), (4, 16), (4, 8)]


def _ref_planes(seed, nf, width, height, pad=16, kind="random"):
    if kind == "random":
        return synth.random_planes(seed, nf, width, height, pad)
    return synth.smooth_planes(seed, nf, width, height, pad)


@pytest.mark.parametrize("taps", [8, 4])
@pytest.mark.parametrize("shape", [(200, 136), (128, 32), (131, 37), (8, 8)])
def test_uni_planes_all_fractions(oracle, taps, shape):
    width, height = shape
    nf = 2
    ref = _ref_planes(200 + taps, nf, width, height)
    dr = to_dev(ref.buf)
    nfrac = 4 if taps == 8 else 8
    for yf in range(nfrac):
        for xf in range(nfrac):
            want = synth.random_planes(201, nf, width, height, 16)
            got = to_dev(want.buf)
            oracle.drv("pred_uni_frames", ptr(want.buf, want.origin), want.pitch, ptr(ref.buf, ref.origin), ref.pitch, width, height, taps, xf, yf, nf,
                       want.frame_stride, ref.frame_stride, threads=8)
            lib.call("pred_uni_frames", dptr(got, want.origin), want.pitch, dptr(dr, ref.origin), ref.pitch, width, height, taps, xf, yf, nf, want.frame_stride,
                     ref.frame_stride)
            assert np.array_equal(to_host(got), want.buf), (xf, yf)  # whole buffer: nothing outside width x height may change


@pytest.mark.parametrize("taps", [8, 4])
def test_generic_plane_kernels_all_fractions(oracle, taps, monkeypatch, experiments):
    """the alignment-agnostic plane kernels (taken when reference rows are not 16-byte aligned), forced on aligned planes"""
    monkeypatch.setenv("HEVCASM_PRED_GENERIC", "1")
    test_uni_planes_all_fractions(oracle, taps, (200, 136))
    test_bi_planes(oracle, taps)


@pytest.mark.parametrize("env", [{"HEVCASM_PRED_STREAM": "ldg"}, {"HEVCASM_PRED_STREAM": "ldg", "HEVCASM_PRED_PATH": "stream"}, {"HEVCASM_PRED_PATH": "tile"},
                                 {"HEVCASM_PRED_HV": "stream"}])
@pytest.mark.parametrize("taps", [8, 4])
def test_fallback_plane_kernels_all_fractions(oracle, taps, env, monkeypatch, experiments):
    """the kernels behind the TMA-fed one: LDG-fed streaming kernel (planes the TMA unit cannot describe) and the shared-memory
    tile kernels, each forced on planes the default dispatch would hand to the TMA kernel"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    test_uni_planes_all_fractions(oracle, taps, (200, 136))
    test_uni_planes_all_fractions(oracle, taps, (131, 37))
    test_bi_planes(oracle, taps)


@pytest.mark.parametrize("grid", [None, "2", "1"])
@pytest.mark.parametrize("taps", [8, 4])
def test_tensor_core_plane_kernels(oracle, taps, grid, monkeypatch, experiments):
    """the tcgen05 kernels (vertical pass as an int8 Toeplitz product for one reference, horizontal pass for two), pinned for
    every plane size and both filters; with 1 or 2 CTAs each CTA walks over several tiles, which exercises the accumulator / stage /
    output-buffer rotation of the producer-consumer pipeline.  Odd widths leave through the byte-store path of the right-hand tile."""
    monkeypatch.setenv("HEVCASM_PRED_HV", "umma")
    if grid:
        monkeypatch.setenv("HEVCASM_PRED_UMMA_GRID", grid)
    for shape in ((200, 136), (131, 37), (8, 8), (464, 400), (700, 130)):
        width, height = shape
        nf = 3
        ref = _ref_planes(220 + taps, nf, width, height)
        dr = to_dev(ref.buf)
        nfrac = 4 if taps == 8 else 8
        for xf, yf in ((1, 1), (2, nfrac - 1), (nfrac - 1, 2)):
            want = synth.random_planes(221, nf, width, height, 16)
            got = to_dev(want.buf)
            oracle.drv("pred_uni_frames", ptr(want.buf, want.origin), want.pitch, ptr(ref.buf, ref.origin), ref.pitch, width, height, taps, xf, yf, nf,
                       want.frame_stride, ref.frame_stride, threads=8)
            lib.call("pred_uni_frames", dptr(got, want.origin), want.pitch, dptr(dr, ref.origin), ref.pitch, width, height, taps, xf, yf, nf, want.frame_stride,
                     ref.frame_stride)
            assert np.array_equal(to_host(got), want.buf), (shape, xf, yf)
    test_bi_planes(oracle, taps)
    if taps == 8 and grid is None:
        test_bi_extremes(oracle)
    monkeypatch.setenv("HEVCASM_PRED_BI", "hfirst")   # the other two-reference kernel: horizontal pass on the tensor cores
    test_bi_planes(oracle, taps)
    if taps == 8 and grid is None:
        test_bi_extremes(oracle)


@pytest.mark.parametrize("taps", [8, 4])
def test_uni_planes_unaligned_pointers(oracle, taps):
    """destination and reference origins at odd byte offsets, odd pitch"""
    width, height, nf = 77, 45, 1
    pitch = 131
    rows = height + 24
    ref = synth.random_bytes(210, rows * pitch).reshape(1, rows, pitch)
    org = 12 * pitch + 13
    nfrac = 4 if taps == 8 else 8
    dr = to_dev(ref)
    for xf, yf in ((1, 0), (0, 1), (nfrac - 1, 2), (0, 0)):
        want = synth.random_bytes(211, rows * pitch).reshape(1, rows, pitch).copy()
        got = to_dev(want)
        oracle.drv("pred_uni_frames", ptr(want, org + 1), pitch, ptr(ref, org), pitch, width, height, taps, xf, yf, nf, rows * pitch, rows * pitch, threads=2)
        lib.call("pred_uni_frames", dptr(got, org + 1), pitch, dptr(dr, org), pitch, width, height, taps, xf, yf, nf, rows * pitch, rows * pitch)
        assert np.array_equal(to_host(got), want), (xf, yf)


@pytest.mark.parametrize("taps", [8, 4])
def test_bi_planes(oracle, taps):
    width, height, nf = 200, 72, 2
    r0 = _ref_planes(220, nf, width, height)
    r1 = _ref_planes(221, nf, width, height, kind="smooth")
    d0, d1 = to_dev(r0.buf), to_dev(r1.buf)
    nfrac = 4 if taps == 8 else 8
    fr = synth.splitmix64(222, 4 * 20) % np.uint64(nfrac)
    cases = [tuple(int(v) for v in fr[4 * i:4 * i + 4]) for i in range(20)] + [(f, f, f, f) for f in range(nfrac)] + [(0, 0, 0, 0), (0, 1, 1, 0)]
    for c in cases:
        want = synth.random_planes(223, nf, width, height, 16)
        got = to_dev(want.buf)
        oracle.drv("pred_bi_frames", ptr(want.buf, want.origin), want.pitch, ptr(r0.buf, r0.origin), ptr(r1.buf, r1.origin), r0.pitch, width, height, taps, *c,
                   nf, want.frame_stride, r0.frame_stride, threads=8)
        lib.call("pred_bi_frames", dptr(got, want.origin), want.pitch, dptr(d0, r0.origin), dptr(d1, r1.origin), r0.pitch, width, height, taps, *c, nf,
                 want.frame_stride, r0.frame_stride)
        assert np.array_equal(to_host(got), want.buf), c


def test_bi_extremes(oracle):
    """0/255 checkerboards drive the bi-pred intermediates to their extremes, where the reference's C wraps to int16
    (SURVEY.md 8(a) divergence table): the GPU must wrap identically."""
    width, height, nf, pad = 64, 64, 1, 16
    pitch = synth.pitch_for(width, pad)
    yy, xx = np.mgrid[0:height + 2 * pad, 0:pitch]
    for pat in (((xx + yy) & 1) * 255, (xx & 1) * 255, (yy & 1) * 255, np.full_like(xx, 255)):
        r = synth.Planes(synth.aligned_copy(pat.astype(np.uint8)[None]), width, height, pad)
        d = to_dev(r.buf)
        for taps, nfrac in ((8, 4), (4, 8)):
            for f in range(nfrac):
                want = synth.random_planes(230, nf, width, height, pad)
                got = to_dev(want.buf)
                c = (f, f, (f + 1) % nfrac, f)
                oracle.drv("pred_bi_frames", ptr(want.buf, want.origin), want.pitch, ptr(r.buf, r.origin), ptr(r.buf, r.origin + 1), r.pitch, width, height,
                           taps, *c, nf, want.frame_stride, r.frame_stride, threads=2)
                lib.call("pred_bi_frames", dptr(got, want.origin), want.pitch, dptr(d, r.origin), dptr(d, r.origin + 1), r.pitch, width, height, taps, *c, nf,
                         want.frame_stride, r.frame_stride)
                assert np.array_equal(to_host(got), want.buf), (taps, c)


def _pu_list(taps, width, height, seed):
    """non-overlapping PUs of every reference partition size with random quarter/eighth-sample motion vectors"""
    scale = 1 if taps == 8 else 2
    parts = [(w // scale, h // scale) for w, h in LUMA_PARTITIONS]
    rng = synth.splitmix64(seed, 4096).astype(np.int64)
    pus, x, y, rowh, k = [], 0, 0, 0, 0
    for rep in range(3):
        for (w, h) in parts:
            if x + w > width:
                x, y, rowh = 0, y + rowh, 0
            if y + h > height:
                continue
            mvx = int(rng[k] % 97) - 48
            mvy = int(rng[k + 1] % 97) - 48
            if rep == 0 and k % 5 == 0:
                mvx &= ~(3 if taps == 8 else 7)     # some full-pel / H-only / V-only vectors
            if rep == 0 and k % 7 == 0:
                mvy &= ~(3 if taps == 8 else 7)
            pus.append((x, y, w, h, mvx, mvy, int(rng[k + 2] % 97) - 48, int(rng[k + 3] % 97) - 48))
            k += 4
            x += w
            rowh = max(rowh, h)
    return np.array(pus, np.int16)


@pytest.mark.parametrize("taps", [8, 4])
def test_uni_pu_list(oracle, taps):
    width, height = 512, 420
    ref = _ref_planes(240 + taps, 1, width, height, pad=32)
    pus8 = _pu_list(taps, width, height, 241)
    pus = np.ascontiguousarray(pus8[:, :6])
    want = synth.random_planes(242, 1, width, height, 32)
    got = to_dev(want.buf)
    oracle.drv("pred_uni_batch", ptr(want.buf, want.origin), want.pitch, ptr(ref.buf, ref.origin), ref.pitch, taps, ptr(pus), len(pus), threads=4)
    dr, dp = to_dev(ref.buf), to_dev(pus)
    lib.call("pred_uni_batch", dptr(got, want.origin), want.pitch, dptr(dr, ref.origin), ref.pitch, taps, dptr(dp), len(pus))
    assert len(pus) > 40
    assert np.array_equal(to_host(got), want.buf)
    lib.call("pred_uni_batch", dptr(got), want.pitch, dptr(dr), ref.pitch, taps, None, 0)  # empty list is a no-op


@pytest.mark.parametrize("taps", [8, 4])
def test_bi_pu_list(oracle, taps):
    width, height = 512, 420
    r0 = _ref_planes(250 + taps, 1, width, height, pad=32)
    r1 = _ref_planes(251 + taps, 1, width, height, pad=32, kind="smooth")
    pus = _pu_list(taps, width, height, 252)
    want = synth.random_planes(253, 1, width, height, 32)
    got = to_dev(want.buf)
    oracle.drv("pred_bi_batch", ptr(want.buf, want.origin), want.pitch, ptr(r0.buf, r0.origin), ptr(r1.buf, r1.origin), r0.pitch, taps, ptr(pus), len(pus),
               threads=4)
    d0, d1, dp = to_dev(r0.buf), to_dev(r1.buf), to_dev(pus)
    lib.call("pred_bi_batch", dptr(got, want.origin), want.pitch, dptr(d0, r0.origin), dptr(d1, r1.origin), r0.pitch, taps, dptr(dp), len(pus))
    assert np.array_equal(to_host(got), want.buf)


def test_full_size_properties():
    """4K plane, size-independent properties: the full-pel position is a copy; on a constant plane every fractional
    position returns the constant (the taps sum to 64); H-only and V-only commute with transposing the content."""
    width, height, pad = 3840, 2160, 16
    ref = synth.random_planes(260, 1, width, height, pad)
    dr = to_dev(ref.buf)
    out = to_dev(np.zeros_like(ref.buf))
    lib.call("pred_uni_frames", dptr(out, ref.origin), ref.pitch, dptr(dr, ref.origin), ref.pitch, width, height, 8, 0, 0, 1, ref.frame_stride, ref.frame_stride)
    assert np.array_equal(to_host(out)[0, pad:pad + height, pad:pad + width], ref.interior(0))
    const = to_dev(np.full_like(ref.buf, 173))
    for taps, xf, yf in ((8, 1, 3), (8, 2, 0), (4, 5, 7), (4, 0, 3)):
        lib.call("pred_uni_frames", dptr(out, ref.origin), ref.pitch, dptr(const, ref.origin), ref.pitch, width, height, taps, xf, yf, 1, ref.frame_stride,
                 ref.frame_stride)
        assert np.all(to_host(out)[0, pad:pad + height, pad:pad + width] == 173)
    # transpose property on a square crop: V-filtering the transposed content == transpose of H-filtering
    n = 1024
    sq = synth.random_planes(261, 1, n, n, pad)
    a = sq.buf[0, :n + 2 * pad, :n + 2 * pad]
    t = synth.aligned_empty((1, n + 2 * pad, sq.pitch), np.uint8)
    t[...] = 0
    t[0, :, :n + 2 * pad] = a.T
    d_a, d_t = to_dev(sq.buf), to_dev(t)
    o_a, o_t = to_dev(np.zeros_like(sq.buf)), to_dev(np.zeros_like(t))
    lib.call("pred_uni_frames", dptr(o_a, sq.origin), sq.pitch, dptr(d_a, sq.origin), sq.pitch, n, n, 8, 3, 0, 1, sq.frame_stride, sq.frame_stride)
    lib.call("pred_uni_frames", dptr(o_t, sq.origin), sq.pitch, dptr(d_t, sq.origin), sq.pitch, n, n, 8, 0, 3, 1, sq.frame_stride, sq.frame_stride)
    ha = to_host(o_a)[0, pad:pad + n, pad:pad + n]
    vt = to_host(o_t)[0, pad:pad + n, pad:pad + n]
    assert np.array_equal(ha, vt.T)


@pytest.mark.parametrize("taps", [8, 4])
def test_bounded_forms_stay_inside_the_footprint(oracle, taps):
    """hevcasm_pred_*_frames_bounded with no slack: the reference planes sit at the very start and end of their device allocation, padded by
    exactly the filter footprint (taps/2-1 before, taps/2 after) - a kernel that read whole 16-byte chunks around it would leave the
    allocation (compute-sanitizer memcheck runs this test).  Results equal the unbounded forms' and the oracle's."""
    import torch
    width, height, nf = 120, 72, 2
    lo, hi = taps // 2 - 1, taps // 2
    pitch, rows = width + lo + hi, height + lo + hi           # tight: no alignment padding at all
    n = nf * rows * pitch
    host = synth.random_bytes(300 + taps, n).reshape(nf, rows, pitch)
    host1 = synth.random_bytes(301 + taps, n).reshape(nf, rows, pitch)
    d0 = torch.from_numpy(host.copy()).cuda()   # exactly n bytes: the footprint ends with the allocation
    d1 = torch.from_numpy(host1.copy()).cuda()
    org = lo * pitch + lo
    nfrac = 4 if taps == 8 else 8
    for xf, yf in ((0, 0), (1, 0), (0, nfrac - 1), (2, 1)):
        want = np.zeros((nf, height, width), np.uint8)
        oracle.drv("pred_uni_frames", ptr(want), width, ptr(host, org), pitch, width, height, taps, xf, yf, nf, width * height, rows * pitch, threads=4)
        got = dev_full(want.shape, np.uint8, 9)
        lib.call("pred_uni_frames_bounded", dptr(got), width, dptr(d0, org), pitch, width, height, taps, xf, yf, nf, width * height, rows * pitch, 0, 0)
        assert np.array_equal(to_host(got), want), (xf, yf)
    want = np.zeros((nf, height, width), np.uint8)
    oracle.drv("pred_bi_frames", ptr(want), width, ptr(host, org), ptr(host1, org), pitch, width, height, taps, 1, 2, nfrac - 1, 0, nf, width * height,
               rows * pitch, threads=4)
    got = dev_full(want.shape, np.uint8, 9)
    lib.call("pred_bi_frames_bounded", dptr(got), width, dptr(d0, org), dptr(d1, org), pitch, width, height, taps, 1, 2, nfrac - 1, 0, nf, width * height,
             rows * pitch, 0, 0)
    assert np.array_equal(to_host(got), want)


@pytest.mark.parametrize("taps", [8, 4])
def test_pu_lists_over_frames(oracle, taps):
    """hevcasm_pred_*_list_frames: the PU lists of several frames in one launch (descriptors with a trailing frame index, frames
    interleaved in the list) = the per-frame list forms of the oracle on each frame"""
    width, height, nf = 512, 420, 3
    r0 = _ref_planes(270 + taps, nf, width, height, pad=32)
    r1 = _ref_planes(271 + taps, nf, width, height, pad=32, kind="smooth")
    per_frame = [_pu_list(taps, width, height, 272 + f) for f in range(nf)]
    for bi in (False, True):
        want = synth.random_planes(273, nf, width, height, 32)
        got = to_dev(want.buf)
        rows = []
        for f, pus in enumerate(per_frame):
            d = np.ascontiguousarray(pus if bi else pus[:, :6])
            wf, rf0, rf1 = want.buf[f], r0.buf[f], r1.buf[f]
            org = want.origin
            if bi:
                oracle.drv("pred_bi_batch", ptr(wf, org), want.pitch, ptr(rf0, r0.origin), ptr(rf1, r1.origin), r0.pitch, taps, ptr(d), len(d), threads=4)
            else:
                oracle.drv("pred_uni_batch", ptr(wf, org), want.pitch, ptr(rf0, r0.origin), r0.pitch, taps, ptr(d), len(d), threads=4)
            rows.append(np.concatenate([d, np.full((len(d), 1), f, np.int16)], axis=1))
        # interleave the frames' lists: PU i of frame 0, PU i of frame 1, ...
        n = min(len(r) for r in rows)
        mixed = np.ascontiguousarray(np.concatenate([np.stack([r[:n] for r in rows], 1).reshape(-1, rows[0].shape[1])] + [r[n:] for r in rows]))
        d0, d1, dp = to_dev(r0.buf), to_dev(r1.buf), to_dev(mixed)
        if bi:
            lib.call("pred_bi_list_frames", dptr(got, want.origin), want.pitch, dptr(d0, r0.origin), dptr(d1, r1.origin), r0.pitch, taps, dptr(dp), len(mixed),
                     want.frame_stride, r0.frame_stride)
        else:
            lib.call("pred_uni_list_frames", dptr(got, want.origin), want.pitch, dptr(d0, r0.origin), r0.pitch, taps, dptr(dp), len(mixed), want.frame_stride,
                     r0.frame_stride)
        assert np.array_equal(to_host(got), want.buf), bi


@pytest.mark.parametrize("taps", [8, 4])
def test_pu_lists_odd_strides(oracle, taps):
    """the PU-list kernels with row strides that are not multiples of 4 (every row of a strip has its own alignment) and odd plane origins"""
    width, height = 512, 420
    pus = _pu_list(taps, width, height, 281)

    def odd(planes, extra, shift):
        """the same planes in a buffer with an odd pitch, origin shifted by `shift` bytes"""
        nf, rows, pitch = planes.buf.shape
        flat = np.zeros(nf * rows * (pitch + extra) + 64, np.uint8)
        view = np.lib.stride_tricks.as_strided(flat[shift:], (nf, rows, pitch), (rows * (pitch + extra), pitch + extra, 1))
        view[...] = planes.buf
        return flat, view, shift + planes.pad * (pitch + extra) + planes.pad, pitch + extra, rows * (pitch + extra)

    r0 = _ref_planes(282 + taps, 1, width, height, pad=32)
    r1 = _ref_planes(283 + taps, 1, width, height, pad=32, kind="smooth")
    w0 = synth.random_planes(284, 1, width, height, 32)
    f0, _, o0, p0, _ = odd(r0, 3, 1)
    f1, _, o1, p1, _ = odd(r1, 3, 1)
    assert p0 == p1 and o0 == o1
    for bi in (False, True):
        fw, vw, ow, pw, _ = odd(w0, 1, 2)
        d = np.ascontiguousarray(pus if bi else pus[:, :6])
        if bi:
            oracle.drv("pred_bi_batch", ptr(fw, ow), pw, ptr(f0, o0), ptr(f1, o1), p0, taps, ptr(d), len(d), threads=4)
        else:
            oracle.drv("pred_uni_batch", ptr(fw, ow), pw, ptr(f0, o0), p0, taps, ptr(d), len(d), threads=4)
        fg = odd(w0, 1, 2)[0]
        dg, d0, d1, dp = to_dev(fg), to_dev(f0), to_dev(f1), to_dev(d)
        if bi:
            lib.call("pred_bi_batch", dptr(dg, ow), pw, dptr(d0, o0), dptr(d1, o1), p0, taps, dptr(dp), len(d))
        else:
            lib.call("pred_uni_batch", dptr(dg, ow), pw, dptr(d0, o0), p0, taps, dptr(dp), len(d))
        assert np.array_equal(to_host(dg), fw), bi
