"""Pins the oracle (oracle/hevc_oracle.c, our restatement) against the reference's own C path compiled in place
(oracle/_ref, built by oracle/build_ref.sh from /root/reference).  Inputs follow the reference's own self-tests
 and widen them: every fractional position, several quantiser settings, full-range coefficients.
CPU only."""
import ctypes as C

import numpy as np
import pytest

from hevcasm_b200 import synth
from hevcasm_b200.abi import HEVCASM_RECT
from oracle.binding import ptr

# reference sad.c:231-240
PARTITIONS = [(64, 64), (64, 48), (64, 32), (64, 16), (48, 64), (32, 64), (32, 32), (32, 24), (32, 16), (32, 8), (24, 32),
              (16, 64), (16, 32), (16, 16), (16, 12), (16, 8), (16, 4), (12, 16), (8, 32), (8, 16), (8, 8), (8, 4), (4, 8)]


def test_interpolation_tables(oracle, reference):
    for taps, nfrac in ((8, 4), (4, 8)):
        for frac in range(nfrac):
            for k in range(taps):
                assert oracle.blk.pred_coefficient(taps, frac, k) == reference.blk.pred_coefficient(taps, frac, k)


def test_transform_matrices_are_hevc(oracle):
    # spot values of the HEVC core transform (H.265 8.6.4.2) + orthogonality-ish sanity of the generated matrices
    t8 = [[oracle.blk.transform_matrix(0, 8, k, x) for x in range(8)] for k in range(8)]
    assert t8[1] == [89, 75, 50, 18, -18, -50, -75, -89]
    assert t8[2] == [83, 36, -36, -83, -83, -36, 36, 83]
    assert t8[7] == [18, -50, 75, -89, 89, -75, 50, -18]
    t32 = np.array([[oracle.blk.transform_matrix(0, 32, k, x) for x in range(32)] for k in range(32)])
    assert list(t32[:, 0]) == [64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64, 61, 57, 54, 50, 46, 43, 38,
                               36, 31, 25, 22, 18, 13, 9, 4]
    g = t32 @ t32.T
    assert np.all(np.abs(g - np.diag(np.diag(g))) < 600) and np.all(np.abs(np.diag(g) - 32 * 4096) < 600)


@pytest.mark.parametrize("w,h", PARTITIONS)
def test_sad_blocks(oracle, reference, w, h):
    # reference sad.c:243-259: 128x128 rand() arrays, ref at an unaligned offset
    src = synth.random_bytes(1, 128 * 128)
    ref = synth.random_bytes(2, 128 * 128)
    rect = HEVCASM_RECT(w, h)
    a = oracle.blk.sad(ptr(src), 64, ptr(ref, 1 + 128), 64, rect)
    b = reference.blk.sad(ptr(src), 64, ptr(ref, 1 + 128), 64, rect)
    assert a == b
    # 4-way, reference sad.c:317-340 offsets
    offs = [1 + 2 * 128, 2 + 1 * 128, 3 + 2 * 128, 2 + 3 * 128]
    refs = (C.c_void_p * 4)(*[ref.ctypes.data + o for o in offs])
    sa, sb = (C.c_int * 4)(), (C.c_int * 4)()
    oracle.blk.sad_multiref_4(ptr(src), 64, refs, 64, sa, rect)
    reference.blk.sad_multiref_4(ptr(src), 64, refs, 64, sb, rect)
    assert list(sa) == list(sb)


@pytest.mark.parametrize("log2", [2, 3, 4, 5, 6])
def test_ssd_blocks(oracle, reference, log2):
    n = 1 << log2
    a = synth.random_bytes(3, 128 * 64)
    b = synth.random_bytes(4, 128 * 64)
    assert oracle.blk.ssd(ptr(a), 2 * n, ptr(b), 2 * n, log2) == reference.blk.ssd(ptr(a), 2 * n, ptr(b), 2 * n, log2)
    z = np.zeros(64 * 128, np.uint8)
    o = np.full(64 * 128, 255, np.uint8)
    assert oracle.blk.ssd(ptr(z), 64, ptr(o), 64, 6) == 64 * 64 * 255 * 255 == reference.blk.ssd(ptr(z), 64, ptr(o), 64, 6)


@pytest.mark.parametrize("taps", [8, 4])
def test_pred_uni_all_fractions(oracle, reference, taps):
    # reference pred_inter.c:461-487 uses an 80x192 buffer and fractions {0,1}; here every fraction, three shapes
    pitch = 192
    ref = synth.random_bytes(5, 96 * pitch)
    nfrac = 4 if taps == 8 else 8
    shapes = [(taps * 8, taps * 8), (taps * 3, taps * 2), (taps, taps // 2 + 2)]
    for w, h in shapes:
        for yf in range(nfrac):
            for xf in range(nfrac):
                da = np.zeros(80 * pitch, np.uint8)
                db = np.zeros(80 * pitch, np.uint8)
                oracle.blk.pred_uni(ptr(da), pitch, ptr(ref, 8 * pitch + 16), pitch, taps, w, h, xf, yf)
                reference.blk.pred_uni(ptr(db), pitch, ptr(ref, 8 * pitch + 16), pitch, taps, w, h, xf, yf)
                A = da.reshape(80, pitch)[:h, :w]
                B = db.reshape(80, pitch)[:h, :w]
                assert np.array_equal(A, B), (w, h, xf, yf)


@pytest.mark.parametrize("taps", [8, 4])
def test_pred_bi_fractions(oracle, reference, taps):
    pitch = 192
    r0 = synth.random_bytes(6, 96 * pitch)
    r1 = synth.random_bytes(7, 96 * pitch)
    nfrac = 4 if taps == 8 else 8
    fr = synth.splitmix64(8, 4 * 24) % np.uint64(nfrac)
    cases = [tuple(int(v) for v in fr[4 * i:4 * i + 4]) for i in range(24)] + [(f, f, f, f) for f in range(nfrac)] + [(0, 0, 0, 0)]
    for (w, h) in [(taps * 8, taps * 8), (taps * 2, taps * 3)]:
        for x0, y0, x1, y1 in cases:
            da = np.zeros(80 * pitch, np.uint8)
            db = np.zeros(80 * pitch, np.uint8)
            args = (pitch, ptr(r0, 8 * pitch + 16), ptr(r1, 8 * pitch + 16), pitch, taps, w, h, x0, y0, x1, y1)
            oracle.blk.pred_bi(ptr(da), *args)
            reference.blk.pred_bi(ptr(db), *args)
            assert np.array_equal(da.reshape(80, pitch)[:h, :w], db.reshape(80, pitch)[:h, :w]), (w, h, x0, y0, x1, y1)


def test_pred_bi_extremes(oracle, reference):
    # checkerboards of 0/255 push the bi-pred intermediates to their extremes (SURVEY.md divergence table)
    pitch = 192
    yy, xx = np.mgrid[0:96, 0:pitch]
    for pat in (((xx + yy) & 1) * 255, ((xx & 1) * 255), np.full((96, pitch), 255), np.zeros((96, pitch))):
        r = pat.astype(np.uint8).reshape(-1)
        for taps, nfrac in ((8, 4), (4, 8)):
            for f in range(nfrac):
                da = np.zeros(80 * pitch, np.uint8)
                db = np.zeros(80 * pitch, np.uint8)
                args = (pitch, ptr(r, 8 * pitch + 16), ptr(r, 8 * pitch + 17), pitch, taps, 16, 16, f, f, (f + 1) % nfrac, f)
                oracle.blk.pred_bi(ptr(da), *args)
                reference.blk.pred_bi(ptr(db), *args)
                assert np.array_equal(da, db)


TR = [(1, 2), (0, 2), (0, 3), (0, 4), (0, 5)]


@pytest.mark.parametrize("trType,log2", TR)
def test_forward_transform(oracle, reference, trType, log2):
    n = 1 << log2
    for seed, lo, hi in ((9, -256, 255), (10, -32768, 32767), (11, -1, 1)):
        for rep in range(8):
            src = synth.random_int16(seed * 100 + rep, 32 * 40, lo, hi)
            ca = np.zeros(n * n, np.int16)
            cb = np.zeros(n * n, np.int16)
            oracle.blk.transform(ptr(ca), ptr(src, 3), 40, trType, log2)
            reference.blk.transform(ptr(cb), ptr(src, 3), 40, trType, log2)
            assert np.array_equal(ca, cb)


@pytest.mark.parametrize("trType,log2", TR)
def test_inverse_transform_add(oracle, reference, trType, log2):
    n = 1 << log2
    for seed, lo, hi in ((12, -32768, 32767), (13, -512, 511), (14, 32767, 32767), (15, -32768, -32768)):
        for rep in range(8):
            co = synth.random_int16(seed * 100 + rep, n * n, lo, hi)
            if rep == 7:  # sparse: DC + one AC
                co[:] = 0
                co[0], co[n + 1] = 1000, -777
            pred = synth.random_bytes(seed * 100 + rep + 50, 32 * 48)
            da = np.zeros(32 * 48, np.uint8)
            db = np.zeros(32 * 48, np.uint8)
            oracle.blk.inverse_transform_add(ptr(da), 48, ptr(pred), 48, ptr(co), trType, log2)
            reference.blk.inverse_transform_add(ptr(db), 48, ptr(pred), 48, ptr(co), trType, log2)
            assert np.array_equal(da, db)


QUANT = [(51, 20, 14), (26214, 18, 171 << 7), (14564, 26, 85 << 7), (32767, 16, 32767), (1, 27, 0), (16384, 21, 171 << 7)]
DEQUANT = [(51, 14), (18432, 6), (640, 5), (32767, 15), (1, 1), (40 << 4, 6)]


@pytest.mark.parametrize("scale,shift,offset", QUANT)
def test_quantize(oracle, reference, scale, shift, offset):
    for n in (16, 64, 256, 1024):
        src = synth.random_int16(16 + n, n)
        src[:4] = [-32768, 32767, 0, -1]
        da, db = np.zeros(n, np.int16), np.zeros(n, np.int16)
        ra = oracle.blk.quantize(ptr(da), ptr(src), scale, shift, offset, n)
        rb = reference.blk.quantize(ptr(db), ptr(src), scale, shift, offset, n)
        assert np.array_equal(da, db) and ra == rb
    z = np.zeros(16, np.int16)
    assert oracle.blk.quantize(ptr(z.copy()), ptr(z), scale, shift, 0, 16) == 0


@pytest.mark.parametrize("scale,shift", DEQUANT)
def test_quantize_inverse(oracle, reference, scale, shift):
    for n in (16, 64, 256, 1024):
        src = synth.random_int16(17 + n, n)
        da, db = np.zeros(n, np.int16), np.zeros(n, np.int16)
        oracle.blk.quantize_inverse(ptr(da), ptr(src), scale, shift, n)
        reference.blk.quantize_inverse(ptr(db), ptr(src), scale, shift, n)
        assert np.array_equal(da, db)


@pytest.mark.parametrize("log2", [2, 3, 4, 5])
def test_quantize_reconstruct(oracle, reference, log2):
    n = 1 << log2
    pred = synth.random_bytes(18, 32 * 32)
    for lo, hi in ((-256, 255), (-32768, 32767)):
        res = synth.random_int16(19 + log2, n * n, lo, hi)
        da, db = np.zeros(32 * 32, np.uint8), np.zeros(32 * 32, np.uint8)
        oracle.blk.quantize_reconstruct(ptr(da), 32, ptr(pred), 32, ptr(res), log2)
        reference.blk.quantize_reconstruct(ptr(db), 32, ptr(pred), 32, ptr(res), log2)
        assert np.array_equal(da, db)


@pytest.mark.parametrize("log2", [1, 2, 3])
def test_hadamard_satd(oracle, reference, log2):
    # reference hadamard.c:200-230 tests 8x8 arrays of rand() bytes; here: many positions, strides and the extremes
    a = synth.random_bytes(20, 96 * 64)
    b = synth.random_bytes(21, 96 * 64)
    for off in (0, 1, 7 + 64, 33 + 5 * 64, 50 + 80 * 64):
        for stride in (64, 96):
            assert oracle.blk.hadamard_satd(ptr(a, off), stride, ptr(b, off + 2), stride, log2) == \
                reference.blk.hadamard_satd(ptr(a, off), stride, ptr(b, off + 2), stride, log2)
    z, f = np.zeros(64 * 64, np.uint8), np.full(64 * 64, 255, np.uint8)
    n = 1 << log2
    assert oracle.blk.hadamard_satd(ptr(z), 64, ptr(f), 64, log2) == reference.blk.hadamard_satd(ptr(z), 64, ptr(f), 64, log2) == (n // 4 + 255 * n * n) // (n // 2)
    chk = ((np.arange(64)[:, None] + np.arange(64)[None, :]) & 1).astype(np.uint8).reshape(-1) * 255
    assert oracle.blk.hadamard_satd(ptr(chk), 64, ptr(z), 64, log2) == reference.blk.hadamard_satd(ptr(chk), 64, ptr(z), 64, log2)


def test_ssd_linear(oracle, reference):
    a = synth.random_bytes(22, 4096)
    b = synth.random_bytes(23, 4096)
    for size in (0, 1, 15, 16, 64, 333, 512, 4000):
        assert oracle.blk.ssd_linear(ptr(a, 3), ptr(b, 5), size) == reference.blk.ssd_linear(ptr(a, 3), ptr(b, 5), size)
