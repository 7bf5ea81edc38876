"""Golden cases: batched entry points on seeded synthetic planes.  Each case returns the SHA-256 of its output buffer(s).

One definition, three backends (see Backend): the reference's own compiled C path (generates tests/golden/golden.json -
make_golden.py, run where /root/reference exists), the CPU oracle (checked on every CPU run) and the CUDA library through
its C ABI (checked by the -m gpu tests).  The committed digests travel to the GPU box; /root/reference does not."""
import ctypes as C
import hashlib

import numpy as np

from hevcasm_b200 import synth
from hevcasm_b200.abi import HEVCASM_RECT

W, H, NF, PAD = 200, 136, 2, 16


class CpuBackend:
    """oracle.binding.CpuLib (oracle or reference): host arrays, drv(name, ...)"""

    def __init__(self, cpu):
        self.cpu = cpu

    def buf(self, a):
        return a

    def ptr(self, h, off=0):
        return C.c_void_p(h.ctypes.data + off * h.itemsize)

    def call(self, name, *args):
        self.cpu.drv(name, *args, threads=4)

    def get(self, h):
        return h


class GpuBackend:
    """libhevcasm_b200.so: device tensors, lib.call(name, ...)"""

    def __init__(self):
        import torch
        from hevcasm_b200 import lib
        self.torch, self.lib = torch, lib

    def buf(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).cuda()

    def ptr(self, h, off=0):
        return C.c_void_p(h.data_ptr() + off * h.element_size())

    def call(self, name, *args):
        self.lib.call(name, *args)

    def get(self, h):
        self.torch.cuda.synchronize()
        return h.cpu().numpy()


def _sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _planes():
    src = synth.smooth_planes(901, NF, W, H, PAD)
    ref = synth.smooth_planes(901, NF, W, H, PAD, shift=(3, -2), noise=6)
    rnd = synth.random_planes(902, NF, W, H, PAD)
    return src, ref, rnd


def cases():
    """yields (name, fn(backend) -> digest)"""
    src, ref, rnd = _planes()

    for (w, h) in ((64, 64), (32, 24), (16, 16), (12, 16), (8, 8), (4, 8)):
        def sad(b, w=w, h=h):
            n = NF * (W // w) * (H // h) * 64
            out = b.buf(np.zeros(n, np.int32))
            ds, dr = b.buf(src.buf), b.buf(ref.buf)
            b.call("sad_sweep_frames", b.ptr(ds, src.origin), src.pitch, b.ptr(dr, ref.origin), ref.pitch, W, H, HEVCASM_RECT(w, h), -4, -4, 8, 8, NF,
                   src.frame_stride, ref.frame_stride, b.ptr(out))
            return _sha(b.get(out))
        yield f"sad_sweep_{w}x{h}", sad

    for log2 in (2, 3, 4, 5, 6):
        def ssd(b, log2=log2):
            n = NF * (W >> log2) * (H >> log2)
            out = b.buf(np.zeros(n, np.int32))
            da, db = b.buf(rnd.buf), b.buf(ref.buf)
            b.call("ssd_frames", b.ptr(da, rnd.origin), rnd.pitch, b.ptr(db, ref.origin), ref.pitch, W, H, log2, NF, rnd.frame_stride, ref.frame_stride, b.ptr(out))
            return _sha(b.get(out))
        yield f"ssd_{1 << log2}", ssd

    for log2 in (1, 2, 3):
        def satd(b, log2=log2):
            n = NF * (W >> log2) * (H >> log2)
            out = b.buf(np.zeros(n, np.int32))
            da, db = b.buf(rnd.buf), b.buf(ref.buf)
            b.call("hadamard_satd_frames", b.ptr(da, rnd.origin), rnd.pitch, b.ptr(db, ref.origin), ref.pitch, W, H, log2, NF, rnd.frame_stride, ref.frame_stride,
                   b.ptr(out))
            return _sha(b.get(out))
        yield f"hadamard_satd_{1 << log2}", satd

    def ssd_linear(b):
        digests = []
        da, db = b.buf(rnd.buf), b.buf(ref.buf)
        for size, runs in ((16, 40), (333, 25), (512, 30)):
            out = b.buf(np.zeros(runs, np.int32))
            b.call("ssd_linear_batch", b.ptr(da, rnd.origin), rnd.pitch, b.ptr(db, ref.origin + 3), ref.pitch, size, runs, b.ptr(out))
            digests.append(_sha(b.get(out)))
        return _sha(np.frombuffer("".join(digests).encode(), np.uint8))
    yield "ssd_linear", ssd_linear

    for taps, nfrac in ((8, 4), (4, 8)):
        def uni(b, taps=taps, nfrac=nfrac):
            dr = b.buf(rnd.buf)
            digests = []
            for yf in range(nfrac):
                for xf in range(nfrac):
                    dst = b.buf(np.zeros_like(rnd.buf))
                    b.call("pred_uni_frames", b.ptr(dst, rnd.origin), rnd.pitch, b.ptr(dr, rnd.origin), rnd.pitch, W, H, taps, xf, yf, NF, rnd.frame_stride,
                           rnd.frame_stride)
                    digests.append(_sha(b.get(dst)))
            return _sha(np.frombuffer("".join(digests).encode(), np.uint8))
        yield f"pred_uni_{taps}tap_all_fractions", uni

        def bi(b, taps=taps, nfrac=nfrac):
            d0, d1 = b.buf(rnd.buf), b.buf(ref.buf)
            digests = []
            for c in ((0, 0, 0, 0), (1, 2, 3, 1), (nfrac - 1, 0, 0, nfrac - 1), (2, 2, 1, 3), (0, 1, nfrac - 2, 0)):
                dst = b.buf(np.zeros_like(rnd.buf))
                b.call("pred_bi_frames", b.ptr(dst, rnd.origin), rnd.pitch, b.ptr(d0, rnd.origin), b.ptr(d1, ref.origin), rnd.pitch, W, H, taps, *c, NF,
                       rnd.frame_stride, rnd.frame_stride)
                digests.append(_sha(b.get(dst)))
            return _sha(np.frombuffer("".join(digests).encode(), np.uint8))
        yield f"pred_bi_{taps}tap", bi

    res = synth.residual_planes(903, NF, W, H)
    full = synth.Planes(synth.random_int16(904, res.buf.size).reshape(res.buf.shape), W, H, 0)
    for trType, log2 in ((1, 2), (0, 2), (0, 3), (0, 4), (0, 5)):
        n = 1 << log2
        nb = NF * (W // n) * (H // n)

        def fwd(b, trType=trType, log2=log2, nb=nb, n=n):
            digests = []
            for r in (res, full):
                dres = b.buf(r.buf)
                co = b.buf(np.zeros(nb * n * n, np.int16))
                b.call("transform_frames", b.ptr(co), b.ptr(dres, r.origin), r.pitch, W, H, log2, trType, NF, r.frame_stride)
                digests.append(_sha(b.get(co)))
            return _sha(np.frombuffer("".join(digests).encode(), np.uint8))
        yield f"transform_{'dst' if trType else 'dct'}{n}", fwd

        def inv(b, trType=trType, log2=log2, nb=nb, n=n):
            digests = []
            for seed, lo, hi in ((905, -32768, 32767), (906, -700, 700)):
                co = synth.random_int16(seed + log2, nb * n * n, lo, hi)
                dst, dp, dc = b.buf(np.zeros_like(rnd.buf)), b.buf(rnd.buf), b.buf(co)
                b.call("inverse_transform_add_frames", b.ptr(dst, rnd.origin), rnd.pitch, b.ptr(dp, rnd.origin), rnd.pitch, b.ptr(dc), W, H, log2, trType, NF,
                       rnd.frame_stride, rnd.frame_stride)
                digests.append(_sha(b.get(dst)))
            return _sha(np.frombuffer("".join(digests).encode(), np.uint8))
        yield f"inverse_transform_add_{'dst' if trType else 'dct'}{n}", inv

    coef = synth.random_int16(907, 64 * 1024)
    coef[coef == -32768] = -32767
    coef[4096:8192] = 0
    for scale, shift, offset in ((51, 20, 14), (26214, 18, 171 << 7), (14564, 26, 85 << 7)):
        def quant(b, scale=scale, shift=shift, offset=offset):
            digests = []
            for npb in (16, 64, 256, 1024):
                nblk = coef.size // npb
                dst, dsrc, cbf = b.buf(np.zeros_like(coef)), b.buf(coef), b.buf(np.zeros(nblk, np.int32))
                b.call("quantize_batch", b.ptr(dst), b.ptr(dsrc), scale, shift, offset, npb, nblk, b.ptr(cbf))
                digests.append(_sha(b.get(dst), (b.get(cbf) != 0).astype(np.uint8)))  # only the truth value of cbf is contractual
            return _sha(np.frombuffer("".join(digests).encode(), np.uint8))
        yield f"quantize_{scale}_{shift}", quant
    for scale, shift in ((51, 14), (18432, 6), (640, 5)):
        def dequant(b, scale=scale, shift=shift):
            dst, dsrc = b.buf(np.zeros_like(coef)), b.buf(coef)
            b.call("quantize_inverse_batch", b.ptr(dst), b.ptr(dsrc), scale, shift, coef.size)
            return _sha(b.get(dst))
        yield f"quantize_inverse_{scale}_{shift}", dequant
    for log2 in (2, 3, 4, 5):
        def recon(b, log2=log2):
            n = 1 << log2
            nb = NF * (W // n) * (H // n)
            r = synth.random_int16(908 + log2, nb * n * n, -300, 300)
            dst, dp, dr = b.buf(np.zeros_like(rnd.buf)), b.buf(rnd.buf), b.buf(r)
            b.call("quantize_reconstruct_frames", b.ptr(dst, rnd.origin), rnd.pitch, b.ptr(dp, rnd.origin), rnd.pitch, b.ptr(dr), W, H, log2, NF, rnd.frame_stride,
                   rnd.frame_stride)
            return _sha(b.get(dst))
        yield f"quantize_reconstruct_{1 << log2}", recon
