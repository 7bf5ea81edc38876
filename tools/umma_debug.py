"""Debug helper: tensor-core two-pass interpolation vs the streaming kernel on one small plane batch; prints where they differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from hevcasm_b200 import lib, synth
from gpu_util import to_dev, dptr, to_host

width, height, nf, taps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ref = synth.random_planes(208, nf, width, height, 16)
dr = to_dev(ref.buf)
def run(xf, yf):
    o = synth.random_planes(201, nf, width, height, 16)
    g = to_dev(o.buf)
    lib.call("pred_uni_frames", dptr(g, o.origin), o.pitch, dptr(dr, ref.origin), ref.pitch, width, height, taps, xf, yf, nf, o.frame_stride, ref.frame_stride)
    return to_host(g), o
for xf, yf in [(1, 1), (2, 3)]:
    os.environ.pop("HEVCASM_PRED_HV", None)
    a, o = run(xf, yf)
    os.environ["HEVCASM_PRED_HV"] = "umma"
    b, _ = run(xf, yf)
    d = np.flatnonzero(a != b)
    print("frac", xf, yf, "pitch", o.pitch, "origin", o.origin, "fs", o.frame_stride, "mismatches", d.size)
    if d.size:
        rel = d - o.origin
        fr = d // o.frame_stride
        rr = (d - fr * o.frame_stride - o.origin)
        ys, xs = rr // o.pitch, rr % o.pitch
        print(" frames", np.unique(fr), "y range", ys.min(), ys.max(), "x range", xs.min(), xs.max())
        print(" first", [(int(f), int(y), int(x), int(a[i]), int(b[i])) for f, y, x, i in list(zip(fr, ys, xs, d))[:12]])
        yy, cnt = np.unique(ys, return_counts=True)
        print(" rows with mismatches", list(zip(yy.tolist(), cnt.tolist()))[:40])
