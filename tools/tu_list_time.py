"""Times the transform-unit list forms class by class (16 4K frames, bench.py's quad-tree tiling): python tools/tu_list_time.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from hevcasm_b200 import lib, synth
from oracle.binding import ptr

W, H, NF, PAD = 3840, 2160, 16, 64
pitch = synth.pitch_for(W, PAD); rows = H + 2 * PAD; org = PAD * pitch + PAD; fs = rows * pitch
g = torch.Generator(device="cuda").manual_seed(7)
a = torch.randint(0, 256, (NF, rows, pitch), dtype=torch.uint8, device="cuda", generator=g)
o8 = torch.empty_like(a)
rp = synth.pitch_for(W, 0, 128)
res = torch.randint(-256, 256, (NF, H, rp), dtype=torch.int16, device="cuda", generator=g)
n = NF * W * H
co = torch.randint(-600, 600, (n,), dtype=torch.int16, device="cuda", generator=g)
co2 = torch.empty((n,), dtype=torch.int16, device="cuda")
def d(t, off=0): return C.c_void_p(t.data_ptr() + off * t.element_size())
tus, counts, covered = bench.tu_buckets(synth, torch, W, H, NF)
start = np.concatenate([[0], np.cumsum(counts)])
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
names = ["4x4 DST", "4x4", "8x8", "16x16", "32x32"]
for c in range(5):
    only = np.zeros(5, np.int32); only[c] = counts[c]
    sub = tus[start[c]:start[c + 1]].contiguous()
    samples = int(counts[c]) * (16, 16, 64, 256, 1024)[c]
    tf = timeit(lambda: lib.call("transform_list_frames", d(co2), d(res), rp, d(sub), ptr(only), H * rp))
    ti = timeit(lambda: lib.call("inverse_transform_add_list_frames", d(o8, org), pitch, d(a, org), pitch, d(co), d(sub), ptr(only), fs, fs))
    print(f"{names[c]:8s} {counts[c]:8d} TUs {samples/1e6:7.1f} Msamples  forward {tf:7.1f} us ({samples*4/tf/1e3/6460.5:.2f} HBM)  inverse {ti:7.1f} us ({samples*4/ti/1e3/6460.5:.2f})")
tf = timeit(lambda: lib.call("transform_list_frames", d(co2), d(res), rp, d(tus), ptr(counts), H * rp))
ti = timeit(lambda: lib.call("inverse_transform_add_list_frames", d(o8, org), pitch, d(a, org), pitch, d(co), d(tus), ptr(counts), fs, fs))
print(f"all      {int(counts.sum()):8d} TUs {covered/1e6:7.1f} Msamples  forward {tf:7.1f} us  inverse {ti:7.1f} us")
