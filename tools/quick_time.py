"""Scratch timing of individual entry points with CUDA events (development aid; bench.py is the contract)."""
import ctypes as C
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from hevcasm_b200 import lib, synth


def dptr(t, off=0):
    return C.c_void_p(t.data_ptr() + off * t.element_size())


def time_call(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    W, H, NF, PAD = 3840, 2160, 16, 64
    pitch = synth.pitch_for(W, PAD)
    rows = H + 2 * PAD
    g = torch.Generator(device="cuda").manual_seed(1)
    src = torch.randint(0, 256, (NF, rows, pitch), dtype=torch.uint8, device="cuda", generator=g)
    ref = torch.randint(0, 256, (NF, rows, pitch), dtype=torch.uint8, device="cuda", generator=g)
    org = PAD * pitch + PAD
    fs = rows * pitch
    res = {}
    outs = [torch.empty((NF * (W // s) * (H // s) * 64,), dtype=torch.int32, device="cuda") for s in (8, 16, 32, 64)]
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def pyr():
        lib.call("sad_sweep_pyramid_frames", dptr(src, org), pitch, dptr(ref, org), pitch, W, H, -4, -4, NF, fs, fs, *[dptr(o) for o in outs], stream=stream)
    ms = time_call(pyr)
    res["sad_pyramid_ms_per_4k_frame"] = ms / NF
    res["sad_pyramid_Gsamples_s"] = NF * W * H / ms / 1e6
    for s, o in zip((8, 16, 32, 64), outs):
        def one():
            lib.call("sad_sweep_frames", dptr(src, org), pitch, dptr(ref, org), pitch, W, H, (s << 8) | s, -4, -4, 8, 8, NF, fs, fs, dptr(o), stream=stream)
        ms = time_call(one)
        res[f"sad_sweep_{s}x{s}_Gsamples_s"] = NF * W * H / ms / 1e6

    def ssd():
        lib.call("ssd_frames", dptr(src, org), pitch, dptr(ref, org), pitch, W, H, 3, NF, fs, fs, dptr(outs[0]), stream=stream)
    ms = time_call(ssd)
    res["ssd8_Gsamples_s"] = NF * W * H / ms / 1e6
    res["ssd8_GBps"] = NF * W * H * 2 / ms / 1e6

    n = NF * W * H
    q_src = torch.randint(-32768, 32767, (n,), dtype=torch.int16, device="cuda", generator=g)
    q_dst = torch.empty_like(q_src)
    cbf = torch.empty((n // 64,), dtype=torch.int32, device="cuda")

    def quant():
        lib.call("quantize_batch", dptr(q_dst), dptr(q_src), 26214, 18, 171 << 7, 64, n // 64, dptr(cbf), stream=stream)
    ms = time_call(quant)
    res["quantize_Gsamples_s"] = n / ms / 1e6
    res["quantize_GBps"] = n * 4 / ms / 1e6

    def dequant():
        lib.call("quantize_inverse_batch", dptr(q_dst), dptr(q_src), 18432, 6, n, stream=stream)
    ms = time_call(dequant)
    res["dequant_Gsamples_s"] = n / ms / 1e6
    res["dequant_GBps"] = n * 4 / ms / 1e6

    a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    b = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    ms = time_call(lambda: b.copy_(a))
    res["torch_copy_GBps"] = 2 * (1 << 30) / ms / 1e6
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
