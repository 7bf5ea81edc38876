"""Runs one entry point a few times on a 16-frame 4K batch (for ncu captures: `ncu ... python tools/run_one.py <name>`)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hevcasm_b200 import lib, synth
if os.environ.get("RUN_ONE_EXP"):   # experiments build: HEVCASM_* switches and the not-adopted kernel variants
    lib.use_experiments()

W, H, NF, PAD = 3840, 2160, int(os.environ.get("RUN_ONE_NF", "16")), 64
name = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pitch = synth.pitch_for(W, PAD); rows = H + 2 * PAD; org = PAD * pitch + PAD; fs = rows * pitch
g = torch.Generator(device="cuda").manual_seed(7)
a = torch.randint(0, 256, (NF, rows, pitch), dtype=torch.uint8, device="cuda", generator=g)
b = torch.randint(0, 256, (NF, rows, pitch), dtype=torch.uint8, device="cuda", generator=g)
o8 = torch.empty_like(a)
n = NF * W * H
rp = synth.pitch_for(W, 0, 128)
res = torch.randint(-256, 256, (NF, H, rp), dtype=torch.int16, device="cuda", generator=g)
co = torch.randint(-600, 600, (n,), dtype=torch.int16, device="cuda", generator=g)
co2 = torch.empty((n,), dtype=torch.int16, device="cuda")
cbf = torch.empty((n // 16,), dtype=torch.int32, device="cuda")
cbf2 = torch.empty((n // 4,), dtype=torch.int32, device="cuda")
def d(t, off=0): return C.c_void_p(t.data_ptr() + off * t.element_size())
sad_outs = [torch.empty((NF * (W // s) * (H // s) * 64,), dtype=torch.int32, device="cuda") for s in (8, 16, 32, 64)]
calls = {
    "sad_pyr": lambda: lib.call("sad_sweep_pyramid_frames", d(a, org), pitch, d(b, org), pitch, W, H, -4, -4, NF, fs, fs, *[d(o) for o in sad_outs]),
    "sad_pyr8": lambda: lib.call("sad_sweep_pyramid_frames", d(a, org), pitch, d(b, org), pitch, W, H, -4, -4, NF, fs, fs, d(sad_outs[0]), None, None, None),
    "sad_sweep8": lambda: lib.call("sad_sweep_frames", d(a, org), pitch, d(b, org), pitch, W, H, (8 << 8) | 8, -4, -4, 8, 8, NF, fs, fs, d(sad_outs[0])),
    "sad_sweep16": lambda: lib.call("sad_sweep_frames", d(a, org), pitch, d(b, org), pitch, W, H, (16 << 8) | 16, -4, -4, 8, 8, NF, fs, fs, d(sad_outs[1])),
    "sad_pyr16": lambda: lib.call("sad_sweep_pyramid_frames", d(a, org), pitch, d(b, org), pitch, W, H, -4, -4, NF, fs, fs, None, d(sad_outs[1]), None, None),
    "pred_hv": lambda: lib.call("pred_uni_frames", d(o8, org), pitch, d(a, org), pitch, W, H, 8, 1, 3, NF, fs, fs),
    "satd2": lambda: lib.call("hadamard_satd_frames", d(a, org), pitch, d(b, org), pitch, W, H, 1, NF, fs, fs, d(cbf2)),
    "satd4": lambda: lib.call("hadamard_satd_frames", d(a, org), pitch, d(b, org), pitch, W, H, 2, NF, fs, fs, d(cbf)),
    "satd8": lambda: lib.call("hadamard_satd_frames", d(a, org), pitch, d(b, org), pitch, W, H, 3, NF, fs, fs, d(cbf)),
    "pred_copy": lambda: lib.call("pred_uni_frames", d(o8, org), pitch, d(a, org), pitch, W, H, 8, 0, 0, NF, fs, fs),
    "pred_h": lambda: lib.call("pred_uni_frames", d(o8, org), pitch, d(a, org), pitch, W, H, 8, 1, 0, NF, fs, fs),
    "pred_v": lambda: lib.call("pred_uni_frames", d(o8, org), pitch, d(a, org), pitch, W, H, 8, 0, 2, NF, fs, fs),
    "pred_chroma_hv": lambda: lib.call("pred_uni_frames", d(o8, org), pitch, d(a, org), pitch, W, H, 4, 3, 5, NF, fs, fs),
    "pred_bi_avg": lambda: lib.call("pred_bi_frames", d(o8, org), pitch, d(a, org), d(b, org), pitch, W, H, 8, 0, 0, 0, 0, NF, fs, fs),
    "pred_bi": lambda: lib.call("pred_bi_frames", d(o8, org), pitch, d(a, org), d(b, org), pitch, W, H, 8, 1, 2, 3, 1, NF, fs, fs),
    "pred_bi_chroma": lambda: lib.call("pred_bi_frames", d(o8, org), pitch, d(a, org), d(b, org), pitch, W, H, 4, 3, 5, 6, 1, NF, fs, fs),
    "pipe8p": lambda: lib.call("residual_from_planes_pipeline_frames", d(o8, org), pitch, d(co2), d(cbf), d(b, org), pitch, d(a, org), pitch, W, H, 3, 0, 26214, 18, 171 << 7, 18432, 6, NF, fs, fs, fs),
    "fwd8": lambda: lib.call("transform_frames", d(co2), d(res), rp, W, H, 3, 0, NF, H * rp),
    "fwd4": lambda: lib.call("transform_frames", d(co2), d(res), rp, W, H, 2, 0, NF, H * rp),
    "fwd16": lambda: lib.call("transform_frames", d(co2), d(res), rp, W, H, 4, 0, NF, H * rp),
    "fwd32": lambda: lib.call("transform_frames", d(co2), d(res), rp, W, H, 5, 0, NF, H * rp),
    "inv4": lambda: lib.call("inverse_transform_add_frames", d(o8, org), pitch, d(a, org), pitch, d(co), W, H, 2, 0, NF, fs, fs),
    "inv8": lambda: lib.call("inverse_transform_add_frames", d(o8, org), pitch, d(a, org), pitch, d(co), W, H, 3, 0, NF, fs, fs),
    "inv16": lambda: lib.call("inverse_transform_add_frames", d(o8, org), pitch, d(a, org), pitch, d(co), W, H, 4, 0, NF, fs, fs),
    "inv32": lambda: lib.call("inverse_transform_add_frames", d(o8, org), pitch, d(a, org), pitch, d(co), W, H, 5, 0, NF, fs, fs),
    "recon16": lambda: lib.call("quantize_reconstruct_frames", d(o8, org), pitch, d(a, org), pitch, d(co), W, H, 4, NF, fs, fs),
    "recon8": lambda: lib.call("quantize_reconstruct_frames", d(o8, org), pitch, d(a, org), pitch, d(co), W, H, 3, NF, fs, fs),
    "pipe16": lambda: lib.call("residual_pipeline_frames", d(o8, org), pitch, d(co2), d(cbf), d(res), rp, d(a, org), pitch, W, H, 4, 0, 26214, 18, 171 << 7, 18432, 6, NF, fs, H * rp, fs),
    "pipe32": lambda: lib.call("residual_pipeline_frames", d(o8, org), pitch, d(co2), d(cbf), d(res), rp, d(a, org), pitch, W, H, 5, 0, 26214, 18, 171 << 7, 18432, 6, NF, fs, H * rp, fs),
    "pipe8": lambda: lib.call("residual_pipeline_frames", d(o8, org), pitch, d(co2), d(cbf), d(res), rp, d(a, org), pitch, W, H, 3, 0, 26214, 18, 171 << 7, 18432, 6, NF, fs, H * rp, fs),
}
import numpy as np
def pu_list(sz, bi=False, seed=5):
    xs, ys = np.meshgrid(np.arange(W // sz) * sz, np.arange(H // sz) * sz)
    if os.environ.get("RUN_ONE_ORDER") == "ctu":   # CTUs (64x64) in raster order, the PUs of a CTU in z-order - the order a codec produces them in
        def morton(u, v):
            m = np.zeros_like(u)
            for bit in range(4):
                m |= ((u >> bit) & 1) << (2 * bit) | ((v >> bit) & 1) << (2 * bit + 1)
            return m
        key = ((ys // 64) * (W // 64 + 1) + xs // 64) * 256 + morton((xs % 64) // sz, (ys % 64) // sz)
        o = np.argsort(key.reshape(-1), kind="stable")
        xs, ys = xs.reshape(-1)[o], ys.reshape(-1)[o]
    n = xs.size
    r = synth.splitmix64(seed, 4 * n).astype(np.int64)
    span = int(os.environ.get("RUN_ONE_MVSPAN", "64"))   # motion vectors in [-span, span] quarter samples; "0" = all zero; "c<N>" = one vector per 64x64 CTU
    if os.environ.get("RUN_ONE_MVCTU"):
        ctu = (ys.reshape(-1) // 64) * 64 + xs.reshape(-1) // 64
        rc = synth.splitmix64(seed + 1, 2 * 4096).astype(np.int64)
        mvx, mvy = rc[ctu % 4096] % (2 * span + 1) - span, rc[4096 + ctu % 4096] % (2 * span + 1) - span
    else:
        mvx, mvy = r[:n] % (2 * span + 1) - span, r[n:2 * n] % (2 * span + 1) - span
    cols = [xs.reshape(-1), ys.reshape(-1), np.full(n, sz), np.full(n, sz), mvx, mvy]
    if bi:
        cols += [r[2 * n:3 * n] % 129 - 64, r[3 * n:] % 129 - 64]
    return torch.from_numpy(np.stack(cols, -1).astype(np.int16)).cuda()
for sz in (8, 16, 32, 64):
    calls[f"pred_list{sz}"] = (lambda pl: (lambda: lib.call("pred_uni_batch", d(o8, org), pitch, d(a, org), pitch, 8, d(pl), pl.shape[0])))(pu_list(sz))
    calls[f"pred_bilist{sz}"] = (lambda pl: (lambda: lib.call("pred_bi_batch", d(o8, org), pitch, d(a, org), d(b, org), pitch, 8, d(pl), pl.shape[0])))(pu_list(sz, True))
def pu_list_frames(sz, bi=False):
    """the per-frame list of every frame of the batch, each descriptor with its frame index (hevcasm_pred_*_list_frames)"""
    one = pu_list(sz, bi).cpu().numpy()
    rows = [np.concatenate([one, np.full((len(one), 1), f, np.int16)], axis=1) for f in range(NF)]
    return torch.from_numpy(np.ascontiguousarray(np.concatenate(rows))).cuda()
if name.startswith("pred_listf") or name.startswith("pred_bilistf"):
    bi = name.startswith("pred_bilistf")
    plf = pu_list_frames(int(name[len("pred_bilistf" if bi else "pred_listf"):]), bi)
    if bi:
        calls[name] = lambda: lib.call("pred_bi_list_frames", d(o8, org), pitch, d(a, org), d(b, org), pitch, 8, d(plf), plf.shape[0], fs, fs)
    else:
        calls[name] = lambda: lib.call("pred_uni_list_frames", d(o8, org), pitch, d(a, org), pitch, 8, d(plf), plf.shape[0], fs, fs)
for _ in range(iters):
    calls[name]()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    calls[name]()
e1.record(); torch.cuda.synchronize()
print(name, "ms/launch", e0.elapsed_time(e1) / iters)
