#!/usr/bin/env python
"""bench.py - the driver's measurement contract for hevcasm_b200.

Two workloads (BASELINE.json configs):

  --config 4k_sad (default, the headline - configs[1]): the 4K motion-estimation SAD sweep - every 8x8, 16x16, 32x32 and 64x64 PU
      of a batch of 3840x2160 8-bit frames against 64 candidate vectors (dx, dy in [-4, 3]^2), all four PU sizes from one pass over
      the frames (hevcasm_sad_sweep_pyramid_frames).  A "step" is one such pass over `--frames` frames per GPU ("scaling": "weak").
  --config 8k64 (configs[4]): a FIXED batch of 64 frames of 7680x4320 dealt to the ranks (hevcasm_b200.shard, "scaling": "strong");
      a step takes every frame through the SAD sweep, one two-pass luma interpolation and the fused 8x8 residual pipeline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config ...]        # this repo's CUDA path (one JSON line on rank 0)
    python bench.py --impl reference [...]                                     # the reference's own C path on the host cores

value    : Gsamples/s with the inputs resident in HBM (CUDA events on the launching stream, max over ranks)
e2e      : the same through the host-memory C-ABI calls (hevcasm_*_host): page-locked host frames in, results back in host memory,
           copies inside the timed region
roofline : HBM roofline of the dominant kernel from its algorithmic bytes, DRAM traffic from the committed ncu capture
kernels  : every other kernel of the path against its own roofline (median and best of >= 100 launches), each with the reference's
           C path on the host cores beside it (cpu)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

W4K, H4K, W8K, H8K, PAD = 3840, 2160, 7680, 4320, 64
PAD_TRAVEL = 16   # samples of padding the host forms copy with each plane: the window [-4, 3] needs 4, the interpolation kernels 16
SAD_OUT_BYTES = 4.0 * 64 * (1 / 64 + 1 / 256 + 1 / 1024 + 1 / 4096)            # int32 outputs of the 4 levels, per sample
SAD_OUT_BYTES_PACKED = 64 * (2 / 64 + 2 / 256 + 4 / 1024 + 4 / 4096)           # uint16 for 8x8 / 16x16, int32 for 32x32 / 64x64
SAD_BYTES_PER_SAMPLE = 2.0 + SAD_OUT_BYTES                                      # + src + ref
PIPE8_BYTES_PER_SAMPLE = 2 + 1 + 2 + 1 + 4 / 64                                 # residual + pred + levels + rec + cbf
METRIC = "Gsamples/s (4K ME SAD sweep: 8x8..64x64 PUs x 64 candidate vectors)"
METRIC_8K = "Gsamples/s (8K x 64 frames: SAD sweep + luma HV interpolation + 8x8 residual pipeline, end to end per sample)"
QP = (26214, 18, 171 << 7, 18432, 6)                                            # q_scale, q_shift, q_offset, iq_scale, iq_shift


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="4k_sad", choices=["4k_sad", "8k64"])
    ap.add_argument("--frames", type=int, default=None, help="4k_sad: 4K frames per GPU per step (default 32); 8k64: size of the fixed batch (default 64)")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel table")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--kernel-launches", type=int, default=200, help="launches per entry of the kernel table (groups of 10; >= 100 for reported numbers, "
                    "small for an ncu launch list)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_counters():
    """per-kernel counters of the committed ncu captures (profiles/r02_kernel_counters.json, written by tools/ncu_counters.py)"""
    p = os.path.join(ROOT, "profiles", "r02_kernel_counters.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


# ---------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (pynvml; nvidia-smi fallback)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        for n in dir(nv):
            if n.startswith("nvmlClocksEventReason") or n.startswith("nvmlClocksThrottleReason"):
                v = getattr(nv, n)
                if isinstance(v, int) and v and (v & (v - 1)) == 0:
                    names.setdefault(v, n.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit and name not in ("GpuIdle", "None", "ApplicationsClocksSetting"):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if self.nv is None or not self.samples:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = [float(v) for v in out.strip().split(",")]
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "how": "nvidia-smi after the run"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "how": "unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz), "reasons": sorted(self.reasons),
                "samples": len(self.samples), "how": "pynvml during the timed region"}


# ---------------------------------------------------------------------------------------------- the reference's C path on the host

def cpu_library():
    from oracle import binding
    ref = binding.reference()
    if ref is not None:
        ref.lib.ref_drv_set_avx2_sad(1)   # libvpx AVX2 intrinsics for the 32x32 / 64x64 four-way SAD (the reference's own SIMD)
        return ref, "reference"
    return binding.oracle(), "port"


CPU_NOTE = "reference C path at -O3 -mavx2 (+ libvpx AVX2 intrinsics for the 32x32/64x64 four-way SAD); the x86 asm needs yasm/nasm, absent here"


def cpu_sad_sweep(cpu, src, ref, width, height, n_frames, threads):
    """The SAD workload on the CPU: four sweeps (8, 16, 32, 64) of 64 candidates through the 4-way SAD function of `cpu`
    (oracle/binding.CpuLib).  Returns (seconds, outputs)."""
    from oracle.binding import ptr
    outs = [np.empty((n_frames * (width // s) * (height // s) * 64,), np.int32) for s in (8, 16, 32, 64)]
    t0 = time.perf_counter()
    for s, o in zip((8, 16, 32, 64), outs):
        cpu.drv("sad_sweep_frames", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, width, height, (s << 8) | s, -4, -4,
                8, 8, n_frames, src.frame_stride, ref.frame_stride, ptr(o), threads=threads)
    return time.perf_counter() - t0, outs


def cpu_8k_frame(cpu, planes, threads):
    """configs[4] for the frames of `planes` on the CPU: SAD sweep (4 sizes x 64 candidates), one HV luma interpolation, 8x8 forward DCT ->
    quantize -> dequantize -> inverse + add.  Returns seconds."""
    from oracle.binding import ptr
    src, ref, res, pred = planes
    W, H, nf = src.width, src.height, src.n_frames
    t, _ = cpu_sad_sweep(cpu, src, ref, W, H, nf, threads)
    out = np.empty_like(src.buf)
    nb = (W // 8) * (H // 8) * nf
    co, lv, dq = (np.empty(nb * 64, np.int16) for _ in range(3))
    cbf = np.empty(nb, np.int32)
    t0 = time.perf_counter()
    cpu.drv("pred_uni_frames", ptr(out, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, W, H, 8, 1, 3, nf, src.frame_stride, ref.frame_stride, threads=threads)
    cpu.drv("transform_frames", ptr(co), ptr(res.buf, res.origin), res.pitch, W, H, 3, 0, nf, res.frame_stride, threads=threads)
    cpu.drv("quantize_batch", ptr(lv), ptr(co), QP[0], QP[1], QP[2], 64, nb, ptr(cbf), threads=threads)
    cpu.drv("quantize_inverse_batch", ptr(dq), ptr(lv), QP[3], QP[4], nb * 64, threads=threads)
    cpu.drv("inverse_transform_add_frames", ptr(out, src.origin), src.pitch, ptr(pred.buf, pred.origin), pred.pitch, ptr(dq), W, H, 3, 0, nf, src.frame_stride,
            pred.frame_stride, threads=threads)
    return t + time.perf_counter() - t0


def host_planes_8k(synth, seed, nf):
    return (synth.random_planes(seed, nf, W8K, H8K, PAD), synth.random_planes(seed + 1, nf, W8K, H8K, PAD), synth.residual_planes(seed + 2, nf, W8K, H8K),
            synth.random_planes(seed + 3, nf, W8K, H8K, PAD))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hevcasm_b200 import synth
    cpu, kind = cpu_library()
    threads = os.cpu_count() or 1
    if args.config == "8k64":
        planes = host_planes_8k(synth, synth.SEED, 1)
        for _ in range(max(args.warmup, 1)):
            cpu_8k_frame(cpu, planes, threads)
        t = sum(cpu_8k_frame(cpu, planes, threads) for _ in range(args.steps))
        value = args.steps * W8K * H8K / t / 1e9
        sample = "1 of the 64 8K frames per step: SAD sweep (4 PU sizes x 64 candidates) + luma HV interpolation + 8x8 DCT / quant / dequant / inverse+add"
        metric, workload = METRIC_8K, "8K x 64 frames, SAD + interpolation + residual pipeline (BASELINE configs[4])"
    else:
        nf = 2
        src = synth.random_planes(synth.SEED, nf, W4K, H4K, PAD)
        ref = synth.random_planes(synth.SEED + 1, nf, W4K, H4K, PAD)
        for _ in range(max(args.warmup, 1)):
            cpu_sad_sweep(cpu, src, ref, W4K, H4K, 1, threads)
        t = sum(cpu_sad_sweep(cpu, src, ref, W4K, H4K, nf, threads)[0] for _ in range(args.steps))
        value = args.steps * nf * W4K * H4K / t / 1e9
        sample = f"{nf} 4K frames per step, all four PU sizes, 64 candidates as 16 four-way calls per PU"
        metric, workload = METRIC, "4K SAD sweep, 8x8/16x16/32x32/64x64 PUs x 64 candidates (BASELINE configs[1])"
    emit({
        "impl": "reference", "metric": metric, "value": value, "unit": "Gsamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong" if args.config == "8k64" else "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload, "sample": sample, "host_threads": threads},
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": threads, "kind": kind, "sample": sample, "note": CPU_NOTE},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ---------------------------------------------------------------------------------------------- GPU arm

def dptr(t, off=0):
    return C.c_void_p(t.data_ptr() + off * t.element_size())


def hptr(a, off=0):
    return C.c_void_p(a.ctypes.data + off * a.itemsize)


def time_kernel(torch, fn, groups=20, per_group=10, warm=3):
    """groups x per_group launches on the current stream, one CUDA-event bracket per group: median / best / mean ms per launch"""
    for _ in range(warm):
        fn()
    # the entry before this one may have left the GPU idle for a second (its CPU baseline): keep launching until ~30 ms of GPU work have
    # run, so that the timed groups start from ramped clocks and a warm instruction cache
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.03:
        for _ in range(per_group):
            fn()
        torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(groups)]
    torch.cuda.synchronize()
    for e0, e1 in ev:
        e0.record()
        for _ in range(per_group):
            fn()
        e1.record()
    torch.cuda.synchronize()
    ms = np.array([e0.elapsed_time(e1) / per_group for e0, e1 in ev])
    return {"ms": float(np.median(ms)), "ms_best": float(ms.min()), "ms_mean": float(ms.mean()), "launches": groups * per_group}


def pu_list(synth, torch, width, height, sizes, nf, bi=False, seed=5):
    """a list of PUs tiling `nf` frames with squares of the given sizes (one size per 64-row band, cycling), random quarter-sample motion
    vectors within +-16 samples: descriptors {x, y, w, h, mvx, mvy[, mvx1, mvy1], frame} as int16"""
    rows = []
    for f in range(nf):
        for band, y0 in enumerate(range(0, height - 63, 64)):
            sz = sizes[(band + f) % len(sizes)]
            xs, ys = np.meshgrid(np.arange(width // sz) * sz, y0 + np.arange(64 // sz) * sz)
            n = xs.size
            rows.append(np.stack([xs.reshape(-1), ys.reshape(-1), np.full(n, sz), np.full(n, sz), np.full(n, f)], -1))
    a = np.concatenate(rows)
    n = len(a)
    r = synth.splitmix64(seed, 4 * n).astype(np.int64)
    mv = [r[k * n:(k + 1) * n] % 129 - 64 for k in range(4 if bi else 2)]
    cols = [a[:, 0], a[:, 1], a[:, 2], a[:, 3]] + mv + [a[:, 4]]
    return torch.from_numpy(np.stack(cols, -1).astype(np.int16)).cuda(), int((a[:, 2] * a[:, 3]).sum())


def tu_buckets(synth, torch, width, height, nf, seed=9):
    """a tiling of `nf` frames into transform units: every 64x64 CTU holds TUs of one size class - 4x4 DST, 4x4, 8x8, 16x16 or 32x32 - chosen at
    random (the classes of a frame interleave at CTU granularity: a class's blocks are 64-byte row segments scattered over the plane); bucketed by size class as the *_list_frames forms take them: entries
    {x, y, frame} as int16, counts [4x4 DST, 4x4 DCT, 8x8, 16x16, 32x32]"""
    cx, cy, ff = np.meshgrid(np.arange(width // 64) * 64, np.arange(height // 64) * 64, np.arange(nf), indexing="ij")
    cx, cy, ff = cx.reshape(-1), cy.reshape(-1), ff.reshape(-1)
    kind = synth.splitmix64(seed, len(cx)).astype(np.int64) % 5      # the size class of the cell
    buckets = []
    for k, nsz in enumerate((4, 4, 8, 16, 32)):
        sel = kind == k
        ox, oy = np.meshgrid(np.arange(0, 64, nsz), np.arange(0, 64, nsz))
        x = (cx[sel, None] + ox.reshape(1, -1)).reshape(-1)
        y = (cy[sel, None] + oy.reshape(1, -1)).reshape(-1)
        f = np.repeat(ff[sel], ox.size)
        buckets.append(np.stack([x, y, f], -1).astype(np.int16))
    counts = np.array([len(b) for b in buckets], np.int32)
    covered = int(sum(len(b) * (16, 16, 64, 256, 1024)[c] for c, b in enumerate(buckets)))
    return torch.from_numpy(np.ascontiguousarray(np.concatenate(buckets))).cuda(), counts, covered


def kernel_table(torch, lib, synth, stream, hbm_peak, counters, with_cpu, launches=200):
    """Every other kernel of the path on a 16-frame 4K batch (working set >> L2): CUDA events, 200 launches each (median and best of 20 groups
    of 10), and beside it the reference's C path for the same call on 1-2 frames on all host cores."""
    from oracle.binding import ptr
    NF = 16
    pitch = synth.pitch_for(W4K, PAD)
    rows = H4K + 2 * PAD
    org, fs = PAD * pitch + PAD, rows * pitch
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randint(0, 256, (NF, rows, pitch), dtype=torch.uint8, device="cuda", generator=g)
    b = torch.randint(0, 256, (NF, rows, pitch), dtype=torch.uint8, device="cuda", generator=g)
    o8 = torch.empty_like(a)
    n = NF * W4K * H4K
    out = {}
    cpu, kind, threads = (None, None, 0)
    if with_cpu:
        cpu, kind = cpu_library()
        threads = os.cpu_count() or 1
        cnf = 2                                           # frames of the CPU sample
        ha, hb = a[:cnf].cpu().numpy(), b[:cnf].cpu().numpy()
        ho = np.empty_like(ha)
        hi32 = np.empty(cnf * (W4K // 2) * (H4K // 2), np.int32)
        hsad = np.empty(cnf * (W4K // 8) * (H4K // 8) * 64, np.int32)

    def cpu_time(call, samples):
        """the reference's C path for one call (warm-up + timed) -> Gsamples/s"""
        if cpu is None:
            return None
        call()
        t0 = time.perf_counter()
        call()
        dt = time.perf_counter() - t0
        return {"gsamples_s": round(samples / dt / 1e9, 4), "cores": threads, "kind": kind, "sample": f"{cnf} 4K frames"}

    def rec(name, fn, samples, bytes_per_sample, bound="hbm", extra=None, cpu_call=None, cpu_samples=None):
        try:
            t = time_kernel(torch, fn, groups=max(1, launches // 10), per_group=min(10, max(1, launches)))
        except Exception as e:  # entry point refused the shape
            out[name] = {"error": str(e)[:100]}
            return
        gs = samples / t["ms"] / 1e6
        r = {"gsamples_s": round(gs, 1), "gsamples_s_best": round(samples / t["ms_best"] / 1e6, 1), "ms": round(t["ms"], 4), "ms_best": round(t["ms_best"], 4),
             "launches": t["launches"], "bytes_per_sample": bytes_per_sample, "gbs": round(gs * bytes_per_sample, 1),
             "hbm_frac": round(gs * bytes_per_sample / hbm_peak, 3), "hbm_frac_best": round(samples / t["ms_best"] / 1e6 * bytes_per_sample / hbm_peak, 3), "bound": bound}
        c = counters.get(name)
        if c:   # instruction mix of the committed ncu capture of this kernel (tools/ncu_counters.py)
            r["ncu"] = c
            if "idp_per_sample" in c:   # integer-pipe view (profiles/r01_pipe_peak.json: 64 IDP/clk/SM x 148 SMs x 1.965 GHz = 18.6 T IDP/s)
                r["idp_pipe_frac"] = round(gs * c["idp_per_sample"] / 1e3 / 18.6, 3)
        if extra:
            r.update(extra)
        if cpu_call is not None and cpu is not None:
            r["cpu"] = cpu_time(cpu_call, cpu_samples if cpu_samples is not None else cnf * W4K * H4K)
            if r["cpu"]:
                r["gpu_over_cpu"] = round(gs / max(r["cpu"]["gsamples_s"], 1e-9), 1)
        out[name] = r

    i32 = torch.empty((NF * (W4K // 2) * (H4K // 2),), dtype=torch.int32, device="cuda")
    best = [torch.empty((NF * (W4K // s) * (H4K // s) * 2,), dtype=torch.int32, device="cuda") for s in (8, 16, 32, 64)]
    rec("sad_pyramid_best (argmin folded in)", lambda: lib.call("sad_sweep_pyramid_best_frames", dptr(a, org), pitch, dptr(b, org), pitch, W4K, H4K, -4, -4, NF, fs, fs,
                                                                 *[dptr(o) for o in best], stream=stream), n, 2 + 8 * (1 / 64 + 1 / 256 + 1 / 1024 + 1 / 4096),
        bound="int-pipe (VABSDIFF4)")
    pk = [torch.empty((NF * (W4K // s) * (H4K // s) * 64,), dtype=torch.uint16 if s < 32 else torch.int32, device="cuda") for s in (8, 16, 32, 64)]
    rec("sad_pyramid_packed (uint16 8x8/16x16)", lambda: lib.call("sad_sweep_pyramid_packed_frames", dptr(a, org), pitch, dptr(b, org), pitch, W4K, H4K, -4, -4, NF, fs,
                                                                   fs, *[dptr(o) for o in pk], stream=stream), n, 2 + SAD_OUT_BYTES_PACKED, bound="int-pipe (VABSDIFF4)")
    del pk
    one = torch.empty((NF * (W4K // 8) * (H4K // 8) * 64,), dtype=torch.int32, device="cuda")
    for s in (8, 16, 32, 64):
        rec(f"sad_sweep_{s}x{s} (single size, 64 candidates)",
            lambda s=s: lib.call("sad_sweep_frames", dptr(a, org), pitch, dptr(b, org), pitch, W4K, H4K, (s << 8) | s, -4, -4, 8, 8, NF, fs, fs, dptr(one), stream=stream),
            n, 2 + 256 / (s * s), bound="int-pipe (VABSDIFF4)",
            cpu_call=(lambda s=s: cpu.drv("sad_sweep_frames", ptr(ha, org), pitch, ptr(hb, org), pitch, W4K, H4K, (s << 8) | s, -4, -4, 8, 8, cnf, fs, fs,
                                          ptr(hsad), threads=threads)) if with_cpu else None)
    del one
    for log2 in (2, 3, 4, 5, 6):
        N = 1 << log2
        rec(f"ssd_{N}x{N}", lambda log2=log2: lib.call("ssd_frames", dptr(a, org), pitch, dptr(b, org), pitch, W4K, H4K, log2, NF, fs, fs, dptr(i32), stream=stream),
            n, 2 + 4 / (N * N),
            cpu_call=(lambda log2=log2: cpu.drv("ssd_frames", ptr(ha, org), pitch, ptr(hb, org), pitch, W4K, H4K, log2, cnf, fs, fs, ptr(hi32), threads=threads)) if with_cpu else None)
    for log2 in (1, 2, 3):
        N = 1 << log2
        rec(f"hadamard_satd_{N}x{N}",
            lambda log2=log2: lib.call("hadamard_satd_frames", dptr(a, org), pitch, dptr(b, org), pitch, W4K, H4K, log2, NF, fs, fs, dptr(i32), stream=stream),
            n, 2 + 4 / (N * N),
            cpu_call=(lambda log2=log2: cpu.drv("hadamard_satd_frames", ptr(ha, org), pitch, ptr(hb, org), pitch, W4K, H4K, log2, cnf, fs, fs, ptr(hi32), threads=threads)) if with_cpu else None)

    # interpolation (whole planes, one fractional position per launch)
    for name, taps, xf, yf in (("pred_uni_luma_copy", 8, 0, 0), ("pred_uni_luma_h", 8, 1, 0), ("pred_uni_luma_v", 8, 0, 2), ("pred_uni_luma_hv", 8, 1, 3),
                               ("pred_uni_luma_hv_half", 8, 2, 2), ("pred_uni_chroma_copy", 4, 0, 0), ("pred_uni_chroma_h", 4, 3, 0), ("pred_uni_chroma_v", 4, 0, 5),
                               ("pred_uni_chroma_hv", 4, 3, 5)):
        two_pass = xf and yf
        extra = {"kernel": "uv::pred_vh_kernel (tcgen05.mma kind::i8 + TMA + TMEM)"} if taps == 8 and two_pass else None
        rec(name, lambda taps=taps, xf=xf, yf=yf: lib.call("pred_uni_frames", dptr(o8, org), pitch, dptr(a, org), pitch, W4K, H4K, taps, xf, yf, NF, fs, fs, stream=stream),
            n, 2, bound="int-pipe (IDP)" if two_pass else "hbm", extra=extra,
            cpu_call=(lambda taps=taps, xf=xf, yf=yf: cpu.drv("pred_uni_frames", ptr(ho, org), pitch, ptr(ha, org), pitch, W4K, H4K, taps, xf, yf, cnf, fs, fs, threads=threads)) if with_cpu else None)
    for name, taps, fr in (("pred_bi_luma_hv", 8, (1, 2, 3, 1)), ("pred_bi_luma_copy", 8, (0, 0, 0, 0)), ("pred_bi_chroma_hv", 4, (3, 5, 6, 1)),
                           ("pred_bi_chroma_copy", 4, (0, 0, 0, 0))):
        rec(name, lambda taps=taps, fr=fr: lib.call("pred_bi_frames", dptr(o8, org), pitch, dptr(a, org), dptr(b, org), pitch, W4K, H4K, taps, *fr, NF, fs, fs, stream=stream),
            n, 3, bound="int-pipe (IDP)" if any(fr) else "hbm",
            cpu_call=(lambda taps=taps, fr=fr: cpu.drv("pred_bi_frames", ptr(ho, org), pitch, ptr(ha, org), ptr(hb, org), pitch, W4K, H4K, taps, *fr, cnf, fs, fs, threads=threads)) if with_cpu else None)

    # prediction-unit lists over the whole batch in ONE launch (per-PU motion vectors, frame index in the descriptor)
    for name, sizes in (("pred_uni_list_8x8", (8,)), ("pred_uni_list_16x16", (16,)), ("pred_uni_list_32x32", (32,)), ("pred_uni_list_64x64", (64,)),
                        ("pred_uni_list_mixed_8..64", (8, 16, 32, 64))):
        if not hasattr(lib.load(), "hevcasm_pred_uni_list_frames"):
            break
        pl, covered = pu_list(synth, torch, W4K, H4K, sizes, NF)
        rec(name, lambda pl=pl: lib.call("pred_uni_list_frames", dptr(o8, org), pitch, dptr(a, org), pitch, 8, dptr(pl), pl.shape[0], fs, fs, stream=stream),
            covered, 2, bound="latency / hbm", extra={"pus": int(pl.shape[0])})
    if hasattr(lib.load(), "hevcasm_pred_bi_list_frames"):
        pl, covered = pu_list(synth, torch, W4K, H4K, (8, 16, 32, 64), NF, bi=True)
        rec("pred_bi_list_mixed_8..64", lambda: lib.call("pred_bi_list_frames", dptr(o8, org), pitch, dptr(a, org), dptr(b, org), pitch, 8, dptr(pl), pl.shape[0], fs, fs,
                                                         stream=stream), covered, 3, bound="int-pipe (IDP)", extra={"pus": int(pl.shape[0])})

    # SAD of every PU of a mixed list at its own integer vector (the refinement step after the sweep), whole batch in one launch
    if hasattr(lib.load(), "hevcasm_sad_list_frames"):
        pl, covered = pu_list(synth, torch, W4K, H4K, (8, 16, 32, 64), NF)
        sl = pl.clone()
        sl[:, 4:6] = torch.div(sl[:, 4:6], 4, rounding_mode="floor")
        sl = sl.contiguous()
        sad_out = torch.empty((sl.shape[0],), dtype=torch.int32, device="cuda")
        rec("sad_pu_list_mixed_8..64", lambda: lib.call("sad_list_frames", dptr(a, org), pitch, dptr(b, org), pitch, dptr(sl), sl.shape[0], fs, fs, dptr(sad_out),
                                                        stream=stream), covered, 2, extra={"pus": int(sl.shape[0])})

    # residual path: int16 residual planes
    rp = synth.pitch_for(W4K, 0, 128)
    res = torch.randint(-256, 256, (NF, H4K, rp), dtype=torch.int16, device="cuda", generator=g)
    co = torch.empty((n,), dtype=torch.int16, device="cuda")
    co2 = torch.empty((n,), dtype=torch.int16, device="cuda")
    cbf = torch.empty((n // 16,), dtype=torch.int32, device="cuda")
    if with_cpu:
        hres = res[:cnf].cpu().numpy()
        hco, hco2 = np.empty(cnf * W4K * H4K, np.int16), np.empty(cnf * W4K * H4K, np.int16)
        hcbf = np.empty(cnf * W4K * H4K // 16, np.int32)
    for log2, tr in ((2, 1), (2, 0), (3, 0), (4, 0), (5, 0)):
        N = 1 << log2
        extra = {"kernel": "ft::fwd_umma_kernel: first stage on tcgen05 (kind::i8 on the raw int16 tile, TMEM), second stage in registers"} if log2 >= 4 else None
        rec(f"fwd_{'dst' if tr else 'dct'}_{N}x{N}", lambda log2=log2, tr=tr: lib.call("transform_frames", dptr(co), dptr(res), rp, W4K, H4K, log2, tr, NF, H4K * rp, stream=stream),
            NF * (W4K // N * N) * (H4K // N * N), 4, extra=extra,
            cpu_call=(lambda log2=log2, tr=tr: cpu.drv("transform_frames", ptr(hco), ptr(hres), rp, W4K, H4K, log2, tr, cnf, H4K * rp, threads=threads)) if with_cpu else None,
            cpu_samples=(cnf * (W4K // N * N) * (H4K // N * N)) if with_cpu else None)
    lib.call("transform_frames", dptr(co), dptr(res), rp, W4K, H4K, 3, 0, NF, H4K * rp, stream=stream)
    if with_cpu:
        cpu.drv("transform_frames", ptr(hco), ptr(hres), rp, W4K, H4K, 3, 0, cnf, H4K * rp, threads=threads)
    nh = (cnf * W4K * H4K) if with_cpu else 0
    rec("quantize", lambda: lib.call("quantize_batch", dptr(co2), dptr(co), QP[0], QP[1], QP[2], 64, n // 64, dptr(cbf), stream=stream), n, 4 + 4 / 64,
        cpu_call=(lambda: cpu.drv("quantize_batch", ptr(hco2), ptr(hco), QP[0], QP[1], QP[2], 64, nh // 64, ptr(hcbf), threads=threads)) if with_cpu else None)
    rec("quantize_inverse", lambda: lib.call("quantize_inverse_batch", dptr(co), dptr(co2), QP[3], QP[4], n, stream=stream), n, 4,
        cpu_call=(lambda: cpu.drv("quantize_inverse_batch", ptr(hco), ptr(hco2), QP[3], QP[4], nh, threads=threads)) if with_cpu else None)
    for log2, tr in ((2, 1), (2, 0), (3, 0), (4, 0), (5, 0)):
        N = 1 << log2
        rec(f"inv_{'dst' if tr else 'dct'}_add_{N}x{N}",
            lambda log2=log2, tr=tr: lib.call("inverse_transform_add_frames", dptr(o8, org), pitch, dptr(a, org), pitch, dptr(co), W4K, H4K, log2, tr, NF, fs, fs, stream=stream),
            NF * (W4K // N * N) * (H4K // N * N), 4,
            cpu_call=(lambda log2=log2, tr=tr: cpu.drv("inverse_transform_add_frames", ptr(ho, org), pitch, ptr(ha, org), pitch, ptr(hco), W4K, H4K, log2, tr, cnf, fs, fs,
                                                        threads=threads)) if with_cpu else None,
            cpu_samples=(cnf * (W4K // N * N) * (H4K // N * N)) if with_cpu else None)
    # transform-unit lists over the whole batch in one call (mixed sizes bucketed by class, frame index per TU): up to five launches
    if hasattr(lib.load(), "hevcasm_inverse_transform_add_list_frames"):
        from oracle.binding import ptr as hptr
        tus, counts, covered = tu_buckets(synth, torch, W4K, H4K, NF)
        rec("fwd_tu_list_mixed_4..32", lambda: lib.call("transform_list_frames", dptr(co2), dptr(res), rp, dptr(tus), hptr(counts), H4K * rp, stream=stream), covered, 4,
            extra={"tus": int(counts.sum()), "by_class": counts.tolist()})
        rec("inv_tu_list_mixed_4..32", lambda: lib.call("inverse_transform_add_list_frames", dptr(o8, org), pitch, dptr(a, org), pitch, dptr(co), dptr(tus), hptr(counts),
                                                        fs, fs, stream=stream), covered, 4, extra={"tus": int(counts.sum()), "by_class": counts.tolist()})
    for log2 in (2, 3, 4, 5):
        N = 1 << log2
        rec(f"quantize_reconstruct_{N}x{N}", lambda log2=log2: lib.call("quantize_reconstruct_frames", dptr(o8, org), pitch, dptr(a, org), pitch, dptr(co), W4K, H4K, log2, NF,
                                                                        fs, fs, stream=stream), NF * (W4K // N * N) * (H4K // N * N), 4,
            cpu_call=(lambda log2=log2: cpu.drv("quantize_reconstruct_frames", ptr(ho, org), pitch, ptr(ha, org), pitch, ptr(hco), W4K, H4K, log2, cnf, fs, fs,
                                                threads=threads)) if with_cpu else None,
            cpu_samples=(cnf * (W4K // N * N) * (H4K // N * N)) if with_cpu else None)
    lv = torch.empty((n,), dtype=torch.int16, device="cuda")

    def cpu_pipeline(log2):
        N = 1 << log2
        nb = (W4K // N) * (H4K // N) * cnf
        cpu.drv("transform_frames", ptr(hco), ptr(hres), rp, W4K, H4K, log2, 0, cnf, H4K * rp, threads=threads)
        cpu.drv("quantize_batch", ptr(hco2), ptr(hco), QP[0], QP[1], QP[2], N * N, nb, ptr(hcbf), threads=threads)
        cpu.drv("quantize_inverse_batch", ptr(hco), ptr(hco2), QP[3], QP[4], nb * N * N, threads=threads)
        cpu.drv("inverse_transform_add_frames", ptr(ho, org), pitch, ptr(ha, org), pitch, ptr(hco), W4K, H4K, log2, 0, cnf, fs, fs, threads=threads)
    for log2 in (2, 3, 4, 5):
        N = 1 << log2
        rec(f"residual_pipeline_{N}x{N}_fused", lambda log2=log2: lib.call("residual_pipeline_frames", dptr(o8, org), pitch, dptr(lv), dptr(cbf), dptr(res), rp, dptr(a, org),
                                                                         pitch, W4K, H4K, log2, 0, *QP, NF, fs, H4K * rp, fs, stream=stream),
            NF * (W4K // N * N) * (H4K // N * N), 6 + 4 / (N * N), cpu_call=(lambda log2=log2: cpu_pipeline(log2)) if with_cpu else None,
            cpu_samples=(cnf * (W4K // N * N) * (H4K // N * N)) if with_cpu else None)
    if hasattr(lib.load(), "hevcasm_residual_from_planes_pipeline_frames"):
        rec("residual_pipeline_8x8_fused_from_src_pred", lambda: lib.call("residual_from_planes_pipeline_frames", dptr(o8, org), pitch, dptr(lv), dptr(cbf), dptr(b, org), pitch,
                                                                          dptr(a, org), pitch, W4K, H4K, 3, 0, *QP, NF, fs, fs, fs, stream=stream), n, 5 + 4 / 64)
    return out


def setup_dist(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun exactly the way the driver does
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    from hevcasm_b200 import lib
    lib.load()  # raises if libhevcasm_b200.so is missing: there is no fallback
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the b200 arm has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner would otherwise precede the JSON line on stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return world, rank, local, torch, dist, barrier, max_over_ranks


def timed_steps(torch, lib, step, args, barrier, local, max_over_ranks):
    """W warm-up steps, then exactly K steps between barrier + synchronize, CUDA events on the launching stream; clocks sampled meanwhile"""
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.05:   # more untimed steps until ~50 ms have passed: clocks ramped before the timed region
        step()
        torch.cuda.synchronize()
    barrier()
    launches0 = lib.launch_count()
    with ClockSampler(local) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    return max_over_ranks(e0.elapsed_time(e1)) / args.steps, lib.launch_count() - launches0, clocks


def run_gpu_4k(args):
    world, rank, local, torch, dist, barrier, max_over_ranks = setup_dist(args)
    from hevcasm_b200 import lib, synth
    hbm_peak, peak_src = peaks()
    counters = ncu_counters()
    NF = args.frames or 32
    # frames are sharded over ranks: every rank owns NF whole frames, no exchange
    src_h = synth.random_planes(synth.SEED + 17 * rank, NF, W4K, H4K, PAD)
    ref_h = synth.random_planes(synth.SEED + 17 * rank + 1, NF, W4K, H4K, PAD)
    pitch, org, fs = src_h.pitch, src_h.origin, src_h.frame_stride
    src_d, ref_d = torch.from_numpy(src_h.buf).cuda(), torch.from_numpy(ref_h.buf).cuda()
    sizes = (8, 16, 32, 64)
    out_elems = [NF * (W4K // s) * (H4K // s) * 64 for s in sizes]
    outs_d = [torch.empty((e,), dtype=torch.int32, device="cuda") for e in out_elems]
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step():
        lib.call("sad_sweep_pyramid_frames", dptr(src_d, org), pitch, dptr(ref_d, org), pitch, W4K, H4K, -4, -4, NF, fs, fs,
                 *[dptr(o) for o in outs_d], stream=stream)

    ms_per_step, launches, clocks = timed_steps(torch, lib, step, args, barrier, local, max_over_ranks)
    samples_per_step = world * NF * W4K * H4K
    value = samples_per_step / ms_per_step / 1e6  # Gsamples/s over all ranks
    per_gpu = value / world

    # ---- e2e through the host-memory C-ABI forms: page-locked host buffers on the GPU's NUMA node, copies inside the timed region
    e2e = e2e_int32 = e2e_best = None
    if not args.no_e2e:
        src_p = lib.pinned_array(src_h.buf.shape, np.uint8, device=local)
        ref_p = lib.pinned_array(ref_h.buf.shape, np.uint8, device=local)
        src_p[...] = src_h.buf
        ref_p[...] = ref_h.buf
        h2d = 2 * NF * (H4K + 2 * PAD_TRAVEL) * (W4K + 2 * PAD_TRAVEL)
        e2e_steps = max(3, min(args.steps, 10))

        def run_host(api, outs, arena):
            with lib.Context(local, arena_bytes=arena) as ctx:
                def one():
                    lib.call_host(api, ctx.handle, hptr(src_p, org), pitch, hptr(ref_p, org), pitch, W4K, H4K, PAD_TRAVEL, -4, -4, NF, fs, fs, *[hptr(o) for o in outs])
                one()
                barrier()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    one()
                barrier()
                dt = time.perf_counter() - t0
            return max_over_ranks(dt)

        # headline e2e: all 64 SADs of every PU, the 8x8 / 16x16 levels as uint16 (exact)
        pk_p = [lib.pinned_array((e,), np.uint16 if s < 32 else np.int32, device=local) for e, s in zip(out_elems, sizes)]
        dt = run_host("sad_sweep_pyramid_packed_frames_host", pk_p, 3 << 30)
        dev = [o.cpu().numpy() for o in outs_d]   # whole arrays: the host results must equal the device-resident results of the same inputs
        same = all(bool(np.array_equal(p.astype(np.int32), d)) for p, d in zip(pk_p, dev))
        e2e = {"value": samples_per_step * e2e_steps / dt / 1e9, "unit": "Gsamples/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": sum(o.nbytes for o in pk_p) * world, "steps": e2e_steps, "api": "hevcasm_sad_sweep_pyramid_packed_frames_host",
               "matches_device_path": same, "compared": "every SAD of every PU of every frame", "host_buffers": "page-locked, bound to the GPU's NUMA node (hevcasm_cuda_host_alloc_near)", "numa_node_rank0": int(lib.load().hevcasm_cuda_device_numa_node(local))}
        del pk_p
        outs_p = [lib.pinned_array((e,), np.int32, device=local) for e in out_elems]
        dt = run_host("sad_sweep_pyramid_frames_host", outs_p, 3 << 30)
        e2e_int32 = {"value": samples_per_step * e2e_steps / dt / 1e9, "unit": "Gsamples/s", "h2d_bytes_per_step": h2d * world,
                     "d2h_bytes_per_step": sum(out_elems) * 4 * world, "api": "hevcasm_sad_sweep_pyramid_frames_host (all levels int32, round 1's headline)",
                     "matches_device_path": all(bool(np.array_equal(p, d)) for p, d in zip(outs_p, dev))}
        del outs_p, dev
        best_p = [lib.pinned_array((NF * (W4K // s) * (H4K // s) * 2,), np.int32, device=local) for s in sizes]
        dt = run_host("sad_sweep_pyramid_best_frames_host", best_p, 1 << 30)
        e2e_best = {"value": samples_per_step * e2e_steps / dt / 1e9, "unit": "Gsamples/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": sum(o.nbytes for o in best_p) * world, "api": "hevcasm_sad_sweep_pyramid_best_frames_host",
                    "note": "min SAD + candidate index per PU instead of all 64 SADs"}

    if rank == 0:
        c = counters.get("sad_pyramid", {})
        traffic = c.get("dram_bytes_per_frame")
        roof = {"bound": "hbm", "kernel": "sad_pyramid_tma_kernel (hevcasm_sad_sweep_pyramid_frames)", "achieved": per_gpu * SAD_BYTES_PER_SAMPLE, "peak": hbm_peak, "unit": "GB/s",
                "frac": per_gpu * SAD_BYTES_PER_SAMPLE / hbm_peak, "traffic": traffic * NF if traffic else None,
                "traffic_source": c.get("source", "no ncu capture committed (profiles/r02_kernel_counters.json missing)"), "peak_source": peak_src,
                "algorithmic_bytes_per_sample": SAD_BYTES_PER_SAMPLE, "algorithmic_bytes_per_launch": SAD_BYTES_PER_SAMPLE * NF * W4K * H4K,
                "int_pipe": {"absdiff_per_sample": 64, "achieved_T_absdiff_s": per_gpu * 64 / 1e3, "peak_T_absdiff_s": 73.5,
                             "frac": per_gpu * 64 / 1e3 / 73.5, "peak_source": "profiles/r01_pipe_peak.json (64 VABSDIFF4/clk/SM)"}}
        line = {"metric": METRIC, "value": value, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": "4K SAD sweep, 8x8/16x16/32x32/64x64 PUs x 64 candidates (BASELINE configs[1])", "frames_per_gpu": NF,
                           "width": W4K, "height": H4K, "candidates": 64, "l2": f"inputs+outputs {int((2 * NF * fs + sum(out_elems) * 4) / 2**20)} MiB per step >> 126 MB L2",
                           "sharding": "whole frames per rank, no collective"},
                "roofline": roof, "gpu_launches": int(launches), "clocks": clocks.summary()}
        if e2e:
            line["e2e"], line["e2e_int32"], line["e2e_best"] = e2e, e2e_int32, e2e_best
        if world == 1 and not args.no_cpu:
            cpu, kind = cpu_library()
            threads = os.cpu_count() or 1
            nf_cpu = min(NF, 8)
            cpu_sad_sweep(cpu, src_h, ref_h, W4K, H4K, 1, threads)
            dt, cpu_outs = cpu_sad_sweep(cpu, src_h, ref_h, W4K, H4K, nf_cpu, threads)
            # parity inside the untimed region: EVERY SAD of the CPU sample's frames against the GPU's device-resident output
            ok = all(bool(np.array_equal(c_, o[:len(c_)].cpu().numpy())) for c_, o in zip(cpu_outs, outs_d))
            line["cpu_baseline"] = {"value": nf_cpu * W4K * H4K / dt / 1e9, "unit": "Gsamples/s", "cores": threads, "kind": kind,
                                    "sample": f"{nf_cpu} of the {NF} 4K frames, all four PU sizes, 16 four-way calls per PU",
                                    "gpu_output_matches": ok, "compared": f"all {sum(len(c_) for c_ in cpu_outs)} SADs of those frames", "note": CPU_NOTE}
        if world == 1 and not args.no_kernels:
            del src_d, ref_d, outs_d
            torch.cuda.empty_cache()
            line["kernels"] = kernel_table(torch, lib, synth, stream, hbm_peak, counters, with_cpu=not args.no_cpu, launches=args.kernel_launches)
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_gpu_8k(args):
    """configs[4]: a fixed batch of 8K frames dealt to the ranks; per frame SAD sweep + HV interpolation + fused 8x8 residual pipeline"""
    world, rank, local, torch, dist, barrier, max_over_ranks = setup_dist(args)
    from hevcasm_b200 import lib, shard, synth
    hbm_peak, peak_src = peaks()
    NFT = args.frames or 64
    f0, f1 = shard.frame_range(NFT, rank, world)
    nf = f1 - f0
    pitch = synth.pitch_for(W8K, PAD)
    rows = H8K + 2 * PAD
    org, fs = PAD * pitch + PAD, rows * pitch
    rp = synth.pitch_for(W8K, 0, 128)
    g = torch.Generator(device="cuda").manual_seed(1000 + f0)
    mk8 = lambda: torch.randint(0, 256, (max(nf, 1), rows, pitch), dtype=torch.uint8, device="cuda", generator=g)   # noqa: E731
    src, ref, pred = mk8(), mk8(), mk8()
    res = torch.randint(-256, 256, (max(nf, 1), H8K, rp), dtype=torch.int16, device="cuda", generator=g)
    ipl, rec = torch.empty_like(src), torch.empty_like(src)
    sizes = (8, 16, 32, 64)
    sad = [torch.empty((max(nf, 1) * (W8K // s) * (H8K // s) * 64,), dtype=torch.uint16 if s < 32 else torch.int32, device="cuda") for s in sizes]
    lv = torch.empty((max(nf, 1) * W8K * H8K,), dtype=torch.int16, device="cuda")
    cbf = torch.empty((max(nf, 1) * (W8K // 8) * (H8K // 8),), dtype=torch.int32, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def k_sad():
        lib.call("sad_sweep_pyramid_packed_frames", dptr(src, org), pitch, dptr(ref, org), pitch, W8K, H8K, -4, -4, nf, fs, fs, *[dptr(o) for o in sad], stream=stream)

    def k_pred():
        lib.call("pred_uni_frames", dptr(ipl, org), pitch, dptr(ref, org), pitch, W8K, H8K, 8, 1, 3, nf, fs, fs, stream=stream)

    def k_pipe():
        lib.call("residual_pipeline_frames", dptr(rec, org), pitch, dptr(lv), dptr(cbf), dptr(res), rp, dptr(pred, org), pitch, W8K, H8K, 3, 0, *QP, nf, fs, H8K * rp, fs,
                 stream=stream)

    def step():
        if nf:
            k_sad(), k_pred(), k_pipe()

    ms_per_step, launches, clocks = timed_steps(torch, lib, step, args, barrier, local, max_over_ranks)
    total = NFT * W8K * H8K
    value = total / ms_per_step / 1e6
    parts = {}
    if nf:
        for name, fn, bps in (("sad_sweep_pyramid_packed", k_sad, 2 + SAD_OUT_BYTES_PACKED), ("pred_uni_luma_hv", k_pred, 2.0), ("residual_pipeline_8x8_fused", k_pipe, PIPE8_BYTES_PER_SAMPLE)):
            t = time_kernel(torch, fn, groups=10, per_group=2, warm=1)
            parts[name] = {"ms": round(t["ms"], 4), "bytes_per_sample": bps, "hbm_frac": round(nf * W8K * H8K * bps / t["ms"] / 1e6 / hbm_peak, 3)}
    bytes_per_sample = 2 + SAD_OUT_BYTES_PACKED + 2.0 + PIPE8_BYTES_PER_SAMPLE

    # ---- e2e through the host forms on a bounded number of this rank's frames (page-locked host memory for all 64 8K frames would be 25 GB)
    e2e = None
    if not args.no_e2e:
        ne = min(nf, 4)
        h2d = d2h = 0
        dt = 0.0
        if ne:
            hp = host_planes_8k(synth, 9000 + 7 * f0, ne)
            def pin(a):
                p = lib.pinned_array(a.shape, a.dtype, device=local)
                p[...] = a
                return p
            hsrc, href, hres, hpred = (pin(p.buf) for p in hp)
            hsad = [lib.pinned_array((ne * (W8K // s) * (H8K // s) * 64,), np.uint16 if s < 32 else np.int32, device=local) for s in sizes]
            hipl, hrec = lib.pinned_array(hsrc.shape, np.uint8, device=local), lib.pinned_array(hsrc.shape, np.uint8, device=local)
            hlv = lib.pinned_array((ne * W8K * H8K,), np.int16, device=local)
            hcbf = lib.pinned_array((ne * (W8K // 8) * (H8K // 8),), np.int32, device=local)
            P0 = hp[0]
            padded = ne * (H8K + 2 * PAD_TRAVEL) * (W8K + 2 * PAD_TRAVEL)
            h2d = 3 * padded + ne * W8K * H8K * 3   # src + ref (SAD), ref (interpolation), residual (int16) + predictor
            d2h = sum(o.nbytes for o in hsad) + 2 * ne * W8K * H8K + hlv.nbytes + hcbf.nbytes
            with lib.Context(local, arena_bytes=3 << 30) as ctx:
                def one():
                    lib.call_host("sad_sweep_pyramid_packed_frames_host", ctx.handle, hptr(hsrc, P0.origin), P0.pitch, hptr(href, P0.origin), P0.pitch, W8K, H8K, PAD_TRAVEL, -4, -4,
                                  ne, P0.frame_stride, P0.frame_stride, *[hptr(o) for o in hsad])
                    lib.call_host("pred_uni_frames_host", ctx.handle, hptr(hipl, P0.origin), P0.pitch, hptr(href, P0.origin), P0.pitch, W8K, H8K, PAD_TRAVEL, 8, 1, 3, ne,
                                  P0.frame_stride, P0.frame_stride)
                    lib.call_host("residual_pipeline_frames_host", ctx.handle, hptr(hrec, P0.origin), P0.pitch, hptr(hlv), hptr(hcbf), hptr(hres, hp[2].origin), hp[2].pitch,
                                  hptr(hpred, P0.origin), P0.pitch, W8K, H8K, 3, 0, *QP, ne, P0.frame_stride, hp[2].frame_stride, P0.frame_stride)
                one()
                barrier()
                t0 = time.perf_counter()
                one()
                one()
                barrier()
                dt = (time.perf_counter() - t0) / 2 / ne * nf   # seconds this rank needs for ITS frames at the measured rate
        else:
            barrier()
            barrier()
        dt = max_over_ranks(dt)
        e2e = {"value": total / dt / 1e9, "unit": "Gsamples/s", "h2d_bytes_per_step": int(h2d / max(ne, 1) * NFT), "d2h_bytes_per_step": int(d2h / max(ne, 1) * NFT),
               "api": "hevcasm_sad_sweep_pyramid_packed_frames_host + hevcasm_pred_uni_frames_host + hevcasm_residual_pipeline_frames_host",
               "sample": f"{ne} of each rank's {nf} frames through the host forms, rate scaled to the rank's share; slowest rank counts"}
    if rank == 0:
        per_gpu = value / world
        line = {"metric": METRIC_8K, "value": value, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": "8K x 64 frames, SAD + interpolation + residual pipeline (BASELINE configs[4])", "frames_total": NFT, "frames_rank0": nf,
                           "width": W8K, "height": H8K, "per_frame": "SAD sweep 8x8..64x64 x 64 candidates (packed outputs), luma 8-tap HV interpolation (1,3), fused 8x8 residual pipeline",
                           "l2": "working set of a step is tens of GB >> 126 MB L2", "sharding": "hevcasm_b200.shard.frame_range: whole frames per rank, no collective"},
                "roofline": {"bound": "hbm", "kernel": "the three kernels of a step (sum of algorithmic bytes)", "achieved": per_gpu * bytes_per_sample, "peak": hbm_peak, "unit": "GB/s",
                             "frac": per_gpu * bytes_per_sample / hbm_peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_sample": bytes_per_sample,
                             "per_kernel_rank0": parts},
                "gpu_launches": int(launches), "clocks": clocks.summary()}
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu:
            cpu, kind = cpu_library()
            threads = os.cpu_count() or 1
            planes = host_planes_8k(synth, synth.SEED, 1)
            cpu_8k_frame(cpu, planes, threads)
            dtc = cpu_8k_frame(cpu, planes, threads)
            line["cpu_baseline"] = {"value": W8K * H8K / dtc / 1e9, "unit": "Gsamples/s", "cores": threads, "kind": kind, "sample": "1 of the 64 8K frames, all three stages",
                                    "note": CPU_NOTE}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


_JSON_OUT = None   # the process's real stdout; fd 1 itself is pointed at stderr so that only the JSON line reaches stdout


def emit(line):
    (_JSON_OUT or sys.stdout).write(json.dumps(line) + "\n")
    (_JSON_OUT or sys.stdout).flush()


def main():
    global _JSON_OUT
    args = parse()
    # libraries print to fd 1 behind Python's back (NCCL's version banner precedes the JSON line under torchrun): keep a private
    # copy of stdout for the one JSON line and send everything else to stderr
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "8k64":
        run_gpu_8k(args)
    else:
        run_gpu_4k(args)


if __name__ == "__main__":
    main()
