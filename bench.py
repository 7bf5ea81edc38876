#!/usr/bin/env python
"""bench.py - the driver's measurement contract for hevcasm_b200.

Headline workload (BASELINE.json configs[1]): the 4K motion-estimation SAD sweep - every 8x8, 16x16, 32x32 and 64x64
PU of a batch of 3840x2160 8-bit frames against 64 candidate vectors (dx, dy in [-4, 3]^2), all four PU sizes from one
pass over the frames (hevcasm_sad_sweep_pyramid_frames).  A "step" is one such pass over `--frames` frames per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path (one JSON line on rank 0)
    python bench.py --impl reference [...]                        # the reference's own C path on the host cores

value    : Gsamples/s (source samples, each compared against all 64 candidates), inputs resident in HBM
e2e      : the same through the host-memory C-ABI call (hevcasm_sad_sweep_pyramid_frames_host): pinned host frames in,
           SAD arrays back in host memory, copies inside the timed region
roofline : HBM roofline of the SAD kernel from its algorithmic bytes, plus the integer-pipe figures
kernels  : the other kernels of the path (SSD, interpolation, transforms, quantisation), each against its own roofline
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

W4K, H4K, PAD = 3840, 2160, 64
SAD_BYTES_PER_SAMPLE = 2.0 + 4.0 * 64 * (1 / 64 + 1 / 256 + 1 / 1024 + 1 / 4096)   # src + ref + int32 outputs of 4 levels
METRIC = "Gsamples/s (4K ME SAD sweep: 8x8..64x64 PUs x 64 candidate vectors)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=32, help="4K frames per GPU per step")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel table")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


# ---------------------------------------------------------------------------------------------- reference arm

def cpu_sad_sweep(cpu, src, ref, n_frames, threads):
    """The headline workload on the CPU: four sweeps (8, 16, 32, 64) of 64 candidates through the 4-way SAD function of
    `cpu` (oracle/binding.CpuLib).  Returns seconds."""
    from oracle.binding import ptr
    outs = [np.empty((n_frames * (W4K // s) * (H4K // s) * 64,), np.int32) for s in (8, 16, 32, 64)]
    t0 = time.perf_counter()
    for s, o in zip((8, 16, 32, 64), outs):
        cpu.drv("sad_sweep_frames", ptr(src.buf, src.origin), src.pitch, ptr(ref.buf, ref.origin), ref.pitch, W4K, H4K, (s << 8) | s, -4, -4,
                8, 8, n_frames, src.frame_stride, ref.frame_stride, ptr(o), threads=threads)
    return time.perf_counter() - t0, outs


def cpu_library():
    from oracle import binding
    ref = binding.reference()
    if ref is not None:
        ref.lib.ref_drv_set_avx2_sad(1)   # libvpx AVX2 intrinsics for the 32x32 / 64x64 four-way SAD (the reference's own SIMD)
        return ref, "reference"
    return binding.oracle(), "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hevcasm_b200 import synth
    cpu, kind = cpu_library()
    threads = os.cpu_count() or 1
    nf = 2
    src = synth.random_planes(synth.SEED, nf, W4K, H4K, PAD)
    ref = synth.random_planes(synth.SEED + 1, nf, W4K, H4K, PAD)
    for _ in range(max(args.warmup, 1)):
        cpu_sad_sweep(cpu, src, ref, 1, threads)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_sad_sweep(cpu, src, ref, nf, threads)
        t += dt
    value = args.steps * nf * W4K * H4K / t / 1e9
    sample = f"{nf} 4K frames per step, all four PU sizes, 64 candidates as 16 four-way calls per PU"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gsamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "4K SAD sweep, 8x8/16x16/32x32/64x64 PUs x 64 candidates (BASELINE configs[1])", "sample": sample,
                   "host_threads": threads},
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": threads, "kind": kind, "sample": sample,
                         "note": "reference C path at -O3 -mavx2 (+ libvpx AVX2 intrinsics for 32x32/64x64); the x86 asm needs yasm/nasm, absent here"},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ---------------------------------------------------------------------------------------------- GPU arm

def dptr(t, off=0):
    return C.c_void_p(t.data_ptr() + off * t.element_size())


def time_on_stream(torch, fn, iters, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters  # ms


def kernel_table(torch, lib, synth, stream, hbm_peak):
    """Every other kernel of the path on a 16-frame 4K batch (working set >> L2), CUDA events, 10 launches each."""
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

    def rec(name, ms, samples, bytes_per_sample, bound="hbm", extra=None):
        gs = samples / ms / 1e6
        r = {"gsamples_s": round(gs, 1), "ms": round(ms, 4), "bytes_per_sample": bytes_per_sample, "gbs": round(gs * bytes_per_sample, 1),
             "hbm_frac": round(gs * bytes_per_sample / hbm_peak, 3), "bound": bound}
        if extra:
            r.update(extra)
        out[name] = r

    i32 = torch.empty((NF * (W4K // 4) * (H4K // 4),), dtype=torch.int32, device="cuda")
    best = [torch.empty((NF * (W4K // s) * (H4K // s) * 2,), dtype=torch.int32, device="cuda") for s in (8, 16, 32, 64)]
    ms = time_on_stream(torch, lambda: lib.call("sad_sweep_pyramid_best_frames", dptr(a, org), pitch, dptr(b, org), pitch, W4K, H4K, -4, -4, NF, fs, fs,
                                                *[dptr(o) for o in best], stream=stream), 10, 3)
    rec("sad_pyramid_best (argmin folded in)", ms, n, 2 + 8 * (1 / 64 + 1 / 256 + 1 / 1024 + 1 / 4096), bound="int-pipe (VABSDIFF4)",
        extra={"absdiff_pipe_frac": round(n / ms / 1e6 * 64 / 1e3 / 73.5, 3)})
    for log2 in (3, 4):
        N = 1 << log2
        ms = time_on_stream(torch, lambda: lib.call("ssd_frames", dptr(a, org), pitch, dptr(b, org), pitch, W4K, H4K, log2, NF, fs, fs, dptr(i32),
                                                    stream=stream), 10, 3)
        rec(f"ssd_{N}x{N}", ms, n, 2 + 4 / (N * N))

    for log2 in (2, 3):
        N = 1 << log2
        ms = time_on_stream(torch, lambda: lib.call("hadamard_satd_frames", dptr(a, org), pitch, dptr(b, org), pitch, W4K, H4K, log2, NF, fs, fs, dptr(i32),
                                                    stream=stream), 10, 3)
        rec(f"hadamard_satd_{N}x{N}", ms, n, 2 + 4 / (N * N))

    # interpolation (whole planes, one fractional position per launch)
    for name, taps, xf, yf in (("pred_uni_luma_copy", 8, 0, 0), ("pred_uni_luma_h", 8, 1, 0), ("pred_uni_luma_v", 8, 0, 2), ("pred_uni_luma_hv", 8, 1, 3),
                               ("pred_uni_chroma_hv", 4, 3, 5)):
        try:
            ms = time_on_stream(torch, lambda: lib.call("pred_uni_frames", dptr(o8, org), pitch, dptr(a, org), pitch, W4K, H4K, taps, xf, yf, NF, fs, fs,
                                                        stream=stream), 10, 3)
            # luma HV: vertical pass on tcgen05 (int8 Toeplitz product, 183 MAC per sample), horizontal pass 4 IDP.2A x 1.09 (tile edges)
            idp = {"pred_uni_luma_copy": 0, "pred_uni_luma_h": 2, "pred_uni_luma_v": 2, "pred_uni_luma_hv": 4.4, "pred_uni_chroma_hv": 3.05}[name]
            # integer-pipe view (profiles/r01_pipe_peak.json: 64 IDP/clk/SM x 148 SMs x 1.965 GHz = 18.6 T IDP/s)
            extra = {"idp_per_sample": idp, "idp_pipe_frac": round(n / ms / 1e6 * idp / 1e3 / 18.6, 3)}
            if name == "pred_uni_luma_hv":
                extra["kernel"] = "uv::pred_vh_kernel (tcgen05.mma kind::i8 + TMA + TMEM; HEVCASM_PRED_HV=stream gives the CUDA-core kernel)"
            rec(name, ms, n, 2, bound="hbm" if idp < 4 else "int-pipe (IDP)", extra=extra)
        except Exception as e:  # entry point not available yet
            out[name] = {"error": str(e)[:80]}
    for name, taps, fr in (("pred_bi_luma_hv", 8, (1, 2, 3, 1)), ("pred_bi_luma_copy", 8, (0, 0, 0, 0))):
        try:
            ms = time_on_stream(torch, lambda: lib.call("pred_bi_frames", dptr(o8, org), pitch, dptr(a, org), dptr(b, org), pitch, W4K, H4K, taps, *fr, NF,
                                                        fs, fs, stream=stream), 10, 3)
            idp = 9.8 if any(fr) else 0   # two references x (vertical pass on tcgen05, horizontal pass 4 IDP.2A) + 1 IDP.2A to combine, x 1.09 tile edges; the all-zero position is a byte average
            rec(name, ms, n, 3, bound="int-pipe (IDP)" if idp else "hbm", extra={"idp_per_sample": idp, "idp_pipe_frac": round(n / ms / 1e6 * idp / 1e3 / 18.6, 3)})
        except Exception as e:
            out[name] = {"error": str(e)[:80]}

    # residual path: int16 residual planes
    rp = synth.pitch_for(W4K, 0, 128)
    res = torch.randint(-256, 256, (NF, H4K, rp), dtype=torch.int16, device="cuda", generator=g)
    co = torch.empty((n,), dtype=torch.int16, device="cuda")
    co2 = torch.empty((n,), dtype=torch.int16, device="cuda")
    cbf = torch.empty((n // 16,), dtype=torch.int32, device="cuda")
    for log2 in (2, 3, 4, 5):
        N = 1 << log2
        ms = time_on_stream(torch, lambda: lib.call("transform_frames", dptr(co), dptr(res), rp, W4K, H4K, log2, 0, NF, H4K * rp, stream=stream), 10, 3)
        extra = {"kernel": "ft::fwd_umma_kernel: first stage on tcgen05 (kind::i8 on the raw int16 tile, TMEM), second stage in registers; "
                           "HEVCASM_FWD_PATH=butterfly gives the CUDA-core kernel"} if log2 >= 4 else None
        rec(f"fwd_dct_{N}x{N}", ms, NF * (W4K // N * N) * (H4K // N * N), 4, extra=extra)
    lib.call("transform_frames", dptr(co), dptr(res), rp, W4K, H4K, 3, 0, NF, H4K * rp, stream=stream)
    ms = time_on_stream(torch, lambda: lib.call("quantize_batch", dptr(co2), dptr(co), 26214, 18, 171 << 7, 64, n // 64, dptr(cbf), stream=stream), 10, 3)
    rec("quantize", ms, n, 4 + 4 / 64)
    ms = time_on_stream(torch, lambda: lib.call("quantize_inverse_batch", dptr(co), dptr(co2), 18432, 6, n, stream=stream), 10, 3)
    rec("quantize_inverse", ms, n, 4)
    for log2 in (2, 3, 4, 5):
        N = 1 << log2
        ms = time_on_stream(torch, lambda: lib.call("inverse_transform_add_frames", dptr(o8, org), pitch, dptr(a, org), pitch, dptr(co), W4K, H4K, log2, 0, NF,
                                                    fs, fs, stream=stream), 10, 3)
        rec(f"inv_dct_add_{N}x{N}", ms, NF * (W4K // N * N) * (H4K // N * N), 4)
    ms = time_on_stream(torch, lambda: lib.call("quantize_reconstruct_frames", dptr(o8, org), pitch, dptr(a, org), pitch, dptr(co), W4K, H4K, 3, NF, fs, fs,
                                                stream=stream), 10, 3)
    rec("quantize_reconstruct_8x8", ms, n, 4)
    try:
        lv = torch.empty((n,), dtype=torch.int16, device="cuda")
        ms = time_on_stream(torch, lambda: lib.call("residual_pipeline_frames", dptr(o8, org), pitch, dptr(lv), dptr(cbf), dptr(res), rp, dptr(a, org), pitch,
                                                    W4K, H4K, 3, 0, 26214, 18, 171 << 7, 18432, 6, NF, fs, H4K * rp, fs, stream=stream), 10, 3)
        rec("residual_pipeline_8x8_fused", ms, n, 6 + 4 / 64)
    except Exception as e:
        out["residual_pipeline_8x8_fused"] = {"error": str(e)[:80]}
    return out


def run_gpu(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun exactly the way the driver does
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    from hevcasm_b200 import lib, synth
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

    hbm_peak, peak_src = peaks()
    NF = args.frames
    # frames are sharded round-robin over ranks: every rank owns NF whole frames, no exchange
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

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = lib.launch_count()
    with ClockSampler(local) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    launches = lib.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    samples_per_step = world * NF * W4K * H4K
    value = samples_per_step / ms_per_step / 1e6  # Gsamples/s over all ranks
    per_gpu = value / world

    # ---- e2e through the host-memory C-ABI form (pinned host buffers, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        src_p = lib.pinned_array(src_h.buf.shape, np.uint8)
        ref_p = lib.pinned_array(ref_h.buf.shape, np.uint8)
        src_p[...] = src_h.buf
        ref_p[...] = ref_h.buf
        outs_p = [lib.pinned_array((e,), np.int32) for e in out_elems]
        h2d = 2 * NF * (H4K + 2 * PAD) * (W4K + 2 * PAD)
        d2h = sum(out_elems) * 4
        with lib.Context(local, arena_bytes=3 << 30) as ctx:
            def e2e_step():
                lib.call_host("sad_sweep_pyramid_frames_host", ctx.handle, C.c_void_p(src_p.ctypes.data + org), pitch, C.c_void_p(ref_p.ctypes.data + org),
                              pitch, W4K, H4K, PAD, -4, -4, NF, fs, fs, *[C.c_void_p(o.ctypes.data) for o in outs_p])
            e2e_steps = max(3, min(args.steps, 10))
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            barrier()
            dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # cheap integrity check: the host results must equal the device-resident results of the same inputs
        same = all(bool(np.array_equal(o_p[:4096], o_d[:4096].cpu().numpy())) for o_p, o_d in zip(outs_p, outs_d))
        e2e = {"value": samples_per_step * e2e_steps / dt / 1e9, "unit": "Gsamples/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "steps": e2e_steps, "api": "hevcasm_sad_sweep_pyramid_frames_host", "matches_device_path": same}

    # ---- the same end to end with the argmin folded in (hevcasm_sad_sweep_pyramid_best_frames_host): 1/32 of the result bytes
    e2e_best = None
    if not args.no_e2e:
        best_p = [lib.pinned_array((NF * (W4K // s) * (H4K // s) * 2,), np.int32) for s in sizes]
        with lib.Context(local, arena_bytes=1 << 30) as ctx:
            def best_step():
                lib.call_host("sad_sweep_pyramid_best_frames_host", ctx.handle, C.c_void_p(src_p.ctypes.data + org), pitch, C.c_void_p(ref_p.ctypes.data + org),
                              pitch, W4K, H4K, PAD, -4, -4, NF, fs, fs, *[C.c_void_p(o.ctypes.data) for o in best_p])
            best_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                best_step()
            barrier()
            dtb = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dtb], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dtb = float(t.item())
        e2e_best = {"value": samples_per_step * e2e_steps / dtb / 1e9, "unit": "Gsamples/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": sum(o.nbytes for o in best_p) * world, "api": "hevcasm_sad_sweep_pyramid_best_frames_host",
                    "note": "min SAD + candidate index per PU instead of all 64 SADs"}

    if rank == 0:
        roof = {"bound": "hbm", "kernel": "sad_pyramid_tma_kernel (hevcasm_sad_sweep_pyramid_frames)", "achieved": per_gpu * SAD_BYTES_PER_SAMPLE, "peak": hbm_peak, "unit": "GB/s",
                "frac": per_gpu * SAD_BYTES_PER_SAMPLE / hbm_peak, "traffic": 55.63e6 * NF, "traffic_source": "ncu dram__bytes_read+write per launch / 8 frames, profiles/r01_sad_pyramid.md", "peak_source": peak_src,
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
            line["e2e"] = e2e
        if e2e_best:
            line["e2e_best"] = e2e_best
        if world == 1 and not args.no_cpu:
            cpu, kind = cpu_library()
            threads = os.cpu_count() or 1
            nf_cpu = min(NF, 8)
            cpu_sad_sweep(cpu, src_h, ref_h, 1, threads)
            dt, cpu_outs = cpu_sad_sweep(cpu, src_h, ref_h, nf_cpu, threads)
            ok = all(bool(np.array_equal(c[:65536], o[:65536].cpu().numpy())) for c, o in zip(cpu_outs, outs_d))
            line["cpu_baseline"] = {"value": nf_cpu * W4K * H4K / dt / 1e9, "unit": "Gsamples/s", "cores": threads, "kind": kind,
                                    "sample": f"{nf_cpu} of the {NF} 4K frames, all four PU sizes, 16 four-way calls per PU",
                                    "gpu_output_matches": ok}
        if world == 1 and not args.no_kernels:
            line["kernels"] = kernel_table(torch, lib, synth, stream, hbm_peak)
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
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
