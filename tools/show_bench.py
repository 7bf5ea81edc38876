"""Prints the numbers of a bench.py JSON line in readable form: python tools/show_bench.py file.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
print("value", round(d["value"], 2), d["unit"], "| ms/step", round(d["ms_per_step"], 4), "| n_gpus", d["n_gpus"], "| scaling", d["scaling"], "| launches", d.get("gpu_launches"))
r = d.get("roofline", {})
if r:
    print("roofline frac", round(r["frac"], 3), "achieved", round(r["achieved"], 1), "peak", r["peak"], "traffic", r.get("traffic"))
    if "per_kernel_rank0" in r:
        print("  per kernel", r["per_kernel_rank0"])
for k in ("e2e", "e2e_int32", "e2e_best", "cpu_baseline"):
    if k in d:
        print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in d[k].items() if a in ("value", "matches_device_path", "gpu_output_matches", "d2h_bytes_per_step", "h2d_bytes_per_step", "cores", "numa_node_rank0", "sample")})
print("clocks", d.get("clocks"))
for k, v in d.get("kernels", {}).items():
    if "error" in v:
        print("  ", k, v)
        continue
    c = v.get("cpu") or {}
    print(f"   {k:48s} {v['ms']:.4f} ms  frac {v['hbm_frac']:.3f} (best {v['hbm_frac_best']:.3f})  {v['gsamples_s']:7.1f} Gs/s  cpu {c.get('gsamples_s')}  x{v.get('gpu_over_cpu')}")
