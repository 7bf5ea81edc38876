"""Distils committed ncu captures into profiles/r02_kernel_counters.json, the file bench.py reads its DRAM-traffic and instruction-mix figures from
(instead of literals in the source).  Run here, no GPU needed:

    python tools/ncu_counters.py <name>=<file.ncu-rep>:<samples per launch>[:<frames per launch>] ...

<name> is the key bench.py looks up: "sad_pyramid" for the headline kernel, otherwise the name of a kernels.* entry."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("NCU_COUNTERS_OUT", os.path.join(ROOT, "profiles", "r02_kernel_counters.json"))


def distil(rep, samples, frames):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    d = dict(zip(rows[0], rows[2]))
    units = dict(zip(rows[0], rows[1]))

    def val(k):
        v = float(d[k].replace(",", ""))
        u = units.get(k, "")
        return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6}.get(u, 1)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    si, ti = hdr.index("Source"), hdr.index("Thread Instructions Executed")
    ops = collections.Counter()
    for r in rows[2:]:
        if len(r) <= ti or not r[ti].isdigit():
            continue
        t = r[si].split()
        op = (t[1] if t[0].startswith("@") else t[0]).rstrip(";").split(".")[0]
        ops[op] += int(r[ti])
    total = sum(ops.values())
    dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    out = {"kernel": d["Kernel Name"][:80], "source": os.environ.get("NCU_COUNTERS_SOURCE", os.path.relpath(rep, ROOT)), "duration_us_under_ncu": round(val("gpu__time_duration.sum"), 2),
           "dram_bytes_per_launch": dram, "dram_bytes_per_sample": round(dram / samples, 4), "thread_instructions_per_sample": round(total / samples, 2),
           "idp_per_sample": round(ops["IDP"] / samples, 3), "registers": int(float(d["launch__registers_per_thread"])),
           "top_ops_per_sample": {k: round(v / samples, 3) for k, v in ops.most_common(8)}}
    if frames:
        out["dram_bytes_per_frame"] = dram / frames
    return out


def main():
    cur = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for a in sys.argv[1:]:
        name, rest = a.split("=", 1)
        parts = rest.split(":")
        rep, samples = parts[0], float(parts[1])
        frames = float(parts[2]) if len(parts) > 2 else 0
        cur[name] = distil(rep, samples, frames)
        print(name, json.dumps(cur[name])[:300])
    json.dump(cur, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
