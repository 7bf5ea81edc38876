"""ncu launch-list CSV (--metrics gpu__time_duration.sum) -> markdown table grouped by kernel and grid:
python tools/launch_list.py launches.csv > profiles/rNN_launches_bench.md"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
i0 = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[i0]
ki, gi, vi, mi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value"), hdr.index("Metric Name")
ui = hdr.index("Metric Unit")
groups = collections.OrderedDict()
for r in rows[i0 + 1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum": continue
    v = float(r[vi].replace(",", ""))
    if r[ui] in ("ns", "nsecond"): v /= 1000.0
    elif r[ui] in ("ms", "msecond"): v *= 1000.0
    groups.setdefault((r[ki], r[gi]), []).append(v)
print("| launches | mean us | grid | kernel |\n|---|---|---|---|")
for (k, g), v in groups.items():
    print(f"| {len(v)} | {sum(v) / len(v):.1f} | {g} | `{k[:110]}` |")
