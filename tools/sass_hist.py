"""Opcode histogram (weighted by executed warp-instructions) and stall summary from an `ncu --page source --csv --print-source sass` dump."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si, ei = hdr.index("Source"), hdr.index("Instructions Executed")
smp = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops = collections.Counter(); samples = collections.Counter(); stalls = collections.Counter()
total = 0
for r in rows[2:]:
    if len(r) <= ei or not r[ei].isdigit(): continue
    op = r[si].split()[0] if not r[si].strip().startswith("@") else r[si].split()[1]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LD", "ST")) and "." in op else "")
    n = int(r[ei] or 0); ops[op] += n; total += n
    samples[op] += int(r[smp] or 0)
    for i in stall_cols: stalls[hdr[i]] += int(r[i] or 0)
print("total warp-instructions", total)
for op, n in ops.most_common(25): print(f"{op:14s} {n:12d} {100*n/total:5.1f}%   samples {samples[op]}")
ts = sum(stalls.values())
print("stall samples:", {k: f"{100*v/ts:.1f}%" for k, v in stalls.most_common(8)})
