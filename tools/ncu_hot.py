"""Top stalled SASS lines of an .ncu-rep: python tools/ncu_hot.py file.ncu-rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
si, ei, smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
out = []
for k, r in enumerate(rows[2:]):
    if len(r) <= ei or not r[ei].isdigit(): continue
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    out.append((int(r[smp] or 0), k, int(r[ei]), r[si][:70], st))
tot = sum(o[0] for o in out)
print("total samples", tot)
for o in sorted(out, reverse=True)[:n]: print(o)
