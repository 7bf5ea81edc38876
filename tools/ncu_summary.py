"""Key metrics + opcode mix from an .ncu-rep (run here, no GPU needed): python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'launch__grid_size',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:72s} {units[i]:10s}", [r[i][:60] for r in rows[2:]])
if len(sys.argv) > 2:   # every metric whose name matches the regex
    import re
    for i, h in enumerate(hdr):
        if re.search(sys.argv[2], h): print(f"{h:72s} {units[i]:10s}", [r[i][:60] for r in rows[2:]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
si, ei, smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops, stalls, total = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    if len(r) <= ei or not r[ei].isdigit(): continue
    t = r[si].split()
    op = t[1] if t[0].startswith("@") else t[0]
    parts = op.rstrip(";").split(".")
    op = parts[0] + ("." + parts[1] if parts[0] in ("LDG", "STG", "LDS", "STS", "IDP") and len(parts) > 1 else "")
    n = int(r[ei]); ops[op] += n; total += n
    for i in stall_cols: stalls[hdr[i]] += int(r[i] or 0)
print("warp-instructions", total)
print("  ".join(f"{op} {100*n/total:.1f}%" for op, n in ops.most_common(22)))
ts = sum(stalls.values()) or 1
print("stalls:", "  ".join(f"{k[6:]} {100*v/ts:.0f}%" for k, v in stalls.most_common(8)))
