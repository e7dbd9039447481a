#!/usr/bin/env python
"""Summarise an .ncu-rep: key throughput metrics, stall reasons, and the hottest SASS lines.  usage: ncu_summary.py rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmalite_cycles_active", "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_xu.avg.pct",
        "sm__inst_executed_pipe_lsu.avg.pct", "sm__inst_executed_pipe_uniform", "sm__inst_executed_pipe_adu", "sm__inst_executed_pipe_cbu",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct", "lts__t_sector_hit_rate.pct", "idc__request", "smsp__average_warp", "smsp__warps_issue_stalled", "smsp__average_warps_issue_stalled"]
for i, n in enumerate(h):
    if any(n.startswith(k) for k in keys) and not n.endswith("_not_issued") and ".min" not in n and ".max" not in n:
        print(f"{n:90s} {u[i]:12s} {v[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ci = {n: i for i, n in enumerate(h)}
body = [r for r in rows[hi + 1:] if len(r) > ci["# Samples"]]
tot = sum(float(r[ci["# Samples"]]) for r in body)
print("total samples", tot, " instructions", len(body))
# cumulative share by stall reason
for k in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_dispatch", "stall_not_selected", "stall_selected", "stall_barrier", "stall_mio", "stall_no_inst", "stall_lg", "stall_branch_resolving"):
    if k in ci:
        print(f"  {k:22s} {sum(float(r[ci[k]]) for r in body)/tot*100:5.1f}%")
order = sorted(range(len(body)), key=lambda i: -float(body[i][ci["# Samples"]]))
for i in order[:top]:
    r = body[i]
    why = max(((k, float(r[ci[k]])) for k in ci if k.startswith("stall_") and "Not Issued" not in k), key=lambda t: t[1])
    print(f"{float(r[ci['# Samples']])/tot*100:5.1f}%  #{i:5d} {r[ci['Source']].strip()[:90]:90s} {why[0]}")
