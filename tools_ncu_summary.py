#!/usr/bin/env python
"""Summarise one .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / profiles/ quote.
usage: python tools_ncu_summary.py gpurun_out/x.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals))
u = dict(zip(hdr, units))


def g(k):
    try:
        return float(d[k].replace(",", ""))
    except Exception:
        return None


keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct", "smsp__issue_active.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio"]
lines = ["kernel: " + d.get("Kernel Name", "?")]
for k in keys:
    if k in d:
        lines.append("%-75s %s %s" % (k, d[k], u.get(k, "")))
stalls = []
for k in hdr:
    if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
        v = g(k)
        if v:
            stalls.append((v, k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
tot = sum(v for v, _ in stalls) or 1
lines.append("warp stall samples (share): " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for v, k in sorted(stalls, reverse=True)[:8]))
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")
