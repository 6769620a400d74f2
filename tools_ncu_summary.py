"""python tools_ncu_summary.py <report.ncu-rep> — prints the few metrics DESIGN.md / profiles/ quote, from `ncu -i ... --page raw --csv`"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in want:
    if k in m:
        print("%-75s %s %s" % (k, m[k][0], m[k][1]))
stalls = {h: float(v[0].replace(",", "")) for h, v in m.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and v[0] not in ("", "n/a")}
tot = sum(stalls.values()) or 1.0
top = sorted(stalls.items(), key=lambda kv: -kv[1])[:9]
print("warp stall samples (share): " + ", ".join("%s %.1f%%" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot) for k, v in top))
