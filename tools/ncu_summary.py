"""Print the handful of ncu metrics we track per round from an .ncu-rep (developer tool; output goes to profiles/)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_shared_loads", "smsp__inst_executed.sum",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=" * 100)
    for k in keys:
        if k in d: print(f"{k:75s} {d[k]} {units[hdr.index(k)]}")
    print("-- stall reasons (warps stalled per issue-active cycle) --")
    for h in sorted(stall, key=lambda h: -float(d[h] or 0))[:8]:
        print(f"  {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):28s} {float(d[h]):.3f}")
