#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep, captured on the GPU box with `ncu --set full`) into a small text file for
profiles/: per captured launch the duration, SM-active share, issue and pipe utilisation, stall reasons, local-memory and DRAM
traffic.  usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.txt "title / command line" """
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# {title}\n# source: {rep} (ncu --set full --clock-control none; per-launch times under ncu are cold-cache and serialised)\n")
        for r in rows[2:]:
            f.write(f"\n== {r[idx['Kernel Name']]}  (launch id {r[idx['ID']]})\n")
            for m in METRICS:
                if m in idx and r[idx[m]] not in ("", "n/a"):
                    f.write(f"{m:88s} {r[idx[m]]} {units[idx[m]]}\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
