"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md quotes.
    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
name_i = hdr.index("Kernel Name")
for d in data:
    if pat not in d[name_i]:
        continue
    print("kernel:", d[name_i][:90])
    for k in KEYS:
        if k in hdr:
            print("  %-86s %s %s" % (k, d[hdr.index(k)], units[hdr.index(k)]))
