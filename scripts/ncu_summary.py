#!/usr/bin/env python
"""Condense `ncu --set full` reports into the few numbers DESIGN.md / bench.py quote.

usage: ncu_summary.py report.ncu-rep [...]  > profiles/xyz.txt
Reads each report through `ncu -i <rep> --page raw --csv` (ncu must be on PATH).
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_read", "dram__bytes_read.sum"),
    ("dram_write", "dram__bytes_write.sum"),
    ("dram_pct_peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct_peak", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("smem_dyn", "launch__shared_mem_per_block_dynamic"),
    ("smem_static", "launch__shared_mem_per_block_static"),
    ("occupancy_theo_pct", "sm__maximum_warps_per_active_cycle_pct"),
    ("occupancy_achieved_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("inst_executed", "smsp__inst_executed.sum"),
    ("ipc_active", "sm__inst_executed.avg.per_cycle_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct"),
    ("stall_long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_short_scoreboard", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall_lg_throttle", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("stall_mio_throttle", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    ("stall_math_pipe", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("smem_bank_conflicts_ld", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
    ("smem_bank_conflicts_st", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
    ("sm_clock_mhz", "sm__cycles_elapsed.avg.per_second"),
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"# {path}: no kernels")
            continue
        hdr, units, data = rows[0], rows[1], rows[2:]
        ki = hdr.index("Kernel Name")
        print(f"# {path}")
        for r in data:
            name = r[ki].replace("void <unnamed>::", "")
            print(f"kernel: {name[:110]}")
            for label, metric in METRICS:
                if metric in hdr:
                    i = hdr.index(metric)
                    print(f"  {label:26s} {r[i]:>18s} {units[i]}")
            print()


if __name__ == "__main__":
    main()
