#!/usr/bin/env python
"""Condense `ncu --set full` reports into the few numbers DESIGN.md / bench.py quote.

usage: ncu_summary.py report.ncu-rep [...]  > profiles/xyz.txt
       ncu_summary.py --traffic KERNEL_SUBSTRING ORBITS_PER_GPU ALGORITHMIC_BYTES report.ncu-rep > profiles/r1_k1_traffic.json
         (DRAM bytes per launch of the matching kernel: what bench.py reports as roofline.traffic)
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


def traffic(substr, orbits, algo_bytes, path):
    import json

    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[0], rows[2:]
    ki, ri, wi, ti = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
    hits = [r for r in data if substr in r[ki]]
    if not hits:
        raise SystemExit(f"no kernel matching {substr!r} in {path}")
    rd = sum(float(r[ri]) for r in hits) / len(hits)
    wr = sum(float(r[wi]) for r in hits) / len(hits)
    print(json.dumps({
        "kernel": hits[0][ki].replace("void <unnamed>::", "").split("(")[0], "launches_averaged": len(hits),
        "orbits_per_gpu": int(orbits), "algorithmic_bytes": int(algo_bytes),
        "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
        "ncu_duration_ns": sum(float(r[ti]) for r in hits) / len(hits),
        "source": f"ncu --set full --clock-control none capture {path.split('/')[-1]} (python bench.py, default workload)",
    }, indent=1))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--traffic":
        return traffic(*sys.argv[2:6])
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
