#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): per kernel duration, DRAM bytes, throughputs, top stall reasons."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=====", r[hdr.index("Kernel Name")][:70], "id", r[0])
        for k in KEYS:
            if k in hdr:
                print(f"   {k:70s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
        st = []
        for k, v in zip(hdr, r):
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(v), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        print("   stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:6]))


if __name__ == "__main__":
    main(sys.argv[1])
