#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of numbers DESIGN.md/bench.py cite.
usage: tools/ncu_summary.py <report.ncu-rep> [kernel-regex]  > profiles/<name>.txt"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pat and not pat.search(r[ki]):
            continue
        print(f"== {r[ki]}  (report {rep})")
        for i, h in enumerate(hdr):
            if h in KEYS or "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    if float(r[i].replace(",", "")) == 0 and "stalled" in h:
                        continue
                except ValueError:
                    pass
                print(f"  {h:90s} {units[i]:16s} {r[i]}")


if __name__ == "__main__":
    main()
