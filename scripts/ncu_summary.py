#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel table kept under profiles/.

usage: ncu -i gpurun_out/prof_full.ncu-rep --page raw --csv > raw.csv ; python scripts/ncu_summary.py raw.csv "header note" > profiles/NAME.csv
"""
import csv
import re
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    kn = hdr.index("Kernel Name")
    if len(sys.argv) > 2:
        print("# " + sys.argv[2])
    out = csv.writer(sys.stdout)
    out.writerow(["Kernel Name"] + ["%s [%s]" % (w, units[i]) if units[i] else w for w, i in cols])
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[kn]).replace("void ", "")
        out.writerow([name] + [r[i] for _, i in cols])


if __name__ == "__main__":
    main()
