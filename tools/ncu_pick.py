#!/usr/bin/env python3
"""Print selected metrics from an `ncu --page raw --csv` export (one kernel per row)."""
import csv, sys
pats = sys.argv[2:] or ["gpu__time_duration.sum", "registers_per_thread", "warps_active.avg.pct", "pipe_fp64_cycles_active.avg.pct",
                        "issue_active.avg.pct", "smsp__inst_executed.sum", "thread_inst_executed_per_inst",
                        "issue_stalled", "dram__bytes_read.sum", "dram__bytes_write.sum", "icc", "inst_issued.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    for h, u, v in zip(hdr, units, r):
        if any(p in h for p in pats) and ".max" not in h and ".min" not in h:
            print("%-90s %s %s" % (h, v, u))
