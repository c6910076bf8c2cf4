#!/usr/bin/env python
"""Summarise an ncu raw-page CSV (ncu -i x.ncu-rep --page raw --csv) into the few lines that matter."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = [k for k in hdr if any(s in k for s in (
    "Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct", "smsp__inst_executed.avg.per_cycle_active",
    "sm__throughput.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__occupancy_limit", "sm__pipe_", "smsp__average_warp", "gpu__dram_throughput", "lts__t_bytes.sum", "smsp__cycles_active.avg",
    "l1tex__data_bank_conflicts_pipe_lsu", "smsp__warp_issue_stalled"))
    and not any(s in k for s in ("_pred_on", ".max_rate", "peak_sustained.", "_realtime", ".min.", ".max.", ".sum.pct", "_elapsed",
                                 ".min ", ".max "))]
for r in rows[2:]:
    print("=" * 100)
    for k in keys:
        v = r[idx[k]]
        if v not in ("0", "", "n/a"):
            print("%-90s %s %s" % (k, v, units[idx[k]]))
