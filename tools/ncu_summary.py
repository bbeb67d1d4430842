#!/usr/bin/env python
"""Condense an ncu --set full report into a small CSV (metric,unit,launch0,launch1,...) of the metrics the
roofline discussion uses.  usage: tools/ncu_summary.py <report.ncu-rep> <out.csv>"""
import csv, io, subprocess, sys
rep, outp = sys.argv[1:3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_static", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "sm__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
want += [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_warp_active.pct")]
want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
with open(outp, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + ["launch%d" % k for k in range(len(data))])
    for m in want:
        if m in hdr:
            c = hdr.index(m)
            w.writerow([m, units[c]] + [r[c] for r in data])
