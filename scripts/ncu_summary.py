#!/usr/bin/env python
"""Summary of one kernel of an ncu raw page (ncu -i X.ncu-rep --page raw --csv): the counters the design notes quote
(time, instructions, issue / warp activity, DRAM traffic, L2 hit rate, stall reasons per issue) plus derived per-row figures.

    python scripts/ncu_summary.py raw.csv ROWS VOCAB ["header line"]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
nrows, vocab = int(sys.argv[2]), int(sys.argv[3])
hdr, units, vals = rows[0], rows[1], rows[2]
ix = {h: i for i, h in enumerate(hdr)}
keep = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
keep += sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"))
if len(sys.argv) > 4: print("# " + sys.argv[4])
for k in keep:
    if k in ix: print("%-100s %-16s %s" % (k, units[ix[k]], vals[ix[k]]))
f = lambda k: float(vals[ix[k]].replace(",", ""))
inst, rd, wr = f("smsp__inst_executed.sum"), f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
rd *= scale[units[ix["dram__bytes_read.sum"]]]; wr *= scale[units[ix["dram__bytes_write.sum"]]]
us = f("gpu__time_duration.sum") * {"us": 1.0, "ms": 1e3, "ns": 1e-3}[units[ix["gpu__time_duration.sum"]]]
print("derived: %d rows; %.0f warp-instructions per row (%.1f thread-instructions per logit); DRAM %.1f KB read + %.1f KB written per row "
      "(algorithmic %.1f KB: x %.2f); %.2f M rows/s under the profiler (cold caches, serialised)"
      % (nrows, inst / nrows, inst * 32 / nrows / vocab, rd / nrows / 1e3, wr / nrows / 1e3, (4 * vocab + 32) / 1e3,
         (rd + wr) / nrows / (4 * vocab + 32), nrows / us))
