#!/usr/bin/env python
"""Digest of an ncu report exported with --page raw --csv and --page source --csv."""
import collections
import csv
import re
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
d = dict(zip(rows[0], rows[2]))
for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
          "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct",
          "lts__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
          "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "smsp__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"):
    print("%-70s %s" % (k, d.get(k)))
rows = list(csv.reader(open(src)))
hdr = rows[1]
data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
op_inst, op_samp = collections.Counter(), collections.Counter()
tot_inst = tot_samp = 0
stalls = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in data:
    s = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", s)
    op = m.group(2).split(".")[0] if m else s[:10]
    n = int(r[ix["Instructions Executed"]] or 0)
    sm = int(r[ix["# Samples"]] or 0)
    op_inst[op] += n
    op_samp[op] += sm
    tot_inst += n
    tot_samp += sm
    for c in stall_cols:
        stalls[c] += int(r[ix[c]] or 0)
print("total warp insts %d samples %d" % (tot_inst, tot_samp))
for op, n in op_inst.most_common(26):
    print("%-10s %8.1fM %6.2f%%  samples %6.2f%%" % (op, n / 1e6, 100.0 * n / tot_inst, 100.0 * op_samp[op] / max(tot_samp, 1)))
print({k: v for k, v in stalls.most_common(8)})
