#!/usr/bin/env python
"""ncu raw page (ncu -i rep --page raw --csv) -> profiles/ncu_summary.json,
the file bench.py reads the dominant kernel's DRAM traffic from.
usage: ncu_summary.py raw.csv out.json "<capture command / note>" """
import csv
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

raw, out, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(raw)))
names, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(names, zip(vals, units)))
KEYS = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__sass_average_branch_targets_threads_uniform.pct")


def to_bytes(v, u):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


m = {k: {"value": d[k][0], "unit": d[k][1]} for k in KEYS if k in d}
# stamp: which kernel this is a capture of (bench.py refuses the traffic figure
# when the sources or the launch shape it runs differ)
hsh = hashlib.sha256()
for f in ("csrc/qr_b200.cu", "csrc/qr_core.cuh", "csrc/qr_kscene.h"):
    hsh.update(open(os.path.join(ROOT, "quadray-engine_b200", f), "rb").read())
try:
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], stdout=subprocess.PIPE,
                            check=True).stdout.decode().strip()
    dirty = subprocess.run(["git", "-C", ROOT, "status", "--porcelain", "quadray-engine_b200/csrc"],
                           stdout=subprocess.PIPE, check=True).stdout.decode().strip()
    if dirty:
        commit += "+uncommitted kernel edits"
except Exception:
    commit = None
kname = d.get("Kernel Name", ("", ""))[0]
mm = re.search(r"<\(?(?:bool\))?(\d+), \(?(?:int\))?(\d+), \(?(?:int\))?(\d+)>", kname)
res = {"capture": note, "kernel": kname, "commit": commit, "kernel_source_sha256": hsh.hexdigest(),
       "threads_per_cta": int(mm.group(2)) if mm else int(float(d.get("launch__block_size", ("0", ""))[0])),
       "metrics": m}
if "dram__bytes_read.sum" in d:
    res["dram_bytes_per_launch"] = to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res)[:400])
