#!/usr/bin/env python
"""Annotated SASS listing of an ncu report: every instruction of the kernel in
program order with its source line, executed count (M warp instructions),
threads per warp and stall samples.

usage: ncu_sass.py <src.csv from `ncu -i rep --page source --csv`> <lib.so>
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

src_csv, lib = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src_csv)))
kname = rows[0][1]
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
m = re.match(r"void (\w+)<\(bool\)(\d), \(int\)(\d+), \(int\)(\d+)>", kname)
mangled = "_Z%d%sILb%sELi%sELi%sEEv9qr_launch" % (len(m.group(1)), m.group(1), m.group(2), m.group(3), m.group(4))
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(td, cubin)], check=True,
                          stdout=subprocess.PIPE).stdout.decode().splitlines()
inside = False
chain = []
fresh = True
k = 0
for ln in sass:
    if ln.startswith(".text."):
        inside = ln.startswith(".text." + mangled + ":")
        continue
    if not inside:
        continue
    mm = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if mm:
        if fresh:
            chain = []
            fresh = False
        chain.append("%s:%s" % (os.path.basename(mm.group(1)).replace("qr_core.cuh", "c").replace("qr_b200.cu", "b"), mm.group(2)))
        continue
    if re.match(r"\s*\.L_x_\d+:", ln):
        print(ln.strip())
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+", ln):
        fresh = True
        r = data[k] if k < len(data) else None
        k += 1
        n = int(r[ix["Instructions Executed"]] or 0) if r else 0
        s = int(r[ix["# Samples"]] or 0) if r else 0
        t = int(r[ix["Thread Instructions Executed"]] or 0) if r else 0
        ins = re.sub(r"/\*[0-9a-f]+\*/", "", ln).strip()
        print("%8.2f %5d %4.1f  %-70s %s" % (n / 1e6, s, t / max(n, 1), ins[:70], "<".join(chain[:3])))
