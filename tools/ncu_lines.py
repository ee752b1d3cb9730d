#!/usr/bin/env python
"""Per-source-line digest of an ncu report (instructions executed, stall samples).

usage: ncu_lines.py <src.csv from `ncu -i rep --page source --csv`> <lib.so> [top N]

The ncu source page lists SASS instructions in order; nvdisasm -g on the cubin
of the same build gives the source line of each instruction.  Joined by order.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

src_csv, lib = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60

rows = list(csv.reader(open(src_csv)))
kname = rows[0][1]
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
m = re.match(r"void (\w+)<\(bool\)(\d), \(int\)(\d+), \(int\)(\d+)(?:, \(int\)(\d+))?>", kname)
if m.group(5) is None:
    mangled = "_Z%d%sILb%sELi%sELi%sEEv9qr_launch" % (len(m.group(1)), m.group(1), m.group(2), m.group(3), m.group(4))
else:
    mangled = "_Z%d%sILb%sELi%sELi%sELi%sEEv9qr_launch" % (len(m.group(1)), m.group(1), m.group(2), m.group(3),
                                                          m.group(4), m.group(5))

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, check=True,
                   stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(td, cubin)], check=True,
                          stdout=subprocess.PIPE).stdout.decode().splitlines()

LEAF = (52, 216)    # qr_core.cuh: arithmetic / load / select helpers, attributed to their caller


def is_leaf(f, l):
    return (f == "qr_core.cuh" and LEAF[0] <= l <= LEAF[1]) or f.endswith(".hpp") or f.endswith(".h") \
        or (f == "qr_core.cuh" and 280 <= l <= 316)


lines = []          # (file, line) per instruction, in order: innermost non-helper frame
inside = False
chain = []
fresh = True
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
        chain.append((os.path.basename(mm.group(1)), int(mm.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+", ln):
        fresh = True
        cur = ("?", 0)
        for c in chain:
            cur = c
            if not is_leaf(*c):
                break
        lines.append(cur)

if len(lines) != len(data):
    sys.stderr.write("warning: %d SASS instructions in cubin, %d in report\n" % (len(lines), len(data)))

by_inst, by_samp, by_thr, by_fp = (collections.Counter(), collections.Counter(), collections.Counter(),
                                   collections.Counter())
tot_i = tot_s = 0
for (f, l), r in zip(lines, data):
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    t = int(r[ix["Thread Instructions Executed"]] or 0)
    by_inst[(f, l)] += n
    by_samp[(f, l)] += s
    by_thr[(f, l)] += t
    op = re.sub(r"^\s*(@!?U?P\d+\s+)?", "", r[ix["Source"]]).split(".")[0].split()[0]
    if op in ("FMUL", "FADD", "FFMA", "MUFU", "FSETP", "FSEL", "FMNMX", "FCHK"):
        by_fp[(f, l)] += n
    tot_i += n
    tot_s += s

print("kernel %s: %d warp instructions, %d samples" % (mangled, tot_i, tot_s))
print("%-22s %10s %7s %7s %6s %6s" % ("line", "inst(M)", "inst%", "samp%", "thr/w", "fp%"))
for key, n in by_inst.most_common(top):
    print("%-16s:%-5d %10.1f %6.2f%% %6.2f%% %6.1f %5.0f%%" % (key[0], key[1], n / 1e6, 100.0 * n / tot_i,
                                                         100.0 * by_samp[key] / max(tot_s, 1),
                                                         by_thr[key] / max(n, 1), 100.0 * by_fp[key] / max(n, 1)))
