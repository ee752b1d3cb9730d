#!/usr/bin/env python
"""A few path-traced frames of a "_pt" fixture through the C ABI (for ncu)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "test18_q_pt"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if name.endswith(".blob"):
    b = np.fromfile(name, dtype=np.uint8)
else:
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    b = np.ascontiguousarray(z["blob"], dtype=np.uint8)
hdr = b[:256].view(np.int32)
n = 4 * int(hdr[6]) * int(hdr[5])
pkg = ge.load_package()
ctx = pkg.Context([0])
ctx.upload(b)
ctx.pt_reset(n)
for _ in range(frames):
    fr = ctx.render_frame()
    print("frame %.3f ms" % ctx.last_render_ms())
ctx.close()
