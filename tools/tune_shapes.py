#!/usr/bin/env python
"""Kernel time of every launch shape (QR_B200_SHAPE) on a fixture. Tuning aid."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "demo03_1080p_a4g"
shapes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1", "2", "3", "4", "5"]
pkg = ge.load_package()
if name.endswith(".blob"):
    # a raw scene blob (oracle harness, QR_DUMP_BLOB): timing only, no reference frame
    import numpy as np
    blob = np.fromfile(name, dtype=np.uint8)
    ref = None
else:
    blob, ref, meta = ge.load_golden(name)
for sh in shapes:
    os.environ["QR_B200_SHAPE"] = sh
    ctx = pkg.Context([0])
    ctx.upload(blob)
    ms = []
    for _ in range(12):
        ctx.render(None)
        ctx.sync()
        ms.append(ctx.last_render_ms())
    got = ctx.render_frame()
    info = ctx.kernel_info()
    print("shape %s threads %d ctas/sm %d regs %d local %d: kernel ms min %.3f med %.3f  pixels != ref %d"
          % (sh, info["threads_per_cta"], info["ctas_per_sm"], info["regs_per_thread"],
             info["local_bytes_per_thread"], min(ms[2:]), sorted(ms[2:])[len(ms[2:]) // 2],
             int((got != ref).sum()) if ref is not None else -1))
    ctx.close()
