#!/usr/bin/env python
"""Kernel time of one GPU's share of the frame (tile rows r, r + N, ...) for
N = 1, 2, 4, 8: what strong scaling loses to launch, staging and the tail."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "demo03_1080p_a4g"
pkg = ge.load_package()
blob, ref, meta = ge.load_golden(name)
h, w = ref.shape
ctx = pkg.Context([0])
ctx.upload(blob)
buf = torch.zeros((h, w), dtype=torch.int32, device="cuda:0")
torch.cuda.synchronize()
full = None
for n in (1, 2, 4, 8, 16):
    best = []
    for r in range(min(n, 4)):
        ms = []
        for _ in range(8):
            ctx.render_rows(buf.data_ptr(), w, r, n)
            ctx.sync()
            ms.append(ctx.last_render_ms())
        best.append(min(ms[2:]))
    t = max(best)
    if n == 1:
        full = t
    print("N=%2d: share kernel %.3f ms (ideal %.3f, efficiency %.0f %%), threads/CTA %d"
          % (n, t, full / n, 100.0 * full / n / t, ctx.kernel_info()["threads_per_cta"]))
ctx.close()
