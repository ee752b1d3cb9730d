#!/usr/bin/env python
"""Work-item timeline of one launch (tuning build `make itemlog`, run on the GPU
box): writes the raw log (start ns, duration ns, SM per item) and prints where
the launch's time goes -- head (first item starts), steady part, tail (from the
moment the queue ran dry to the last warp's end).

usage: item_log.py [fixture] [n_ranks] [out_prefix]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("QR_B200_LIB", os.path.join(ROOT, "quadray-engine_b200", "lib", "libquadray_b200_itemlog.so"))
import numpy as np  # noqa: E402
import __graft_entry__ as ge  # noqa: E402
import torch  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "demo03_1080p_a4g"
ranks = int(sys.argv[2]) if len(sys.argv) > 2 else 1
prefix = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "itemlog")
pkg = ge.load_package()
blob, ref, meta = ge.load_golden(name)
h, w = ref.shape
ctx = pkg.Context([0])
ctx.upload(blob)
buf = torch.zeros((h, w), dtype=torch.int32, device="cuda:0")
torch.cuda.synchronize()
for _ in range(5):
    ctx.render_rows(buf.data_ptr(), w, 0, ranks)
    ctx.sync()
plain = ctx.last_render_ms()
path = "%s_%s_n%d.bin" % (prefix, name, ranks)
os.environ["QR_B200_ITEM_LOG"] = path
ctx.render_rows(buf.data_ptr(), w, 0, ranks)
ctx.sync()
logged = ctx.last_render_ms()
del os.environ["QR_B200_ITEM_LOG"]
info = ctx.kernel_info()
ctx.close()

rec = np.fromfile(path, dtype=np.dtype([("t0", "<u8"), ("dt", "<u4"), ("sm", "<u4")]))
ok = rec["dt"] > 0
t0 = rec["t0"][ok].astype(np.int64)
dt = rec["dt"][ok].astype(np.int64)
base = t0.min()
t0 -= base
t1 = t0 + dt
span = t1.max()
last_start = t0.max()
slots = info["sm_count"] * info["threads_per_cta"] // 32 * info["ctas_per_sm"]
print("fixture %s, share 1/%d: %d items (%d logged), %d warp slots" % (name, ranks, len(rec), int(ok.sum()), slots))
print("kernel ms: %.3f plain, %.3f with the log" % (plain, logged))
print("first item start .. last item end: %.1f us; last item STARTS at %.1f us (tail %.1f us = %.1f %%)"
      % (span / 1e3, last_start / 1e3, (span - last_start) / 1e3, 100.0 * (span - last_start) / span))
print("item duration us: mean %.1f, median %.1f, p90 %.1f, p99 %.1f, max %.1f"
      % (dt.mean() / 1e3, np.median(dt) / 1e3, np.percentile(dt, 90) / 1e3, np.percentile(dt, 99) / 1e3, dt.max() / 1e3))
busy = dt.sum()
print("sum of item durations / (slots x span) = %.3f" % (busy / (slots * span)))
# how many warps are still at work, in 10 us steps over the last 150 us
for back in range(150, -1, -10):
    t = span - back * 1000
    n = int(((t0 <= t) & (t1 > t)).sum())
    print("  t = end - %3d us: %4d items in flight" % (back, n))
# duration by tile row (locality preserving cost map)
if ranks == 1:
    rows = (h + 7) // 8
    per_row = len(rec) // rows
    d = rec["dt"][: rows * per_row].reshape(rows, per_row).astype(np.float64)
    m = d.mean(axis=1) / 1e3
    mx = d.max(axis=1) / 1e3
    print("tile rows: mean item us (max) from the top of the frame:")
    print(" ".join("%.0f(%.0f)" % (a, b) for a, b in zip(m, mx)))
