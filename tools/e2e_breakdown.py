#!/usr/bin/env python
"""Wall-clock split of the end-to-end path (upload / render+D2H) per chunk count. Tuning aid."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
blob, ref, meta = ge.load_golden("demo03_1080p_a4g")
blob = np.ascontiguousarray(blob)
h, w = ref.shape
for chunks in sys.argv[1:] or ["1", "2", "4", "8"]:
    os.environ["QR_B200_CHUNKS"] = chunks
    ctx = pkg.Context([0])
    pinned = torch.zeros((h, w), dtype=torch.int32).pin_memory()
    pageable = np.zeros((h, w), dtype=np.uint32)
    for frame, label in ((pinned.numpy().view(np.uint32), "pinned"), (pageable, "pageable")):
        tu, tr = [], []
        for i in range(25):
            t0 = time.perf_counter()
            ctx.upload(blob)
            t1 = time.perf_counter()
            ctx.render(frame, w)
            t2 = time.perf_counter()
            if i >= 5:
                tu.append(t1 - t0)
                tr.append(t2 - t1)
        ctx.render(None)
        ctx.sync()
        print("chunks %s %-8s upload %.3f ms  render+D2H %.3f ms  (kernel alone %.3f ms)  diff %d"
              % (chunks, label, 1e3 * np.median(tu), 1e3 * np.median(tr), ctx.last_render_ms(),
                 int((frame != ref).sum())))
    ctx.close()
