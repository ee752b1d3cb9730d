#!/usr/bin/env python
"""Re-count rays and algorithmic IEEE operations of every pixel fixture with
the current device core compiled for the host (tests/hostsim) and rewrite the
"rays" / "ieee_ops" entries of the fixtures' meta.  Blobs and frames are kept."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden as mg  # noqa: E402

for fn in sorted(os.listdir(mg.OUT)):
    if not fn.endswith(".npz") or fn.endswith("h.npz"):
        continue
    path = os.path.join(mg.OUT, fn)
    z = np.load(path)
    blob, frame = z["blob"], z["frame"]
    meta = json.loads(bytes(z["meta"]).decode())
    hframe, rays, ops = mg.deferred_counts(np.ascontiguousarray(blob), meta["x_res"], meta["y_res"], meta["fsaa"])
    assert int((hframe != frame).sum()) == meta["oracle_packet1_mismatch_vs_ref"], fn
    changed = rays != meta.get("rays") or ops != meta.get("ieee_ops")
    meta["rays"], meta["ieee_ops"] = rays, ops
    if changed:
        np.savez_compressed(path, blob=blob, frame=frame,
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
    print("%-22s rays %10d  ops %12d %s" % (fn[:-4], rays["total"], ops["total"], "(updated)" if changed else ""))
