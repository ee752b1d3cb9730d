#!/usr/bin/env python
"""Render fixtures with the CHECKED build of the library and print its
violation counters (run with QR_B200_LIB pointing at libquadray_b200_checked.so;
tests/test_gpu_checked_build.py does).  usage: checked_run.py fixture [fixture ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
out = {}
for name in sys.argv[1:]:
    blob, ref, meta = ge.load_golden(name)
    ctx = pkg.Context([0])
    ctx.upload(blob)
    ctx.check_counters()
    got = ctx.render_frame()
    t = ctx.dump_hits()
    c = ctx.check_counters()
    info = ctx.kernel_info()
    out[name] = {"counters": c, "pixels_differ": int((got != ref).sum()), "threads": info["threads_per_cta"],
                 "staged": info["scene_in_smem"], "device_tiling": info["device_tiling"],
                 "hits_finite": float((t < 1e30).mean())}
    ctx.close()
print(json.dumps(out))
