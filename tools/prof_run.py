#!/usr/bin/env python
"""Render a fixture a few times through the C ABI (for ncu / sanitizer runs).
usage: prof_run.py [fixture] [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "demo03_1080p_a4g"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pkg = ge.load_package()
blob, ref, meta = ge.load_golden(name)
ctx = pkg.Context([0])
ctx.upload(blob)
for _ in range(frames):
    ctx.render(None)
    ctx.sync()
    print("kernel ms", ctx.last_render_ms())
got = ctx.render_frame()
print("pixels != reference:", int((got != ref).sum()), "of", got.size)
ctx.close()
