#!/usr/bin/env python
"""BASELINE.json config 5 at the stated size on the GPUs of one box: a generated
cloud of N random quadrics under 8-ary bounding-volume arrays
(apps/qr_synth_scene.h), 7680 x 4320, 4xAA.

 1. build/qr_b200_harness renders the scene through the unmodified engine (list
    building on the host, flatten, C ABI) on all GPUs of QR_B200_DEVICES; the
    frame's row CRC-32s are compared with the reference's (tests/golden fixture)
 2. the scene blob of that frame (QR_B200_DUMP_BLOB) is then rendered a few
    times through the C ABI from Python: frame time and Mrays/s with the
    engine's host work out of the way.

usage (GPU box): config5.py <n_gpus> [fixture]      prints one JSON object"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 1
fixture = sys.argv[2] if len(sys.argv) > 2 else "synth100k_8k_a4c"
rowcrc, meta = ge.load_golden_crc(fixture)
blob_path = "/tmp/qr_config5.blob"
frame_path = "/tmp/qr_config5.raw"
env = dict(os.environ)
env.update({"QR_B200_DEVICES": ",".join(str(i) for i in range(n_gpus)), "QR_B200_DUMP_BLOB": blob_path,
            "QR_B200_TIMING": "1"})
t0 = time.perf_counter()
p = subprocess.run([os.path.join(ROOT, "build", "qr_b200_harness")] + meta["args"].split()
                   + ["-t", str(len(os.sched_getaffinity(0))), "-o", frame_path],
                   env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=3000)
wall = time.perf_counter() - t0
assert p.returncode == 0, p.stderr.decode()[-2000:]
info = json.loads(p.stdout.decode().strip().splitlines()[-1])
frame = np.fromfile(frame_path, dtype=np.uint32).reshape(info["y_res"], info["x_res"])
bad = int((ge.row_crcs(frame) != rowcrc).sum())
res = {"fixture": fixture, "args": meta["args"], "n_gpus": n_gpus,
       "engine_api": {"ms_per_call": info["ms_med"], "wall_s": wall, "rows_differ_vs_reference": bad,
                      "host_threads": info["threads"]},
       "reference": {"ms_per_call": meta.get("ref_ms"), "threads": meta.get("ref_threads"),
                     "note": "build container, when the fixture was made"}}

pkg = ge.load_package()
blob = np.fromfile(blob_path, dtype=np.uint8)
ctx = pkg.Context(list(range(n_gpus)))
t0 = time.perf_counter()
ctx.upload(blob)
ctx.sync()
res["c_abi"] = {"blob_bytes": int(blob.size), "pack_and_upload_s": time.perf_counter() - t0}
got = ctx.render_frame()
res["c_abi"]["rows_differ_vs_reference"] = int((ge.row_crcs(got) != rowcrc).sum())
ctx.ray_counts()
ts = []
for _ in range(5):
    ctx.sync()
    t1 = time.perf_counter()
    ctx.render(None)
    ctx.sync()
    ts.append((time.perf_counter() - t1) * 1e3)
counts = ctx.ray_counts()
rays = sum(counts.values()) / 5.0
ms = sorted(ts)[len(ts) // 2]
res["c_abi"].update({"frame_ms": ms, "frame_ms_all": ts, "rays_per_frame": rays, "rays": counts,
                     "Mrays_per_s": rays / (ms * 1e-3) / 1e6,
                     "primary_Msamples_per_s": (got.size << 2) / (ms * 1e-3) / 1e6,
                     "kernel": ctx.kernel_info()})
ctx.close()
print(json.dumps(res))
