#!/usr/bin/env python
"""Path-tracer frame time through the drop-in API, on the GPU
box: rt_Scene::set_pton(1) + render() x N, the unmodified reference on all host
cores beside the B200 backend, same scene / size / frame count, and that both
show the same picture (row CRC-32s of the last frame).  One JSON object."""
import json
import os
import subprocess
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "qr_ref_harness")
B200 = os.path.join(ROOT, "build", "qr_b200_harness")


def run(binary, args, out):
    p = subprocess.run([binary] + args + ["-o", out], check=True, stdout=subprocess.PIPE)
    return json.loads(p.stdout.decode().strip().splitlines()[-1])


def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "test18"
    x, y = (sys.argv[2], sys.argv[3]) if len(sys.argv) > 3 else ("1920", "1080")
    frames = sys.argv[4] if len(sys.argv) > 4 else "8"
    ncpu = len(os.sched_getaffinity(0))
    base = ["-s", scene, "-x", x, "-y", y, "-a", "2", "-r", "-g", "-Q", "-d", "0", "-f", frames, "-w", "2"]
    res = {"scene": scene, "x_res": int(x), "y_res": int(y), "fsaa": "4x", "frames": int(frames) + 2,
           "what": "ms per rt_Scene::render() call with the path tracer on (one more sample per pixel sample), median"}
    crc = {}
    for name, binary, extra in (("reference_%dthr" % ncpu, REF, ["-t", str(ncpu)]), ("b200", B200, ["-t", str(ncpu)])):
        out = "/tmp/pt_%s.raw" % name
        j = run(binary, base + extra, out)
        res[name] = {"ms_med": j["ms_med"], "ms_min": j["ms_min"], "threads": j["threads"], "simd": j["simd"]}
        fr = np.fromfile(out, dtype=np.uint32).reshape(int(y), int(x))
        crc[name] = [zlib.crc32(r.tobytes()) for r in fr]
        res[name]["lit"] = float((fr != 0).mean())
    a, b = list(crc.values())
    res["rows_differ"] = int(sum(1 for p, q in zip(a, b) if p != q))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
