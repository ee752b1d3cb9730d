#!/usr/bin/env python
"""Frame time through the drop-in API (rt_Scene::render(time): update phases +
render0), animated RooT default scene, 1080p 4xAA + gamma -- the "update +
render" number of SURVEY.md 8d/8f next to the render-only bench.py figure.

Runs on the GPU box: the reference harness on all host cores, the B200 harness
synchronously (1 and N update threads) and pipelined (QR_B200_PIPELINE=1:
update of frame N + 1 overlaps the GPU's frame N).  Prints one JSON object.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "qr_ref_harness")
B200 = os.path.join(ROOT, "build", "qr_b200_harness")
NOTILE = "0x0230FFB9"       # RT_OPTS_FULL without RT_OPTS_TILING | RT_OPTS_TILING_EXT1


def run(binary, args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([binary] + args, env=e, check=True, stdout=subprocess.PIPE).stdout.decode()
    return json.loads(out.strip().splitlines()[-1])


def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "demo03"
    frames = sys.argv[2] if len(sys.argv) > 2 else "200"
    ncpu = len(os.sched_getaffinity(0))
    base = ["-s", scene, "-x", "1920", "-y", "1080", "-a", "2", "-g", "-d", "16", "-f", frames, "-w", "10"]
    res = {"scene": scene, "x_res": 1920, "y_res": 1080, "fsaa": "4x", "host_cores": ncpu, "frames": int(frames),
           "what": "ms per rt_Scene::render(time) call, animated (time += 16 ms per frame), median / min"}
    cases = [
        ("reference_%dthr" % ncpu, REF, ["-t", str(ncpu)], {}),
        ("b200_sync_1thr", B200, [], {}),
        ("b200_sync_%dthr" % ncpu, B200, ["-t", str(ncpu)], {}),
        ("b200_pipelined_1thr", B200, [], {"QR_B200_PIPELINE": "1"}),
        ("b200_pipelined_%dthr" % ncpu, B200, ["-t", str(ncpu)], {"QR_B200_PIPELINE": "1"}),
        # engine tiling off (RT_OPTS_TILING / _EXT1 cleared), tile lists built on the device
        ("b200_sync_%dthr_device_tiling" % ncpu, B200, ["-t", str(ncpu), "-p", NOTILE],
         {"QR_B200_EXPECT_DEVICE_TILING": "1"}),
        ("b200_pipelined_%dthr_device_tiling" % ncpu, B200, ["-t", str(ncpu), "-p", NOTILE],
         {"QR_B200_PIPELINE": "1", "QR_B200_EXPECT_DEVICE_TILING": "1"}),
        ("reference_%dthr_tiling_off" % ncpu, REF, ["-t", str(ncpu), "-p", NOTILE], {}),
    ]
    for name, binary, extra, env in cases:
        if not os.path.exists(binary):
            res[name] = None
            continue
        j = run(binary, base + extra, env)
        res[name] = {"ms_med": j["ms_med"], "ms_min": j["ms_min"], "ms_mean": j["ms_mean"], "threads": j["threads"]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
