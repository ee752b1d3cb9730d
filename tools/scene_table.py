#!/usr/bin/env python
"""ms per rt_Scene::render(time) call for every test / demo scene at 1080p 4xAA
(BASELINE.json config 3 at size): the unmodified reference on all host cores
next to the B200 drop-in (synchronous and pipelined), same harness, same
options.  Runs on the GPU box; prints one JSON object.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "qr_ref_harness")
B200 = os.path.join(ROOT, "build", "qr_b200_harness")


def run(binary, args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([binary] + args, env=e, check=True, stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL).stdout.decode()
    return json.loads(out.strip().splitlines()[-1])


def main():
    frames = sys.argv[1] if len(sys.argv) > 1 else "60"
    ncpu = len(os.sched_getaffinity(0))
    scenes = ["test%02d" % i for i in range(1, 19)] + ["demo01", "demo02", "demo03"]
    res = {"x_res": 1920, "y_res": 1080, "fsaa": "4x", "host_cores": ncpu, "frames": int(frames),
           "what": "median ms per rt_Scene::render(time) call, time += 16 ms per frame, %d update threads" % ncpu,
           "scenes": {}}
    for s in scenes:
        base = ["-s", s, "-x", "1920", "-y", "1080", "-a", "2", "-d", "16", "-f", frames, "-w", "5", "-t", str(ncpu)]
        if s.startswith("test"):
            base += ["-p", "full"]
        else:
            base += ["-g"]
        r = run(REF, base)
        a = run(B200, base)
        b = run(B200, base, {"QR_B200_PIPELINE": "1"})
        res["scenes"][s] = {"reference_ms": r["ms_med"], "b200_sync_ms": a["ms_med"], "b200_pipelined_ms": b["ms_med"],
                            "speedup_pipelined": round(r["ms_med"] / b["ms_med"], 1)}
        sys.stderr.write("%s %s\n" % (s, res["scenes"][s]))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
