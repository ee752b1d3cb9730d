#!/usr/bin/env python
"""Randomised differential run on the GPU box: N scene / option combinations
rendered by the unmodified reference (oracle/_ref/qr_ref_harness, CPU) and by
the drop-in backend (build/qr_b200_harness); prints one JSON object with the
cases and the number of differing pixels of each.
A case gets 30 s per binary, the run stops after "budget_s" seconds and prints
what it has.
usage: fuzz_diff.py [n_cases] [seed] [budget_s]"""
import json
import os
import random
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "qr_ref_harness")
B200 = os.path.join(ROOT, "build", "qr_b200_harness")


def render(binary, args, path, env=None):
    e = dict(os.environ)
    e.update(env or {})
    try:
        p = subprocess.run([binary] + args + ["-q", "-o", path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e,
                           timeout=30)
    except subprocess.TimeoutExpired:
        return "timeout"
    if p.returncode != 0:
        return "error: " + p.stderr.decode(errors="replace")[-200:]
    info = json.loads(p.stdout.decode().strip().splitlines()[-1])
    return np.fromfile(path, dtype=np.uint32).reshape(info["y_res"], info["x_res"])


def case(rng):
    kind = rng.choice(["synth", "synth", "demo", "test", "test"])
    a = []
    if kind == "synth":
        a += ["-s", "synth", "-N", str(rng.choice([50, 200, 700, 2000])), "-S", str(rng.randrange(1, 10 ** 6)),
              "-E", str(rng.choice([6, 12, 25, 40])), "-R", str(rng.choice([0, 1, 1])),
              "-M", str(rng.choice([0, 0, 150, 400, 800]))]
    elif kind == "demo":
        a += ["-s", "demo%02d" % rng.randrange(1, 4), "-b", str(rng.randrange(0, 60000))]
    else:
        a += ["-s", "test%02d" % rng.randrange(1, 19)]
    a += ["-x", str(rng.randrange(97, 520)), "-y", str(rng.randrange(64, 300)), "-a", str(rng.randrange(0, 3))]
    if rng.random() < 0.5:
        a.append("-g")
    if rng.random() < 0.5:
        a.append("-r")
    p = rng.random()
    if p < 0.15 and kind != "synth":
        a += ["-p", "none"]
    elif p < 0.35:
        a += ["-p", "0x0230FFB9"]          # host tiling off: tile lists built on the device
    elif p < 0.5:
        a += ["-p", "full"]
    env = {}
    if rng.random() < 0.25:
        a += ["-t", "4", "-f", "3", "-d", "333"]
        if rng.random() < 0.5:
            env["QR_B200_PIPELINE"] = "1"
            a[-3] = "4"                    # the pipelined backend shows frame N - 1
    return a, env


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 20261019)
    budget = float(sys.argv[3]) if len(sys.argv) > 3 else 300.0
    t_start = time.time()
    out = {"cases": [], "differing_cases": 0, "failed_to_run": 0}
    with tempfile.TemporaryDirectory() as td:
        for _ in range(n):
            if time.time() - t_start > budget:
                break
            a, env = case(rng)
            ref_args = list(a)
            if env.get("QR_B200_PIPELINE"):
                i = ref_args.index("-f")
                ref_args[i + 1] = "3"
            want = render(REF, ref_args, os.path.join(td, "r.raw"))
            got = render(B200, a, os.path.join(td, "g.raw"), env)
            if isinstance(want, str) or isinstance(got, str):
                out["failed_to_run"] += 1
                out["cases"].append({"args": " ".join(a), "env": env, "ran": False,
                                     "reference": want if isinstance(want, str) else "ok",
                                     "b200": got if isinstance(got, str) else "ok"})
                continue
            d = int((want != got).sum()) if want.shape == got.shape else -1
            out["cases"].append({"args": " ".join(a), "env": env, "differ": d, "lit": round(float((want != 0).mean()), 3)})
            if d != 0:
                out["differing_cases"] += 1
    print(json.dumps(out))


if __name__ == "__main__":
    main()
