"""Differential tests on the GPU box: the UNMODIFIED reference (oracle/_ref/
qr_ref_harness, built in the build container from /root/reference, CPU, the
checker) and the drop-in backend (build/qr_b200_harness) render the same scene
through the same public API; the frames must be identical.  These go beyond the
committed fixtures: generated quadric clouds of other seeds / sizes / mirror
shares, the animated demo scenes at other times, odd resolutions."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "oracle", "_ref", "qr_ref_harness")
B200 = os.path.join(ROOT, "build", "qr_b200_harness")
need = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(B200)),
                          reason="needs oracle/_ref/qr_ref_harness and build/qr_b200_harness (built where /root/reference is)")


def render(binary, args, path, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([binary] + args.split() + ["-q", "-o", path], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       timeout=600, env=e)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    info = json.loads(p.stdout.decode().strip().splitlines()[-1])
    return np.fromfile(path, dtype=np.uint32).reshape(info["y_res"], info["x_res"])


def both(args, tmp_path, env=None):
    want = render(REF, args, str(tmp_path / "ref.raw"))
    got = render(B200, args, str(tmp_path / "b200.raw"), env)
    return want, got


CLOUDS = [
    "-s synth -N 300 -S 11 -E 12 -x 320 -y 200 -a 2",
    "-s synth -N 1500 -S 12 -E 30 -M 150 -x 320 -y 200",
    "-s synth -N 800 -S 13 -E 20 -R 0 -x 333 -y 211 -a 1",
    "-s synth -N 2500 -S 14 -E 35 -M 400 -x 320 -y 200 -a 2 -r -g",
    "-s synth -N 64 -S 15 -E 6 -M 500 -x 256 -y 160 -a 2 -g",
    "-s synth -N 5000 -S 16 -E 40 -x 400 -y 240 -p none",
    "-s synth -N 1200 -S 17 -E 25 -M 250 -x 320 -y 200 -a 2 -p 0x0230FFB9",
]


@need
@pytest.mark.parametrize("args", CLOUDS)
def test_generated_quadric_clouds(tmp_path, args):
    want, got = both(args, tmp_path)
    assert want.any()
    assert int((got != want).sum()) == 0, args


ANIMATED = [
    "-s demo01 -x 400 -y 240 -a 2 -g -b 3700",
    "-s demo01 -x 400 -y 240 -r -g -b 12345",
    "-s demo02 -x 400 -y 240 -a 2 -g -b 2500",
    "-s demo02 -x 403 -y 237 -a 1 -r -g -b 9100",
    "-s demo03 -x 400 -y 240 -a 2 -g -b 5000",
    "-s demo03 -x 400 -y 240 -a 2 -r -g -b 20000",
    "-s demo03 -x 640 -y 360 -g -b 31415 -p 0x0230FFB9",
]


@need
@pytest.mark.parametrize("args", ANIMATED)
def test_demo_scenes_at_other_times(tmp_path, args):
    want, got = both(args, tmp_path)
    assert int((got != want).sum()) == 0, args


@need
def test_animation_sequence_synchronous_and_threads(tmp_path):
    """Five consecutive animated frames (update phases on 4 host threads): the
    last one equals the reference's last one."""
    args = "-s demo03 -x 400 -y 240 -a 2 -g -b 1000 -d 250 -f 5 -t 4"
    want, got = both(args, tmp_path)
    assert int((got != want).sum()) == 0


@need
@pytest.mark.parametrize("args", ["-s test18 -x 200 -y 120 -a 2 -r -Q -d 0 -f 5",
                                  "-s test13 -x 160 -y 96 -r -g -Q -d 0 -f 3",
                                  "-s test18 -x 203 -y 77 -a 1 -Q -d 0 -f 4 -p none"])
def test_path_tracer_other_sizes(tmp_path, args):
    want, got = both(args, tmp_path)
    assert int((got != want).sum()) == 0, args
