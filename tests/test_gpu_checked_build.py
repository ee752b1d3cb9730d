"""The checked build of the library (quadray-engine_b200/Makefile `checked`,
-DQR_CHECKED): the render kernel verifies its own element cursors, surface
offsets, tile indices, pixel stores, stack levels and scratch reads and counts
violations.  compute-sanitizer is not available on the GPU pool
(profiles/r02d_compute_sanitizer.txt); this is what stands in for memcheck /
initcheck on the paths SURVEY.md section 5 names: staged and unstaged
(QR_B200_NOSTAGE=1) scene, the 640 / 768 / 1024-thread shapes, device-built
tile lists, deep recursion (test17/18), custom clippers (test16), generated
quadric clouds."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

CHECKED = os.path.join(ROOT, "quadray-engine_b200", "lib", "libquadray_b200_checked.so")
FIXTURES = ["test14_full_a4", "test17_full_a4", "test18_full_a4", "test16_full_a4rg", "synth1k_a4",
            "demo03_a4g_nt", "test05_odd", "demo02_a4rg"]


@pytest.mark.skipif(not os.path.exists(CHECKED), reason="checked library not built")
@pytest.mark.parametrize("nostage,shape", [("0", "2"), ("0", "3"), ("0", "7"), ("1", "3")])
def test_checked_build_counts_no_violation(nostage, shape):
    env = dict(os.environ)
    env.update({"QR_B200_LIB": CHECKED, "QR_B200_NOSTAGE": nostage, "QR_B200_SHAPE": shape})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "checked_run.py")] + FIXTURES, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    res = json.loads(p.stdout.decode().strip().splitlines()[-1])
    for name in FIXTURES:
        r = res[name]
        assert r["counters"] == [0] * 8, (name, r)
        assert r["pixels_differ"] == 0, (name, r)
        if nostage == "1":
            assert r["staged"] == 0, (name, r)
        elif not name.startswith("synth"):          # (the generated clouds do not fit shared memory)
            assert r["staged"] == 1, (name, r)
    assert res["demo03_a4g_nt"]["device_tiling"] == 1


@pytest.mark.skipif(not os.path.exists(CHECKED), reason="checked library not built")
def test_checked_build_counters_can_fire():
    """QR_B200_CHECK_SELFTEST=1 trips every check once: the counters are wired."""
    env = dict(os.environ)
    env.update({"QR_B200_LIB": CHECKED, "QR_B200_CHECK_SELFTEST": "1"})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "checked_run.py"), "test05_odd"], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    res = json.loads(p.stdout.decode().strip().splitlines()[-1])
    assert res["test05_odd"]["counters"][:7] == [1] * 7, res


@pytest.mark.skipif(not os.path.exists(CHECKED), reason="checked library not built")
def test_normal_build_has_no_counters(pkg):
    c = pkg.Context([0])
    with pytest.raises(pkg.QuadRayError):
        c.check_counters()
    c.close()
