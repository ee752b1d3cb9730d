"""The drop-in path on a B200: the reference's ENGINE (core/engine, core/system,
compiled unmodified from /root/reference by quadray-engine_b200/Makefile) linked
with this repo's replacement tracer translation unit and libquadray_b200.so.

  build/qr_b200_harness   headless stand-in for RooT: rt_Platform + rt_Scene
                          construct / render / get_frame through the public API
  build/core_test_b200    the reference's own test/core_test.cpp, unmodified

Both binaries are built in the build container (they need the reference's
sources) and travel to the GPU box; nothing here reads /root/reference."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN_CRC, ROOT

pytestmark = pytest.mark.gpu

HARNESS = os.path.join(ROOT, "build", "qr_b200_harness")
CORE_TEST = os.path.join(ROOT, "build", "core_test_b200")


def run_harness(args, tmp_path, env=None):
    out = str(tmp_path / "frame.raw")
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([HARNESS] + args + ["-o", out], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       timeout=600, env=e)
    assert p.returncode == 0, p.stderr.decode()
    info = json.loads(p.stdout.decode().strip().splitlines()[-1])
    frame = np.fromfile(out, dtype=np.uint32).reshape(info["y_res"], info["x_res"])
    return frame, info


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="build/qr_b200_harness was not built")
@pytest.mark.parametrize("name", ["test01_full", "test14_full_a4", "test17_full_a4", "test16_none_a4",
                                  "test05_odd", "demo02_a4g", "demo03_a4g"])
def test_scene_api_renders_the_reference_frame(entry, tmp_path, name):
    """rt_Scene::render() -> render0 (tracer_b200.cpp) -> flatten -> C ABI ->
    kernel -> get_frame(): the frame equals the unmodified reference's."""
    _, ref, meta = entry.load_golden(name)
    frame, info = run_harness(meta["args"].split(), tmp_path)
    assert info["simd"] == "512x2v2"                   # what switch0 answers
    assert frame.shape == ref.shape
    assert int((frame != ref).sum()) == 0, meta["args"]


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="build/qr_b200_harness was not built")
@pytest.mark.parametrize("name", ["demo03_a4g_nt", "demo02_a4g_nt", "test14_full_nt", "test16_a4rg_nt", "test05_odd_nt"])
def test_scene_api_with_device_side_tiling(entry, tmp_path, name):
    """the application switches the engine's host tiling off
    (RT_OPTS_TILING), the flattener sends the bounding boxes along and the
    device builds the tile lists; the frame is the one the reference renders
    with ITS tiling on.  Synchronous and pipelined."""
    _, ref, meta = entry.load_golden(name)
    env = {"QR_B200_EXPECT_DEVICE_TILING": "1"}
    frame, _ = run_harness(meta["blob_args"].split(), tmp_path, env)
    assert int((frame != ref).sum()) == 0, meta["blob_args"]
    env["QR_B200_PIPELINE"] = "1"
    frame, _ = run_harness(meta["blob_args"].split() + ["-t", "4", "-f", "3", "-d", "0"], tmp_path, env)
    assert int((frame != ref).sum()) == 0, meta["blob_args"]


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="build/qr_b200_harness was not built")
@pytest.mark.parametrize("name", GOLDEN_CRC)
def test_generated_quadric_clouds_at_size(entry, tmp_path, name):
    """BASELINE.json config 5: 10k / 100k random quadrics under 8-ary
    bounding-volume arrays (apps/qr_synth_scene.h), up to 7680x4320 4xAA.  The
    harness generates the scene, the engine builds its lists, the backend
    flattens and renders; every row of the frame has the CRC-32 of the
    reference's row (fixture made by tools/make_golden.py from the unmodified
    reference), i.e. the frames are identical.  The scene does not fit shared
    memory, so this is also the unstaged (L1/L2) kernel at size."""
    rowcrc, meta = entry.load_golden_crc(name)
    frame, info = run_harness(meta["args"].split(), tmp_path)
    assert frame.shape == (meta["y_res"], meta["x_res"])
    bad = np.nonzero(entry.row_crcs(frame) != rowcrc)[0]
    assert bad.size == 0, (meta["args"], bad[:10].tolist())


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="build/qr_b200_harness was not built")
def test_scene_api_animation_and_update_phases(tmp_path):
    """Several frames with the engine's update phases running between them
    (animated demo scene): the drop-in keeps rendering and frames change."""
    a, _ = run_harness(["-s", "demo01", "-a", "2", "-f", "1", "-b", "0"], tmp_path)
    b, _ = run_harness(["-s", "demo01", "-a", "2", "-f", "3", "-b", "0", "-d", "500"], tmp_path)
    assert a.shape == b.shape and a.any() and b.any()
    assert int((a != b).sum()) > 0


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="build/qr_b200_harness was not built")
def test_scene_api_pipelined_mode_lags_one_frame(tmp_path):
    """QR_B200_PIPELINE=1 (update of frame N + 1 overlaps the GPU's frame N):
    the caller sees F0 F0 F1 F2 ..., every frame identical to the synchronous
    run's; host threads run the update phases (-t 4)."""
    args = ["-s", "demo01", "-a", "2", "-b", "0", "-d", "500", "-t", "4"]
    sync3, _ = run_harness(args + ["-f", "3"], tmp_path)                     # F2
    sync4, _ = run_harness(args + ["-f", "4"], tmp_path)                     # F3
    pipe4, _ = run_harness(args + ["-f", "4"], tmp_path, {"QR_B200_PIPELINE": "1"})
    pipe5, _ = run_harness(args + ["-f", "5"], tmp_path, {"QR_B200_PIPELINE": "1"})
    assert int((sync3 != sync4).sum()) > 0
    assert np.array_equal(pipe4, sync3)
    assert np.array_equal(pipe5, sync4)
    one, _ = run_harness(args + ["-f", "1"], tmp_path, {"QR_B200_PIPELINE": "1"})
    first, _ = run_harness(args + ["-f", "1"], tmp_path)
    assert np.array_equal(one, first)                                         # the first frame is synchronous


@pytest.mark.skipif(not os.path.exists(CORE_TEST), reason="build/core_test_b200 was not built")
def test_reference_core_test_passes_on_the_gpu(tmp_path):
    """test/core_test.cpp, unmodified: 18 scenes, RT_OPTS_NONE vs RT_OPTS_FULL
    frames compared with its own tolerance (core_test.cpp:96-145)."""
    (tmp_path / "dump").mkdir()
    p = subprocess.run([CORE_TEST], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=900)
    text = p.stdout.decode(errors="replace")
    assert p.returncode == 0, text[-2000:]
    # one "Time N" / "Time F" pair per scene, no "Frames differ", no exception
    assert len(re.findall(r"Time F", text)) == 18, text[-2000:]
    assert "Frames differ" not in text and "Exception" not in text, text[-2000:]
    assert "simd =  512x2v2" in text or "512x2v2" in text
