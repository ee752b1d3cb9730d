"""Path tracer: rt_Scene::set_pton, render0 with pt_on.

The pins are frames of the UNMODIFIED reference after N accumulated frames
(tools/make_golden.py, "_pt" fixtures, 512x2v2 target).  With pt_on the
reference's result depends on its SIMD width (seeds advance per tentative hit,
packet-wide branches have unmasked side effects), so the oracle is compared at
packet = 32 and the GPU renders packets: one warp = one packet (csrc/qr_pt.cuh).

CPU: the oracle against the reference's frames; the packet tracer compiled for
the host (a "warp" of one lane) against the oracle at packet = 1, in frames,
seeds and colour planes.  GPU: the C ABI and the drop-in harness against the
reference's frames, the device's seed / colour planes against the oracle's.
"""
import ctypes
import os
import subprocess
import json

import numpy as np
import pytest

from conftest import GOLDEN_PT, ROOT


class PtState(ctypes.Structure):
    _fields_ = [("pseed", ctypes.c_void_p), ("ptr_r", ctypes.c_void_p), ("ptr_g", ctypes.c_void_p),
                ("ptr_b", ctypes.c_void_p), ("pts_c", ctypes.c_float)]


@pytest.fixture(scope="module")
def oracle(entry):
    lib = entry.load_oracle()
    lib.qr_oracle_render_pt.restype = ctypes.c_int
    lib.qr_oracle_render_pt.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(PtState)]
    lib.qr_oracle_pt_seed.restype = None
    lib.qr_oracle_pt_seed.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    return lib


@pytest.fixture(scope="module")
def hostsim():
    path = os.path.join(ROOT, "tests", "hostsim", "libqr_hostsim.so")
    if not os.path.exists(path):
        pytest.skip("tests/hostsim/libqr_hostsim.so not built (run __graft_entry__.build())")
    lib = ctypes.CDLL(path)
    lib.qr_hostsim_render_pt.restype = ctypes.c_int
    lib.qr_hostsim_render_pt.argtypes = ([ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4
                                         + [ctypes.POINTER(ctypes.c_float)])
    return lib


def geometry(blob):
    hdr = np.ascontiguousarray(blob[:256]).view(np.int32)
    w, h, row = int(hdr[4]), int(hdr[5]), int(hdr[6])
    return w, h, row, 4 * row * h


def fresh_state(oracle, n):
    seeds = np.zeros(n, np.uint32)
    oracle.qr_oracle_pt_seed(seeds.ctypes.data, n)
    return seeds, np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)


_memo = {}


def oracle_frames(oracle, blob, frames, packet, rows=None):
    """Frame after 1 and after "frames" frames, and the final state.  "rows":
    only that band of the frame (packets never leave their row, and a row's
    seeds and colours are its own, so a band accumulates like the frame)."""
    b = np.ascontiguousarray(blob, dtype=np.uint8)
    key = (b.tobytes()[:4096], b.size, frames, packet, rows)
    if key in _memo:
        return _memo[key]
    w, h, row, n = geometry(b)
    y0, y1 = rows if rows is not None else (0, h)
    sd, pr, pg, pb = fresh_state(oracle, n)
    st = PtState(sd.ctypes.data, pr.ctypes.data, pg.ctypes.data, pb.ctypes.data, 0.0)
    fr = np.zeros((h, w), np.uint32)
    first = None
    for k in range(frames):
        assert oracle.qr_oracle_render_pt(b.ctypes.data, b.size, fr.ctypes.data, w, packet, y0, y1, ctypes.byref(st)) == 0
        if k == 0:
            first = fr.copy()
    assert st.pts_c == float(frames)
    _memo[key] = (first, fr, (sd, pr, pg, pb))
    return _memo[key]


@pytest.mark.parametrize("name", GOLDEN_PT)
def test_oracle_path_tracer_renders_the_reference_frames(entry, oracle, name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    assert int(np.ascontiguousarray(z["blob"][12:16]).view(np.uint32)[0]) & 0x100      # QR_BLOB_PT
    # a band of the frame here (the CPU suite's time); the GPU tests compare whole frames and states
    h = meta["y_res"]
    band = (h // 3, h // 3 + max(8, h // 4))
    first, last, _ = oracle_frames(oracle, z["blob"], meta["frames"], 32, band)
    assert int((first[band[0]:band[1]] != z["frame1"][band[0]:band[1]]).sum()) == 0
    assert int((last[band[0]:band[1]] != z["frame"][band[0]:band[1]]).sum()) == 0


@pytest.mark.parametrize("name,lanes", [("test18_a4_pt", 8), ("test18_a4_pt", 16), ("test02_a2rg_pt", 8),
                                        ("test02_a2rg_pt", 16), ("test17_r_pt", 16), ("demo02_rg_pt", 8)])
def test_oracle_follows_the_reference_at_other_simd_widths(oracle, name, lanes):
    """The reference's 256x1v2 (8 lanes) and 512x1v2 (16 lanes) targets render
    their own path-traced frames; the oracle at packet = 8 / 16 renders those."""
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    h = meta["y_res"]
    band = (h // 3, h // 3 + max(8, h // 4))
    _, last, _ = oracle_frames(oracle, z["blob"], meta["frames"], lanes, band)
    assert int((last[band[0]:band[1]] != z["frame_w%d" % lanes][band[0]:band[1]]).sum()) == 0


def test_path_traced_frames_depend_on_the_packet_width(oracle):
    """The property that forces a packet tracer on the GPU: with every sample
    deciding alone (packet = 1) a few pixels differ from the 32-lane target's."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "test17_r_pt.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    _, last1, _ = oracle_frames(oracle, z["blob"], meta["frames"], 1)
    assert int((last1 != z["frame"]).sum()) > 0


@pytest.mark.parametrize("name", ["test18_a4_pt", "test17_r_pt", "test02_a2rg_pt", "test05_odd_pt", "demo02_rg_pt"])
def test_packet_tracer_host_build_equals_oracle_packet_1(oracle, hostsim, name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    b = np.ascontiguousarray(z["blob"], dtype=np.uint8)
    w, h, row, n = geometry(b)
    _, want, (wsd, wr, wg, wb) = oracle_frames(oracle, b, meta["frames"], 1)
    sd, pr, pg, pb = fresh_state(oracle, n)
    fr = np.zeros((h, w), np.uint32)
    cnt = ctypes.c_float(0.0)
    for _ in range(meta["frames"]):
        assert hostsim.qr_hostsim_render_pt(b.ctypes.data, b.size, fr.ctypes.data, w, 0, h, sd.ctypes.data,
                                            pr.ctypes.data, pg.ctypes.data, pb.ctypes.data, ctypes.byref(cnt)) == 0
    assert int((fr != want).sum()) == 0
    assert np.array_equal(sd, wsd)
    for a, c in ((pr, wr), (pg, wg), (pb, wb)):
        assert np.array_equal(a.view(np.uint32), c.view(np.uint32))


# ------------------------------------------------------------------ GPU ---

@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_PT)
def test_gpu_path_tracer_renders_the_reference_frames(entry, pkg, oracle, name):
    """C ABI: upload (QR_BLOB_PT) + qr_pt_reset + N frames = the reference's frame
    after N frames; seeds and colour planes = the oracle's at packet 32."""
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    b = np.ascontiguousarray(z["blob"], dtype=np.uint8)
    w, h, row, n = geometry(b)
    ctx = pkg.Context([0])
    try:
        ctx.upload(b)
        ctx.pt_reset(n)
        launches0 = ctx.launch_count()
        first = None
        for k in range(meta["frames"]):
            ctx.upload(b)
            fr = ctx.render_frame()
            if k == 0:
                first = fr.copy()
        assert ctx.pt_frames() == meta["frames"]
        assert ctx.launch_count() - launches0 >= meta["frames"]          # the GPU path ran (a frame may be launched in chunks)
        sd, pr, pg, pb = ctx.pt_fetch(n)
    finally:
        ctx.close()
    assert int((first != z["frame1"]).sum()) == 0
    assert int((fr != z["frame"]).sum()) == 0
    _, _, (wsd, wr, wg, wb) = oracle_frames(oracle, b, meta["frames"], 32)
    assert np.array_equal(sd, wsd)
    for a, c in ((pr, wr), (pg, wg), (pb, wb)):
        assert np.array_equal(a.view(np.uint32), c.view(np.uint32))


@pytest.mark.gpu
def test_gpu_path_tracer_state_is_required_and_resettable(entry, pkg):
    z = np.load(os.path.join(ROOT, "tests", "golden", "test18_a4_pt.npz"))
    b = np.ascontiguousarray(z["blob"], dtype=np.uint8)
    w, h, row, n = geometry(b)
    ctx = pkg.Context([0])
    try:
        ctx.upload(b)
        with pytest.raises(RuntimeError):
            ctx.render_frame()                  # no qr_pt_reset yet
        ctx.pt_reset(n)
        a = ctx.render_frame()
        ctx.render_frame()
        ctx.pt_reset(n)                         # set_pton off / on: the same first frame again
        c = ctx.render_frame()
        assert ctx.pt_frames() == 1
        with pytest.raises(RuntimeError):
            ctx.dump_hits()
    finally:
        ctx.close()
    assert np.array_equal(a, z["frame1"]) and np.array_equal(c, z["frame1"])


@pytest.mark.gpu
def test_gpu_path_tracer_on_several_devices(entry, pkg):
    """One context over several GPUs: every device keeps the seed / colour planes
    of the tile rows it is dealt.  Needs >= 2 devices."""
    import torch
    nd = torch.cuda.device_count()
    if nd < 2:
        pytest.skip("needs at least 2 GPUs")
    for name in ("test18_q_pt", "test05_odd_pt"):
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        meta = json.loads(bytes(z["meta"]).decode())
        b = np.ascontiguousarray(z["blob"], dtype=np.uint8)
        w, h, row, n = geometry(b)
        ctx = pkg.Context(list(range(nd)))
        try:
            ctx.upload(b)
            ctx.pt_reset(n)
            for _ in range(meta["frames"]):
                ctx.upload(b)
                fr = ctx.render_frame()
        finally:
            ctx.close()
        assert int((fr != z["frame"]).sum()) == 0, name


HARNESS = os.path.join(ROOT, "build", "qr_b200_harness")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(HARNESS), reason="build/qr_b200_harness was not built")
@pytest.mark.parametrize("name", ["test18_a4_pt", "test18_q_pt", "test05_odd_pt", "demo03_a4rg_pt"])
def test_scene_api_path_tracer(tmp_path, name):
    """rt_Scene::set_pton(1) + render() x N through the drop-in backend: the
    frame of the unmodified reference after N frames."""
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    out = str(tmp_path / "f.raw")
    p = subprocess.run([HARNESS] + meta["args"].split() + ["-Q", "-d", "0", "-f", str(meta["frames"]), "-o", out],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()
    frame = np.fromfile(out, dtype=np.uint32).reshape(meta["y_res"], meta["x_res"])
    assert int((frame != z["frame"]).sum()) == 0
