"""Device-side tiling (quadray-engine_b200/csrc/qr_tiling.cuh)
on the CPU: the very functions the two tiling kernels run, compiled for the
host (tests/hostsim).

The reference's engine tiles on the host (stile + merge, engine.cpp:1956-2128,
3129-3232); with RT_OPTS_TILING off it leaves every tile head at the camera
list, and the backend culls per tile itself.  Parity argument: tiling is
conservative culling, so a per-tile SUPERSET of the reference's list in the
same order renders the same pixels.  Checked here, tile by tile, against the
lists the reference built for the same frames."""
import ctypes
import os

import numpy as np
import pytest

import qr_blob
from conftest import ROOT

PAIRS = [("demo03_a4g", "demo03_a4g_nt"), ("demo01_a4g", "demo01_a4g_nt"), ("demo02_a4g", "demo02_a4g_nt"),
         ("test14_full", "test14_full_nt"), ("test12_full", "test12_full_nt"), ("test05_odd", "test05_odd_nt")]


@pytest.fixture(scope="module")
def hostsim():
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "hostsim", "libqr_hostsim.so"))
    lib.qr_hostsim_tile_list.restype = ctypes.c_int
    lib.qr_hostsim_tile_list.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    return lib


def surf_key(srf_row):
    """Geometry of a surface record (list / material indices differ between
    two flattenings of the same scene): words 0..42 hold pos .. a_sgn, minus
    the index-valued trnode (19) and clip_head (23)."""
    w = srf_row[:44].copy()
    w[19] = 0
    w[23] = 0
    return w.tobytes()


def reference_tile_surfaces(sec, t):
    """Leaf surfaces of tile t in the list the reference's engine built."""
    el, srf = sec["elem"], sec["surf"]
    out = []
    e = int(sec["tiles"][t])
    while e >= 0:
        s = int(el[e, 2])
        if s >= 0 and srf[s, 43] >= 0:          # srf_t[3] < 0: array (trnode element)
            out.append(s)
        e = int(el[e, 3])
    return out


@pytest.mark.parametrize("tiled,untiled", PAIRS)
def test_device_tile_lists_are_ordered_supersets_of_the_reference_s(entry, hostsim, tiled, untiled):
    bt, _, _ = entry.load_golden(tiled)
    bu, _, mu = entry.load_golden(untiled)
    st, su = qr_blob.sections(bt.tobytes()), qr_blob.sections(bu.tobytes())
    assert su["header"]["off_bounds"] != 0 and su["header"]["n_bounds"] == su["header"]["n_surf"]
    assert len(set(su["tiles"].tolist())) == 1, "the engine ran with RT_OPTS_TILING off"
    assert st["header"]["n_tiles"] == su["header"]["n_tiles"]
    b = np.ascontiguousarray(bu)
    buf = np.zeros(4096, dtype=np.int32)
    ref_total = dev_total = 0
    for t in range(su["header"]["n_tiles"]):
        n = hostsim.qr_hostsim_tile_list(b.ctypes.data, b.size, t, buf.ctypes.data, buf.size)
        assert 0 <= n <= buf.size, (untiled, t, n)
        dev = [surf_key(su["surf"][i]) for i in buf[:n]]
        ref = [surf_key(st["surf"][i]) for i in reference_tile_surfaces(st, t)]
        # ref is a sub-sequence of dev: same order, nothing the reference tests is missing
        it = iter(dev)
        assert all(k in it for k in ref), (untiled, t)
        ref_total += len(ref)
        dev_total += n
    # ... and not much more than the reference tests (bounding rectangle + half a tile)
    assert dev_total <= 1.6 * ref_total + su["header"]["n_tiles"], (ref_total, dev_total)


def test_tiled_blobs_are_left_alone(entry, hostsim):
    b, _, _ = entry.load_golden("demo03_a4g")
    b = np.ascontiguousarray(b)
    buf = np.zeros(16, dtype=np.int32)
    assert hostsim.qr_hostsim_tile_list(b.ctypes.data, b.size, 0, buf.ctypes.data, 16) == -1
