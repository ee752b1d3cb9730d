"""The product's device core (quadray-engine_b200/csrc/qr_core.cuh) compiled
for the host (tests/hostsim) against the oracle with per-sample semantics.

This proves on the CPU, for every fixture, that the kernel's algorithm --
closest hit first, shade once, explicit continuation stack -- gives exactly the
pixels of the reference's shade-every-passing-surface order.  The same source
is what nvcc compiles for sm_100a; the GPU tests repeat the comparison there.
"""
import ctypes
import os

import numpy as np
import pytest

from conftest import GOLDEN_SMALL, ROOT


@pytest.fixture(scope="module")
def hostsim():
    path = os.path.join(ROOT, "tests", "hostsim", "libqr_hostsim.so")
    if not os.path.exists(path):
        import __graft_entry__ as ge
        ge.build()
    lib = ctypes.CDLL(path)
    lib.qr_hostsim_render.restype = ctypes.c_int
    lib.qr_hostsim_render.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return lib


def run_hostsim(lib, blob, want_t=False):
    b = np.ascontiguousarray(blob, dtype=np.uint8)
    hdr = b[:256].view(np.int32)
    w, h, fsaa = int(hdr[4]), int(hdr[5]), int(hdr[7])
    frame = np.zeros((h, w), dtype=np.uint32)
    t = np.zeros((h, w, 1 << fsaa), dtype=np.float32) if want_t else None
    rays = (ctypes.c_uint64 * 4)()
    rc = lib.qr_hostsim_render(b.ctypes.data, b.size, frame.ctypes.data, w,
                               t.ctypes.data if want_t else None, 0, h, rays)
    assert rc == 0
    return frame, t, [int(x) for x in rays]


@pytest.mark.parametrize("name", GOLDEN_SMALL)
def test_core_matches_oracle_bit_exact(entry, hostsim, name):
    blob, ref, meta = entry.load_golden(name)
    want, _, _ = entry.oracle_render(blob, packet=1)
    got, _, rays = run_hostsim(hostsim, blob)
    assert int((got != want).sum()) == 0, meta["args"]
    assert rays[0] == ref.size << meta["fsaa"]
    # deferred shading never casts more secondary rays than shading every
    # passing surface does
    imm = meta["oracle_immediate_rays"]
    assert rays[1] <= imm["rays_shadow"] and rays[2] <= imm["rays_reflect"] and rays[3] <= imm["rays_refract"]


def test_core_primary_hit_distance_matches_oracle(entry, hostsim):
    """dump mode: per-sample primary hit distance, north star 1e-5 relative;
    the restatements agree exactly."""
    for name in ("test17_full_a4", "test15_full_a2", "test14_full"):
        blob, _, _ = entry.load_golden(name)
        _, t_want, _ = entry.oracle_render(blob, packet=1, want_t=True)
        _, t_got, _ = run_hostsim(hostsim, blob, want_t=True)
        assert np.array_equal(t_got.view(np.uint32), t_want.view(np.uint32)), name


def test_core_primary_hit_distance_matches_reference(entry, hostsim):
    """... and the reference's own ctx_T_BUF(0) (fixtures "_T": patched scratch
    build of the reference, oracle/Makefile `tdump`, tracer.cpp:5161)."""
    from conftest import GOLDEN_T
    assert GOLDEN_T
    for name in GOLDEN_T:
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        frame, t_got, _ = run_hostsim(hostsim, z["blob"], want_t=True)
        want = z["t"]
        t_got = t_got.reshape(want.shape)
        assert np.abs(t_got - want).max() <= 1e-5 * np.abs(want).max()
        assert np.array_equal(t_got.view(np.uint32), want.view(np.uint32)), name
        assert np.array_equal(frame, z["frame"]), name
