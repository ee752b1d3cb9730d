"""The oracle (oracle/render0_oracle.c) against the golden frames rendered by
the UNMODIFIED reference (tests/golden/*.npz, made by tools/make_golden.py).

This is what pins the oracle: SURVEY.md 8c -- the reference ships no golden
images, so frames of the reference itself are the known answers.
"""
import numpy as np
import pytest

from conftest import GOLDEN_HASHED, GOLDEN_SMALL, GOLDEN_T

T_REL_TOL = 1e-5                # north star: per-ray hit distance, relative

# emulating the reference's 32-lane packets must reproduce its frame bit for bit
PACKET32 = ["test01_full", "test05_full", "test12_full", "test15_full", "test16_full",
            "test17_full_a4", "test18_full_a4", "test15_full_a2", "test14_none",
            "test05_odd", "demo01_a4g", "demo03_a4g"]


@pytest.mark.parametrize("name", PACKET32)
def test_oracle_packet32_is_bit_identical_to_reference(entry, name):
    blob, ref, meta = entry.load_golden(name)
    got, _, stats = entry.oracle_render(blob, packet=32)
    assert got.shape == ref.shape
    assert int((got != ref).sum()) == 0, meta["args"]
    assert stats["rays_primary"] >= ref.size << meta["fsaa"]


@pytest.mark.parametrize("name", GOLDEN_SMALL)
def test_oracle_per_sample_semantics_match_reference(entry, name):
    """packet=1 (what the GPU computes) vs the reference frame: the north-star
    bar is >= 99.9 % identical and the rest within 1 LSB; here they are equal."""
    blob, ref, meta = entry.load_golden(name)
    got, _, _ = entry.oracle_render(blob, packet=1)
    diff = got != ref
    frac = diff.mean()
    assert frac <= 1e-3, (name, frac)
    if diff.any():
        for sh in (0, 8, 16):
            d = np.abs(((got >> sh) & 255).astype(int) - ((ref >> sh) & 255).astype(int))
            assert d.max() <= 1, (name, sh, d.max())


@pytest.mark.parametrize("name", GOLDEN_T)
def test_oracle_hit_distance_is_the_reference_s(entry, name):
    """Dump mode pinned to the reference: ctx_T_BUF(0) as the reference itself
    stores it at XX_end (tracer.cpp:5161; patched scratch build, oracle/Makefile
    `tdump`) against the oracle's, with the reference's packets and per sample."""
    z = np.load(entry.os.path.join(entry.GOLDEN_DIR, name + ".npz"))
    blob, ref, want = z["blob"], z["frame"], z["t"]
    for packet in (32, 1):
        got, t, _ = entry.oracle_render(blob, packet=packet, want_t=True)
        assert int((got != ref).sum()) == 0
        t = t.reshape(want.shape)
        rel = np.abs(t - want) / np.abs(want)
        assert rel.max() <= T_REL_TOL, (name, packet, float(rel.max()))
        assert np.array_equal(t.view(np.uint32), want.view(np.uint32)), (name, packet)   # in fact identical


def test_oracle_other_packet_widths_agree(entry):
    """128/256/512-bit targets of the reference give identical frames
    (SURVEY.md section 6); so do the emulated widths."""
    blob, ref, _ = entry.load_golden("test17_full_a4")
    for packet in (4, 8, 16, 64):
        got, _, _ = entry.oracle_render(blob, packet=packet)
        assert int((got != ref).sum()) == 0, packet


def test_oracle_row_range_and_dump(entry):
    blob, ref, meta = entry.load_golden("test01_full")
    got, t, _ = entry.oracle_render(blob, packet=1, y0=96, y1=160, want_t=True)
    assert np.array_equal(got[96:160], ref[96:160])
    assert not got[:96].any() and not got[160:].any()
    band = t[96:160]
    assert np.isfinite(band).any() and (band[np.isfinite(band)] > 0).all()


def test_oracle_rejects_bad_blob(entry):
    blob, _, _ = entry.load_golden("test01_full")
    bad = blob.copy()
    bad[0] ^= 0xFF
    with pytest.raises(RuntimeError):
        entry.oracle_render(bad)
    with pytest.raises(RuntimeError):
        entry.oracle_render(blob[:100])


@pytest.mark.parametrize("name", GOLDEN_HASHED)
def test_oracle_rows_of_full_size_fixtures(entry, name):
    """The full-size fixtures (1080p / 4K, reference frame kept as row CRCs):
    a few rows through the oracle reproduce the reference's rows exactly."""
    blob, rowcrc, meta = entry.load_golden_hashed(name)
    y0 = (meta["y_res"] // 2) & ~7
    got, _, _ = entry.oracle_render(blob, packet=32, y0=y0, y1=y0 + 4)
    assert np.array_equal(entry.row_crcs(got[y0:y0 + 4]), rowcrc[y0:y0 + 4]), meta["args"]




@pytest.mark.parametrize("name", ["test12_full", "test03_full", "test16_none_a4", "demo02_a4g", "synth1k_a4",
                                  "synth400_metal"])
def test_flattener_of_this_tree_feeds_the_reference_frame(entry, tmp_path, name):
    """Host logic without a GPU: the reference ENGINE + this tree's flattener
    (quadray-engine_b200/host/qr_flatten.cpp) + the oracle, linked as
    oracle/_ref/qr_oracle_harness, renders the reference's frame -- so a change
    of the flattener (element order, prefetching, pointer map) is checked on
    CPU before it meets a GPU.  The binary needs the reference's sources to be
    built (build container); it is skipped where it does not exist."""
    import json
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "oracle", "_ref", "qr_oracle_harness")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/qr_oracle_harness was not built")
    _, ref, meta = entry.load_golden(name)
    out = str(tmp_path / "frame.raw")
    env = dict(os.environ)
    env["QR_ORACLE_PACKET"] = "32"
    p = subprocess.run([exe] + meta["args"].split() + ["-o", out], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-1000:]
    info = json.loads(p.stdout.decode().strip().splitlines()[-1])
    got = np.fromfile(out, dtype=np.uint32).reshape(info["y_res"], info["x_res"])
    assert got.shape == ref.shape
    assert int((got != ref).sum()) == 0, meta["args"]
