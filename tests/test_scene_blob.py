"""Structural checks of the flattened scene blob (include/qr_scene_blob.h) as
produced by the host flattener quadray-engine_b200/host/qr_flatten.cpp."""
import numpy as np
import pytest

import qr_blob
from conftest import GOLDEN_SMALL


@pytest.mark.parametrize("name", GOLDEN_SMALL)
def test_blob_indices_are_in_range(entry, name):
    blob, frame, meta = entry.load_golden(name)
    s = qr_blob.sections(blob.tobytes())
    h = s["header"]
    assert (h["x_res"], h["y_res"]) == (meta["x_res"], meta["y_res"])
    assert h["fsaa"] == meta["fsaa"] and 0 <= h["depth"] <= 10
    assert h["tls_row"] * h["tile_w"] >= h["x_res"] and h["tls_col"] * h["tile_h"] >= h["y_res"]
    assert h["n_tiles"] == h["tls_row"] * h["tls_col"]
    ne, ns, nm, nl = h["n_elem"], h["n_surf"], h["n_mat"], h["n_lgt"]
    el, sf, mt = s["elem"], s["surf"], s["mat"]
    assert ((s["tiles"] >= -1) & (s["tiles"] < ne)).all()
    assert ((el[:, 3] >= -1) & (el[:, 3] < ne)).all()          # next
    assert ((el[:, 1] >= -1) & (el[:, 1] < ne)).all()          # data_p
    assert (el[:, 2] < max(ns, nl)).all()
    # surface record: trnode, clip list, materials, lists
    assert ((sf[:, 19] >= -1) & (sf[:, 19] < ns)).all()        # trnode
    assert ((sf[:, 23] >= -1) & (sf[:, 23] < ne)).all()        # clip_head
    assert ((sf[:, 32:35] >= 0) & (sf[:, 32:35] < 6)).all()    # a_map I,J,K
    assert ((sf[:, 35] >= 0) & (sf[:, 35] <= 3)).all()         # a_map L
    assert np.isin(sf[:, 36:39], (0, 1)).all() and np.isin(sf[:, 39], (0, 3)).all()
    assert ((sf[:, 44:46] >= -1) & (sf[:, 44:46] < nm)).all()  # materials
    assert ((sf[:, 48:52] >= -1) & (sf[:, 48:52] < ne)).all()  # light / surface lists
    # the field shift and the transform flag go together (the kernels rely on it)
    real = sf[:, 43] != 9
    assert ((sf[real, 35] != 0) == (sf[real, 39] != 0)).all()
    # materials: texture rectangle inside the texel pool
    size = (mt[:, 4].astype(np.int64) + 1) * (mt[:, 5].astype(np.int64) + 1)
    assert ((mt[:, 7] >= 0) & (mt[:, 7] + size <= h["n_texels"])).all()
    assert (mt[:, 6] == np.log2(mt[:, 4] + 1).astype(int)).all()


def test_lists_terminate(entry):
    blob, _, _ = entry.load_golden("test14_full")
    s = qr_blob.sections(blob.tobytes())
    el = s["elem"]
    for head in list(s["tiles"][:200]) + list(s["surf"][:, 48:52].ravel()):
        n, e = 0, int(head)
        while e >= 0:
            e = int(el[e, 3])
            n += 1
            assert n <= len(el)
