"""Parser for the scene blob of include/qr_scene_blob.h (tests / tools only)."""
import struct
import numpy as np

MAGIC = 0x31425251
VERSION = 3

_HDR_FMT = "<4I" + "12i" + "I4f3i" + "f3f3f3f4f4f3ffIi" + "iIiIiIiIiIiI4i"
_HDR_NAMES = (
    ["magic", "version", "total_bytes", "flags"]
    + ["x_res", "y_res", "x_row", "fsaa", "depth", "tile_w", "tile_h", "tls_row", "tls_col", "lst_head", "pad0a", "pad0b"]
    + ["ctx_flags", "t_min", "org0", "org1", "org2", "off_bounds", "n_bounds", "p1c"]
    + ["cam_t_max"] + [f"dir{i}" for i in range(3)] + [f"hor{i}" for i in range(3)] + [f"ver{i}" for i in range(3)]
    + [f"hor_a{i}" for i in range(4)] + [f"ver_a{i}" for i in range(4)] + [f"amb{i}" for i in range(3)]
    + ["cam_clamp", "cam_cmask", "pad2"]
    + ["n_surf", "off_surf", "n_mat", "off_mat", "n_lgt", "off_lgt", "n_elem", "off_elem",
       "n_tiles", "off_tiles", "n_texels", "off_texels", "p3a", "p3b", "p3c", "p3d"]
)
assert struct.calcsize(_HDR_FMT) == 256, struct.calcsize(_HDR_FMT)


def parse_header(buf):
    vals = struct.unpack_from(_HDR_FMT, buf, 0)
    h = dict(zip(_HDR_NAMES, vals))
    if h["magic"] != MAGIC:
        raise ValueError("not a scene blob")
    if h["version"] != VERSION:
        raise ValueError("blob version %d != %d" % (h["version"], VERSION))
    return h


def sections(buf):
    h = parse_header(buf)
    a = np.frombuffer(buf, dtype=np.uint8)
    out = {"header": h}
    out["surf"] = a[h["off_surf"]: h["off_surf"] + h["n_surf"] * 256].view(np.int32).reshape(-1, 64)
    out["mat"] = a[h["off_mat"]: h["off_mat"] + h["n_mat"] * 128].view(np.int32).reshape(-1, 32)
    out["lgt"] = a[h["off_lgt"]: h["off_lgt"] + h["n_lgt"] * 64].view(np.int32).reshape(-1, 16)
    out["elem"] = a[h["off_elem"]: h["off_elem"] + h["n_elem"] * 16].view(np.int32).reshape(-1, 4)
    out["tiles"] = a[h["off_tiles"]: h["off_tiles"] + h["n_tiles"] * 4].view(np.int32)
    out["texels"] = a[h["off_texels"]: h["off_texels"] + h["n_texels"] * 4].view(np.uint32)
    return out


if __name__ == "__main__":
    import sys
    b = open(sys.argv[1], "rb").read()
    s = sections(b)
    h = s["header"]
    print({k: h[k] for k in ("x_res", "y_res", "x_row", "fsaa", "depth", "tile_w", "tile_h", "tls_row", "tls_col",
                             "n_surf", "n_mat", "n_lgt", "n_elem", "n_tiles", "n_texels", "total_bytes")})
    # elements reachable from tiles vs the rest
    el = s["elem"]
    seen = np.zeros(len(el), bool)
    for t in s["tiles"]:
        e = t
        while e >= 0 and not seen[e]:
            seen[e] = True
            e = el[e, 3]
    print("tile elems", int(seen.sum()), "other elems", int((~seen).sum()))
    srf = s["surf"]
    print("a_map[L] vs a_sgn[L] consistent:", bool(np.all((srf[:, 35] != 0) == (srf[:, 39] != 0))),
          "tags", np.unique(srf[:, 43], return_counts=True))
