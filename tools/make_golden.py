#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/.

Runs in the build container only (it needs /root/reference through the two
binaries oracle/Makefile builds):

  oracle/_ref/qr_ref_harness     the UNMODIFIED reference core -> reference frame
  oracle/_ref/qr_oracle_harness  reference engine + this repo's flattener
                                 -> scene blob of the very same frame

Each fixture is tests/golden/<name>.npz with
  blob   uint8   scene blob (include/qr_scene_blob.h)
  frame  uint32  y_res x x_res reference frame (0x00RRGGBB), rendered by the
                 reference's auto-selected SIMD target (512x2v2 here)
  meta   json    harness arguments, reference target, ray counts

The reference ships no golden images, so these are the pins.
"""
import ctypes
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "qr_ref_harness")
ORC = os.path.join(ROOT, "oracle", "_ref", "qr_oracle_harness")
OUT = os.path.join(ROOT, "tests", "golden")

CASES = {}
for i in range(1, 19):
    CASES["test%02d_full" % i] = "-s test%02d -p full" % i
CASES.update({
    "test01_full_a4": "-s test01 -p full -a 2",
    "test14_full_a4": "-s test14 -p full -a 2",
    "test15_full_a2": "-s test15 -p full -a 1",
    "test17_full_a4": "-s test17 -p full -a 2",
    "test18_full_a4": "-s test18 -p full -a 2",
    "test14_none":    "-s test14 -p none",
    "test16_none_a4": "-s test16 -p none -a 2",
    "test05_odd":     "-s test05 -p full -x 403 -y 250 -a 2",
    "demo01_a4g":     "-s demo01 -a 2 -g",
    "demo02_a4g":     "-s demo02 -a 2 -g",
    "demo03_a4g":     "-s demo03 -a 2 -g -b 1000",
    "demo03_1080p_a4g": "-s demo03 -x 1920 -y 1080 -a 2 -g",
    # BASELINE.json config 5 in miniature: generated quadric clouds under 8-ary
    # bounding-volume arrays (apps/qr_synth_scene.h), with and without mirrors
    "synth1k_a4":     "-s synth -N 1000 -E 25 -x 480 -y 270 -a 2",
    "synth400_metal": "-s synth -N 400 -E 16 -M 300 -S 7 -x 400 -y 240",
    # BASELINE.json config 3 as stated: Fresnel (and gamma) props ON, i.e.
    # RT_OPTS_FRESNEL / RT_OPTS_GAMMA cleared (format.h:59-60, 73-75)
    "test02_full_rg":   "-s test02 -p full -r -g",
    "test09_full_rg":   "-s test09 -p full -r -g",
    "test12_full_rg":   "-s test12 -p full -r -g",
    "test14_full_rg":   "-s test14 -p full -r -g",
    "test15_full_a4rg": "-s test15 -p full -a 2 -r -g",
    "test16_full_a4rg": "-s test16 -p full -a 2 -r -g",
    "test17_full_rg":   "-s test17 -p full -r -g",
    "test18_full_a2rg": "-s test18 -p full -a 1 -r -g",
    "demo01_a4rg":      "-s demo01 -a 2 -r -g",
    "demo02_a4rg":      "-s demo02 -a 2 -r -g",
    "demo03_a4rg":      "-s demo03 -a 2 -r -g -b 1000",
})


# Device-side tiling: the BLOB comes from the engine with
# RT_OPTS_TILING off (every tile head is the camera list, bounding boxes sent
# along, "_nt"), the FRAME from the reference with its own host tiling on.
NOTILE = " -p 0x0230FFB9"       # RT_OPTS_FULL without TILING / TILING_EXT1
CASES_NT = {
    "demo01_a4g_nt":     "-s demo01 -a 2 -g",
    "demo02_a4g_nt":     "-s demo02 -a 2 -g",
    "demo03_a4g_nt":     "-s demo03 -a 2 -g -b 1000",
    "test05_odd_nt":     "-s test05 -x 403 -y 250 -a 2",
    "test12_full_nt":    "-s test12",
    "test14_full_nt":    "-s test14",
    "test15_a2_nt":      "-s test15 -a 1",
    "test16_a4rg_nt":    "-s test16 -a 2 -r -g",
    "test17_a4_nt":      "-s test17 -a 2",
}
CASES.update({k: v + " -p full" for k, v in CASES_NT.items()})
BLOB_ARGS = {k: v + NOTILE for k, v in CASES_NT.items()}


# Full-size cases (BASELINE.json configs 3 and 4): the reference frame is kept
# as per-row CRC-32s instead of pixels ("h" suffix) so the fixtures stay small.
CASES_HASHED = {
    "test02_1080p_a4h": "-s test02 -p full -x 1920 -y 1080 -a 2",
    "test09_1080p_h":   "-s test09 -p full -x 1920 -y 1080",
    "test12_1080p_a4h": "-s test12 -p full -x 1920 -y 1080 -a 2",
    "test14_1080p_h":   "-s test14 -p full -x 1920 -y 1080",
    "test15_1080p_a4h": "-s test15 -p full -x 1920 -y 1080 -a 2",
    "test16_1080p_a4h": "-s test16 -p full -x 1920 -y 1080 -a 2",
    "test17_1080p_a4h": "-s test17 -p full -x 1920 -y 1080 -a 2",
    "test18_1080p_a4h": "-s test18 -p full -x 1920 -y 1080 -a 2",
    "demo03_4k_a4gh":   "-s demo03 -x 3840 -y 2160 -a 2 -g",
    # config 3 with the Fresnel + gamma props on, at the stated 1920x1080
    "test02_1080p_a4rgh": "-s test02 -p full -x 1920 -y 1080 -a 2 -r -g",
    "test09_1080p_a4rgh": "-s test09 -p full -x 1920 -y 1080 -a 2 -r -g",
    "test12_1080p_a4rgh": "-s test12 -p full -x 1920 -y 1080 -a 2 -r -g",
    "test14_1080p_a4rgh": "-s test14 -p full -x 1920 -y 1080 -a 2 -r -g",
    "test15_1080p_a4rgh": "-s test15 -p full -x 1920 -y 1080 -a 2 -r -g",
    "test16_1080p_a4rgh": "-s test16 -p full -x 1920 -y 1080 -a 2 -r -g",
    "test17_1080p_a4rgh": "-s test17 -p full -x 1920 -y 1080 -a 2 -r -g",
    "test18_1080p_a4rgh": "-s test18 -p full -x 1920 -y 1080 -a 2 -r -g",
    "demo01_1080p_a4rgh": "-s demo01 -x 1920 -y 1080 -a 2 -r -g",
    "demo02_1080p_a4rgh": "-s demo02 -x 1920 -y 1080 -a 2 -r -g",
    "demo03_1080p_a4rgh": "-s demo03 -x 1920 -y 1080 -a 2 -r -g",
    # device-side tiling at the bench resolution and at 4K
    "demo03_1080p_a4g_nth": "-s demo03 -x 1920 -y 1080 -a 2 -g -p full",
    "demo03_4k_a4g_nth":    "-s demo03 -x 3840 -y 2160 -a 2 -g -p full",
}
BLOB_ARGS.update({"demo03_1080p_a4g_nth": "-s demo03 -x 1920 -y 1080 -a 2 -g" + NOTILE,
                  "demo03_4k_a4g_nth": "-s demo03 -x 3840 -y 2160 -a 2 -g" + NOTILE})


# Config 5 at size: only the reference frame's row CRC-32s are kept ("c"
# suffix, no blob); the GPU side is rendered through the drop-in harness, which
# generates the very same scene (tests/test_gpu_dropin.py).
CASES_CRC = {
    "synth10k_1080p_a4c": "-s synth -N 10000 -x 1920 -y 1080 -a 2",
    "synth10k_8k_a4c":    "-s synth -N 10000 -x 7680 -y 4320 -a 2",
    "synth100k_1080p_c":  "-s synth -N 100000 -x 1920 -y 1080",
    # config 5 at the stated size: 100 k quadrics, 7680 x 4320, 4xAA
    "synth100k_8k_a4c":   "-s synth -N 100000 -x 7680 -y 4320 -a 2",
    # the same clouds with SHARED GLOBAL lists (RT_OPTS_RENDER / SHADOW* / 2SIDED* off,
    # engine.cpp:2170-2173, 2483-2486): the engine builds no per-surface lists (its O(n^2)
    # part), every ray walks the camera list's bounding-volume tree
    "synth100k_1080p_glc":  "-s synth -N 100000 -x 1920 -y 1080 -p 0x023001BF",
    "synth10k_8k_a4_glc":   "-s synth -N 10000 -x 7680 -y 4320 -a 2 -p 0x023001BF",
}


# Dump-mode pins ("_T" suffix): primary hit distance per sample written by the
# reference itself -- oracle/_ref/qr_ref_tdump, the reference with one store
# added at XX_end in a scratch copy (oracle/Makefile `tdump`, tracer.cpp:5161).
CASES_T = {
    "test17_q_a4_T": "-s test17 -p full -x 400 -y 240 -a 2",
    "test15_q_a2_T": "-s test15 -p full -x 400 -y 240 -a 1",
    "test14_q_T":    "-s test14 -p full -x 400 -y 240",
    "synth400_q_T":  "-s synth -N 400 -E 16 -M 300 -S 7 -x 400 -y 240",      # has misses (T stays at the camera's t_max)
    "test16_q_T":    "-s test16 -p full -x 400 -y 240 -r -g",
    "demo03_q_a4_T": "-s demo03 -x 400 -y 240 -a 2 -g",
}


# Path tracer ("_pt" suffix): the frame the unmodified reference
# shows after N accumulated frames of rt_Scene::set_pton(1) at a fixed time
# (harness -Q -f N -d 0), and the scene blob (QR_BLOB_PT set).  The reference's
# path-traced frames depend on its SIMD width: these are the 512x2v2 target's.
CASES_PT = {
    "test18_a4_pt":    ("-s test18 -x 160 -y 96 -a 2", 3),
    "test18_a2rg_pt":  ("-s test18 -x 160 -y 96 -a 1 -r -g", 3),
    "test18_q_pt":     ("-s test18 -x 320 -y 200 -r", 4),
    "test17_r_pt":     ("-s test17 -x 160 -y 96 -r", 3),       # packet 1 differs here (Fresnel split under mixed TIR)
    "test02_a2rg_pt":  ("-s test02 -x 160 -y 96 -a 1 -r -g", 3),
    "test05_odd_pt":   ("-s test05 -x 99 -y 40 -a 2", 2),      # x_res not a multiple of the packet
    "test16_none_pt":  ("-s test16 -x 48 -y 30 -a 1 -p none", 2),
    "demo02_rg_pt":    ("-s demo02 -x 96 -y 56 -r -g", 3),
    "demo03_a4rg_pt":  ("-s demo03 -x 96 -y 56 -a 2 -r -g", 2),
}


PT_WIDTHS = ("test18_a4_pt", "test02_a2rg_pt", "test17_r_pt", "demo02_rg_pt")


def main_pt(names):
    for name in names:
        a, frames = CASES_PT[name]
        args = a.split() + ["-Q", "-d", "0"]
        with tempfile.TemporaryDirectory() as td:
            rf, r1, of, bf = (os.path.join(td, n) for n in ("r.raw", "r1.raw", "o.raw", "s.blob"))
            jr = run([REF] + args + ["-f", str(frames), "-o", rf])
            run([REF] + args + ["-f", "1", "-o", r1])
            run([ORC] + args + ["-f", "1", "-o", of], {"QR_DUMP_BLOB": bf, "QR_ORACLE_PACKET": "32"})
            w, h = jr["x_res"], jr["y_res"]
            frame = np.fromfile(rf, dtype=np.uint32).reshape(h, w)
            frame1 = np.fromfile(r1, dtype=np.uint32).reshape(h, w)
            assert np.array_equal(frame1, np.fromfile(of, dtype=np.uint32).reshape(h, w)), name
            blob = np.fromfile(bf, dtype=np.uint8)
            # the same frames from the reference's narrower AVX targets (8 and 16 lanes): the
            # path tracer's result depends on the SIMD width, the oracle follows it at packet = S
            extra = {}
            if name in PT_WIDTHS:
                for lanes, target in ((8, "-n 256 -k 1 -v 2"), (16, "-n 512 -k 1 -v 2")):
                    jw = run([REF] + args + target.split() + ["-f", str(frames), "-o", rf])
                    assert jw["simd"] == {8: "256x1v2", 16: "512x1v2"}[lanes], jw["simd"]
                    extra["frame_w%d" % lanes] = np.fromfile(rf, dtype=np.uint32).reshape(h, w)
        meta = {"name": name, "args": a, "frames": frames, "x_res": w, "y_res": h, "fsaa": jr["fsaa"],
                "opts": jr["opts"], "ref_simd": jr["simd"], "lit": float((frame != 0).mean())}
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, blob=blob, frame=frame, frame1=frame1,
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **extra)
        print("%-20s %4dx%-4d %d frames, lit %.3f  npz %7d B" % (name, w, h, frames, meta["lit"], os.path.getsize(path)))


def main_t(names):
    tdump = os.path.join(ROOT, "oracle", "_ref", "qr_ref_tdump")
    for name in names:
        args = CASES_T[name].split()
        with tempfile.TemporaryDirectory() as td:
            rf, tf, of, bf = (os.path.join(td, n) for n in ("r.raw", "t.raw", "o.raw", "s.blob"))
            jr = run([tdump] + args + ["-o", rf, "-T", tf])
            ju = run([REF] + args + ["-o", of])
            w, h = jr["x_res"], jr["y_res"]
            frame = np.fromfile(rf, dtype=np.uint32).reshape(h, w)
            # the added store does not change the frame the reference renders
            assert np.array_equal(frame, np.fromfile(of, dtype=np.uint32).reshape(h, w)), name
            t = np.fromfile(tf, dtype=np.float32).reshape(h, w << jr["fsaa"])
            run([ORC] + args + ["-o", of], {"QR_DUMP_BLOB": bf, "QR_ORACLE_PACKET": "1", "QR_ORACLE_ROWS": "1"})
            blob = np.fromfile(bf, dtype=np.uint8)
        meta = {"name": name, "args": CASES_T[name], "x_res": w, "y_res": h, "fsaa": jr["fsaa"],
                "opts": jr["opts"], "ref_simd": jr["simd"], "hit_fraction": float(np.isfinite(t).mean())}
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, blob=blob, frame=frame, t=t,
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        print("%-20s %4dx%-4d hit %.3f  npz %7d B" % (name, w, h, meta["hit_fraction"], os.path.getsize(path)))


def main_crc(names):
    for name in names:
        args = CASES_CRC[name].split()
        with tempfile.TemporaryDirectory() as td:
            rf = os.path.join(td, "r.raw")
            jr = run([REF] + args + ["-t", str(os.cpu_count()), "-o", rf])
            w, h = jr["x_res"], jr["y_res"]
            frame = np.fromfile(rf, dtype=np.uint32).reshape(h, w)
        meta = {"name": name, "args": CASES_CRC[name], "x_res": w, "y_res": h, "fsaa": jr["fsaa"],
                "opts": jr["opts"], "ref_simd": jr["simd"], "ref_ms": jr["ms_min"], "ref_threads": jr["threads"],
                "covered": float((frame != 0).mean())}
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, rowcrc=row_crcs(frame),
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        print("%-20s %4dx%-4d ref %.0f ms  npz %7d B" % (name, w, h, jr["ms_min"], os.path.getsize(path)))


def row_crcs(frame):
    import zlib
    return np.array([zlib.crc32(np.ascontiguousarray(r).tobytes()) for r in frame], dtype=np.uint32)


def main_hashed(names):
    for name in names:
        args = CASES_HASHED[name].split()
        with tempfile.TemporaryDirectory() as td:
            rf, of, bf = (os.path.join(td, n) for n in ("r.raw", "o.raw", "s.blob"))
            jr = run([REF] + args + ["-o", rf])
            # the oracle harness only has to flatten: one row is enough
            run([ORC] + BLOB_ARGS.get(name, CASES_HASHED[name]).split() + ["-o", of],
                {"QR_DUMP_BLOB": bf, "QR_ORACLE_PACKET": "1", "QR_ORACLE_ROWS": "1"})
            w, h = jr["x_res"], jr["y_res"]
            frame = np.fromfile(rf, dtype=np.uint32).reshape(h, w)
            blob = np.fromfile(bf, dtype=np.uint8)
        meta = {"name": name, "args": CASES_HASHED[name], "x_res": w, "y_res": h, "fsaa": jr["fsaa"],
                "opts": jr["opts"], "ref_simd": jr["simd"]}
        if name in BLOB_ARGS:
            meta["blob_args"] = BLOB_ARGS[name]
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, blob=blob, rowcrc=row_crcs(frame),
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        print("%-20s %4dx%-4d blob %8d B  npz %7d B" % (name, w, h, blob.size, os.path.getsize(path)))


def run(cmd, env=None):
    e = dict(os.environ)
    if env:
        e.update(env)
    out = subprocess.run(cmd, env=e, check=True, stdout=subprocess.PIPE).stdout.decode()
    return json.loads(out.strip().splitlines()[-1])


def deferred_counts(blob, w, h, fsaa):
    """Rays and IEEE operations of the shade-once algorithm the GPU runs,
    counted by the device core compiled for the host (tests/hostsim)."""
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "hostsim", "libqr_hostsim.so"))
    lib.qr_hostsim_render.restype = ctypes.c_int
    lib.qr_hostsim_render.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    frame = np.zeros((h, w), dtype=np.uint32)
    rays = (ctypes.c_uint64 * 4)()
    ops = (ctypes.c_uint64 * 4)()
    lib.qr_hostsim_ops(ops)
    rc = lib.qr_hostsim_render(blob.ctypes.data, blob.size, frame.ctypes.data, w, None, 0, h, rays)
    assert rc == 0
    lib.qr_hostsim_ops(ops)
    r = dict(zip(("primary", "shadow", "reflect", "refract"), [int(x) for x in rays]))
    r["total"] = sum(r.values())
    o = dict(zip(("addsub", "mul", "div", "sqrt"), [int(x) for x in ops]))
    o["total"] = sum(o.values())
    return frame, r, o


def main(names):
    os.makedirs(OUT, exist_ok=True)
    for name in names:
        args = CASES[name].split()
        with tempfile.TemporaryDirectory() as td:
            rf, of, bf, sf = (os.path.join(td, n) for n in ("r.raw", "o.raw", "s.blob", "st.json"))
            jr = run([REF] + args + ["-o", rf])
            jo = run([ORC] + BLOB_ARGS.get(name, CASES[name]).split() + ["-o", of],
                     {"QR_DUMP_BLOB": bf, "QR_ORACLE_PACKET": "1", "QR_ORACLE_STATS": sf})
            w, h = jr["x_res"], jr["y_res"]
            frame = np.fromfile(rf, dtype=np.uint32).reshape(h, w)
            oframe = np.fromfile(of, dtype=np.uint32).reshape(h, w)
            blob = np.fromfile(bf, dtype=np.uint8)
            stats = json.load(open(sf))
        hframe, rays, ops = deferred_counts(blob, w, h, jr["fsaa"])
        assert np.array_equal(hframe, oframe), name
        meta = {
            "name": name, "args": CASES[name], "x_res": w, "y_res": h, "fsaa": jr["fsaa"],
            "opts": jr["opts"], "ref_simd": jr["simd"],
            "oracle_packet1_mismatch_vs_ref": int((frame != oframe).sum()),
            "oracle_immediate_rays": stats,
            "rays": rays,           # shade-once algorithm (what the GPU casts)
            "ieee_ops": ops,        # algorithmic IEEE fp32 operations per frame
        }
        if name in BLOB_ARGS:
            meta["blob_args"] = BLOB_ARGS[name]
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, blob=blob, frame=frame, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        print("%-20s %4dx%-4d blob %7d B  npz %7d B  oracle(packet=1) != ref: %d px"
              % (name, w, h, blob.size, os.path.getsize(path), meta["oracle_packet1_mismatch_vs_ref"]))


if __name__ == "__main__":
    names = sys.argv[1:] or (list(CASES) + list(CASES_HASHED) + list(CASES_CRC) + list(CASES_T) + list(CASES_PT))
    main([n for n in names if n in CASES])
    main_hashed([n for n in names if n in CASES_HASHED])
    main_crc([n for n in names if n in CASES_CRC])
    main_t([n for n in names if n in CASES_T])
    main_pt([n for n in names if n in CASES_PT])
