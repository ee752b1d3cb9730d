#!/usr/bin/env python
"""bench.py -- headline benchmark of the render0 path.

Workload (BASELINE.json configs[1]): RooT's default demo scene scn_demo03,
1920x1080, 4x antialiasing + gamma, camera 0, time 0 -- the scene blob of
tests/golden/demo03_1080p_a4g.npz (flattened from the reference engine by
tools/make_golden.py; synthetic in the sense that nothing is loaded from a
dataset: the scene is the reference's statically linked demo data).

A step = one render0 pass over one frame (8 294 400 primary samples).

  value   total rays / s (primary + shadow + reflection + refraction rays of the
          shade-once algorithm, 36.2 M per frame) with the scene already in
          HBM; device time from CUDA events on the launching stream
  e2e     same metric through the C ABI with HOST buffers: qr_scene_upload
          (pinned staging + H2D) + qr_render (kernel + D2H of the frame)
  roofline  algorithmic IEEE fp32 operations per frame (counted by the device
          core compiled for the host, tools/make_golden.py) / kernel time,
          against the measured non-FMA FP32 rate of this GPU (qr_fp32_peak)
  cpu_baseline  the UNMODIFIED reference (oracle/_ref/qr_ref_harness, AVX-512
          if the host has it, all host threads) on the same frame, render-only

  --impl reference   times only the reference's CPU implementation.

N > 1 (torchrun, one rank per GPU): strong scaling of one frame.  Tile rows are
dealt round-robin (rank r renders rows r, r + N, ...); every rank stores its
pixels straight into rank 0's framebuffer over NVLink (the buffer is shared
through a CUDA IPC handle, qr_frame_ipc_*); the last warp of every rank's
kernel bumps a counter behind that framebuffer (over NVLink as well) and rank
0's stream waits for it (qr_render_rows_notify / qr_wait_notify): no
collective on the data path.  --gather nccl renders into a local buffer and
gathers the rows with one NCCL gather instead.
e2e at N > 1: every rank uploads the scene from host memory and renders its
rows straight into ONE page-locked host frame shared by all ranks (POSIX
shared memory), two frames in flight; the pixels leave every GPU over its own
PCIe link.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "demo03_1080p_a4g"
WORKLOAD_DESC = "scn_demo03 (RooT default demo scene) 1920x1080 4xAA + gamma, camera 0, t=0"
METRIC = "Mrays/s at 1080p 4xAA (frame ms = ms_per_step)"
UNIT = "Mrays/s"
REF_HARNESS = os.path.join(ROOT, "oracle", "_ref", "qr_ref_harness")
REF_ARGS = ["-s", "demo03", "-x", "1920", "-y", "1080", "-a", "2", "-g", "-u"]
REF_ARGS_UPDATE = ["-s", "demo03", "-x", "1920", "-y", "1080", "-a", "2", "-g"]


def make_config(meta, world, gather):
    """The workload description: IDENTICAL in both arms (the driver compares
    it), so it holds no measured value."""
    return {
        "workload": WORKLOAD_DESC, "x_res": meta["x_res"], "y_res": meta["y_res"], "fsaa": "4x", "gamma": True,
        "rays_per_frame": meta["rays"]["total"],
        "primary_samples_per_frame": meta["rays"]["primary"],
        "n_gpus": world,
        "l2": "GPU arm: flushed between timed iterations (256 MB fill); reference arm: host caches as they are",
        "sharding": "single GPU" if world == 1 else
                    "GPU arm: tile rows dealt round-robin over %d ranks (%s); reference arm: rank 0's host cores"
                    % (world, "P2P stores into rank 0's framebuffer over NVLink" if gather == "p2p" else "NCCL gather"),
        "timing": "GPU arm: CUDA events per step on the launching stream, summed, max over ranks; "
                  "reference arm: wall clock per frame inside the harness, median",
    }


# ---------------------------------------------------------------- helpers ---

class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for raw in self.proc.stdout:
            self.lines.append(raw.decode(errors="replace").strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def kernel_source_hash():
    """sha256 over the sources the render kernel is compiled from."""
    import hashlib
    hsh = hashlib.sha256()
    for f in ("csrc/qr_b200.cu", "csrc/qr_core.cuh", "csrc/qr_kscene.h"):
        hsh.update(open(os.path.join(ROOT, "quadray-engine_b200", f), "rb").read())
    return hsh.hexdigest()


def load_workload():
    import __graft_entry__ as ge
    blob, ref_frame, meta = ge.load_golden(WORKLOAD)
    return ge, blob, ref_frame, meta


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(frames, warmup, threads, ref_args=None, extra=()):
    """The unmodified reference on the host cores; returns dict or None."""
    if not os.path.exists(REF_HARNESS):
        return None
    cmd = [REF_HARNESS] + list(ref_args or REF_ARGS) + ["-t", str(threads), "-f", str(frames), "-w", str(warmup)] + list(extra)
    try:
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900, check=True)
        return json.loads(out.stdout.decode().strip().splitlines()[-1])
    except Exception as exc:      # binary not runnable on this host
        sys.stderr.write("reference harness failed: %r\n" % (exc,))
        return None


def run_port(blob, meta, rows):
    """Fallback CPU baseline: the oracle port, one thread, a band of rows."""
    import __graft_entry__ as ge
    t0 = time.perf_counter()
    _, _, st = ge.oracle_render(blob, packet=32, y0=0, y1=rows)
    dt = time.perf_counter() - t0
    frac = rows / float(meta["y_res"])
    return dt / frac


def cpu_baseline(blob, meta, frames, warmup, variants=False):
    """cpu_baseline object: the reference timed on this box's host cores."""
    rays = meta["rays"]["total"]
    threads = min(host_threads(), 120)
    ref = run_reference(frames, warmup, threads)
    if ref is not None:
        ms = ref["ms_med"]
        out = {"value": rays / (ms * 1e-3) / 1e6, "unit": UNIT, "cores": ref["threads"],
               "kind": "reference", "frame_ms": ms, "frame_ms_min": ref["ms_min"],
               "simd": ref["simd"],
               "sample": "%d frames (+%d warm-up) of the same 1080p 4xAA demo03 frame, render-only "
                         "(lists frozen with RT_OPTS_UPDATE_EXT0), %d pinned threads, target %s"
                         % (frames, warmup, ref["threads"], ref["simd"])}
        if variants:
            # the plain AVX-512 target, one thread, and
            # update + render (what rt_Scene::render costs a caller) beside it
            var = {}
            for name, ra, th, extra, nfr in (
                    ("avx512_512x1v2_all_threads_render_only", REF_ARGS, threads, ["-n", "512", "-k", "1", "-v", "2"], 20),
                    ("best_target_1_thread_render_only", REF_ARGS, 1, [], 5),
                    ("best_target_all_threads_update_and_render", REF_ARGS_UPDATE, threads, [], 20)):
                r = run_reference(nfr, 2, th, ra, extra)
                if r is not None:
                    var[name] = {"frame_ms": r["ms_med"], "threads": r["threads"], "simd": r["simd"],
                                 "value": rays / (r["ms_med"] * 1e-3) / 1e6, "frames": nfr}
            out["variants"] = var
        return out, ms
    rows = 64
    sec = run_port(blob, meta, rows)
    return {"value": rays / sec / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
            "frame_ms": sec * 1e3,
            "sample": "oracle port (32-lane packet emulation), rows 0..%d of the frame, scaled" % rows}, sec * 1e3


# -------------------------------------------------------- reference arm ---

_JSON_FD = None


def emit(line):
    """The JSON line, on the real stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main_reference(args, rank, world):
    if rank != 0:
        return 0
    ge, blob, _, meta = load_workload()
    rays = meta["rays"]["total"]
    base, ms = cpu_baseline(blob, meta, max(args.steps, 1), max(args.warmup, 0))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(meta, max(args.gpus, 1), args.gather),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# --------------------------------------------------------------- our arm ---

class _DevArray(object):
    """Raw device pointer -> torch view (through __cuda_array_interface__)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i4",
                                         "data": (int(ptr), False), "version": 2}


def main_gpu(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 backend has no CPU fallback")

    ge, blob, ref_frame, meta = load_workload()
    pkg = ge.load_package()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    ctx = pkg.Context([local_rank])
    ctx.upload(blob)
    ctx.sync()
    hd = ctx.header
    h, w, x_row = hd["y_res"], hd["x_res"], max(hd["x_row"], hd["x_res"])
    tile_h = hd["tile_h"]
    tls_col = (h + tile_h - 1) // tile_h
    slots = (tls_col + world - 1) // world              # tile rows per rank, rounded up

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    host_frame = torch.zeros((h, x_row), dtype=torch.int32).pin_memory()
    qstream = torch.cuda.ExternalStream(ctx.stream(0), device=dev)
    cur = torch.cuda.current_stream(dev)
    token = torch.zeros(1, dtype=torch.int32, device=dev)

    # ---- where the pixels go ------------------------------------------------
    gather = args.gather if world > 1 else "none"
    frame_d = None                      # rank 0 (or single GPU): the assembled frame, (h, x_row) int32
    remote_ptr = None
    if gather == "p2p":
        # rank 0's library-owned framebuffer, opened by everybody else
        handle = torch.zeros(64, dtype=torch.uint8, device=dev)
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            if rank == 0:
                handle.copy_(torch.frombuffer(bytearray(ctx.frame_ipc_export()), dtype=torch.uint8))
        except Exception as exc:
            sys.stderr.write("rank 0: CUDA IPC export failed (%r)\n" % (exc,))
            ok.zero_()
        dist.broadcast(handle, src=0)
        try:
            if rank == 0:
                frame_ptr, _ = ctx.frame_device()
            else:
                remote_ptr = ctx.frame_ipc_open(handle.cpu().numpy().tobytes())
                frame_ptr = remote_ptr
        except Exception as exc:                    # IPC not permitted on this box
            sys.stderr.write("rank %d: CUDA IPC open failed (%r)\n" % (rank, exc))
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if remote_ptr is not None:
                ctx.frame_ipc_close(remote_ptr)
                remote_ptr = None
            gather = "nccl"
            if rank == 0:
                sys.stderr.write("bench.py: falling back to the NCCL gather\n")
        elif rank == 0:
            frame_d = torch.as_tensor(_DevArray(frame_ptr, (h, x_row)), device=dev)
    notify_ptr = None
    frames_done = [0]
    if gather == "p2p":
        # completion counter 0 behind rank 0's framebuffer, as every rank maps it
        notify_ptr = ctx.frame_notify_slot(frame_ptr, 0)
        if rank == 0:
            torch.as_tensor(_DevArray(notify_ptr, (1,)), device=dev).zero_()
        torch.cuda.synchronize(dev)
        dist.barrier()
    if gather == "nccl":
        # frame padded to whole slots; tile row t = k * world + r sits at view[k, r]
        full_d = torch.zeros((slots * world * tile_h, x_row), dtype=torch.int32, device=dev)
        frame_ptr = full_d.data_ptr()
        frame_d = full_d[:h]
        view = full_d.view(slots, world, tile_h, x_row)
        mine_d = torch.zeros((slots, tile_h, x_row), dtype=torch.int32, device=dev)
        gather_d = [torch.zeros_like(mine_d) for _ in range(world)] if rank == 0 else None
    if gather == "none":
        frame_d = torch.zeros((h, x_row), dtype=torch.int32, device=dev)
        frame_ptr = frame_d.data_ptr()

    def finish_frame():
        """After this rank's kernel: make the frame complete on rank 0 (the one
        exchange step per frame; --gather nccl only, the P2P path signals from
        the kernel)."""
        if gather == "nccl":
            mine_d.copy_(view[:, rank])
            dist.gather(mine_d, gather_list=gather_d, dst=0)
            if rank == 0:
                for r in range(1, world):
                    view[:, r].copy_(gather_d[r])

    def render_shard():
        """This rank's tile rows of one frame; with the P2P gather the kernel
        bumps the counter behind rank 0's frame and rank 0's stream waits until
        all ranks have (frames so far * ranks)."""
        if gather == "p2p":
            frames_done[0] += 1
            ctx.render_rows_notify(frame_ptr, x_row, rank, world, notify_ptr)
            if rank == 0:
                ctx.wait_notify(notify_ptr, frames_done[0] * world)
        else:
            ctx.render_rows(frame_ptr, x_row, rank, world)

    def step_device(ev0, ev1):
        flush.fill_(rank + 1)                       # evict L2 between timed iterations
        qstream.wait_stream(cur)
        ev0.record(qstream)
        render_shard()
        if gather == "nccl":
            cur.wait_stream(qstream)
            finish_frame()
            ev1.record(cur)
        else:
            ev1.record(qstream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing -------------------------------------------
    for _ in range(max(args.warmup, 3)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        step_device(e0, e1)
    barrier()
    ctx.ray_counts()
    launches0 = ctx.launch_count()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for e0, e1 in evs:
        step_device(e0, e1)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    kern_ms = None
    launches = ctx.launch_count() - launches0
    counts = ctx.ray_counts()

    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    rays_t = torch.tensor([float(sum(counts.values()))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays_t, op=dist.ReduceOp.SUM)
    dev_ms = float(t.item())
    rays_measured = float(rays_t.item()) / args.steps
    rays = meta["rays"]["total"]

    if os.environ.get("QR_BENCH_DIAG"):
        # diagnostics (stderr): this rank's kernel-only time of its share, frame by frame,
        # with the ranks in lockstep (barrier per frame) and with the L2 flush as in the timed loop
        ks = []
        for _ in range(10):
            flush.fill_(rank + 1)
            barrier()
            qstream.wait_stream(cur)
            render_shard()
            ctx.sync()
            ks.append(ctx.last_render_ms())
            barrier()
        sys.stderr.write("diag rank %d: share kernel ms min %.3f med %.3f max %.3f; timed loop mean %.3f ms/step\n"
                         % (rank, min(ks), sorted(ks)[len(ks) // 2], max(ks), dev_ms / args.steps))

    # parity of what was just timed (rank 0 holds the assembled frame)
    parity = None
    if rank == 0:
        got = frame_d[:, :w].cpu().numpy().view(np.uint32)
        parity = {"pixels_differ_vs_reference_cpu_frame": int((got != ref_frame).sum()), "pixels": int(got.size)}

    # ---- kernel-only duration for the roofline (N = 1) ----------------------
    roofline = None
    peak = None
    if world == 1:
        ks = []
        for _ in range(min(args.steps, 20)):
            flush.fill_(1)
            qstream.wait_stream(cur)
            ctx.render_rows(frame_ptr, x_row, 0, 1)
            ctx.sync()
            ks.append(ctx.last_render_ms())
        kern_ms = statistics.mean(ks)
        peak = ctx.fp32_peak()
        ops = meta["ieee_ops"]["total"]
        # DRAM traffic of the kernel comes from the committed ncu capture -- only
        # if that capture is of THIS kernel: same sources (hash), same launch shape
        traffic, traffic_note = None, "no profiles/ncu_summary.json"
        ncu_json = os.path.join(ROOT, "profiles", "ncu_summary.json")
        if os.path.exists(ncu_json):
            try:
                nj = json.load(open(ncu_json))
                kinfo = ctx.kernel_info()
                if nj.get("kernel_source_sha256") != kernel_source_hash():
                    traffic_note = "ncu_summary.json was captured from other kernel sources (%s...)" \
                                   % str(nj.get("kernel_source_sha256"))[:12]
                elif nj.get("threads_per_cta") != kinfo["threads_per_cta"]:
                    traffic_note = "ncu_summary.json was captured at another launch shape"
                else:
                    traffic = nj.get("dram_bytes_per_launch")
                    traffic_note = "ncu --set full, commit %s, %s" % (nj.get("commit"), nj.get("kernel"))
            except Exception as exc:
                traffic_note = "ncu_summary.json unreadable: %r" % (exc,)
        achieved = ops / (kern_ms * 1e-3) / 1e12
        roofline = {
            "bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_note,
            "kernel": "qr_render_kernel", "kernel_ms": kern_ms,
            "algorithmic_ops_per_launch": ops,
            "peak_source": "measured on this GPU by qr_fp32_peak (separately rounded FMUL+FADD, no FMA; "
                           "MEASURED_PEAKS.json has no FP32 figure); nominal 148 SM x 128 lanes x 1.965 GHz = 37.2",
            "hbm_note": "framebuffer write is %.1f MB per launch = %.1f GB/s, far from the HBM bound"
                        % (h * x_row * 4 / 1e6, h * x_row * 4 / (kern_ms * 1e-3) / 1e9),
        }
        # the HBM view of the same kernel, against the driver-measured copy bandwidth
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            hbm_src = "MEASURED_PEAKS.json"
        except Exception:
            pass
        hbm_bytes = h * w * 4 + int(blob.size)          # framebuffer out + scene image in, per launch
        roofline["hbm"] = {"algorithmic_bytes_per_launch": hbm_bytes,
                           "achieved": hbm_bytes / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                           "frac": hbm_bytes / (kern_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src}

    # ---- end to end through the C ABI with host buffers ----------------------
    blob_h = np.ascontiguousarray(blob)
    hf = host_frame.numpy().view(np.uint32)
    h2d = int(blob_h.size)
    d2h = int(h * w * 4) if rank == 0 else 0

    e2e_path_n = None
    if world == 1:
        def step_e2e():
            ctx.upload(blob_h)
            ctx.render(hf, x_row)

        for _ in range(3):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_parity = int((hf[:, :w] != ref_frame).sum())
    else:
        # ONE page-locked host frame (two of them: two frames in flight) shared by
        # all ranks through POSIX shared memory; every rank renders its tile rows
        # straight into it, so the pixels leave each GPU over its own PCIe link
        # while its kernel runs.  Per step and rank: qr_scene_upload (pack +
        # pinned H2D, into the scene slot the GPU is not reading) +
        # qr_render_rows(shared host frame) + wait for the PREVIOUS frame's
        # kernel + a flag in the shared segment; rank 0 sees the previous frame
        # complete when every rank's flag has reached it, and reads it.
        fbytes = h * x_row * 4
        shm_name = "/dev/shm/qr_b200_bench_%s_%s" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "x"))
        ring = None
        if rank == 0:
            ring = pkg.SharedFrameRing(shm_name, h, x_row, world, rank, create=True)
        dist.barrier()
        if rank != 0:
            ring = pkg.SharedFrameRing(shm_name, h, x_row, world, rank, create=False)
        seg_dev = ctx.host_register(ring.host_ptr, ring.nbytes)
        dist.barrier()
        ctx.pipeline(True)
        evq = [torch.cuda.Event(), torch.cuda.Event()]
        seen = [0]

        def run_sharded(n):
            """n frames (numbered from seen + 1); returns after the last one is
            complete in host memory."""
            base = seen[0]
            for i in range(n + 1):
                k = base + i + 1
                if i < n:
                    ctx.upload(blob_h)
                    ring.wait_free(k)                       # frame k - 2 has been consumed
                    ctx.render_rows(seg_dev + ring.frame_offset(k), x_row, rank, world)
                    evq[k & 1].record(qstream)
                if i > 0:
                    evq[(k - 1) & 1].synchronize()          # this rank's rows of frame k - 1 are in host memory
                    ring.mark_done(k - 1)                   # ... says so to rank 0
                    if rank == 0:
                        ring.wait_complete(k - 1)
                        _ = int(ring.frame(k - 1)[h // 2, w // 2])          # the step's result is read on the host
                        ring.release(k - 1)                 # ... and its buffer handed back
            seen[0] = base + n

        run_sharded(3)
        barrier()
        t0 = time.perf_counter()
        run_sharded(args.steps)
        e2e_s = time.perf_counter() - t0
        barrier()
        e2e_parity = int((ring.frame(seen[0])[:, :w] != ref_frame).sum()) if rank == 0 else None
        ctx.pipeline(False)
        ctx.upload(blob_h)
        ctx.sync()
        ctx.host_unregister(ring.host_ptr)
        d2h = int(sum(min(tile_h, h - t * tile_h) for t in range(rank, tls_col, world)) * w * 4)
        e2e_path_n = ("per step and rank: qr_scene_upload(host blob: pack + pinned H2D, two scene slots) + "
                      "qr_render_rows(its tile rows, stored by the kernel straight into ONE page-locked host frame "
                      "shared by all ranks, over each GPU's own PCIe link) + wait for the previous frame; "
                      "two frames in flight; rank 0 reads the frame when every rank has flagged it")
        dist.barrier()
        ring.close(unlink=(rank == 0))
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    bytes_t = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(bytes_t, op=dist.ReduceOp.SUM)
    e2e_s = float(t.item())
    h2d, d2h = int(bytes_t[0].item()), int(bytes_t[1].item())

    # the same through the pipelined calls (N = 1): scene N + 1 is packed and
    # copied while frame N renders; every step still uploads its scene from
    # host memory and delivers its frame into the host buffer
    e2e_pipe_s = None
    if world == 1:
        ctx.pipeline(True)
        def run_pipe(n):
            prev = None
            for _ in range(n):
                if prev is not None:
                    ctx.render_fetch(prev, hf)          # D2H of the previous frame beside the pack + H2D
                ctx.upload(blob_h)
                tk = ctx.render_begin()
                if prev is not None:
                    ctx.render_end(prev, hf, fetched=True)
                prev = tk
            ctx.render_end(prev, hf)

        run_pipe(3)
        hf[:] = 0
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        run_pipe(args.steps)
        e2e_pipe_s = time.perf_counter() - t0
        e2e_pipe_parity = int((hf[:, :w] != ref_frame).sum())
        ctx.pipeline(False)
        ctx.upload(blob_h)

    # ---- the single-process multi-GPU context (what the drop-in render0 uses
    # with QR_B200_DEVICES=0,1,..): rank 0 opens one over all GPUs of the job
    # while the other ranks idle, renders the 1080p workload and the 4K frame
    # of BASELINE.json config 4, and reports parity and frame time
    single = None
    if world > 1:
        barrier()
        # the other ranks wait on the HOST (gloo): an NCCL barrier would keep a
        # spinning kernel on their GPUs, which rank 0's context is rendering on
        idle = dist.new_group(backend="gloo")
        if rank == 0:
            try:
                c2 = pkg.Context(list(range(world)))
                single = {"devices": world}
                for tag, fx in (("1080p", None), ("4k", "demo03_4k_a4gh")):
                    if fx is None:
                        b2, want, crc = blob, ref_frame, None
                    else:
                        b2, crc, _m = ge.load_golden_hashed(fx)
                        want = None
                    c2.upload(b2)
                    got2 = c2.render_frame()
                    ts = []
                    for _ in range(8):
                        c2.sync()
                        t1 = time.perf_counter()
                        c2.render(None)
                        c2.sync()
                        ts.append((time.perf_counter() - t1) * 1e3)
                    if want is not None:
                        differ = int((got2 != want).sum())
                    else:
                        differ = int((ge.row_crcs(got2) != crc).sum())
                    single[tag] = {"frame_ms": statistics.median(ts),
                                   ("pixels_differ" if want is not None else "rows_differ"): differ}
                c2.close()
            except Exception as exc:
                single = {"error": repr(exc)}
        dist.barrier(group=idle)
        barrier()

    info = ctx.kernel_info()
    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base, _ = cpu_baseline(blob, meta, 100, 3, variants=True)

    if rank == 0:
        ms_per_step = dev_ms / args.steps
        sharding = {"none": "single GPU",
                    "p2p": "tile rows dealt round-robin; every rank stores into rank 0's framebuffer over NVLink "
                           "(CUDA IPC); the last warp of each rank's kernel bumps a counter behind that frame, rank 0's "
                           "stream waits for it (no collective)",
                    "nccl": "tile rows dealt round-robin; one NCCL gather to rank 0 per frame"}[gather]
        line = {
            "metric": METRIC, "value": rays * args.steps / (dev_ms * 1e-3) / 1e6, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "fps": 1e3 / ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "ours",
            "config": make_config(meta, world, gather),
            "rays_per_frame_measured": rays_measured,
            "primary_Msamples_per_s": meta["rays"]["primary"] * args.steps / (dev_ms * 1e-3) / 1e6,
            "sharding": sharding,
            "e2e": {"value": rays * args.steps / e2e_s / 1e6, "unit": UNIT,
                    "ms_per_step": e2e_s / args.steps * 1e3,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "path": "qr_scene_upload(host blob: pack + pinned H2D) + qr_render(page-locked host frame, written by the kernel over PCIe)"
                            if world == 1 else e2e_path_n,
                    "pixels_differ_vs_reference_cpu_frame": e2e_parity},
            "gpu_launches": launches,
            "clocks": clocks,
            "parity": parity,
            "kernel": info,
            "wall_s_timed_region": wall,
        }
        if e2e_pipe_s is not None:
            # throughput of the pipelined calls is the end-to-end figure; the
            # synchronous call sequence (one frame's latency) stays beside it
            sync = line["e2e"]
            line["e2e"] = {
                "value": rays * args.steps / e2e_pipe_s / 1e6, "unit": UNIT,
                "ms_per_step": e2e_pipe_s / args.steps * 1e3,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "path": "per step: qr_render_fetch(previous frame -> page-locked host frame, DMA) + "
                        "qr_scene_upload(host blob: pack + pinned H2D) + qr_render_begin + qr_render_end(previous "
                        "frame); two frames in flight",
                "pixels_differ_vs_reference_cpu_frame": e2e_pipe_parity,
                "synchronous": {"value": sync["value"], "ms_per_step": sync["ms_per_step"], "path": sync["path"],
                                "pixels_differ_vs_reference_cpu_frame": sync["pixels_differ_vs_reference_cpu_frame"]},
            }
        if single is not None:
            line["single_process_ctx"] = single
        if roofline is not None:
            line["roofline"] = roofline
        if base is not None:
            line["cpu_baseline"] = base
        emit(line)

    if remote_ptr is not None:
        ctx.frame_ipc_close(remote_ptr)
    if world > 1:
        dist.barrier()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: how the tile rows reach rank 0 (default: P2P stores over NVLink)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # stdout carries the ONE JSON line and nothing else: libraries that write to
    # file descriptor 1 on their own (NCCL prints its version there) go to stderr
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if world == 1 and args.gpus > 1:
        sys.stderr.write("bench.py: --gpus %d without torchrun: running the single-process multi-GPU "
                         "context is not the benchmark contract; launch with torch.distributed.run\n" % args.gpus)

    if args.impl == "reference":
        return main_reference(args, rank, world)
    return main_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
