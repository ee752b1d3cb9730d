"""quadray-engine_b200 -- B200-native backend for QuadRay's core/tracer render0 path.

The product is native: CUDA kernels + C ABI in ``lib/libquadray_b200.so``
(``include/quadray_b200.h``) behind the replacement tracer translation unit
``host/tracer_b200.cpp``.  This module is only the thin ctypes binding that the
tests and ``bench.py`` use to call the very same C ABI from Python; it holds no
rendering logic and has no CPU path -- loading fails loudly when the library
has not been built, and ``Context()`` fails when there is no CUDA device.

The directory name contains a dash, so import it with
``__graft_entry__.load_package()`` (importlib by path).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# QR_B200_LIB: another build of the same library (kernel A/B runs: tools/tune_shapes.py)
LIB_PATH = os.environ.get("QR_B200_LIB") or os.path.join(_HERE, "lib", "libquadray_b200.so")

QR_OK, QR_E_ARG, QR_E_BLOB, QR_E_CUDA, QR_E_NODEV, QR_E_STATE = 0, -1, -2, -3, -4, -5

# every entry point include/quadray_b200.h declares
SYMBOLS = (
    "qr_init", "qr_shutdown", "qr_last_error", "qr_scene_upload", "qr_render",
    "qr_render_device", "qr_render_rows", "qr_sync", "qr_frame_device",
    "qr_frame_ipc_export", "qr_frame_ipc_open", "qr_frame_ipc_close", "qr_dump_hits",
    "qr_ray_counts", "qr_last_render_ms", "qr_stream", "qr_launch_count",
    "qr_kernel_query", "qr_fp32_peak", "qr_pipeline", "qr_render_begin", "qr_render_fetch", "qr_render_end",
    "qr_frame_notify_slot", "qr_render_rows_notify", "qr_wait_notify", "qr_host_register", "qr_host_unregister",
    "qr_check_counters",
    "qr_pt_reset", "qr_pt_frames", "qr_pt_fetch",
)


class QuadRayError(RuntimeError):
    def __init__(self, code, text):
        RuntimeError.__init__(self, "quadray_b200 error %d: %s" % (code, text))
        self.code = code


class KernelInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "sm_count", "threads_per_cta", "ctas_per_sm", "regs_per_thread",
        "local_bytes_per_thread", "smem_static_bytes", "smem_dynamic_bytes",
        "scene_in_smem", "device_tiling")]


_lib = None


def load_library():
    """dlopen the C-ABI library; raises if it was not built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run __graft_entry__.build() "
                          "(nvcc, sm_100a); there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    lib.qr_init.argtypes = [ctypes.POINTER(ci), ci, ctypes.POINTER(vp)]
    lib.qr_init.restype = ci
    lib.qr_shutdown.argtypes = [vp]
    lib.qr_shutdown.restype = None
    lib.qr_last_error.argtypes = [vp]
    lib.qr_last_error.restype = ctypes.c_char_p
    lib.qr_scene_upload.argtypes = [vp, vp, sz]
    lib.qr_scene_upload.restype = ci
    lib.qr_render.argtypes = [vp, vp, ci]
    lib.qr_render.restype = ci
    lib.qr_render_device.argtypes = [vp, vp, ci, ci, ci]
    lib.qr_render_device.restype = ci
    lib.qr_render_rows.argtypes = [vp, vp, ci, ci, ci]
    lib.qr_render_rows.restype = ci
    lib.qr_check_counters.argtypes = [vp, ctypes.POINTER(ctypes.c_uint32)]
    lib.qr_check_counters.restype = ci
    lib.qr_pt_reset.argtypes = [vp, vp, ctypes.c_size_t]
    lib.qr_pt_reset.restype = ci
    lib.qr_pt_frames.argtypes = [vp]
    lib.qr_pt_frames.restype = ci
    lib.qr_pt_fetch.argtypes = [vp, vp, vp, vp, vp, ctypes.c_size_t]
    lib.qr_pt_fetch.restype = ci
    lib.qr_host_register.argtypes = [vp, vp, sz, ctypes.POINTER(vp)]
    lib.qr_host_register.restype = ci
    lib.qr_host_unregister.argtypes = [vp, vp]
    lib.qr_host_unregister.restype = ci
    lib.qr_frame_notify_slot.argtypes = [vp, vp, ci, ctypes.POINTER(vp)]
    lib.qr_frame_notify_slot.restype = ci
    lib.qr_render_rows_notify.argtypes = [vp, vp, ci, ci, ci, vp]
    lib.qr_render_rows_notify.restype = ci
    lib.qr_wait_notify.argtypes = [vp, vp, ctypes.c_uint32]
    lib.qr_wait_notify.restype = ci
    lib.qr_frame_ipc_export.argtypes = [vp, vp]
    lib.qr_frame_ipc_export.restype = ci
    lib.qr_frame_ipc_open.argtypes = [vp, vp, ctypes.POINTER(vp)]
    lib.qr_frame_ipc_open.restype = ci
    lib.qr_frame_ipc_close.argtypes = [vp, vp]
    lib.qr_frame_ipc_close.restype = ci
    lib.qr_sync.argtypes = [vp]
    lib.qr_pipeline.argtypes = [vp, ci]
    lib.qr_pipeline.restype = ci
    lib.qr_render_begin.argtypes = [vp, ctypes.POINTER(ci)]
    lib.qr_render_begin.restype = ci
    lib.qr_render_fetch.argtypes = [vp, ci, vp, ci]
    lib.qr_render_fetch.restype = ci
    lib.qr_render_end.argtypes = [vp, ci, vp, ci]
    lib.qr_render_end.restype = ci
    lib.qr_sync.restype = ci
    lib.qr_frame_device.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(ci)]
    lib.qr_frame_device.restype = ci
    lib.qr_dump_hits.argtypes = [vp, vp]
    lib.qr_dump_hits.restype = ci
    lib.qr_ray_counts.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64)]
    lib.qr_ray_counts.restype = ci
    lib.qr_last_render_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    lib.qr_last_render_ms.restype = ci
    lib.qr_stream.argtypes = [vp, ci]
    lib.qr_stream.restype = vp
    lib.qr_launch_count.argtypes = [vp]
    lib.qr_launch_count.restype = ctypes.c_uint64
    lib.qr_kernel_query.argtypes = [vp, ctypes.POINTER(KernelInfo)]
    lib.qr_kernel_query.restype = ci
    lib.qr_fp32_peak.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    lib.qr_fp32_peak.restype = ci
    _lib = lib
    return lib


def _ptr(a):
    """void* of a numpy array / bytes-like / int device pointer."""
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(ctypes.c_void_p)
    return ctypes.cast(ctypes.c_char_p(bytes(a)), ctypes.c_void_p)


class Context(object):
    """One qr_ctx: the GPUs that render a frame together."""

    def __init__(self, devices=None):
        self.lib = load_library()
        self.h = ctypes.c_void_p()
        if devices is None:
            rc = self.lib.qr_init(None, 0, ctypes.byref(self.h))
        else:
            arr = (ctypes.c_int * len(devices))(*devices)
            rc = self.lib.qr_init(arr, len(devices), ctypes.byref(self.h))
        if rc != QR_OK:
            text = self.lib.qr_last_error(None).decode()
            self.h = ctypes.c_void_p()
            raise QuadRayError(rc, text)
        self.header = None

    def _check(self, rc):
        if rc != QR_OK:
            raise QuadRayError(rc, self.lib.qr_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.qr_shutdown(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, blob):
        """blob: bytes or uint8 ndarray (host memory)."""
        if isinstance(blob, np.ndarray):
            b = np.ascontiguousarray(blob, dtype=np.uint8)
            self._check(self.lib.qr_scene_upload(self.h, _ptr(b), b.size))
            raw = b[:256].tobytes()
        else:
            raw = bytes(blob)
            self._check(self.lib.qr_scene_upload(self.h, raw, len(raw)))
        hdr = np.frombuffer(raw[:256], dtype=np.int32)
        self.header = {"x_res": int(hdr[4]), "y_res": int(hdr[5]), "x_row": int(hdr[6]),
                       "fsaa": int(hdr[7]), "depth": int(hdr[8]), "tile_w": int(hdr[9]),
                       "tile_h": int(hdr[10]), "tls_row": int(hdr[11]), "tls_col": int(hdr[12])}

    def render(self, frame=None, stride=None):
        """Render into a host uint32 array (rows "stride" pixels apart) or,
        with frame=None, asynchronously into GPU 0's framebuffer."""
        if frame is None:
            self._check(self.lib.qr_render(self.h, None, 0))
            return None
        if stride is None:
            stride = frame.strides[0] // 4 if frame.ndim == 2 else self.header["x_res"]
        self._check(self.lib.qr_render(self.h, _ptr(frame), int(stride)))
        return frame

    def pipeline(self, on=True):
        """Pipelined frames: upload alternates between two scene slots,
        render_begin queues a frame, render_end collects it."""
        self._check(self.lib.qr_pipeline(self.h, 1 if on else 0))

    def render_begin(self):
        t = ctypes.c_int(-1)
        self._check(self.lib.qr_render_begin(self.h, ctypes.byref(t)))
        return (t.value, dict(self.header))

    def render_fetch(self, ticket, frame):
        """Start delivering the frame of "ticket" into "frame"; returns at once."""
        t, _ = ticket
        self._check(self.lib.qr_render_fetch(self.h, t, _ptr(frame), frame.strides[0] // 4))

    def render_end(self, ticket, frame=None, fetched=False):
        """Wait for the frame of "ticket"; fetched=True completes a render_fetch."""
        t, h = ticket
        if fetched:
            self._check(self.lib.qr_render_end(self.h, t, None, 0))
            return frame
        if frame is None:
            frame = np.zeros((h["y_res"], h["x_res"]), dtype=np.uint32)
        self._check(self.lib.qr_render_end(self.h, t, _ptr(frame), frame.strides[0] // 4))
        return frame

    def render_frame(self):
        h = self.header
        out = np.zeros((h["y_res"], h["x_res"]), dtype=np.uint32)
        return self.render(out, h["x_res"])

    def render_device(self, dev_ptr, stride, y0, y1):
        self._check(self.lib.qr_render_device(self.h, ctypes.c_void_p(dev_ptr), int(stride), int(y0), int(y1)))

    def render_rows(self, dev_ptr, stride, tile_row0, tile_row_step):
        """Tile rows tile_row0, tile_row0 + step, ... into a device buffer of
        full-frame geometry (possibly another GPU's, see frame_ipc_open)."""
        self._check(self.lib.qr_render_rows(self.h, ctypes.c_void_p(dev_ptr), int(stride),
                                            int(tile_row0), int(tile_row_step)))

    def pt_reset(self, n_slots, pseed=None):
        """Path tracer: new seed plane (the reference's own when pseed is None),
        colour planes and frame count at zero (rt_Scene::set_pton(1))."""
        if pseed is not None:
            pseed = np.ascontiguousarray(pseed, dtype=np.uint32)
            assert pseed.size == int(n_slots)
        self._check(self.lib.qr_pt_reset(self.h, pseed.ctypes.data if pseed is not None else None, int(n_slots)))

    def pt_frames(self):
        return int(self.lib.qr_pt_frames(self.h))

    def pt_fetch(self, n_slots):
        """(seeds, red, green, blue) planes of GPU 0."""
        sd = np.empty(int(n_slots), dtype=np.uint32)
        pl = [np.empty(int(n_slots), dtype=np.float32) for _ in range(3)]
        self._check(self.lib.qr_pt_fetch(self.h, sd.ctypes.data, pl[0].ctypes.data, pl[1].ctypes.data,
                                         pl[2].ctypes.data, int(n_slots)))
        return sd, pl[0], pl[1], pl[2]

    def check_counters(self):
        """Violation counters of the checked build (raises in the normal build)."""
        c = (ctypes.c_uint32 * 8)()
        self._check(self.lib.qr_check_counters(self.h, c))
        return [int(x) for x in c]

    def host_register(self, host_ptr, nbytes):
        """Page-lock a host buffer (e.g. a shared-memory frame) for this context's
        GPU; returns the device pointer the kernels store through."""
        out = ctypes.c_void_p()
        self._check(self.lib.qr_host_register(self.h, ctypes.c_void_p(host_ptr), int(nbytes), ctypes.byref(out)))
        return out.value

    def host_unregister(self, host_ptr):
        self._check(self.lib.qr_host_unregister(self.h, ctypes.c_void_p(host_ptr)))

    def frame_notify_slot(self, frame_dev_ptr, index):
        """Device address of completion counter "index" behind the library's
        framebuffer as mapped at frame_dev_ptr (own or opened through IPC)."""
        out = ctypes.c_void_p()
        self._check(self.lib.qr_frame_notify_slot(self.h, ctypes.c_void_p(frame_dev_ptr), int(index),
                                                  ctypes.byref(out)))
        return out.value

    def render_rows_notify(self, dev_ptr, stride, tile_row0, tile_row_step, notify_ptr):
        """render_rows + the kernel adds 1 to *notify_ptr when its pixels are visible."""
        self._check(self.lib.qr_render_rows_notify(self.h, ctypes.c_void_p(dev_ptr), int(stride),
                                                   int(tile_row0), int(tile_row_step),
                                                   ctypes.c_void_p(notify_ptr)))

    def wait_notify(self, notify_ptr, target):
        """The context's stream waits until the counter has reached target."""
        self._check(self.lib.qr_wait_notify(self.h, ctypes.c_void_p(notify_ptr), ctypes.c_uint32(target & 0xFFFFFFFF)))

    def frame_ipc_export(self):
        """64-byte CUDA IPC handle of GPU 0's framebuffer (bytes)."""
        buf = ctypes.create_string_buffer(64)
        self._check(self.lib.qr_frame_ipc_export(self.h, buf))
        return buf.raw

    def frame_ipc_open(self, handle):
        p = ctypes.c_void_p()
        self._check(self.lib.qr_frame_ipc_open(self.h, ctypes.create_string_buffer(bytes(handle), 64),
                                               ctypes.byref(p)))
        return p.value

    def frame_ipc_close(self, dev_ptr):
        self._check(self.lib.qr_frame_ipc_close(self.h, ctypes.c_void_p(dev_ptr)))

    def sync(self):
        self._check(self.lib.qr_sync(self.h))

    def frame_device(self):
        p, s = ctypes.c_void_p(), ctypes.c_int()
        self._check(self.lib.qr_frame_device(self.h, ctypes.byref(p), ctypes.byref(s)))
        return p.value, s.value

    def dump_hits(self):
        h = self.header
        out = np.zeros((h["y_res"], h["x_res"], 1 << h["fsaa"]), dtype=np.float32)
        self._check(self.lib.qr_dump_hits(self.h, _ptr(out)))
        return out

    def ray_counts(self):
        c = (ctypes.c_uint64 * 4)()
        self._check(self.lib.qr_ray_counts(self.h, c))
        return {"primary": int(c[0]), "shadow": int(c[1]), "reflect": int(c[2]), "refract": int(c[3])}

    def last_render_ms(self):
        ms = ctypes.c_float()
        self._check(self.lib.qr_last_render_ms(self.h, ctypes.byref(ms)))
        return float(ms.value)

    def stream(self, index=0):
        return self.lib.qr_stream(self.h, index)

    def launch_count(self):
        return int(self.lib.qr_launch_count(self.h))

    def fp32_peak(self):
        """Measured non-FMA FP32 rate of GPU 0, 1e12 ops/s."""
        t = ctypes.c_double()
        self._check(self.lib.qr_fp32_peak(self.h, ctypes.byref(t)))
        return float(t.value)

    def kernel_info(self):
        k = KernelInfo()
        self._check(self.lib.qr_kernel_query(self.h, ctypes.byref(k)))
        return {n: getattr(k, n) for n, _ in KernelInfo._fields_}


def rank_tile_rows(y_res, tile_h, rank, world):
    """Tile rows of rank "rank" out of "world" (SURVEY.md 8e): rows rank,
    rank + world, ... -- the round-robin deal qr_render uses across the GPUs
    of one context and qr_render_rows(rank, world) renders."""
    tls_col = (y_res + tile_h - 1) // tile_h
    return list(range(rank, tls_col, world))


def tile_row_span(y_res, tile_h, tile_row):
    """Frame rows [y0, y1) of one tile row."""
    y0 = tile_row * tile_h
    return y0, min(y0 + tile_h, y_res)


class SharedFrameRing(object):
    """Two host frames shared by the ranks of one box (one process per GPU), in
    ONE POSIX shared-memory segment -- the role RooT's XShm image plays for the
    reference's worker threads.  Every rank page-locks its mapping
    (Context.host_register) and renders its tile rows straight into frame
    k & 1; the segment also carries the hand-shake:

        flags[rank]  last frame whose rows this rank has completed
        flags[world] last frame rank 0 has consumed (its buffer is free again)

    Frames are numbered from 1.  Protocol per frame k on every rank:
        wait_free(k); render rows into frame(k); ... ; mark_done(k)
    and on rank 0:  wait_complete(k); read frame(k); release(k).
    Host-side logic only (no CUDA here): tests/test_sharding_gloo.py runs it on
    the CPU with the oracle standing in for the device."""

    def __init__(self, path, y_res, x_row, world, rank, create):
        import mmap
        self.path, self.world, self.rank = path, int(world), int(rank)
        self.fbytes = int(y_res) * int(x_row) * 4
        self.nbytes = 2 * self.fbytes + 4096
        flags = os.O_RDWR | ((os.O_CREAT | os.O_TRUNC) if create else 0)
        fd = os.open(path, flags, 0o600)
        try:
            if create:
                os.ftruncate(fd, self.nbytes)
            self.map = mmap.mmap(fd, self.nbytes)
        finally:
            os.close(fd)
        self.bytes = np.frombuffer(self.map, dtype=np.uint8)
        self.frames = [self.bytes[k * self.fbytes:(k + 1) * self.fbytes].view(np.uint32).reshape(int(y_res), int(x_row))
                       for k in range(2)]
        self.flags = self.bytes[2 * self.fbytes:2 * self.fbytes + 8 * (self.world + 1)].view(np.int64)
        if create:
            self.flags[:] = 0

    @property
    def host_ptr(self):
        return self.bytes.ctypes.data

    def frame(self, k):
        return self.frames[k & 1]

    def frame_offset(self, k):
        return (k & 1) * self.fbytes

    def wait_free(self, k):
        """Frame k's buffer is free once frame k - 2 has been consumed."""
        while int(self.flags[self.world]) < k - 2:
            pass

    def mark_done(self, k):
        self.flags[self.rank] = k

    def wait_complete(self, k):
        while int(self.flags[:self.world].min()) < k:
            pass

    def release(self, k):
        self.flags[self.world] = k

    def close(self, unlink=False):
        self.frames, self.flags, self.bytes = None, None, None
        try:
            self.map.close()
        except BufferError:
            pass
        if unlink:
            try:
                os.unlink(self.path)
            except OSError:
                pass
