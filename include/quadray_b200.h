/*
 * quadray_b200.h -- C ABI of libquadray_b200.so, the B200 (sm_100a) backend
 * for QuadRay's core/tracer render0 path.
 *
 * The library is the drop-in boundary: plain C, plain pointers and sizes, no
 * exceptions, no torch types.  It is what the replacement tracer translation
 * unit (quadray-engine_b200/host/tracer_b200.cpp) calls from inside
 *
 *     rt_void rt_Platform::render0(rt_SIMD_INFOX *s_inf)   core/engine/engine.h:127
 *                                                          core/tracer/tracer.cpp:5992-6104
 *
 * i.e. it replaces the 21 per-SIMD-width instantiations of render0
 * (core/tracer/tracer.cpp:1081-5405, tracer_*.cpp) and their run-time
 * dispatch.  The engine-side structures (rt_SIMD_SURFACE, rt_ELEM lists, ...)
 * are flattened by the caller into the index-based blob of qr_scene_blob.h.
 *
 * Every function returns 0 on success or a negative QR_E_* code;
 * qr_last_error() gives the text.  Nothing here falls back to the CPU: if no
 * CUDA device is usable qr_init fails.
 */
#ifndef QUADRAY_B200_H
#define QUADRAY_B200_H

#include <stddef.h>
#include <stdint.h>

#include "qr_scene_blob.h"

#ifdef __cplusplus
extern "C" {
#endif

#define QR_OK            0
#define QR_E_ARG        (-1)    /* bad argument */
#define QR_E_BLOB       (-2)    /* malformed scene blob */
#define QR_E_CUDA       (-3)    /* CUDA runtime error */
#define QR_E_NODEV      (-4)    /* no usable CUDA device */
#define QR_E_STATE      (-5)    /* call order (e.g. render before upload) */

typedef struct qr_ctx qr_ctx;

/*
 * Create a context on "ndev" CUDA devices (devices == NULL: device 0..ndev-1;
 * ndev == 0: one device, the current one).  One stream, one pinned staging
 * area and one device arena per GPU: the per-GPU streams take the place of the
 * reference's worker threads (rt_FUNC_INIT, core/engine/engine.h:71;
 * root/RooT_linux.cpp:632-727).
 */
int qr_init(const int *devices, int ndev, qr_ctx **out);

/* rt_FUNC_TERM counterpart (engine.h:72): drains the streams, frees everything. */
void qr_shutdown(qr_ctx *ctx);

/* Text of the last error of this context (ctx == NULL: of the last failed
 * qr_init on this thread).  Never NULL. */
const char *qr_last_error(const qr_ctx *ctx);

/*
 * Hand over the scene for the next frame(s): the flattened equivalent of the
 * pointer graph reachable from rt_SIMD_INFOX (core/tracer/tracer.h:150-216,
 * core/engine/engine.cpp:3600-3627).  The blob is copied during the call
 * (pinned staging, async H2D to every GPU of the context), so the caller may
 * release its per-frame pools right after (engine.cpp:3317-3323).
 */
int qr_scene_upload(qr_ctx *ctx, const void *blob, size_t bytes);

/*
 * Render the uploaded scene: what pfm->render0(s_inf) does for ALL thread
 * indices at once (core/engine/engine.cpp:3284-3298, 3627).  Tile rows are
 * dealt round-robin to the GPUs of the context; peers store their pixels
 * straight into GPU 0's framebuffer over NVLink (P2P stores from the kernel,
 * peer copies when P2P is unavailable).
 *   frame  != NULL: host framebuffer of y_res rows, "stride" pixels apart
 *                   (0x00RRGGBB, rt_Scene::get_frame(), engine.cpp:3774-3777);
 *                   the call returns when it holds the finished frame.  On
 *                   one GPU a page-locked frame (cudaMallocHost /
 *                   cudaHostRegister, or QR_B200_PIN_FRAME=1) is written by the
 *                   kernel itself over PCIe while it renders -- no D2H pass;
 *                   any other frame goes through pinned staging in two chunks
 *                   whose copy overlaps the rendering of the next.
 *   frame  == NULL: render only (asynchronous; see qr_sync / qr_frame_device).
 * Alignment: frame buffers (host or device, here and in qr_render_device /
 * qr_render_rows / qr_render_fetch) need 4-byte alignment only.  The kernel
 * issues 128-bit stores where the ADDRESS of a group of four pixels is
 * 16-byte aligned and 32-bit stores elsewhere, so a frame at base + 4 or an
 * odd stride renders correctly (a little slower).
 */
int qr_render(qr_ctx *ctx, uint32_t *frame, int stride);

/*
 * Render rows [y0, y1) (y0 a multiple of tile_h) into a caller-owned DEVICE
 * buffer of full-frame geometry that GPU 0 of the context can address,
 * asynchronously on the context's stream.  Used when the consumer is on the
 * GPU.
 */
int qr_render_device(qr_ctx *ctx, uint32_t *frame_dev, int stride, int y0, int y1);

/*
 * Render the tile rows tile_row0, tile_row0 + tile_row_step, ... (tile row r
 * covers frame rows [r * tile_h, (r + 1) * tile_h)) into a DEVICE buffer of
 * full-frame geometry, asynchronously on the context's stream.  This is the
 * unit of multi-GPU sharding: rank r of N calls it with (r, N) -- the
 * interleave the reference's worker threads use for scanlines (index,
 * index + thnum, ...; core/tracer/tracer.cpp:1142-1151, 5383-5394), at tile
 * row granularity.  frame_dev may be a peer GPU's buffer (qr_frame_ipc_open or
 * a peer-enabled pointer): the pixels then travel over NVLink as the kernel
 * stores them and no gather pass is needed.
 */
int qr_render_rows(qr_ctx *ctx, uint32_t *frame_dev, int stride, int tile_row0, int tile_row_step);

/*
 * Completion signal of sharded frames without a collective.  The framebuffer
 * of the library (qr_frame_device, and what qr_frame_ipc_open maps on the other
 * ranks) is followed by 16 counters; qr_frame_notify_slot returns the device
 * address of one of them for a given mapping of that buffer.
 *   qr_render_rows_notify  as qr_render_rows; in addition the kernel adds 1
 *                          (system scope) to *notify_dev when every pixel of
 *                          this launch is visible to the owner of the frame --
 *                          from a peer GPU the counter is reached over NVLink,
 *                          like the pixels
 *   qr_wait_notify         the context's stream waits until the counter has
 *                          reached "target" (counters only ever grow; the
 *                          comparison is wrap-around safe), e.g. target =
 *                          frames so far * ranks on the rank that owns the frame
 * This replaces the reference's worker join (rt_FUNC_RENDER returning on all
 * threads, core/engine/engine.h:71-74) across GPUs.
 */
int qr_frame_notify_slot(qr_ctx *ctx, uint32_t *frame_dev, int index, uint32_t **slot_dev);

/*
 * A HOST framebuffer shared by several processes (one rank per GPU, e.g. a
 * POSIX shared-memory segment, the role RooT's XShm image plays): every rank
 * page-locks its mapping and renders its tile rows straight into it with
 * qr_render_rows(dev_ptr, ...) -- the pixels leave every GPU over its own PCIe
 * link while the kernel runs, no gather to one GPU and no D2H pass.
 */
int qr_host_register(qr_ctx *ctx, void *host, size_t bytes, uint32_t **dev_ptr);
int qr_host_unregister(qr_ctx *ctx, void *host);
int qr_render_rows_notify(qr_ctx *ctx, uint32_t *frame_dev, int stride, int tile_row0, int tile_row_step,
                          uint32_t *notify_dev);
int qr_wait_notify(qr_ctx *ctx, uint32_t *notify_dev, uint32_t target);

/*
 * Pipelined frames: the engine's update phases of frame N + 1 (rt_Scene::render
 * phases 0.5-3 and the tiling merge, core/engine/engine.cpp:2976-3276 -- all
 * host work) run while the GPUs render frame N, which the reference's
 * update-barrier-render loop (root/RooT_linux.cpp:775-792) cannot do.
 *
 *   qr_pipeline(ctx, 1)        scenes alternate between two device slots, so
 *                              qr_scene_upload does not wait for the frame in
 *                              flight (only for the one begun two uploads ago)
 *   qr_render_begin(ctx, &t)   queue the frame of the scene uploaded last into
 *                              a page-locked frame owned by the library;
 *                              returns at once with a ticket (0 or 1)
 *   qr_render_fetch(ctx, t, frame, stride)
 *                              optional: start moving that frame to "frame" as
 *                              soon as it is rendered and return at once -- by
 *                              DMA straight into a page-locked frame (an
 *                              application frame is page-locked on first sight
 *                              unless QR_B200_PIN_FRAME=0), else into pinned
 *                              staging; the caller overlaps its host work
 *   qr_render_end(ctx, t, frame, stride)
 *                              wait until "frame" holds that frame (fetching
 *                              it now unless a fetch is under way; geometry of
 *                              the scene it was begun with; frame == NULL:
 *                              complete the fetch under way, or wait and drop)
 * At most two frames are in flight.  qr_pipeline(ctx, 0) drains and returns to
 * the synchronous protocol (upload, qr_render).
 */
int qr_pipeline(qr_ctx *ctx, int on);
int qr_render_begin(qr_ctx *ctx, int *ticket);
int qr_render_fetch(qr_ctx *ctx, int ticket, uint32_t *frame, int stride);
int qr_render_end(qr_ctx *ctx, int ticket, uint32_t *frame, int stride);

/* Wait for all queued work of the context. */
int qr_sync(qr_ctx *ctx);

/* GPU 0's framebuffer for the uploaded scene (x_row-strided; allocated on
 * first use, holds the last qr_render) and its stride. */
int qr_frame_device(qr_ctx *ctx, const uint32_t **frame_dev, int *stride);

/*
 * One process per GPU: the rank that owns the framebuffer exports it
 * (cudaIpcMemHandle_t, 64 bytes), the other ranks open it and pass the
 * pointer to qr_render_rows, so every rank stores its tile rows straight into
 * the owner's framebuffer over NVLink.  The owner must not re-upload a scene
 * of another frame geometry while the handle is open elsewhere.
 */
int qr_frame_ipc_export(qr_ctx *ctx, void *handle64);
int qr_frame_ipc_open(qr_ctx *ctx, const void *handle64, uint32_t **frame_dev);
int qr_frame_ipc_close(qr_ctx *ctx, uint32_t *frame_dev);

/*
 * "Dump mode": per primary sample hit distance (ctx_T_BUF at XX_end,
 * core/tracer/tracer.cpp:5161), y_res * x_res * (1 << fsaa) floats, sample-
 * major within a pixel; +inf where nothing was hit.  Renders one frame.
 */
int qr_dump_hits(qr_ctx *ctx, float *t_out);

/*
 * Rays cast by the renders since the last call (then reset):
 * counts[0] primary samples, [1] shadow, [2] reflection, [3] refraction rays,
 * counted where the reference pushes a context (tracer.cpp:2801-2831,
 * 3486-3532, 3831-3866), for the surface that ends up visible.
 */
int qr_ray_counts(qr_ctx *ctx, uint64_t counts[4]);

/* Device time (ms) of the kernels of the last qr_render / qr_render_device,
 * from CUDA events on the launching stream; max over the GPUs. */
int qr_last_render_ms(qr_ctx *ctx, float *ms);

/* cudaStream_t of GPU "index" of the context (for event timing by callers). */
void *qr_stream(qr_ctx *ctx, int index);

/* Number of kernels this library has launched since qr_init. */
uint64_t qr_launch_count(const qr_ctx *ctx);

/* Static facts about the kernel build: registers, resident CTAs per SM, ... */
typedef struct qr_kernel_info
{
    int sm_count;
    int threads_per_cta;
    int ctas_per_sm;
    int regs_per_thread;
    int local_bytes_per_thread;
    int smem_static_bytes;
    int smem_dynamic_bytes;   /* for the currently uploaded scene, 0 if none */
    int scene_in_smem;        /* 1 when surfaces/materials/lights are staged */
    int device_tiling;        /* 1 when the tile lists of the current scene were built on the device
                                 (the engine ran with RT_OPTS_TILING off and the blob carries bounds) */
} qr_kernel_info;

int qr_kernel_query(qr_ctx *ctx, qr_kernel_info *info);

/*
 * Path tracer (rt_Scene::set_pton, render0 with pt_on:
 * core/tracer/tracer.cpp:1112-1136, 1218-1285, 2339-2701, 3428-3466,
 * 5176-5219).  A blob whose header carries QR_BLOB_PT (the flattener sets it
 * from rt_SIMD_INFOX::pt_on) is rendered by the packet tracer of csrc/qr_pt.cuh
 * -- one warp = one 32-lane packet of the reference's 512x2v2 target, frames
 * bit-identical to that target's -- through the same render calls (not the
 * pipelined ones, not qr_dump_hits); x_row must be a multiple of 32 >> fsaa.
 * Every frame adds one sample per pixel sample to a running mean.
 *
 * qr_pt_reset   = rt_Scene::reset_pseed + reset_color + inf_PTS_C = 0
 *                 (engine.cpp:3670-3700): "pseed" is the scene's seed plane of
 *                 n_slots = 4 * x_row * y_res words (NULL: filled here as
 *                 reset_pseed fills it), the colour planes start at zero.
 *                 The state lives on the device(s) from then on.
 * qr_pt_frames  frames accumulated since the reset.
 * qr_pt_fetch   copies GPU 0's seed / colour planes out (any pointer may be
 *                 NULL); for tests and for a caller that wants rt_Scene's host
 *                 planes current.
 */
int qr_pt_reset(qr_ctx *ctx, const uint32_t *pseed, size_t n_slots);
int qr_pt_frames(qr_ctx *ctx);
int qr_pt_fetch(qr_ctx *ctx, uint32_t *pseed, float *ptr_r, float *ptr_g, float *ptr_b, size_t n_slots);

/*
 * The checked build of the library (make checked: -DQR_CHECKED,
 * lib/libquadray_b200_checked.so) counts out-of-range element cursors, surface
 * offsets, tile indices, pixel stores, stack levels and scratch reads that
 * were not written first, instead of faulting; this returns the counters since
 * the last call and clears them.  QR_E_STATE in the normal build.
 */
int qr_check_counters(qr_ctx *ctx, uint32_t counters[8]);

/*
 * Measured ceiling of the pipe that bounds render0 on this GPU: separately
 * rounded FP32 multiplies and adds (no FMA contraction, as bit parity with the
 * reference demands), in 1e12 operations per second, best of a few launches
 * of a register-resident FMUL/FADD loop on GPU 0 at the clocks of the moment.
 * It is the denominator of the roofline fraction bench.py reports.
 */
int qr_fp32_peak(qr_ctx *ctx, double *tera_ops);

#ifdef __cplusplus
}
#endif

#endif /* QUADRAY_B200_H */
