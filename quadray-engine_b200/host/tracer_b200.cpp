/*
 * tracer_b200.cpp -- replacement for the reference's core/tracer translation
 * units (tracer.cpp + the 21 tracer_*.cpp re-includes).
 *
 * It defines the three "backend global entry points" that rt_Platform
 * declares and the reference implements in tracer.cpp (core/engine/engine.h:
 * 123-127, core/tracer/tracer.h:109-112):
 *
 *     rt_si32 rt_Platform::switch0(rt_SIMD_INFOX*, rt_si32 simd)   tracer.cpp:5827-5876
 *     rt_void rt_Platform::update0(rt_SIMD_SURFACE*)               tracer.cpp:5774-5808
 *     rt_void rt_Platform::render0(rt_SIMD_INFOX*)                 tracer.cpp:5992-6104
 *
 * so the engine (core/engine, core/system), the scene format (format.h) and
 * the applications (RooT, core_test) link unchanged.  render0 flattens the
 * engine's per-frame pointer graph into the scene blob (qr_flatten.cpp) and
 * hands it to libquadray_b200.so through the C ABI of include/quadray_b200.h;
 * the frame is back in s_inf->frame when it returns, as the callers expect
 * (engine.cpp:3774-3777, test/core_test.cpp:984, root/RooT.h:597-631).
 *
 * There is no CPU path in this file: without a usable B200 render0 throws
 * rt_Exception.
 */

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <sys/time.h>

#include "tracer.h"
#include "format.h"
#include <mutex>
#include "engine.h"

#include "qr_flatten.h"
#include "quadray_b200.h"

/*
 * Per-surface solver / normal / clipper dispatch tags and the conic flag,
 * same rules as tracer.cpp:5774-5808:
 *   srf_t[0] solver   1 plane, 2 quadric, 3 two-plane (hypercylinder, w == 0)
 *   srf_t[1] material 1 plane, 2 with linear terms (parabolic family), 3 pure quadratic
 *   srf_t[2] clipper  same classes as srf_t[1]
 *   msc_p[1] conic    1 cone / hyperboloid with w == 0, 2 hypercylinder with w == 0
 */
rt_void rt_Platform::update0(rt_SIMD_SURFACE *s_srf)
{
    rt_ui32 tag = (rt_ui32)(rt_word)s_srf->srf_t[3];

    if (tag >= RT_TAG_SURFACE_MAX)
    {
        return;
    }

    rt_bool quad = tag > RT_TAG_PLANE;
    rt_bool flat = s_srf->sci_w[0] == 0.0f;
    rt_bool para = tag == RT_TAG_PARABOLOID || tag == RT_TAG_PARACYLINDER
                || tag == RT_TAG_HYPERPARABOLOID;

    s_srf->srf_t[0] = !quad ? 1 : (tag == RT_TAG_HYPERCYLINDER && flat) ? 3 : 2;
    s_srf->srf_t[1] = !quad ? 1 : para ? 2 : 3;
    s_srf->srf_t[2] = !quad ? 1 : para ? 2 : 3;
    s_srf->msc_p[1] = (tag == RT_TAG_CONE || (tag == RT_TAG_HYPERBOLOID && flat))
                    ? (rt_pntr)1
                    : (tag == RT_TAG_HYPERCYLINDER && flat) ? (rt_pntr)2
                    : (rt_pntr)0;
}

/*
 * The B200 backend has one target: a warp of 32 fp32 lanes.  In the
 * reference's vocabulary (n_simd x k_size, engine.cpp:628-639) that is the
 * paired 512-bit target 512x2, width 32 = RT_SIMD_WIDTH of this build, so
 * get_fsaa_max(), x_row and tile_w rounding (engine.cpp:570, 668-681,
 * 2833-2839) stay coherent.  A request for any other target is answered with
 * this one; the applications then report "not supported" themselves
 * (test/core_test.cpp:907-918).
 */
rt_si32 rt_Platform::switch0(rt_SIMD_INFOX *s_inf, rt_si32 simd)
{
    s_mask = 0x02000000;
    s_mode = 0x02000000;
    return simd_init(4, 2, 2);
}

/* process-wide GPU context, created by the first render0 */
static qr_ctx      *g_ctx = RT_NULL;
static qr_Flattener g_flat;
static char         g_err[600];
static bool         g_pipelined = false;
static double       g_tb = 0.0;         /* QR_B200_TIMING: frame queued at */

/* pipelined mode: the frame in flight and the buffer it belongs to */
static struct
{
    int      ticket;
    bool     fetching;
    rt_pntr  frame;
    rt_si32  row, w, h;
}
g_prev = { -1, false, RT_NULL, 0, 0, 0 };

static double qr_now_ms()
{
    timeval tm;
    gettimeofday(&tm, NULL);
    return tm.tv_sec * 1000.0 + tm.tv_usec / 1000.0;
}

static rt_void qr_throw(rt_pstr what, const qr_ctx *ctx)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", what, qr_last_error(ctx));
    throw rt_Exception(g_err);
}

static rt_void qr_atexit()
{
    if (g_ctx != RT_NULL)
    {
        qr_shutdown(g_ctx);
        g_ctx = RT_NULL;
    }
}

/*
 * QR_B200_DEVICES="0,1,2,3" selects the GPUs that share a frame by tile-row
 * bands (default: the current device only).
 */
static rt_void qr_context()
{
    if (g_ctx != RT_NULL)
    {
        return;
    }

    int devs[16], ndev = 0;
    const char *env = getenv("QR_B200_DEVICES");
    if (env != RT_NULL)
    {
        const char *p = env;
        while (*p != 0 && ndev < 16)
        {
            char *end = RT_NULL;
            long d = strtol(p, &end, 10);
            if (end == p)
            {
                break;
            }
            devs[ndev++] = (int)d;
            p = *end == ',' ? end + 1 : end;
        }
    }

    if (qr_init(ndev > 0 ? devs : RT_NULL, ndev, &g_ctx) != QR_OK)
    {
        g_ctx = RT_NULL;
        qr_throw("B200 backend init failed", RT_NULL);
    }
    atexit(qr_atexit);

    const char *pipe = getenv("QR_B200_PIPELINE");
    if (pipe != RT_NULL && pipe[0] == '1')
    {
        if (qr_pipeline(g_ctx, 1) != QR_OK)
        {
            qr_throw("B200 pipelined mode failed", g_ctx);
        }
        g_pipelined = true;
    }
}

/*
 * Called once per worker index by render_slice (engine.cpp:3458-3628).  The
 * whole frame is one GPU submission, so index 0 does the work and the other
 * indices of the same frame return at once (they would otherwise render the
 * interleaved rows index, index + thnum, ... tracer.cpp:1142-1151).
 */
rt_void rt_Platform::render0(rt_SIMD_INFOX *s_inf)
{
    if (s_inf->index != 0)
    {
        return;
    }

    /* The reference's render0 is re-entrant per rt_SIMD_INFOX; here one GPU
     * context, one flattener and (pipelined mode) one frame in flight serve the
     * process.  Scenes of several rt_Platform instances that render
     * concurrently take turns. */
    static std::mutex g_render_mtx;
    std::lock_guard<std::mutex> render_lock(g_render_mtx);

    qr_context();

    /* QR_B200_TIMING=1: wall time of the three steps, every 64th frame, on
     * stderr in the style of the reference's "Time" lines */
    static const bool timing = getenv("QR_B200_TIMING") != RT_NULL;
    static unsigned frame_no = 0;
    const double t0 = timing ? qr_now_ms() : 0.0;

    /* pipelined mode: the previous frame starts its way into the caller's
     * buffer now, so the transfer runs beside the flattening and packing */
    const bool same = g_prev.frame == s_inf->frame && g_prev.row == s_inf->frm_row
                   && g_prev.w == s_inf->frm_w && g_prev.h == s_inf->frm_h;
    if (g_pipelined && g_prev.ticket >= 0 && same && !g_prev.fetching)
    {
        /* (a frame stays "fetching" if a later step of that call threw: the
         * next call completes it instead of starting it again) */
        if (qr_render_fetch(g_ctx, g_prev.ticket, (uint32_t *)s_inf->frame, (int)s_inf->frm_row) != QR_OK)
        {
            qr_throw("B200 frame fetch failed", g_ctx);
        }
        g_prev.fetching = true;
    }

    size_t bytes = 0;
    const uint8_t *blob = g_flat.build(s_inf, &bytes);
    const double t1 = timing ? qr_now_ms() : 0.0;

    {
        /* QR_B200_DUMP_BLOB=path: the scene blob of the first frame goes to a
         * file (tools/config5.py times big scenes through the C ABI with it) */
        static const char *dump = getenv("QR_B200_DUMP_BLOB");
        static bool dumped = false;
        if (dump != RT_NULL && !dumped)
        {
            dumped = true;
            FILE *f = fopen(dump, "wb");
            if (f != RT_NULL)
            {
                fwrite(blob, 1, bytes, f);
                fclose(f);
            }
        }
    }

    /*
     * Path tracer (rt_Scene::set_pton): the seed and colour planes live on the
     * device from the first frame on.  rt_Scene resets its host planes when the
     * mode is switched on (engine.cpp:3731-3741); that shows here as the first
     * pt_on frame, other planes, or the marker this backend leaves in the host's
     * red plane (nothing else reads it while the backend renders) wiped by
     * reset_color.  (Switched off and on again with no frame in between, the
     * reference keeps counting samples; here the count restarts with the planes.)
     */
    static struct { bool on; const void *pseed; size_t slots; } g_pt = { false, RT_NULL, 0 };
    static const rt_ui32 pt_mark = 0x7FC0B200u;
    if (s_inf->pt_on != 0)
    {
        if (g_pipelined)
        {
            throw rt_Exception("B200 backend: no path tracer in pipelined mode (QR_B200_PIPELINE)");
        }
        const size_t slots = (size_t)4 * (size_t)s_inf->frm_row * (size_t)s_inf->frm_h;
        const bool fresh = !g_pt.on || g_pt.pseed != s_inf->pseed || g_pt.slots != slots
                        || memcmp(s_inf->ptr_r, &pt_mark, sizeof(pt_mark)) != 0;
        if (fresh && qr_pt_reset(g_ctx, (const uint32_t *)s_inf->pseed, slots) != QR_OK)
        {
            qr_throw("B200 path-tracer reset failed", g_ctx);
        }
        g_pt.on = true; g_pt.pseed = s_inf->pseed; g_pt.slots = slots;
    }
    else
    {
        g_pt.on = false;
    }

    if (qr_scene_upload(g_ctx, blob, bytes) != QR_OK)
    {
        qr_throw("B200 scene upload failed", g_ctx);
    }
    const double t2 = timing ? qr_now_ms() : 0.0;
    {
        /* QR_B200_EXPECT_DEVICE_TILING=1 (tests): the application runs the
         * engine with RT_OPTS_TILING off and expects the tile lists to be
         * built on the device */
        static const bool expect = getenv("QR_B200_EXPECT_DEVICE_TILING") != RT_NULL;
        if (expect)
        {
            qr_kernel_info ki;
            if (qr_kernel_query(g_ctx, &ki) != QR_OK || !ki.device_tiling)
            {
                throw rt_Exception("B200 backend: device-side tiling expected, but the scene came tiled by the engine");
            }
        }
    }
    if (g_pipelined)
    {
        /*
         * QR_B200_PIPELINE=1: this call queues frame N and hands back frame
         * N - 1, so the engine's update phases of the next frame overlap the
         * GPU.  The very first frame (and the first after a change of frame
         * buffer or geometry) is rendered synchronously, the call after it
         * leaves the caller's buffer as it is: the caller sees F0 F0 F1 F2 ...
         */
        int ticket = -1;
        if (qr_render_begin(g_ctx, &ticket) != QR_OK)
        {
            qr_throw("B200 render failed", g_ctx);
        }
        g_tb = timing ? qr_now_ms() : 0.0;
        int rc = QR_OK;
        if (g_prev.ticket >= 0)
        {
            /* completes the fetch started above, or drops a frame that belongs
             * to another buffer / geometry */
            rc = qr_render_end(g_ctx, g_prev.ticket, RT_NULL, 0);
            g_prev.ticket = -1;
            g_prev.fetching = false;
        }
        if (rc == QR_OK && !same)
        {
            /* nothing older belongs in this buffer: wait for this very frame */
            rc = qr_render_end(g_ctx, ticket, (uint32_t *)s_inf->frame, (int)s_inf->frm_row);
        }
        else
        if (rc == QR_OK)
        {
            g_prev.ticket = ticket;
        }
        if (rc != QR_OK)
        {
            qr_throw("B200 render failed", g_ctx);
        }
        g_prev.frame = s_inf->frame; g_prev.row = s_inf->frm_row;
        g_prev.w = s_inf->frm_w;     g_prev.h = s_inf->frm_h;
    }
    else
    if (qr_render(g_ctx, (uint32_t *)s_inf->frame, (int)s_inf->frm_row) != QR_OK)
    {
        qr_throw("B200 render failed", g_ctx);
    }
    if (s_inf->pt_on != 0)
    {
        memcpy(s_inf->ptr_r, &pt_mark, sizeof(pt_mark));
        RT_SIMD_SET(s_inf->pts_c, (rt_real)qr_pt_frames(g_ctx));
    }
    else
    {
        RT_SIMD_SET(s_inf->pts_c, 0.0f);
    }
    static double t_last = 0.0, sum[6] = { 0, 0, 0, 0, 0, 0 };
    if (timing)
    {
        /* running means over all frames but the first few (warm-up), printed
         * every 64th frame */
        const double t3 = qr_now_ms();
        if (frame_no >= 8)
        {
            sum[0] += t1 - t0; sum[1] += t2 - t1; sum[2] += t3 - t2;
            sum[3] += t0 - t_last; sum[4] += g_tb - t2; sum[5] += t3 - g_tb;
        }
        if (frame_no >= 8 && ((frame_no - 7) & 63) == 0)
        {
            const double n = (double)(frame_no - 7);
            if (g_pipelined)
            {
                fprintf(stderr, "B200 render0 (pipelined, mean of %d): since last render0 %.3f ms (engine update + app), "
                                "queue frame %.3f ms, collect previous frame %.3f ms\n",
                                (int)n, sum[3] / n, sum[4] / n, sum[5] / n);
            }
            fprintf(stderr, "B200 render0 (mean of %d): flatten %.3f ms, upload %.3f ms, render + frame copy %.3f ms "
                            "(blob %u bytes)\n", (int)n, sum[0] / n, sum[1] / n, sum[2] / n, (unsigned)bytes);
        }
        frame_no++;
    }
    if (timing)
    {
        t_last = qr_now_ms();
    }
}

/*
 * engine.cpp:4102-4117, 4162-4308 reference the Fresnel plotters of the
 * 128v4 target (debug aid behind core_test -z / RooT F-keys, outside the
 * render0 path).  Scalar statements of the same formulas (the reference's own
 * RT_PLOT_FUNCS_REF code, tracer.cpp:5413-5434, 5509-5533, 5617-5633,
 * 5675-5700, shows them in C); they keep plot_funcs() working.
 */
namespace simd_128v4
{

/* unpolarised dielectric reflectance; i_cos is the (negative) ray.normal dot */
rt_void plot_fresnel(rt_SIMD_INFOP *s_inf)
{
    for (int lane = 0; lane < 4; lane++)
    {
        const float eta = s_inf->c_rfr[lane];
        const float ci = -s_inf->i_cos[lane];
        const float st2 = eta * eta * (1.0f - ci * ci);
        float refl = 1.0f;                      /* total inner reflection */
        if (!(st2 > 1.0f))
        {
            const float ct = sqrtf(1.0f - st2);
            const float rs = (eta * ci - ct) / (eta * ci + ct);
            const float rp = (ci - eta * ct) / (ci + eta * ct);
            refl = 0.5f * (rs * rs + rp * rp);
        }
        s_inf->o_rfl[lane] = refl;
    }
}

/* Schlick's polynomial approximation of the same curve */
rt_void plot_schlick(rt_SIMD_INFOP *s_inf)
{
    for (int lane = 0; lane < 4; lane++)
    {
        const float eta = s_inf->c_rfr[lane];
        float cx = -s_inf->i_cos[lane];
        float base = (eta - 1.0f) / (eta + 1.0f);
        base = base * base;
        bool tir = false;
        if (eta > 1.0f)
        {
            const float st2 = eta * eta * (1.0f - cx * cx);
            tir = st2 > 1.0f;
            if (!tir) cx = sqrtf(1.0f - st2);
        }
        const float w = 1.0f - cx;
        s_inf->o_rfl[lane] = tir ? 1.0f : base + (1.0f - base) * (w * w) * (w * w) * w;
    }
}

/* conductor reflectance, approximate form used by render0 (tracer.cpp:3729-3751) */
rt_void plot_fresnel_metal_fast(rt_SIMD_INFOP *s_inf)
{
    for (int lane = 0; lane < 4; lane++)
    {
        const float c  = -s_inf->i_cos[lane];
        const float n  = s_inf->c_rcp[lane];
        const float nk = n * n + s_inf->ext_2[lane];
        const float c2 = c * c;
        const float tw = 2.0f * n * c;
        const float rs = (nk - tw + c2) / (nk + tw + c2);
        const float rp = (nk * c2 - tw + 1.0f) / (nk * c2 + tw + 1.0f);
        s_inf->o_rfl[lane] = 0.5f * (rs + rp);
    }
}

/* conductor reflectance, full form (compiled out of render0: RT_FEAT_FRESNEL_METAL_SLOW is 0) */
rt_void plot_fresnel_metal_slow(rt_SIMD_INFOP *s_inf)
{
    for (int lane = 0; lane < 4; lane++)
    {
        const float c   = -s_inf->i_cos[lane];
        const float n2  = s_inf->c_rcp[lane] * s_inf->c_rcp[lane];
        const float k2  = s_inf->ext_2[lane];
        const float c2  = c * c, s2 = 1.0f - c * c;
        const float u   = n2 - k2 - s2;
        const float ab  = sqrtf(u * u + 4.0f * n2 * k2);
        const float a   = sqrtf(0.5f * (ab + u));
        const float rs  = (ab + c2 - 2.0f * a * c) / (ab + c2 + 2.0f * a * c);
        const float q   = c2 * ab + s2 * s2, z = 2.0f * a * c * s2;
        const float rp  = rs * (q - z) / (q + z);
        s_inf->o_rfl[lane] = 0.5f * (rs + rp);
    }
}

}
