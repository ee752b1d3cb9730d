/*
 * tracer_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * A stand-in for the reference's core/tracer translation units that routes
 * rt_Platform::render0 through the host flattener (the product's
 * quadray-engine_b200/host/qr_flatten.cpp) and then through the C restatement
 * oracle/render0_oracle.c.  Linked with the unmodified reference engine into
 * oracle/_ref/qr_oracle_harness, it is how
 *   - the flattener + restatement pair is pinned against frames of the
 *     unmodified reference (tools/make_golden.py, tests/test_oracle_golden.py)
 *   - scene blobs for the GPU-side tests are produced (QR_DUMP_BLOB=path).
 * The product never links this file.
 */

#include <stdio.h>
#include <sys/time.h>
#include <stdlib.h>
#include <string.h>

#include "tracer.h"
#include "format.h"
#include "engine.h"

#include "qr_flatten.h"
#include "render0_oracle.h"

/* same rules as core/tracer/tracer.cpp:5774-5808 */
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

/* one fixed 32-lane target, reported as 512x2v2 (tracer.cpp:5827-5876) */
rt_si32 rt_Platform::switch0(rt_SIMD_INFOX *s_inf, rt_si32 simd)
{
    s_mask = 0x02000000;
    s_mode = 0x02000000;
    return simd_init(4, 2, 2);
}

static qr_Flattener g_flat;

rt_void rt_Platform::render0(rt_SIMD_INFOX *s_inf)
{
    if (s_inf->index != 0)
    {
        return;
    }

    size_t bytes = 0;
    static double flat_ms = 0.0;
    static int flat_n = 0;
    timeval fa, fb;
    gettimeofday(&fa, NULL);
    const uint8_t *blob = g_flat.build(s_inf, &bytes);
    gettimeofday(&fb, NULL);
    /* QR_FLATTEN_TIME=1: running mean of the flattener's time per frame, as the
     * product pays it (lists fresh from the update phases on other cores) */
    if (getenv("QR_FLATTEN_TIME") != NULL)
    {
        flat_ms += (fb.tv_sec - fa.tv_sec) * 1e3 + (fb.tv_usec - fa.tv_usec) / 1e3;
        if ((++flat_n % 50) == 0)
        {
            fprintf(stderr, "flatten: mean %.3f ms over %d frames\n", flat_ms / flat_n, flat_n);
        }
    }

    /* QR_FLATTEN_REPEAT=n: time n more flattener passes of this frame (the
     * product's host-side cost, measurable without a GPU) */
    const char *rep = getenv("QR_FLATTEN_REPEAT");
    if (rep != NULL && atoi(rep) > 0)
    {
        timeval a, b;
        gettimeofday(&a, NULL);
        for (int i = 0; i < atoi(rep); i++)
        {
            blob = g_flat.build(s_inf, &bytes);
        }
        gettimeofday(&b, NULL);
        fprintf(stderr, "flatten: %.3f ms per pass (%zu bytes)\n",
                ((b.tv_sec - a.tv_sec) * 1e3 + (b.tv_usec - a.tv_usec) / 1e3) / atoi(rep), bytes);
    }

    const char *dump = getenv("QR_DUMP_BLOB");
    if (dump != NULL && dump[0] != 0)
    {
        FILE *f = fopen(dump, "wb");
        if (f != NULL)
        {
            fwrite(blob, 1, bytes, f);
            fclose(f);
        }
    }

    const char *ps = getenv("QR_ORACLE_PACKET");
    int packet = ps != NULL ? atoi(ps) : 1;

    /* QR_ORACLE_ROWS=n: render only the first n rows (fixture generation
     * that is after the blob, not the frame) */
    const char *rs = getenv("QR_ORACLE_ROWS");
    int rows = rs != NULL ? atoi(rs) : (int)s_inf->frm_h;
    if (rows < 0 || rows > (int)s_inf->frm_h) rows = (int)s_inf->frm_h;

    qr_oracle_stats st;
    memset(&st, 0, sizeof(st));
    int rc;
    if (s_inf->pt_on != 0)
    {
        /* path tracer: the scene's own seed / colour planes, the sample count
         * where render0 keeps it (tracer.cpp:1112-1136) */
        qr_oracle_pt pt;
        pt.pseed = (uint32_t *)s_inf->pseed;
        pt.ptr_r = (float *)s_inf->ptr_r;
        pt.ptr_g = (float *)s_inf->ptr_g;
        pt.ptr_b = (float *)s_inf->ptr_b;
        pt.pts_c = s_inf->pts_c[0];
        rc = qr_oracle_render_pt(blob, bytes, (uint32_t *)s_inf->frame,
                                 (int)s_inf->frm_row, packet, 0, rows, &pt);
        RT_SIMD_SET(s_inf->pts_c, pt.pts_c);
    }
    else
    {
        RT_SIMD_SET(s_inf->pts_c, 0.0f);
        rc = qr_oracle_render(blob, bytes, (uint32_t *)s_inf->frame,
                              (int)s_inf->frm_row, packet, NULL,
                              0, rows, &st);
    }
    if (rc != 0)
    {
        throw rt_Exception("qr_oracle_render failed");
    }

    const char *sf = getenv("QR_ORACLE_STATS");
    if (sf != NULL && sf[0] != 0)
    {
        FILE *f = fopen(sf, "w");
        if (f != NULL)
        {
            fprintf(f, "{\"rays_primary\": %llu, \"rays_shadow\": %llu, "
                       "\"rays_reflect\": %llu, \"rays_refract\": %llu, "
                       "\"surf_visits\": %llu, \"shaded_hits\": %llu, "
                       "\"tex_nonplane\": %llu, \"blob_bytes\": %llu}\n",
                    (unsigned long long)st.rays_primary,
                    (unsigned long long)st.rays_shadow,
                    (unsigned long long)st.rays_reflect,
                    (unsigned long long)st.rays_refract,
                    (unsigned long long)st.surf_visits,
                    (unsigned long long)st.shaded_hits,
                    (unsigned long long)st.tex_nonplane,
                    (unsigned long long)bytes);
            fclose(f);
        }
    }
}

/* engine.cpp:4102-4117 references these debug plotters of the 128v4 target */
namespace simd_128v4
{
rt_void plot_fresnel(rt_SIMD_INFOP *s_inf) {}
rt_void plot_schlick(rt_SIMD_INFOP *s_inf) {}
rt_void plot_fresnel_metal_fast(rt_SIMD_INFOP *s_inf) {}
rt_void plot_fresnel_metal_slow(rt_SIMD_INFOP *s_inf) {}
}
