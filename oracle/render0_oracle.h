/*
 * render0_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's render0 (core/tracer/tracer.cpp:1081-5405)
 * operating on the flattened scene blob (include/qr_scene_blob.h).  It exists
 * to check the CUDA path; nothing in the product may call it.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * load this library.
 */
#ifndef RENDER0_ORACLE_H
#define RENDER0_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qr_oracle_stats
{
    uint64_t rays_primary;    /* samples traced from the camera */
    uint64_t rays_shadow;     /* lanes entering a shadow context */
    uint64_t rays_reflect;    /* lanes entering a reflection context */
    uint64_t rays_refract;    /* lanes entering a transparency context */
    uint64_t surf_visits;     /* lane x list-element visits (all contexts) */
    uint64_t shaded_hits;     /* lanes passing into MT_mat (non-shadow) */
    uint64_t tex_nonplane;    /* shaded textured hits on non-plane surfaces */
} qr_oracle_stats;

/*
 * Render one frame.
 *   blob/bytes   scene blob
 *   frame        y_res rows of "stride" pixels (0x00RRGGBB); only x < x_res
 *                is written
 *   packet       emulated SIMD width S of the reference target (4..64, a
 *                multiple of 4): every packet-wide early-out of the macro
 *                assembler (CHECK_MASK) is taken exactly as a reference build
 *                of that width takes it.  packet = 1 selects per-sample
 *                semantics (each sample decides alone) -- what a scalar
 *                one-thread-per-sample GPU kernel computes.
 *   t_out        optional, per primary sample T_BUF at XX_end ("dump mode"),
 *                y_res * x_res * (1 << fsaa) floats, sample-major within pixel
 *   y0, y1       row range [y0, y1)
 * Returns 0 or a negative error code.
 */
int qr_oracle_render(const void *blob, size_t bytes, uint32_t *frame,
                     int stride, int packet, float *t_out,
                     int y0, int y1, qr_oracle_stats *stats);

/*
 * Path-tracer state of a scene (rt_Scene::pseed, ptr_r/g/b, engine.cpp:
 * 2875-2901; rt_SIMD_INFOX::pts_c): 4 * x_row * y_res entries per plane, slot
 * ((y * x_row + x) << fsaa) + sample.  reset = rt_Scene::reset_pseed /
 * reset_color (engine.cpp:3670-3700) and pts_c = 0.
 */
typedef struct qr_oracle_pt
{
    uint32_t *pseed;
    float    *ptr_r, *ptr_g, *ptr_b;
    float     pts_c;
} qr_oracle_pt;

/* the seed plane as rt_Scene::reset_pseed fills it */
void qr_oracle_pt_seed(uint32_t *pseed, size_t n);

/*
 * One more frame of the path tracer (rt_Scene::set_pton(1), render0 with
 * pt_on, tracer.cpp:1112-1136, 1218-1285, 2339-2701, 3428-3466, 5176-5219).
 * With packet = S the result is the reference's of that SIMD width; PT results
 * depend on the width (unmasked side effects of packet-wide branches).
 */
int qr_oracle_render_pt(const void *blob, size_t bytes, uint32_t *frame,
                        int stride, int packet, int y0, int y1, qr_oracle_pt *pt);

#ifdef __cplusplus
}
#endif

#endif /* RENDER0_ORACLE_H */
