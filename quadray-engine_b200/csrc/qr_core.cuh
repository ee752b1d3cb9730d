/*
 * qr_core.cuh -- per-sample ray tracer core of the B200 render0 path.
 *
 * One thread owns one ray sample and runs this state machine to completion:
 *
 *      +--> WALK  (one list walk: closest hit, or any-hit for a shadow ray)
 *      |      |
 *      |      v
 *      |    SHADE (normal, texel, ambient) -> LIGHTS (one shadow WALK each)
 *      |      -> TRANSPARENCY (push a refraction child -> WALK)
 *      |      -> REFLECTION   (push a reflection child -> WALK)
 *      +------ pop / mix colours back up to the primary ray
 *
 * All rays of a warp -- primary, shadow, reflected, refracted -- funnel
 * through the same WALK loop, so the warp re-converges on the hot code after
 * every shading step.  Recursion of the reference (context stack bumped by
 * RT_STACK_STEP, tracer.cpp:2806, 3507, 3841) becomes an explicit per-thread
 * stack of qr_frame records.
 *
 * Semantics are those of the reference's render0 (core/tracer/tracer.cpp:
 * 1081-5405) evaluated for ONE lane: every packet-wide early-out (CHECK_MASK)
 * is decided by the sample alone.  The reference shades every surface that
 * passes the running depth test in list order and lets the nearest overwrite
 * the colour (tracer.cpp:1602-1605, 641-662); here the nearest candidate is
 * found first (ties keep the earliest, as t_buf > t_val is strict) and shaded
 * once -- same pixels, no overdraw shading.  oracle/render0_oracle.c with
 * packet = 1 is the executable statement of these semantics.
 *
 * Arithmetic: every add/sub/mul/div/sqrt is a separately rounded IEEE-754
 * binary32 operation in the order of the reference asm (never contracted to
 * FMA), rsq = 1/sqrt and rcp = 1/x (rtconf.h:164-193), compare predicates as
 * in rtarch_x32_512x2v2.h:706-880 (cgt = NLE, cge = NLT, cne = NEQ_UQ).
 *
 * The file compiles as CUDA device code (the product) and as plain C++ host
 * code (tests/hostsim, CPU-only unit tests of this logic).
 */
#ifndef QR_CORE_CUH
#define QR_CORE_CUH

#include <stdint.h>
#include "qr_scene_blob.h"

#if defined(__CUDACC__)
#define QR_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#include <string.h>
#define QR_HD static inline
#endif

/* ---- rounded arithmetic ---------------------------------------------------- */

#if defined(__CUDA_ARCH__)
QR_HD float qr_add(float a, float b) { return __fadd_rn(a, b); }
QR_HD float qr_sub(float a, float b) { return __fsub_rn(a, b); }
QR_HD float qr_mul(float a, float b) { return __fmul_rn(a, b); }
QR_HD float qr_div(float a, float b) { return __fdiv_rn(a, b); }
QR_HD float qr_sqrt(float a)         { return __fsqrt_rn(a); }
QR_HD float    qr_u2f(uint32_t u)    { return __uint_as_float(u); }
QR_HD uint32_t qr_f2u(float f)       { return __float_as_uint(f); }
/* cvmps / cvnps with the x86 "integer indefinite" for out-of-range inputs */
QR_HD int32_t qr_cvm(float x)
{
    return (x >= -2147483648.0f && x < 2147483648.0f) ? __float2int_rd(x)
                                                      : (int32_t)0x80000000u;
}
QR_HD int32_t qr_cvn(float x)
{
    return (x >= -2147483648.0f && x < 2147483648.0f) ? __float2int_rn(x)
                                                      : (int32_t)0x80000000u;
}
#else
/* host build (tests/hostsim): QR_COUNT_OPS tallies the algorithmic IEEE
 * operations -- the numerator of the FP32 roofline bench.py reports */
#ifdef QR_COUNT_OPS
extern unsigned long long qr_ops[4];    /* add/sub, mul, div, sqrt */
#define QR_OP(i) (qr_ops[i]++)
#else
#define QR_OP(i) ((void)0)
#endif
QR_HD float qr_add(float a, float b) { QR_OP(0); volatile float r = a + b; return r; }
QR_HD float qr_sub(float a, float b) { QR_OP(0); volatile float r = a - b; return r; }
QR_HD float qr_mul(float a, float b) { QR_OP(1); volatile float r = a * b; return r; }
QR_HD float qr_div(float a, float b) { QR_OP(2); volatile float r = a / b; return r; }
QR_HD float qr_sqrt(float a)         { QR_OP(3); return sqrtf(a); }
QR_HD float    qr_u2f(uint32_t u)    { float f; memcpy(&f, &u, 4); return f; }
QR_HD uint32_t qr_f2u(float f)       { uint32_t u; memcpy(&u, &f, 4); return u; }
QR_HD int32_t qr_cvm(float x)
{
    return (x >= -2147483648.0f && x < 2147483648.0f) ? (int32_t)floorf(x)
                                                      : (int32_t)0x80000000u;
}
QR_HD int32_t qr_cvn(float x)
{
    return (x >= -2147483648.0f && x < 2147483648.0f) ? (int32_t)lrintf(x)
                                                      : (int32_t)0x80000000u;
}
#endif

QR_HD float qr_rsq(float a)  { return qr_div(1.0f, qr_sqrt(a)); }
QR_HD float qr_neg(float a)  { return qr_u2f(qr_f2u(a) ^ 0x80000000u); }
QR_HD float qr_abs(float a)  { return qr_u2f(qr_f2u(a) & 0x7FFFFFFFu); }
/* x ^ sign-bit when "flip" is set (srf_SBASE / srf_SMASK select) */
QR_HD float qr_sgn(float a, int flip) { return flip ? qr_neg(a) : a; }

/* compare predicates of the reference target */
QR_HD bool qr_gt(float a, float b) { return !(a <= b); }    /* cgt = NLE */
QR_HD bool qr_ge(float a, float b) { return !(a <  b); }    /* cge = NLT */

/* ---- scene view ------------------------------------------------------------ */

struct qr_view
{
    const qr_blob_header *h;
    const qr_surface     *surfs;
    const qr_material    *mats;
    const qr_light       *lgts;
    const qr_elem        *elems;
    const int32_t        *tiles;
    const uint32_t       *texels;
};

QR_HD void qr_view_init(qr_view &v, const void *blob)
{
    const uint8_t *b = (const uint8_t *)blob;
    const qr_blob_header *h = (const qr_blob_header *)blob;
    v.h      = h;
    v.surfs  = (const qr_surface  *)(b + h->off_surf);
    v.mats   = (const qr_material *)(b + h->off_mat);
    v.lgts   = (const qr_light    *)(b + h->off_lgt);
    v.elems  = (const qr_elem     *)(b + h->off_elem);
    v.tiles  = (const int32_t     *)(b + h->off_tiles);
    v.texels = (const uint32_t    *)(b + h->off_texels);
}

/* ray counters, SURVEY.md 8(d): one ray = one list walk for one sample */
struct qr_counters
{
    uint32_t shadow, reflect, refract;
};

/* continuation of a level that waits for a child ray */
struct qr_frame
{
    float   col[3];          /* COL of the level so far */
    float   ray[3];          /* RAY_X/Y/Z of the level */
    float   hit[3];          /* HIT_X/Y/Z */
    float   nrm[3];          /* NRM_X/Y/Z */
    float   loc[3];          /* NRM_I/J/K: stored local hit (tracer.cpp:2272-2282) */
    float   c_trn, c_rfl;    /* ctx_C_TRN / ctx_C_RFL */
    int32_t ei;              /* list element of the surface being shaded */
    int32_t flg;             /* ctx_LOCAL(FLG): side | props */
    int32_t stage;           /* 0: child is the refraction ray, 1: reflection */
};

#define QR_MODE_CLOSEST 0
#define QR_MODE_SHADOW  1

/*
 * 3x3 transform, tracer.cpp:1447-1479 / 1512-1548 / 2063-2095: diagonal
 * products first, then the off-diagonal terms of each row in column order;
 * a_map[L] == 1 keeps the diagonal only (scaling fast path).
 */
QR_HD void qr_xform(const qr_surface &s, float v1, float v2, float v3,
                    float &o4, float &o5, float &o6)
{
    float x4 = qr_mul(s.tci[0], v1);
    float x5 = qr_mul(s.tcj[1], v2);
    float x6 = qr_mul(s.tck[2], v3);
    if (s.a_map[3] != 1)
    {
        x4 = qr_add(x4, qr_mul(s.tci[1], v2));
        x4 = qr_add(x4, qr_mul(s.tci[2], v3));
        x5 = qr_add(x5, qr_mul(s.tcj[0], v1));
        x5 = qr_add(x5, qr_mul(s.tcj[2], v3));
        x6 = qr_add(x6, qr_mul(s.tck[0], v1));
        x6 = qr_add(x6, qr_mul(s.tck[1], v2));
    }
    o4 = x4; o5 = x5; o6 = x6;
}

/* value of a clipper's implicit function at point p (clipper space),
 * PL_clp 4198-4208, QD_clp 4910-4951, TP_clp 4341-4370 */
QR_HD bool qr_clip_eval(const qr_surface &cs, int side_data,
                        const float *px /* NRM_X/Y/Z */, const float *pi /* NRM_I/J/K */,
                        bool &valid)
{
    const float *p = cs.a_sgn[3] ? pi : px;
    float v;
    valid = true;
    if (cs.srf_t[2] == 1)
    {
        const int k = cs.a_map[2];
        v = qr_sgn(k < 3 ? px[k] : pi[k - 3], cs.a_sgn[2]);
    }
    else
    if (cs.srf_t[2] == 2)
    {
        float a1 = qr_mul(qr_add(cs.scj[0], cs.scj[0]), p[0]);
        float a4 = qr_sub(qr_mul(qr_mul(p[0], p[0]), cs.sci[0]), a1);
        float a2 = qr_mul(qr_add(cs.scj[1], cs.scj[1]), p[1]);
        float a5 = qr_sub(qr_mul(qr_mul(p[1], p[1]), cs.sci[1]), a2);
        float a3 = qr_mul(qr_add(cs.scj[2], cs.scj[2]), p[2]);
        float a6 = qr_sub(qr_mul(qr_mul(p[2], p[2]), cs.sci[2]), a3);
        a4 = qr_sub(a4, cs.sci[3]);
        a4 = qr_add(a4, a5);
        v  = qr_add(a4, a6);
    }
    else
    if (cs.srf_t[2] == 3)
    {
        float a4 = qr_mul(qr_mul(p[0], p[0]), cs.sci[0]);
        float a5 = qr_mul(qr_mul(p[1], p[1]), cs.sci[1]);
        float a6 = qr_mul(qr_mul(p[2], p[2]), cs.sci[2]);
        a4 = qr_sub(a4, cs.sci[3]);
        a4 = qr_add(a4, a5);
        v  = qr_add(a4, a6);
    }
    else
    {
        valid = false;
        return true;
    }
    /* APPLY_CLIP 488-496 */
    return side_data < 0 ? qr_ge(v, 0.0f) : (v <= 0.0f);
}

/* per-walk transform-caching state (ctx DFF / RAY_IJK / LOCAL(OBJ)) */
struct qr_walk_state
{
    float dff[6];            /* DFF_X/Y/Z, DFF_I/J/K */
    float rayi[3];           /* RAY_I/J/K */
    int   l_obj;             /* ctx_LOCAL(OBJ): trnode's last element */
};

/*
 * CC_clp, tracer.cpp:1597-2160, for one candidate root "t" of element "ei".
 * On success "loc" holds the (possibly adjusted) local hit point.
 *   dmask / amask / side: XMISC(PTR) & DMASK lane, AMASK lane, LOCAL(FLG) side
 */
QR_HD bool qr_clip(const qr_view &v, const qr_surface &s, const qr_walk_state &w,
                   const float *org, const float *ray, float t_min, float t_buf,
                   float t, bool dmask, uint32_t amask, int side, float *loc)
{
    bool m = true;
    m = m && qr_gt(t_buf, t);
    m = m && (t_min < t);

    float hit[3];
    hit[0] = qr_add(qr_mul(ray[0], t), org[0]);
    hit[1] = qr_add(qr_mul(ray[1], t), org[1]);
    hit[2] = qr_add(qr_mul(ray[2], t), org[2]);

    const int shift = s.a_sgn[3];
    if (s.a_map[3] != 0)
    {
        loc[0] = qr_add(qr_mul(w.rayi[0], t), w.dff[3]);
        loc[1] = qr_add(qr_mul(w.rayi[1], t), w.dff[4]);
        loc[2] = qr_add(qr_mul(w.rayi[2], t), w.dff[5]);
    }
    else
    {
        loc[0] = qr_sub(hit[0], s.pos[0]);
        loc[1] = qr_sub(hit[1], s.pos[1]);
        loc[2] = qr_sub(hit[2], s.pos[2]);
    }

    /* conic singularity solver 1706-1856 (lane semantics: hmask decides) */
    if (s.conic != 0 && dmask)
    {
        const int iI = s.a_map[0] - shift, iJ = s.a_map[1] - shift, iK = s.a_map[2] - shift;
        float a0 = qr_mul(loc[iI], loc[iI]);
        if (s.conic != 2)
        {
            a0 = qr_add(a0, qr_mul(loc[iJ], loc[iJ]));
        }
        a0 = qr_add(a0, qr_mul(loc[iK], loc[iK]));
        if (a0 < s.t_eps)
        {
            const float *df = w.dff + shift;
            uint32_t q1 = (qr_f2u(df[iI]) & 0x80000000u) ^ 0x3F800000u;
            uint32_t q2 = 0;
            float q3 = s.sci[iI];
            float q4 = 1.0f;
            if (s.conic != 2)
            {
                q2 = (qr_f2u(df[iJ]) & 0x80000000u) ^ 0x3F800000u;
                q3 = qr_add(q3, s.sci[iJ]);
                q4 = qr_add(q4, 1.0f);
            }
            q3 = qr_div(q3, s.sci[iK]);
            q3 = qr_neg(q3);
            float q6 = q3;
            q3 = qr_sqrt(q3);
            q6 = qr_add(q6, q4);
            q4 = qr_rsq(q6);
            q4 = qr_mul(q4, s.t_eps);
            float p1 = qr_mul(qr_u2f(q1), q4);
            float p2 = qr_mul(qr_u2f(q2), q4);
            float p3 = qr_mul(q3, q4);
            const uint32_t ts = side ? 0x80000000u : 0u;
            uint32_t u3 = qr_f2u(p3) ^ (qr_f2u(df[iK]) & 0x80000000u);
            u3 ^= (ts & amask) ^ amask;
            const uint32_t tsn = (ts | amask) ^ amask;
            loc[iI] = qr_u2f(qr_f2u(p1) ^ tsn);
            if (s.conic != 2)
            {
                loc[iJ] = qr_u2f(qr_f2u(p2) ^ tsn);
            }
            loc[iK] = qr_u2f(u3);
        }
    }

    /* axis min/max 1874-1927 */
    const int mm = s.minmax_t;
    if (mm & 1)  m = m && (s.min[0] <= loc[0]);
    if (mm & 8)  m = m && qr_ge(s.max[0], loc[0]);
    if (mm & 2)  m = m && (s.min[1] <= loc[1]);
    if (mm & 16) m = m && qr_ge(s.max[1], loc[1]);
    if (mm & 4)  m = m && (s.min[2] <= loc[2]);
    if (mm & 32) m = m && qr_ge(s.max[2], loc[2]);

    /* custom clippers 1931-2151.  The reference evaluates the whole list for
     * the packet; a lone sample may stop as soon as its mask is clear and no
     * accumulator is open (a cleared mask can only come back through an
     * accum enter/leave pair). */
    int di = s.clip_head;
    if (di == QR_NIL)
    {
        return m;
    }

    float nx[3] = {0.0f, 0.0f, 0.0f};           /* NRM_X/Y/Z */
    float ni[3] = {0.0f, 0.0f, 0.0f};           /* NRM_I/J/K */
    bool  acc = false, in_acc = false;          /* C_ACC, inside enter..leave */
    int   redx = QR_NIL;
    bool  last = true;                          /* stale Xmm4 stand-in */

    for (; di != QR_NIL; di = v.elems[di].next)
    {
        if (!m && !in_acc)
        {
            return false;
        }
        const qr_elem ce = v.elems[di];

        if (ce.simd == QR_NIL)
        {
            if (ce.data_i > 0)
            {
                m = !m && acc;                  /* annpx: ~mask & C_ACC */
                in_acc = false;
            }
            else
            {
                acc = m;
                m = s.c_def != 0;
                in_acc = true;
            }
            continue;
        }

        const qr_surface &cs = v.surfs[ce.simd];
        bool have_local = false;

        if (cs.srf_t[3] >= 0)
        {
            if (redx != QR_NIL)
            {
                ni[0] = qr_sub(nx[0], cs.pos[0]);
                ni[1] = qr_sub(nx[1], cs.pos[1]);
                ni[2] = qr_sub(nx[2], cs.pos[2]);
                if (di == redx) redx = QR_NIL;
                have_local = true;
            }
        }
        else
        if (ce.simd == s.trnode)
        {
            nx[0] = qr_add(loc[0], s.pos[0]);
            nx[1] = qr_add(loc[1], s.pos[1]);
            nx[2] = qr_add(loc[2], s.pos[2]);
            redx = ce.data_p;
            continue;
        }

        if (!have_local)
        {
            nx[0] = qr_sub(hit[0], cs.pos[0]);
            nx[1] = qr_sub(hit[1], cs.pos[1]);
            nx[2] = qr_sub(hit[2], cs.pos[2]);
            if (cs.a_map[3] != 0)
            {
                float o4, o5, o6;
                qr_xform(cs, nx[0], nx[1], nx[2], o4, o5, o6);
                if (cs.srf_t[3] < 0)
                {
                    nx[0] = o4; nx[1] = o5; nx[2] = o6;
                    redx = ce.data_p;
                    continue;
                }
                ni[0] = o4; ni[1] = o5; ni[2] = o6;
            }
        }

        bool valid;
        bool r = qr_clip_eval(cs, ce.data_i, nx, ni, valid);
        if (valid) last = r;
        m = m && last;
    }
    return m;
}

/*
 * Shadow applicability of a hit, CHECK_SHAD 549-589: light-emitting and
 * fully-transparent non-refractive surfaces do not cast shadows.
 */
QR_HD bool qr_casts_shadow(int props)
{
    if (props & QR_PROP_LIGHT) return false;
    if ((props & QR_PROP_TRANSP) && !(props & QR_PROP_REFRACT)) return false;
    return true;
}

/*
 * One list walk, OO_cyc 1341 .. OO_out 5142, for one sample.
 *   mode CLOSEST: returns true when something was hit; t_buf / best_* updated
 *   mode SHADOW : returns true when the sample is in shadow (first occluder)
 * "ploc" is the stored local hit of the originating level (NRM_I/J/K of the
 * previous context), used when the ray starts on the surface being tested.
 */
QR_HD bool qr_walk(const qr_view &v, int head, int mode,
                   const float *org, const float *ray, float t_min, float t_max,
                   int p_obj, int p_flg, const float *ploc,
                   float &t_buf, int &best_ei, int &best_side, float *best_loc)
{
    qr_walk_state w;
    w.dff[0] = w.dff[1] = w.dff[2] = w.dff[3] = w.dff[4] = w.dff[5] = 0.0f;
    w.rayi[0] = w.rayi[1] = w.rayi[2] = 0.0f;
    w.l_obj = QR_NIL;

    t_buf = t_max;
    best_ei = QR_NIL;
    best_side = 0;

    for (int ei = head; ei != QR_NIL; ei = v.elems[ei].next)
    {
        const qr_elem e = v.elems[ei];
        const int si = e.simd;
        const qr_surface &s = v.surfs[si];
        const bool same = (si == p_obj);
        const int shift = s.a_sgn[3];

        /* 1352-1373 */
        if (same)
        {
            w.dff[shift + 0] = ploc[0];
            w.dff[shift + 1] = ploc[1];
            w.dff[shift + 2] = ploc[2];
        }

        if (!(s.srf_t[3] < 0) && w.l_obj != QR_NIL)
        {
            /* 1385-1417: transform caching under a trnode */
            if (!same)
            {
                w.dff[3] = qr_sub(w.dff[0], s.pos[0]);
                w.dff[4] = qr_sub(w.dff[1], s.pos[1]);
                w.dff[5] = qr_sub(w.dff[2], s.pos[2]);
            }
            if (ei == w.l_obj) w.l_obj = QR_NIL;
        }
        else
        {
            /* OO_dff 1419-1556 */
            bool do_ray = same;
            if (!same)
            {
                w.dff[0] = qr_sub(org[0], s.pos[0]);
                w.dff[1] = qr_sub(org[1], s.pos[1]);
                w.dff[2] = qr_sub(org[2], s.pos[2]);
                if (s.a_map[3] != 0)
                {
                    float o4, o5, o6;
                    qr_xform(s, w.dff[0], w.dff[1], w.dff[2], o4, o5, o6);
                    if (s.srf_t[3] < 0)
                    {
                        w.dff[0] = o4; w.dff[1] = o5; w.dff[2] = o6;
                        w.l_obj = e.data_p;
                    }
                    else
                    {
                        w.dff[3] = o4; w.dff[4] = o5; w.dff[5] = o6;
                    }
                    do_ray = true;
                }
            }
            if (do_ray)
            {
                qr_xform(s, ray[0], ray[1], ray[2], w.rayi[0], w.rayi[1], w.rayi[2]);
            }
        }

        const float *lray = shift ? w.rayi : ray;       /* RAY at a_sgn[L] */
        const float *ldff = w.dff + shift;              /* DFF at a_sgn[L] */

        /* AR_ptr 3955-4054: bounding volume of an array */
        if (e.data_i == 1)
        {
            float x1 = lray[0];
            float x0 = qr_mul(s.sci[0], x1);
            float x5 = ldff[0];
            float q7 = qr_mul(s.sci[0], x5);
            float x3 = x1;
            x1 = qr_mul(x1, x0); x3 = qr_mul(x3, q7); x5 = qr_mul(x5, q7);

            float x2 = lray[1];
            x0 = qr_mul(s.sci[1], x2);
            float x6 = ldff[1];
            q7 = qr_mul(s.sci[1], x6);
            float x4 = x2;
            x2 = qr_mul(x2, x0); x4 = qr_mul(x4, q7); x6 = qr_mul(x6, q7);
            x1 = qr_add(x1, x2); x3 = qr_add(x3, x4); x5 = qr_add(x5, x6);

            x2 = lray[2];
            x0 = qr_mul(s.sci[2], x2);
            x6 = ldff[2];
            q7 = qr_mul(s.sci[2], x6);
            x4 = x2;
            x2 = qr_mul(x2, x0); x4 = qr_mul(x4, q7); x6 = qr_mul(x6, q7);
            x1 = qr_add(x1, x2); x3 = qr_add(x3, x4); x5 = qr_add(x5, x6);

            x5 = qr_sub(x5, s.sci[3]);
            x5 = qr_mul(x5, x1);
            x3 = qr_mul(x3, x3);
            x3 = qr_sub(x3, x5);
            if (!(0.0f <= x3))
            {
                ei = e.data_p;                          /* AR_skp */
                if (ei == w.l_obj) w.l_obj = QR_NIL;
            }
            continue;
        }

        const int tag = s.srf_t[0];
        if (tag == 0) continue;

        float loc[3];

        if (tag == 1)
        {
            /* PL_ptr 4062-4136 */
            if (same) continue;
            const int k = s.a_map[2];
            const float dk = qr_neg(qr_sgn(k < 3 ? w.dff[k] : w.dff[k], s.a_sgn[2]));
            const float rk = qr_sgn(k < 3 ? ray[k] : w.rayi[k - 3], s.a_sgn[2]);
            if (!(0.0f != rk)) continue;
            const float t = qr_div(dk, rk);
            if (!qr_clip(v, s, w, org, ray, t_min, t_buf, t, false, 0u, 0, loc)) continue;
            const int side = (rk < 0.0f) ? QR_FLAG_SIDE_OUTER : QR_FLAG_SIDE_INNER;
            if (mode == QR_MODE_SHADOW)
            {
                if (qr_casts_shadow(s.props[side])) return true;
                continue;
            }
            t_buf = t; best_ei = ei; best_side = side;
            best_loc[0] = loc[0]; best_loc[1] = loc[1]; best_loc[2] = loc[2];
            continue;
        }

        float a_val, b_val, c_val, d_val;

        if (tag == 3)
        {
            /* TP_ptr 4216-4277 */
            const int iI = s.a_map[0], iK = s.a_map[2];
            const float sci_i = s.sci[iI - shift], sci_k = s.sci[iK - shift];
            const float ri = iI < 3 ? ray[iI] : w.rayi[iI - 3], di = w.dff[iI];
            const float rk = iK < 3 ? ray[iK] : w.rayi[iK - 3], dk = w.dff[iK];
            float a5 = qr_sub(qr_mul(di, rk), qr_mul(dk, ri));
            a5 = qr_mul(a5, a5);
            a5 = qr_mul(a5, sci_i);
            a5 = qr_mul(a5, sci_k);
            d_val = qr_abs(a5);
            b_val = qr_add(qr_mul(qr_mul(sci_i, di), ri), qr_mul(qr_mul(sci_k, dk), rk));
            c_val = qr_add(qr_mul(qr_mul(di, di), sci_i), qr_mul(qr_mul(dk, dk), sci_k));
            a_val = qr_add(qr_mul(qr_mul(ri, ri), sci_i), qr_mul(qr_mul(rk, rk), sci_k));
        }
        else
        {
            /* QD_ptr 4378-4447 */
            float a1 = lray[0];
            float a0 = qr_mul(s.sci[0], a1);
            float a5 = ldff[0];
            float a7 = qr_sub(qr_mul(s.sci[0], a5), s.scj[0]);
            float a3 = qr_mul(a1, a7);
            a1 = qr_mul(a1, a0);
            a7 = qr_sub(a7, s.scj[0]);
            a5 = qr_mul(a5, a7);

            float a2 = lray[1];
            a0 = qr_mul(s.sci[1], a2);
            float a6 = ldff[1];
            a7 = qr_sub(qr_mul(s.sci[1], a6), s.scj[1]);
            float a4 = qr_mul(a2, a7);
            a2 = qr_mul(a2, a0);
            a7 = qr_sub(a7, s.scj[1]);
            a6 = qr_mul(a6, a7);

            a1 = qr_add(a1, a2); a3 = qr_add(a3, a4); a5 = qr_add(a5, a6);

            a2 = lray[2];
            a0 = qr_mul(s.sci[2], a2);
            a6 = ldff[2];
            a7 = qr_sub(qr_mul(s.sci[2], a6), s.scj[2]);
            a4 = qr_mul(a2, a7);
            a2 = qr_mul(a2, a0);
            a7 = qr_sub(a7, s.scj[2]);
            a6 = qr_mul(a6, a7);

            a1 = qr_add(a1, a2); a3 = qr_add(a3, a4); a5 = qr_add(a5, a6);

            a5 = qr_sub(a5, s.sci[3]);
            a_val = a1;
            b_val = a3;
            c_val = a5;
            d_val = qr_sub(qr_mul(a3, a3), qr_mul(a5, a1));
        }

        /* QD_rts 4449-4547 */
        if (!(0.0f <= d_val)) continue;
        const float b = qr_neg(b_val);
        const bool dmask = d_val < s.d_eps;
        const float sd = qr_u2f(qr_f2u(qr_sqrt(d_val)) ^ (qr_f2u(b) & 0x80000000u));
        const float bd = qr_add(b, sd);
        const bool m_pos = 0.0f <= sd;
        float t1n = m_pos ? c_val : bd;
        float t1d = m_pos ? bd : a_val;
        float t2n = m_pos ? bd : c_val;
        float t2d = m_pos ? a_val : bd;
        const uint32_t amask = qr_f2u(a_val) & 0x80000000u;
        float t1 = 0.0f, t2 = 0.0f;
        bool k1 = true, k2 = true;

        if (dmask)
        {
            /* 4572-4623: near-zero determinant, both roots up front */
            if (t1n == 0.0f) t1d = 1.0f;
            if (t2n == 0.0f) t2d = 1.0f;
            t1 = qr_div(t1n, t1d);
            t2 = qr_div(t2n, t2d);
            k1 = t1d != 0.0f;
            k2 = t2d != 0.0f;
            float a2 = qr_u2f(qr_f2u(qr_sub(t1, t2)) ^ amask);
            const bool fm = 0.0f <= a2;
            a2 = fm ? a2 : 0.0f;
            float a5 = qr_abs(qr_mul(fm ? s.t_eps : 0.0f, t1));
            a2 = qr_sub(qr_mul(a2, -0.5f), a5);
            uint32_t u2 = qr_f2u(a2) ^ amask;
            if (!(k1 && k2)) u2 = 0;
            t1 = qr_add(t1, qr_u2f(u2));
            t2 = qr_sub(t2, qr_u2f(u2));
        }

        /* QD_srt 4646-4824, one lane: the side tried first follows the sign of
         * "a"; a hit on the first side ends the surface (overdraw check) */
        const int first = qr_gt(0.0f, a_val) ? QR_FLAG_SIDE_INNER : QR_FLAG_SIDE_OUTER;
        const int pf = p_flg & (QR_FLAG_SIDE | QR_FLAG_PASS);
        for (int pass = 0; pass < 2; pass++)
        {
            const int side = pass == 0 ? first : (first ^ 1);
            /* CHECK_SIDE 531-540 */
            if (same && (pf == 1 - side || pf == 2 + side)) continue;
            float t; bool k;
            if (side == QR_FLAG_SIDE_OUTER)
            {
                if (dmask) { t = t1; k = k1; }
                else { t = qr_div(t1n, t1d); k = t1d != 0.0f; }
            }
            else
            {
                if (dmask) { t = t2; k = k2; }
                else { t = qr_div(t2n, t2d); k = t2d != 0.0f; }
            }
            if (!k) continue;
            if (!qr_clip(v, s, w, org, ray, t_min, t_buf, t, dmask, amask, side, loc)) continue;
            if (mode == QR_MODE_SHADOW)
            {
                if (qr_casts_shadow(s.props[side])) return true;
                break;
            }
            t_buf = t; best_ei = ei; best_side = side;
            best_loc[0] = loc[0]; best_loc[1] = loc[1]; best_loc[2] = loc[2];
            break;
        }
    }

    return mode == QR_MODE_SHADOW ? false : best_ei != QR_NIL;
}

/* texel -> linear colour, PAINT_COLX 664-673 */
QR_HD float qr_unpack(uint32_t texel, int sh, const qr_material &m, int props)
{
    float c = (float)(int32_t)((texel >> sh) & m.cmask);
    c = qr_div(c, m.clamp);
    if (props & QR_PROP_GAMMA) c = qr_mul(c, c);
    return c;
}

/* fixed-point 28.4 power, tracer.cpp:2981-3039 */
QR_HD float qr_pow_28_4(float x, uint32_t l_pow)
{
    uint32_t eax = l_pow & 0xF;
    float x2 = x, x4 = x, x1 = 1.0f;
    if (eax != 0)
    {
        do
        {
            x4 = qr_sqrt(x4);
            const uint32_t esi = 0x8 & eax;
            eax = (eax << 1) & 0xF;
            if (esi != 0) x1 = qr_mul(x1, x4);
        }
        while (eax != 0);
    }
    eax = l_pow >> 4;
    if (eax != 0)
    {
        const float x3 = x1;
        x1 = 1.0f;
        do
        {
            const uint32_t esi = 1 & eax;
            eax = eax >> 1;
            if (esi != 0) x1 = qr_mul(x1, x2);
            x2 = qr_mul(x2, x2);
        }
        while (eax != 0);
        x1 = qr_mul(x1, x3);
    }
    return x1;
}

/* normalise + dot with the normal, tracer.cpp:3216-3246 / 3620-3653 */
QR_HD float qr_norm_dot(const float *ray, const float *nrm, float *a)
{
    float s0 = qr_mul(ray[0], ray[0]);
    s0 = qr_add(s0, qr_mul(ray[1], ray[1]));
    s0 = qr_add(s0, qr_mul(ray[2], ray[2]));
    const float inv = qr_rsq(s0);
    a[0] = qr_mul(ray[0], inv);
    a[1] = qr_mul(ray[1], inv);
    a[2] = qr_mul(ray[2], inv);
    float d = qr_mul(a[0], nrm[0]);
    d = qr_add(d, qr_mul(a[1], nrm[1]));
    d = qr_add(d, qr_mul(a[2], nrm[2]));
    return d;
}

/* exact dielectric Fresnel, tracer.cpp:3385-3400 / 3781-3796 */
QR_HD float qr_fresnel(float c, float rfr, float x0, float x7)
{
    float a1 = c;
    float a2 = qr_sub(qr_mul(a1, rfr), x7);
    const float a7 = qr_mul(x7, rfr);
    const float a3 = qr_sub(a1, a7);
    a1 = qr_add(a1, a7);
    float a0 = qr_div(x0, a2);
    a1 = qr_div(a1, a3);
    a0 = qr_mul(a0, a0);
    a1 = qr_mul(a1, a1);
    a0 = qr_add(a0, a1);
    a0 = qr_mul(a0, -0.5f);
    return qr_abs(a0);
}

/*
 * Trace one primary sample.  "stack" needs QR_STACK_DEPTH frames.
 * Returns the sample colour (before clamp / AA / gamma) and the primary T_BUF.
 */
QR_HD void qr_trace_sample(const qr_view &v, int px, int py, int lane4,
                           qr_frame *stack, float *out_col, float *out_t,
                           qr_counters *cnt)
{
    const qr_blob_header &h = *v.h;

    /* current ray */
    float org[3], ray[3];
    float t_min, t_max;
    int   head, mode, p_obj, p_flg;
    int   lvl = 0;

    /* shading state of the current level */
    float lray[3] = {0, 0, 0};      /* RAY of the level (ray[] may hold a shadow ray) */
    float hit[3] = {0, 0, 0}, nrm[3] = {0, 0, 0}, loc[3] = {0, 0, 0};
    float tex[3] = {0, 0, 0}, col[3] = {0, 0, 0};
    float dot = 0.0f, c_trn = 0.0f, c_rfl = 0.0f;
    float xr[3] = {0, 0, 0};
    int   cur_ei = QR_NIL, l_flg = 0, li = QR_NIL;

    /* walk results */
    float t_buf; int best_ei, best_side; float best_loc[3] = {0, 0, 0};

    /* 1287-1322: primary ray; hor_i / ver_i are exact integers */
    {
        float hs = qr_add((float)px, h.hor_a[lane4]);
        float vs = qr_add((float)py, h.ver_a[lane4]);
        hs = qr_add(hs, 0.0f);
        vs = qr_add(vs, 0.0f);
        for (int k = 0; k < 3; k++)
        {
            float a = qr_mul(h.hor[k], hs);
            float b = qr_mul(h.ver[k], vs);
            a = qr_add(a, b);
            ray[k] = qr_add(a, h.dir[k]);
            org[k] = h.org[k];
        }
        t_min = h.t_min;
        t_max = h.cam_t_max;
        int tx = px / h.tile_w;
        if (tx >= h.tls_row) tx = h.tls_row - 1;
        head = v.tiles[(py / h.tile_h) * h.tls_row + tx];
        mode = QR_MODE_CLOSEST;
        p_obj = QR_NIL;
        p_flg = (int)h.ctx_flags;
    }

    float primary_t = t_max;
    int resume;                     /* 0 none, 1 after refraction, 2 after reflection */

    for (;;)
    {
        /* ---------------- WALK ---------------- */
        const float *ploc = mode == QR_MODE_SHADOW ? loc
                          : (lvl > 0 ? stack[lvl - 1].loc : loc);
        const bool res = qr_walk(v, head, mode, org, ray, t_min, t_max,
                                 p_obj, p_flg, ploc,
                                 t_buf, best_ei, best_side, best_loc);
        bool lights_phase = false;
        resume = 0;

        if (mode == QR_MODE_SHADOW)
        {
            /* LT_ret 2833-3151: light contribution unless occluded */
            const qr_elem le = v.elems[li];
            if (!res)
            {
                const qr_light &lg = v.lgts[le.simd];
                const qr_surface &s = v.surfs[v.elems[cur_ei].simd];
                const qr_material &m = v.mats[s.mat[l_flg & 1]];
                const int props = l_flg;

                /* ray[] holds NEW_X/Y/Z = light vector */
                float x4 = qr_mul(ray[0], ray[0]);
                x4 = qr_add(x4, qr_mul(ray[1], ray[1]));
                x4 = qr_add(x4, qr_mul(ray[2], ray[2]));
                const float r2 = x4;
                float d, x6;
                if (props & QR_PROP_DIFFUSE)
                {
                    d = dot;
                    x6 = x4;
                    const float x5 = qr_rsq(x4);
                    x4 = qr_mul(x5, x6);
                    x6 = qr_mul(x6, lg.a_qdr);
                    x4 = qr_mul(x4, lg.a_lnr);
                    x6 = qr_add(x6, lg.a_cnt);
                    x6 = qr_add(x6, x4);
                    x4 = qr_rsq(x6);
                    x6 = d;
                    d = qr_mul(d, x4);
                    d = qr_mul(d, x5);
                    d = qr_mul(d, m.l_dff);
                }
                else
                {
                    x6 = dot;
                    d = 0.0f;
                }

                bool spec_done = false;
                float spc = 0.0f;
                if (props & QR_PROP_SPECULAR)
                {
                    float x1 = ray[0], x2 = ray[1], x3 = ray[2];
                    float a4 = qr_mul(x6, nrm[0]);
                    x1 = qr_sub(x1, a4); x1 = qr_sub(x1, a4);
                    float a5 = qr_mul(x6, nrm[1]);
                    x2 = qr_sub(x2, a5); x2 = qr_sub(x2, a5);
                    float a6 = qr_mul(x6, nrm[2]);
                    x3 = qr_sub(x3, a6); x3 = qr_sub(x3, a6);
                    a4 = lray[0]; x1 = qr_mul(x1, a4); a4 = qr_mul(a4, a4);
                    a5 = lray[1]; x2 = qr_mul(x2, a5); a5 = qr_mul(a5, a5);
                    a6 = lray[2]; x3 = qr_mul(x3, a6); a6 = qr_mul(a6, a6);
                    a6 = qr_add(a6, a4);
                    a6 = qr_add(a6, a5);
                    x1 = qr_add(x1, x2);
                    x1 = qr_add(x1, x3);
                    if (0.0f < x1)
                    {
                        spec_done = true;
                        x1 = qr_mul(x1, qr_rsq(a6));
                        x1 = qr_mul(x1, qr_rsq(r2));
                        x1 = qr_pow_28_4(x1, m.l_pow);
                        spc = qr_mul(x1, m.l_spc);
                    }
                }

                if (spec_done && !(props & QR_PROP_METAL))
                {
                    /* LT_mtl 3090-3149 */
                    for (int k = 0; k < 3; k++)
                    {
                        float x1 = qr_mul(tex[k], d);
                        x1 = qr_mul(x1, lg.col[k]);
                        x1 = qr_add(x1, qr_mul(lg.col[k], spc));
                        col[k] = qr_add(x1, col[k]);
                    }
                }
                else
                {
                    /* LT_spc 3047-3084 */
                    if (spec_done) d = qr_add(d, spc);
                    for (int k = 0; k < 3; k++)
                    {
                        float x1 = qr_mul(tex[k], lg.col[k]);
                        x1 = qr_mul(x1, d);
                        col[k] = qr_add(x1, col[k]);
                    }
                }
            }
            li = le.next;
            lights_phase = true;
        }
        else
        {
            if (lvl == 0) primary_t = t_buf;

            if (!res)
            {
                /* nothing hit: COL of this level stays 0 */
                col[0] = col[1] = col[2] = 0.0f;
                resume = -1;
            }
            else
            {
                /* ---------------- SHADE ---------------- */
                cur_ei = best_ei;
                const qr_elem e = v.elems[cur_ei];
                const qr_surface &s = v.surfs[e.simd];
                const int side = best_side;
                l_flg = side | s.props[side];           /* FETCH_PROP */
                const int props = l_flg;
                const int shift = s.a_sgn[3];
                const qr_material &m = v.mats[s.mat[side]];

                lray[0] = ray[0]; lray[1] = ray[1]; lray[2] = ray[2];
                hit[0] = qr_add(qr_mul(ray[0], t_buf), org[0]);
                hit[1] = qr_add(qr_mul(ray[1], t_buf), org[1]);
                hit[2] = qr_add(qr_mul(ray[2], t_buf), org[2]);
                loc[0] = best_loc[0]; loc[1] = best_loc[1]; loc[2] = best_loc[2];

                const int kind = s.srf_t[0] == 1 ? 1 : s.srf_t[1];
                float tex_uv[2] = {0.0f, 0.0f};
                float nl[3] = {0.0f, 0.0f, 0.0f};       /* normal, local fields */

                if (kind == 1)
                {
                    /* PL_mat 4149-4193 */
                    if (props & QR_PROP_TEXTURE)
                    {
                        tex_uv[0] = qr_sgn(loc[s.a_map[0] - shift], s.a_sgn[0]);
                        tex_uv[1] = qr_sgn(loc[s.a_map[1] - shift], s.a_sgn[1]);
                    }
                    if (props & QR_PROP_NORMAL)
                    {
                        const uint32_t u = (0x3F800000u ^ (side ? 0x80000000u : 0u))
                                         ^ (s.a_sgn[2] ? 0x80000000u : 0u);
                        nl[s.a_map[0] - shift] = 0.0f;
                        nl[s.a_map[1] - shift] = 0.0f;
                        nl[s.a_map[2] - shift] = qr_u2f(u);
                    }
                }
                else
                if (props & QR_PROP_NORMAL)
                {
                    /* QD_mat 4855-4899 / TP_mat 4290-4330 */
                    float x4 = qr_mul(loc[0], s.sci[0]);
                    float x5 = qr_mul(loc[1], s.sci[1]);
                    float x6 = qr_mul(loc[2], s.sci[2]);
                    if (kind == 2)
                    {
                        x4 = qr_sub(x4, s.scj[0]);
                        x5 = qr_sub(x5, s.scj[1]);
                        x6 = qr_sub(x6, s.scj[2]);
                    }
                    float x1 = qr_mul(x4, x4);
                    x1 = qr_add(x1, qr_mul(x5, x5));
                    x1 = qr_add(x1, qr_mul(x6, x6));
                    float x0 = qr_rsq(x1);
                    if (side) x0 = qr_neg(x0);
                    nl[0] = qr_mul(x4, x0);
                    nl[1] = qr_mul(x5, x0);
                    nl[2] = qr_mul(x6, x0);
                }

                if (props & QR_PROP_NORMAL)
                {
                    if (s.a_map[3] != 0)
                    {
                        /* MT_nrm 2184-2259: transposed matrix of the trnode */
                        const qr_surface &t = v.surfs[s.trnode];
                        float x4 = qr_mul(t.tci[0], nl[0]);
                        float x5 = qr_mul(t.tcj[1], nl[1]);
                        float x6 = qr_mul(t.tck[2], nl[2]);
                        bool renorm = true;
                        if (t.a_map[3] != 1)
                        {
                            x4 = qr_add(x4, qr_mul(t.tcj[0], nl[1]));
                            x4 = qr_add(x4, qr_mul(t.tck[0], nl[2]));
                            x5 = qr_add(x5, qr_mul(t.tci[1], nl[0]));
                            x5 = qr_add(x5, qr_mul(t.tck[1], nl[2]));
                            x6 = qr_add(x6, qr_mul(t.tci[2], nl[0]));
                            x6 = qr_add(x6, qr_mul(t.tcj[2], nl[1]));
                            if (t.a_map[3] == 2) renorm = false;
                        }
                        if (renorm)
                        {
                            float x1 = qr_mul(x4, x4);
                            x1 = qr_add(x1, qr_mul(x5, x5));
                            x1 = qr_add(x1, qr_mul(x6, x6));
                            const float x0 = qr_rsq(x1);
                            x4 = qr_mul(x4, x0); x5 = qr_mul(x5, x0); x6 = qr_mul(x6, x0);
                        }
                        nrm[0] = x4; nrm[1] = x5; nrm[2] = x6;
                    }
                    else
                    {
                        nrm[0] = nl[0]; nrm[1] = nl[1]; nrm[2] = nl[2];
                    }
                }

                /* MT_mat 2286-2327: texel */
                uint32_t p = 0;
                if (props & QR_PROP_TEXTURE)
                {
                    float tx = tex_uv[m.t_map[0]];
                    float ty = tex_uv[m.t_map[1]];
                    tx = qr_sub(tx, m.xoffs);
                    ty = qr_sub(ty, m.yoffs);
                    tx = qr_mul(tx, m.xscal);
                    ty = qr_mul(ty, m.yscal);
                    const uint32_t ix = (uint32_t)qr_cvm(tx) & m.xmask;
                    const uint32_t iy = ((uint32_t)qr_cvm(ty) & m.ymask) << m.yshft;
                    p = ix + iy;
                }
                const uint32_t texel = v.texels[m.tex + p];
                tex[0] = qr_unpack(texel, 16, m, props);
                tex[1] = qr_unpack(texel, 8, m, props);
                tex[2] = qr_unpack(texel, 0, m, props);

                if (props & QR_PROP_LIGHT)
                {
                    /* LT_set 3164-3177 */
                    col[0] = tex[0]; col[1] = tex[1]; col[2] = tex[2];
                    li = QR_NIL;
                }
                else
                {
                    /* ambient 2721-2756 */
                    col[0] = qr_mul(tex[0], h.amb[0]);
                    col[1] = qr_mul(tex[1], h.amb[1]);
                    col[2] = qr_mul(tex[2], h.amb[2]);
                    li = s.lst_lgt[side];
                }
                lights_phase = true;
            }
        }

        if (lights_phase)
        {
            /* LT_cyc 2762-2831: next light that sees the front of the surface */
            bool go_shadow = false;
            while (li != QR_NIL)
            {
                const qr_elem le = v.elems[li];
                const qr_light &lg = v.lgts[le.simd];
                float x1 = qr_sub(lg.pos[0], hit[0]);
                float x2 = qr_sub(lg.pos[1], hit[1]);
                float x3 = qr_sub(lg.pos[2], hit[2]);
                float d = qr_mul(x1, nrm[0]);
                d = qr_add(d, qr_mul(x2, nrm[1]));
                d = qr_add(d, qr_mul(x3, nrm[2]));
                if (0.0f < d)
                {
                    dot = d;
                    org[0] = hit[0]; org[1] = hit[1]; org[2] = hit[2];
                    ray[0] = x1; ray[1] = x2; ray[2] = x3;
                    t_min = 0.0f;
                    t_max = lg.t_max;
                    head = le.data_p;
                    mode = QR_MODE_SHADOW;
                    p_obj = v.elems[cur_ei].simd;
                    p_flg = l_flg | QR_FLAG_PASS_BACK | QR_FLAG_SHAD;
                    if (cnt) cnt->shadow++;
                    go_shadow = true;
                    break;
                }
                li = le.next;
            }
            if (go_shadow) continue;
            resume = 0;
        }

        /* ------------- TRANSPARENCY / REFLECTION / unwinding ------------- */
        for (;;)
        {
            const qr_surface *sp = 0;
            const qr_material *mp = 0;
            int props = 0, side = 0;

            if (resume == -1)
            {
                /* return colour "col" of the finished level to its parent */
                if (lvl == 0) goto done;
                lvl--;
                const qr_frame &f = stack[lvl];
                const float cc[3] = { col[0], col[1], col[2] };
                col[0] = f.col[0]; col[1] = f.col[1]; col[2] = f.col[2];
                lray[0] = f.ray[0]; lray[1] = f.ray[1]; lray[2] = f.ray[2];
                hit[0] = f.hit[0]; hit[1] = f.hit[1]; hit[2] = f.hit[2];
                nrm[0] = f.nrm[0]; nrm[1] = f.nrm[1]; nrm[2] = f.nrm[2];
                loc[0] = f.loc[0]; loc[1] = f.loc[1]; loc[2] = f.loc[2];
                c_trn = f.c_trn; c_rfl = f.c_rfl;
                cur_ei = f.ei; l_flg = f.flg;
                if (f.stage == 0)
                {
                    /* TR_ret 3534-3552 */
                    xr[0] = qr_mul(cc[0], c_trn);
                    xr[1] = qr_mul(cc[1], c_trn);
                    xr[2] = qr_mul(cc[2], c_trn);
                    resume = 1;
                }
                else
                {
                    /* RF_ret 3868-3884 */
                    xr[0] = qr_mul(cc[0], c_rfl);
                    xr[1] = qr_mul(cc[1], c_rfl);
                    xr[2] = qr_mul(cc[2], c_rfl);
                    resume = 2;
                }
            }

            sp = &v.surfs[v.elems[cur_ei].simd];
            side = l_flg & 1;
            props = l_flg;
            mp = &v.mats[sp->mat[side]];
            const qr_surface &s = *sp;
            const qr_material &m = *mp;

            if (resume == 0)
            {
                /* TRANSPARENCY 3185-3532 */
                c_trn = m.c_trn;
                c_rfl = m.c_rfl;
                xr[0] = xr[1] = xr[2] = 0.0f;
                bool push = false;
                float nw[3] = {0.0f, 0.0f, 0.0f};

                if (!(props & QR_PROP_OPAQUE))
                {
                    bool go = true;
                    float x0 = 0.0f, x4 = 0.0f, x7 = 0.0f;
                    if ((props & QR_PROP_REFRACT) || (props & QR_PROP_FRESNEL))
                    {
                        float a[3];
                        x4 = qr_norm_dot(lray, nrm, a);
                        x0 = qr_mul(x4, m.c_rfr);
                        x7 = qr_mul(x0, x0);
                        x7 = qr_add(x7, 1.0f);
                        x7 = qr_sub(x7, m.rfr_2);
                        if ((props & QR_PROP_FRESNEL) && !(0.0f <= x7))
                        {
                            /* TR_tir 3280-3295 */
                            c_trn = 0.0f;
                            c_rfl = qr_add(m.c_rfl, m.c_trn);
                            go = false;
                        }
                        if (go)
                        {
                            x7 = qr_sqrt(x7);
                            x0 = qr_add(x0, x7);
                            if (props & QR_PROP_REFRACT)
                            {
                                nw[0] = qr_sub(qr_mul(a[0], m.c_rfr), qr_mul(nrm[0], x0));
                                nw[1] = qr_sub(qr_mul(a[1], m.c_rfr), qr_mul(nrm[1], x0));
                                nw[2] = qr_sub(qr_mul(a[2], m.c_rfr), qr_mul(nrm[2], x0));
                            }
                            else
                            {
                                nw[0] = lray[0]; nw[1] = lray[1]; nw[2] = lray[2];
                            }
                        }
                    }
                    else
                    {
                        nw[0] = lray[0]; nw[1] = lray[1]; nw[2] = lray[2];
                    }
                    if (go && (props & QR_PROP_FRESNEL))
                    {
                        /* TR_ini 3385-3424 */
                        float a0 = qr_fresnel(x4, m.c_rfr, x0, x7);
                        a0 = qr_mul(a0, m.c_trn);
                        c_trn = qr_sub(m.c_trn, a0);
                        c_rfl = qr_add(m.c_rfl, a0);
                    }
                    if (go && lvl < h.depth)
                    {
                        push = true;
                    }
                }

                if (push)
                {
                    qr_frame &f = stack[lvl];
                    f.col[0] = col[0]; f.col[1] = col[1]; f.col[2] = col[2];
                    f.ray[0] = lray[0]; f.ray[1] = lray[1]; f.ray[2] = lray[2];
                    f.hit[0] = hit[0]; f.hit[1] = hit[1]; f.hit[2] = hit[2];
                    f.nrm[0] = nrm[0]; f.nrm[1] = nrm[1]; f.nrm[2] = nrm[2];
                    f.loc[0] = loc[0]; f.loc[1] = loc[1]; f.loc[2] = loc[2];
                    f.c_trn = c_trn; f.c_rfl = c_rfl;
                    f.ei = cur_ei; f.flg = l_flg; f.stage = 0;
                    org[0] = hit[0]; org[1] = hit[1]; org[2] = hit[2];
                    ray[0] = nw[0]; ray[1] = nw[1]; ray[2] = nw[2];
                    t_min = 0.0f;
                    t_max = h.cam_t_max;
                    head = s.lst_srf[side ^ 1];         /* FETCH_IPTR */
                    mode = QR_MODE_CLOSEST;
                    p_obj = v.elems[cur_ei].simd;
                    p_flg = l_flg | QR_FLAG_PASS_THRU;
                    lvl++;
                    if (cnt) cnt->refract++;
                    break;
                }
                resume = 1;
            }

            if (resume == 1)
            {
                /* TR_mix 3564-3598 */
                float x0 = qr_sub(1.0f, m.c_trn);
                x0 = qr_sub(x0, m.c_rfl);
                if (!(0.0f <= x0)) x0 = 0.0f;
                col[0] = qr_add(xr[0], qr_mul(col[0], x0));
                col[1] = qr_add(xr[1], qr_mul(col[1], x0));
                col[2] = qr_add(xr[2], qr_mul(col[2], x0));

                /* REFLECTIONS 3604-3866 */
                bool go = (props & QR_PROP_REFLECT) != 0;
                if (!go && !(props & QR_PROP_OPAQUE) && (props & QR_PROP_FRESNEL)) go = true;
                if (!go)
                {
                    resume = -1;
                    continue;
                }

                float a[3], nw[3];
                const float d = qr_norm_dot(lray, nrm, a);
                for (int k = 0; k < 3; k++)
                {
                    const float nd = qr_mul(nrm[k], d);
                    nw[k] = qr_sub(qr_sub(a[k], nd), nd);
                }

                if ((props & QR_PROP_FRESNEL) && (props & QR_PROP_OPAQUE))
                {
                    float a0 = d;
                    if (props & QR_PROP_METAL)
                    {
                        /* 3729-3751 */
                        float a6 = m.c_rcp;
                        float a4 = qr_mul(a0, a6);
                        a4 = qr_add(a4, a4);
                        a0 = qr_mul(a0, a0);
                        a6 = qr_mul(a6, a6);
                        a6 = qr_add(a6, m.ext_2);
                        float a1 = qr_mul(a0, a6);
                        a0 = qr_add(a0, a6);
                        a1 = qr_add(a1, 1.0f);
                        const float a2 = qr_sub(a0, a4), a3 = qr_sub(a1, a4);
                        a0 = qr_add(a0, a4);
                        a1 = qr_add(a1, a4);
                        a0 = qr_div(a0, a2);
                        a1 = qr_div(a1, a3);
                        a0 = qr_add(a0, a1);
                        a0 = qr_abs(qr_mul(a0, -0.5f));
                    }
                    else
                    {
                        /* RF_mtl 3767-3796 */
                        float y0 = qr_mul(a0, m.c_rfr);
                        float y7 = qr_mul(y0, y0);
                        y7 = qr_add(y7, 1.0f);
                        y7 = qr_sub(y7, m.rfr_2);
                        y7 = qr_sqrt(y7);
                        y0 = qr_add(y0, y7);
                        a0 = qr_fresnel(a0, m.c_rfr, y0, y7);
                    }
                    /* RF_pre 3806-3815 */
                    a0 = qr_sub(a0, 1.0f);
                    a0 = qr_mul(a0, m.c_rfl);
                    c_rfl = qr_add(m.c_rfl, a0);
                }

                xr[0] = xr[1] = xr[2] = 0.0f;
                if (lvl < h.depth)
                {
                    qr_frame &f = stack[lvl];
                    f.col[0] = col[0]; f.col[1] = col[1]; f.col[2] = col[2];
                    f.ray[0] = lray[0]; f.ray[1] = lray[1]; f.ray[2] = lray[2];
                    f.hit[0] = hit[0]; f.hit[1] = hit[1]; f.hit[2] = hit[2];
                    f.nrm[0] = nrm[0]; f.nrm[1] = nrm[1]; f.nrm[2] = nrm[2];
                    f.loc[0] = loc[0]; f.loc[1] = loc[1]; f.loc[2] = loc[2];
                    f.c_trn = c_trn; f.c_rfl = c_rfl;
                    f.ei = cur_ei; f.flg = l_flg; f.stage = 1;
                    org[0] = hit[0]; org[1] = hit[1]; org[2] = hit[2];
                    ray[0] = nw[0]; ray[1] = nw[1]; ray[2] = nw[2];
                    t_min = 0.0f;
                    t_max = h.cam_t_max;
                    head = s.lst_srf[side];             /* FETCH_XPTR */
                    mode = QR_MODE_CLOSEST;
                    p_obj = v.elems[cur_ei].simd;
                    p_flg = l_flg | QR_FLAG_PASS_BACK;
                    lvl++;
                    if (cnt) cnt->reflect++;
                    break;
                }
                resume = 2;
            }

            if (resume == 2)
            {
                /* RF_mix 3888-3908 */
                col[0] = qr_add(xr[0], col[0]);
                col[1] = qr_add(xr[1], col[1]);
                col[2] = qr_add(xr[2], col[2]);
                resume = -1;
                continue;
            }
        }
    }

done:
    out_col[0] = col[0];
    out_col[1] = col[1];
    out_col[2] = col[2];
    *out_t = primary_t;
}

/* ---- epilogue helpers, XX_end 5221-5343 / FRAME_SIMD 988-1006 -------------- */

QR_HD float qr_clamp1(float c)                  /* minps: source on NaN */
{
    return c < 1.0f ? c : 1.0f;
}

QR_HD uint32_t qr_pack(const qr_blob_header &h, float r, float g, float b)
{
    float c[3] = { r, g, b };
    uint32_t pix = 0;
    for (int k = 0; k < 3; k++)
    {
        float x = c[k];
        if (h.ctx_flags & QR_PROP_GAMMA) x = qr_sqrt(x);
        x = qr_mul(x, h.cam_clamp);
        const uint32_t iv = (uint32_t)qr_cvn(x) & h.cam_cmask;
        pix |= iv << (k == 0 ? 16 : k == 1 ? 8 : 0);
    }
    return pix;
}

#endif /* QR_CORE_CUH */
