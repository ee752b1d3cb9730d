/*
 * qr_pt.cuh -- the path-tracer branch of render0 as a
 * PACKET tracer: one warp is one 32-lane packet of the reference's 512x2v2
 * target, one thread one SIMD lane.
 *
 * Why not the per-sample kernel of qr_core.cuh: with pt_on the reference's
 * result depends on its SIMD width.  Every lane owns a random-number seed
 * (rt_Scene::pseed, engine.cpp:2875-2901) that advances only while the lane is
 * active, and render0 shades a surface the moment a packet finds it closer
 * than what it had (tracer.cpp:2339-2701: a diffuse bounce per tentative hit,
 * Russian roulette below the fifth level, one of reflection / refraction below
 * the second, 3428-3466) -- so the seeds advance per tentative hit, and the
 * packet-wide early-outs (CHECK_MASK) decide whether code with unmasked side
 * effects runs for a lane at all (a lane in total inner reflection takes part
 * in the Fresnel split only if another lane of its packet is not).  A kernel
 * in which every sample decides alone renders a few different pixels per frame
 * (oracle at packet 1 vs 32); bit parity needs the packet.  So this file keeps
 * the macro-assembler's structure: the context stack as an array of fields per
 * lane (a context is a window of 64 fields, the next one starts 54 further, so
 * T_NEW / HIT / NEW of a level ARE T_MIN / ORG / RAY of the next; tracer.h:
 * 426-665), masks instead of branches, and warp votes where the reference
 * tests a whole register (CHECK_MASK NONE / FULL, tracer.cpp:454-470).
 * Control flow is uniform across the warp by construction, which is what lets
 * the mutually recursive walk / material / clip functions vote with the full
 * mask.  Every floating-point operation is the reference's, in its order,
 * separately rounded (the library is built with --fmad=false, IEEE division
 * and square root); fmaf is the reference's VFMADD231PS.
 *
 * It reads the scene BLOB as the flattener wrote it (include/qr_scene_blob.h),
 * not the compiled image of qr_kscene.h: the path tracer is a convergence
 * mode, its cost is the bounces.
 *
 * Host build (tests/hostsim): a "warp" of one lane -- every sample decides
 * alone, which is what the oracle computes at packet = 1.
 */
#ifndef QR_PT_CUH
#define QR_PT_CUH

#include <stdint.h>
#include <math.h>
#include "qr_scene_blob.h"

#if defined(__CUDACC__)
#define QR_PT_D   __device__ __forceinline__
#define QR_PT_REC __device__ __noinline__
#else
#define QR_PT_D   static inline
#define QR_PT_REC static
#endif

namespace qr_pt
{

#define NF     64                       /* fields per context */
#define STEP   54                       /* RT_STACK_STEP / (Q*0x10) */
#define LEVELS (QR_STACK_DEPTH + 3)
#define NIL    QR_NIL
#define SMASK  0x80000000u
#define ONES   0xFFFFFFFFu
/* one pass: "continue" leaves the block, as it left the lane loop it was */
#define QR_LANE for (int qr_lane_once = 0; qr_lane_once < 1; qr_lane_once++)

typedef union { float f; uint32_t u; int32_t i; } W;

/* field numbers = tracer.h:426-662 offsets / (Q*0x10) */
enum
{
    T_MIN = 0, ORG = 1, RAY = 4, DFF = 10, TEX_U = 16, TEX_V = 17,
    C_PTR = 18, C_BUF = 19, TEX = 20, COL = 23, C_ACC = 26, F_RFL = 27,
    T_VAL = 28, T_BUF = 29, TMASK = 30, WMASK = 31, XMASK = 32,
    XTMP1 = 33, XTMP2 = 34, NRM = 36, AMASK = 46, DMASK = 47,
    F_RND = 48, F_PRB = 49, M_TRN = 50, M_RFL = 51, C_TRN = 52, C_RFL = 53,
    T_NEW = 54, HIT = 55, NEW = 58
};

/* packed scalar fields of one context: PARAM, LOCAL, XMISC (tracer.h:561-592) */
struct lvl_t
{
    int p_flg, p_lst, p_obj;            /* PARAM(FLG/LST/OBJ) */
    int l_flg, l_lst, l_obj;            /* LOCAL(FLG/LST/OBJ) */
    int x_ptr, x_flg, x_tag;            /* XMISC(PTR/FLG/TAG) */
};

/* one lane's view of the packet's state (thread-local) */
struct R
{
    const qr_blob_header *h;
    const qr_surface  *surfs;
    const qr_material *mats;
    const qr_light    *lgts;
    const qr_elem     *elems;
    const int32_t     *tiles;
    const uint32_t    *texels;
    int   depth;                        /* inf_DEPTH */
    int   pt;                           /* pt_on (always set here; the non-PT code paths stay for reference) */
    uint32_t *seed;                     /* this lane's slot of the seed plane */
    lvl_t lv[LEVELS];
    W     mem[LEVELS * STEP + NF];
};

#define FLD(c, f) (r->mem[(c) + (f)])

/* ---- instruction semantics (core/config/rtarch_x32_512x2v2.h:706-880) ------ */

QR_PT_D uint32_t m_lt(float a, float b) { return a <  b ? ONES : 0; } /* clt */
QR_PT_D uint32_t m_le(float a, float b) { return a <= b ? ONES : 0; } /* cle */
QR_PT_D uint32_t m_gt(float a, float b) { return !(a <= b) ? ONES : 0; } /* cgt = NLE */
QR_PT_D uint32_t m_ge(float a, float b) { return !(a <  b) ? ONES : 0; } /* cge = NLT */
QR_PT_D uint32_t m_eq(float a, float b) { return a == b ? ONES : 0; } /* ceq */
QR_PT_D uint32_t m_ne(float a, float b) { return a != b ? ONES : 0; } /* cne = NEQ_UQ */

QR_PT_D float    u2f(uint32_t u) { W w; w.u = u; return w.f; }
QR_PT_D uint32_t f2u(float f)    { W w; w.f = f; return w.u; }

QR_PT_D float rsq(float x) { return 1.0f / sqrtf(x); }

/* sinps_rr / cosps_rr, tracer.cpp:1032-1057 */
QR_PT_D float sin_ps(float x)
{
    const float t1 = x * x;
    float xd = x;
    float xs = xd * t1;
    xd = fmaf(xs, -0.1666666666666666666666666666666666666666666f, xd);
    xs = xs * t1;
    xd = fmaf(xs, +0.0083333333333333333333333333333333333333333f, xd);
    xs = xs * t1;
    xd = fmaf(xs, -0.0001984126984126984126984126984126984126984f, xd);
    xs = xs * t1;
    xd = fmaf(xs, +0.0000027557319223985890652557319223985890652f, xd);
    return xd;
}

QR_PT_D float cos_ps(float x)
{
    const float t1 = x * x;
    float xd = 1.0f;
    float xs = xd * t1;
    xd = fmaf(xs, -0.5f, xd);
    xs = xs * t1;
    xd = fmaf(xs, +0.0416666666666666666666666666666666666666666f, xd);
    xs = xs * t1;
    xd = fmaf(xs, -0.0013888888888888888888888888888888888888888f, xd);
    xs = xs * t1;
    xd = fmaf(xs, +0.0000248015873015873015873015873015873015873f, xd);
    return xd;
}

/* cvnps: round to nearest even, x86 "integer indefinite" out of range */
QR_PT_D int32_t cvn(float x)
{
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return (int32_t)0x80000000u;
    return (int32_t)rintf(x);
}

/* cvmps: round towards minus infinity */
QR_PT_D int32_t cvm(float x)
{
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return (int32_t)0x80000000u;
    return (int32_t)floorf(x);
}

/* CHECK_MASK NONE / FULL over the packet (tracer.cpp:454-470) */
#if defined(__CUDA_ARCH__)
QR_PT_D int pk_none(const W m) { return !__any_sync(0xFFFFFFFFu, m.u != 0u); }
QR_PT_D int pk_full(const W m) { return __all_sync(0xFFFFFFFFu, m.u == ONES); }
#else
QR_PT_D int pk_none(const W m) { return m.u == 0u; }
QR_PT_D int pk_full(const W m) { return m.u == ONES; }
#endif

/*
 * GET_RANDOM 1014-1028, RT_PRNG = LCG24 (tracer.h:53, engine.cpp:867-873): the
 * lane's seed advances s = s * 214013 + 2531011 (mod 2^32) and is stored only
 * if the lane is in TMASK; the number is bits 8..31 of the new seed over 2^24
 * (a lane outside TMASK gets a number too, from a seed that stays).
 */
QR_PT_D W get_random(R *r, int c)
{
    uint32_t sd = *r->seed;
    sd = sd * 214013u + 2531011u;
    if (FLD(c, TMASK).u) *r->seed = sd;
    W x;
    x.f = (float)(int32_t)((sd >> 8) & 0xFFFFFFu) / 16777216.0f;
    return x;
}

/*
 * 3x3 transform as written at tracer.cpp:1447-1479 (diff), 1512-1548 (ray),
 * 2063-2095 (clip): diagonal products first, then the two off-diagonal terms
 * of each row in column order; a_map[L] == 1 keeps the diagonal only.
 */
QR_PT_D void xform(const qr_surface *s, float v1, float v2, float v3,
                         float *o4, float *o5, float *o6)
{
    float x4 = s->tci[0] * v1;
    float x5 = s->tcj[1] * v2;
    float x6 = s->tck[2] * v3;
    if (s->a_map[3] != 1)
    {
        x4 = x4 + s->tci[1] * v2;
        x4 = x4 + s->tci[2] * v3;
        x5 = x5 + s->tcj[0] * v1;
        x5 = x5 + s->tcj[2] * v3;
        x6 = x6 + s->tck[0] * v1;
        x6 = x6 + s->tck[1] * v2;
    }
    *o4 = x4; *o5 = x5; *o6 = x6;
}

QR_PT_REC void walk(R *r, int lvl, int ei);

/* ---- custom clipping subroutine: tracer.cpp:1597-2160 (CC_clp .. CC_out) -- */
/* x7 in/out is the running tmask (Xmm7); returns with it updated */

QR_PT_REC void clip(R *r, int lvl, int ei, W &x7)
{
    const int c = lvl * STEP;
    lvl_t *L = &r->lv[lvl];
    const qr_elem *e = &r->elems[ei];
    const qr_surface *s = &r->surfs[e->simd];
    const int shift = s->a_sgn[3];
    W x4, x5, x6;

    /* 1599-1681: depth test, near plane, hit point, local hit */
    QR_LANE
    {
        float t = FLD(c, T_VAL).f;
        x7.u &= m_gt(FLD(c, T_BUF).f, t);
        x7.u &= m_lt(FLD(c, T_MIN).f, t);

        float hx = FLD(c, RAY + 0).f * t; hx = hx + FLD(c, ORG + 0).f;
        float hy = FLD(c, RAY + 1).f * t; hy = hy + FLD(c, ORG + 1).f;
        float hz = FLD(c, RAY + 2).f * t; hz = hz + FLD(c, ORG + 2).f;
        FLD(c, HIT + 0).f = hx;
        FLD(c, HIT + 1).f = hy;
        FLD(c, HIT + 2).f = hz;

        if (s->a_map[3] != 0)
        {
            float li = FLD(c, RAY + 3).f * t; li = li + FLD(c, DFF + 3).f;
            float lj = FLD(c, RAY + 4).f * t; lj = lj + FLD(c, DFF + 4).f;
            float lk = FLD(c, RAY + 5).f * t; lk = lk + FLD(c, DFF + 5).f;
            FLD(c, NEW + 3).f = li;
            FLD(c, NEW + 4).f = lj;
            FLD(c, NEW + 5).f = lk;
            x4.f = li; x5.f = lj; x6.f = lk;
        }
        else
        {
            hx = hx - s->pos[0];
            hy = hy - s->pos[1];
            hz = hz - s->pos[2];
            FLD(c, NEW + 0).f = hx;
            FLD(c, NEW + 1).f = hy;
            FLD(c, NEW + 2).f = hz;
            x4.f = hx; x5.f = hy; x6.f = hz;
        }
    }

    /* 1706-1856: conic singularity solver */
    if (s->conic != 0 && L->x_ptr != 0)
    {
        const int iI = s->a_map[0], iJ = s->a_map[1], iK = s->a_map[2];
        W x0;
        QR_LANE
        {
            float a1 = FLD(c, NEW + iI).f; a1 = a1 * a1;
            float a0 = a1;
            if (s->conic != 2)
            {
                float a2 = FLD(c, NEW + iJ).f; a2 = a2 * a2;
                a0 = a0 + a2;
            }
            float a3 = FLD(c, NEW + iK).f; a3 = a3 * a3;
            a0 = a0 + a3;
            x0.u = m_lt(a0, s->t_eps) & FLD(c, DMASK).u;
        }
        if (!pk_none(x0))
        {
            QR_LANE
            {
                uint32_t hm = x0.u;
                float one = 1.0f;
                uint32_t q1 = (FLD(c, DFF + iI).u & SMASK) ^ f2u(one);
                uint32_t q2 = 0;
                float q3 = s->sci[iI - shift];
                float q4 = one;
                if (s->conic != 2)
                {
                    q2 = (FLD(c, DFF + iJ).u & SMASK) ^ f2u(one);
                    q3 = q3 + s->sci[iJ - shift];
                    q4 = q4 + one;
                }
                q3 = q3 / s->sci[iK - shift];
                q3 = u2f(f2u(q3) ^ SMASK);
                float q6 = q3;
                q3 = sqrtf(q3);
                q6 = q6 + q4;
                q4 = rsq(q6);
                q4 = q4 * s->t_eps;
                float p1 = u2f(q1) * q4;
                float p2 = u2f(q2) * q4;
                float p3 = q3 * q4;

                uint32_t am = FLD(c, AMASK).u;
                uint32_t ts = (L->l_flg & 1) ? SMASK : 0;   /* srf_SBASE + FLG*Q*16 */
                uint32_t sk = FLD(c, DFF + iK).u & SMASK;
                uint32_t u3 = f2u(p3) ^ sk;
                u3 ^= (ts & am) ^ am;
                uint32_t tsn = (ts | am) ^ am;
                uint32_t u1 = f2u(p1) ^ tsn;
                uint32_t u2 = f2u(p2) ^ tsn;

                FLD(c, NEW + iI).u = ((FLD(c, NEW + iI).u | hm) ^ hm) | (u1 & hm);
                if (s->conic != 2)
                {
                    FLD(c, NEW + iJ).u = ((FLD(c, NEW + iJ).u | hm) ^ hm) | (u2 & hm);
                }
                FLD(c, NEW + iK).u = ((FLD(c, NEW + iK).u | hm) ^ hm) | (u3 & hm);
            }
            QR_LANE
            {
                x4 = FLD(c, NEW + shift + 0);
                x5 = FLD(c, NEW + shift + 1);
                x6 = FLD(c, NEW + shift + 2);
            }
        }
    }

    /* 1874-1927: axis min/max clipping on the (un-mapped) local point */
    QR_LANE
    {
        if (s->minmax_t & 1)  x7.u &= m_le(s->min[0], x4.f);
        if (s->minmax_t & 8)  x7.u &= m_ge(s->max[0], x4.f);
        if (s->minmax_t & 2)  x7.u &= m_le(s->min[1], x5.f);
        if (s->minmax_t & 16) x7.u &= m_ge(s->max[1], x5.f);
        if (s->minmax_t & 4)  x7.u &= m_le(s->min[2], x6.f);
        if (s->minmax_t & 32) x7.u &= m_ge(s->max[2], x6.f);
    }

    /* 1931-2151: custom clippers */
    L->l_lst = s->trnode;
    int redx = NIL;

    for (int di = s->clip_head; di != NIL; di = r->elems[di].next)
    {
        const qr_elem *ce = &r->elems[di];

        if (ce->simd == NIL)                        /* 1948-1962: accum marker */
        {
            if (ce->data_i > 0)
            {
                QR_LANE x7.u = ~x7.u & FLD(c, C_ACC).u;
            }
            else
            {
                QR_LANE
                {
                    FLD(c, C_ACC).u = x7.u;
                    x7.u = s->c_def;
                }
            }
            continue;
        }

        const qr_surface *cs = &r->surfs[ce->simd];
        int have_local = 0;

        if (cs->srf_t[3] >= 0)
        {
            if (redx != NIL)                        /* 1976-2004: cached trnode */
            {
                QR_LANE
                {
                    FLD(c, NRM + 3).f = FLD(c, NRM + 0).f - cs->pos[0];
                    FLD(c, NRM + 4).f = FLD(c, NRM + 1).f - cs->pos[1];
                    FLD(c, NRM + 5).f = FLD(c, NRM + 2).f - cs->pos[2];
                }
                if (di == redx) redx = NIL;
                have_local = 1;
            }
        }
        else
        if (ce->simd == L->l_lst)                   /* 2006-2037: same trnode */
        {
            QR_LANE
            {
                FLD(c, NRM + 0).f = FLD(c, NEW + 3).f + s->pos[0];
                FLD(c, NRM + 1).f = FLD(c, NEW + 4).f + s->pos[1];
                FLD(c, NRM + 2).f = FLD(c, NEW + 5).f + s->pos[2];
            }
            redx = ce->data_p;
            continue;
        }

        if (!have_local)                            /* 2039-2125: CC_dff */
        {
            int cached = 0;
            QR_LANE
            {
                float d1 = FLD(c, HIT + 0).f - cs->pos[0];
                float d2 = FLD(c, HIT + 1).f - cs->pos[1];
                float d3 = FLD(c, HIT + 2).f - cs->pos[2];
                FLD(c, NRM + 0).f = d1;
                FLD(c, NRM + 1).f = d2;
                FLD(c, NRM + 2).f = d3;
                if (cs->a_map[3] != 0)
                {
                    float o4, o5, o6;
                    xform(cs, d1, d2, d3, &o4, &o5, &o6);
                    if (cs->srf_t[3] < 0)
                    {
                        FLD(c, NRM + 0).f = o4;
                        FLD(c, NRM + 1).f = o5;
                        FLD(c, NRM + 2).f = o6;
                        cached = 1;
                    }
                    else
                    {
                        FLD(c, NRM + 3).f = o4;
                        FLD(c, NRM + 4).f = o5;
                        FLD(c, NRM + 5).f = o6;
                    }
                }
            }
            if (cached)
            {
                redx = ce->data_p;
                continue;
            }
        }

        /* CC_trm 2127-2140: clipper evaluators */
        const int cshift = cs->a_sgn[3];
        if (cs->srf_t[2] == 1)                      /* PL_clp 4198-4208 */
        {
            const int k = cs->a_map[2];
            const uint32_t sg = cs->a_sgn[2] ? SMASK : 0;
            QR_LANE
            {
                float v = u2f(FLD(c, NRM + k).u ^ sg);
                x4.u = ce->data_i < 0 ? m_ge(v, 0.0f) : m_le(v, 0.0f);
            }
        }
        else
        if (cs->srf_t[2] == 2)                      /* QD_clp 4910-4951 */
        {
            QR_LANE
            {
                float dx = FLD(c, NRM + cshift + 0).f;
                float dy = FLD(c, NRM + cshift + 1).f;
                float dz = FLD(c, NRM + cshift + 2).f;
                float a1 = cs->scj[0] + cs->scj[0]; a1 = a1 * dx;
                float a4 = dx * dx; a4 = a4 * cs->sci[0]; a4 = a4 - a1;
                float a2 = cs->scj[1] + cs->scj[1]; a2 = a2 * dy;
                float a5 = dy * dy; a5 = a5 * cs->sci[1]; a5 = a5 - a2;
                float a3 = cs->scj[2] + cs->scj[2]; a3 = a3 * dz;
                float a6 = dz * dz; a6 = a6 * cs->sci[2]; a6 = a6 - a3;
                a4 = a4 - cs->sci[3];
                a4 = a4 + a5;
                a4 = a4 + a6;
                x4.u = ce->data_i < 0 ? m_ge(a4, 0.0f) : m_le(a4, 0.0f);
            }
        }
        else
        if (cs->srf_t[2] == 3)                      /* TP_clp 4341-4370 */
        {
            QR_LANE
            {
                float dx = FLD(c, NRM + cshift + 0).f;
                float dy = FLD(c, NRM + cshift + 1).f;
                float dz = FLD(c, NRM + cshift + 2).f;
                float a4 = dx * dx; a4 = a4 * cs->sci[0];
                float a5 = dy * dy; a5 = a5 * cs->sci[1];
                float a6 = dz * dz; a6 = a6 * cs->sci[2];
                a4 = a4 - cs->sci[3];
                a4 = a4 + a5;
                a4 = a4 + a6;
                x4.u = ce->data_i < 0 ? m_ge(a4, 0.0f) : m_le(a4, 0.0f);
            }
        }
        /* CC_ret 2138-2140 (srf_t[2] == 0 would use a stale Xmm4) */
        QR_LANE x7.u &= x4.u;
    }
}

/* ---- material: tracer.cpp:4139-4193, 4280-4336, 4845-4905, 2166-3947 ------ */
/* kind: 1 PL_mat, 2 QD_mat, 3 TP_mat.  Returns 1 when the walk must stop
 * (OO_out from CHECK_SHAD), 0 to return to the calling solver (SR_rt*). */

QR_PT_REC int material(R *r, int lvl, int ei, int kind)
{
    const int c = lvl * STEP;
    lvl_t *L = &r->lv[lvl];
    const qr_elem *e = &r->elems[ei];
    const int si = e->simd;
    const qr_surface *s = &r->surfs[si];
    const int shift = s->a_sgn[3];

    /* FETCH_PROP 597-604 */
    const int side = L->l_flg & 1;
    const uint32_t tside = (L->l_flg & 1) ? SMASK : 0;
    L->l_flg |= s->props[side];
    const int props = L->l_flg;

    /* CHECK_SHAD 549-589 */
    if (L->p_flg & QR_FLAG_SHAD)
    {
        if (props & QR_PROP_LIGHT) return 0;
        if ((props & QR_PROP_TRANSP) && !(props & QR_PROP_REFRACT)) return 0;
        QR_LANE FLD(c, C_BUF).u |= FLD(c, TMASK).u;
        if (pk_full(FLD(c, C_BUF))) return 1;
        return 0;
    }

    int have_nrm = 0;

    if (kind == 1)
    {
        /* PL_mat 4149-4193 */
        if (props & QR_PROP_TEXTURE)
        {
            const uint32_t sgi = s->a_sgn[0] ? SMASK : 0, sgj = s->a_sgn[1] ? SMASK : 0;
            QR_LANE
            {
                FLD(c, TEX_U).u = FLD(c, NEW + s->a_map[0]).u ^ sgi;
                FLD(c, TEX_V).u = FLD(c, NEW + s->a_map[1]).u ^ sgj;
            }
        }
        if (props & QR_PROP_NORMAL)
        {
            const uint32_t sgk = s->a_sgn[2] ? SMASK : 0;
            QR_LANE
            {
                FLD(c, NRM + s->a_map[0]).u = 0;
                FLD(c, NRM + s->a_map[1]).u = 0;
                FLD(c, NRM + s->a_map[2]).u = (f2u(1.0f) ^ tside) ^ sgk;
            }
            have_nrm = 1;
        }
    }
    else
    {
        /* QD_mat 4855-4899 / TP_mat 4290-4330 */
        if (props & QR_PROP_NORMAL)
        {
            QR_LANE
            {
                float x4 = FLD(c, NEW + shift + 0).f * s->sci[0];
                float x5 = FLD(c, NEW + shift + 1).f * s->sci[1];
                float x6 = FLD(c, NEW + shift + 2).f * s->sci[2];
                if (kind == 2)
                {
                    x4 = x4 - s->scj[0];
                    x5 = x5 - s->scj[1];
                    x6 = x6 - s->scj[2];
                }
                float x1 = x4 * x4, x2 = x5 * x5, x3 = x6 * x6;
                x1 = x1 + x2;
                x1 = x1 + x3;
                float x0 = rsq(x1);
                x0 = u2f(f2u(x0) ^ tside);
                FLD(c, NRM + shift + 0).f = x4 * x0;
                FLD(c, NRM + shift + 1).f = x5 * x0;
                FLD(c, NRM + shift + 2).f = x6 * x0;
            }
            have_nrm = 1;
        }
    }

    /* MT_nrm 2184-2263: transform normal with the trnode's transposed matrix */
    if (have_nrm && s->a_map[3] != 0)
    {
        const qr_surface *t = &r->surfs[s->trnode];
        QR_LANE
        {
            float n1 = FLD(c, NRM + 3).f, n2 = FLD(c, NRM + 4).f, n3 = FLD(c, NRM + 5).f;
            float x4 = t->tci[0] * n1;
            float x5 = t->tcj[1] * n2;
            float x6 = t->tck[2] * n3;
            int renorm = 1;
            if (t->a_map[3] != 1)
            {
                x4 = x4 + t->tcj[0] * n2;
                x4 = x4 + t->tck[0] * n3;
                x5 = x5 + t->tci[1] * n1;
                x5 = x5 + t->tck[1] * n3;
                x6 = x6 + t->tci[2] * n1;
                x6 = x6 + t->tcj[2] * n2;
                if (t->a_map[3] == 2) renorm = 0;
            }
            if (renorm)
            {
                float x1 = x4 * x4, x2 = x5 * x5, x3 = x6 * x6;
                x1 = x1 + x2;
                x1 = x1 + x3;
                float x0 = rsq(x1);
                x4 = x4 * x0; x5 = x5 * x0; x6 = x6 * x0;
            }
            FLD(c, NRM + 0).f = x4;
            FLD(c, NRM + 1).f = x5;
            FLD(c, NRM + 2).f = x6;
        }
    }

    /* MT_mat 2267-2327 */
    QR_LANE
    {
        W a = FLD(c, NEW + shift + 0), b = FLD(c, NEW + shift + 1), d = FLD(c, NEW + shift + 2);
        FLD(c, NRM + 3) = a;
        FLD(c, NRM + 4) = b;
        FLD(c, NRM + 5) = d;
    }
    L->l_lst = ei;

    const qr_material *m = &r->mats[s->mat[side]];

    QR_LANE
    {
        uint32_t p = 0;
        if (props & QR_PROP_TEXTURE)
        {
            float tx = FLD(c, TEX_U + m->t_map[0]).f;
            float ty = FLD(c, TEX_U + m->t_map[1]).f;
            tx = tx - m->xoffs;
            ty = ty - m->yoffs;
            tx = tx * m->xscal;
            ty = ty * m->yscal;
            uint32_t ix = (uint32_t)cvm(tx) & m->xmask;
            uint32_t iy = ((uint32_t)cvm(ty) & m->ymask) << m->yshft;
            p = (ix + iy) << 2;
        }
        FLD(c, C_PTR).u = p;
    }
    /* PAINT_FRAG 653-662 */
    QR_LANE
    {
        if (FLD(c, TMASK).u == 0) continue;
        FLD(c, T_BUF) = FLD(c, T_VAL);
        FLD(c, C_BUF).u = r->texels[m->tex + (FLD(c, C_PTR).u >> 2)];
    }
    /* PAINT_COLX 664-673 */
    QR_LANE
    {
        uint32_t cb = FLD(c, C_BUF).u;
        for (int k = 0; k < 3; k++)
        {
            int sh = k == 0 ? 16 : k == 1 ? 8 : 0;
            float v = (float)(int32_t)((cb >> sh) & m->cmask);
            v = v / m->clamp;
            if (props & QR_PROP_GAMMA) v = v * v;
            FLD(c, TEX + k).f = v;
        }
    }

    /* LIGHTS 2333-3179 */
    QR_LANE
    {
        FLD(c, F_RND) = FLD(c, TMASK);
        FLD(c, F_PRB) = FLD(c, TMASK);
    }

    if (r->pt != 0)
    {
        /* path tracer, 2339-2701: instead of the lights one diffuse bounce */
        W x1, x2, x3;
        QR_LANE x1.u = x2.u = x3.u = 0;

        int bounce = (props & QR_PROP_DIFFUSE) != 0;        /* CHECK_PROP(PT_mix, RT_PROP_DIFFUSE) */
        if (bounce && !(r->depth > QR_STACK_DEPTH - 5))
        {
            /* 2352-2396 (RT_FEAT_PT_SPLIT_DEPTH): deeper levels go on with the
             * probability of the brightest colour channel */
            W x0, x4;
            QR_LANE
            {
                float a = FLD(c, TEX + 0).f;             /* maxps: the source when unordered */
                a = a > FLD(c, TEX + 1).f ? a : FLD(c, TEX + 1).f;
                a = a > FLD(c, TEX + 2).f ? a : FLD(c, TEX + 2).f;
                x4.f = a;
            }
            x0 = get_random(r, c);
            QR_LANE
            {
                x0.u = m_lt(x0.f, x4.f) & FLD(c, F_PRB).u;
                FLD(c, F_PRB) = x0;
                FLD(c, TMASK) = x0;
            }
            if (pk_none(x0))
            {
                bounce = 0;                                 /* PT_chk -> PT_mix */
            }
            else
            {
                QR_LANE
                {
                    const float x5 = 1.0f / x4.f;        /* rcpps, all lanes */
                    for (int k = 0; k < 3; k++) FLD(c, TEX + k).f = FLD(c, TEX + k).f * x5;
                }
            }
        }
        if (bounce)
        {
            /* 2398-2530: orthonormal basis around the normal (its fields are
             * borrowed ones: TEX_U, TEX_V, C_PTR, C_ACC, F_RFL, T_VAL) */
            W x6, x0;
            QR_LANE
            {
                const float n1 = FLD(c, NRM + 0).f, n2 = FLD(c, NRM + 1).f, n3 = FLD(c, NRM + 2).f;
                const float r4 = FLD(c, RAY + 0).f, r5 = FLD(c, RAY + 1).f, r6 = FLD(c, RAY + 2).f;
                float a0, a7;
                a0 = n2 * r6; a7 = n3 * r5; float u4 = a0 - a7;
                a0 = n3 * r4; a7 = n1 * r6; float u5 = a0 - a7;
                a0 = n1 * r5; a7 = n2 * r4; float u6 = a0 - a7;
                float s1 = u4 * u4, s2 = u5 * u5, s3 = u6 * u6;
                s1 = s1 + s2;
                s1 = s1 + s3;
                const float inv = rsq(s1);
                u4 = u4 * inv; u5 = u5 * inv; u6 = u6 * inv;
                FLD(c, TEX_U).f = u4; FLD(c, TEX_V).f = u5; FLD(c, C_PTR).f = u6;
                a0 = n2 * u6; a7 = n3 * u5; FLD(c, C_ACC).f = a0 - a7;
                a0 = n3 * u4; a7 = n1 * u6; FLD(c, F_RFL).f = a0 - a7;
                a0 = n1 * u5; a7 = n2 * u4; FLD(c, T_VAL).f = a0 - a7;
            }
            /* 2532-2590: cosine-weighted direction over the hemisphere */
            x0 = get_random(r, c);
            QR_LANE
            {
                x6.f = x0.f;
                float a0 = 1.0f - x6.f;
                x6.f = sqrtf(x6.f);
                a0 = sqrtf(a0);
                x1.f = FLD(c, NRM + 0).f * a0;
                x2.f = FLD(c, NRM + 1).f * a0;
                x3.f = FLD(c, NRM + 2).f * a0;
            }
            x0 = get_random(r, c);
            QR_LANE
            {
                const float pi = (float)3.14159265358979323846;     /* mat_GPC10, object.cpp:4130 */
                float a0 = x0.f + x0.f;
                a0 = a0 * pi;
                a0 = a0 - pi;
                float a4 = cos_ps(a0);
                a4 = a4 * x6.f;
                x1.f = x1.f + FLD(c, TEX_U).f * a4;
                x2.f = x2.f + FLD(c, TEX_V).f * a4;
                x3.f = x3.f + FLD(c, C_PTR).f * a4;
                a4 = sin_ps(a0);
                a4 = a4 * x6.f;
                x1.f = x1.f + FLD(c, C_ACC).f * a4;
                x2.f = x2.f + FLD(c, F_RFL).f * a4;
                x3.f = x3.f + FLD(c, T_VAL).f * a4;
                FLD(c, NEW + 0) = x1;
                FLD(c, NEW + 1) = x2;
                FLD(c, NEW + 2) = x3;
                FLD(c, T_NEW).u = 0;
                x1.u = x2.u = x3.u = 0;
            }
            if (r->depth != 0)
            {
                /* 2599-2660: the bounce, one level down (PT_ret is tag 4) */
                const int cc = c + STEP;
                lvl_t *C = &r->lv[lvl + 1];
                r->depth -= 1;
                C->p_flg = L->l_flg | QR_FLAG_PASS_BACK;
                C->p_lst = s->mat[side];
                C->p_obj = si;
                QR_LANE
                {
                    FLD(cc, WMASK) = FLD(c, TMASK);
                    FLD(cc, T_BUF).f = r->h->cam_t_max;
                    FLD(cc, C_BUF).u = 0;
                    FLD(cc, COL + 0).u = 0;
                    FLD(cc, COL + 1).u = 0;
                    FLD(cc, COL + 2).u = 0;
                    FLD(cc, T_MIN).u = 0;
                }
                C->l_flg = 0; C->l_lst = NIL; C->l_obj = NIL;
                walk(r, lvl + 1, s->lst_srf[side]);
                r->depth += 1;
                QR_LANE
                {
                    x1.f = FLD(cc, COL + 0).f * m->l_dff;
                    x2.f = FLD(cc, COL + 1).f * m->l_dff;
                    x3.f = FLD(cc, COL + 2).f * m->l_dff;
                    x1.f = x1.f * FLD(c, TEX + 0).f;
                    x2.f = x2.f * FLD(c, TEX + 1).f;
                    x3.f = x3.f * FLD(c, TEX + 2).f;
                }
            }
        }
        /* PT_mix 2664-2699: self-emission, then the radiance of the hit lanes */
        QR_LANE
        {
            x1.f = x1.f + m->col[0];
            x2.f = x2.f + m->col[1];
            x3.f = x3.f + m->col[2];
            FLD(c, TMASK) = FLD(c, F_RND);
            if (FLD(c, TMASK).u == 0) continue;
            FLD(c, COL + 0) = x1;
            FLD(c, COL + 1) = x2;
            FLD(c, COL + 2) = x3;
        }
    }
    else
    if (props & QR_PROP_LIGHT)
    {
        /* LT_set 3164-3177 */
        QR_LANE
        {
            if (FLD(c, TMASK).u == 0) continue;
            for (int k = 0; k < 3; k++) FLD(c, COL + k) = FLD(c, TEX + k);
        }
    }
    else
    {
        /* ambient 2721-2756 */
        QR_LANE
        {
            if (FLD(c, TMASK).u == 0) continue;
            for (int k = 0; k < 3; k++)
                FLD(c, COL + k).f = FLD(c, TEX + k).f * r->h->amb[k];
        }

        /* LT_cyc 2760-3156 */
        for (int li = s->lst_lgt[side]; li != NIL; li = r->elems[li].next)
        {
            const qr_elem *le = &r->elems[li];
            const qr_light *lg = &r->lgts[le->simd];
            W x0, x7;

            QR_LANE
            {
                float x1 = lg->pos[0] - FLD(c, HIT + 0).f;
                FLD(c, NEW + 0).f = x1;
                x1 = x1 * FLD(c, NRM + 0).f;
                float x2 = lg->pos[1] - FLD(c, HIT + 1).f;
                FLD(c, NEW + 1).f = x2;
                x2 = x2 * FLD(c, NRM + 1).f;
                float x3 = lg->pos[2] - FLD(c, HIT + 2).f;
                FLD(c, NEW + 2).f = x3;
                x3 = x3 * FLD(c, NRM + 2).f;
                float d = x1 + x2;
                d = d + x3;
                x0.f = d;
                x7.u = m_lt(0.0f, d) & FLD(c, TMASK).u;
            }
            if (pk_none(x7)) continue;

            /* shadows 2794-2850 */
            {
                const int cc = c + STEP;
                lvl_t *C = &r->lv[lvl + 1];
                QR_LANE
                {
                    x7.u = m_eq(x7.f, 0.0f);   /* ceqps with 0: inverted lmask */
                    FLD(c, C_PTR).f = x0.f;
                }
                r->depth -= 1;
                C->p_flg = L->l_flg | QR_FLAG_PASS_BACK | QR_FLAG_SHAD;
                C->p_lst = li;
                C->p_obj = si;
                QR_LANE
                {
                    FLD(cc, WMASK) = FLD(c, TMASK);
                    FLD(cc, T_BUF).f = lg->t_max;
                    FLD(cc, C_BUF) = x7;
                    FLD(cc, COL + 0).u = 0;
                    FLD(cc, COL + 1).u = 0;
                    FLD(cc, COL + 2).u = 0;
                    FLD(cc, T_MIN).u = 0;
                }
                C->l_flg = 0; C->l_lst = NIL; C->l_obj = NIL;
                walk(r, lvl + 1, le->data_p);
                QR_LANE x7 = FLD(cc, C_BUF);
                r->depth += 1;
            }
            if (pk_full(x7)) continue;

            const int do_dff = (props & QR_PROP_DIFFUSE) != 0;
            const int do_spc = (props & QR_PROP_SPECULAR) != 0;
            W x1s;                /* specular term (Xmm1), 0 if skipped */
            W x2m;
            float lx, ly, lz, x6v, len2;

            QR_LANE
            {
                float d = FLD(c, C_PTR).f;
                x7.u = m_eq(x7.f, 0.0f);       /* invert shadow mask */

                float x1 = FLD(c, NEW + 0).f, x2 = FLD(c, NEW + 1).f, x3 = FLD(c, NEW + 2).f;
                float x4 = x1 * x1, x5 = x2 * x2, x6 = x3 * x3;
                x4 = x4 + x5;
                x4 = x4 + x6;
                FLD(c, C_PTR).f = x4;

                if (do_dff)
                {
                    /* 2876-2918 */
                    d = u2f(f2u(d) & x7.u);
                    x6 = x4;
                    x5 = rsq(x4);
                    x4 = x5 * x6;
                    x6 = x6 * lg->a_qdr;
                    x4 = x4 * lg->a_lnr;
                    x6 = x6 + lg->a_cnt;
                    x6 = x6 + x4;
                    x4 = rsq(x6);
                    x6 = d;
                    d = d * x4;
                    d = d * x5;
                    d = d * m->l_dff;
                }
                else
                {
                    x6 = d;
                    d = 0.0f;
                }
                x0.f = d;
                x6v = x6;
                lx = x1; ly = x2; lz = x3;
            }

            int spec_done = 0;
            if (do_spc)
            {
                /* 2935-2973 */
                QR_LANE
                {
                    float x1 = lx, x2 = ly, x3 = lz;
                    float x4 = x6v * FLD(c, NRM + 0).f;
                    x1 = x1 - x4; x1 = x1 - x4;
                    float x5 = x6v * FLD(c, NRM + 1).f;
                    x2 = x2 - x5; x2 = x2 - x5;
                    float x6 = x6v * FLD(c, NRM + 2).f;
                    x3 = x3 - x6; x3 = x3 - x6;

                    x4 = FLD(c, RAY + 0).f; x1 = x1 * x4; x4 = x4 * x4;
                    x5 = FLD(c, RAY + 1).f; x2 = x2 * x5; x5 = x5 * x5;
                    x6 = FLD(c, RAY + 2).f; x3 = x3 * x6; x6 = x6 * x6;
                    x6 = x6 + x4;
                    x6 = x6 + x5;
                    x1 = x1 + x2;
                    x1 = x1 + x3;
                    uint32_t mm = m_lt(0.0f, x1) & x7.u;
                    x2m.u = mm;
                    x1s.u = f2u(x1) & mm;
                    len2 = x6;
                }
                if (!pk_none(x2m))
                {
                    /* 2975-3041 */
                    spec_done = 1;
                    QR_LANE
                    {
                        float x1 = x1s.f;
                        float x4 = FLD(c, C_PTR).f;
                        float x5 = rsq(len2);
                        x1 = x1 * x5;
                        x5 = rsq(x4);
                        x1 = x1 * x5;

                        uint32_t eax = m->l_pow & 0xF;
                        float x2 = x1;
                        x4 = x1;
                        x1 = 1.0f;
                        if (eax != 0)
                        {
                            do
                            {
                                x4 = sqrtf(x4);
                                uint32_t esi = 0x8 & eax;
                                eax = (eax << 1) & 0xF;
                                if (esi != 0) x1 = x1 * x4;
                            }
                            while (eax != 0);
                        }
                        eax = m->l_pow >> 4;
                        if (eax != 0)
                        {
                            float x3 = x1;
                            x1 = 1.0f;
                            do
                            {
                                uint32_t esi = 1 & eax;
                                eax = eax >> 1;
                                if (esi != 0) x1 = x1 * x2;
                                x2 = x2 * x2;
                            }
                            while (eax != 0);
                            x1 = x1 * x3;
                        }
                        x1 = x1 * m->l_spc;
                        x1s.f = x1;
                    }
                }
            }

            if (spec_done && !(props & QR_PROP_METAL))
            {
                /* LT_mtl 3090-3149: "plain" diffuse-specular blending */
                QR_LANE
                {
                    if (FLD(c, TMASK).u == 0) continue;
                    for (int k = 0; k < 3; k++)
                    {
                        float x1 = FLD(c, TEX + k).f;
                        float x4 = lg->col[k];
                        x1 = x1 * x0.f;
                        x1 = x1 * x4;
                        x4 = x4 * x1s.f;
                        x1 = x1 + x4;
                        x1 = x1 + FLD(c, COL + k).f;
                        FLD(c, COL + k).f = x1;
                    }
                }
            }
            else
            {
                /* LT_spc 3047-3084: "metal" blending (also when specular is off) */
                QR_LANE
                {
                    float d = x0.f;
                    if (spec_done) d = d + x1s.f;
                    if (FLD(c, TMASK).u == 0) continue;
                    for (int k = 0; k < 3; k++)
                    {
                        float x1 = FLD(c, TEX + k).f;
                        x1 = x1 * lg->col[k];
                        x1 = x1 * d;
                        x1 = x1 + FLD(c, COL + k).f;
                        FLD(c, COL + k).f = x1;
                    }
                }
            }
        }
    }

    /* TRANSPARENCY 3185-3598 */
    W xr, xg, xb;
    {
        QR_LANE
        {
            FLD(c, C_TRN).f = m->c_trn;
            FLD(c, C_RFL).f = m->c_rfl;
            W t = FLD(c, F_PRB);
            FLD(c, TMASK) = t;
            FLD(c, M_TRN) = t;
            FLD(c, M_RFL) = t;
            xr.u = xg.u = xb.u = 0;
        }

        int traced = 0;
        if (!pk_none(FLD(c, TMASK)) && !(props & QR_PROP_OPAQUE))
        {
            int go = 1;
            const int rfi = (props & QR_PROP_REFRACT) || (props & QR_PROP_FRESNEL);
            W x0, x4, x6, x7;

            if (rfi)
            {
                /* TR_rfi 3212-3260 */
                W x1, x2, x3;
                QR_LANE
                {
                    float a1 = FLD(c, RAY + 0).f, a2 = FLD(c, RAY + 1).f, a3 = FLD(c, RAY + 2).f;
                    float s0 = a1 * a1;
                    s0 = s0 + a2 * a2;
                    s0 = s0 + a3 * a3;
                    float inv = rsq(s0);
                    a1 = a1 * inv; a2 = a2 * inv; a3 = a3 * inv;
                    float d = a1 * FLD(c, NRM + 0).f;
                    d = d + a2 * FLD(c, NRM + 1).f;
                    d = d + a3 * FLD(c, NRM + 2).f;
                    x1.f = a1; x2.f = a2; x3.f = a3;
                    x4.f = d;
                    x6.f = m->c_rfr;
                    float b0 = d * m->c_rfr;
                    float b7 = b0 * b0;
                    b7 = b7 + 1.0f;
                    b7 = b7 - m->rfr_2;
                    x0.f = b0; x7.f = b7;
                }
                if (props & QR_PROP_FRESNEL)
                {
                    /* 3266-3295: total inner reflection */
                    QR_LANE
                    {
                        uint32_t mk = m_le(0.0f, x7.f) & FLD(c, M_TRN).u;
                        FLD(c, M_TRN).u = mk;
                        FLD(c, TMASK).u = mk;
                    }
                    if (pk_none(FLD(c, M_TRN)))
                    {
                        QR_LANE
                        {
                            FLD(c, C_TRN).u = 0;
                            FLD(c, C_RFL).f = m->c_rfl + m->c_trn;
                        }
                        go = 0;
                    }
                }
                if (go)
                {
                    /* TR_cnt 3297-3347 */
                    QR_LANE
                    {
                        x7.f = sqrtf(x7.f);
                        x0.f = x0.f + x7.f;
                    }
                    if (props & QR_PROP_REFRACT)
                    {
                        QR_LANE
                        {
                            float x5 = FLD(c, NRM + 0).f * x0.f;
                            FLD(c, NEW + 0).f = x1.f * x6.f - x5;
                            x5 = FLD(c, NRM + 1).f * x0.f;
                            FLD(c, NEW + 1).f = x2.f * x6.f - x5;
                            x5 = FLD(c, NRM + 2).f * x0.f;
                            FLD(c, NEW + 2).f = x3.f * x6.f - x5;
                        }
                    }
                    else
                    {
                        QR_LANE
                            for (int k = 0; k < 3; k++) FLD(c, NEW + k) = FLD(c, RAY + k);
                    }
                }
            }
            else
            {
                /* TR_rfe 3336-3347: propagate ray */
                QR_LANE
                    for (int k = 0; k < 3; k++) FLD(c, NEW + k) = FLD(c, RAY + k);
            }

            if (go && (props & QR_PROP_FRESNEL))
            {
                /* TR_ini 3385-3424: exact dielectric Fresnel */
                QR_LANE
                {
                    float a1 = x4.f;
                    float a2 = a1 * x6.f;
                    a2 = a2 - x7.f;
                    float a7 = x7.f * x6.f;
                    float a3 = a1;
                    a1 = a1 + a7;
                    a3 = a3 - a7;
                    float a0 = x0.f / a2;
                    a1 = a1 / a3;
                    a0 = a0 * a0;
                    a1 = a1 * a1;
                    a0 = a0 + a1;
                    a0 = a0 * -0.5f;
                    uint32_t u0 = f2u(a0) & 0x7FFFFFFFu;
                    uint32_t mk = FLD(c, M_TRN).u;
                    u0 &= mk;
                    a0 = u2f(u0) * m->c_trn;
                    u0 = f2u(a0) | (~mk & f2u(m->c_trn));
                    a0 = u2f(u0);
                    FLD(c, C_TRN).f = m->c_trn - a0;
                    FLD(c, C_RFL).f = m->c_rfl + a0;
                }
                if (r->pt != 0 && !(r->depth > QR_STACK_DEPTH - 2))
                {
                    /* 3428-3466 (RT_FEAT_PT_SPLIT_FRESNEL): below the first two
                     * levels follow ONE of the two rays, chosen with probability
                     * 0.25 + 0.5 * reflectance share, and weigh it up */
                    W x0r;
                    x0r = get_random(r, c);
                    QR_LANE
                    {
                        const float a4 = FLD(c, C_TRN).f;
                        float a5 = FLD(c, C_RFL).f;
                        float a6 = a5;
                        float a7 = a4 + a5;
                        a5 = a5 / a7;
                        a7 = 0.5f;
                        a5 = a5 * a7;
                        a7 = a7 * a7;
                        a7 = a7 + a5;
                        const float rn = x0r.f;
                        const uint32_t mt = m_ge(rn, a7) & FLD(c, M_TRN).u;
                        FLD(c, M_TRN).u = mt;
                        const uint32_t mr = m_lt(rn, a7) & FLD(c, M_RFL).u;
                        FLD(c, M_RFL).u = mr;
                        a5 = a4;
                        const float a2 = 1.0f - a7;
                        a5 = a5 / a2;
                        a6 = a6 / a7;
                        FLD(c, C_TRN).u = f2u(a5) & mt;
                        FLD(c, C_RFL).u = f2u(a6) & mr;
                    }
                }
            }

            if (go && !pk_none(FLD(c, M_TRN)))
            {
                /* TR_frn 3472-3552 */
                QR_LANE
                {
                    FLD(c, TMASK) = FLD(c, M_TRN);
                    FLD(c, T_NEW).u = 0;
                }
                if (r->depth != 0)
                {
                    const int cc = c + STEP;
                    lvl_t *C = &r->lv[lvl + 1];
                    r->depth -= 1;
                    C->p_flg = L->l_flg | QR_FLAG_PASS_THRU;
                    C->p_lst = s->mat[side];
                    C->p_obj = si;
                    QR_LANE
                    {
                        FLD(cc, WMASK) = FLD(c, TMASK);
                        FLD(cc, T_BUF).f = r->h->cam_t_max;
                        FLD(cc, C_BUF).u = 0;
                        FLD(cc, COL + 0).u = 0;
                        FLD(cc, COL + 1).u = 0;
                        FLD(cc, COL + 2).u = 0;
                        FLD(cc, T_MIN).u = 0;
                    }
                    C->l_flg = 0; C->l_lst = NIL; C->l_obj = NIL;
                    walk(r, lvl + 1, s->lst_srf[side ^ 1]);
                    r->depth += 1;
                    QR_LANE
                    {
                        float t = FLD(c, C_TRN).f;
                        xr.f = FLD(cc, COL + 0).f * t;
                        xg.f = FLD(cc, COL + 1).f * t;
                        xb.f = FLD(cc, COL + 2).f * t;
                    }
                    traced = 1;
                }
            }
        }
        (void)traced;

        /* TR_mix 3564-3598 */
        QR_LANE
        {
            float x0 = 1.0f - m->c_trn;
            x0 = x0 - m->c_rfl;
            x0 = u2f(f2u(x0) & m_le(0.0f, x0));
            float a = FLD(c, COL + 0).f * x0;
            float b = FLD(c, COL + 1).f * x0;
            float d = FLD(c, COL + 2).f * x0;
            a = xr.f + a;
            b = xg.f + b;
            d = xb.f + d;
            FLD(c, TMASK) = FLD(c, F_RND);
            if (FLD(c, TMASK).u == 0) continue;
            FLD(c, COL + 0).f = a;
            FLD(c, COL + 1).f = b;
            FLD(c, COL + 2).f = d;
        }
    }

    /* REFLECTIONS 3604-3930 */
    {
        int go = (props & QR_PROP_REFLECT) != 0;
        if (!go && !(props & QR_PROP_OPAQUE) && (props & QR_PROP_FRESNEL)) go = 1;
        if (go && pk_none(FLD(c, M_RFL))) go = 0;

        if (go)
        {
            W x0;
            QR_LANE
            {
                FLD(c, TMASK) = FLD(c, M_RFL);
                float a1 = FLD(c, RAY + 0).f, a2 = FLD(c, RAY + 1).f, a3 = FLD(c, RAY + 2).f;
                float a4 = FLD(c, NRM + 0).f, a5 = FLD(c, NRM + 1).f, a6 = FLD(c, NRM + 2).f;
                float s0 = a1 * a1;
                s0 = s0 + a2 * a2;
                s0 = s0 + a3 * a3;
                float inv = rsq(s0);
                a1 = a1 * inv; a2 = a2 * inv; a3 = a3 * inv;
                float d = a1 * a4;
                d = d + a2 * a5;
                d = d + a3 * a6;
                a4 = a4 * d; a1 = a1 - a4; a1 = a1 - a4;
                a5 = a5 * d; a2 = a2 - a5; a2 = a2 - a5;
                a6 = a6 * d; a3 = a3 - a6; a3 = a3 - a6;
                FLD(c, NEW + 0).f = a1;
                FLD(c, NEW + 1).f = a2;
                FLD(c, NEW + 2).f = a3;
                x0.f = d;
            }

            if ((props & QR_PROP_FRESNEL) && (props & QR_PROP_OPAQUE))
            {
                QR_LANE
                {
                    float a0 = x0.f;
                    if (props & QR_PROP_METAL)
                    {
                        /* 3729-3751: Fresnel for metals, fast */
                        float a6 = m->c_rcp;
                        float a4 = a0 * a6;
                        a4 = a4 + a4;
                        a0 = a0 * a0;
                        a6 = a6 * a6;
                        a6 = a6 + m->ext_2;
                        float a1 = a0 * a6;
                        a0 = a0 + a6;
                        a1 = a1 + 1.0f;
                        float a2 = a0, a3 = a1;
                        a0 = a0 + a4;
                        a1 = a1 + a4;
                        a2 = a2 - a4;
                        a3 = a3 - a4;
                        a0 = a0 / a2;
                        a1 = a1 / a3;
                        a0 = a0 + a1;
                        a0 = a0 * -0.5f;
                        a0 = u2f(f2u(a0) & 0x7FFFFFFFu);
                    }
                    else
                    {
                        /* RF_mtl 3767-3796: Fresnel for plain opaque */
                        float a4 = a0;
                        float a6 = m->c_rfr;
                        a0 = a0 * a6;
                        float a7 = a0 * a0;
                        a7 = a7 + 1.0f;
                        a7 = a7 - m->rfr_2;
                        a7 = sqrtf(a7);
                        a0 = a0 + a7;
                        float a1 = a4;
                        float a2 = a1 * a6;
                        a2 = a2 - a7;
                        a7 = a7 * a6;
                        float a3 = a1;
                        a1 = a1 + a7;
                        a3 = a3 - a7;
                        a0 = a0 / a2;
                        a1 = a1 / a3;
                        a0 = a0 * a0;
                        a1 = a1 * a1;
                        a0 = a0 + a1;
                        a0 = a0 * -0.5f;
                        a0 = u2f(f2u(a0) & 0x7FFFFFFFu);
                    }
                    /* RF_pre 3806-3815 */
                    a0 = a0 - 1.0f;
                    a0 = a0 * m->c_rfl;
                    FLD(c, C_RFL).f = m->c_rfl + a0;
                }
            }

            /* RF_frn 3819-3884 */
            QR_LANE
            {
                FLD(c, T_NEW).u = 0;
                xr.u = xg.u = xb.u = 0;
            }
            if (r->depth != 0)
            {
                const int cc = c + STEP;
                lvl_t *C = &r->lv[lvl + 1];
                r->depth -= 1;
                C->p_flg = L->l_flg | QR_FLAG_PASS_BACK;
                C->p_lst = s->mat[side];
                C->p_obj = si;
                QR_LANE
                {
                    FLD(cc, WMASK) = FLD(c, TMASK);
                    FLD(cc, T_BUF).f = r->h->cam_t_max;
                    FLD(cc, C_BUF).u = 0;
                    FLD(cc, COL + 0).u = 0;
                    FLD(cc, COL + 1).u = 0;
                    FLD(cc, COL + 2).u = 0;
                    FLD(cc, T_MIN).u = 0;
                }
                C->l_flg = 0; C->l_lst = NIL; C->l_obj = NIL;
                walk(r, lvl + 1, s->lst_srf[side]);
                r->depth += 1;
                QR_LANE
                {
                    float t = FLD(c, C_RFL).f;
                    xr.f = FLD(cc, COL + 0).f * t;
                    xg.f = FLD(cc, COL + 1).f * t;
                    xb.f = FLD(cc, COL + 2).f * t;
                }
            }
            /* RF_mix 3888-3908 */
            QR_LANE
            {
                float a = xr.f + FLD(c, COL + 0).f;
                float b = xg.f + FLD(c, COL + 1).f;
                float d = xb.f + FLD(c, COL + 2).f;
                FLD(c, TMASK) = FLD(c, F_RND);
                if (FLD(c, TMASK).u == 0) continue;
                FLD(c, COL + 0).f = a;
                FLD(c, COL + 1).f = b;
                FLD(c, COL + 2).f = d;
            }
        }
    }

    return 0;
}

/* material redirect, QD_mtr 4826-4842 */
QR_PT_REC int material_redirect(R *r, int lvl, int ei)
{
    const qr_surface *s = &r->surfs[r->elems[ei].simd];
    return material(r, lvl, ei, s->srf_t[1]);
}

/* ---- list walk: OO_cyc 1341 .. OO_out 5142 -------------------------------- */

QR_PT_REC void walk(R *r, int lvl, int ei)
{
    const int c = lvl * STEP;
    lvl_t *L = &r->lv[lvl];
    lvl_t *P = lvl > 0 ? &r->lv[lvl - 1] : NULL;
    (void)P;

    for (; ei != NIL; ei = r->elems[ei].next)
    {
        const qr_elem *e = &r->elems[ei];
        const int si = e->simd;
        const qr_surface *s = &r->surfs[si];
        const int same = (si == L->p_obj);
        const int shift = s->a_sgn[3];


        /* 1352-1373: reuse stored local hit of the previous context */
        if (same)
        {
            const int pc = c - STEP;
            QR_LANE
            {
                W a = FLD(pc, NRM + 3), b = FLD(pc, NRM + 4), d = FLD(pc, NRM + 5);
                FLD(c, DFF + shift + 0) = a;
                FLD(c, DFF + shift + 1) = b;
                FLD(c, DFF + shift + 2) = d;
            }
        }

        int do_ray = 0;

        if (!(s->srf_t[3] < 0) && L->l_obj != NIL)
        {
            /* 1385-1417: transform caching under a trnode */
            if (!same)
            {
                QR_LANE
                {
                    FLD(c, DFF + 3).f = FLD(c, DFF + 0).f - s->pos[0];
                    FLD(c, DFF + 4).f = FLD(c, DFF + 1).f - s->pos[1];
                    FLD(c, DFF + 5).f = FLD(c, DFF + 2).f - s->pos[2];
                }
            }
            if (ei == L->l_obj) L->l_obj = NIL;
        }
        else
        {
            /* OO_dff 1419-1506 */
            if (same)
            {
                do_ray = 1;
            }
            else
            {
                QR_LANE
                {
                    FLD(c, DFF + 0).f = FLD(c, ORG + 0).f - s->pos[0];
                    FLD(c, DFF + 1).f = FLD(c, ORG + 1).f - s->pos[1];
                    FLD(c, DFF + 2).f = FLD(c, ORG + 2).f - s->pos[2];
                }
                if (s->a_map[3] != 0)
                {
                    const int dst = s->srf_t[3] < 0 ? 0 : 3;
                    QR_LANE
                    {
                        float o4, o5, o6;
                        xform(s, FLD(c, DFF + 0).f, FLD(c, DFF + 1).f, FLD(c, DFF + 2).f, &o4, &o5, &o6);
                        FLD(c, DFF + dst + 0).f = o4;
                        FLD(c, DFF + dst + 1).f = o5;
                        FLD(c, DFF + dst + 2).f = o6;
                    }
                    if (s->srf_t[3] < 0) L->l_obj = e->data_p;
                    do_ray = 1;
                }
            }
            if (do_ray)
            {
                /* OO_ray 1508-1556 */
                QR_LANE
                {
                    float o4, o5, o6;
                    xform(s, FLD(c, RAY + 0).f, FLD(c, RAY + 1).f, FLD(c, RAY + 2).f, &o4, &o5, &o6);
                    FLD(c, RAY + 3).f = o4;
                    FLD(c, RAY + 4).f = o5;
                    FLD(c, RAY + 5).f = o6;
                }
            }
        }

        /* OO_trm 1558-1570 / AR_ptr 3955-4054: bounding volume of an array */
        if (e->data_i == 1)
        {
            W x7;
            QR_LANE
            {
                float x1 = FLD(c, RAY + shift + 0).f;
                float x0 = s->sci[0] * x1;
                float x5 = FLD(c, DFF + shift + 0).f;
                float q7 = s->sci[0] * x5;
                float x3 = x1;
                x1 = x1 * x0; x3 = x3 * q7; x5 = x5 * q7;

                float x2 = FLD(c, RAY + shift + 1).f;
                x0 = s->sci[1] * x2;
                float x6 = FLD(c, DFF + shift + 1).f;
                q7 = s->sci[1] * x6;
                float x4 = x2;
                x2 = x2 * x0; x4 = x4 * q7; x6 = x6 * q7;
                x1 = x1 + x2; x3 = x3 + x4; x5 = x5 + x6;

                x2 = FLD(c, RAY + shift + 2).f;
                x0 = s->sci[2] * x2;
                x6 = FLD(c, DFF + shift + 2).f;
                q7 = s->sci[2] * x6;
                x4 = x2;
                x2 = x2 * x0; x4 = x4 * q7; x6 = x6 * q7;
                x1 = x1 + x2; x3 = x3 + x4; x5 = x5 + x6;

                x5 = x5 - s->sci[3];
                x5 = x5 * x1;
                x3 = x3 * x3;
                x3 = x3 - x5;
                x7.u = m_le(0.0f, x3) & FLD(c, WMASK).u;
            }
            if (pk_none(x7))
            {
                /* AR_skp 4038-4054 */
                ei = e->data_p;
                if (ei == L->l_obj) L->l_obj = NIL;
            }
            continue;
        }

        const int tag = s->srf_t[0];
        if (tag == 0) continue;                     /* trnode element */

        W x1, x3, x4, x6, x7, x0;

        if (tag == 1)
        {
            /* PL_ptr 4062-4136 */
            if (same) continue;
            const int k = s->a_map[2];
            const uint32_t sg = s->a_sgn[2] ? SMASK : 0;
            QR_LANE
            {
                uint32_t dk = (FLD(c, DFF + k).u ^ sg) ^ SMASK;
                float rk = u2f(FLD(c, RAY + k).u ^ sg);
                x7.u = m_ne(0.0f, rk) & FLD(c, WMASK).u;
                FLD(c, T_VAL).f = u2f(dk) / rk;
            }
            clip(r, lvl, ei, x7);
            if (pk_none(x7)) continue;
            QR_LANE
            {
                FLD(c, XMASK) = x7;
                float rk = u2f(FLD(c, RAY + k).u ^ sg);
                x7.u &= m_lt(rk, 0.0f);
                FLD(c, TMASK) = x7;
            }
            if (!pk_none(x7))
            {
                L->l_flg = QR_FLAG_SIDE_OUTER;
                if (material(r, lvl, ei, 1)) return;
            }
            QR_LANE x7.u = FLD(c, TMASK).u ^ FLD(c, XMASK).u;
            if (pk_none(x7)) continue;
            QR_LANE FLD(c, TMASK) = x7;
            L->l_flg = QR_FLAG_SIDE_INNER;
            if (material(r, lvl, ei, 1)) return;
            continue;
        }

        if (tag == 3)
        {
            /* TP_ptr 4216-4277 */
            const int iI = s->a_map[0], iK = s->a_map[2];
            const float sci_i = s->sci[iI - shift], sci_k = s->sci[iK - shift];
            QR_LANE
            {
                float ri = FLD(c, RAY + iI).f, di = FLD(c, DFF + iI).f;
                float rk = FLD(c, RAY + iK).f, dk = FLD(c, DFF + iK).f;
                float a0 = di, a7 = dk;
                float a6 = dk * ri;
                float a5 = di * rk;
                a5 = a5 - a6;
                a5 = a5 * a5;
                a5 = a5 * sci_i;
                a5 = a5 * sci_k;
                a5 = u2f(f2u(a5) & 0x7FFFFFFFu);
                float a3 = sci_i * a0;
                float a4 = sci_k * a7;
                a3 = a3 * ri;
                a4 = a4 * rk;
                a3 = a3 + a4;
                a0 = a0 * a0;
                a7 = a7 * a7;
                a0 = a0 * sci_i;
                a7 = a7 * sci_k;
                a0 = a0 + a7;
                float a1 = ri * ri;
                float a2 = rk * rk;
                a1 = a1 * sci_i;
                a2 = a2 * sci_k;
                a1 = a1 + a2;
                x1.f = a1; x4.f = a3; x6.f = a0; x3.f = a5;
            }
        }
        else
        {
            /* QD_ptr 4378-4447 */
            QR_LANE
            {
                float a1 = FLD(c, RAY + shift + 0).f;
                float a0 = s->sci[0] * a1;
                float a5 = FLD(c, DFF + shift + 0).f;
                float a7 = s->sci[0] * a5;
                a7 = a7 - s->scj[0];
                float a3 = a1;
                a1 = a1 * a0;
                a3 = a3 * a7;
                a7 = a7 - s->scj[0];
                a5 = a5 * a7;

                float a2 = FLD(c, RAY + shift + 1).f;
                a0 = s->sci[1] * a2;
                float a6 = FLD(c, DFF + shift + 1).f;
                a7 = s->sci[1] * a6;
                a7 = a7 - s->scj[1];
                float a4 = a2;
                a2 = a2 * a0;
                a4 = a4 * a7;
                a7 = a7 - s->scj[1];
                a6 = a6 * a7;

                a1 = a1 + a2; a3 = a3 + a4; a5 = a5 + a6;

                a2 = FLD(c, RAY + shift + 2).f;
                a0 = s->sci[2] * a2;
                a6 = FLD(c, DFF + shift + 2).f;
                a7 = s->sci[2] * a6;
                a7 = a7 - s->scj[2];
                a4 = a2;
                a2 = a2 * a0;
                a4 = a4 * a7;
                a7 = a7 - s->scj[2];
                a6 = a6 * a7;

                a1 = a1 + a2; a3 = a3 + a4; a5 = a5 + a6;

                a5 = a5 - s->sci[3];
                a6 = a5;
                a5 = a5 * a1;
                a4 = a3;
                a3 = a3 * a3;
                a3 = a3 - a5;
                x1.f = a1; x4.f = a4; x6.f = a6; x3.f = a3;
            }
        }

        /* QD_rts 4449-4547 */
        QR_LANE x7.u = m_le(0.0f, x3.f) & FLD(c, WMASK).u;
        if (pk_none(x7)) continue;

        QR_LANE
        {
            float b = u2f(f2u(x4.f) ^ SMASK);
            float d = x3.f;
            FLD(c, DMASK).u = m_lt(d, s->d_eps) & x7.u;
            uint32_t bs = SMASK & f2u(b);
            float sd = u2f(f2u(sqrtf(d)) ^ bs);
            float bd = b + sd;
            uint32_t m_pos = m_le(0.0f, sd);
            uint32_t m_neg = m_gt(0.0f, sd);
            uint32_t cu = f2u(x6.f), bu = f2u(bd), au = f2u(x1.f);
            x6.u = (cu & m_neg) | (bu & m_pos);          /* t2nmr */
            x4.u = (bu & m_neg) | (cu & m_pos);          /* t1nmr */
            x3.u = (bu & m_neg) | (au & m_pos);          /* t2dnm */
            x0.u = (au & m_pos) | (au & m_neg);          /* a_val */
            x1.u = (au & m_neg) | (bu & m_pos);          /* t1dnm */
        }

        /* 4572-4623: root sorting for near-zero determinant */
        L->x_ptr = 0;
        if (!pk_none(FLD(c, DMASK)))
        {
            L->x_ptr = 1;
            QR_LANE
            {
                FLD(c, AMASK).u = SMASK & x0.u;
                uint32_t z1 = m_eq(x4.f, 0.0f);
                x1.u = ((x1.u | z1) ^ z1) | (z1 & f2u(1.0f));
                uint32_t z2 = m_eq(x6.f, 0.0f);
                x3.u = ((x3.u | z2) ^ z2) | (z2 & f2u(1.0f));
                float t1 = x4.f / x1.f;
                float t2 = x6.f / x3.f;
                uint32_t k1 = m_ne(x1.f, 0.0f);
                uint32_t k2 = m_ne(x3.f, 0.0f);
                uint32_t am = FLD(c, AMASK).u;
                float a2 = t1 - t2;
                a2 = u2f(f2u(a2) ^ am);
                uint32_t fm = m_le(0.0f, a2);
                a2 = u2f(f2u(a2) & fm);
                float a5 = u2f(fm & f2u(s->t_eps));
                a5 = a5 * t1;
                a5 = u2f(f2u(a5) & 0x7FFFFFFFu);
                a2 = a2 * -0.5f;
                a2 = a2 - a5;
                uint32_t u2 = f2u(a2) ^ am;
                u2 &= k1; u2 &= k2; u2 &= FLD(c, DMASK).u;
                t1 = t1 + u2f(u2);
                t2 = t2 - u2f(u2);
                x4.f = t1; x6.f = t2;
                x1.u = k1; x3.u = k2;
            }
        }

        /* QD_srt 4646-4824: side loop */
        {
            W x5;
            L->x_flg = 2;
            L->x_tag = 0;
            QR_LANE x5.u = m_gt(0.0f, x0.f) & x7.u;
            int state;                  /* 1 = rc1, 2 = rc2 */
            if (pk_none(x5)) state = 1;
            else
            {
                QR_LANE x5.u ^= x7.u;
                if (pk_none(x5)) state = 2;
                else { L->x_tag = 1; state = 1; }
            }

            int stop = 0;               /* 1 = OO_end, 2 = OO_out */
            while (!stop)
            {
                if (state == 1)
                {
                    /* QD_rc1 4695-4740 */
                    L->x_flg -= 1;
                    int skip = 0;
                    if (same)
                    {
                        int f = L->p_flg & (QR_FLAG_SIDE | QR_FLAG_PASS);
                        if (f == 1 - QR_FLAG_SIDE_OUTER || f == 2 + QR_FLAG_SIDE_OUTER) skip = 1;
                    }
                    if (skip)
                    {
                        /* QD_rt2 */
                        if (L->x_flg == 0) { stop = 1; break; }
                        state = 2;
                        continue;
                    }
                    if (L->x_ptr == 0)
                    {
                        QR_LANE
                        {
                            x4.f = x4.f / x1.f;
                            x1.u = m_ne(x1.f, 0.0f);
                        }
                    }
                    QR_LANE
                    {
                        FLD(c, XTMP1) = x6;
                        FLD(c, XTMP2) = x3;
                        FLD(c, XMASK) = x7;
                        x7.u &= x1.u;
                        FLD(c, T_VAL) = x4;
                    }
                    L->l_flg = QR_FLAG_SIDE_OUTER;
                    clip(r, lvl, ei, x7);
                    int hit = !pk_none(x7);
                    if (hit)
                    {
                        QR_LANE FLD(c, TMASK) = x7;
                        if (material_redirect(r, lvl, ei)) { stop = 2; break; }
                        if (L->x_flg == 0) { stop = 1; break; }
                        if (L->x_tag == 0)
                        {
                            QR_LANE x7.u = FLD(c, TMASK).u ^ FLD(c, XMASK).u;
                            if (pk_none(x7)) { stop = 1; break; }
                        }
                    }
                    /* QD_rs2 4767-4775 */
                    QR_LANE
                    {
                        x6 = FLD(c, XTMP1);
                        x3 = FLD(c, XTMP2);
                        x7 = FLD(c, XMASK);
                    }
                    if (L->x_flg == 0) { stop = 1; break; }
                    state = 2;
                }
                else
                {
                    /* QD_rc2 4777-4824 */
                    L->x_flg -= 1;
                    int skip = 0;
                    if (same)
                    {
                        int f = L->p_flg & (QR_FLAG_SIDE | QR_FLAG_PASS);
                        if (f == 1 - QR_FLAG_SIDE_INNER || f == 2 + QR_FLAG_SIDE_INNER) skip = 1;
                    }
                    if (skip)
                    {
                        /* QD_rt1 */
                        if (L->x_flg == 0) { stop = 1; break; }
                        state = 1;
                        continue;
                    }
                    if (L->x_ptr == 0)
                    {
                        QR_LANE
                        {
                            x6.f = x6.f / x3.f;
                            x3.u = m_ne(x3.f, 0.0f);
                        }
                    }
                    QR_LANE
                    {
                        FLD(c, XTMP1) = x4;
                        FLD(c, XTMP2) = x1;
                        FLD(c, XMASK) = x7;
                        x7.u &= x3.u;
                        FLD(c, T_VAL) = x6;
                    }
                    L->l_flg = QR_FLAG_SIDE_INNER;
                    clip(r, lvl, ei, x7);
                    int hit = !pk_none(x7);
                    if (hit)
                    {
                        QR_LANE FLD(c, TMASK) = x7;
                        if (material_redirect(r, lvl, ei)) { stop = 2; break; }
                        if (L->x_flg == 0) { stop = 1; break; }
                        if (L->x_tag == 0)
                        {
                            QR_LANE x7.u = FLD(c, TMASK).u ^ FLD(c, XMASK).u;
                            if (pk_none(x7)) { stop = 1; break; }
                        }
                    }
                    /* QD_rs1 4685-4693 */
                    QR_LANE
                    {
                        x4 = FLD(c, XTMP1);
                        x1 = FLD(c, XTMP2);
                        x7 = FLD(c, XMASK);
                    }
                    if (L->x_flg == 0) { stop = 1; break; }
                    state = 1;
                }
            }
            if (stop == 2) return;
        }
    }
}

/*
 * One lane of one packet: the primary sample of pixel (px, y), AA slot gl & 3,
 * from the jitter (1218-1285) to the running mean in the colour planes
 * (5176-5219).  px0 is the packet's first pixel (its tile's list is the
 * packet's, 1328-1333), "slot" this lane's entry of the seed / colour planes,
 * ((y * x_row + x) << fsaa) + lane.
 */
QR_PT_D void trace_lane(R *r, int y, int px, int gl, int px0, uint32_t *pseed, float *ptr_r, float *ptr_g,
                        float *ptr_b, size_t slot, float pts_o, float pts_u, float col[3])
{
    const qr_blob_header *h = r->h;
    r->depth = h->depth;
    r->seed = pseed + slot;
    lvl_t *L = &r->lv[0];
    L->p_flg = (int)h->ctx_flags;
    L->p_lst = NIL; L->p_obj = NIL;
    L->l_flg = 0; L->l_lst = NIL; L->l_obj = NIL;

    /* 1218-1285 (RT_FEAT_PT_RANDOM_SAMPLE): tent-filtered jitter, all lanes draw */
    float jit[2];
    FLD(0, TMASK).u = ONES;
    for (int pass = 0; pass < 2; pass++)
    {
        const W x0 = get_random(r, 0);
        const float a0 = x0.f + x0.f;
        const uint32_t lt = m_lt(a0, 1.0f);
        float a3 = sqrtf(a0);
        a3 = a3 - 1.0f;
        float a5 = 2.0f - a0;
        a5 = sqrtf(a5);
        const float b2 = 1.0f - a5;
        float j = u2f((f2u(a3) & lt) | (~lt & f2u(b2)));
        j = j * 0.5f;
        if (h->fsaa != 0) j = j * 0.5f;
        jit[pass] = j;
    }

    /* 1287-1322: ray init; hor_i / ver_i are exact integers (engine.cpp:3613-3624) */
    float hs = (float)px + h->hor_a[gl & 3];
    float vs = (float)y + h->ver_a[gl & 3];
    hs = hs + jit[0];
    vs = vs + jit[1];
    for (int k = 0; k < 3; k++)
    {
        float a = h->hor[k] * hs;
        const float bb = h->ver[k] * vs;
        a = a + bb;
        a = a + h->dir[k];
        FLD(0, RAY + k).f = a;
        FLD(0, ORG + k).f = h->org[k];
        FLD(0, COL + k).u = 0;
    }
    FLD(0, T_MIN).f = h->t_min;
    FLD(0, WMASK).u = ONES;
    FLD(0, T_BUF).f = h->cam_t_max;
    FLD(0, C_BUF).u = 0;

    int tx = px0 / h->tile_w;
    if (tx >= h->tls_row) tx = h->tls_row - 1;
    walk(r, 0, r->tiles[(y / h->tile_h) * h->tls_row + tx]);

    /* 5176-5219: running mean over the frames since set_pton */
    float *acc[3] = { ptr_r, ptr_g, ptr_b };
    for (int k = 0; k < 3; k++)
    {
        float a0 = FLD(0, COL + k).f * pts_o;
        const float a1 = acc[k][slot] * pts_u;
        a0 = a0 + a1;
        acc[k][slot] = a0;
        col[k] = a0;
    }
}

QR_PT_D void init(R *r, const void *blob)
{
    const uint8_t *b = (const uint8_t *)blob;
    const qr_blob_header *h = (const qr_blob_header *)blob;
    r->h      = h;
    r->surfs  = (const qr_surface  *)(b + h->off_surf);
    r->mats   = (const qr_material *)(b + h->off_mat);
    r->lgts   = (const qr_light    *)(b + h->off_lgt);
    r->elems  = (const qr_elem     *)(b + h->off_elem);
    r->tiles  = (const int32_t     *)(b + h->off_tiles);
    r->texels = (const uint32_t    *)(b + h->off_texels);
    r->pt = 1;
    r->depth = 0;
    r->seed = 0;
    for (int i = 0; i < LEVELS * STEP + NF; i++) r->mem[i].u = 0;
    for (int i = 0; i < LEVELS; i++)
    {
        lvl_t z = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        r->lv[i] = z;
    }
}

#undef FLD
#undef QR_LANE
#undef NF
#undef STEP
#undef LEVELS
#undef NIL
#undef SMASK
#undef ONES

} /* namespace qr_pt */

#endif /* QR_PT_CUH */
