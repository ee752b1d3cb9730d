/*
 * render0_oracle.c -- TEST INFRASTRUCTURE ONLY (see render0_oracle.h).
 *
 * A plain-C restatement of the reference's render0 macro-assembler routine,
 * core/tracer/tracer.cpp:1081-5405 (macros 454-1070), for RT_ELEMENT=32.
 * Every block below cites the tracer.cpp lines it follows; the order of every
 * floating-point operation is the order of the asm, each mulps/addps/subps/
 * divps/sqrps is one separately rounded IEEE-754 binary32 operation (no FMA:
 * built with -ffp-contract=off), rsqps = 1.0f / sqrtf(x) and rcpps = 1.0f / x
 * (core/config/rtconf.h:164-168, 188-193 with RT_SIMD_COMPAT_RCP/RSQ = 1),
 * compares follow core/config/rtarch_x32_512x2v2.h:706-880 (cgt = NLE and
 * cge = NLT are true on NaN, cne is true on NaN).
 *
 * The context stack is reproduced literally: one array of SIMD fields, a
 * context is a window of 64 fields and the next context starts 54 fields
 * (RT_STACK_STEP = Q*0x360, tracer.h:665) further, so T_NEW/HIT/NEW of one
 * level ARE T_MIN/ORG/RAY of the next one, and stale fields behave as in the
 * reference.  A "packet" of n lanes stands for the S lanes of one SIMD
 * register set; CHECK_MASK early-outs are taken over the packet.  With n = 1
 * every sample decides alone (the semantics of a one-thread-per-sample GPU
 * kernel); with n = S the result is bit-identical to the reference build of
 * that SIMD width.
 *
 * PARITY PINNING: checked against frames rendered by the unmodified reference
 * (oracle/_ref/qr_ref_harness) for all 18 test scenes and the 3 demo scenes,
 * see tests/test_oracle_golden.py and tools/make_golden.py.
 */

#include <math.h>
#include <string.h>
#include <stdlib.h>

#include "qr_scene_blob.h"
#include "render0_oracle.h"

#define MAXS   64
#define NF     64                       /* fields per context */
#define STEP   54                       /* RT_STACK_STEP / (Q*0x10) */
#define LEVELS (QR_STACK_DEPTH + 3)
#define NIL    QR_NIL

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

#define SMASK 0x80000000u
#define ONES  0xFFFFFFFFu

/* packed scalar fields of one context: PARAM, LOCAL, XMISC (tracer.h:561-592) */
typedef struct lvl_t
{
    int p_flg, p_lst, p_obj;            /* PARAM(FLG/LST/OBJ) */
    int l_flg, l_lst, l_obj;            /* LOCAL(FLG/LST/OBJ) */
    int x_ptr, x_flg, x_tag;            /* XMISC(PTR/FLG/TAG) */
} lvl_t;

typedef struct R
{
    const qr_blob_header *h;
    const qr_surface  *surfs;
    const qr_material *mats;
    const qr_light    *lgts;
    const qr_elem     *elems;
    const int32_t     *tiles;
    const uint32_t    *texels;
    int   n;                            /* lanes in the current packet */
    int   depth;                        /* inf_DEPTH */
    lvl_t lv[LEVELS];
    W     mem[LEVELS * STEP + NF][MAXS];
    qr_oracle_stats st;
    /* path tracer (pt_on): seed / colour planes of the frame, the slot of the
     * current packet's lane 0 in them (inf_PRNGS), 1 / samples and 1 - that */
    qr_oracle_pt *pt;
    size_t prngs;
    float  pts_o, pts_u;
} R;

#define FLD(c, f) (r->mem[(c) + (f)])

/* ---- instruction semantics ------------------------------------------------ */

static inline uint32_t m_lt(float a, float b) { return a <  b ? ONES : 0; } /* clt */
static inline uint32_t m_le(float a, float b) { return a <= b ? ONES : 0; } /* cle */
static inline uint32_t m_gt(float a, float b) { return !(a <= b) ? ONES : 0; } /* cgt = NLE */
static inline uint32_t m_ge(float a, float b) { return !(a <  b) ? ONES : 0; } /* cge = NLT */
static inline uint32_t m_eq(float a, float b) { return a == b ? ONES : 0; } /* ceq */
static inline uint32_t m_ne(float a, float b) { return a != b ? ONES : 0; } /* cne = NEQ_UQ */

static inline float    u2f(uint32_t u) { W w; w.u = u; return w.f; }
static inline uint32_t f2u(float f)    { W w; w.f = f; return w.u; }

static inline float rsq(float x) { return 1.0f / sqrtf(x); }

/* sinps_rr / cosps_rr, tracer.cpp:1032-1057: power series, the terms added by
 * fused multiply-adds (fmaps3ld is VFMADD231PS on the AVX-512 targets) */
static inline float sin_ps(float x)
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

static inline float cos_ps(float x)
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
static inline int32_t cvn(float x)
{
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return (int32_t)0x80000000u;
    return (int32_t)lrintf(x);
}

/* cvmps: round towards minus infinity */
static inline int32_t cvm(float x)
{
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return (int32_t)0x80000000u;
    return (int32_t)floorf(x);
}

static int none(const W *m, int n)
{
    for (int l = 0; l < n; l++) if (m[l].u) return 0;
    return 1;
}

static int full(const W *m, int n)
{
    for (int l = 0; l < n; l++) if (m[l].u != ONES) return 0;
    return 1;
}

/*
 * 3x3 transform as written at tracer.cpp:1447-1479 (diff), 1512-1548 (ray),
 * 2063-2095 (clip): diagonal products first, then the two off-diagonal terms
 * of each row in column order; a_map[L] == 1 keeps the diagonal only.
 */
static inline void xform(const qr_surface *s, float v1, float v2, float v3,
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

static void walk(R *r, int lvl, int ei);

/* ---- custom clipping subroutine: tracer.cpp:1597-2160 (CC_clp .. CC_out) -- */
/* x7 in/out is the running tmask (Xmm7); returns with it updated */

static void clip(R *r, int lvl, int ei, W *x7)
{
    const int c = lvl * STEP, n = r->n;
    lvl_t *L = &r->lv[lvl];
    const qr_elem *e = &r->elems[ei];
    const qr_surface *s = &r->surfs[e->simd];
    const int shift = s->a_sgn[3];
    W x4[MAXS], x5[MAXS], x6[MAXS];

    /* 1599-1681: depth test, near plane, hit point, local hit */
    for (int l = 0; l < n; l++)
    {
        float t = FLD(c, T_VAL)[l].f;
        x7[l].u &= m_gt(FLD(c, T_BUF)[l].f, t);
        x7[l].u &= m_lt(FLD(c, T_MIN)[l].f, t);

        float hx = FLD(c, RAY + 0)[l].f * t; hx = hx + FLD(c, ORG + 0)[l].f;
        float hy = FLD(c, RAY + 1)[l].f * t; hy = hy + FLD(c, ORG + 1)[l].f;
        float hz = FLD(c, RAY + 2)[l].f * t; hz = hz + FLD(c, ORG + 2)[l].f;
        FLD(c, HIT + 0)[l].f = hx;
        FLD(c, HIT + 1)[l].f = hy;
        FLD(c, HIT + 2)[l].f = hz;

        if (s->a_map[3] != 0)
        {
            float li = FLD(c, RAY + 3)[l].f * t; li = li + FLD(c, DFF + 3)[l].f;
            float lj = FLD(c, RAY + 4)[l].f * t; lj = lj + FLD(c, DFF + 4)[l].f;
            float lk = FLD(c, RAY + 5)[l].f * t; lk = lk + FLD(c, DFF + 5)[l].f;
            FLD(c, NEW + 3)[l].f = li;
            FLD(c, NEW + 4)[l].f = lj;
            FLD(c, NEW + 5)[l].f = lk;
            x4[l].f = li; x5[l].f = lj; x6[l].f = lk;
        }
        else
        {
            hx = hx - s->pos[0];
            hy = hy - s->pos[1];
            hz = hz - s->pos[2];
            FLD(c, NEW + 0)[l].f = hx;
            FLD(c, NEW + 1)[l].f = hy;
            FLD(c, NEW + 2)[l].f = hz;
            x4[l].f = hx; x5[l].f = hy; x6[l].f = hz;
        }
    }

    /* 1706-1856: conic singularity solver */
    if (s->conic != 0 && L->x_ptr != 0)
    {
        const int iI = s->a_map[0], iJ = s->a_map[1], iK = s->a_map[2];
        W x0[MAXS];
        for (int l = 0; l < n; l++)
        {
            float a1 = FLD(c, NEW + iI)[l].f; a1 = a1 * a1;
            float a0 = a1;
            if (s->conic != 2)
            {
                float a2 = FLD(c, NEW + iJ)[l].f; a2 = a2 * a2;
                a0 = a0 + a2;
            }
            float a3 = FLD(c, NEW + iK)[l].f; a3 = a3 * a3;
            a0 = a0 + a3;
            x0[l].u = m_lt(a0, s->t_eps) & FLD(c, DMASK)[l].u;
        }
        if (!none(x0, n))
        {
            for (int l = 0; l < n; l++)
            {
                uint32_t hm = x0[l].u;
                float one = 1.0f;
                uint32_t q1 = (FLD(c, DFF + iI)[l].u & SMASK) ^ f2u(one);
                uint32_t q2 = 0;
                float q3 = s->sci[iI - shift];
                float q4 = one;
                if (s->conic != 2)
                {
                    q2 = (FLD(c, DFF + iJ)[l].u & SMASK) ^ f2u(one);
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

                uint32_t am = FLD(c, AMASK)[l].u;
                uint32_t ts = (L->l_flg & 1) ? SMASK : 0;   /* srf_SBASE + FLG*Q*16 */
                uint32_t sk = FLD(c, DFF + iK)[l].u & SMASK;
                uint32_t u3 = f2u(p3) ^ sk;
                u3 ^= (ts & am) ^ am;
                uint32_t tsn = (ts | am) ^ am;
                uint32_t u1 = f2u(p1) ^ tsn;
                uint32_t u2 = f2u(p2) ^ tsn;

                FLD(c, NEW + iI)[l].u = ((FLD(c, NEW + iI)[l].u | hm) ^ hm) | (u1 & hm);
                if (s->conic != 2)
                {
                    FLD(c, NEW + iJ)[l].u = ((FLD(c, NEW + iJ)[l].u | hm) ^ hm) | (u2 & hm);
                }
                FLD(c, NEW + iK)[l].u = ((FLD(c, NEW + iK)[l].u | hm) ^ hm) | (u3 & hm);
            }
            for (int l = 0; l < n; l++)
            {
                x4[l] = FLD(c, NEW + shift + 0)[l];
                x5[l] = FLD(c, NEW + shift + 1)[l];
                x6[l] = FLD(c, NEW + shift + 2)[l];
            }
        }
    }

    /* 1874-1927: axis min/max clipping on the (un-mapped) local point */
    for (int l = 0; l < n; l++)
    {
        if (s->minmax_t & 1)  x7[l].u &= m_le(s->min[0], x4[l].f);
        if (s->minmax_t & 8)  x7[l].u &= m_ge(s->max[0], x4[l].f);
        if (s->minmax_t & 2)  x7[l].u &= m_le(s->min[1], x5[l].f);
        if (s->minmax_t & 16) x7[l].u &= m_ge(s->max[1], x5[l].f);
        if (s->minmax_t & 4)  x7[l].u &= m_le(s->min[2], x6[l].f);
        if (s->minmax_t & 32) x7[l].u &= m_ge(s->max[2], x6[l].f);
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
                for (int l = 0; l < n; l++) x7[l].u = ~x7[l].u & FLD(c, C_ACC)[l].u;
            }
            else
            {
                for (int l = 0; l < n; l++)
                {
                    FLD(c, C_ACC)[l].u = x7[l].u;
                    x7[l].u = s->c_def;
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
                for (int l = 0; l < n; l++)
                {
                    FLD(c, NRM + 3)[l].f = FLD(c, NRM + 0)[l].f - cs->pos[0];
                    FLD(c, NRM + 4)[l].f = FLD(c, NRM + 1)[l].f - cs->pos[1];
                    FLD(c, NRM + 5)[l].f = FLD(c, NRM + 2)[l].f - cs->pos[2];
                }
                if (di == redx) redx = NIL;
                have_local = 1;
            }
        }
        else
        if (ce->simd == L->l_lst)                   /* 2006-2037: same trnode */
        {
            for (int l = 0; l < n; l++)
            {
                FLD(c, NRM + 0)[l].f = FLD(c, NEW + 3)[l].f + s->pos[0];
                FLD(c, NRM + 1)[l].f = FLD(c, NEW + 4)[l].f + s->pos[1];
                FLD(c, NRM + 2)[l].f = FLD(c, NEW + 5)[l].f + s->pos[2];
            }
            redx = ce->data_p;
            continue;
        }

        if (!have_local)                            /* 2039-2125: CC_dff */
        {
            int cached = 0;
            for (int l = 0; l < n; l++)
            {
                float d1 = FLD(c, HIT + 0)[l].f - cs->pos[0];
                float d2 = FLD(c, HIT + 1)[l].f - cs->pos[1];
                float d3 = FLD(c, HIT + 2)[l].f - cs->pos[2];
                FLD(c, NRM + 0)[l].f = d1;
                FLD(c, NRM + 1)[l].f = d2;
                FLD(c, NRM + 2)[l].f = d3;
                if (cs->a_map[3] != 0)
                {
                    float o4, o5, o6;
                    xform(cs, d1, d2, d3, &o4, &o5, &o6);
                    if (cs->srf_t[3] < 0)
                    {
                        FLD(c, NRM + 0)[l].f = o4;
                        FLD(c, NRM + 1)[l].f = o5;
                        FLD(c, NRM + 2)[l].f = o6;
                        cached = 1;
                    }
                    else
                    {
                        FLD(c, NRM + 3)[l].f = o4;
                        FLD(c, NRM + 4)[l].f = o5;
                        FLD(c, NRM + 5)[l].f = o6;
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
            for (int l = 0; l < n; l++)
            {
                float v = u2f(FLD(c, NRM + k)[l].u ^ sg);
                x4[l].u = ce->data_i < 0 ? m_ge(v, 0.0f) : m_le(v, 0.0f);
            }
        }
        else
        if (cs->srf_t[2] == 2)                      /* QD_clp 4910-4951 */
        {
            for (int l = 0; l < n; l++)
            {
                float dx = FLD(c, NRM + cshift + 0)[l].f;
                float dy = FLD(c, NRM + cshift + 1)[l].f;
                float dz = FLD(c, NRM + cshift + 2)[l].f;
                float a1 = cs->scj[0] + cs->scj[0]; a1 = a1 * dx;
                float a4 = dx * dx; a4 = a4 * cs->sci[0]; a4 = a4 - a1;
                float a2 = cs->scj[1] + cs->scj[1]; a2 = a2 * dy;
                float a5 = dy * dy; a5 = a5 * cs->sci[1]; a5 = a5 - a2;
                float a3 = cs->scj[2] + cs->scj[2]; a3 = a3 * dz;
                float a6 = dz * dz; a6 = a6 * cs->sci[2]; a6 = a6 - a3;
                a4 = a4 - cs->sci[3];
                a4 = a4 + a5;
                a4 = a4 + a6;
                x4[l].u = ce->data_i < 0 ? m_ge(a4, 0.0f) : m_le(a4, 0.0f);
            }
        }
        else
        if (cs->srf_t[2] == 3)                      /* TP_clp 4341-4370 */
        {
            for (int l = 0; l < n; l++)
            {
                float dx = FLD(c, NRM + cshift + 0)[l].f;
                float dy = FLD(c, NRM + cshift + 1)[l].f;
                float dz = FLD(c, NRM + cshift + 2)[l].f;
                float a4 = dx * dx; a4 = a4 * cs->sci[0];
                float a5 = dy * dy; a5 = a5 * cs->sci[1];
                float a6 = dz * dz; a6 = a6 * cs->sci[2];
                a4 = a4 - cs->sci[3];
                a4 = a4 + a5;
                a4 = a4 + a6;
                x4[l].u = ce->data_i < 0 ? m_ge(a4, 0.0f) : m_le(a4, 0.0f);
            }
        }
        /* CC_ret 2138-2140 (srf_t[2] == 0 would use a stale Xmm4) */
        for (int l = 0; l < n; l++) x7[l].u &= x4[l].u;
    }
}

/*
 * GET_RANDOM 1014-1028, RT_PRNG = LCG24 (tracer.h:53, engine.cpp:867-873): the
 * seed of every lane advances s = s * 214013 + 2531011 (mod 2^32) and is stored
 * for the lanes of TMASK only; the number is bits 8..31 of the new seed over
 * 2^24.  Lanes outside TMASK get a number too (from a seed that stays).
 */
static void get_random(R *r, int c, W *x0)
{
    for (int l = 0; l < r->n; l++)
    {
        uint32_t sd = r->pt->pseed[r->prngs + (size_t)l];
        sd = sd * 214013u + 2531011u;
        if (FLD(c, TMASK)[l].u) r->pt->pseed[r->prngs + (size_t)l] = sd;
        const float a = (float)(int32_t)((sd >> 8) & 0xFFFFFFu);
        const float b = (float)(int32_t)0xFFFFFF + 1.0f;
        x0[l].f = a / b;
    }
}

/* ---- material: tracer.cpp:4139-4193, 4280-4336, 4845-4905, 2166-3947 ------ */
/* kind: 1 PL_mat, 2 QD_mat, 3 TP_mat.  Returns 1 when the walk must stop
 * (OO_out from CHECK_SHAD), 0 to return to the calling solver (SR_rt*). */

static int material(R *r, int lvl, int ei, int kind)
{
    const int c = lvl * STEP, n = r->n;
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
        for (int l = 0; l < n; l++) FLD(c, C_BUF)[l].u |= FLD(c, TMASK)[l].u;
        if (full(FLD(c, C_BUF), n)) return 1;
        return 0;
    }

    int have_nrm = 0;

    if (kind == 1)
    {
        /* PL_mat 4149-4193 */
        if (props & QR_PROP_TEXTURE)
        {
            const uint32_t sgi = s->a_sgn[0] ? SMASK : 0, sgj = s->a_sgn[1] ? SMASK : 0;
            for (int l = 0; l < n; l++)
            {
                FLD(c, TEX_U)[l].u = FLD(c, NEW + s->a_map[0])[l].u ^ sgi;
                FLD(c, TEX_V)[l].u = FLD(c, NEW + s->a_map[1])[l].u ^ sgj;
            }
        }
        if (props & QR_PROP_NORMAL)
        {
            const uint32_t sgk = s->a_sgn[2] ? SMASK : 0;
            for (int l = 0; l < n; l++)
            {
                FLD(c, NRM + s->a_map[0])[l].u = 0;
                FLD(c, NRM + s->a_map[1])[l].u = 0;
                FLD(c, NRM + s->a_map[2])[l].u = (f2u(1.0f) ^ tside) ^ sgk;
            }
            have_nrm = 1;
        }
    }
    else
    {
        /* QD_mat 4855-4899 / TP_mat 4290-4330 */
        if (props & QR_PROP_NORMAL)
        {
            for (int l = 0; l < n; l++)
            {
                float x4 = FLD(c, NEW + shift + 0)[l].f * s->sci[0];
                float x5 = FLD(c, NEW + shift + 1)[l].f * s->sci[1];
                float x6 = FLD(c, NEW + shift + 2)[l].f * s->sci[2];
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
                FLD(c, NRM + shift + 0)[l].f = x4 * x0;
                FLD(c, NRM + shift + 1)[l].f = x5 * x0;
                FLD(c, NRM + shift + 2)[l].f = x6 * x0;
            }
            have_nrm = 1;
        }
    }

    /* MT_nrm 2184-2263: transform normal with the trnode's transposed matrix */
    if (have_nrm && s->a_map[3] != 0)
    {
        const qr_surface *t = &r->surfs[s->trnode];
        for (int l = 0; l < n; l++)
        {
            float n1 = FLD(c, NRM + 3)[l].f, n2 = FLD(c, NRM + 4)[l].f, n3 = FLD(c, NRM + 5)[l].f;
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
            FLD(c, NRM + 0)[l].f = x4;
            FLD(c, NRM + 1)[l].f = x5;
            FLD(c, NRM + 2)[l].f = x6;
        }
    }

    /* MT_mat 2267-2327 */
    for (int l = 0; l < n; l++)
    {
        W a = FLD(c, NEW + shift + 0)[l], b = FLD(c, NEW + shift + 1)[l], d = FLD(c, NEW + shift + 2)[l];
        FLD(c, NRM + 3)[l] = a;
        FLD(c, NRM + 4)[l] = b;
        FLD(c, NRM + 5)[l] = d;
    }
    L->l_lst = ei;

    const qr_material *m = &r->mats[s->mat[side]];

    for (int l = 0; l < n; l++)
    {
        uint32_t p = 0;
        if (props & QR_PROP_TEXTURE)
        {
            float tx = FLD(c, TEX_U + m->t_map[0])[l].f;
            float ty = FLD(c, TEX_U + m->t_map[1])[l].f;
            tx = tx - m->xoffs;
            ty = ty - m->yoffs;
            tx = tx * m->xscal;
            ty = ty * m->yscal;
            uint32_t ix = (uint32_t)cvm(tx) & m->xmask;
            uint32_t iy = ((uint32_t)cvm(ty) & m->ymask) << m->yshft;
            p = (ix + iy) << 2;
        }
        FLD(c, C_PTR)[l].u = p;
    }
    /* PAINT_FRAG 653-662 */
    for (int l = 0; l < n; l++)
    {
        if (FLD(c, TMASK)[l].u == 0) continue;
        FLD(c, T_BUF)[l] = FLD(c, T_VAL)[l];
        FLD(c, C_BUF)[l].u = r->texels[m->tex + (FLD(c, C_PTR)[l].u >> 2)];
        r->st.shaded_hits++;
        if (kind != 1 && (props & QR_PROP_TEXTURE)) r->st.tex_nonplane++;
    }
    /* PAINT_COLX 664-673 */
    for (int l = 0; l < n; l++)
    {
        uint32_t cb = FLD(c, C_BUF)[l].u;
        for (int k = 0; k < 3; k++)
        {
            int sh = k == 0 ? 16 : k == 1 ? 8 : 0;
            float v = (float)(int32_t)((cb >> sh) & m->cmask);
            v = v / m->clamp;
            if (props & QR_PROP_GAMMA) v = v * v;
            FLD(c, TEX + k)[l].f = v;
        }
    }

    /* LIGHTS 2333-3179 */
    for (int l = 0; l < n; l++)
    {
        FLD(c, F_RND)[l] = FLD(c, TMASK)[l];
        FLD(c, F_PRB)[l] = FLD(c, TMASK)[l];
    }

    if (r->pt != NULL)
    {
        /* path tracer, 2339-2701: instead of the lights one diffuse bounce */
        W x1[MAXS], x2[MAXS], x3[MAXS];
        for (int l = 0; l < n; l++) x1[l].u = x2[l].u = x3[l].u = 0;

        int bounce = (props & QR_PROP_DIFFUSE) != 0;        /* CHECK_PROP(PT_mix, RT_PROP_DIFFUSE) */
        if (bounce && !(r->depth > QR_STACK_DEPTH - 5))
        {
            /* 2352-2396 (RT_FEAT_PT_SPLIT_DEPTH): deeper levels go on with the
             * probability of the brightest colour channel */
            W x0[MAXS], x4[MAXS];
            for (int l = 0; l < n; l++)
            {
                float a = FLD(c, TEX + 0)[l].f;             /* maxps: the source when unordered */
                a = a > FLD(c, TEX + 1)[l].f ? a : FLD(c, TEX + 1)[l].f;
                a = a > FLD(c, TEX + 2)[l].f ? a : FLD(c, TEX + 2)[l].f;
                x4[l].f = a;
            }
            get_random(r, c, x0);
            for (int l = 0; l < n; l++)
            {
                x0[l].u = m_lt(x0[l].f, x4[l].f) & FLD(c, F_PRB)[l].u;
                FLD(c, F_PRB)[l] = x0[l];
                FLD(c, TMASK)[l] = x0[l];
            }
            if (none(x0, n))
            {
                bounce = 0;                                 /* PT_chk -> PT_mix */
            }
            else
            {
                for (int l = 0; l < n; l++)
                {
                    const float x5 = 1.0f / x4[l].f;        /* rcpps, all lanes */
                    for (int k = 0; k < 3; k++) FLD(c, TEX + k)[l].f = FLD(c, TEX + k)[l].f * x5;
                }
            }
        }
        if (bounce)
        {
            /* 2398-2530: orthonormal basis around the normal (its fields are
             * borrowed ones: TEX_U, TEX_V, C_PTR, C_ACC, F_RFL, T_VAL) */
            W x6[MAXS], x0[MAXS];
            for (int l = 0; l < n; l++)
            {
                const float n1 = FLD(c, NRM + 0)[l].f, n2 = FLD(c, NRM + 1)[l].f, n3 = FLD(c, NRM + 2)[l].f;
                const float r4 = FLD(c, RAY + 0)[l].f, r5 = FLD(c, RAY + 1)[l].f, r6 = FLD(c, RAY + 2)[l].f;
                float a0, a7;
                a0 = n2 * r6; a7 = n3 * r5; float u4 = a0 - a7;
                a0 = n3 * r4; a7 = n1 * r6; float u5 = a0 - a7;
                a0 = n1 * r5; a7 = n2 * r4; float u6 = a0 - a7;
                float s1 = u4 * u4, s2 = u5 * u5, s3 = u6 * u6;
                s1 = s1 + s2;
                s1 = s1 + s3;
                const float inv = rsq(s1);
                u4 = u4 * inv; u5 = u5 * inv; u6 = u6 * inv;
                FLD(c, TEX_U)[l].f = u4; FLD(c, TEX_V)[l].f = u5; FLD(c, C_PTR)[l].f = u6;
                a0 = n2 * u6; a7 = n3 * u5; FLD(c, C_ACC)[l].f = a0 - a7;
                a0 = n3 * u4; a7 = n1 * u6; FLD(c, F_RFL)[l].f = a0 - a7;
                a0 = n1 * u5; a7 = n2 * u4; FLD(c, T_VAL)[l].f = a0 - a7;
            }
            /* 2532-2590: cosine-weighted direction over the hemisphere */
            get_random(r, c, x0);
            for (int l = 0; l < n; l++)
            {
                x6[l].f = x0[l].f;
                float a0 = 1.0f - x6[l].f;
                x6[l].f = sqrtf(x6[l].f);
                a0 = sqrtf(a0);
                x1[l].f = FLD(c, NRM + 0)[l].f * a0;
                x2[l].f = FLD(c, NRM + 1)[l].f * a0;
                x3[l].f = FLD(c, NRM + 2)[l].f * a0;
            }
            get_random(r, c, x0);
            for (int l = 0; l < n; l++)
            {
                const float pi = (float)3.14159265358979323846;     /* mat_GPC10, object.cpp:4130 */
                float a0 = x0[l].f + x0[l].f;
                a0 = a0 * pi;
                a0 = a0 - pi;
                float a4 = cos_ps(a0);
                a4 = a4 * x6[l].f;
                x1[l].f = x1[l].f + FLD(c, TEX_U)[l].f * a4;
                x2[l].f = x2[l].f + FLD(c, TEX_V)[l].f * a4;
                x3[l].f = x3[l].f + FLD(c, C_PTR)[l].f * a4;
                a4 = sin_ps(a0);
                a4 = a4 * x6[l].f;
                x1[l].f = x1[l].f + FLD(c, C_ACC)[l].f * a4;
                x2[l].f = x2[l].f + FLD(c, F_RFL)[l].f * a4;
                x3[l].f = x3[l].f + FLD(c, T_VAL)[l].f * a4;
                FLD(c, NEW + 0)[l] = x1[l];
                FLD(c, NEW + 1)[l] = x2[l];
                FLD(c, NEW + 2)[l] = x3[l];
                FLD(c, T_NEW)[l].u = 0;
                x1[l].u = x2[l].u = x3[l].u = 0;
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
                for (int l = 0; l < n; l++)
                {
                    FLD(cc, WMASK)[l] = FLD(c, TMASK)[l];
                    FLD(cc, T_BUF)[l].f = r->h->cam_t_max;
                    FLD(cc, C_BUF)[l].u = 0;
                    FLD(cc, COL + 0)[l].u = 0;
                    FLD(cc, COL + 1)[l].u = 0;
                    FLD(cc, COL + 2)[l].u = 0;
                    FLD(cc, T_MIN)[l].u = 0;
                    if (FLD(c, TMASK)[l].u) r->st.rays_reflect++;
                }
                C->l_flg = 0; C->l_lst = NIL; C->l_obj = NIL;
                walk(r, lvl + 1, s->lst_srf[side]);
                r->depth += 1;
                for (int l = 0; l < n; l++)
                {
                    x1[l].f = FLD(cc, COL + 0)[l].f * m->l_dff;
                    x2[l].f = FLD(cc, COL + 1)[l].f * m->l_dff;
                    x3[l].f = FLD(cc, COL + 2)[l].f * m->l_dff;
                    x1[l].f = x1[l].f * FLD(c, TEX + 0)[l].f;
                    x2[l].f = x2[l].f * FLD(c, TEX + 1)[l].f;
                    x3[l].f = x3[l].f * FLD(c, TEX + 2)[l].f;
                }
            }
        }
        /* PT_mix 2664-2699: self-emission, then the radiance of the hit lanes */
        for (int l = 0; l < n; l++)
        {
            x1[l].f = x1[l].f + m->col[0];
            x2[l].f = x2[l].f + m->col[1];
            x3[l].f = x3[l].f + m->col[2];
            FLD(c, TMASK)[l] = FLD(c, F_RND)[l];
            if (FLD(c, TMASK)[l].u == 0) continue;
            FLD(c, COL + 0)[l] = x1[l];
            FLD(c, COL + 1)[l] = x2[l];
            FLD(c, COL + 2)[l] = x3[l];
        }
    }
    else
    if (props & QR_PROP_LIGHT)
    {
        /* LT_set 3164-3177 */
        for (int l = 0; l < n; l++)
        {
            if (FLD(c, TMASK)[l].u == 0) continue;
            for (int k = 0; k < 3; k++) FLD(c, COL + k)[l] = FLD(c, TEX + k)[l];
        }
    }
    else
    {
        /* ambient 2721-2756 */
        for (int l = 0; l < n; l++)
        {
            if (FLD(c, TMASK)[l].u == 0) continue;
            for (int k = 0; k < 3; k++)
                FLD(c, COL + k)[l].f = FLD(c, TEX + k)[l].f * r->h->amb[k];
        }

        /* LT_cyc 2760-3156 */
        for (int li = s->lst_lgt[side]; li != NIL; li = r->elems[li].next)
        {
            const qr_elem *le = &r->elems[li];
            const qr_light *lg = &r->lgts[le->simd];
            W x0[MAXS], x7[MAXS];

            for (int l = 0; l < n; l++)
            {
                float x1 = lg->pos[0] - FLD(c, HIT + 0)[l].f;
                FLD(c, NEW + 0)[l].f = x1;
                x1 = x1 * FLD(c, NRM + 0)[l].f;
                float x2 = lg->pos[1] - FLD(c, HIT + 1)[l].f;
                FLD(c, NEW + 1)[l].f = x2;
                x2 = x2 * FLD(c, NRM + 1)[l].f;
                float x3 = lg->pos[2] - FLD(c, HIT + 2)[l].f;
                FLD(c, NEW + 2)[l].f = x3;
                x3 = x3 * FLD(c, NRM + 2)[l].f;
                float d = x1 + x2;
                d = d + x3;
                x0[l].f = d;
                x7[l].u = m_lt(0.0f, d) & FLD(c, TMASK)[l].u;
            }
            if (none(x7, n)) continue;

            /* shadows 2794-2850 */
            {
                const int cc = c + STEP;
                lvl_t *C = &r->lv[lvl + 1];
                for (int l = 0; l < n; l++)
                {
                    x7[l].u = m_eq(x7[l].f, 0.0f);   /* ceqps with 0: inverted lmask */
                    FLD(c, C_PTR)[l].f = x0[l].f;
                }
                r->depth -= 1;
                C->p_flg = L->l_flg | QR_FLAG_PASS_BACK | QR_FLAG_SHAD;
                C->p_lst = li;
                C->p_obj = si;
                for (int l = 0; l < n; l++)
                {
                    FLD(cc, WMASK)[l] = FLD(c, TMASK)[l];
                    FLD(cc, T_BUF)[l].f = lg->t_max;
                    FLD(cc, C_BUF)[l] = x7[l];
                    FLD(cc, COL + 0)[l].u = 0;
                    FLD(cc, COL + 1)[l].u = 0;
                    FLD(cc, COL + 2)[l].u = 0;
                    FLD(cc, T_MIN)[l].u = 0;
                    if (FLD(c, TMASK)[l].u && !x7[l].u) r->st.rays_shadow++;
                }
                C->l_flg = 0; C->l_lst = NIL; C->l_obj = NIL;
                walk(r, lvl + 1, le->data_p);
                for (int l = 0; l < n; l++) x7[l] = FLD(cc, C_BUF)[l];
                r->depth += 1;
            }
            if (full(x7, n)) continue;

            const int do_dff = (props & QR_PROP_DIFFUSE) != 0;
            const int do_spc = (props & QR_PROP_SPECULAR) != 0;
            W x1s[MAXS];                /* specular term (Xmm1), 0 if skipped */
            W x2m[MAXS];
            float lx[MAXS], ly[MAXS], lz[MAXS], x6v[MAXS], len2[MAXS];

            for (int l = 0; l < n; l++)
            {
                float d = FLD(c, C_PTR)[l].f;
                x7[l].u = m_eq(x7[l].f, 0.0f);       /* invert shadow mask */

                float x1 = FLD(c, NEW + 0)[l].f, x2 = FLD(c, NEW + 1)[l].f, x3 = FLD(c, NEW + 2)[l].f;
                float x4 = x1 * x1, x5 = x2 * x2, x6 = x3 * x3;
                x4 = x4 + x5;
                x4 = x4 + x6;
                FLD(c, C_PTR)[l].f = x4;

                if (do_dff)
                {
                    /* 2876-2918 */
                    d = u2f(f2u(d) & x7[l].u);
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
                x0[l].f = d;
                x6v[l] = x6;
                lx[l] = x1; ly[l] = x2; lz[l] = x3;
            }

            int spec_done = 0;
            if (do_spc)
            {
                /* 2935-2973 */
                for (int l = 0; l < n; l++)
                {
                    float x1 = lx[l], x2 = ly[l], x3 = lz[l];
                    float x4 = x6v[l] * FLD(c, NRM + 0)[l].f;
                    x1 = x1 - x4; x1 = x1 - x4;
                    float x5 = x6v[l] * FLD(c, NRM + 1)[l].f;
                    x2 = x2 - x5; x2 = x2 - x5;
                    float x6 = x6v[l] * FLD(c, NRM + 2)[l].f;
                    x3 = x3 - x6; x3 = x3 - x6;

                    x4 = FLD(c, RAY + 0)[l].f; x1 = x1 * x4; x4 = x4 * x4;
                    x5 = FLD(c, RAY + 1)[l].f; x2 = x2 * x5; x5 = x5 * x5;
                    x6 = FLD(c, RAY + 2)[l].f; x3 = x3 * x6; x6 = x6 * x6;
                    x6 = x6 + x4;
                    x6 = x6 + x5;
                    x1 = x1 + x2;
                    x1 = x1 + x3;
                    uint32_t mm = m_lt(0.0f, x1) & x7[l].u;
                    x2m[l].u = mm;
                    x1s[l].u = f2u(x1) & mm;
                    len2[l] = x6;
                }
                if (!none(x2m, n))
                {
                    /* 2975-3041 */
                    spec_done = 1;
                    for (int l = 0; l < n; l++)
                    {
                        float x1 = x1s[l].f;
                        float x4 = FLD(c, C_PTR)[l].f;
                        float x5 = rsq(len2[l]);
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
                        x1s[l].f = x1;
                    }
                }
            }

            if (spec_done && !(props & QR_PROP_METAL))
            {
                /* LT_mtl 3090-3149: "plain" diffuse-specular blending */
                for (int l = 0; l < n; l++)
                {
                    if (FLD(c, TMASK)[l].u == 0) continue;
                    for (int k = 0; k < 3; k++)
                    {
                        float x1 = FLD(c, TEX + k)[l].f;
                        float x4 = lg->col[k];
                        x1 = x1 * x0[l].f;
                        x1 = x1 * x4;
                        x4 = x4 * x1s[l].f;
                        x1 = x1 + x4;
                        x1 = x1 + FLD(c, COL + k)[l].f;
                        FLD(c, COL + k)[l].f = x1;
                    }
                }
            }
            else
            {
                /* LT_spc 3047-3084: "metal" blending (also when specular is off) */
                for (int l = 0; l < n; l++)
                {
                    float d = x0[l].f;
                    if (spec_done) d = d + x1s[l].f;
                    if (FLD(c, TMASK)[l].u == 0) continue;
                    for (int k = 0; k < 3; k++)
                    {
                        float x1 = FLD(c, TEX + k)[l].f;
                        x1 = x1 * lg->col[k];
                        x1 = x1 * d;
                        x1 = x1 + FLD(c, COL + k)[l].f;
                        FLD(c, COL + k)[l].f = x1;
                    }
                }
            }
        }
    }

    /* TRANSPARENCY 3185-3598 */
    W xr[MAXS], xg[MAXS], xb[MAXS];
    {
        for (int l = 0; l < n; l++)
        {
            FLD(c, C_TRN)[l].f = m->c_trn;
            FLD(c, C_RFL)[l].f = m->c_rfl;
            W t = FLD(c, F_PRB)[l];
            FLD(c, TMASK)[l] = t;
            FLD(c, M_TRN)[l] = t;
            FLD(c, M_RFL)[l] = t;
            xr[l].u = xg[l].u = xb[l].u = 0;
        }

        int traced = 0;
        if (!none(FLD(c, TMASK), n) && !(props & QR_PROP_OPAQUE))
        {
            int go = 1;
            const int rfi = (props & QR_PROP_REFRACT) || (props & QR_PROP_FRESNEL);
            W x0[MAXS], x4[MAXS], x6[MAXS], x7[MAXS];

            if (rfi)
            {
                /* TR_rfi 3212-3260 */
                W x1[MAXS], x2[MAXS], x3[MAXS];
                for (int l = 0; l < n; l++)
                {
                    float a1 = FLD(c, RAY + 0)[l].f, a2 = FLD(c, RAY + 1)[l].f, a3 = FLD(c, RAY + 2)[l].f;
                    float s0 = a1 * a1;
                    s0 = s0 + a2 * a2;
                    s0 = s0 + a3 * a3;
                    float inv = rsq(s0);
                    a1 = a1 * inv; a2 = a2 * inv; a3 = a3 * inv;
                    float d = a1 * FLD(c, NRM + 0)[l].f;
                    d = d + a2 * FLD(c, NRM + 1)[l].f;
                    d = d + a3 * FLD(c, NRM + 2)[l].f;
                    x1[l].f = a1; x2[l].f = a2; x3[l].f = a3;
                    x4[l].f = d;
                    x6[l].f = m->c_rfr;
                    float b0 = d * m->c_rfr;
                    float b7 = b0 * b0;
                    b7 = b7 + 1.0f;
                    b7 = b7 - m->rfr_2;
                    x0[l].f = b0; x7[l].f = b7;
                }
                if (props & QR_PROP_FRESNEL)
                {
                    /* 3266-3295: total inner reflection */
                    for (int l = 0; l < n; l++)
                    {
                        uint32_t mk = m_le(0.0f, x7[l].f) & FLD(c, M_TRN)[l].u;
                        FLD(c, M_TRN)[l].u = mk;
                        FLD(c, TMASK)[l].u = mk;
                    }
                    if (none(FLD(c, M_TRN), n))
                    {
                        for (int l = 0; l < n; l++)
                        {
                            FLD(c, C_TRN)[l].u = 0;
                            FLD(c, C_RFL)[l].f = m->c_rfl + m->c_trn;
                        }
                        go = 0;
                    }
                }
                if (go)
                {
                    /* TR_cnt 3297-3347 */
                    for (int l = 0; l < n; l++)
                    {
                        x7[l].f = sqrtf(x7[l].f);
                        x0[l].f = x0[l].f + x7[l].f;
                    }
                    if (props & QR_PROP_REFRACT)
                    {
                        for (int l = 0; l < n; l++)
                        {
                            float x5 = FLD(c, NRM + 0)[l].f * x0[l].f;
                            FLD(c, NEW + 0)[l].f = x1[l].f * x6[l].f - x5;
                            x5 = FLD(c, NRM + 1)[l].f * x0[l].f;
                            FLD(c, NEW + 1)[l].f = x2[l].f * x6[l].f - x5;
                            x5 = FLD(c, NRM + 2)[l].f * x0[l].f;
                            FLD(c, NEW + 2)[l].f = x3[l].f * x6[l].f - x5;
                        }
                    }
                    else
                    {
                        for (int l = 0; l < n; l++)
                            for (int k = 0; k < 3; k++) FLD(c, NEW + k)[l] = FLD(c, RAY + k)[l];
                    }
                }
            }
            else
            {
                /* TR_rfe 3336-3347: propagate ray */
                for (int l = 0; l < n; l++)
                    for (int k = 0; k < 3; k++) FLD(c, NEW + k)[l] = FLD(c, RAY + k)[l];
            }

            if (go && (props & QR_PROP_FRESNEL))
            {
                /* TR_ini 3385-3424: exact dielectric Fresnel */
                for (int l = 0; l < n; l++)
                {
                    float a1 = x4[l].f;
                    float a2 = a1 * x6[l].f;
                    a2 = a2 - x7[l].f;
                    float a7 = x7[l].f * x6[l].f;
                    float a3 = a1;
                    a1 = a1 + a7;
                    a3 = a3 - a7;
                    float a0 = x0[l].f / a2;
                    a1 = a1 / a3;
                    a0 = a0 * a0;
                    a1 = a1 * a1;
                    a0 = a0 + a1;
                    a0 = a0 * -0.5f;
                    uint32_t u0 = f2u(a0) & 0x7FFFFFFFu;
                    uint32_t mk = FLD(c, M_TRN)[l].u;
                    u0 &= mk;
                    a0 = u2f(u0) * m->c_trn;
                    u0 = f2u(a0) | (~mk & f2u(m->c_trn));
                    a0 = u2f(u0);
                    FLD(c, C_TRN)[l].f = m->c_trn - a0;
                    FLD(c, C_RFL)[l].f = m->c_rfl + a0;
                }
                if (r->pt != NULL && !(r->depth > QR_STACK_DEPTH - 2))
                {
                    /* 3428-3466 (RT_FEAT_PT_SPLIT_FRESNEL): below the first two
                     * levels follow ONE of the two rays, chosen with probability
                     * 0.25 + 0.5 * reflectance share, and weigh it up */
                    W x0r[MAXS];
                    get_random(r, c, x0r);
                    for (int l = 0; l < n; l++)
                    {
                        const float a4 = FLD(c, C_TRN)[l].f;
                        float a5 = FLD(c, C_RFL)[l].f;
                        float a6 = a5;
                        float a7 = a4 + a5;
                        a5 = a5 / a7;
                        a7 = 0.5f;
                        a5 = a5 * a7;
                        a7 = a7 * a7;
                        a7 = a7 + a5;
                        const float rn = x0r[l].f;
                        const uint32_t mt = m_ge(rn, a7) & FLD(c, M_TRN)[l].u;
                        FLD(c, M_TRN)[l].u = mt;
                        const uint32_t mr = m_lt(rn, a7) & FLD(c, M_RFL)[l].u;
                        FLD(c, M_RFL)[l].u = mr;
                        a5 = a4;
                        const float a2 = 1.0f - a7;
                        a5 = a5 / a2;
                        a6 = a6 / a7;
                        FLD(c, C_TRN)[l].u = f2u(a5) & mt;
                        FLD(c, C_RFL)[l].u = f2u(a6) & mr;
                    }
                }
            }

            if (go && !none(FLD(c, M_TRN), n))
            {
                /* TR_frn 3472-3552 */
                for (int l = 0; l < n; l++)
                {
                    FLD(c, TMASK)[l] = FLD(c, M_TRN)[l];
                    FLD(c, T_NEW)[l].u = 0;
                }
                if (r->depth != 0)
                {
                    const int cc = c + STEP;
                    lvl_t *C = &r->lv[lvl + 1];
                    r->depth -= 1;
                    C->p_flg = L->l_flg | QR_FLAG_PASS_THRU;
                    C->p_lst = s->mat[side];
                    C->p_obj = si;
                    for (int l = 0; l < n; l++)
                    {
                        FLD(cc, WMASK)[l] = FLD(c, TMASK)[l];
                        FLD(cc, T_BUF)[l].f = r->h->cam_t_max;
                        FLD(cc, C_BUF)[l].u = 0;
                        FLD(cc, COL + 0)[l].u = 0;
                        FLD(cc, COL + 1)[l].u = 0;
                        FLD(cc, COL + 2)[l].u = 0;
                        FLD(cc, T_MIN)[l].u = 0;
                        if (FLD(c, TMASK)[l].u) r->st.rays_refract++;
                    }
                    C->l_flg = 0; C->l_lst = NIL; C->l_obj = NIL;
                    walk(r, lvl + 1, s->lst_srf[side ^ 1]);
                    r->depth += 1;
                    for (int l = 0; l < n; l++)
                    {
                        float t = FLD(c, C_TRN)[l].f;
                        xr[l].f = FLD(cc, COL + 0)[l].f * t;
                        xg[l].f = FLD(cc, COL + 1)[l].f * t;
                        xb[l].f = FLD(cc, COL + 2)[l].f * t;
                    }
                    traced = 1;
                }
            }
        }
        (void)traced;

        /* TR_mix 3564-3598 */
        for (int l = 0; l < n; l++)
        {
            float x0 = 1.0f - m->c_trn;
            x0 = x0 - m->c_rfl;
            x0 = u2f(f2u(x0) & m_le(0.0f, x0));
            float a = FLD(c, COL + 0)[l].f * x0;
            float b = FLD(c, COL + 1)[l].f * x0;
            float d = FLD(c, COL + 2)[l].f * x0;
            a = xr[l].f + a;
            b = xg[l].f + b;
            d = xb[l].f + d;
            FLD(c, TMASK)[l] = FLD(c, F_RND)[l];
            if (FLD(c, TMASK)[l].u == 0) continue;
            FLD(c, COL + 0)[l].f = a;
            FLD(c, COL + 1)[l].f = b;
            FLD(c, COL + 2)[l].f = d;
        }
    }

    /* REFLECTIONS 3604-3930 */
    {
        int go = (props & QR_PROP_REFLECT) != 0;
        if (!go && !(props & QR_PROP_OPAQUE) && (props & QR_PROP_FRESNEL)) go = 1;
        if (go && none(FLD(c, M_RFL), n)) go = 0;

        if (go)
        {
            W x0[MAXS];
            for (int l = 0; l < n; l++)
            {
                FLD(c, TMASK)[l] = FLD(c, M_RFL)[l];
                float a1 = FLD(c, RAY + 0)[l].f, a2 = FLD(c, RAY + 1)[l].f, a3 = FLD(c, RAY + 2)[l].f;
                float a4 = FLD(c, NRM + 0)[l].f, a5 = FLD(c, NRM + 1)[l].f, a6 = FLD(c, NRM + 2)[l].f;
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
                FLD(c, NEW + 0)[l].f = a1;
                FLD(c, NEW + 1)[l].f = a2;
                FLD(c, NEW + 2)[l].f = a3;
                x0[l].f = d;
            }

            if ((props & QR_PROP_FRESNEL) && (props & QR_PROP_OPAQUE))
            {
                for (int l = 0; l < n; l++)
                {
                    float a0 = x0[l].f;
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
                    FLD(c, C_RFL)[l].f = m->c_rfl + a0;
                }
            }

            /* RF_frn 3819-3884 */
            for (int l = 0; l < n; l++)
            {
                FLD(c, T_NEW)[l].u = 0;
                xr[l].u = xg[l].u = xb[l].u = 0;
            }
            if (r->depth != 0)
            {
                const int cc = c + STEP;
                lvl_t *C = &r->lv[lvl + 1];
                r->depth -= 1;
                C->p_flg = L->l_flg | QR_FLAG_PASS_BACK;
                C->p_lst = s->mat[side];
                C->p_obj = si;
                for (int l = 0; l < n; l++)
                {
                    FLD(cc, WMASK)[l] = FLD(c, TMASK)[l];
                    FLD(cc, T_BUF)[l].f = r->h->cam_t_max;
                    FLD(cc, C_BUF)[l].u = 0;
                    FLD(cc, COL + 0)[l].u = 0;
                    FLD(cc, COL + 1)[l].u = 0;
                    FLD(cc, COL + 2)[l].u = 0;
                    FLD(cc, T_MIN)[l].u = 0;
                    if (FLD(c, TMASK)[l].u) r->st.rays_reflect++;
                }
                C->l_flg = 0; C->l_lst = NIL; C->l_obj = NIL;
                walk(r, lvl + 1, s->lst_srf[side]);
                r->depth += 1;
                for (int l = 0; l < n; l++)
                {
                    float t = FLD(c, C_RFL)[l].f;
                    xr[l].f = FLD(cc, COL + 0)[l].f * t;
                    xg[l].f = FLD(cc, COL + 1)[l].f * t;
                    xb[l].f = FLD(cc, COL + 2)[l].f * t;
                }
            }
            /* RF_mix 3888-3908 */
            for (int l = 0; l < n; l++)
            {
                float a = xr[l].f + FLD(c, COL + 0)[l].f;
                float b = xg[l].f + FLD(c, COL + 1)[l].f;
                float d = xb[l].f + FLD(c, COL + 2)[l].f;
                FLD(c, TMASK)[l] = FLD(c, F_RND)[l];
                if (FLD(c, TMASK)[l].u == 0) continue;
                FLD(c, COL + 0)[l].f = a;
                FLD(c, COL + 1)[l].f = b;
                FLD(c, COL + 2)[l].f = d;
            }
        }
    }

    return 0;
}

/* material redirect, QD_mtr 4826-4842 */
static int material_redirect(R *r, int lvl, int ei)
{
    const qr_surface *s = &r->surfs[r->elems[ei].simd];
    return material(r, lvl, ei, s->srf_t[1]);
}

/* ---- list walk: OO_cyc 1341 .. OO_out 5142 -------------------------------- */

static void walk(R *r, int lvl, int ei)
{
    const int c = lvl * STEP, n = r->n;
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

        for (int l = 0; l < n; l++) if (FLD(c, WMASK)[l].u) r->st.surf_visits++;

        /* 1352-1373: reuse stored local hit of the previous context */
        if (same)
        {
            const int pc = c - STEP;
            for (int l = 0; l < n; l++)
            {
                W a = FLD(pc, NRM + 3)[l], b = FLD(pc, NRM + 4)[l], d = FLD(pc, NRM + 5)[l];
                FLD(c, DFF + shift + 0)[l] = a;
                FLD(c, DFF + shift + 1)[l] = b;
                FLD(c, DFF + shift + 2)[l] = d;
            }
        }

        int do_ray = 0;

        if (!(s->srf_t[3] < 0) && L->l_obj != NIL)
        {
            /* 1385-1417: transform caching under a trnode */
            if (!same)
            {
                for (int l = 0; l < n; l++)
                {
                    FLD(c, DFF + 3)[l].f = FLD(c, DFF + 0)[l].f - s->pos[0];
                    FLD(c, DFF + 4)[l].f = FLD(c, DFF + 1)[l].f - s->pos[1];
                    FLD(c, DFF + 5)[l].f = FLD(c, DFF + 2)[l].f - s->pos[2];
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
                for (int l = 0; l < n; l++)
                {
                    FLD(c, DFF + 0)[l].f = FLD(c, ORG + 0)[l].f - s->pos[0];
                    FLD(c, DFF + 1)[l].f = FLD(c, ORG + 1)[l].f - s->pos[1];
                    FLD(c, DFF + 2)[l].f = FLD(c, ORG + 2)[l].f - s->pos[2];
                }
                if (s->a_map[3] != 0)
                {
                    const int dst = s->srf_t[3] < 0 ? 0 : 3;
                    for (int l = 0; l < n; l++)
                    {
                        float o4, o5, o6;
                        xform(s, FLD(c, DFF + 0)[l].f, FLD(c, DFF + 1)[l].f, FLD(c, DFF + 2)[l].f, &o4, &o5, &o6);
                        FLD(c, DFF + dst + 0)[l].f = o4;
                        FLD(c, DFF + dst + 1)[l].f = o5;
                        FLD(c, DFF + dst + 2)[l].f = o6;
                    }
                    if (s->srf_t[3] < 0) L->l_obj = e->data_p;
                    do_ray = 1;
                }
            }
            if (do_ray)
            {
                /* OO_ray 1508-1556 */
                for (int l = 0; l < n; l++)
                {
                    float o4, o5, o6;
                    xform(s, FLD(c, RAY + 0)[l].f, FLD(c, RAY + 1)[l].f, FLD(c, RAY + 2)[l].f, &o4, &o5, &o6);
                    FLD(c, RAY + 3)[l].f = o4;
                    FLD(c, RAY + 4)[l].f = o5;
                    FLD(c, RAY + 5)[l].f = o6;
                }
            }
        }

        /* OO_trm 1558-1570 / AR_ptr 3955-4054: bounding volume of an array */
        if (e->data_i == 1)
        {
            W x7[MAXS];
            for (int l = 0; l < n; l++)
            {
                float x1 = FLD(c, RAY + shift + 0)[l].f;
                float x0 = s->sci[0] * x1;
                float x5 = FLD(c, DFF + shift + 0)[l].f;
                float q7 = s->sci[0] * x5;
                float x3 = x1;
                x1 = x1 * x0; x3 = x3 * q7; x5 = x5 * q7;

                float x2 = FLD(c, RAY + shift + 1)[l].f;
                x0 = s->sci[1] * x2;
                float x6 = FLD(c, DFF + shift + 1)[l].f;
                q7 = s->sci[1] * x6;
                float x4 = x2;
                x2 = x2 * x0; x4 = x4 * q7; x6 = x6 * q7;
                x1 = x1 + x2; x3 = x3 + x4; x5 = x5 + x6;

                x2 = FLD(c, RAY + shift + 2)[l].f;
                x0 = s->sci[2] * x2;
                x6 = FLD(c, DFF + shift + 2)[l].f;
                q7 = s->sci[2] * x6;
                x4 = x2;
                x2 = x2 * x0; x4 = x4 * q7; x6 = x6 * q7;
                x1 = x1 + x2; x3 = x3 + x4; x5 = x5 + x6;

                x5 = x5 - s->sci[3];
                x5 = x5 * x1;
                x3 = x3 * x3;
                x3 = x3 - x5;
                x7[l].u = m_le(0.0f, x3) & FLD(c, WMASK)[l].u;
            }
            if (none(x7, n))
            {
                /* AR_skp 4038-4054 */
                ei = e->data_p;
                if (ei == L->l_obj) L->l_obj = NIL;
            }
            continue;
        }

        const int tag = s->srf_t[0];
        if (tag == 0) continue;                     /* trnode element */

        W x1[MAXS], x3[MAXS], x4[MAXS], x6[MAXS], x7[MAXS], x0[MAXS];

        if (tag == 1)
        {
            /* PL_ptr 4062-4136 */
            if (same) continue;
            const int k = s->a_map[2];
            const uint32_t sg = s->a_sgn[2] ? SMASK : 0;
            for (int l = 0; l < n; l++)
            {
                uint32_t dk = (FLD(c, DFF + k)[l].u ^ sg) ^ SMASK;
                float rk = u2f(FLD(c, RAY + k)[l].u ^ sg);
                x7[l].u = m_ne(0.0f, rk) & FLD(c, WMASK)[l].u;
                FLD(c, T_VAL)[l].f = u2f(dk) / rk;
            }
            clip(r, lvl, ei, x7);
            if (none(x7, n)) continue;
            for (int l = 0; l < n; l++)
            {
                FLD(c, XMASK)[l] = x7[l];
                float rk = u2f(FLD(c, RAY + k)[l].u ^ sg);
                x7[l].u &= m_lt(rk, 0.0f);
                FLD(c, TMASK)[l] = x7[l];
            }
            if (!none(x7, n))
            {
                L->l_flg = QR_FLAG_SIDE_OUTER;
                if (material(r, lvl, ei, 1)) return;
            }
            for (int l = 0; l < n; l++) x7[l].u = FLD(c, TMASK)[l].u ^ FLD(c, XMASK)[l].u;
            if (none(x7, n)) continue;
            for (int l = 0; l < n; l++) FLD(c, TMASK)[l] = x7[l];
            L->l_flg = QR_FLAG_SIDE_INNER;
            if (material(r, lvl, ei, 1)) return;
            continue;
        }

        if (tag == 3)
        {
            /* TP_ptr 4216-4277 */
            const int iI = s->a_map[0], iK = s->a_map[2];
            const float sci_i = s->sci[iI - shift], sci_k = s->sci[iK - shift];
            for (int l = 0; l < n; l++)
            {
                float ri = FLD(c, RAY + iI)[l].f, di = FLD(c, DFF + iI)[l].f;
                float rk = FLD(c, RAY + iK)[l].f, dk = FLD(c, DFF + iK)[l].f;
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
                x1[l].f = a1; x4[l].f = a3; x6[l].f = a0; x3[l].f = a5;
            }
        }
        else
        {
            /* QD_ptr 4378-4447 */
            for (int l = 0; l < n; l++)
            {
                float a1 = FLD(c, RAY + shift + 0)[l].f;
                float a0 = s->sci[0] * a1;
                float a5 = FLD(c, DFF + shift + 0)[l].f;
                float a7 = s->sci[0] * a5;
                a7 = a7 - s->scj[0];
                float a3 = a1;
                a1 = a1 * a0;
                a3 = a3 * a7;
                a7 = a7 - s->scj[0];
                a5 = a5 * a7;

                float a2 = FLD(c, RAY + shift + 1)[l].f;
                a0 = s->sci[1] * a2;
                float a6 = FLD(c, DFF + shift + 1)[l].f;
                a7 = s->sci[1] * a6;
                a7 = a7 - s->scj[1];
                float a4 = a2;
                a2 = a2 * a0;
                a4 = a4 * a7;
                a7 = a7 - s->scj[1];
                a6 = a6 * a7;

                a1 = a1 + a2; a3 = a3 + a4; a5 = a5 + a6;

                a2 = FLD(c, RAY + shift + 2)[l].f;
                a0 = s->sci[2] * a2;
                a6 = FLD(c, DFF + shift + 2)[l].f;
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
                x1[l].f = a1; x4[l].f = a4; x6[l].f = a6; x3[l].f = a3;
            }
        }

        /* QD_rts 4449-4547 */
        for (int l = 0; l < n; l++) x7[l].u = m_le(0.0f, x3[l].f) & FLD(c, WMASK)[l].u;
        if (none(x7, n)) continue;

        for (int l = 0; l < n; l++)
        {
            float b = u2f(f2u(x4[l].f) ^ SMASK);
            float d = x3[l].f;
            FLD(c, DMASK)[l].u = m_lt(d, s->d_eps) & x7[l].u;
            uint32_t bs = SMASK & f2u(b);
            float sd = u2f(f2u(sqrtf(d)) ^ bs);
            float bd = b + sd;
            uint32_t m_pos = m_le(0.0f, sd);
            uint32_t m_neg = m_gt(0.0f, sd);
            uint32_t cu = f2u(x6[l].f), bu = f2u(bd), au = f2u(x1[l].f);
            x6[l].u = (cu & m_neg) | (bu & m_pos);          /* t2nmr */
            x4[l].u = (bu & m_neg) | (cu & m_pos);          /* t1nmr */
            x3[l].u = (bu & m_neg) | (au & m_pos);          /* t2dnm */
            x0[l].u = (au & m_pos) | (au & m_neg);          /* a_val */
            x1[l].u = (au & m_neg) | (bu & m_pos);          /* t1dnm */
        }

        /* 4572-4623: root sorting for near-zero determinant */
        L->x_ptr = 0;
        if (!none(FLD(c, DMASK), n))
        {
            L->x_ptr = 1;
            for (int l = 0; l < n; l++)
            {
                FLD(c, AMASK)[l].u = SMASK & x0[l].u;
                uint32_t z1 = m_eq(x4[l].f, 0.0f);
                x1[l].u = ((x1[l].u | z1) ^ z1) | (z1 & f2u(1.0f));
                uint32_t z2 = m_eq(x6[l].f, 0.0f);
                x3[l].u = ((x3[l].u | z2) ^ z2) | (z2 & f2u(1.0f));
                float t1 = x4[l].f / x1[l].f;
                float t2 = x6[l].f / x3[l].f;
                uint32_t k1 = m_ne(x1[l].f, 0.0f);
                uint32_t k2 = m_ne(x3[l].f, 0.0f);
                uint32_t am = FLD(c, AMASK)[l].u;
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
                u2 &= k1; u2 &= k2; u2 &= FLD(c, DMASK)[l].u;
                t1 = t1 + u2f(u2);
                t2 = t2 - u2f(u2);
                x4[l].f = t1; x6[l].f = t2;
                x1[l].u = k1; x3[l].u = k2;
            }
        }

        /* QD_srt 4646-4824: side loop */
        {
            W x5[MAXS];
            L->x_flg = 2;
            L->x_tag = 0;
            for (int l = 0; l < n; l++) x5[l].u = m_gt(0.0f, x0[l].f) & x7[l].u;
            int state;                  /* 1 = rc1, 2 = rc2 */
            if (none(x5, n)) state = 1;
            else
            {
                for (int l = 0; l < n; l++) x5[l].u ^= x7[l].u;
                if (none(x5, n)) state = 2;
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
                        for (int l = 0; l < n; l++)
                        {
                            x4[l].f = x4[l].f / x1[l].f;
                            x1[l].u = m_ne(x1[l].f, 0.0f);
                        }
                    }
                    for (int l = 0; l < n; l++)
                    {
                        FLD(c, XTMP1)[l] = x6[l];
                        FLD(c, XTMP2)[l] = x3[l];
                        FLD(c, XMASK)[l] = x7[l];
                        x7[l].u &= x1[l].u;
                        FLD(c, T_VAL)[l] = x4[l];
                    }
                    L->l_flg = QR_FLAG_SIDE_OUTER;
                    clip(r, lvl, ei, x7);
                    int hit = !none(x7, n);
                    if (hit)
                    {
                        for (int l = 0; l < n; l++) FLD(c, TMASK)[l] = x7[l];
                        if (material_redirect(r, lvl, ei)) { stop = 2; break; }
                        if (L->x_flg == 0) { stop = 1; break; }
                        if (L->x_tag == 0)
                        {
                            for (int l = 0; l < n; l++) x7[l].u = FLD(c, TMASK)[l].u ^ FLD(c, XMASK)[l].u;
                            if (none(x7, n)) { stop = 1; break; }
                        }
                    }
                    /* QD_rs2 4767-4775 */
                    for (int l = 0; l < n; l++)
                    {
                        x6[l] = FLD(c, XTMP1)[l];
                        x3[l] = FLD(c, XTMP2)[l];
                        x7[l] = FLD(c, XMASK)[l];
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
                        for (int l = 0; l < n; l++)
                        {
                            x6[l].f = x6[l].f / x3[l].f;
                            x3[l].u = m_ne(x3[l].f, 0.0f);
                        }
                    }
                    for (int l = 0; l < n; l++)
                    {
                        FLD(c, XTMP1)[l] = x4[l];
                        FLD(c, XTMP2)[l] = x1[l];
                        FLD(c, XMASK)[l] = x7[l];
                        x7[l].u &= x3[l].u;
                        FLD(c, T_VAL)[l] = x6[l];
                    }
                    L->l_flg = QR_FLAG_SIDE_INNER;
                    clip(r, lvl, ei, x7);
                    int hit = !none(x7, n);
                    if (hit)
                    {
                        for (int l = 0; l < n; l++) FLD(c, TMASK)[l] = x7[l];
                        if (material_redirect(r, lvl, ei)) { stop = 2; break; }
                        if (L->x_flg == 0) { stop = 1; break; }
                        if (L->x_tag == 0)
                        {
                            for (int l = 0; l < n; l++) x7[l].u = FLD(c, TMASK)[l].u ^ FLD(c, XMASK)[l].u;
                            if (none(x7, n)) { stop = 1; break; }
                        }
                    }
                    /* QD_rs1 4685-4693 */
                    for (int l = 0; l < n; l++)
                    {
                        x4[l] = FLD(c, XTMP1)[l];
                        x1[l] = FLD(c, XTMP2)[l];
                        x7[l] = FLD(c, XMASK)[l];
                    }
                    if (L->x_flg == 0) { stop = 1; break; }
                    state = 1;
                }
            }
            if (stop == 2) return;
        }
    }
}

/* ---- frame loops: YY_cyc / XX_cyc 1142-1322, epilogue XX_end 5161-5343 ---- */

static int render(const void *blob, size_t bytes, uint32_t *frame,
                  int stride, int packet, float *t_out,
                  int y0, int y1, qr_oracle_stats *stats, qr_oracle_pt *pt);

/* rt_Scene::reset_pseed, engine.cpp:3651-3685 (RT_PRNG != LCG48: a 48-bit LCG
 * seeds the plane, low 32 bits kept) */
void qr_oracle_pt_seed(uint32_t *pseed, size_t n)
{
    uint64_t seed = 1;
    for (size_t k = 0; k < n; k++)
    {
        seed = (seed * 25214903917ull + 11ull) & 0x0000FFFFFFFFFFFFull;
        pseed[k] = (uint32_t)seed;
    }
}

int qr_oracle_render(const void *blob, size_t bytes, uint32_t *frame,
                     int stride, int packet, float *t_out,
                     int y0, int y1, qr_oracle_stats *stats)
{
    return render(blob, bytes, frame, stride, packet, t_out, y0, y1, stats, NULL);
}

int qr_oracle_render_pt(const void *blob, size_t bytes, uint32_t *frame,
                        int stride, int packet, int y0, int y1, qr_oracle_pt *pt)
{
    if (pt == NULL || pt->pseed == NULL || pt->ptr_r == NULL || pt->ptr_g == NULL || pt->ptr_b == NULL) return -6;
    return render(blob, bytes, frame, stride, packet, NULL, y0, y1, NULL, pt);
}

static int render(const void *blob, size_t bytes, uint32_t *frame,
                  int stride, int packet, float *t_out,
                  int y0, int y1, qr_oracle_stats *stats, qr_oracle_pt *pt)
{
    const qr_blob_header *h = (const qr_blob_header *)blob;
    if (bytes < sizeof(*h) || h->magic != QR_BLOB_MAGIC) return -1;
    if (h->version != QR_BLOB_VERSION || h->total_bytes > bytes) return -2;
    if (packet != 1 && (packet < 4 || packet > MAXS || (packet & 3))) return -3;
    if (h->fsaa < 0 || h->fsaa > 2) return -4;

    R *r = (R *)calloc(1, sizeof(R));
    if (r == NULL) return -5;

    const uint8_t *b = (const uint8_t *)blob;
    r->h      = h;
    r->surfs  = (const qr_surface  *)(b + h->off_surf);
    r->mats   = (const qr_material *)(b + h->off_mat);
    r->lgts   = (const qr_light    *)(b + h->off_lgt);
    r->elems  = (const qr_elem     *)(b + h->off_elem);
    r->tiles  = (const int32_t     *)(b + h->off_tiles);
    r->texels = (const uint32_t    *)(b + h->off_texels);

    const int fsaa = h->fsaa;
    const int spp = 1 << fsaa;
    const int G = packet < 4 ? 4 : packet;      /* lanes per epilogue group */
    const int gpix = G >> fsaa;                 /* pixels per group */
    static const int lane_px[3][4] = { {0, 1, 2, 3}, {0, 0, 1, 1}, {0, 0, 0, 0} };

    if (y0 < 0) y0 = 0;
    if (y1 > h->y_res) y1 = h->y_res;

    r->pt = pt;
    if (pt != NULL)
    {
        /* 1112-1124: one more sample per pixel sample; its weight, the rest's */
        pt->pts_c = pt->pts_c + 1.0f;
        r->pts_o = 1.0f / pt->pts_c;
        r->pts_u = 1.0f - r->pts_o;
    }

    for (int y = y0; y < y1; y++)
    {
        for (int x = 0; x < h->x_res; x += gpix)
        {
            float colr[MAXS], colg[MAXS], colb[MAXS], tb[MAXS];
            /* 1168-1174, 5182: the packet's slots in the seed / colour planes,
             * (y * x_row + x) << fsaa + lane */
            const size_t slot0 = ((size_t)y * (size_t)h->x_row + (size_t)x) << fsaa;

            for (int g0 = 0; g0 < G; g0 += packet)
            {
                const int n = packet;
                r->n = n;
                r->depth = h->depth;
                lvl_t *L = &r->lv[0];
                L->p_flg = (int)h->ctx_flags;
                L->p_lst = NIL; L->p_obj = NIL;
                L->l_flg = 0; L->l_lst = NIL; L->l_obj = NIL;

                int px0 = 0;
                W jx[MAXS], jy[MAXS];
                for (int l = 0; l < n; l++) jx[l].u = jy[l].u = 0;
                if (pt != NULL)
                {
                    /* 1218-1285 (RT_FEAT_PT_RANDOM_SAMPLE): tent-filtered jitter
                     * of the sample position, all lanes draw */
                    r->prngs = slot0 + (size_t)g0;
                    for (int l = 0; l < n; l++) FLD(0, TMASK)[l].u = ONES;
                    for (int pass = 0; pass < 2; pass++)
                    {
                        W x0[MAXS];
                        get_random(r, 0, x0);
                        for (int l = 0; l < n; l++)
                        {
                            float a0 = x0[l].f + x0[l].f;
                            const float a2 = a0;
                            const uint32_t lt = m_lt(a0, 1.0f);
                            float a3 = sqrtf(a2);
                            a3 = a3 - 1.0f;
                            float a5 = 2.0f - a2;
                            a5 = sqrtf(a5);
                            const float b2 = 1.0f - a5;
                            const uint32_t u = (f2u(a3) & lt) | (~lt & f2u(b2));
                            (pass == 0 ? jx : jy)[l].u = u;
                        }
                    }
                    for (int l = 0; l < n; l++)
                    {
                        jx[l].f = jx[l].f * 0.5f;
                        jy[l].f = jy[l].f * 0.5f;
                        if (fsaa != 0)
                        {
                            jx[l].f = jx[l].f * 0.5f;
                            jy[l].f = jy[l].f * 0.5f;
                        }
                    }
                }
                for (int l = 0; l < n; l++)
                {
                    const int gl = g0 + l;          /* lane within the group */
                    const int px = x + (gl >> 2) * (4 >> fsaa) + lane_px[fsaa][gl & 3];
                    if (l == 0) px0 = px;
                    /* 1287-1322: ray init; hor_i/ver_i are exact integers
                     * (engine.cpp:3613-3624), the PT jitter terms are zero
                     * unless the path tracer is on */
                    float hs = (float)px + h->hor_a[gl & 3];
                    float vs = (float)y + h->ver_a[gl & 3];
                    hs = hs + jx[l].f;
                    vs = vs + jy[l].f;
                    for (int k = 0; k < 3; k++)
                    {
                        float a = h->hor[k] * hs;
                        float bb = h->ver[k] * vs;
                        a = a + bb;
                        a = a + h->dir[k];
                        FLD(0, RAY + k)[l].f = a;
                        FLD(0, ORG + k)[l].f = h->org[k];
                        FLD(0, COL + k)[l].u = 0;
                    }
                    FLD(0, T_MIN)[l].f = h->t_min;
                    FLD(0, WMASK)[l].u = ONES;
                    FLD(0, T_BUF)[l].f = h->cam_t_max;
                    FLD(0, C_BUF)[l].u = 0;
                    r->st.rays_primary++;
                }
                /* 1328-1333: tile of the packet's first pixel */
                int tx = px0 / h->tile_w;
                if (tx >= h->tls_row) tx = h->tls_row - 1;
                const int head = r->tiles[(y / h->tile_h) * h->tls_row + tx];
                walk(r, 0, head);

                for (int l = 0; l < n; l++)
                {
                    colr[g0 + l] = FLD(0, COL + 0)[l].f;
                    colg[g0 + l] = FLD(0, COL + 1)[l].f;
                    colb[g0 + l] = FLD(0, COL + 2)[l].f;
                    tb[g0 + l]   = FLD(0, T_BUF)[l].f;
                }
            }

            if (pt != NULL)
            {
                /* 5176-5219: running mean over the frames since set_pton */
                float *acc[3] = { pt->ptr_r, pt->ptr_g, pt->ptr_b };
                float *col[3] = { colr, colg, colb };
                for (int k = 0; k < 3; k++)
                {
                    for (int l = 0; l < G; l++)
                    {
                        float a0 = col[k][l] * r->pts_o;
                        const float a1 = acc[k][slot0 + (size_t)l] * r->pts_u;
                        a0 = a0 + a1;
                        acc[k][slot0 + (size_t)l] = a0;
                        col[k][l] = a0;
                    }
                }
            }

            /* XX_end 5221-5234: clamp (minps returns the source on NaN) */
            float *cc[3] = { colr, colg, colb };
            for (int k = 0; k < 3; k++)
                for (int l = 0; l < G; l++)
                    cc[k][l] = cc[k][l] < 1.0f ? cc[k][l] : 1.0f;

            /* AA_cyc 5241-5308: halve, then add adjacent pairs, per pass */
            int cnt = G;
            for (int p = 0; p < fsaa; p++)
            {
                for (int k = 0; k < 3; k++)
                {
                    for (int l = 0; l < cnt; l++) cc[k][l] = cc[k][l] * 0.5f;
                    for (int l = 0; l < cnt / 2; l++) cc[k][l] = cc[k][2 * l] + cc[k][2 * l + 1];
                }
                cnt >>= 1;
            }

            /* FRAME_SIMD 988-1006 */
            for (int l = 0; l < gpix; l++)
            {
                uint32_t pix = 0;
                for (int k = 0; k < 3; k++)
                {
                    float v = cc[k][l];
                    if (h->ctx_flags & QR_PROP_GAMMA) v = sqrtf(v);
                    v = v * h->cam_clamp;
                    uint32_t iv = (uint32_t)cvn(v) & h->cam_cmask;
                    pix |= iv << (k == 0 ? 16 : k == 1 ? 8 : 0);
                }
                if (x + l < h->x_res && frame != NULL)
                {
                    frame[(size_t)y * stride + x + l] = pix;
                }
            }

            if (t_out != NULL)
            {
                for (int gl = 0; gl < G; gl++)
                {
                    const int px = x + (gl >> 2) * (4 >> fsaa) + lane_px[fsaa][gl & 3];
                    const int sm = fsaa == 0 ? 0 : fsaa == 1 ? (gl & 1) : (gl & 3);
                    if (px < h->x_res)
                    {
                        t_out[((size_t)y * h->x_res + px) * spp + sm] = tb[gl];
                    }
                }
            }
        }
    }

    if (stats != NULL) *stats = r->st;
    free(r);
    return 0;
}
