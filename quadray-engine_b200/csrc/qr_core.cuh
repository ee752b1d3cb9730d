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
#include "qr_kscene.h"

#if defined(__CUDACC__)
#define QR_HD __host__ __device__ __forceinline__
/* paths that are rare in most scenes live out of line, so that the hot list
 * walk stays compact in the instruction cache */
#define QR_HD_COLD __host__ __device__ __noinline__
#else
#include <math.h>
#include <string.h>
#define QR_HD static inline
#define QR_HD_COLD static
#endif

/* ---- checked build --------------------------------------------------------- */
/*
 * -DQR_CHECKED (make checked -> lib/libquadray_b200_checked.so): the kernel
 * checks its own accesses and counts violations instead of faulting -- what
 * stands in for compute-sanitizer, which the GPU pool does not allow
 * (profiles/r02d_compute_sanitizer.txt).  tests/test_gpu_checked_build.py runs
 * fixtures through it and expects all counters at zero.
 *   0 element cursor outside the image's element array (incl. device-built tile lists)
 *   1 surface record offset outside the surface table
 *   2 best-hit record read without having been written by the walk that reported a hit
 *   3 continuation stack deeper than the scene's depth
 *   4 pixel store outside the frame
 *   5 tile index outside the tile table
 *   6 shading state read back from the scratch that was not parked by this thread
 */
#define QR_CHECK_COUNTERS 8
#if defined(QR_CHECKED) && defined(__CUDACC__)
/* (single translation unit: defined here) */
struct qr_check_limits_t { unsigned long long elems_lo, elems_hi; uint32_t surf_bytes, n_tiles; };
static __constant__ qr_check_limits_t qr_check_limits;
static __device__ unsigned int qr_check_count[QR_CHECK_COUNTERS];
#endif
#if defined(QR_CHECKED) && defined(__CUDA_ARCH__)
#define QR_CHECK(cond, code) do { if (!(cond)) atomicAdd(&qr_check_count[code], 1u); } while (0)
#define QR_CHECKED_ONLY(x) x
#else
#define QR_CHECK(cond, code) ((void)0)
#define QR_CHECKED_ONLY(x)
#endif
#define QR_POISON 0x7FC0DEADu   /* a NaN no computation here produces */

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

/* ---- scene view (packed image, qr_kscene.h) -------------------------------- */

/*
 * SH = true : the kscene prefix sits in shared memory; bases are 32-bit
 *             shared-window addresses and quads are read with ld.shared.v4
 *             (one LDS.128, address = register + immediate)
 * SH = false: bases are byte pointers (global memory on the device, plain
 *             memory in the host build)
 */
template <bool SH> struct qr_hot;

template <> struct qr_hot<false>
{
    typedef const uint8_t *base_t;
#if defined(__CUDACC__)
    static __host__ __device__ __forceinline__ qr_f4 ld(base_t b, uint32_t off)
#else
    static inline qr_f4 ld(base_t b, uint32_t off)
#endif
    {
        return *(const qr_f4 *)(b + off);
    }
};

#if defined(__CUDACC__)
template <> struct qr_hot<true>
{
    typedef uint32_t base_t;
    static __device__ __forceinline__ qr_f4 ld(base_t b, uint32_t off)
    {
        qr_f4 r;
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
            : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(b + off));
        return r;
    }
};
#endif

/*
 * Per-thread scratch: a few 16-byte quads in shared memory (device) that take
 * values which must survive a list walk but are not needed inside it, so they
 * do not occupy registers in the hot loop:
 *   BEST  best hit of the walk (surface | side, local hit point): written by
 *         the walk when a nearer candidate passes clipping, read once by the
 *         shader afterwards
 *   LOC   stored local hit of the level the ray starts from (NRM_I/J/K of the
 *         previous context, tracer.cpp:1352-1373): read by the walk only when
 *         it meets the surface the ray left
 *   COL RAY NRM TEX   shading state of a level while its shadow ray is walked
 *         (COL shares its quad with BEST: the best-hit record is consumed by
 *         the shader before it parks anything, and shadow walks write no
 *         BEST)
 *   MISC  primary T_BUF and the ray counters (shadow, reflection, refraction)
 *   WORG WRAY  the ray in world space while the walk works on a copy that it
 *         transforms in place (open transform node / surface with a matrix);
 *         read back when a node closes, by the clipper for the world hit
 *         point inside a node, and by the shader after the walk
 * Quad q of thread t sits at base + q * stride + t * 16: a warp's 128-bit
 * access covers 512 contiguous bytes (no bank conflicts).
 */
#define QR_SC_BEST   0      /* closest-hit walks only ...                                   */
#define QR_SC_COL    0      /* ... shadow walks only: the two never hold live data together */
#define QR_SC_LOC    1
#define QR_SC_RAY    2
#define QR_SC_NRM    3
#define QR_SC_TEX    4
#define QR_SC_MISC   5
#define QR_SC_WORG   6      /* world origin / direction of the ray being walked: parked for */
#define QR_SC_WRAY   7      /* the walk, which transforms its copy in place inside nodes */
#define QR_SC_QUADS  8

#if defined(__CUDACC__)
/* (the host pass of nvcc parses these too; it never calls them) */
struct qr_scratch { uint32_t addr, stride; };   /* shared-window address of quad 0, bytes between quads */
QR_HD void qr_sc_st(const qr_scratch s, uint32_t q, float a, float b, float c, float d)
{
#if defined(__CUDA_ARCH__)
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "r"(s.addr + q * s.stride), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
#endif
}
QR_HD qr_f4 qr_sc_ld(const qr_scratch s, uint32_t q)
{
    qr_f4 r = {0.0f, 0.0f, 0.0f, 0.0f};
#if defined(__CUDA_ARCH__)
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(s.addr + q * s.stride) : "memory");
#endif
    return r;
}
QR_HD void qr_sc_st1(const qr_scratch s, uint32_t q, uint32_t w, uint32_t val)
{
#if defined(__CUDA_ARCH__)
    asm volatile("st.shared.b32 [%0], %1;" :: "r"(s.addr + q * s.stride + w * 4u), "r"(val) : "memory");
#endif
}
QR_HD uint32_t qr_sc_ld1(const qr_scratch s, uint32_t q, uint32_t w)
{
    uint32_t val = 0;
#if defined(__CUDA_ARCH__)
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(val) : "r"(s.addr + q * s.stride + w * 4u) : "memory");
#endif
    return val;
}
#else
struct qr_scratch { qr_f4 *quads; };
QR_HD void qr_sc_st(const qr_scratch s, uint32_t q, float a, float b, float c, float d)
{
    s.quads[q].x = a; s.quads[q].y = b; s.quads[q].z = c; s.quads[q].w = d;
}
QR_HD qr_f4 qr_sc_ld(const qr_scratch s, uint32_t q) { return s.quads[q]; }
QR_HD void qr_sc_st1(const qr_scratch s, uint32_t q, uint32_t w, uint32_t val)
{
    memcpy((float *)&s.quads[q] + w, &val, 4);
}
QR_HD uint32_t qr_sc_ld1(const qr_scratch s, uint32_t q, uint32_t w)
{
    uint32_t val; memcpy(&val, (const float *)&s.quads[q] + w, 4); return val;
}
#endif
QR_HD void qr_sc_inc(const qr_scratch s, uint32_t q, uint32_t w)
{
    qr_sc_st1(s, q, w, qr_sc_ld1(s, q, w) + 1u);
}

template <bool SH>
struct qr_view
{
    const qr_blob_header *h;
    typename qr_hot<SH>::base_t surf;   /* QR_KSURF_QUADS per surface */
    typename qr_hot<SH>::base_t shade;  /* QR_KSHADE_QUADS per surface */
    typename qr_hot<SH>::base_t mat;    /* QR_KMAT_QUADS per material */
    typename qr_hot<SH>::base_t lgt;    /* QR_KLGT_QUADS per light */
    const qr_kelem       *elems;
    const int32_t        *tiles;
    const uint32_t       *texels;
};

/* quad "q" of the surface record at byte offset "so" (= surface index * 128) */
#if defined(QR_CHECKED) && defined(__CUDA_ARCH__)
#define QR_SURF(v, so, q)  (((uint32_t)(so) < qr_check_limits.surf_bytes ? (void)0 : (void)atomicAdd(&qr_check_count[1], 1u)), \
                            qr_hot<SH>::ld((v).surf, ((uint32_t)(so) < qr_check_limits.surf_bytes ? (uint32_t)(so) : 0u) + (q) * 16u))
#else
#define QR_SURF(v, so, q)  qr_hot<SH>::ld((v).surf,  (uint32_t)(so) + (q) * 16u)
#endif
/* shading record of the same surface (32 B each) */
#define QR_SHADE(v, so, q) qr_hot<SH>::ld((v).shade, ((uint32_t)(so) >> 2) + (q) * 16u)
#define QR_MAT(v, i, q)    qr_hot<SH>::ld((v).mat,   (uint32_t)(i) * (QR_KMAT_QUADS * 16u) + (q) * 16u)
#define QR_LGT(v, i, q)    qr_hot<SH>::ld((v).lgt,   (uint32_t)(i) * (QR_KLGT_QUADS * 16u) + (q) * 16u)

/* surfaces are named by the byte offset of their record; "no surface": */
#define QR_SO_NIL  0xFFFFFF80u

/* host / global-memory view: "img" is the whole kscene image */
QR_HD void qr_view_init(qr_view<false> &v, const void *img)
{
    const uint8_t *b = (const uint8_t *)img;
    const qr_blob_header *h = (const qr_blob_header *)img;
    v.h      = h;
    v.surf   = b + h->off_surf;
    v.shade  = b + (uint32_t)h->pad3[0];
    v.mat    = b + h->off_mat;
    v.lgt    = b + h->off_lgt;
    v.elems  = (const qr_kelem *)(b + h->off_elem);
    v.tiles  = (const int32_t *)(b + h->off_tiles);
    v.texels = (const uint32_t *)(b + h->off_texels);
}

/* continuation of a level that waits for a child ray */
struct qr_frame
{
    /* a level that waits for its REFLECTION child (the last thing it does)
     * needs only the first five words back: colour so far and C_RFL.  Writing
     * and reading the whole frame made thread-local memory the main user of
     * L1 (the list elements share it): L1 hit rate 72 % -> 95 % */
    float    col[3];         /* COL of the level so far */
    float    c_rfl;          /* ctx_C_RFL */
    int32_t  stage;          /* 0: child is the refraction ray, 1: reflection */
    float    c_trn;          /* ctx_C_TRN */
    uint32_t so;             /* surface being shaded */
    int32_t  flg;            /* ctx_LOCAL(FLG): side | props */
    float    ray[3];         /* RAY_X/Y/Z of the level */
    float    hit[3];         /* HIT_X/Y/Z */
    float    nrm[3];         /* NRM_X/Y/Z */
    float    loc[3];         /* NRM_I/J/K: stored local hit (tracer.cpp:2272-2282) */
};

#define QR_MODE_CLOSEST 0
#define QR_MODE_SHADOW  1

/* register-resident 3-vectors: component select without local memory */
QR_HD float qr_pick3(uint32_t i, float a0, float a1, float a2)
{
    return i == 0 ? a0 : (i == 1 ? a1 : a2);
}

QR_HD void qr_put3(uint32_t i, float v, float &a0, float &a1, float &a2)
{
    if (i == 0) a0 = v; else if (i == 1) a1 = v; else a2 = v;
}

/*
 * 3x3 transform, tracer.cpp:1447-1479 / 1512-1548 / 2063-2095: diagonal
 * products first, then the off-diagonal terms of each row in column order;
 * a_map[L] == 1 keeps the diagonal only (scaling fast path).
 */
QR_HD void qr_xform(const qr_f4 q5, const qr_f4 q6, const float tck_z, uint32_t trm,
                    float v1, float v2, float v3, float &o4, float &o5, float &o6)
{
    float x4 = qr_mul(q5.x, v1);
    float x5 = qr_mul(q6.x, v2);
    float x6 = qr_mul(tck_z, v3);
    if (trm != 1)
    {
        x4 = qr_add(x4, qr_mul(q5.y, v2));
        x4 = qr_add(x4, qr_mul(q5.z, v3));
        x5 = qr_add(x5, qr_mul(q5.w, v1));
        x5 = qr_add(x5, qr_mul(q6.y, v3));
        x6 = qr_add(x6, qr_mul(q6.z, v1));
        x6 = qr_add(x6, qr_mul(q6.w, v2));
    }
    o4 = x4; o5 = x5; o6 = x6;
}

/*
 * Conic singularity solver, tracer.cpp:1706-1856, for one sample whose
 * determinant is near zero (lane semantics: hmask decides): a local hit point
 * closer than t_eps to the apex of a cone / zero hyperboloid / zero
 * hypercylinder is moved onto the surface next to it.
 */
template <bool SH>
QR_HD_COLD qr_f4 qr_conic_fix(const typename qr_hot<SH>::base_t surf, uint32_t so, uint32_t d,
                              float ld0, float ld1, float ld2, uint32_t amask, int side,
                              float lx, float ly, float lz)
{
    struct { typename qr_hot<SH>::base_t surf; } v = { surf };
    const uint32_t conic = QR_D_CONIC(d);
    const uint32_t iI = QR_D_MAP(d, 0), iJ = QR_D_MAP(d, 1), iK = QR_D_MAP(d, 2);
    const qr_f4 q1 = QR_SURF(v, so, 1);
    const float t_eps = QR_SURF(v, so, 7).z;
    const float li = qr_pick3(iI, lx, ly, lz), lj = qr_pick3(iJ, lx, ly, lz), lk = qr_pick3(iK, lx, ly, lz);
    float a0 = qr_mul(li, li);
    if (conic != 2)
    {
        a0 = qr_add(a0, qr_mul(lj, lj));
    }
    a0 = qr_add(a0, qr_mul(lk, lk));
    if (a0 < t_eps)
    {
        const float dfi = qr_pick3(iI, ld0, ld1, ld2), dfj = qr_pick3(iJ, ld0, ld1, ld2), dfk = qr_pick3(iK, ld0, ld1, ld2);
        const float sci_i = qr_pick3(iI, q1.x, q1.y, q1.z), sci_j = qr_pick3(iJ, q1.x, q1.y, q1.z), sci_k = qr_pick3(iK, q1.x, q1.y, q1.z);
        const uint32_t u1 = (qr_f2u(dfi) & 0x80000000u) ^ 0x3F800000u;
        uint32_t u2 = 0;
        float q3 = sci_i;
        float q4 = 1.0f;
        if (conic != 2)
        {
            u2 = (qr_f2u(dfj) & 0x80000000u) ^ 0x3F800000u;
            q3 = qr_add(q3, sci_j);
            q4 = qr_add(q4, 1.0f);
        }
        q3 = qr_div(q3, sci_k);
        q3 = qr_neg(q3);
        float q6 = q3;
        q3 = qr_sqrt(q3);
        q6 = qr_add(q6, q4);
        q4 = qr_rsq(q6);
        q4 = qr_mul(q4, t_eps);
        const float p1 = qr_mul(qr_u2f(u1), q4);
        const float p2 = qr_mul(qr_u2f(u2), q4);
        const float p3 = qr_mul(q3, q4);
        const uint32_t ts = side ? 0x80000000u : 0u;
        uint32_t u3 = qr_f2u(p3) ^ (qr_f2u(dfk) & 0x80000000u);
        u3 ^= (ts & amask) ^ amask;
        const uint32_t tsn = (ts | amask) ^ amask;
        qr_put3(iI, qr_u2f(qr_f2u(p1) ^ tsn), lx, ly, lz);
        if (conic != 2)
        {
            qr_put3(iJ, qr_u2f(qr_f2u(p2) ^ tsn), lx, ly, lz);
        }
        qr_put3(iK, qr_u2f(u3), lx, ly, lz);
    }
    qr_f4 r;
    r.x = lx; r.y = ly; r.z = lz; r.w = 0.0f;
    return r;
}

/*
 * Custom clippers of a surface, tracer.cpp:1931-2151, for one candidate hit:
 * (hx, hy, hz) is the hit point in the world, (lx, ly, lz) in the surface's
 * space, (q0x, q0y, q0z) the surface's position.  Out of line (see
 * QR_HD_COLD); the view's members it needs come by value.
 */
template <bool SH>
QR_HD_COLD bool qr_clip_custom(const typename qr_hot<SH>::base_t surf, const qr_kelem *elems, uint32_t so,
                               float q0x, float q0y, float q0z, float hx, float hy, float hz,
                               float lx, float ly, float lz)
{
    struct { typename qr_hot<SH>::base_t surf; const qr_kelem *elems; } v = { surf, elems };
    struct { float x, y, z; } q0 = { q0x, q0y, q0z };
    /* custom clippers 1931-2151.  The reference evaluates the whole list for
     * the packet; a lone sample may stop as soon as its mask is clear and no
     * accumulator is open (a cleared mask can only come back through an
     * accum enter/leave pair). */
    const uint32_t s_trnode = qr_f2u(QR_SURF(v, so, 4).w);
    const uint32_t c_def = qr_f2u(QR_SURF(v, so, 7).w);
    float nx = 0.0f, ny = 0.0f, nz = 0.0f;      /* NRM_X/Y/Z */
    float ni = 0.0f, nj = 0.0f, nk = 0.0f;      /* NRM_I/J/K */
    bool  m = true, acc = false, in_acc = false;
    int32_t redx = QR_NIL;
    bool  last = true;

    for (uint32_t di = qr_f2u(QR_SURF(v, so, 3).w); ; di++)
    {
        const qr_kelem ce = v.elems[di];
        if (ce.w == QR_KEND) break;
        if (!m && !in_acc)
        {
            return false;
        }

        if (ce.w & QR_KC_ACCUM)
        {
            if (ce.w & QR_KC_NEG)
            {
                acc = m;
                m = c_def != 0;
                in_acc = true;
            }
            else
            {
                m = !m && acc;                  /* annpx: ~mask & C_ACC */
                in_acc = false;
            }
            continue;
        }

        const uint32_t co = QR_K_SURF_OFF(ce.w);
        const qr_f4 c0 = QR_SURF(v, co, 0);
        const uint32_t cd = qr_f2u(c0.w);
        bool have_local = false;

        if (!(cd & QR_D_ARRAY_MASK))
        {
            if (redx != QR_NIL)
            {
                ni = qr_sub(nx, c0.x);
                nj = qr_sub(ny, c0.y);
                nk = qr_sub(nz, c0.z);
                if ((int32_t)di == redx) redx = QR_NIL;
                have_local = true;
            }
        }
        else
        if (co == s_trnode)
        {
            nx = qr_add(lx, q0.x);
            ny = qr_add(ly, q0.y);
            nz = qr_add(lz, q0.z);
            redx = ce.aux;
            continue;
        }

        if (!have_local)
        {
            nx = qr_sub(hx, c0.x);
            ny = qr_sub(hy, c0.y);
            nz = qr_sub(hz, c0.z);
            if (cd & QR_D_TRM_MASK)
            {
                float o4, o5, o6;
                qr_xform(QR_SURF(v, co, 5), QR_SURF(v, co, 6), QR_SURF(v, co, 7).x, QR_D_TRM(cd),
                         nx, ny, nz, o4, o5, o6);
                if (cd & QR_D_ARRAY_MASK)
                {
                    nx = o4; ny = o5; nz = o6;
                    redx = ce.aux;
                    continue;
                }
                ni = o4; nj = o5; nk = o6;
            }
        }

        /* clipper evaluators: PL_clp 4198-4208, QD_clp 4910-4951, TP_clp 4341-4370 */
        const bool csh = (cd & QR_D_SHIFT_MASK) != 0;
        const float p0 = csh ? ni : nx, p1 = csh ? nj : ny, p2 = csh ? nk : nz;
        const uint32_t ctag = QR_D_CLIP(cd);
        if (ctag != 0)
        {
            float val;
            if (ctag == 1)
            {
                val = qr_sgn(qr_pick3(QR_D_MAP(cd, 2), p0, p1, p2), QR_D_SGN(cd, 2));
            }
            else
            {
                const qr_f4 c1 = QR_SURF(v, co, 1);
                float a4 = qr_mul(qr_mul(p0, p0), c1.x);
                float a5 = qr_mul(qr_mul(p1, p1), c1.y);
                float a6 = qr_mul(qr_mul(p2, p2), c1.z);
                if (ctag == 2)
                {
                    const qr_f4 c2 = QR_SURF(v, co, 2);
                    a4 = qr_sub(a4, qr_mul(qr_add(c2.x, c2.x), p0));
                    a5 = qr_sub(a5, qr_mul(qr_add(c2.y, c2.y), p1));
                    a6 = qr_sub(a6, qr_mul(qr_add(c2.z, c2.z), p2));
                }
                a4 = qr_sub(a4, c1.w);
                a4 = qr_add(a4, a5);
                val = qr_add(a4, a6);
            }
            /* APPLY_CLIP 488-496 */
            last = (ce.w & QR_KC_NEG) ? qr_ge(val, 0.0f) : (val <= 0.0f);
        }
        m = m && last;
    }
    return m;
}

/*
 * CC_clp, tracer.cpp:1597-2160, for one candidate root "t" of the surface at
 * "so" (descriptor "d", first quad "q0") that already passed the depth tests
 * t_buf > t and t_min < t (1602-1610).  (lr, ld) are the ray / diff in the
 * surface's field set (world or trnode space).  On success lx/ly/lz hold the
 * (possibly adjusted) local hit point.
 */
template <bool SH, typename V>
QR_HD bool qr_clip(const V &v, uint32_t so, uint32_t d, const qr_f4 q0,
                   const qr_scratch sc, bool xf, float bo0, float bo1, float bo2,
                   float lr0, float lr1, float lr2, float ld0, float ld1, float ld2,
                   float t, bool dmask, uint32_t amask, int side,
                   float &lx, float &ly, float &lz)
{
    /* the world ray: the current frame itself, or parked in the scratch while
     * the walk is inside a transform node / a surface's own matrix */
    const qr_f4 wo = qr_sc_ld(sc, QR_SC_WORG), wr = qr_sc_ld(sc, QR_SC_WRAY);
    const float ox = wo.x, oy = wo.y, oz = wo.z, rx = wr.x, ry = wr.y, rz = wr.z;
    const float hx = qr_add(qr_mul(rx, t), ox);
    const float hy = qr_add(qr_mul(ry, t), oy);
    const float hz = qr_add(qr_mul(rz, t), oz);

    if (d & QR_D_TRM_MASK)
    {
        lx = qr_add(qr_mul(lr0, t), ld0);
        ly = qr_add(qr_mul(lr1, t), ld1);
        lz = qr_add(qr_mul(lr2, t), ld2);
    }
    else
    {
        lx = qr_sub(hx, q0.x);
        ly = qr_sub(hy, q0.y);
        lz = qr_sub(hz, q0.z);
    }

    /* conic singularity solver 1706-1856 (lane semantics: hmask decides) */
    if ((d & QR_D_CONIC_MASK) && dmask)
    {
        const qr_f4 f = qr_conic_fix<SH>(v.surf, so, d, ld0, ld1, ld2, amask, side, lx, ly, lz);
        lx = f.x; ly = f.y; lz = f.z;
    }

    /* axis min/max 1874-1927; an axis that is switched off holds -inf / +inf
     * (qr_kscene.h), so all six compares run unconditionally */
    if (d & QR_D_MM_MASK)
    {
        const qr_f4 q3 = QR_SURF(v, so, 3), q4 = QR_SURF(v, so, 4);
        const bool m = (q3.x <= lx) && (q3.y <= ly) && (q3.z <= lz)
                    && qr_ge(q4.x, lx) && qr_ge(q4.y, ly) && qr_ge(q4.z, lz);
        if (!m) return false;
    }

    if (!(d & QR_D_HASCLIP_MASK))
    {
        return true;
    }

    return qr_clip_custom<SH>(v.surf, v.elems, so, q0.x, q0.y, q0.z, hx, hy, hz, lx, ly, lz);
}

/*
 * Shadow applicability of a hit, CHECK_SHAD 549-589: light-emitting and
 * fully-transparent non-refractive surfaces do not cast shadows.
 */
QR_HD bool qr_casts_shadow(uint32_t props)
{
    if (props & QR_PROP_LIGHT) return false;
    if ((props & QR_PROP_TRANSP) && !(props & QR_PROP_REFRACT)) return false;
    return true;
}

QR_HD uint32_t qr_side_props(uint32_t packed, int side)
{
    return side ? (packed >> 16) : (packed & 0xFFFFu);
}

/*
 * Cursor into the element array.  On the device the array never straddles a
 * 4 GB boundary (qr_scene_upload sees to that), so the upper address word is a
 * constant and stepping / skipping is ONE 32-bit add on the lower word.
 */
#if defined(__CUDA_ARCH__)
struct qr_ecur { unsigned long long p; };
QR_HD qr_ecur qr_e_at(const qr_kelem *base, uint32_t idx)
{
    qr_ecur c;
    c.p = (unsigned long long)(base + idx);
    return c;
}
QR_HD qr_kelem qr_e_ld(const qr_ecur c)
{
    qr_kelem e;
#if defined(QR_CHECKED) && defined(__CUDA_ARCH__)
    if (c.p < qr_check_limits.elems_lo || c.p + sizeof(qr_kelem) > qr_check_limits.elems_hi || (c.p & 7) != 0)
    {
        atomicAdd(&qr_check_count[0], 1u);
        e.w = QR_KEND; e.aux = 0;
        return e;
    }
    /* (device-built tile lists are written by the launch before: no .nc) */
    asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(e.w), "=r"(e.aux) : "l"(c.p));
    return e;
#endif
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(e.w), "=r"(e.aux) : "l"(c.p));
    return e;
}
QR_HD void qr_e_skip(qr_ecur &c, int32_t bytes)
{
    /* 32-bit add on the lower word of the 64-bit register pair */
    asm("{\n\t.reg .b32 lo, hi;\n\t"
        "mov.b64 {lo, hi}, %0;\n\t"
        "add.u32 lo, lo, %1;\n\t"
        "mov.b64 %0, {lo, hi};\n\t}"
        : "+l"(c.p) : "r"(bytes));
}
#else
struct qr_ecur { const uint8_t *p; };
QR_HD qr_ecur qr_e_at(const qr_kelem *base, uint32_t idx) { qr_ecur c; c.p = (const uint8_t *)(base + idx); return c; }
QR_HD qr_kelem qr_e_ld(const qr_ecur c) { return *(const qr_kelem *)c.p; }
QR_HD void qr_e_skip(qr_ecur &c, int32_t bytes) { c.p += bytes; }
#endif

/*
 * One candidate root "t" of the surface at "so" on "side": depth tests
 * (CC_clp 1602-1610), clipping, then the hit is taken.  Returns true when the
 * surface is done (hit taken, or a shadow walk met a transparent occluder).
 *   closest walk: t_buf = t, best-hit record into the scratch
 *   shadow walk : an occluder sets t_buf = -inf and parks the cursor in front
 *                 of element 0 (an END), so the walk ends at its next step
 */
template <bool SH, typename V>
QR_HD bool qr_candidate(const V &v, const qr_scratch sc, uint32_t w, uint32_t so, int mode,
                        float bo0, float bo1, float bo2, float cr0, float cr1, float cr2,
                        float ld0, float ld1, float ld2, bool have_ld,
                        float t, bool dmask, uint32_t amask, int side,
                        float t_min, float &t_buf, qr_ecur &cur)
{
    if (!qr_gt(t_buf, t)) return false;
    if (!(t_min < t)) return false;

    /* the record's first quad is read here rather than kept alive through
     * the solver */
    const qr_f4 q0 = QR_SURF(v, so, 0);
    const uint32_t d = qr_f2u(q0.w);
    if (!have_ld && (d & QR_D_TRM_MASK))
    {
        /* planes X / Y / Z solve with one coordinate; inside a node the
         * clipper wants the whole local origin */
        ld0 = qr_sub(bo0, q0.x); ld1 = qr_sub(bo1, q0.y); ld2 = qr_sub(bo2, q0.z);
    }
    float lx, ly, lz;
    if (!qr_clip<SH>(v, so, d, q0, sc, (w & QR_KF_NODE) != 0, bo0, bo1, bo2, cr0, cr1, cr2,
                     ld0, ld1, ld2, t, dmask, amask, side, lx, ly, lz)) return false;
    if (mode == QR_MODE_SHADOW)
    {
        if (qr_casts_shadow(qr_side_props(qr_f2u(QR_SURF(v, so, 2).w), side)))
        {
            t_buf = qr_u2f(0xFF800000u);
            cur = qr_e_at(v.elems, 0);
            qr_e_skip(cur, -(int32_t)sizeof(qr_kelem));
        }
        return true;
    }
    t_buf = t;
    qr_sc_st(sc, QR_SC_BEST, qr_u2f(so | (uint32_t)side), lx, ly, lz);
    return true;
}

/*
 * One list walk, OO_cyc 1341 .. OO_out 5142, for one sample.  Returns t_buf:
 *   mode CLOSEST: below the t_max it started from when something was hit; the
 *                 scratch quad QR_SC_BEST (surface | side, local hit) is set
 *   mode SHADOW : -inf when the sample is in shadow (first occluder)
 * The scratch quad QR_SC_LOC holds the stored local hit of the originating
 * level (NRM_I/J/K of the previous context), used when the ray starts on the
 * surface tested (p_obj, tracer.cpp:1352-1373).
 *
 * (bo, cr) are the ray origin and direction in the CURRENT frame: the world,
 * or the space of the open transform node (transform caching, tracer.cpp:
 * 1377-1421, 1483-1500), or for one element the space of a surface with its
 * own matrix.  Which elements open / close a frame is compiled into the
 * element stream (qr_kscene.h); the world ray waits in the scratch meanwhile.
 *
 * The loop is written for a small instruction count per element: kinds are
 * tested most frequent first, nothing but the ray, t_buf and the cursor lives
 * across an iteration, the cursor is one 32-bit word, and the next element is
 * loaded where it is known (the other warps of the SM hide the latency).
 */
#if defined(__CUDACC__) && defined(QR_WALK_NOINLINE)
#define QR_WALK_FN __device__ __noinline__
#else
#define QR_WALK_FN QR_HD
#endif
template <bool SH>
QR_WALK_FN float qr_walk(const typename qr_hot<SH>::base_t surf, const qr_kelem *elems, uint32_t head, int mode,
                   float ox, float oy, float oz, float rx, float ry, float rz,
                   float t_min, float t_max, uint32_t p_obj, int p_flg,
                   const qr_scratch sc)
{
    struct { typename qr_hot<SH>::base_t surf; const qr_kelem *elems; } v = { surf, elems };
    qr_sc_st(sc, QR_SC_WORG, ox, oy, oz, 0.0f);
    qr_sc_st(sc, QR_SC_WRAY, rx, ry, rz, 0.0f);
    float bo0 = ox, bo1 = oy, bo2 = oz;
    float cr0 = rx, cr1 = ry, cr2 = rz;
    const int  pf = p_flg & (QR_FLAG_SIDE | QR_FLAG_PASS);
    /* a root t <= 0 (or NaN) can never pass t_min < t when t_min >= 0 */
    const bool no_neg = !(t_min < 0.0f);
    float t_buf = t_max;
    QR_CHECKED_ONLY(if (mode == QR_MODE_CLOSEST) qr_sc_st1(sc, QR_SC_BEST, 0, QR_POISON);)

    qr_ecur cur = qr_e_at(v.elems, head);
    qr_kelem e = qr_e_ld(cur);

    for (;;)
    {
        const uint32_t w = e.w;
        const uint32_t kind = QR_K_KIND(w);
        const uint32_t so = QR_K_SURF_OFF(w);

        if (kind == QR_K_BV)
        {
            /* AR_ptr 3955-4054: bounding volume of an array */
            const qr_f4 q0 = QR_SURF(v, so, 0);
            const qr_f4 q1 = QR_SURF(v, so, 1);
            const float ld0 = qr_sub(bo0, q0.x), ld1 = qr_sub(bo1, q0.y), ld2 = qr_sub(bo2, q0.z);
            float x1 = cr0;
            float x0 = qr_mul(q1.x, x1);
            float x5 = ld0;
            float q7 = qr_mul(q1.x, x5);
            float x3 = x1;
            x1 = qr_mul(x1, x0); x3 = qr_mul(x3, q7); x5 = qr_mul(x5, q7);

            float x2 = cr1;
            x0 = qr_mul(q1.y, x2);
            float x6 = ld1;
            q7 = qr_mul(q1.y, x6);
            float x4 = x2;
            x2 = qr_mul(x2, x0); x4 = qr_mul(x4, q7); x6 = qr_mul(x6, q7);
            x1 = qr_add(x1, x2); x3 = qr_add(x3, x4); x5 = qr_add(x5, x6);

            x2 = cr2;
            x0 = qr_mul(q1.z, x2);
            x6 = ld2;
            q7 = qr_mul(q1.z, x6);
            x4 = x2;
            x2 = qr_mul(x2, x0); x4 = qr_mul(x4, q7); x6 = qr_mul(x6, q7);
            x1 = qr_add(x1, x2); x3 = qr_add(x3, x4); x5 = qr_add(x5, x6);

            x5 = qr_sub(x5, q1.w);
            x5 = qr_mul(x5, x1);
            x3 = qr_mul(x3, x3);
            x3 = qr_sub(x3, x5);
            /* AR_skp: on a miss continue behind the array's last leaf */
            const bool miss = !(0.0f <= x3);
            qr_e_skip(cur, miss ? e.aux : (int32_t)sizeof(qr_kelem));
            e = qr_e_ld(cur);
            continue;
        }

        if (kind <= QR_K_PLANE_G)
        {
            /* the leaves */
            const bool same = (so == p_obj);
            do
            {
                if (kind <= QR_K_PLANE_Z)
                {
                    /* PL_ptr 4062-4136; the element carries the axis K,
                     * a_sgn[K] and pos[K].  The reference divides
                     * (d ^ s ^ sign) by (r ^ s), s = a_sgn[K]: the quotient
                     * and the sign test do not depend on s, the side does */
                    if (same) break;
                    const bool px = kind == QR_K_PLANE_X, py = kind == QR_K_PLANE_Y;
                    const float ok = px ? bo0 : (py ? bo1 : bo2);
                    const float rk = px ? cr0 : (py ? cr1 : cr2);
                    const float dk = qr_sub(ok, qr_u2f((uint32_t)e.aux));
                    if (!(0.0f != rk)) break;
                    if (no_neg && (int32_t)(qr_f2u(dk) ^ qr_f2u(rk)) >= 0) break;
                    const float t = qr_div(qr_neg(dk), rk);
                    const uint32_t sg = (w << (31 - 6)) & 0x80000000u;  /* QR_KF_SGN -> sign bit */
                    const int side = (qr_u2f(qr_f2u(rk) ^ sg) < 0.0f) ? QR_FLAG_SIDE_OUTER : QR_FLAG_SIDE_INNER;
                    qr_candidate<SH>(v, sc, w, so, mode, bo0, bo1, bo2, cr0, cr1, cr2, 0.0f, 0.0f, 0.0f, false,
                                     t, false, 0u, side, t_min, t_buf, cur);
                    break;
                }

                float ld0, ld1, ld2;
                const qr_f4 q0 = QR_SURF(v, so, 0);
                if (same || (w & QR_KF_OWN))
                {
                    /* the rarer prologues */
                    ld0 = ld1 = ld2 = 0.0f;
                    if (w & QR_KF_OWN)
                    {
                        /* OO_dff 1429-1556: surface with its own matrix */
                        const qr_f4 q5t = QR_SURF(v, so, 5), q6 = QR_SURF(v, so, 6);
                        const float tckz = QR_SURF(v, so, 7).x;
                        const uint32_t trm = QR_D_TRM(qr_f2u(q0.w));
                        if (!same)
                        {
                            qr_xform(q5t, q6, tckz, trm, qr_sub(bo0, q0.x), qr_sub(bo1, q0.y), qr_sub(bo2, q0.z),
                                     ld0, ld1, ld2);
                        }
                        float n0, n1, n2;
                        qr_xform(q5t, q6, tckz, trm, cr0, cr1, cr2, n0, n1, n2);
                        cr0 = n0; cr1 = n1; cr2 = n2;
                    }
                    else
                    {
                        ld0 = qr_sub(bo0, q0.x); ld1 = qr_sub(bo1, q0.y); ld2 = qr_sub(bo2, q0.z);
                    }
                    if (same)
                    {
                        /* 1352-1373: secondary ray leaving this very surface
                         * reuses the stored local hit as its local diff */
                        const qr_f4 pl = qr_sc_ld(sc, QR_SC_LOC);
                        ld0 = pl.x; ld1 = pl.y; ld2 = pl.z;
                    }
                    if (kind == QR_K_PLANE_G)
                    {
                        /* PL_ptr with the axis map and sign of the descriptor */
                        if (same) break;
                        const uint32_t d = qr_f2u(q0.w);
                        const uint32_t k = QR_D_MAP(d, 2);
                        const uint32_t sg = (d << (31 - 14)) & 0x80000000u;     /* a_sgn[K] */
                        const float dk = qr_pick3(k, ld0, ld1, ld2);
                        const float rk = qr_pick3(k, cr0, cr1, cr2);
                        if (!(0.0f != rk)) break;
                        if (no_neg && (int32_t)(qr_f2u(dk) ^ qr_f2u(rk)) >= 0) break;
                        const float t = qr_div(qr_neg(dk), rk);
                        const int side = (qr_u2f(qr_f2u(rk) ^ sg) < 0.0f) ? QR_FLAG_SIDE_OUTER : QR_FLAG_SIDE_INNER;
                        qr_candidate<SH>(v, sc, w, so, mode, bo0, bo1, bo2, cr0, cr1, cr2, ld0, ld1, ld2, true,
                                         t, false, 0u, side, t_min, t_buf, cur);
                        break;
                    }
                }
                else
                {
                    ld0 = qr_sub(bo0, q0.x);
                    ld1 = qr_sub(bo1, q0.y);
                    ld2 = qr_sub(bo2, q0.z);
                }

                float a_val, b_val, c_val, d_val;
                const qr_f4 q1 = QR_SURF(v, so, 1);

                if (kind == QR_K_TWOPLANE)
                {
                    /* TP_ptr 4216-4277 */
                    const uint32_t d = qr_f2u(q0.w);
                    const uint32_t iI = QR_D_MAP(d, 0), iK = QR_D_MAP(d, 2);
                    const float sci_i = qr_pick3(iI, q1.x, q1.y, q1.z), sci_k = qr_pick3(iK, q1.x, q1.y, q1.z);
                    const float ri = qr_pick3(iI, cr0, cr1, cr2), di = qr_pick3(iI, ld0, ld1, ld2);
                    const float rk = qr_pick3(iK, cr0, cr1, cr2), dk = qr_pick3(iK, ld0, ld1, ld2);
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
                    const qr_f4 q2 = QR_SURF(v, so, 2);
                    float a7 = qr_sub(qr_mul(q1.x, ld0), q2.x);
                    float a3 = qr_mul(cr0, a7);
                    float a1 = qr_mul(cr0, qr_mul(q1.x, cr0));
                    a7 = qr_sub(a7, q2.x);
                    float a5 = qr_mul(ld0, a7);

                    a7 = qr_sub(qr_mul(q1.y, ld1), q2.y);
                    float a4 = qr_mul(cr1, a7);
                    float a2 = qr_mul(cr1, qr_mul(q1.y, cr1));
                    a7 = qr_sub(a7, q2.y);
                    float a6 = qr_mul(ld1, a7);

                    a1 = qr_add(a1, a2); a3 = qr_add(a3, a4); a5 = qr_add(a5, a6);

                    a7 = qr_sub(qr_mul(q1.z, ld2), q2.z);
                    a4 = qr_mul(cr2, a7);
                    a2 = qr_mul(cr2, qr_mul(q1.z, cr2));
                    a7 = qr_sub(a7, q2.z);
                    a6 = qr_mul(ld2, a7);

                    a1 = qr_add(a1, a2); a3 = qr_add(a3, a4); a5 = qr_add(a5, a6);

                    a5 = qr_sub(a5, q1.w);
                    a_val = a1;
                    b_val = a3;
                    c_val = a5;
                    d_val = qr_sub(qr_mul(a3, a3), qr_mul(a5, a1));
                }

                /* QD_rts 4449-4547 */
                if (!(0.0f <= d_val)) break;
                const qr_f4 q7 = QR_SURF(v, so, 7);
                const float b = qr_neg(b_val);
                const bool dmask = d_val < q7.y;
                const float sd = qr_u2f(qr_f2u(qr_sqrt(d_val)) ^ (qr_f2u(b) & 0x80000000u));
                const float bd = qr_add(b, sd);
                const bool m_pos = 0.0f <= sd;
                float t1n = m_pos ? c_val : bd;     /* lazily divided roots (outer, inner) */
                float t1d = m_pos ? bd : a_val;
                float t2n = m_pos ? bd : c_val;
                float t2d = m_pos ? a_val : bd;
                const uint32_t amask = qr_f2u(a_val) & 0x80000000u;
                float t1 = 0.0f, t2 = 0.0f;         /* the roots when dmask */
                bool  k1 = true, k2 = true;

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
                    const float a5 = qr_abs(qr_mul(fm ? q7.z : 0.0f, t1));
                    a2 = qr_sub(qr_mul(a2, -0.5f), a5);
                    uint32_t u2 = qr_f2u(a2) ^ amask;
                    if (!(k1 && k2)) u2 = 0;
                    t1 = qr_add(t1, qr_u2f(u2));
                    t2 = qr_sub(t2, qr_u2f(u2));
                }

                /* QD_srt 4646-4824, one lane: the side tried first follows the
                 * sign of "a"; a hit on the first side ends the surface */
                const int first = qr_gt(0.0f, a_val) ? QR_FLAG_SIDE_INNER : QR_FLAG_SIDE_OUTER;

#pragma unroll 1
                for (int pass = 0; pass < 2; pass++)
                {
                    const int side = first ^ pass;
                    /* CHECK_SIDE 531-540 */
                    if (same && (pf == 1 - side || pf == 2 + side)) continue;
                    const bool outer = side == QR_FLAG_SIDE_OUTER;
                    float t; bool k;
                    if (dmask)
                    {
                        t = outer ? t1 : t2;
                        k = outer ? k1 : k2;
                    }
                    else
                    {
                        const float nn = outer ? t1n : t2n;
                        const float dd = outer ? t1d : t2d;
                        /* a quotient with the sign bit set cannot pass t_min < t */
                        if (no_neg && (int32_t)(qr_f2u(nn) ^ qr_f2u(dd)) < 0) continue;
                        t = qr_div(nn, dd);
                        k = dd != 0.0f;
                    }
                    if (!k) continue;
                    if (qr_candidate<SH>(v, sc, w, so, mode, bo0, bo1, bo2, cr0, cr1, cr2, ld0, ld1, ld2, true,
                                         t, dmask, amask, side, t_min, t_buf, cur)) break;
                }
            }
            while (0);

            qr_e_skip(cur, (int32_t)sizeof(qr_kelem));
            e = qr_e_ld(cur);
            continue;
        }

        /* the rare kinds */
        if (kind == QR_K_END) break;
        if (kind == QR_K_OPEN)
        {
            /* array with a matrix: transform origin diff and ray once for the
             * elements up to the node's last one (1483-1496) */
            const qr_f4 q0 = QR_SURF(v, so, 0);
            const qr_f4 q5 = QR_SURF(v, so, 5), q6 = QR_SURF(v, so, 6);
            const float tckz = QR_SURF(v, so, 7).x;
            const uint32_t trm = QR_D_TRM(qr_f2u(q0.w));
            float n0, n1, n2;
            qr_xform(q5, q6, tckz, trm, qr_sub(bo0, q0.x), qr_sub(bo1, q0.y), qr_sub(bo2, q0.z),
                     n0, n1, n2);
            bo0 = n0; bo1 = n1; bo2 = n2;
            qr_xform(q5, q6, tckz, trm, cr0, cr1, cr2, n0, n1, n2);
            cr0 = n0; cr1 = n1; cr2 = n2;
        }
        else
        if (kind == QR_K_CLOSE)
        {
            /* behind the last element of the open transform node / an own
             * matrix: back to the world (word loads, straight into the
             * ray's registers) */
            bo0 = qr_u2f(qr_sc_ld1(sc, QR_SC_WORG, 0)); bo1 = qr_u2f(qr_sc_ld1(sc, QR_SC_WORG, 1));
            bo2 = qr_u2f(qr_sc_ld1(sc, QR_SC_WORG, 2));
            cr0 = qr_u2f(qr_sc_ld1(sc, QR_SC_WRAY, 0)); cr1 = qr_u2f(qr_sc_ld1(sc, QR_SC_WRAY, 1));
            cr2 = qr_u2f(qr_sc_ld1(sc, QR_SC_WRAY, 2));
        }
        qr_e_skip(cur, kind == QR_K_JUMP ? e.aux : (int32_t)sizeof(qr_kelem));
        e = qr_e_ld(cur);
    }

    /* a closest-hit walk lowered t_buf below the t_max it started from when
     * something was hit (the depth test is strict, 1602-1605); a shadow walk
     * marks an occluder with t_buf = -inf */
    return t_buf;
}

/* texel -> linear colour, PAINT_COLX 664-673 */
QR_HD float qr_unpack(uint32_t texel, int sh, uint32_t cmask, float clamp, uint32_t props)
{
    float c = (float)(int32_t)((texel >> sh) & cmask);
    c = qr_div(c, clamp);
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
QR_HD float qr_norm_dot(float rx, float ry, float rz, float nx, float ny, float nz,
                        float &ax, float &ay, float &az)
{
    float s0 = qr_mul(rx, rx);
    s0 = qr_add(s0, qr_mul(ry, ry));
    s0 = qr_add(s0, qr_mul(rz, rz));
    const float inv = qr_rsq(s0);
    ax = qr_mul(rx, inv);
    ay = qr_mul(ry, inv);
    az = qr_mul(rz, inv);
    float d = qr_mul(ax, nx);
    d = qr_add(d, qr_mul(ay, ny));
    d = qr_add(d, qr_mul(az, nz));
    return d;
}

/* exact dielectric Fresnel, tracer.cpp:3385-3400 / 3781-3796 */
QR_HD float qr_fresnel(float c, float rfr, float x0, float x7)
{
    float a1 = c;
    const float a2 = qr_sub(qr_mul(a1, rfr), x7);
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
 * Returns the sample colour (before clamp / AA / gamma); the primary T_BUF is
 * left in the scratch (QR_SC_MISC.x).
 *
 * Register discipline: the only values that live across a list walk are the
 * ray itself and a few scalars.  The shading state of the level that casts a
 * shadow ray is parked in the thread's scratch quads for the duration of the
 * walk and read back afterwards; the state of a level that waits for a child
 * ray sits in the continuation stack.  That keeps the walk -- where nearly
 * all time goes -- at a register count that lets more warps be resident.
 */
template <bool SH>
QR_HD void qr_trace_sample(const qr_view<SH> &v, int px, int py, int lane4,
                           qr_frame *stack, const qr_scratch sc,
                           float &out_r, float &out_g, float &out_b)
{
    const qr_blob_header &h = *v.h;

    /* current ray; after a hit is shaded (ox, oy, oz) is the hit point HIT_X/Y/Z */
    float ox, oy, oz, rx, ry, rz;
    float t_min, t_max;
    uint32_t head, p_obj;
    int   mode, p_flg;
    int   lvl = 0;

    /* shading state of the current level */
    float lrx, lry, lrz;                /* RAY of the level (rx.. may hold a shadow ray) */
    float nx, ny, nz, lcx, lcy, lcz;
    float tr, tg, tb, cr, cg, cb;
    float dot, c_trn = 0.0f, c_rfl = 0.0f;
    float xr = 0.0f, xg = 0.0f, xb = 0.0f;
    uint32_t cur_so, li;
    int   l_flg;

    /* 1287-1322: primary ray; hor_i / ver_i are exact integers */
    {
        float hs = qr_add((float)px, h.hor_a[lane4]);
        float vs = qr_add((float)py, h.ver_a[lane4]);
        hs = qr_add(hs, 0.0f);
        vs = qr_add(vs, 0.0f);
        rx = qr_add(qr_add(qr_mul(h.hor[0], hs), qr_mul(h.ver[0], vs)), h.dir[0]);
        ry = qr_add(qr_add(qr_mul(h.hor[1], hs), qr_mul(h.ver[1], vs)), h.dir[1]);
        rz = qr_add(qr_add(qr_mul(h.hor[2], hs), qr_mul(h.ver[2], vs)), h.dir[2]);
        ox = h.org[0]; oy = h.org[1]; oz = h.org[2];
        t_min = h.t_min;
        t_max = h.cam_t_max;
        int tx = px / h.tile_w;
        if (tx >= h.tls_row) tx = h.tls_row - 1;
        QR_CHECK((uint32_t)((py / h.tile_h) * h.tls_row + tx) < (uint32_t)h.n_tiles, 5);
        head = (uint32_t)v.tiles[(py / h.tile_h) * h.tls_row + tx];
        mode = QR_MODE_CLOSEST;
        p_obj = QR_SO_NIL;
        p_flg = (int)h.ctx_flags;
    }

    for (;;)
    {
        /* ---------------- WALK ---------------- */
        const float t_buf = qr_walk<SH>(v.surf, v.elems, head, mode, ox, oy, oz, rx, ry, rz, t_min, t_max,
                                        p_obj, p_flg, sc);
        const bool res = mode == QR_MODE_SHADOW ? t_buf < 0.0f : t_buf < h.cam_t_max;
        {
            /* the ray comes back from where the walk parked it, so it does not
             * occupy registers during the walk */
            const qr_f4 wo = qr_sc_ld(sc, QR_SC_WORG), wr = qr_sc_ld(sc, QR_SC_WRAY);
            ox = wo.x; oy = wo.y; oz = wo.z;
            rx = wr.x; ry = wr.y; rz = wr.z;
        }
        int resume = 0;                 /* 0 none, 1 after refraction, 2 after reflection, -1 return */

        if (mode == QR_MODE_SHADOW)
        {
            /* back from a shadow ray: the level's state returns from the scratch */
            {
                const qr_f4 a = qr_sc_ld(sc, QR_SC_COL), b = qr_sc_ld(sc, QR_SC_RAY);
                const qr_f4 c = qr_sc_ld(sc, QR_SC_NRM), d = qr_sc_ld(sc, QR_SC_LOC);
                const qr_f4 e = qr_sc_ld(sc, QR_SC_TEX);
                QR_CHECK(qr_f2u(e.w) == (QR_POISON ^ 1u), 6);
                QR_CHECKED_ONLY(qr_sc_st1(sc, QR_SC_TEX, 3, 0u);)
                cr = a.x; cg = a.y; cb = a.z; dot = a.w;
                lrx = b.x; lry = b.y; lrz = b.z; li = qr_f2u(b.w);
                nx = c.x; ny = c.y; nz = c.z;
                lcx = d.x; lcy = d.y; lcz = d.z;
                tr = e.x; tg = e.y; tb = e.z;
                cur_so = p_obj;
                l_flg = p_flg & ~(QR_FLAG_PASS | QR_FLAG_SHAD);
            }

            /* LT_ret 2833-3151: light contribution unless occluded */
            if (!res)
            {
                const uint32_t lgt = v.elems[li].w;
                const qr_f4 l1 = QR_LGT(v, lgt, 1), l2 = QR_LGT(v, lgt, 2);
                const qr_f4 sh0 = QR_SHADE(v, cur_so, 0);
                const int mi = (int)qr_f2u((l_flg & 1) ? sh0.y : sh0.x);
                const qr_f4 m2 = QR_MAT(v, mi, 2);
                const uint32_t props = (uint32_t)l_flg;

                /* (rx, ry, rz) holds NEW_X/Y/Z = light vector */
                float x4 = qr_mul(rx, rx);
                x4 = qr_add(x4, qr_mul(ry, ry));
                x4 = qr_add(x4, qr_mul(rz, rz));
                const float r2 = x4;
                float dd, x6;
                if (props & QR_PROP_DIFFUSE)
                {
                    dd = dot;
                    x6 = x4;
                    const float x5 = qr_rsq(x4);
                    x4 = qr_mul(x5, x6);
                    x6 = qr_mul(x6, l1.w);
                    x4 = qr_mul(x4, l2.x);
                    x6 = qr_add(x6, l2.y);
                    x6 = qr_add(x6, x4);
                    x4 = qr_rsq(x6);
                    x6 = dd;
                    dd = qr_mul(dd, x4);
                    dd = qr_mul(dd, x5);
                    dd = qr_mul(dd, m2.x);
                }
                else
                {
                    x6 = dot;
                    dd = 0.0f;
                }

                bool spec_done = false;
                float spc = 0.0f;
                if (props & QR_PROP_SPECULAR)
                {
                    float x1 = rx, x2 = ry, x3 = rz;
                    float a4 = qr_mul(x6, nx);
                    x1 = qr_sub(x1, a4); x1 = qr_sub(x1, a4);
                    float a5 = qr_mul(x6, ny);
                    x2 = qr_sub(x2, a5); x2 = qr_sub(x2, a5);
                    float a6 = qr_mul(x6, nz);
                    x3 = qr_sub(x3, a6); x3 = qr_sub(x3, a6);
                    a4 = lrx; x1 = qr_mul(x1, a4); a4 = qr_mul(a4, a4);
                    a5 = lry; x2 = qr_mul(x2, a5); a5 = qr_mul(a5, a5);
                    a6 = lrz; x3 = qr_mul(x3, a6); a6 = qr_mul(a6, a6);
                    a6 = qr_add(a6, a4);
                    a6 = qr_add(a6, a5);
                    x1 = qr_add(x1, x2);
                    x1 = qr_add(x1, x3);
                    if (0.0f < x1)
                    {
                        spec_done = true;
                        x1 = qr_mul(x1, qr_rsq(a6));
                        x1 = qr_mul(x1, qr_rsq(r2));
                        x1 = qr_pow_28_4(x1, qr_f2u(m2.z));
                        spc = qr_mul(x1, m2.y);
                    }
                }

                if (spec_done && !(props & QR_PROP_METAL))
                {
                    /* LT_mtl 3090-3149 */
                    cr = qr_add(qr_add(qr_mul(qr_mul(tr, dd), l1.x), qr_mul(l1.x, spc)), cr);
                    cg = qr_add(qr_add(qr_mul(qr_mul(tg, dd), l1.y), qr_mul(l1.y, spc)), cg);
                    cb = qr_add(qr_add(qr_mul(qr_mul(tb, dd), l1.z), qr_mul(l1.z, spc)), cb);
                }
                else
                {
                    /* LT_spc 3047-3084 */
                    if (spec_done) dd = qr_add(dd, spc);
                    cr = qr_add(qr_mul(qr_mul(tr, l1.x), dd), cr);
                    cg = qr_add(qr_mul(qr_mul(tg, l1.y), dd), cg);
                    cb = qr_add(qr_mul(qr_mul(tb, l1.z), dd), cb);
                }
            }
            li = li + 1;
        }
        else
        {
            if (lvl == 0) qr_sc_st1(sc, QR_SC_MISC, 0, qr_f2u(t_buf));

            if (!res)
            {
                /* nothing hit: COL of this level stays 0 */
                cr = cg = cb = 0.0f;
                resume = -1;
                /* (dead values; keeps every path of the state machine defined) */
                lrx = lry = lrz = nx = ny = nz = lcx = lcy = lcz = tr = tg = tb = 0.0f;
                cur_so = QR_SO_NIL; l_flg = 0; li = 0;
            }
            else
            {
                /* ---------------- SHADE ---------------- */
                const qr_f4 hr = qr_sc_ld(sc, QR_SC_BEST);
                QR_CHECK(qr_f2u(hr.x) != QR_POISON, 2);
                cur_so = qr_f2u(hr.x) & ~127u;
                const qr_f4 q1 = QR_SURF(v, cur_so, 1), q2 = QR_SURF(v, cur_so, 2);
                const uint32_t d = qr_f2u(QR_SURF(v, cur_so, 0).w);
                const int side = (int)(qr_f2u(hr.x) & 1u);
                const uint32_t props = (uint32_t)side | qr_side_props(qr_f2u(q2.w), side);   /* FETCH_PROP */
                l_flg = (int)props;
                const qr_f4 s0 = QR_SHADE(v, cur_so, 0);
                const int mi = (int)qr_f2u(side ? s0.y : s0.x);

                lrx = rx; lry = ry; lrz = rz;
                ox = qr_add(qr_mul(rx, t_buf), ox);
                oy = qr_add(qr_mul(ry, t_buf), oy);
                oz = qr_add(qr_mul(rz, t_buf), oz);
                lcx = hr.y; lcy = hr.z; lcz = hr.w;

                const uint32_t kind = QR_D_TAG(d) == 1 ? 1u : QR_D_KIND(d);
                float tex_u = 0.0f, tex_v = 0.0f;
                float n0 = 0.0f, n1 = 0.0f, n2 = 0.0f;  /* normal, local fields */

                if (kind == 1)
                {
                    /* PL_mat 4149-4193 */
                    if (props & QR_PROP_TEXTURE)
                    {
                        tex_u = qr_sgn(qr_pick3(QR_D_MAP(d, 0), lcx, lcy, lcz), QR_D_SGN(d, 0));
                        tex_v = qr_sgn(qr_pick3(QR_D_MAP(d, 1), lcx, lcy, lcz), QR_D_SGN(d, 1));
                    }
                    if (props & QR_PROP_NORMAL)
                    {
                        const uint32_t u = (0x3F800000u ^ (side ? 0x80000000u : 0u))
                                         ^ (QR_D_SGN(d, 2) ? 0x80000000u : 0u);
                        qr_put3(QR_D_MAP(d, 2), qr_u2f(u), n0, n1, n2);
                    }
                }
                else
                if (props & QR_PROP_NORMAL)
                {
                    /* QD_mat 4855-4899 / TP_mat 4290-4330 */
                    float x4 = qr_mul(lcx, q1.x);
                    float x5 = qr_mul(lcy, q1.y);
                    float x6 = qr_mul(lcz, q1.z);
                    if (kind == 2)
                    {
                        x4 = qr_sub(x4, q2.x);
                        x5 = qr_sub(x5, q2.y);
                        x6 = qr_sub(x6, q2.z);
                    }
                    float x1 = qr_mul(x4, x4);
                    x1 = qr_add(x1, qr_mul(x5, x5));
                    x1 = qr_add(x1, qr_mul(x6, x6));
                    float x0 = qr_rsq(x1);
                    if (side) x0 = qr_neg(x0);
                    n0 = qr_mul(x4, x0);
                    n1 = qr_mul(x5, x0);
                    n2 = qr_mul(x6, x0);
                }

                /* a material without RT_PROP_NORMAL never reads the normal */
                nx = n0; ny = n1; nz = n2;
                if ((props & QR_PROP_NORMAL) && (d & QR_D_TRM_MASK))
                {
                    /* MT_nrm 2184-2259: transposed matrix of the trnode */
                    const uint32_t ti = qr_f2u(QR_SURF(v, cur_so, 4).w);
                    const qr_f4 t5 = QR_SURF(v, ti, 5), t6 = QR_SURF(v, ti, 6);
                    const float tck_z = QR_SURF(v, ti, 7).x;
                    const uint32_t ttrm = QR_D_TRM(qr_f2u(QR_SURF(v, ti, 0).w));
                    float x4 = qr_mul(t5.x, n0);
                    float x5 = qr_mul(t6.x, n1);
                    float x6 = qr_mul(tck_z, n2);
                    bool renorm = true;
                    if (ttrm != 1)
                    {
                        x4 = qr_add(x4, qr_mul(t5.w, n1));
                        x4 = qr_add(x4, qr_mul(t6.z, n2));
                        x5 = qr_add(x5, qr_mul(t5.y, n0));
                        x5 = qr_add(x5, qr_mul(t6.w, n2));
                        x6 = qr_add(x6, qr_mul(t5.z, n0));
                        x6 = qr_add(x6, qr_mul(t6.y, n1));
                        if (ttrm == 2) renorm = false;
                    }
                    if (renorm)
                    {
                        float x1 = qr_mul(x4, x4);
                        x1 = qr_add(x1, qr_mul(x5, x5));
                        x1 = qr_add(x1, qr_mul(x6, x6));
                        const float x0 = qr_rsq(x1);
                        x4 = qr_mul(x4, x0); x5 = qr_mul(x5, x0); x6 = qr_mul(x6, x0);
                    }
                    nx = x4; ny = x5; nz = x6;
                }

                /* MT_mat 2286-2327: texel */
                const qr_f4 m1 = QR_MAT(v, mi, 1), m4 = QR_MAT(v, mi, 4);
                uint32_t p = 0;
                if (props & QR_PROP_TEXTURE)
                {
                    const qr_f4 m0 = QR_MAT(v, mi, 0);
                    const uint32_t ys = qr_f2u(m1.z);
                    float tx = (ys & 0x100u) ? tex_v : tex_u;
                    float ty = (ys & 0x200u) ? tex_v : tex_u;
                    tx = qr_sub(tx, m0.z);
                    ty = qr_sub(ty, m0.w);
                    tx = qr_mul(tx, m0.x);
                    ty = qr_mul(ty, m0.y);
                    const uint32_t ix = (uint32_t)qr_cvm(tx) & qr_f2u(m1.x);
                    const uint32_t iy = ((uint32_t)qr_cvm(ty) & qr_f2u(m1.y)) << (ys & 0xFFu);
                    p = ix + iy;
                }
                const uint32_t texel = v.texels[(int)qr_f2u(m1.w) + p];
                tr = qr_unpack(texel, 16, qr_f2u(m4.z), m4.y, props);
                tg = qr_unpack(texel, 8, qr_f2u(m4.z), m4.y, props);
                tb = qr_unpack(texel, 0, qr_f2u(m4.z), m4.y, props);

                if (props & QR_PROP_LIGHT)
                {
                    /* LT_set 3164-3177 */
                    cr = tr; cg = tg; cb = tb;
                    li = 0;
                }
                else
                {
                    /* ambient 2721-2756 */
                    cr = qr_mul(tr, h.amb[0]);
                    cg = qr_mul(tg, h.amb[1]);
                    cb = qr_mul(tb, h.amb[2]);
                    li = qr_f2u(side ? s0.w : s0.z);
                }
            }
        }

        if (resume == 0)
        {
            /* LT_cyc 2762-2831: next light that sees the front of the surface */
            bool go_shadow = false;
            for (;;)
            {
                const qr_kelem le = v.elems[li];
                if (le.w == QR_KEND) break;
                const qr_f4 l0 = QR_LGT(v, le.w, 0);
                const float x1 = qr_sub(l0.x, ox);
                const float x2 = qr_sub(l0.y, oy);
                const float x3 = qr_sub(l0.z, oz);
                float dd = qr_mul(x1, nx);
                dd = qr_add(dd, qr_mul(x2, ny));
                dd = qr_add(dd, qr_mul(x3, nz));
                if (0.0f < dd)
                {
                    /* shadow ray: park the level, (ox, oy, oz) already is the hit */
                    qr_sc_st(sc, QR_SC_COL, cr, cg, cb, dd);
                    qr_sc_st(sc, QR_SC_RAY, lrx, lry, lrz, qr_u2f(li));
                    qr_sc_st(sc, QR_SC_NRM, nx, ny, nz, 0.0f);
                    qr_sc_st(sc, QR_SC_LOC, lcx, lcy, lcz, 0.0f);
#if defined(QR_CHECKED) && defined(__CUDA_ARCH__)
                    qr_sc_st(sc, QR_SC_TEX, tr, tg, tb, qr_u2f(QR_POISON ^ 1u));    /* "parked" */
#else
                    qr_sc_st(sc, QR_SC_TEX, tr, tg, tb, 0.0f);
#endif
                    rx = x1; ry = x2; rz = x3;
                    t_min = 0.0f;
                    t_max = l0.w;
                    head = (uint32_t)le.aux;
                    mode = QR_MODE_SHADOW;
                    p_obj = cur_so;
                    p_flg = l_flg | QR_FLAG_PASS_BACK | QR_FLAG_SHAD;
                    qr_sc_inc(sc, QR_SC_MISC, 1);
                    go_shadow = true;
                    break;
                }
                li = li + 1;
            }
            if (go_shadow) continue;
        }

        /* ------------- TRANSPARENCY / REFLECTION / unwinding ------------- */
        bool walk_again = false;
        for (;;)
        {
            if (resume == -1)
            {
                /* return colour of the finished level to its parent */
                if (lvl == 0) break;
                lvl--;
                const qr_frame &f = stack[lvl];
                const float ccr = cr, ccg = cg, ccb = cb;
                if (f.stage != 0)
                {
                    /* RF_ret 3868-3884 and RF_mix 3888-3908: the level is
                     * finished with this; nothing else of it is needed */
                    const float rfl = f.c_rfl;
                    cr = qr_add(qr_mul(ccr, rfl), f.col[0]);
                    cg = qr_add(qr_mul(ccg, rfl), f.col[1]);
                    cb = qr_add(qr_mul(ccb, rfl), f.col[2]);
                    continue;
                }
                cr = f.col[0]; cg = f.col[1]; cb = f.col[2];
                lrx = f.ray[0]; lry = f.ray[1]; lrz = f.ray[2];
                ox = f.hit[0]; oy = f.hit[1]; oz = f.hit[2];
                nx = f.nrm[0]; ny = f.nrm[1]; nz = f.nrm[2];
                lcx = f.loc[0]; lcy = f.loc[1]; lcz = f.loc[2];
                c_trn = f.c_trn; c_rfl = f.c_rfl;
                cur_so = f.so; l_flg = f.flg;
                /* TR_ret 3534-3552 */
                xr = qr_mul(ccr, c_trn); xg = qr_mul(ccg, c_trn); xb = qr_mul(ccb, c_trn);
                resume = 1;
            }

            const int side = l_flg & 1;
            const uint32_t props = (uint32_t)l_flg;
            const qr_f4 s0 = QR_SHADE(v, cur_so, 0);
            const qr_f4 s1 = QR_SHADE(v, cur_so, 1);
            const int mi = (int)qr_f2u(side ? s0.y : s0.x);
            const qr_f4 m3 = QR_MAT(v, mi, 3);
            const float m_c_rfl = QR_MAT(v, mi, 2).w;
            const float m_c_trn = m3.x, m_c_rfr = m3.y, m_rfr_2 = m3.z, m_c_rcp = m3.w;

            bool push = false;
            int  push_stage = 0, push_flg = 0;
            uint32_t push_head = 0;
            float nwx = 0.0f, nwy = 0.0f, nwz = 0.0f;

            if (resume == 0)
            {
                /* TRANSPARENCY 3185-3532 */
                c_trn = m_c_trn;
                c_rfl = m_c_rfl;
                xr = xg = xb = 0.0f;

                if (!(props & QR_PROP_OPAQUE))
                {
                    bool go = true;
                    float x0 = 0.0f, x4 = 0.0f, x7 = 0.0f;
                    if ((props & QR_PROP_REFRACT) || (props & QR_PROP_FRESNEL))
                    {
                        float ax, ay, az;
                        x4 = qr_norm_dot(lrx, lry, lrz, nx, ny, nz, ax, ay, az);
                        x0 = qr_mul(x4, m_c_rfr);
                        x7 = qr_mul(x0, x0);
                        x7 = qr_add(x7, 1.0f);
                        x7 = qr_sub(x7, m_rfr_2);
                        if ((props & QR_PROP_FRESNEL) && !(0.0f <= x7))
                        {
                            /* TR_tir 3280-3295 */
                            c_trn = 0.0f;
                            c_rfl = qr_add(m_c_rfl, m_c_trn);
                            go = false;
                        }
                        if (go)
                        {
                            x7 = qr_sqrt(x7);
                            x0 = qr_add(x0, x7);
                            if (props & QR_PROP_REFRACT)
                            {
                                nwx = qr_sub(qr_mul(ax, m_c_rfr), qr_mul(nx, x0));
                                nwy = qr_sub(qr_mul(ay, m_c_rfr), qr_mul(ny, x0));
                                nwz = qr_sub(qr_mul(az, m_c_rfr), qr_mul(nz, x0));
                            }
                            else
                            {
                                nwx = lrx; nwy = lry; nwz = lrz;
                            }
                        }
                    }
                    else
                    {
                        nwx = lrx; nwy = lry; nwz = lrz;
                    }
                    if (go && (props & QR_PROP_FRESNEL))
                    {
                        /* TR_ini 3385-3424 */
                        float a0 = qr_fresnel(x4, m_c_rfr, x0, x7);
                        a0 = qr_mul(a0, m_c_trn);
                        c_trn = qr_sub(m_c_trn, a0);
                        c_rfl = qr_add(m_c_rfl, a0);
                    }
                    if (go && lvl < h.depth)
                    {
                        push = true;
                        push_stage = 0;
                        push_head = qr_f2u(side ? s1.x : s1.y);         /* FETCH_IPTR */
                        push_flg = l_flg | QR_FLAG_PASS_THRU;
                        qr_sc_inc(sc, QR_SC_MISC, 3);
                    }
                }
                if (!push) resume = 1;
            }

            if (resume == 1)
            {
                /* TR_mix 3564-3598 */
                float x0 = qr_sub(1.0f, m_c_trn);
                x0 = qr_sub(x0, m_c_rfl);
                if (!(0.0f <= x0)) x0 = 0.0f;
                cr = qr_add(xr, qr_mul(cr, x0));
                cg = qr_add(xg, qr_mul(cg, x0));
                cb = qr_add(xb, qr_mul(cb, x0));

                /* REFLECTIONS 3604-3866 */
                bool go = (props & QR_PROP_REFLECT) != 0;
                if (!go && !(props & QR_PROP_OPAQUE) && (props & QR_PROP_FRESNEL)) go = true;
                if (!go)
                {
                    resume = -1;
                    continue;
                }

                float ax, ay, az;
                const float dd = qr_norm_dot(lrx, lry, lrz, nx, ny, nz, ax, ay, az);
                {
                    float nd = qr_mul(nx, dd); nwx = qr_sub(qr_sub(ax, nd), nd);
                    nd = qr_mul(ny, dd);       nwy = qr_sub(qr_sub(ay, nd), nd);
                    nd = qr_mul(nz, dd);       nwz = qr_sub(qr_sub(az, nd), nd);
                }

                if ((props & QR_PROP_FRESNEL) && (props & QR_PROP_OPAQUE))
                {
                    float a0 = dd;
                    if (props & QR_PROP_METAL)
                    {
                        /* 3729-3751 */
                        float a6 = m_c_rcp;
                        float a4 = qr_mul(a0, a6);
                        a4 = qr_add(a4, a4);
                        a0 = qr_mul(a0, a0);
                        a6 = qr_mul(a6, a6);
                        a6 = qr_add(a6, QR_MAT(v, mi, 4).x);
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
                        float y0 = qr_mul(a0, m_c_rfr);
                        float y7 = qr_mul(y0, y0);
                        y7 = qr_add(y7, 1.0f);
                        y7 = qr_sub(y7, m_rfr_2);
                        y7 = qr_sqrt(y7);
                        y0 = qr_add(y0, y7);
                        a0 = qr_fresnel(a0, m_c_rfr, y0, y7);
                    }
                    /* RF_pre 3806-3815 */
                    a0 = qr_sub(a0, 1.0f);
                    a0 = qr_mul(a0, m_c_rfl);
                    c_rfl = qr_add(m_c_rfl, a0);
                }

                xr = xg = xb = 0.0f;
                if (lvl < h.depth)
                {
                    push = true;
                    push_stage = 1;
                    push_head = qr_f2u(side ? s1.y : s1.x);             /* FETCH_XPTR */
                    push_flg = l_flg | QR_FLAG_PASS_BACK;
                    qr_sc_inc(sc, QR_SC_MISC, 2);
                }
                else
                {
                    resume = 2;
                }
            }

            if (push)
            {
                QR_CHECK(lvl >= 0 && lvl <= QR_STACK_DEPTH && lvl <= h.depth, 3);
                qr_frame &f = stack[lvl];
                f.col[0] = cr; f.col[1] = cg; f.col[2] = cb;
                f.c_rfl = c_rfl; f.stage = push_stage;
                if (push_stage == 0)
                {
                    /* a level that waits for its refraction child goes on
                     * afterwards (TR_mix, reflection): all of it */
                    f.ray[0] = lrx; f.ray[1] = lry; f.ray[2] = lrz;
                    f.hit[0] = ox; f.hit[1] = oy; f.hit[2] = oz;
                    f.nrm[0] = nx; f.nrm[1] = ny; f.nrm[2] = nz;
                    f.loc[0] = lcx; f.loc[1] = lcy; f.loc[2] = lcz;
                    f.c_trn = c_trn;
                    f.so = cur_so; f.flg = l_flg;
                }
                /* the child's walk finds this level's local hit in the scratch */
                qr_sc_st(sc, QR_SC_LOC, lcx, lcy, lcz, 0.0f);
                rx = nwx; ry = nwy; rz = nwz;
                t_min = 0.0f;
                t_max = h.cam_t_max;
                head = push_head;
                mode = QR_MODE_CLOSEST;
                p_obj = cur_so;
                p_flg = push_flg;
                lvl++;
                walk_again = true;
                break;
            }

            if (resume == 2)
            {
                /* RF_mix 3888-3908 */
                cr = qr_add(xr, cr);
                cg = qr_add(xg, cg);
                cb = qr_add(xb, cb);
                resume = -1;
            }
        }
        if (!walk_again) break;
    }

    out_r = cr; out_g = cg; out_b = cb;
}

/* ---- epilogue helpers, XX_end 5221-5343 / FRAME_SIMD 988-1006 -------------- */

QR_HD float qr_clamp1(float c)                  /* minps: source on NaN */
{
    return c < 1.0f ? c : 1.0f;
}

QR_HD uint32_t qr_pack(const qr_blob_header &h, float r, float g, float b)
{
    const bool gamma = (h.ctx_flags & QR_PROP_GAMMA) != 0;
    const float clampv = h.cam_clamp;
    const uint32_t cmask = h.cam_cmask;
    if (gamma) { r = qr_sqrt(r); g = qr_sqrt(g); b = qr_sqrt(b); }
    const uint32_t ir = (uint32_t)qr_cvn(qr_mul(r, clampv)) & cmask;
    const uint32_t ig = (uint32_t)qr_cvn(qr_mul(g, clampv)) & cmask;
    const uint32_t ib = (uint32_t)qr_cvn(qr_mul(b, clampv)) & cmask;
    return (ir << 16) | (ig << 8) | ib;
}

#endif /* QR_CORE_CUH */
