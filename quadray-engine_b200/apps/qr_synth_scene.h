/*
 * qr_synth_scene.h: generated scene of N random quadrics under bounding-volume
 * arrays (BASELINE.json config 5; SURVEY.md 8d "C5").
 *
 * The scene is produced in the reference's statically-linkable data format
 * (core/engine/format.h:170-760: rt_SCENE / rt_OBJECT / rt_OBJ / rt_RELATION and
 * the per-quadric structs), only at run time instead of as C initialisers, so
 * the very same generator feeds the unmodified reference core and the B200
 * backend through the public rt_Scene constructor (engine.h:325-330).
 *
 *   - N quadrics, tag uniform in RT_TAG_CYLINDER..RT_TAG_HYPERPARABOLOID (1..8),
 *     centres U[-E, E]^3, radii / shape parameters U[0.3, 1], clipped to a box
 *     of half-size U[1, 2] per axis, optionally rotated (own matrix);
 *   - std::mt19937(seed), every draw through one 24-bit uniform helper so the
 *     scene does not depend on the standard library's distributions;
 *   - leaves are sorted along a Morton curve and grouped 8-ary into arrays
 *     carrying {-1, RT_REL_BOUND_ARRAY, -1} (format.h:163), recursively up to
 *     one root: the reference's "tree accelerator" (bounding-volume elements
 *     with skip pointers, tracer.cpp:3955-4054);
 *   - one camera outside the cloud looking at its centre, one point light
 *     without distance attenuation.
 */

#ifndef QR_SYNTH_SCENE_H
#define QR_SYNTH_SCENE_H

#include <stdint.h>
#include <string.h>
#include <random>
#include <vector>
#include <algorithm>

#include "format.h"
#include "all_mat.h"

namespace qr_synth
{

struct Params
{
    int      n;         /* number of quadrics */
    unsigned seed;
    float    extent;    /* E: centres in [-E, E]^3 */
    int      rotate;    /* 1: random rotation per quadric */
    int      metal;     /* per-mille of quadrics with a reflective material */
};

static rt_CAMERA cm_synth =
{
    RT_CAM(PLAIN),
    RT_COL(0xFFFFFFFF),
    { 0.15 },                   /* amb */
    { 1.0 },                    /* pov */
    { 0.5, 0.5, 0.5 },
    { 1.5, 1.5, 1.5 },
};

static rt_LIGHT lt_synth =
{
    RT_LGT(PLAIN),
    RT_COL(0xFFFFFFFF),
    { 0.05, 1.0 },              /* amb, src */
    { 0.0, 1.0, 0.0, 0.0 },     /* rng, cnt, lnr, qdr: no attenuation */
};

/* storage that the scene graph points into; lives until the process ends */
struct Store
{
    std::vector<rt_CYLINDER>        cl;
    std::vector<rt_SPHERE>          sp;
    std::vector<rt_CONE>            cn;
    std::vector<rt_PARABOLOID>      pb;
    std::vector<rt_HYPERBOLOID>     hb;
    std::vector<rt_PARACYLINDER>    pc;
    std::vector<rt_HYPERCYLINDER>   hc;
    std::vector<rt_HYPERPARABOLOID> hp;
    std::vector<std::vector<rt_OBJECT> > arrays;    /* child arrays, bottom-up */
    rt_RELATION                     rel_bound[1];
    rt_OBJECT                       cam[1], lgt[1];
    std::vector<rt_OBJECT>          tree;
};

struct Rng
{
    std::mt19937 g;
    explicit Rng(unsigned s) : g(s) {}
    /* uniform in [0, 1) with 24 bits, exact in binary32 */
    float u01() { return (float)(g() >> 8) * (1.0f / 16777216.0f); }
    float uni(float lo, float hi) { return lo + (hi - lo) * u01(); }
    int   pick(int n) { return (int)(g() % (unsigned)n); }
};

static uint32_t morton_spread(uint32_t v)
{
    v &= 0x3FF;
    v = (v | (v << 16)) & 0x030000FF;
    v = (v | (v << 8))  & 0x0300F00F;
    v = (v | (v << 4))  & 0x030C30C3;
    v = (v | (v << 2))  & 0x09249249;
    return v;
}

static void identity(rt_OBJECT &o)
{
    memset(&o, 0, sizeof(o));
    o.trm.scl[0] = o.trm.scl[1] = o.trm.scl[2] = 1.0f;
}

static void fill_surface(rt_SURFACE &s, Rng &r, rt_MATERIAL *outer)
{
    for (int k = 0; k < 3; k++)
    {
        const float h = r.uni(1.0f, 2.0f);
        s.min[k] = -h;
        s.max[k] = +h;
    }
    rt_SIDE *sd[2] = { &s.side_outer, &s.side_inner };
    for (int i = 0; i < 2; i++)
    {
        sd[i]->scl[0] = sd[i]->scl[1] = 1.0f;
        sd[i]->rot = 0.0f;
        sd[i]->pos[0] = sd[i]->pos[1] = 0.0f;
    }
    s.side_outer.pmat = outer;
    s.side_inner.pmat = &mt_plain01_gray01;
}

static rt_OBJ make_obj(rt_si32 tag, rt_void *p)
{
    rt_OBJ o;
    memset(&o, 0, sizeof(o));
    o.tag = tag;
    o.pobj = p;
    o.obj_num = 1;
    return o;
}

static rt_OBJ make_arr(std::vector<rt_OBJECT> &a, rt_RELATION *rel, int rel_num)
{
    rt_OBJ o;
    memset(&o, 0, sizeof(o));
    o.tag = RT_TAG_ARRAY;
    o.pobj = a.data();
    o.obj_num = (rt_si32)a.size();
    o.prel = rel;
    o.rel_num = rel_num;
    return o;
}

/*
 * Build the scene.  The returned rt_SCENE (and everything it points to) stays
 * valid for the life of the process.
 */
static rt_SCENE build(const Params &p)
{
    static rt_MATERIAL *plain[] =
    {
        &mt_plain01_blue01, &mt_plain01_cyan01, &mt_plain01_green01, &mt_plain01_orange01,
        &mt_plain01_pink01, &mt_plain01_red01, &mt_plain01_white01, &mt_plain02_orange01,
    };
    static rt_MATERIAL *metal[] =
    {
        &mt_metal01_cyan01, &mt_metal01_pink01, &mt_metal02_orange01, &mt_metal03_nickel01,
    };

    Store *st = new Store();
    Rng r(p.seed);
    const int n = p.n;

    /* the vectors must not reallocate once objects point into them */
    st->cl.reserve(n); st->sp.reserve(n); st->cn.reserve(n); st->pb.reserve(n);
    st->hb.reserve(n); st->pc.reserve(n); st->hc.reserve(n); st->hp.reserve(n);

    struct Leaf { uint32_t key; rt_OBJECT obj; };
    std::vector<Leaf> leaves(n);

    for (int i = 0; i < n; i++)
    {
        rt_OBJECT &o = leaves[i].obj;
        identity(o);
        const int tag = 1 + r.pick(8);
        float c[3];
        for (int k = 0; k < 3; k++)
        {
            c[k] = r.uni(-p.extent, p.extent);
            o.trm.pos[k] = c[k];
            o.trm.rot[k] = p.rotate ? (float)(r.pick(24) * 15) : 0.0f;
        }
        rt_MATERIAL *m = r.pick(1000) < p.metal ? metal[r.pick(4)] : plain[r.pick(8)];
        const float a = r.uni(0.3f, 1.0f), b = r.uni(0.3f, 1.0f);

        switch (tag)
        {
            case RT_TAG_CYLINDER:
            {
                rt_CYLINDER q; memset(&q, 0, sizeof(q)); fill_surface(q.srf, r, m);
                q.rad = a;
                st->cl.push_back(q); o.obj = make_obj(tag, &st->cl.back());
            }
            break;
            case RT_TAG_SPHERE:
            {
                rt_SPHERE q; memset(&q, 0, sizeof(q)); fill_surface(q.srf, r, m);
                q.rad = a;
                st->sp.push_back(q); o.obj = make_obj(tag, &st->sp.back());
            }
            break;
            case RT_TAG_CONE:
            {
                rt_CONE q; memset(&q, 0, sizeof(q)); fill_surface(q.srf, r, m);
                q.rat = a;
                st->cn.push_back(q); o.obj = make_obj(tag, &st->cn.back());
            }
            break;
            case RT_TAG_PARABOLOID:
            {
                rt_PARABOLOID q; memset(&q, 0, sizeof(q)); fill_surface(q.srf, r, m);
                q.par = a;
                st->pb.push_back(q); o.obj = make_obj(tag, &st->pb.back());
            }
            break;
            case RT_TAG_HYPERBOLOID:
            {
                rt_HYPERBOLOID q; memset(&q, 0, sizeof(q)); fill_surface(q.srf, r, m);
                q.rat = a; q.hyp = b;
                st->hb.push_back(q); o.obj = make_obj(tag, &st->hb.back());
            }
            break;
            case RT_TAG_PARACYLINDER:
            {
                rt_PARACYLINDER q; memset(&q, 0, sizeof(q)); fill_surface(q.srf, r, m);
                q.par = a;
                st->pc.push_back(q); o.obj = make_obj(tag, &st->pc.back());
            }
            break;
            case RT_TAG_HYPERCYLINDER:
            {
                rt_HYPERCYLINDER q; memset(&q, 0, sizeof(q)); fill_surface(q.srf, r, m);
                q.rat = a; q.hyp = b;
                st->hc.push_back(q); o.obj = make_obj(tag, &st->hc.back());
            }
            break;
            default:
            {
                rt_HYPERPARABOLOID q; memset(&q, 0, sizeof(q)); fill_surface(q.srf, r, m);
                q.pr1 = a; q.pr2 = b;
                st->hp.push_back(q); o.obj = make_obj(RT_TAG_HYPERPARABOLOID, &st->hp.back());
            }
            break;
        }

        uint32_t g[3];
        for (int k = 0; k < 3; k++)
        {
            float t = (c[k] + p.extent) / (2.0f * p.extent) * 1023.0f;
            g[k] = (uint32_t)(t < 0.0f ? 0.0f : t > 1023.0f ? 1023.0f : t);
        }
        leaves[i].key = morton_spread(g[0]) | (morton_spread(g[1]) << 1) | (morton_spread(g[2]) << 2);
    }

    std::stable_sort(leaves.begin(), leaves.end(),
                     [](const Leaf &x, const Leaf &y) { return x.key < y.key; });

    st->rel_bound[0].obj1 = -1;
    st->rel_bound[0].rel  = RT_REL_BOUND_ARRAY;
    st->rel_bound[0].obj2 = -1;

    /* 8-ary grouping, bottom-up; "level" holds the nodes still to be grouped */
    std::vector<rt_OBJECT> level(n);
    for (int i = 0; i < n; i++) level[i] = leaves[i].obj;

    size_t total_arrays = 0;
    for (size_t m = n; m > 8; m = (m + 7) / 8) total_arrays += (m + 7) / 8;
    st->arrays.reserve(total_arrays + 1);

    while (level.size() > 8)
    {
        std::vector<rt_OBJECT> up;
        for (size_t i = 0; i < level.size(); i += 8)
        {
            const size_t e = std::min(level.size(), i + 8);
            st->arrays.push_back(std::vector<rt_OBJECT>(level.begin() + i, level.begin() + e));
            rt_OBJECT node;
            identity(node);
            node.obj = make_arr(st->arrays.back(), st->rel_bound, 1);
            up.push_back(node);
        }
        level.swap(up);
    }
    st->arrays.push_back(level);
    std::vector<rt_OBJECT> &top = st->arrays.back();

    /* camera: default orientation looks down -Z (test scenes use rot X = -90
     * +- a few degrees to look along +Y); stand back so the cloud fits */
    identity(st->cam[0]);
    st->cam[0].trm.rot[0] = -90.0f;
    st->cam[0].trm.pos[1] = -2.3f * p.extent;
    {
        rt_OBJ o; memset(&o, 0, sizeof(o));
        o.tag = RT_TAG_CAMERA; o.pobj = &cm_synth; o.obj_num = 1;
        st->cam[0].obj = o;
    }
    identity(st->lgt[0]);
    st->lgt[0].trm.pos[0] = -1.5f * p.extent;
    st->lgt[0].trm.pos[1] = -2.0f * p.extent;
    st->lgt[0].trm.pos[2] = +1.5f * p.extent;
    {
        rt_OBJ o; memset(&o, 0, sizeof(o));
        o.tag = RT_TAG_LIGHT; o.pobj = &lt_synth; o.obj_num = 1;
        st->lgt[0].obj = o;
    }

    st->tree.resize(3);
    identity(st->tree[0]);
    st->tree[0].obj = make_arr(top, st->rel_bound, 1);
    st->tree[1] = st->lgt[0];
    st->tree[2] = st->cam[0];

    rt_SCENE sc;
    memset(&sc, 0, sizeof(sc));
    sc.root = make_arr(st->tree, RT_NULL, 0);
    sc.opts = RT_OPTS_PT;
    return sc;
}

} /* namespace qr_synth */

#endif /* QR_SYNTH_SCENE_H */
