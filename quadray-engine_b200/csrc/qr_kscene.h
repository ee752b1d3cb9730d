/*
 * qr_kscene.h -- the kernels' packed scene image ("kscene").
 *
 * The scene blob of include/qr_scene_blob.h is the interchange format of the
 * C ABI: one field per word, easy to produce and to check.  The kernels read
 * a tighter image of the same data, built once per upload while the blob is
 * copied into pinned staging (qr_scene_upload):
 *
 *   - every record is a few 16-byte quads, so one traversal step is three
 *     128-bit shared-memory loads instead of ~15 scalar ones
 *   - all small integers of a surface (solver / material / clipper tags, axis
 *     maps and signs, transform class, min-max toggles) sit in ONE descriptor
 *     word, tested with bit ops
 *   - fields only shading needs (materials, light and reflection lists) are
 *     split off the traversal record
 *
 * Image layout:  header | ksurf | kshade | kmat | klgt || elems | tiles | texels
 * The part left of "||" (kscene prefix) is what a CTA stages in shared memory.
 * The header is a qr_blob_header whose section table is rewritten for the
 * image (flags = QR_KSCENE_FLAG, pad3[0] = offset of kshade).
 */
#ifndef QR_KSCENE_H
#define QR_KSCENE_H

#include <stdint.h>
#include <string.h>
#include <utility>
#include <vector>
#include "qr_scene_blob.h"

#define QR_KSCENE_FLAG   0x4B534331u    /* "1CSK" */

#define QR_KSURF_QUADS   8              /* 128 B per surface */
#define QR_KSHADE_QUADS  2              /*  32 B per surface */
#define QR_KMAT_QUADS    5              /*  80 B per material */
#define QR_KLGT_QUADS    3              /*  48 B per light */

#if defined(__CUDACC__)
struct __align__(16) qr_f4 { float x, y, z, w; };
#else
struct alignas(16) qr_f4 { float x, y, z, w; };
#endif

/*
 * List element of the image ("compiled" rt_ELEM).  The reference tracks, while
 * it walks a surface list, whether a transform node is open (ctx_LOCAL(OBJ),
 * tracer.cpp:1377-1421, 1492-1496, 4047-4053); that state only depends on the
 * list, not on the ray, so it is resolved here, once per upload:
 *   surface lists  simd = surface, aux = element to continue at when the
 *                  bounding volume is missed (next of the array's last leaf),
 *                  op  = QR_OP_BV | QR_OP_CACHED
 *   light lists    simd = light, aux = head of the light's shadow list
 *   clip lists     simd = clipper surface or QR_NIL (accum marker),
 *                  aux = trnode's last element (array clippers), op = clip
 *                  side / accum marker (rt_ELEM.data, +-1)
 */
#if defined(__CUDACC__)
struct __align__(16) qr_kelem { int32_t simd, next, aux, op; };
#else
struct alignas(16) qr_kelem { int32_t simd, next, aux, op; };
#endif

#define QR_OP_BV        1   /* bounding-volume element of an array (elm.data & 3 == 1) */
#define QR_OP_CACHED    2   /* child of the open transform node: diff = node diff - pos */
#define QR_OP_OPEN      4   /* array with a matrix: opens a transform node */
#define QR_OP_OWNTRM    8   /* surface with its own matrix, outside any open node */
#define QR_OP_CLOSE    16   /* last element of the open transform node */
#define QR_OP_SKIPCLOSE 32  /* a missed bounding volume skips past the node's last element */

/*
 * ksurf quads (traversal):
 *   q0  pos.x  pos.y  pos.z  desc
 *   q1  sci.x  sci.y  sci.z  sci.w
 *   q2  scj.x  scj.y  scj.z  props (outer | inner << 16)
 *   q3  min.x  min.y  min.z  clip_head
 *   q4  max.x  max.y  max.z  trnode
 *   q5  tci.x  tci.y  tci.z  tcj.x
 *   q6  tcj.y  tcj.z  tck.x  tck.y
 *   q7  tck.z  d_eps  t_eps  c_def
 * kshade quads (shading):
 *   s0  mat[0] mat[1] lst_lgt[0] lst_lgt[1]
 *   s1  lst_srf[0] lst_srf[1] 0 0
 * kmat quads:
 *   m0  xscal yscal xoffs yoffs
 *   m1  xmask ymask (yshft | t_map[0] << 8 | t_map[1] << 9) tex
 *   m2  l_dff l_spc l_pow c_rfl
 *   m3  c_trn c_rfr rfr_2 c_rcp
 *   m4  ext_2 clamp cmask 0
 * klgt quads:
 *   l0  pos.x pos.y pos.z t_max
 *   l1  col.r col.g col.b a_qdr
 *   l2  a_lnr a_cnt 0 0
 *
 * desc bits:
 *   1:0 solver srf_t[0]   3:2 material kind srf_t[1]   5:4 clipper srf_t[2]
 *   7:6 conic             9:8 a_map[L] (transform class)
 *   10  field shift (a_sgn[L] != 0)      11  array (srf_t[3] < 0)
 *   14:12 a_sgn[I,J,K]    16:15 / 18:17 / 20:19  a_map[I,J,K] - shift
 *   26:21 min/max toggles 27 has custom clippers
 */
#define QR_D_TAG(d)     ((d) & 3u)
#define QR_D_KIND(d)    (((d) >> 2) & 3u)
#define QR_D_CLIP(d)    (((d) >> 4) & 3u)
#define QR_D_CONIC(d)   (((d) >> 6) & 3u)
#define QR_D_TRM(d)     (((d) >> 8) & 3u)
#define QR_D_SHIFT(d)   (((d) >> 10) & 1u)
#define QR_D_ARRAY(d)   (((d) >> 11) & 1u)
#define QR_D_SGN(d, i)  (((d) >> (12 + (i))) & 1u)
#define QR_D_MAP(d, i)  (((d) >> (15 + 2 * (i))) & 3u)
#define QR_D_MM(d)      (((d) >> 21) & 63u)
#define QR_D_HASCLIP(d) (((d) >> 27) & 1u)
#define QR_D_CONIC_MASK   (3u << 6)
#define QR_D_TRM_MASK     (3u << 8)
#define QR_D_SHIFT_MASK   (1u << 10)
#define QR_D_ARRAY_MASK   (1u << 11)
#define QR_D_MM_MASK      (63u << 21)
#define QR_D_HASCLIP_MASK (1u << 27)

/* host-side packer (plain inline functions; never called from device code) */

static inline uint32_t qr_k_align16(uint32_t v) { return (v + 15u) & ~15u; }

static inline float qr_k_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* bytes of the kscene image of a (validated) blob */
static inline size_t qr_kscene_size(const void *blob)
{
    const qr_blob_header *h = (const qr_blob_header *)blob;
    uint32_t off = sizeof(qr_blob_header);
    off = qr_k_align16(off + (uint32_t)h->n_surf * QR_KSURF_QUADS * 16);
    off = qr_k_align16(off + (uint32_t)h->n_surf * QR_KSHADE_QUADS * 16);
    off = qr_k_align16(off + (uint32_t)h->n_mat * QR_KMAT_QUADS * 16);
    off = qr_k_align16(off + (uint32_t)h->n_lgt * QR_KLGT_QUADS * 16);
    off = qr_k_align16(off + (uint32_t)h->n_elem * sizeof(qr_elem));
    off = qr_k_align16(off + (uint32_t)h->n_tiles * sizeof(int32_t));
    off = qr_k_align16(off + (uint32_t)h->n_texels * sizeof(uint32_t));
    return off;
}

/*
 * Build the kscene image of "blob" in "out" (qr_kscene_size bytes, 16-aligned).
 * Returns 0, or -1 when a surface list is not well nested (an element would be
 * reached both inside and outside an open transform node).
 */
static inline int qr_kscene_pack(const void *blob, void *out)
{
    const uint8_t *b = (const uint8_t *)blob;
    uint8_t *o = (uint8_t *)out;
    const qr_blob_header *h = (const qr_blob_header *)blob;
    const qr_surface  *sf = (const qr_surface  *)(b + h->off_surf);
    const qr_material *mt = (const qr_material *)(b + h->off_mat);
    const qr_light    *lg = (const qr_light    *)(b + h->off_lgt);
    const qr_elem     *el = (const qr_elem     *)(b + h->off_elem);
    const int32_t     *tl = (const int32_t     *)(b + h->off_tiles);

    qr_blob_header k = *h;
    uint32_t off = sizeof(qr_blob_header);
    k.flags = QR_KSCENE_FLAG;
    k.off_surf = off;   off = qr_k_align16(off + (uint32_t)h->n_surf * QR_KSURF_QUADS * 16);
    k.pad3[0] = (int32_t)off; off = qr_k_align16(off + (uint32_t)h->n_surf * QR_KSHADE_QUADS * 16);
    k.off_mat = off;    off = qr_k_align16(off + (uint32_t)h->n_mat * QR_KMAT_QUADS * 16);
    k.off_lgt = off;    off = qr_k_align16(off + (uint32_t)h->n_lgt * QR_KLGT_QUADS * 16);
    k.off_elem = off;   off = qr_k_align16(off + (uint32_t)h->n_elem * sizeof(qr_elem));
    k.off_tiles = off;  off = qr_k_align16(off + (uint32_t)h->n_tiles * sizeof(int32_t));
    k.off_texels = off; off = qr_k_align16(off + (uint32_t)h->n_texels * sizeof(uint32_t));
    k.total_bytes = off;
    memcpy(o, &k, sizeof(k));

    qr_f4 *ks = (qr_f4 *)(o + k.off_surf);
    qr_f4 *kh = (qr_f4 *)(o + (uint32_t)k.pad3[0]);
    for (int i = 0; i < h->n_surf; i++)
    {
        const qr_surface &s = sf[i];
        const uint32_t shift = s.a_sgn[3] != 0 ? 1u : 0u;
        const int sub = shift ? 3 : 0;
        uint32_t d = 0;
        d |= ((uint32_t)s.srf_t[0] & 3u);
        d |= ((uint32_t)s.srf_t[1] & 3u) << 2;
        d |= ((uint32_t)s.srf_t[2] & 3u) << 4;
        d |= ((uint32_t)s.conic & 3u) << 6;
        d |= ((uint32_t)s.a_map[3] & 3u) << 8;
        d |= shift << 10;
        d |= (s.srf_t[3] < 0 ? 1u : 0u) << 11;
        for (int a = 0; a < 3; a++)
        {
            d |= (s.a_sgn[a] != 0 ? 1u : 0u) << (12 + a);
            d |= ((uint32_t)(s.a_map[a] - sub) & 3u) << (15 + 2 * a);
        }
        d |= ((uint32_t)s.minmax_t & 63u) << 21;
        d |= (s.clip_head != QR_NIL ? 1u : 0u) << 27;

        const uint32_t props = ((uint32_t)s.props[0] & 0xFFFFu) | (((uint32_t)s.props[1] & 0xFFFFu) << 16);
        qr_f4 *q = ks + (size_t)i * QR_KSURF_QUADS;
        q[0].x = s.pos[0]; q[0].y = s.pos[1]; q[0].z = s.pos[2]; q[0].w = qr_k_bits(d);
        q[1].x = s.sci[0]; q[1].y = s.sci[1]; q[1].z = s.sci[2]; q[1].w = s.sci[3];
        q[2].x = s.scj[0]; q[2].y = s.scj[1]; q[2].z = s.scj[2]; q[2].w = qr_k_bits(props);
        q[3].x = s.min[0]; q[3].y = s.min[1]; q[3].z = s.min[2]; q[3].w = qr_k_bits((uint32_t)s.clip_head);
        q[4].x = s.max[0]; q[4].y = s.max[1]; q[4].z = s.max[2]; q[4].w = qr_k_bits((uint32_t)s.trnode);
        q[5].x = s.tci[0]; q[5].y = s.tci[1]; q[5].z = s.tci[2]; q[5].w = s.tcj[0];
        q[6].x = s.tcj[1]; q[6].y = s.tcj[2]; q[6].z = s.tck[0]; q[6].w = s.tck[1];
        q[7].x = s.tck[2]; q[7].y = s.d_eps;  q[7].z = s.t_eps;  q[7].w = qr_k_bits(s.c_def);

        qr_f4 *g = kh + (size_t)i * QR_KSHADE_QUADS;
        g[0].x = qr_k_bits((uint32_t)s.mat[0]);     g[0].y = qr_k_bits((uint32_t)s.mat[1]);
        g[0].z = qr_k_bits((uint32_t)s.lst_lgt[0]); g[0].w = qr_k_bits((uint32_t)s.lst_lgt[1]);
        g[1].x = qr_k_bits((uint32_t)s.lst_srf[0]); g[1].y = qr_k_bits((uint32_t)s.lst_srf[1]);
        g[1].z = 0.0f; g[1].w = 0.0f;
    }

    qr_f4 *km = (qr_f4 *)(o + k.off_mat);
    for (int i = 0; i < h->n_mat; i++)
    {
        const qr_material &m = mt[i];
        qr_f4 *q = km + (size_t)i * QR_KMAT_QUADS;
        const uint32_t ys = (m.yshft & 0xFFu) | ((uint32_t)(m.t_map[0] & 1) << 8) | ((uint32_t)(m.t_map[1] & 1) << 9);
        q[0].x = m.xscal; q[0].y = m.yscal; q[0].z = m.xoffs; q[0].w = m.yoffs;
        q[1].x = qr_k_bits(m.xmask); q[1].y = qr_k_bits(m.ymask); q[1].z = qr_k_bits(ys); q[1].w = qr_k_bits((uint32_t)m.tex);
        q[2].x = m.l_dff; q[2].y = m.l_spc; q[2].z = qr_k_bits(m.l_pow); q[2].w = m.c_rfl;
        q[3].x = m.c_trn; q[3].y = m.c_rfr; q[3].z = m.rfr_2; q[3].w = m.c_rcp;
        q[4].x = m.ext_2; q[4].y = m.clamp; q[4].z = qr_k_bits(m.cmask); q[4].w = 0.0f;
    }

    qr_f4 *kl = (qr_f4 *)(o + k.off_lgt);
    for (int i = 0; i < h->n_lgt; i++)
    {
        const qr_light &l = lg[i];
        qr_f4 *q = kl + (size_t)i * QR_KLGT_QUADS;
        q[0].x = l.pos[0]; q[0].y = l.pos[1]; q[0].z = l.pos[2]; q[0].w = l.t_max;
        q[1].x = l.col[0]; q[1].y = l.col[1]; q[1].z = l.col[2]; q[1].w = l.a_qdr;
        q[2].x = l.a_lnr;  q[2].y = l.a_cnt;  q[2].z = 0.0f;     q[2].w = 0.0f;
    }

    /* ---- compile the lists ---- */
    qr_kelem *ke = (qr_kelem *)(o + k.off_elem);
    const int ne = h->n_elem;
    for (int i = 0; i < ne; i++)
    {
        ke[i].simd = el[i].simd;
        ke[i].next = el[i].next;
        ke[i].aux  = el[i].data_p;
        ke[i].op   = el[i].data_i;      /* light / clip lists keep the raw data */
    }
    {
        /* state[i]: the open trnode's last element when element i of a surface
         * list is processed (QR_NIL none), -2 = not reached yet */
        std::vector<int32_t> state((size_t)ne, -2);
        std::vector<int32_t> roots;
        for (int t = 0; t < h->n_tiles; t++) roots.push_back(tl[t]);
        for (int i = 0; i < h->n_surf; i++)
        {
            roots.push_back(sf[i].lst_srf[0]);
            roots.push_back(sf[i].lst_srf[1]);
            for (int sd = 0; sd < 2; sd++)
            {
                int guard = 0;
                for (int li = sf[i].lst_lgt[sd]; li != QR_NIL && guard <= ne; li = el[li].next, guard++)
                {
                    roots.push_back(el[li].data_p);
                }
            }
        }
        /* (element, state) pairs still to expand */
        std::vector<std::pair<int32_t, int32_t> > work;
        for (size_t r = 0; r < roots.size(); r++)
        {
            if (roots[r] != QR_NIL) work.push_back(std::make_pair(roots[r], (int32_t)QR_NIL));
        }
        while (!work.empty())
        {
            int32_t i = work.back().first, lobj = work.back().second;
            work.pop_back();
            while (i != QR_NIL)
            {
                if (i < 0 || i >= ne) return -1;
                if (state[i] != -2)
                {
                    if (state[i] != lobj) return -1;
                    break;                              /* suffix already compiled */
                }
                state[i] = lobj;
                const qr_elem &e = el[i];
                if (e.simd < 0 || e.simd >= h->n_surf) return -1;
                const qr_surface &s = sf[e.simd];
                const bool is_array = s.srf_t[3] < 0;
                int32_t op = 0;
                if (!is_array && lobj != QR_NIL)
                {
                    if (s.a_sgn[3] == 0) return -1;     /* child without the field shift */
                    op |= QR_OP_CACHED;
                    if (i == lobj)
                    {
                        op |= QR_OP_CLOSE;
                        lobj = QR_NIL;
                    }
                }
                else
                if (is_array && s.a_map[3] != 0)
                {
                    op |= QR_OP_OPEN;
                    lobj = e.data_p;                    /* tracer.cpp:1492-1496 */
                }
                else
                if (is_array && lobj != QR_NIL)
                {
                    return -1;                          /* plain array inside an open node */
                }
                else
                if (!is_array && s.a_map[3] != 0)
                {
                    op |= QR_OP_OWNTRM;
                }
                int32_t aux = QR_NIL;
                if (e.data_i == 1)
                {
                    op |= QR_OP_BV;
                    if (e.data_p == lobj) op |= QR_OP_SKIPCLOSE;
                    if (e.data_p < 0 || e.data_p >= ne) return -1;
                    aux = el[e.data_p].next;            /* tracer.cpp:4042-4054 */
                    const int32_t after = e.data_p == lobj ? (int32_t)QR_NIL : lobj;
                    if (aux != QR_NIL) work.push_back(std::make_pair(aux, after));
                }
                ke[i].op = op;
                ke[i].aux = aux;
                i = e.next;
            }
        }
    }
    memcpy(o + k.off_tiles,  b + h->off_tiles,  (size_t)h->n_tiles * sizeof(int32_t));
    memcpy(o + k.off_texels, b + h->off_texels, (size_t)h->n_texels * sizeof(uint32_t));
    return 0;
}

#endif /* QR_KSCENE_H */
