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
#include <unordered_map>
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
 * List element of the image ("compiled" rt_ELEM), 8 bytes.  Lists are laid
 * out SEQUENTIALLY: the successor of element i is element i + 1, a list ends
 * with an END element, element 0 of the array is an END (the NULL list) and a
 * tail shared with a list emitted earlier is entered through a JUMP.  So the
 * walker can load element i + 1 before it has looked at element i.
 *
 * The reference tracks, while it walks a surface list, whether a transform
 * node is open (ctx_LOCAL(OBJ), tracer.cpp:1377-1421, 1492-1496, 4047-4053);
 * that state only depends on the list, not on the ray, so it is resolved
 * here, once per upload, into the element stream itself:
 *   - an array with a matrix becomes an OPEN element (transform the ray into
 *     the node's space, tracer.cpp:1483-1496); the node's last element is
 *     followed by a CLOSE element (back to the world ray), which is also
 *     where a bounding volume INSIDE the node that ends with the node lands
 *     when it is missed (tracer.cpp:4047-4053);
 *   - a surface with its own matrix outside any node (OO_dff, 1429-1556)
 *     carries OWN (the walk transforms the ray for this one element) and is
 *     followed by a CLOSE as well;
 *   - every element inside a node (or OWN) carries the NODE flag: the world
 *     ray is parked in the thread's scratch (qr_core.cuh);
 *   - a plane's solver only needs ONE coordinate of its position and the
 *     axis: the axis is part of the kind, the coordinate sits in aux, so a
 *     plane that is missed costs no access to its surface record at all;
 *   - aux of a bounding volume / a JUMP is the distance in BYTES from the
 *     element itself to the slot to continue at (a missed volume: the slot
 *     behind its array's last leaf in this list, tracer.cpp:4042-4054).
 *
 *   surface lists  w = surface << 7 | flags | kind
 *   light lists    w = light, aux = head of the light's shadow list
 *   clip lists     w = clipper surface << 7 | QR_KC_* bits,
 *                  aux = trnode's last element (array clippers)
 */
#if defined(__CUDACC__)
struct __align__(8) qr_kelem { uint32_t w; int32_t aux; };
#else
struct alignas(8) qr_kelem { uint32_t w; int32_t aux; };
#endif

#define QR_KEND         0xFFFFFFFFu     /* END element (kind bits = 15) */

#define QR_K_BV         0   /* bounding volume of an array (elm.data & 3 == 1); aux = skip distance */
#define QR_K_PLANE_X    1   /* srf_t[0] == 1, a_map[K] = 0 / 1 / 2; aux = pos[K] (float bits) */
#define QR_K_PLANE_Y    2
#define QR_K_PLANE_Z    3
#define QR_K_QUADRIC    4   /* srf_t[0] == 2 */
#define QR_K_TWOPLANE   5   /* srf_t[0] == 3 */
#define QR_K_PLANE_G    6   /* a plane with its own matrix: axis and sign from the descriptor */
#define QR_K_OPEN       7   /* array with a matrix: transform the ray into the node */
#define QR_K_CLOSE      8   /* back to the world ray */
#define QR_K_NOP        9   /* array without a matrix / surface without a solver */
#define QR_K_JUMP       10  /* continue at aux bytes from here */
#define QR_K_END        15
#define QR_K_KIND(w)    ((w) & 15u)

#define QR_KF_NODE      16u /* world ray in the scratch, local ray in registers (inside a node, or OWN) */
#define QR_KF_SPARE     32u
#define QR_KF_SGN       64u /* planes X / Y / Z: a_sgn[K] */
#define QR_KF_OWN       64u /* quadric / two-plane: its own matrix, outside any node */
#define QR_K_SURF_OFF(w) ((w) & ~127u)  /* byte offset of the surface record (128 B each) */

/*
 * Device-side tiling (qr_tiling.cuh): when the engine left the tiling to the
 * backend, the camera list's leaves are tabled here, in list order, with the
 * transform node they sit in; the device writes one compiled list per tile
 * from them.  Image sections behind the texels:
 *   kleaf table | qr_bound per surface | tile rectangle per leaf (device) || tile lists (device, not uploaded)
 * header: pad3[1] = off_kleaf, pad3[2] = n_kleaf, pad3[3] = off_rects,
 *         pad0[0] = off_tlists, pad0[1] = elements per tile list (capacity)
 */
#if defined(__CUDACC__)
struct __align__(16) qr_kleaf { uint32_t w; int32_t aux; uint32_t open; uint32_t bound; };
#else
struct alignas(16) qr_kleaf { uint32_t w; int32_t aux; uint32_t open; uint32_t bound; };
#endif
#define QR_KLEAF_NO_NODE  0xFFFFFFFFu   /* open: record offset of the leaf's open node, or none */
#define QR_KTILE_MAX_LEAVES   8192u     /* beyond: the untiled list is walked (bounding volumes cull) */
#define QR_KTILE_MAX_BYTES    (1536ull << 20)

#define QR_KC_NEG        1u /* clip lists: rt_ELEM.data < 0 (inner side / accum enter) */
#define QR_KC_ACCUM      2u /* clip lists: accum marker (no surface) */

/*
 * ksurf quads (traversal):
 *   q0  pos.x  pos.y  pos.z  desc
 *   q1  sci.x  sci.y  sci.z  sci.w
 *   q2  scj.x  scj.y  scj.z  props (outer | inner << 16)
 *   q3  min.x  min.y  min.z  clip_head      (min/max: -inf/+inf when switched off)
 *   q4  max.x  max.y  max.z  trnode (byte offset of its record)
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

/* host-side packer (plain inline code; never called from device code) */

static inline uint32_t qr_k_align16(uint32_t v) { return (v + 15u) & ~15u; }

static inline float qr_k_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/*
 * Two steps: plan() validates the lists of a (header-checked) blob, compiles
 * the per-element state and fixes the sequential element order, so bytes() is
 * exact; write() then builds the image straight into the destination (the
 * pinned staging buffer of qr_scene_upload).
 */
class qr_kpacker
{
    public:

    /* 0, or -1 when a list is malformed / not well nested (an element would be
     * reached both inside and outside an open transform node) */
    int plan(const void *blob_)
    {
        blob = (const uint8_t *)blob_;
        h  = (const qr_blob_header *)blob;
        sf = (const qr_surface  *)(blob + h->off_surf);
        el = (const qr_elem     *)(blob + h->off_elem);
        tl = (const int32_t     *)(blob + h->off_tiles);
        ne = h->n_elem;

        /* what the list compiler needs of a surface, in one word:
         * bits 1:0 srf_t[0], 2 array, 3 has a matrix, 4 field shift,
         * 6:5 a_map[K] (field shift taken off), 7 a_sgn[K] */
        sinfo.resize((size_t)h->n_surf);
        for (int i = 0; i < h->n_surf; i++)
        {
            const uint32_t mk = (uint32_t)(sf[i].a_map[2] - (sf[i].a_sgn[3] != 0 ? 3 : 0)) & 3u;
            sinfo[i] = (uint8_t)(((uint32_t)sf[i].srf_t[0] & 3u) | (sf[i].srf_t[3] < 0 ? 4u : 0u)
                               | (sf[i].a_map[3] != 0 ? 8u : 0u) | (sf[i].a_sgn[3] != 0 ? 16u : 0u)
                               | (mk << 5) | (sf[i].a_sgn[2] != 0 ? 128u : 0u));
        }

        out.clear();
        out.reserve((size_t)ne + (size_t)ne / 4u + (size_t)h->n_tiles + 4u * (size_t)h->n_surf + 64u);
        nidx.assign((size_t)ne, -1);
        follow.assign((size_t)ne, -1);
        state.assign((size_t)ne, QR_NIL);
        heads.clear();
        fix.clear();
        out.push_back(make(QR_KEND, 0));                /* element 0: the NULL list */

        k_tiles.resize((size_t)h->n_tiles);
        for (int t = 0; t < h->n_tiles; t++)
        {
            if ((k_tiles[t] = emit_surf_list(tl[t])) < 0) return -1;
        }
        k_srf.assign((size_t)h->n_surf * 2, 0);
        k_lgt.assign((size_t)h->n_surf * 2, 0);
        k_clip.assign((size_t)h->n_surf, 0);
        for (int i = 0; i < h->n_surf; i++)
        {
            for (int sd = 0; sd < 2; sd++)
            {
                if ((k_srf[2 * i + sd] = emit_surf_list(sf[i].lst_srf[sd])) < 0) return -1;
                if ((k_lgt[2 * i + sd] = emit_copy_list(sf[i].lst_lgt[sd], true)) < 0) return -1;
            }
            if ((k_clip[i] = emit_copy_list(sf[i].clip_head, false)) < 0) return -1;
        }
        out.push_back(make(QR_KEND, 0));                /* pad: element i + 1 is always loadable */

        /* bounding-volume skip slots: the slot right behind the array's last
         * leaf, in the emission that holds that leaf */
        for (size_t i = 0; i < fix.size(); i++)
        {
            const int32_t o = fix[i].leaf;
            if (o < 0 || o >= ne || follow[o] < 0) return -1;
            /* the volume is met in the node state its last leaf is met in; a
             * volume inside a node that ends with the node lands on a CLOSE */
            const int32_t after = state[o] == o ? QR_NIL : state[o];    /* node state behind the leaf */
            const int32_t want = fix[i].state == o ? QR_NIL : fix[i].state;
            if (after != want) return -1;
            if (fix[i].state == o && out[(size_t)follow[o]].w != QR_K_CLOSE) return -1;
            /* a volume outside the node need not stop at the node's CLOSE */
            int32_t to = follow[o];
            if (fix[i].state == QR_NIL && out[(size_t)to].w == QR_K_CLOSE) to++;
            out[fix[i].at].aux = (to - fix[i].at) * (int32_t)sizeof(qr_kelem);
        }

        /* materials are deduplicated by content: the engine keeps one record
         * per surface side, most of them copies (RooT's default scene: 346
         * records, a few dozen distinct), and the material table is part of
         * the prefix a CTA stages in shared memory */
        {
            const qr_material *mt = (const qr_material *)(blob + h->off_mat);
            mat_map.assign((size_t)h->n_mat, 0);
            mat_uniq.clear();
            mat_bucket.clear();
            for (int i = 0; i < h->n_mat; i++)
            {
                uint64_t hv = 1469598103934665603ull;
                const uint32_t *w = (const uint32_t *)&mt[i];
                for (size_t j = 0; j < sizeof(qr_material) / 4; j++)
                {
                    hv = (hv ^ w[j]) * 1099511628211ull;
                }
                std::vector<int32_t> &b = mat_bucket[hv];
                int32_t u = -1;
                for (size_t j = 0; j < b.size() && u < 0; j++)
                {
                    if (memcmp(&mt[mat_uniq[(size_t)b[j]]], &mt[i], sizeof(qr_material)) == 0) u = b[j];
                }
                if (u < 0)
                {
                    u = (int32_t)mat_uniq.size();
                    mat_uniq.push_back(i);
                    b.push_back(u);
                }
                mat_map[(size_t)i] = u;
            }
        }

        k = *h;
        uint32_t off = sizeof(qr_blob_header);
        k.flags = QR_KSCENE_FLAG;
        k.n_mat = (int32_t)mat_uniq.size();
        k.off_surf = off;   off = qr_k_align16(off + (uint32_t)h->n_surf * QR_KSURF_QUADS * 16);
        k.pad3[0] = (int32_t)off; off = qr_k_align16(off + (uint32_t)h->n_surf * QR_KSHADE_QUADS * 16);
        k.off_mat = off;    off = qr_k_align16(off + (uint32_t)mat_uniq.size() * QR_KMAT_QUADS * 16);
        k.off_lgt = off;    off = qr_k_align16(off + (uint32_t)h->n_lgt * QR_KLGT_QUADS * 16);
        k.off_elem = off;   off = qr_k_align16(off + (uint32_t)out.size() * (uint32_t)sizeof(qr_kelem));
        k.n_elem = (int32_t)out.size();
        k.off_tiles = off;  off = qr_k_align16(off + (uint32_t)h->n_tiles * sizeof(int32_t));
        k.off_texels = off; off = qr_k_align16(off + (uint32_t)h->n_texels * sizeof(uint32_t));
        k.off_bounds = 0; k.n_bounds = 0;
        k.pad0[0] = 0; k.pad0[1] = 0; k.pad3[1] = 0; k.pad3[2] = 0; k.pad3[3] = 0;
        dev_bytes = off;

        /* device-side tiling: the engine left every tile head at the camera
         * list and sent the bounding boxes along */
        leaves.clear();
        if (plan_tiling())
        {
            k.pad3[1] = (int32_t)off; off = qr_k_align16(off + (uint32_t)(leaves.size() * sizeof(qr_kleaf)));
            k.pad3[2] = (int32_t)leaves.size();
            k.off_bounds = off;       off = qr_k_align16(off + (uint32_t)h->n_bounds * (uint32_t)sizeof(qr_bound));
            k.n_bounds = h->n_bounds;
            k.pad3[3] = (int32_t)off; off = qr_k_align16(off + (uint32_t)(leaves.size() * 16u));
            k.pad0[0] = (int32_t)off;
            k.pad0[1] = (int32_t)tile_cap;
            dev_bytes = (size_t)off + (size_t)h->n_tiles * tile_cap * sizeof(qr_kelem);
        }
        k.total_bytes = off;
        return 0;
    }

    /* bytes of the image that are written here and copied to the device */
    size_t bytes() const { return k.total_bytes; }

    /* bytes the device buffer needs: the image plus the tile lists the device writes */
    size_t device_bytes() const { return dev_bytes; }

    /* does the device build the tile lists (qr_tiling.cuh)? */
    bool device_tiling() const { return !leaves.empty(); }

    /* offset of the list elements = size of the part staged in shared memory */
    uint32_t prefix_bytes() const { return k.off_elem; }

    void write(void *out_) const
    {
        uint8_t *o = (uint8_t *)out_;
        const qr_material *mt = (const qr_material *)(blob + h->off_mat);
        const qr_light    *lg = (const qr_light    *)(blob + h->off_lgt);
        const float inf = qr_k_bits(0x7F800000u), ninf = qr_k_bits(0xFF800000u);

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

            /* an axis clipper that is switched off can never reject: the
             * tests are min <= x and !(max < x), tracer.cpp:1874-1927 */
            const int mm = s.minmax_t;
            const uint32_t props = ((uint32_t)s.props[0] & 0xFFFFu) | (((uint32_t)s.props[1] & 0xFFFFu) << 16);
            qr_f4 *q = ks + (size_t)i * QR_KSURF_QUADS;
            q[0].x = s.pos[0]; q[0].y = s.pos[1]; q[0].z = s.pos[2]; q[0].w = qr_k_bits(d);
            q[1].x = s.sci[0]; q[1].y = s.sci[1]; q[1].z = s.sci[2]; q[1].w = s.sci[3];
            q[2].x = s.scj[0]; q[2].y = s.scj[1]; q[2].z = s.scj[2]; q[2].w = qr_k_bits(props);
            q[3].x = (mm & 1) ? s.min[0] : ninf; q[3].y = (mm & 2) ? s.min[1] : ninf;
            q[3].z = (mm & 4) ? s.min[2] : ninf; q[3].w = qr_k_bits((uint32_t)k_clip[i]);
            q[4].x = (mm & 8) ? s.max[0] : inf;  q[4].y = (mm & 16) ? s.max[1] : inf;
            q[4].z = (mm & 32) ? s.max[2] : inf; q[4].w = qr_k_bits((uint32_t)s.trnode << 7);
            q[5].x = s.tci[0]; q[5].y = s.tci[1]; q[5].z = s.tci[2]; q[5].w = s.tcj[0];
            q[6].x = s.tcj[1]; q[6].y = s.tcj[2]; q[6].z = s.tck[0]; q[6].w = s.tck[1];
            q[7].x = s.tck[2]; q[7].y = s.d_eps;  q[7].z = s.t_eps;  q[7].w = qr_k_bits(s.c_def);

            qr_f4 *g = kh + (size_t)i * QR_KSHADE_QUADS;
            g[0].x = qr_k_bits((uint32_t)kmat_of(s.mat[0])); g[0].y = qr_k_bits((uint32_t)kmat_of(s.mat[1]));
            g[0].z = qr_k_bits((uint32_t)k_lgt[2 * i]);     g[0].w = qr_k_bits((uint32_t)k_lgt[2 * i + 1]);
            g[1].x = qr_k_bits((uint32_t)k_srf[2 * i]);     g[1].y = qr_k_bits((uint32_t)k_srf[2 * i + 1]);
            g[1].z = 0.0f; g[1].w = 0.0f;
        }

        qr_f4 *km = (qr_f4 *)(o + k.off_mat);
        for (size_t i = 0; i < mat_uniq.size(); i++)
        {
            const qr_material &m = mt[mat_uniq[i]];
            qr_f4 *q = km + i * QR_KMAT_QUADS;
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

        memcpy(o + k.off_elem, out.data(), out.size() * sizeof(qr_kelem));
        memcpy(o + k.off_tiles, k_tiles.data(), (size_t)h->n_tiles * sizeof(int32_t));
        memcpy(o + k.off_texels, blob + h->off_texels, (size_t)h->n_texels * sizeof(uint32_t));
        if (!leaves.empty())
        {
            memcpy(o + (uint32_t)k.pad3[1], leaves.data(), leaves.size() * sizeof(qr_kleaf));
            memcpy(o + k.off_bounds, blob + h->off_bounds, (size_t)h->n_bounds * sizeof(qr_bound));
            memset(o + (uint32_t)k.pad3[3], 0, leaves.size() * 16u);
        }
    }

    private:

    static qr_kelem make(uint32_t w, int32_t aux) { qr_kelem e; e.w = w; e.aux = aux; return e; }

    /*
     * Device-side tiling applies when the blob carries bounding boxes and every
     * tile head is the same list (the camera list): table its leaves with the
     * node each one sits in, by walking the COMPILED list -- bounding-volume
     * elements are dropped as the reference's tile lists drop them
     * (engine.cpp:3160-3167), OPEN / CLOSE become the leaf's "open" field.
     */
    bool plan_tiling()
    {
        if (h->off_bounds == 0 || h->n_bounds != h->n_surf || h->n_tiles < 2) return false;
        if ((uint64_t)h->off_bounds + (uint64_t)h->n_bounds * sizeof(qr_bound) > h->total_bytes) return false;
        for (int t = 1; t < h->n_tiles; t++)
        {
            if (tl[t] != tl[0]) return false;
        }
        if (k_tiles[0] <= 0) return false;
        uint32_t open = QR_KLEAF_NO_NODE;
        size_t guard = 0;
        for (size_t i = (size_t)k_tiles[0]; ; )
        {
            if (i >= out.size() || guard++ > out.size()) { leaves.clear(); return false; }
            const qr_kelem e = out[i];
            const uint32_t kind = QR_K_KIND(e.w);
            if (e.w == QR_KEND) break;
            if (kind == QR_K_JUMP) { i = (size_t)((int64_t)i + e.aux / (int32_t)sizeof(qr_kelem)); continue; }
            if (kind == QR_K_OPEN)  open = QR_K_SURF_OFF(e.w);
            else
            if (kind == QR_K_CLOSE) open = QR_KLEAF_NO_NODE;
            else
            if (kind != QR_K_BV && kind != QR_K_NOP)
            {
                qr_kleaf lf;
                lf.w = e.w; lf.aux = e.aux;
                lf.open = (e.w & QR_KF_NODE) ? open : QR_KLEAF_NO_NODE;
                lf.bound = QR_K_SURF_OFF(e.w) >> 7;
                const uint32_t kd = kind;
                const bool own = kd == QR_K_PLANE_G || ((kd == QR_K_QUADRIC || kd == QR_K_TWOPLANE) && (e.w & QR_KF_OWN));
                if (own) lf.open = QR_KLEAF_NO_NODE;
                leaves.push_back(lf);
            }
            i++;
        }
        if (leaves.empty() || leaves.size() > QR_KTILE_MAX_LEAVES) { leaves.clear(); return false; }
        /* elements a tile list can need: its leaves, an OPEN and a CLOSE per run
         * of leaves of one node (a sub-sequence never has more runs), a CLOSE
         * per leaf with its own matrix, the END */
        size_t runs = 0, own = 0;
        for (size_t i = 0; i < leaves.size(); i++)
        {
            if (leaves[i].open != QR_KLEAF_NO_NODE && (i == 0 || leaves[i].open != leaves[i - 1].open)) runs++;
            const uint32_t kd = QR_K_KIND(leaves[i].w);
            if (kd == QR_K_PLANE_G || ((kd == QR_K_QUADRIC || kd == QR_K_TWOPLANE) && (leaves[i].w & QR_KF_OWN))) own++;
        }
        tile_cap = (uint32_t)(leaves.size() + 2 * runs + own + 1);
        if ((uint64_t)h->n_tiles * tile_cap * sizeof(qr_kelem) > QR_KTILE_MAX_BYTES) { leaves.clear(); return false; }
        return true;
    }

    /* index of a blob material in the deduplicated table */
    int32_t kmat_of(int32_t m) const
    {
        return m >= 0 && m < h->n_mat ? mat_map[(size_t)m] : m;
    }

    /*
     * Surface list: compiled and emitted once; a tail shared with a list
     * emitted earlier is entered through a JUMP.  "lobj" is the last element
     * of the open transform node while the list is walked (QR_NIL: none) --
     * the state the reference keeps in ctx_LOCAL(OBJ); state[i] records it
     * per element so that every way of reaching an element agrees on it.
     * follow[i] is the slot behind element i: where a missed bounding volume
     * whose array ends with i continues.
     */
    int32_t emit_surf_list(int32_t head)
    {
        if (head == QR_NIL) return 0;
        if (head < 0 || head >= ne) return -1;
        if (nidx[head] >= 0) return state[head] == QR_NIL ? nidx[head] : -1;
        const int32_t first = (int32_t)out.size();
        int32_t lobj = QR_NIL;
        for (int32_t i = head; ; )
        {
            if (i == QR_NIL)
            {
                if (lobj != QR_NIL) return -1;          /* list ends inside an open node */
                out.push_back(make(QR_KEND, 0));
                break;
            }
            if (i < 0 || i >= ne) return -1;
            if (nidx[i] >= 0)
            {
                if (state[i] != lobj) return -1;
                const int32_t at = (int32_t)out.size();
                out.push_back(make(QR_K_JUMP, (nidx[i] - at) * (int32_t)sizeof(qr_kelem)));
                break;
            }
            const qr_elem &e = el[i];
            if (e.simd < 0 || e.simd >= h->n_surf) return -1;
            const uint32_t si = sinfo[e.simd];
            const bool is_array = (si & 4u) != 0, has_mtx = (si & 8u) != 0;
            const uint32_t rec = (uint32_t)e.simd << 7;
            const uint32_t tag = is_array ? 0u : (si & 3u);
            nidx[i] = (int32_t)out.size();
            state[i] = lobj;

            uint32_t f = QR_K_NOP;
            int32_t  aux = 0;
            bool     close_after = false;
            if (e.data_i == 1)
            {
                /* bounding volume of an array; aux is fixed up at the end */
                if (has_mtx && is_array) return -1;     /* a bounding volume has no matrix of its own */
                if (e.data_p < 0 || e.data_p >= ne) return -1;
                f = QR_K_BV;
                qr_kfix x;
                x.at = (int32_t)out.size();
                x.leaf = e.data_p;                      /* tracer.cpp:4042-4054 */
                x.state = lobj;
                fix.push_back(x);
            }
            else
            if (is_array && has_mtx)
            {
                /* nodes do not nest; should one ever open inside another,
                 * it starts from the world as the reference's walk does */
                if (lobj != QR_NIL) out.push_back(make(QR_K_CLOSE, 0));
                f = QR_K_OPEN;
                lobj = e.data_p;                        /* tracer.cpp:1492-1496 */
                if (lobj < 0 || lobj >= ne) return -1;
            }
            else
            if (is_array)
            {
                if (lobj != QR_NIL) return -1;          /* plain array inside an open node */
            }
            else
            {
                const bool in_node = lobj != QR_NIL;
                const bool own = !in_node && has_mtx && tag != 0;
                if (in_node && !(si & 16u)) return -1;  /* child without the field shift */
                if (tag == 1u)
                {
                    /* PL_ptr 4062-4136: axis, sign and pos[K] travel with the element */
                    const uint32_t mk = (si >> 5) & 3u;
                    if (mk > 2u) return -1;
                    if (own)
                    {
                        f = QR_K_PLANE_G | QR_KF_OWN;
                    }
                    else
                    {
                        f = (QR_K_PLANE_X + mk) | ((si & 128u) ? QR_KF_SGN : 0u);
                        memcpy(&aux, &sf[e.simd].pos[mk], 4);
                    }
                }
                else
                if (tag == 2u) f = QR_K_QUADRIC | (own ? QR_KF_OWN : 0u);
                else
                if (tag == 3u) f = QR_K_TWOPLANE | (own ? QR_KF_OWN : 0u);
                if (in_node || own) f |= QR_KF_NODE;
                if (own) close_after = true;
                if (in_node && i == lobj)
                {
                    close_after = true;
                    lobj = QR_NIL;
                }
            }
            out.push_back(make(rec | f, aux));
            /* a skip from inside the node lands ON the CLOSE */
            follow[i] = (int32_t)out.size();
            if (close_after) out.push_back(make(QR_K_CLOSE, 0));
            i = e.next;
        }
        return first;
    }

    /* light / clip list: one private sequential copy per distinct head */
    int32_t emit_copy_list(int32_t head, bool lights)
    {
        if (head == QR_NIL) return 0;
        if (head < 0 || head >= ne) return -1;
        {
            std::unordered_map<int32_t, int32_t>::const_iterator it = heads.find(head);
            if (it != heads.end()) return it->second;
        }
        /* shadow lists first: their indices go into the copy */
        shadow.clear();
        int guard = 0;
        if (lights)
        {
            for (int32_t i = head; i != QR_NIL; i = el[i].next)
            {
                if (i < 0 || i >= ne || guard++ > ne) return -1;
                const int32_t sh = emit_surf_list(el[i].data_p);
                if (sh < 0) return -1;
                shadow.push_back(sh);
            }
        }
        const int32_t first = (int32_t)out.size();
        size_t n = 0;
        guard = 0;
        for (int32_t i = head; i != QR_NIL; i = el[i].next, n++)
        {
            if (i < 0 || i >= ne || guard++ > ne) return -1;
            const qr_elem &e = el[i];
            if (lights)
            {
                if (e.simd < 0 || e.simd >= h->n_lgt) return -1;
                out.push_back(make((uint32_t)e.simd, shadow[n]));
                continue;
            }
            if (e.simd == QR_NIL)
            {
                out.push_back(make(QR_KC_ACCUM | (e.data_i < 0 ? QR_KC_NEG : 0u), 0));
                continue;
            }
            if (e.simd < 0 || e.simd >= h->n_surf) return -1;
            /* array clipper: data_p = the trnode's last element, further down this list */
            int32_t last = 0;
            if (e.data_p != QR_NIL)
            {
                int32_t pos = 0, g2 = 0;
                int32_t j = head;
                for (; j != QR_NIL && j != e.data_p; j = el[j].next, pos++)
                {
                    if (j < 0 || j >= ne || g2++ > ne) return -1;
                }
                if (j == QR_NIL) return -1;
                last = first + pos;
            }
            out.push_back(make(((uint32_t)e.simd << 7) | (e.data_i < 0 ? QR_KC_NEG : 0u), last));
        }
        out.push_back(make(QR_KEND, 0));
        heads[head] = first;
        return first;
    }

    const uint8_t        *blob;
    const qr_blob_header *h;
    const qr_surface     *sf;
    const qr_elem        *el;
    const int32_t        *tl;
    int                   ne;
    qr_blob_header        k;
    struct qr_kfix { int32_t at, leaf, state; };        /* BV slot, the array's last leaf, node state at the BV */
    std::vector<uint8_t>  sinfo;
    std::vector<int32_t>  state, nidx, follow, k_tiles, k_srf, k_lgt, k_clip, shadow;
    std::vector<int32_t>  mat_map, mat_uniq;            /* blob material -> table index, table index -> blob material */
    std::unordered_map<uint64_t, std::vector<int32_t> > mat_bucket;
    std::vector<qr_kelem> out;
    std::vector<qr_kleaf> leaves;                       /* camera list's leaves (device tiling), or empty */
    uint32_t              tile_cap;
    size_t                dev_bytes;
    std::vector<qr_kfix>  fix;
    std::unordered_map<int32_t, int32_t> heads;
};

#endif /* QR_KSCENE_H */
