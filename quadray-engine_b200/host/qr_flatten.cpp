/*
 * qr_flatten.cpp -- engine pointer graph -> index-based scene blob.
 *
 * Field-by-field this follows SURVEY.md appendix B (what render0 reads):
 *   rt_SIMD_INFOX     core/tracer/tracer.h:150-216
 *   rt_SIMD_CONTEXT   tracer.h:426-665   (level-0 inputs, engine.cpp:3588-3596)
 *   rt_SIMD_CAMERA    tracer.h:677-755   (engine.cpp:3559-3584)
 *   rt_SIMD_LIGHT     tracer.h:765-811   (object.cpp:590-670)
 *   rt_SIMD_SURFACE   tracer.h:821-969   (object.cpp:712-759, 2472-2503)
 *   rt_SIMD_MATERIAL  tracer.h:979-1078  (object.cpp:4094-4151)
 *   rt_ELEM           tracer.h:127-141   (encodings: engine.cpp:1120-1132,
 *                                         1681-1695, 1845-1865, 3201-3208)
 * The byte-offset tricks of the macro assembler (a_map/a_sgn/t_map hold
 * offsets of SIMD fields, object.cpp:2489-2497, 4099-4100) are turned back
 * into small indices here.
 */

#include <string.h>
#include <utility>

#include "tracer.h"
#include "format.h"
#include "engine.h"

#include "qr_flatten.h"

#define QR_FIELD (RT_SIMD_QUADS * 16)   /* bytes of one SIMD field */

/*
 * The lists and SIMD structs were written a moment ago by the update phases on
 * OTHER cores, so nearly every record this thread touches is a cache-to-cache
 * miss; what can be named ahead of time is prefetched.
 */
#if defined(__GNUC__)
#define QR_PREFETCH(p) __builtin_prefetch((p), 0, 1)
#else
#define QR_PREFETCH(p) ((void)0)
#endif

static inline void qr_prefetch_range(const void *p, size_t bytes)
{
    const char *c = (const char *)p;
    for (size_t o = 0; o < bytes; o += 64)
    {
        QR_PREFETCH(c + o);
    }
}

/*
 * Index of a list element, assigned on first sight.  Elements are only NAMED
 * here and read later, in index order, by drain(): every record is then
 * known a few iterations before it is read and can be prefetched, and no
 * chain is chased through cold memory.  (The blob's element order is
 * breadth-first as a result; the kernel image re-sequences the lists anyway,
 * qr_kscene.h.)
 */
int32_t qr_Flattener::list_head(const rt_ELEM *head, ListKind kind)
{
    if (head == RT_NULL)
    {
        return QR_NIL;
    }
    const int32_t next = (int32_t)elem_src.size();
    const int32_t idx = elem_idx.insert(head, next);
    if (idx == next)
    {
        elem_src.push_back(head);
        elem_kind.push_back(kind);
    }
    return idx;
}

int32_t qr_Flattener::surface(const rt_SIMD_SURFACE *s)
{
    if (s == RT_NULL)
    {
        return QR_NIL;
    }
    const int32_t next = (int32_t)surf_src.size();
    const int32_t idx = surf_idx.insert(s, next);
    if (idx == next)
    {
        surf_src.push_back(s);
        /* what the element loop asks of a surface, kept beside the index so
         * the loop does not touch the 4.7 KB record once per element */
        surf_array.push_back(s->srf_t[3] < 0);
    }
    return idx;
}

int32_t qr_Flattener::material(const rt_SIMD_MATERIAL *m)
{
    if (m == RT_NULL)
    {
        return QR_NIL;
    }
    const int32_t known = mat_idx.find(m);
    if (known >= 0)
    {
        return known;
    }

    qr_material r;
    memset(&r, 0, sizeof(r));
    r.xscal = m->xscal[0];
    r.yscal = m->yscal[0];
    r.xoffs = m->xoffs[0];
    r.yoffs = m->yoffs[0];
    r.xmask = (uint32_t)m->xmask[0];
    r.ymask = (uint32_t)m->ymask[0];
    r.yshft = (uint32_t)m->yshft[0];
    r.t_map[0] = m->t_map[0] / QR_FIELD;
    r.t_map[1] = m->t_map[1] / QR_FIELD;
    r.l_dff = m->l_dff[0];
    r.l_spc = m->l_spc[0];
    r.l_pow = m->l_pow[0];
    r.c_rfl = m->c_rfl[0];
    r.c_trn = m->c_trn[0];
    r.c_rfr = m->c_rfr[0];
    r.rfr_2 = m->rfr_2[0];
    r.c_rcp = m->c_rcp[0];
    r.ext_2 = m->ext_2[0];
    r.clamp = m->clamp[0];
    r.cmask = (uint32_t)m->cmask[0];
    r.col[0] = m->col_r[0]; r.col[1] = m->col_g[0]; r.col[2] = m->col_b[0];

    /* texture: (xmask+1) x (ymask+1) texels of 0x00RRGGBB, power of two
     * (object.cpp:4113-4129); a plain colour is a 1x1 texture */
    const rt_ui32 *tex = (const rt_ui32 *)m->tex_p[0];
    if (tex == RT_NULL)
    {
        throw rt_Exception("null texture pointer in qr_Flattener");
    }
    r.tex = tex_idx.find(tex);
    if (r.tex < 0)
    {
        size_t n = (size_t)(r.xmask + 1) * (size_t)(r.ymask + 1);
        r.tex = (int32_t)texels.size();
        tex_idx.insert(tex, r.tex);
        texels.insert(texels.end(), tex, tex + n);
    }

    int32_t idx = (int32_t)mats.size();
    mat_idx.insert(m, idx);
    mats.push_back(r);
    return idx;
}

int32_t qr_Flattener::light(const rt_SIMD_LIGHT *l)
{
    if (l == RT_NULL)
    {
        return QR_NIL;
    }
    const int32_t known = lgt_idx.find(l);
    if (known >= 0)
    {
        return known;
    }

    qr_light r;
    memset(&r, 0, sizeof(r));
    r.t_max  = l->t_max[0];
    r.pos[0] = l->pos_x[0];
    r.pos[1] = l->pos_y[0];
    r.pos[2] = l->pos_z[0];
    r.col[0] = l->col_r[0];
    r.col[1] = l->col_g[0];
    r.col[2] = l->col_b[0];
    r.a_qdr  = l->a_qdr[0];
    r.a_lnr  = l->a_lnr[0];
    r.a_cnt  = l->a_cnt[0];

    int32_t idx = (int32_t)lgts.size();
    lgt_idx.insert(l, idx);
    lgts.push_back(r);
    return idx;
}

/*
 * Resolve everything reachable from the lists and surfaces queued so far.
 */
void qr_Flattener::drain()
{
    size_t elem_done = elems.size();

    while (elem_done < elem_src.size() || surf_done < surf_src.size())
    {
        for (; elem_done < elem_src.size(); elem_done++)
        {
            if (elem_done + 8 < elem_src.size())
            {
                QR_PREFETCH(elem_src[elem_done + 8]);
            }
            const rt_ELEM *e = elem_src[elem_done];
            ListKind kind = elem_kind[elem_done];
            qr_elem r;
            r.data_i = 0;
            r.data_p = QR_NIL;
            r.simd   = QR_NIL;
            r.next   = list_head(e->next, kind);

            if (kind == LIST_SURF)
            {
                const rt_SIMD_SURFACE *s = (const rt_SIMD_SURFACE *)e->simd;
                r.simd   = surface(s);
                r.data_i = (int32_t)(e->data & 3);
                const rt_ELEM *last = (const rt_ELEM *)(e->data & ~(rt_cell)3);
                if (last != RT_NULL && s != RT_NULL
                &&  (r.data_i == 1 || surf_array[(size_t)r.simd]))
                {
                    r.data_p = list_head(last, LIST_SURF);
                }
            }
            else
            if (kind == LIST_LIGHT)
            {
                r.simd   = light((const rt_SIMD_LIGHT *)e->simd);
                r.data_p = list_head((const rt_ELEM *)e->data, LIST_SURF);
            }
            else
            {
                const rt_SIMD_SURFACE *s = (const rt_SIMD_SURFACE *)e->simd;
                r.simd = surface(s);
                if (s != RT_NULL && surf_array[(size_t)r.simd])
                {
                    r.data_p = list_head((const rt_ELEM *)e->data, LIST_CLIP);
                }
                else
                {
                    r.data_i = (int32_t)e->data;
                }
            }

            elems.push_back(r);
        }

        for (; surf_done < surf_src.size(); surf_done++)
        {
            /* two records ahead: the surface; one ahead (its pointers are in
             * cache by now): its two materials */
            if (surf_done + 2 < surf_src.size())
            {
                qr_prefetch_range(surf_src[surf_done + 2], sizeof(rt_SIMD_SURFACE));
            }
            if (surf_done + 1 < surf_src.size())
            {
                const rt_SIMD_SURFACE *n = surf_src[surf_done + 1];
                if (n->mat_p[0] != RT_NULL) qr_prefetch_range(n->mat_p[0], sizeof(rt_SIMD_MATERIAL));
                if (n->mat_p[2] != RT_NULL) qr_prefetch_range(n->mat_p[2], sizeof(rt_SIMD_MATERIAL));
            }
            const rt_SIMD_SURFACE *s = surf_src[surf_done];
            qr_surface r;
            memset(&r, 0, sizeof(r));

            r.pos[0] = s->pos_x[0]; r.pos[1] = s->pos_y[0]; r.pos[2] = s->pos_z[0];
            r.min[0] = s->min_x[0]; r.min[1] = s->min_y[0]; r.min[2] = s->min_z[0];
            r.max[0] = s->max_x[0]; r.max[1] = s->max_y[0]; r.max[2] = s->max_z[0];
            r.d_eps  = s->d_eps[0];
            r.t_eps  = s->t_eps[0];
            r.minmax_t = 0;
            for (int a = 0; a < 3; a++)
            {
                r.minmax_t |= (s->min_t[a] != 0 ? 1 : 0) << a;
                r.minmax_t |= (s->max_t[a] != 0 ? 1 : 0) << (3 + a);
            }
            r.tci[0] = s->tci_x[0]; r.tci[1] = s->tci_y[0]; r.tci[2] = s->tci_z[0];
            r.tcj[0] = s->tcj_x[0]; r.tcj[1] = s->tcj_y[0]; r.tcj[2] = s->tcj_z[0];
            r.tck[0] = s->tck_x[0]; r.tck[1] = s->tck_y[0]; r.tck[2] = s->tck_z[0];
            r.sci[0] = s->sci_x[0]; r.sci[1] = s->sci_y[0]; r.sci[2] = s->sci_z[0];
            r.sci[3] = s->sci_w[0];
            r.scj[0] = s->scj_x[0]; r.scj[1] = s->scj_y[0]; r.scj[2] = s->scj_z[0];
            r.c_def  = (uint32_t)s->c_def[0];

            /* byte offsets of SIMD fields -> indices, object.cpp:2489-2497 */
            r.a_map[RT_I] = s->a_map[RT_I] / QR_FIELD;
            r.a_map[RT_J] = s->a_map[RT_J] / QR_FIELD;
            r.a_map[RT_K] = s->a_map[RT_K] / QR_FIELD;
            r.a_map[RT_L] = s->a_map[RT_L];
            r.a_sgn[RT_I] = s->a_sgn[RT_I] / QR_FIELD;
            r.a_sgn[RT_J] = s->a_sgn[RT_J] / QR_FIELD;
            r.a_sgn[RT_K] = s->a_sgn[RT_K] / QR_FIELD;
            r.a_sgn[RT_L] = s->a_sgn[RT_L] / QR_FIELD;

            for (int t = 0; t < 4; t++)
            {
                r.srf_t[t] = s->srf_t[t];
            }
            r.conic     = (int32_t)(rt_cell)s->msc_p[1];
            r.clip_head = list_head((const rt_ELEM *)s->msc_p[2], LIST_CLIP);
            r.trnode    = surface((const rt_SIMD_SURFACE *)s->msc_p[3]);

            r.mat[0]    = material((const rt_SIMD_MATERIAL *)s->mat_p[0]);
            r.props[0]  = (int32_t)(rt_cell)s->mat_p[1];
            r.mat[1]    = material((const rt_SIMD_MATERIAL *)s->mat_p[2]);
            r.props[1]  = (int32_t)(rt_cell)s->mat_p[3];

            r.lst_lgt[0] = list_head((const rt_ELEM *)s->lst_p[0], LIST_LIGHT);
            r.lst_srf[0] = list_head((const rt_ELEM *)s->lst_p[1], LIST_SURF);
            r.lst_lgt[1] = list_head((const rt_ELEM *)s->lst_p[2], LIST_LIGHT);
            r.lst_srf[1] = list_head((const rt_ELEM *)s->lst_p[3], LIST_SURF);

            surfs.push_back(r);
        }
    }

}

static uint32_t align16(uint32_t v)
{
    return (v + 15u) & ~15u;
}

const uint8_t *qr_Flattener::build(const rt_SIMD_INFOX *s_inf, size_t *bytes)
{
    const rt_SIMD_CONTEXT *s_ctx = (const rt_SIMD_CONTEXT *)s_inf->ctx;
    const rt_SIMD_CAMERA  *s_cam = (const rt_SIMD_CAMERA *)s_inf->cam;

    if (s_ctx == RT_NULL || s_cam == RT_NULL || s_inf->tiles == RT_NULL)
    {
        throw rt_Exception("null-pointer in qr_Flattener::build");
    }

    elem_idx.clear(); surf_idx.clear(); mat_idx.clear();
    lgt_idx.clear();  tex_idx.clear();
    {
        /* last frame's sizes are the best guess for this frame's */
        const size_t ne = elem_src.size() + 1024;
        elem_src.clear(); elem_kind.clear(); surf_src.clear(); surf_array.clear();
        elem_src.reserve(ne); elem_kind.reserve(ne); elems.reserve(ne);
    }
    elems.clear(); surfs.clear(); mats.clear(); lgts.clear();
    texels.clear(); tiles.clear();
    surf_done = 0;

    qr_blob_header h;
    memset(&h, 0, sizeof(h));
    h.magic   = QR_BLOB_MAGIC;
    h.version = QR_BLOB_VERSION;
    h.flags   = s_inf->pt_on != 0 ? QR_BLOB_PT : 0u;

    h.x_res   = (int32_t)s_inf->frm_w;
    h.y_res   = (int32_t)s_inf->frm_h;
    h.x_row   = (int32_t)s_inf->frm_row;
    h.fsaa    = (int32_t)s_inf->fsaa;
    h.depth   = (int32_t)s_inf->depth;
    h.tile_w  = (int32_t)s_inf->tile_w;
    h.tile_h  = (int32_t)s_inf->tile_h;
    h.tls_row = (int32_t)s_inf->tls_row;
    h.tls_col = (h.y_res + h.tile_h - 1) / h.tile_h;
    h.lst_head = QR_NIL; /* inf_LST is only read when RT_FEAT_TILING == 0 */

    h.ctx_flags = (uint32_t)s_ctx->param[1];
    h.t_min  = s_ctx->t_min[0];
    h.org[0] = s_ctx->org_x[0];
    h.org[1] = s_ctx->org_y[0];
    h.org[2] = s_ctx->org_z[0];

    h.cam_t_max = s_cam->t_max[0];
    h.dir[0] = s_cam->dir_x[0]; h.dir[1] = s_cam->dir_y[0]; h.dir[2] = s_cam->dir_z[0];
    h.hor[0] = s_cam->hor_x[0]; h.hor[1] = s_cam->hor_y[0]; h.hor[2] = s_cam->hor_z[0];
    h.ver[0] = s_cam->ver_x[0]; h.ver[1] = s_cam->ver_y[0]; h.ver[2] = s_cam->ver_z[0];
    for (int i = 0; i < 4; i++)
    {
        h.hor_a[i] = s_cam->hor_a[i];
        h.ver_a[i] = s_cam->ver_a[i];
    }
    h.amb[0] = s_cam->col_r[0];
    h.amb[1] = s_cam->col_g[0];
    h.amb[2] = s_cam->col_b[0];
    h.cam_clamp = s_cam->clamp[0];
    h.cam_cmask = (uint32_t)s_cam->cmask[0];

    /* the integer pixel indices of lane i (engine.cpp:3465-3550) are
     * re-derived on the device from fsaa; make sure they are what we expect */
    {
        const int lane[3][4] = { {0, 1, 2, 3}, {0, 0, 1, 1}, {0, 0, 0, 0} };
        if (h.fsaa < 0 || h.fsaa > 2)
        {
            throw rt_Exception("unsupported fsaa mode in qr_Flattener::build");
        }
        for (int i = 0; i < 4; i++)
        {
            if (s_inf->hor_c[i] != (rt_real)lane[h.fsaa][i])
            {
                throw rt_Exception("unexpected hor_c in qr_Flattener::build");
            }
        }
    }

    /* tiles */
    rt_ELEM **tl = (rt_ELEM **)s_inf->tiles;
    int32_t n_tiles = h.tls_row * h.tls_col;
    tiles.resize(n_tiles);

    /* RT_OPTS_TILING off: every tile head is the camera list
     * (engine.cpp:3236-3248).  The bounding boxes the engine's stile() would
     * have projected travel with the blob, and the device culls per tile. */
    bool untiled = n_tiles > 1 && tl[0] != RT_NULL;
    for (int32_t t = 1; t < n_tiles && untiled; t++)
    {
        untiled = tl[t] == tl[0];
    }
    if (untiled)
    {
        /* one list, named once */
        const int32_t head = list_head(tl[0], LIST_SURF);
        for (int32_t t = 0; t < n_tiles; t++)
        {
            tiles[t] = head;
        }
    }
    else
    {
        for (int32_t t = 0; t < n_tiles; t++)
        {
            if (t + 8 < n_tiles && tl[t + 8] != RT_NULL)
            {
                QR_PREFETCH(tl[t + 8]);
            }
            tiles[t] = list_head(tl[t], LIST_SURF);
        }
    }
    std::vector<std::pair<int32_t, const rt_BOUND *> > clist_bounds;
    if (untiled)
    {
        for (const rt_ELEM *e = tl[0]; e != RT_NULL; e = e->next)
        {
            const rt_SIMD_SURFACE *s = (const rt_SIMD_SURFACE *)e->simd;
            if (s == RT_NULL || e->temp == RT_NULL)
            {
                untiled = false;
                break;
            }
            clist_bounds.push_back(std::make_pair(surface(s), (const rt_BOUND *)e->temp));
        }
    }
    drain();

    bounds.clear();
    if (untiled)
    {
        qr_bound none;
        memset(&none, 0, sizeof(none));
        none.n = -1;
        bounds.assign(surfs.size(), none);
        for (size_t i = 0; i < clist_bounds.size(); i++)
        {
            const rt_BOUND *b = clist_bounds[i].second;
            qr_bound &r = bounds[(size_t)clist_bounds[i].first];
            /* more vertices than a box has: treated as unbounded */
            r.n = b->verts_num >= 0 && b->verts_num <= 8 && (b->verts != RT_NULL || b->verts_num == 0)
                ? (int32_t)b->verts_num : 0;
            for (int k = 0; k < r.n; k++)
            {
                r.v[k][0] = b->verts[k].pos[RT_X];
                r.v[k][1] = b->verts[k].pos[RT_Y];
                r.v[k][2] = b->verts[k].pos[RT_Z];
            }
        }
    }

    h.n_surf   = (int32_t)surfs.size();
    h.n_mat    = (int32_t)mats.size();
    h.n_lgt    = (int32_t)lgts.size();
    h.n_elem   = (int32_t)elems.size();
    h.n_tiles  = n_tiles;
    h.n_texels = (int32_t)texels.size();

    uint32_t off = sizeof(qr_blob_header);
    h.off_surf   = off; off = align16(off + h.n_surf   * sizeof(qr_surface));
    h.off_mat    = off; off = align16(off + h.n_mat    * sizeof(qr_material));
    h.off_lgt    = off; off = align16(off + h.n_lgt    * sizeof(qr_light));
    h.off_elem   = off; off = align16(off + h.n_elem   * sizeof(qr_elem));
    h.off_tiles  = off; off = align16(off + h.n_tiles  * sizeof(int32_t));
    h.off_texels = off; off = align16(off + h.n_texels * sizeof(uint32_t));
    if (!bounds.empty())
    {
        h.off_bounds = off;
        h.n_bounds = (int32_t)bounds.size();
        off = align16(off + (uint32_t)(bounds.size() * sizeof(qr_bound)));
    }
    h.total_bytes = off;

    blob.resize(off);
    {
        /* the sections are copied below; only the alignment gaps need zeroing */
        const uint32_t ends[7] = { (uint32_t)sizeof(h),
            h.off_surf   + h.n_surf   * (uint32_t)sizeof(qr_surface),
            h.off_mat    + h.n_mat    * (uint32_t)sizeof(qr_material),
            h.off_lgt    + h.n_lgt    * (uint32_t)sizeof(qr_light),
            h.off_elem   + h.n_elem   * (uint32_t)sizeof(qr_elem),
            h.off_tiles  + h.n_tiles  * (uint32_t)sizeof(int32_t),
            h.off_texels + h.n_texels * (uint32_t)sizeof(uint32_t) };
        const uint32_t nexts[7] = { h.off_surf, h.off_mat, h.off_lgt, h.off_elem, h.off_tiles, h.off_texels,
                                    h.n_bounds ? h.off_bounds : off };
        for (int k = 0; k < 7; k++)
        {
            if (nexts[k] > ends[k]) memset(&blob[ends[k]], 0, nexts[k] - ends[k]);
        }
    }
    memcpy(&blob[0], &h, sizeof(h));
    if (h.n_surf)   memcpy(&blob[h.off_surf],   &surfs[0],  h.n_surf   * sizeof(qr_surface));
    if (h.n_mat)    memcpy(&blob[h.off_mat],    &mats[0],   h.n_mat    * sizeof(qr_material));
    if (h.n_lgt)    memcpy(&blob[h.off_lgt],    &lgts[0],   h.n_lgt    * sizeof(qr_light));
    if (h.n_elem)   memcpy(&blob[h.off_elem],   &elems[0],  h.n_elem   * sizeof(qr_elem));
    if (h.n_tiles)  memcpy(&blob[h.off_tiles],  &tiles[0],  h.n_tiles  * sizeof(int32_t));
    if (h.n_texels) memcpy(&blob[h.off_texels], &texels[0], h.n_texels * sizeof(uint32_t));
    if (h.n_bounds) memcpy(&blob[h.off_bounds], &bounds[0], bounds.size() * sizeof(qr_bound));

    *bytes = blob.size();
    return &blob[0];
}
