/*
 * qr_tiling.cuh -- screen tiling on the device.
 *
 * The reference's engine culls the camera list per screen tile on the host,
 * every frame: rt_SceneThread::stile projects each surface's bounding box onto
 * the tile buffer (core/engine/engine.cpp:1956-2128, tiling() 962-1110), and
 * rt_Scene::render merges the per-surface tile spans into one list per tile
 * (engine.cpp:3129-3232): the camera list's order restricted to the surfaces
 * whose projection touches the tile, bounding-volume elements dropped,
 * surfaces of a transform node grouped behind one trnode element.  With ~8 000
 * tiles at 1080p that is ~160 k list elements to build, to flatten and to pack
 * per frame -- host time that caps the frame rate through rt_Scene::render.
 *
 * Here the engine runs with RT_OPTS_TILING off (every tile head is the camera
 * list, engine.cpp:3236-3248), the flattener sends the bounding-box vertices
 * along (qr_bound, include/qr_scene_blob.h), and two small kernels do the rest
 * per frame:
 *
 *   qr_tile_rect        per leaf of the camera list: the tile rectangle its
 *                       projected box can touch
 *   qr_tile_list_build  per tile: the list in the reference's order and
 *                       grouping, in the walk's compiled element format
 *
 * Parity: tiling is conservative culling -- a surface left out of a tile's
 * list cannot be hit by a primary ray of that tile -- and the order of the
 * survivors is the camera list's, so any SUPERSET of the reference's spans
 * renders the reference's pixels (the reference itself renders identical
 * frames with tiling on and off).  The rectangle computed here is a superset:
 * the bounding rectangle of the projected vertices contains every projected
 * edge the reference rasterises, it is widened by half a tile where the
 * reference widens by RT_TILE_THRESHOLD = 0.2 tile (engine.h:46), and a box
 * that reaches behind the screen plane is cut off a little further back than
 * where the reference cuts it.
 *
 * The file compiles for the device and, for the CPU tests, for the host
 * (tests/hostsim).
 */
#ifndef QR_TILING_CUH
#define QR_TILING_CUH

#include <stdint.h>
#include <math.h>
#include "qr_scene_blob.h"
#include "qr_kscene.h"

#if defined(__CUDACC__)
#define QR_TD __host__ __device__ __forceinline__
#else
#define QR_TD static inline
#endif

/* inclusive tile rectangle; x0 > x1: touches no tile */
struct qr_tile_rect_t { int32_t x0, y0, x1, y1; };

/*
 * Tile rectangle of a bounding box.  Primary rays are
 *     org + t * (hor * px + ver * py + dir),   t >= 0
 * (engine.cpp:3561-3584; hor / ver are the per-pixel steps, dir aims at pixel
 * (0, 0)), hor, ver and the view normal are mutually orthogonal.  A point P
 * in front of the eye is seen at
 *     q = (P - org) * (dir . n) / ((P - org) . n) - dir,  px = q . hor / |hor|^2
 */
QR_TD qr_tile_rect_t qr_tile_rect(const qr_blob_header &h, const qr_bound &b)
{
    qr_tile_rect_t all = { 0, 0, h.tls_row - 1, h.tls_col - 1 };
    qr_tile_rect_t none = { 1, 1, 0, 0 };
    if (b.n < 0) return none;           /* not in the camera list */
    if (b.n == 0 || b.n > 8) return all;

    /* view normal = hor x ver (not normalised: only ratios are used) */
    const float nx = h.hor[1] * h.ver[2] - h.hor[2] * h.ver[1];
    const float ny = h.hor[2] * h.ver[0] - h.hor[0] * h.ver[2];
    const float nz = h.hor[0] * h.ver[1] - h.hor[1] * h.ver[0];
    const float hh = h.hor[0] * h.hor[0] + h.hor[1] * h.hor[1] + h.hor[2] * h.hor[2];
    const float vv = h.ver[0] * h.ver[0] + h.ver[1] * h.ver[1] + h.ver[2] * h.ver[2];
    float zd = h.dir[0] * nx + h.dir[1] * ny + h.dir[2] * nz;      /* screen plane distance * |n| */
    float sgn = 1.0f;
    if (zd < 0.0f) { zd = -zd; sgn = -1.0f; }
    if (!(zd > 0.0f) || !(hh > 0.0f) || !(vv > 0.0f)) return all;

    /* distances along the view normal in world units: screen plane at "pov";
     * what lies in front of (or within the reference's RT_CLIP_THRESHOLD = 0.01
     * behind, object.h:39) the screen plane is projected, the rest is cut off
     * at a plane just behind it -- the reference cuts at the screen plane
     * itself (engine.cpp:2049-2087), so every projected edge here contains the
     * reference's */
    const float nlen = sqrtf(nx * nx + ny * ny + nz * nz);
    if (!(nlen > 0.0f)) return all;
    const float pov = zd / nlen;
    const float margin = 0.05f * pov > 0.011f ? 0.05f * pov : 0.011f;
    const float thr = pov - margin;
    if (!(thr > 0.25f * pov)) return all;

    float xmin = 3.0e38f, xmax = -3.0e38f, ymin = 3.0e38f, ymax = -3.0e38f;
    float z[8], d[8][3];
    bool  front[8], any = false;
    for (int k = 0; k < b.n; k++)
    {
        d[k][0] = b.v[k][0] - h.org[0]; d[k][1] = b.v[k][1] - h.org[1]; d[k][2] = b.v[k][2] - h.org[2];
        z[k] = (d[k][0] * nx + d[k][1] * ny + d[k][2] * nz) * sgn / nlen;
        if (!(z[k] == z[k])) return all;
        front[k] = z[k] >= thr;
    }
    /* the box cut by the plane z = thr is convex: its outline is spanned by
     * the vertices in front and the points where segments between a vertex in
     * front and one behind cross the plane (crossings of box diagonals lie
     * inside the outline, so all pairs may be tried without a table of edges) */
    for (int i = 0; i < b.n; i++)
    {
        for (int j = i; j < b.n; j++)
        {
            float px3[3], zz;
            if (i == j)
            {
                if (!front[i]) continue;
                px3[0] = d[i][0]; px3[1] = d[i][1]; px3[2] = d[i][2]; zz = z[i];
            }
            else
            {
                if (front[i] == front[j]) continue;
                const int a = front[i] ? i : j, c = front[i] ? j : i;      /* a in front, c behind */
                const float t = (thr - z[c]) / (z[a] - z[c]);
                px3[0] = d[c][0] + (d[a][0] - d[c][0]) * t;
                px3[1] = d[c][1] + (d[a][1] - d[c][1]) * t;
                px3[2] = d[c][2] + (d[a][2] - d[c][2]) * t;
                zz = thr;
            }
            const float s = pov / zz;
            const float qx = px3[0] * s - h.dir[0], qy = px3[1] * s - h.dir[1], qz = px3[2] * s - h.dir[2];
            const float px = (qx * h.hor[0] + qy * h.hor[1] + qz * h.hor[2]) / hh;
            const float py = (qx * h.ver[0] + qy * h.ver[1] + qz * h.ver[2]) / vv;
            if (!(px == px) || !(py == py)) return all;
            xmin = px < xmin ? px : xmin; xmax = px > xmax ? px : xmax;
            ymin = py < ymin ? py : ymin; ymax = py > ymax ? py : ymax;
            any = true;
        }
    }
    if (!any) return none;              /* wholly behind the screen plane: no primary ray starts there */
    /* pixels -> tiles with half a tile of margin on every side: the reference
     * widens by RT_TILE_THRESHOLD = 0.2 tile, the samples of a pixel sit within
     * 0.33 pixel (at most 0.125 tile) of its integer coordinates */
    const float tw = (float)h.tile_w, th = (float)h.tile_h;
    float fx0 = xmin / tw - 0.5f, fx1 = xmax / tw + 0.5f, fy0 = ymin / th - 0.5f, fy1 = ymax / th + 0.5f;
    const float big = 1.0e9f;
    fx0 = fx0 < -big ? -big : fx0; fy0 = fy0 < -big ? -big : fy0;
    fx1 = fx1 > big ? big : fx1;   fy1 = fy1 > big ? big : fy1;
    qr_tile_rect_t r;
    r.x0 = (int32_t)fx0 - (fx0 < 0.0f ? 1 : 0); r.y0 = (int32_t)fy0 - (fy0 < 0.0f ? 1 : 0);     /* floor */
    r.x1 = (int32_t)fx1 - (fx1 < 0.0f ? 1 : 0); r.y1 = (int32_t)fy1 - (fy1 < 0.0f ? 1 : 0);
    if (r.x0 < 0) r.x0 = 0;
    if (r.y0 < 0) r.y0 = 0;
    if (r.x1 > h.tls_row - 1) r.x1 = h.tls_row - 1;
    if (r.y1 > h.tls_col - 1) r.y1 = h.tls_col - 1;
    return r;
}

/* does the element own its matrix (qr_kscene.h: it is followed by a CLOSE)? */
QR_TD bool qr_kleaf_own(uint32_t w)
{
    const uint32_t kind = QR_K_KIND(w);
    return kind == QR_K_PLANE_G || ((kind == QR_K_QUADRIC || kind == QR_K_TWOPLANE) && (w & QR_KF_OWN) != 0);
}

/*
 * The list of tile (tx, ty) into "out" (room for the capacity the packer
 * computed): the leaves of the camera list whose rectangle holds the tile, in
 * the camera list's order; a run of leaves of one transform node is opened and
 * closed once (the trnode element of engine.cpp:3190-3208).  Returns the
 * number of elements written, END included.
 */
QR_TD uint32_t qr_tile_list_build(const qr_kleaf *leaf, uint32_t n_leaf, const qr_tile_rect_t *rect,
                                  int32_t tx, int32_t ty, qr_kelem *out)
{
    uint32_t n = 0;
    uint32_t open = QR_KLEAF_NO_NODE;
    for (uint32_t i = 0; i < n_leaf; i++)
    {
        const qr_tile_rect_t r = rect[i];
        if (tx < r.x0 || tx > r.x1 || ty < r.y0 || ty > r.y1) continue;
        const qr_kleaf lf = leaf[i];
        if (lf.open != open)
        {
            if (open != QR_KLEAF_NO_NODE) { out[n].w = QR_K_CLOSE; out[n].aux = 0; n++; }
            if (lf.open != QR_KLEAF_NO_NODE) { out[n].w = lf.open | QR_K_OPEN; out[n].aux = 0; n++; }
            open = lf.open;
        }
        out[n].w = lf.w; out[n].aux = lf.aux; n++;
        if (qr_kleaf_own(lf.w)) { out[n].w = QR_K_CLOSE; out[n].aux = 0; n++; }
    }
    if (open != QR_KLEAF_NO_NODE) { out[n].w = QR_K_CLOSE; out[n].aux = 0; n++; }
    out[n].w = QR_KEND; out[n].aux = 0; n++;
    return n;
}

#endif /* QR_TILING_CUH */
