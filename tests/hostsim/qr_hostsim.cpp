/*
 * qr_hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the product's device core (quadray-engine_b200/csrc/qr_core.cuh)
 * as plain host C++ so that its logic can be unit-tested against the oracle
 * in a container without a GPU (tests/test_core_hostsim.py, -m "not gpu").
 * It is not a fallback: nothing in the product links or loads this library,
 * and libquadray_b200.so fails loudly without a CUDA device.
 */
#include <stdlib.h>
#include <stdint.h>
#define QR_COUNT_OPS
unsigned long long qr_ops[4] = {0, 0, 0, 0};
#include "qr_core.cuh"
#include "qr_tiling.cuh"
#include "qr_pt.cuh"

/* device-tiling emulation: tiles built / elements written since the last call */
static uint64_t qr_tiling_tiles = 0, qr_tiling_elems = 0;
extern "C" void qr_hostsim_tiling(uint64_t out[2])
{
    out[0] = qr_tiling_tiles; out[1] = qr_tiling_elems;
    qr_tiling_tiles = qr_tiling_elems = 0;
}

/*
 * Surfaces (blob indices) of the list the device-side tiling builds for tile
 * "tile" of an untiled blob, in list order; -1 when the blob does not qualify.
 */
extern "C" int qr_hostsim_tile_list(const void *blob, size_t bytes, int tile, int32_t *surf_out, int cap)
{
    const qr_blob_header *h = (const qr_blob_header *)blob;
    if (bytes < sizeof(*h) || h->magic != QR_BLOB_MAGIC || h->total_bytes > bytes) return -1;
    qr_kpacker pk;
    if (pk.plan(blob) != 0 || !pk.device_tiling() || tile < 0 || tile >= h->n_tiles) return -1;
    std::vector<uint8_t> img((pk.bytes() + 63) & ~(size_t)63);
    pk.write(img.data());
    const qr_blob_header *kh = (const qr_blob_header *)img.data();
    const qr_kleaf *leaf = (const qr_kleaf *)(img.data() + (uint32_t)kh->pad3[1]);
    const qr_bound *bnd = (const qr_bound *)(img.data() + kh->off_bounds);
    const uint32_t nl = (uint32_t)kh->pad3[2];
    std::vector<qr_tile_rect_t> rect(nl);
    for (uint32_t i = 0; i < nl; i++) rect[i] = qr_tile_rect(*kh, bnd[leaf[i].bound]);
    std::vector<qr_kelem> out((size_t)kh->pad0[1]);
    const uint32_t n = qr_tile_list_build(leaf, nl, rect.data(), tile % kh->tls_row, tile / kh->tls_row, out.data());
    int m = 0;
    for (uint32_t i = 0; i < n; i++)
    {
        const uint32_t kind = QR_K_KIND(out[i].w);
        if (out[i].w == QR_KEND || kind == QR_K_OPEN || kind == QR_K_CLOSE) continue;
        if (m < cap) surf_out[m] = (int32_t)(QR_K_SURF_OFF(out[i].w) >> 7);
        m++;
    }
    return m;
}

/* algorithmic IEEE op counts since the last call: add/sub, mul, div, sqrt */
extern "C" void qr_hostsim_ops(uint64_t out[4])
{
    for (int i = 0; i < 4; i++) { out[i] = qr_ops[i]; qr_ops[i] = 0; }
}

extern "C" int qr_hostsim_render(const void *blob, size_t bytes, uint32_t *frame,
                                 int stride, float *t_out, int y0, int y1,
                                 uint64_t *rays /* [4]: primary, shadow, reflect, refract */)
{
    const qr_blob_header *h = (const qr_blob_header *)blob;
    if (bytes < sizeof(*h) || h->magic != QR_BLOB_MAGIC) return -1;
    if (h->version != QR_BLOB_VERSION || h->total_bytes > bytes) return -2;

    /* the product path: blob -> packed kscene image (qr_kscene.h) -> core */
    qr_kpacker pk;
    if (pk.plan(blob) != 0) return -4;
    const size_t kbytes = pk.device_bytes();
    void *kimg = aligned_alloc(64, (kbytes + 63) & ~(size_t)63);
    if (kimg == NULL) return -3;
    pk.write(kimg);
    if (pk.device_tiling())
    {
        /* what qr_tile_rect_kernel / qr_tile_list_kernel do on the device */
        uint8_t *img = (uint8_t *)kimg;
        const qr_blob_header *kh = (const qr_blob_header *)img;
        const qr_kleaf *leaf = (const qr_kleaf *)(img + (uint32_t)kh->pad3[1]);
        const qr_bound *bnd = (const qr_bound *)(img + kh->off_bounds);
        qr_tile_rect_t *rect = (qr_tile_rect_t *)(img + (uint32_t)kh->pad3[3]);
        const uint32_t nl = (uint32_t)kh->pad3[2], cap = (uint32_t)kh->pad0[1];
        for (uint32_t i = 0; i < nl; i++) rect[i] = qr_tile_rect(*kh, bnd[leaf[i].bound]);
        for (uint32_t t = 0; t < (uint32_t)kh->n_tiles; t++)
        {
            const uint32_t first = ((uint32_t)kh->pad0[0] - kh->off_elem) / (uint32_t)sizeof(qr_kelem) + t * cap;
            const uint32_t n = qr_tile_list_build(leaf, nl, rect, (int32_t)(t % (uint32_t)kh->tls_row),
                                                  (int32_t)(t / (uint32_t)kh->tls_row),
                                                  (qr_kelem *)(img + kh->off_elem) + first);
            if (n > cap) { free(kimg); return -5; }
            ((int32_t *)(img + kh->off_tiles))[t] = (int32_t)first;
            qr_tiling_elems += n;
        }
        qr_tiling_tiles += (uint64_t)kh->n_tiles;
    }
    qr_view<false> v;
    qr_view_init(v, kimg);

    const int fsaa = h->fsaa, spp = 1 << fsaa;
    static const int lane_px[3][4] = { {0, 1, 2, 3}, {0, 0, 1, 1}, {0, 0, 0, 0} };
    qr_frame stack[QR_STACK_DEPTH + 1];
    qr_f4 quads[QR_SC_QUADS];
    memset(quads, 0, sizeof(quads));
    qr_scratch sc;
    sc.quads = quads;
    uint64_t primary = 0;

    if (y0 < 0) y0 = 0;
    if (y1 > h->y_res) y1 = h->y_res;

    /* groups of 4 lanes = 4 >> fsaa pixels, as in the reference packets */
    for (int y = y0; y < y1; y++)
    {
        for (int x = 0; x < h->x_res; x += 4 >> fsaa)
        {
            float c[3][4], tb[4];
            for (int l = 0; l < 4; l++)
            {
                const int px = x + lane_px[fsaa][l];
                float col[3];
                qr_trace_sample<false>(v, px, y, l, stack, sc, col[0], col[1], col[2]);
                tb[l] = quads[QR_SC_MISC].x;
                primary++;
                for (int k = 0; k < 3; k++) c[k][l] = qr_clamp1(col[k]);
            }
            int n = 4;
            for (int p = 0; p < fsaa; p++)
            {
                for (int k = 0; k < 3; k++)
                {
                    for (int l = 0; l < n; l++) c[k][l] = qr_mul(c[k][l], 0.5f);
                    for (int l = 0; l < n / 2; l++) c[k][l] = qr_add(c[k][2 * l], c[k][2 * l + 1]);
                }
                n >>= 1;
            }
            for (int l = 0; l < (4 >> fsaa); l++)
            {
                if (x + l < h->x_res && frame != NULL)
                {
                    frame[(size_t)y * stride + x + l] = qr_pack(*h, c[0][l], c[1][l], c[2][l]);
                }
            }
            if (t_out != NULL)
            {
                for (int l = 0; l < 4; l++)
                {
                    const int px = x + lane_px[fsaa][l];
                    const int sm = fsaa == 0 ? 0 : fsaa == 1 ? (l & 1) : l;
                    if (px < h->x_res) t_out[((size_t)y * h->x_res + px) * spp + sm] = tb[l];
                }
            }
        }
    }
    free(kimg);
    if (rays != NULL)
    {
        rays[0] = primary;
        for (int k = 1; k < 4; k++) rays[k] = qr_sc_ld1(sc, QR_SC_MISC, (uint32_t)k);
    }
    return 0;
}

/*
 * Path tracer (csrc/qr_pt.cuh) with a "warp" of one lane: one more frame into
 * the caller's seed / colour planes.  Lanes are grouped in fours for the
 * anti-aliasing reduce exactly as above.
 */
extern "C" int qr_hostsim_render_pt(const void *blob, size_t bytes, uint32_t *frame, int stride,
                                    int y0, int y1, uint32_t *pseed, float *ptr_r, float *ptr_g,
                                    float *ptr_b, float *pts_c)
{
    const qr_blob_header *h = (const qr_blob_header *)blob;
    if (bytes < sizeof(*h) || h->magic != QR_BLOB_MAGIC) return -1;
    if (h->version != QR_BLOB_VERSION || h->total_bytes > bytes) return -2;
    static qr_pt::R r;
    qr_pt::init(&r, blob);

    *pts_c = *pts_c + 1.0f;
    const float pts_o = 1.0f / *pts_c;
    const float pts_u = 1.0f - pts_o;

    const int fsaa = h->fsaa;
    static const int lane_px[3][4] = { {0, 1, 2, 3}, {0, 0, 1, 1}, {0, 0, 0, 0} };
    if (y0 < 0) y0 = 0;
    if (y1 > h->y_res) y1 = h->y_res;
    for (int y = y0; y < y1; y++)
    {
        for (int x = 0; x < h->x_res; x += 4 >> fsaa)
        {
            float c[3][4];
            const size_t slot0 = ((size_t)y * (size_t)h->x_row + (size_t)x) << fsaa;
            for (int l = 0; l < 4; l++)
            {
                const int px = x + lane_px[fsaa][l];
                float col[3];
                qr_pt::trace_lane(&r, y, px, l, px, pseed, ptr_r, ptr_g, ptr_b, slot0 + (size_t)l, pts_o, pts_u, col);
                for (int k = 0; k < 3; k++) c[k][l] = qr_clamp1(col[k]);
            }
            int n = 4;
            for (int p = 0; p < fsaa; p++)
            {
                for (int k = 0; k < 3; k++)
                {
                    for (int l = 0; l < n; l++) c[k][l] = qr_mul(c[k][l], 0.5f);
                    for (int l = 0; l < n / 2; l++) c[k][l] = qr_add(c[k][2 * l], c[k][2 * l + 1]);
                }
                n >>= 1;
            }
            for (int l = 0; l < (4 >> fsaa); l++)
            {
                if (x + l < h->x_res && frame != NULL)
                {
                    frame[(size_t)y * stride + x + l] = qr_pack(*h, c[0][l], c[1][l], c[2][l]);
                }
            }
        }
    }
    return 0;
}
