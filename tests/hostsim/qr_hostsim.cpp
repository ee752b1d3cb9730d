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
    const size_t kbytes = pk.bytes();
    void *kimg = aligned_alloc(64, (kbytes + 63) & ~(size_t)63);
    if (kimg == NULL) return -3;
    pk.write(kimg);
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
