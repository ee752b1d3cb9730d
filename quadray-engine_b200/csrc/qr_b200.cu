/*
 * qr_b200.cu -- sm_100a kernels and the C ABI of libquadray_b200.so.
 *
 * Kernel qr_render_kernel: a persistent grid (resident CTAs x SM count) whose
 * WARPS pull work from a device-side tile queue.  A work item is one band of
 * one screen tile (tile = tile_w x tile_h pixels, core/engine/engine.h:38-39):
 * the warp walks the band in packets of 32 samples -- a block of 4 x (8 >> fsaa)
 * pixels times 1 << fsaa samples, in the lane order of the reference's packets
 * (core/engine/engine.cpp:3465-3550) -- so all 32 rays of a warp share the
 * tile's surface list and start next to each other on the screen in BOTH
 * directions (a 4 x 2 pixel block at 4xAA diverges less than 8 x 1).  This
 * replaces the scanline interleave across worker threads
 * (core/tracer/tracer.cpp:1142-1151, 5383-5394).
 *
 * Scene staging: header + surfaces + materials + lights (the blob prefix up to
 * the list elements) are copied into shared memory once per CTA with one TMA
 * bulk copy (cp.async.bulk + mbarrier complete_tx); list elements, tile heads
 * and texels are read through L1/L2.
 *
 * Epilogue (XX_end, tracer.cpp:5161-5343): clamp, AA halving + pairwise adds
 * by warp shuffles, gamma, pack; four packed pixels per 128-bit store.
 */

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <new>
#include <sched.h>
#include <thread>
#include <mutex>
#include <condition_variable>

#include "quadray_b200.h"
#include "qr_core.cuh"
#include "qr_tiling.cuh"
#include "qr_pt.cuh"

/*
 * Launch shapes (threads per CTA, resident CTAs per SM the register budget is
 * sized for).  One CTA stages one copy of the scene prefix, so a few large
 * CTAs leave more of the 256 KB L1/shared array to the L1 cache than many
 * small ones.  QR_B200_SHAPE=<index> overrides the default (tuning only).
 */
struct qr_shape { int threads, ctas; };
static const qr_shape g_shapes[] = { {256, 2}, {512, 1}, {640, 1}, {768, 1}, {384, 1}, {128, 4},
                                     {896, 1}, {1024, 1} };
#define QR_N_SHAPES     8
#define QR_DEFAULT_SHAPE 2
#define QR_BIG_SHAPE     6          /* 896 threads at 72 registers: 28 warps per SM (1080p 4xAA demo scene:
                                       1.46 ms; 1.53 at 768 x 80, 1.48 at 1024 x 64, 1.6 at 640 x 96) */
#define QR_BIG_FRAME_ITEMS 60000    /* work items (32 samples each) per GPU from which QR_BIG_SHAPE pays */

/* ------------------------------------------------------------------ PTX --- */

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src,
                                             uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
                 " [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    uint32_t ok = 0;
    while (!ok)
    {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    }
}

/* --------------------------------------------------------------- kernel --- */

struct qr_launch
{
    const uint8_t      *blob;       /* scene blob in global memory */
    uint32_t           *frame;      /* framebuffer (this GPU) */
    int                 stride;     /* pixels per framebuffer row */
    int                 ty0;        /* first tile row of this launch */
    int                 ty_step;    /* distance between its tile rows (1 = contiguous band) */
    int                 n_trows;    /* number of tile rows: ty0, ty0 + ty_step, ... */
    uint32_t            stage_bytes;/* blob prefix staged in smem, 0 = none */
    unsigned int       *queue;      /* work-item counter: never reset, this launch's items start at ... */
    unsigned int        queue_base; /* ... this value (every warp draws exactly one item too many, so
                                       the host knows where the counter stands after the launch) */
    unsigned long long *rays;       /* [4] ray counters */
    float              *t_out;      /* dump mode, or NULL */
    unsigned int       *notify;     /* or NULL: +1 (system scope) when all pixels of this launch are visible */
    unsigned int       *done;       /* warps of this launch that have finished (notify != NULL) */
    unsigned int        n_warps;    /* warps of this launch */
};

extern __shared__ __align__(128) uint8_t qr_smem[];

#if defined(QR_ITEMLOG)
/* tuning build only (make itemlog): start, duration and SM of every work item */
struct qr_itemlog_t { unsigned long long t0; unsigned int dt, sm; };
static __device__ qr_itemlog_t *qr_itemlog_ptr;
__device__ __forceinline__ unsigned long long qr_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif

/*
 * STAGED = true: the kscene prefix (header, surfaces, shading records,
 * materials, lights) is copied to shared memory by TMA and every access to it
 * is derived from the qr_smem symbol, so the compiler emits LDS.128.
 * STAGED = false (prefix larger than shared memory): everything through L1/L2.
 */
/* staged view: hot sections addressed in the shared window, the rest in global memory */
__device__ __forceinline__ void qr_view_setup(qr_view<true> &v, const uint8_t *img)
{
    const qr_blob_header *h = (const qr_blob_header *)qr_smem;
    const uint32_t base = smem_u32(qr_smem);
    v.h      = h;
    v.surf   = base + h->off_surf;
    v.shade  = base + (uint32_t)h->pad3[0];
    v.mat    = base + h->off_mat;
    v.lgt    = base + h->off_lgt;
    v.elems  = (const qr_kelem *)(img + h->off_elem);
    v.tiles  = (const int32_t *)(img + h->off_tiles);
    v.texels = (const uint32_t *)(img + h->off_texels);
}

__device__ __forceinline__ void qr_view_setup(qr_view<false> &v, const uint8_t *img)
{
    qr_view_init(v, img);
}

template <bool STAGED, int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS)
qr_render_kernel(const qr_launch p)
{
    __shared__ __align__(8) uint64_t bar;

    const int lane = threadIdx.x & 31;

    qr_view<STAGED> v;
    if (STAGED)
    {
        /* one elected thread issues the TMA bulk copies, all wait on the mbarrier */
        if (threadIdx.x == 0)
        {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            mbar_expect_tx(&bar, p.stage_bytes);
            uint32_t off = 0;
            while (off < p.stage_bytes)
            {
                uint32_t n = p.stage_bytes - off;
                if (n > 32768u) n = 32768u;
                tma_bulk_g2s(qr_smem + off, p.blob + off, n, &bar);
                off += n;
            }
        }
        mbar_wait(&bar, 0);
    }
    qr_view_setup(v, p.blob);

    const qr_blob_header &h = *v.h;
    const int fsaa  = h.fsaa;
    const int bh    = 8 >> fsaa;                    /* a packet is a block of 4 x bh pixels */
    const int x_res = h.x_res, y_res = h.y_res;
    const int tiles_x = h.tls_row;
    const int tile_w = h.tile_w, tile_h = h.tile_h;
    const int pk_per_row = (tile_w + 3) / 4;        /* blocks per band of a tile */
    const int bands = (tile_h + bh - 1) / bh;       /* bands of bh rows per tile */
    const unsigned int per_tile = (unsigned int)bands * (unsigned int)pk_per_row;
    const unsigned int n_items = (unsigned int)p.n_trows * (unsigned int)tiles_x * per_tile;

    /* lane -> (pixel within packet, sample, AA pattern slot), engine.cpp:3465-3550 */
    const int lpx   = lane >> fsaa;
    const int smp   = lane & ((1 << fsaa) - 1);
    const int lane4 = fsaa == 0 ? (lane & 3)
                    : fsaa == 1 ? (((lpx & 1) << 1) | smp)
                    : smp;

    qr_frame stack[QR_STACK_DEPTH + 1];
    /* per-thread scratch quads behind the staged scene (qr_core.cuh) */
    qr_scratch sc;
    sc.addr = smem_u32(qr_smem) + p.stage_bytes + threadIdx.x * 16u;
    sc.stride = THREADS * 16u;
    /* opaque to the optimiser (a volatile round trip through its own scratch):
     * one register, instead of S2R + LEA from the thread index at every use */
    asm volatile("st.volatile.shared.b32 [%0], %0;\n\t"
                 "ld.volatile.shared.b32 %0, [%0];" : "+r"(sc.addr) :: "memory");
    qr_sc_st(sc, QR_SC_MISC, 0.0f, 0.0f, 0.0f, 0.0f);

    /* one work item = one packet (a 4 x bh pixel block of a tile); the queue
     * is read one item ahead, so the atomic's latency hides behind a trace */
    unsigned int item = 0;
    if (lane == 0) item = atomicAdd(p.queue, 1u) - p.queue_base;
    item = __shfl_sync(0xFFFFFFFFu, item, 0);

    while (item < n_items)
    {
#if defined(QR_ITEMLOG)
        const unsigned long long il_t0 = qr_globaltimer();
        const unsigned int il_item = item;
#endif
        const unsigned int tile = item / per_tile;
        const unsigned int sub  = item % per_tile;
        const int brow = (int)(sub / (unsigned int)pk_per_row);
        const int pk   = (int)(sub % (unsigned int)pk_per_row);
        const int ty   = p.ty0 + (int)(tile / (unsigned int)tiles_x) * p.ty_step;
        const int tx   = (int)(tile % (unsigned int)tiles_x);
        const int y0   = ty * tile_h + brow * bh;   /* first row of the block */
        const int x0   = tx * tile_w + pk * 4;      /* first pixel column of the block */
        const int y    = y0 + (lpx >> 2);
        /* rows of this band that exist (inside the tile and the frame) */
        int rows = tile_h - brow * bh;
        if (rows > bh) rows = bh;
        if (rows > y_res - y0) rows = y_res - y0;

        if (y0 < y_res && x0 < x_res)
        {
            const int px = x0 + (lpx & 3);

            float col[3] = {0.0f, 0.0f, 0.0f};
            const bool live = px < x_res && (lpx >> 2) < rows;
            if (live)
            {
                qr_trace_sample<STAGED>(v, px, y, lane4, stack, sc, col[0], col[1], col[2]);
                if (p.t_out != NULL)
                {
                    p.t_out[(((size_t)y * x_res + px) << fsaa) + smp] =
                        qr_u2f(qr_sc_ld1(sc, QR_SC_MISC, 0));
                }
            }
            __syncwarp();

            /* XX_end: clamp, then AA passes of halve + add adjacent pairs */
            float r = qr_clamp1(col[0]), g = qr_clamp1(col[1]), b = qr_clamp1(col[2]);
            for (int pass = 0; pass < fsaa; pass++)
            {
                const int d = 1 << pass;
                r = qr_mul(r, 0.5f); g = qr_mul(g, 0.5f); b = qr_mul(b, 0.5f);
                const float r2 = __shfl_xor_sync(0xFFFFFFFFu, r, d);
                const float g2 = __shfl_xor_sync(0xFFFFFFFFu, g, d);
                const float b2 = __shfl_xor_sync(0xFFFFFFFFu, b, d);
                /* the lower lane of the pair is the left operand of the add */
                if ((lane & d) == 0)
                {
                    r = qr_add(r, r2); g = qr_add(g, g2); b = qr_add(b, b2);
                }
                else
                {
                    r = qr_add(r2, r); g = qr_add(g2, g); b = qr_add(b2, b);
                }
            }
            const uint32_t pix = qr_pack(h, r, g, b);

            /* pixel q of the block (row q >> 2, column q & 3) sits in lane
             * q << fsaa; lanes 0..bh-1 collect the four pixels of one block
             * row each and issue one 128-bit store */
            uint32_t q4[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                q4[j] = __shfl_sync(0xFFFFFFFFu, pix, (((lane << 2) + j) << fsaa) & 31);
            }
            if (lane < rows)
            {
                const int xq = x0;
                uint32_t *dst = p.frame + (size_t)(y0 + lane) * p.stride + xq;
                QR_CHECK(y0 + lane < y_res && xq < x_res && p.stride >= x_res, 4);
                /* one 128-bit store when the ADDRESS allows it (a caller's frame
                 * or stride need not be 16-byte aligned) */
                if (xq + 3 < x_res && (((size_t)dst & 15) == 0))
                {
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(q4[0], q4[1], q4[2], q4[3]);
                }
                else
                {
                    for (int j = 0; j < 4; j++)
                    {
                        if (xq + j < x_res) dst[j] = q4[j];
                    }
                }
            }
        }
#if defined(QR_ITEMLOG)
        if (lane == 0 && qr_itemlog_ptr != NULL)
        {
            unsigned int sm;
            asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
            qr_itemlog_t rec = { il_t0, (unsigned int)(qr_globaltimer() - il_t0), sm };
            qr_itemlog_ptr[il_item] = rec;
        }
#endif
        /* drawn when it is needed, not one item ahead: an item reserved early
         * waits behind the one in progress while other warps run dry, which
         * doubles the tail of a launch; the atomic's latency is hidden by the
         * other warps of the SM */
        if (lane == 0) item = atomicAdd(p.queue, 1u) - p.queue_base;
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
    }

    /* ray counters: warp-reduce, one atomic per warp and kind */
    unsigned int c1 = qr_sc_ld1(sc, QR_SC_MISC, 1), c2 = qr_sc_ld1(sc, QR_SC_MISC, 2),
                 c3 = qr_sc_ld1(sc, QR_SC_MISC, 3);
    for (int d = 16; d > 0; d >>= 1)
    {
        c1 += __shfl_xor_sync(0xFFFFFFFFu, c1, d);
        c2 += __shfl_xor_sync(0xFFFFFFFFu, c2, d);
        c3 += __shfl_xor_sync(0xFFFFFFFFu, c3, d);
    }
    if (lane == 0)
    {
        atomicAdd(&p.rays[1], (unsigned long long)c1);
        atomicAdd(&p.rays[2], (unsigned long long)c2);
        atomicAdd(&p.rays[3], (unsigned long long)c3);
    }

    /* completion signal: the last warp of the launch to get here tells the
     * owner of the frame (possibly another GPU: the word sits behind rank 0's
     * framebuffer, reached over NVLink like the pixels) that every pixel of
     * this launch is visible */
    if (p.notify != NULL)
    {
        __threadfence_system();
        __syncwarp();
        if (lane == 0)
        {
            if (atomicAdd(p.done, 1u) == p.n_warps - 1u)
            {
                *p.done = 0u;
                __threadfence();
                atomicAdd_system(p.notify, 1u);
            }
        }
    }
}

/*
 * Path tracer (qr_pt.cuh): one warp = one packet of the reference's 512x2v2
 * target, 32 >> fsaa pixels of one row; the warps of a persistent grid draw
 * packets from a counter.  Every lane of a packet is traced, also those right
 * of x_res (the reference traces them: they vote in the packet-wide tests).
 */
struct qr_pt_launch
{
    const uint8_t *blob;        /* the scene blob as flattened (not the kscene image) */
    uint32_t      *frame;
    int            stride;
    int            ty0, ty_step, n_trows;
    uint32_t      *pseed;       /* seed and colour planes, slot ((y * x_row + x) << fsaa) + lane */
    float         *ptr_r, *ptr_g, *ptr_b;
    float          pts_o, pts_u;/* 1 / frames so far, 1 - that */
    unsigned int  *queue;       /* zero at launch */
};

#define QR_PT_THREADS 128
#define QR_PT_STACK   (16 * 1024)   /* bytes per thread: 13 contexts deep walk -> material -> walk recursion */

#ifndef QR_PT_MINBLOCKS
#define QR_PT_MINBLOCKS 8     /* 64 registers: 32 warps per SM (the measured configuration) */
#endif
__global__ void __launch_bounds__(QR_PT_THREADS, QR_PT_MINBLOCKS)
qr_pt_kernel(const qr_pt_launch p)
{
    qr_pt::R r;
    qr_pt::init(&r, p.blob);
    const qr_blob_header &h = *r.h;
    const int lane = threadIdx.x & 31;
    const int fsaa = h.fsaa;
    const int ppp = 32 >> fsaa;                             /* pixels per packet */
    const unsigned int per_row = (unsigned int)((h.x_res + ppp - 1) / ppp);
    const unsigned int n_items = (unsigned int)p.n_trows * (unsigned int)h.tile_h * per_row;
    const int lpx = (lane >> 2) * (4 >> fsaa) + (fsaa == 0 ? (lane & 3) : fsaa == 1 ? ((lane & 3) >> 1) : 0);

    for (;;)
    {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(p.queue, 1u);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        if (item >= n_items) break;
        const unsigned int rw = item / per_row;
        const int x = (int)(item % per_row) * ppp;
        const int y = (p.ty0 + (int)(rw / (unsigned int)h.tile_h) * p.ty_step) * h.tile_h
                    + (int)(rw % (unsigned int)h.tile_h);
        if (y >= h.y_res) continue;

        float col[3];
        const size_t slot = (((size_t)y * (size_t)h.x_row + (size_t)x) << fsaa) + (size_t)lane;
        qr_pt::trace_lane(&r, y, x + lpx, lane, x, p.pseed, p.ptr_r, p.ptr_g, p.ptr_b, slot,
                          p.pts_o, p.pts_u, col);
        __syncwarp();

        /* XX_end 5221-5343 as in qr_render_kernel */
        float cr = qr_clamp1(col[0]), cg = qr_clamp1(col[1]), cb = qr_clamp1(col[2]);
        for (int pass = 0; pass < fsaa; pass++)
        {
            const int d = 1 << pass;
            cr = qr_mul(cr, 0.5f); cg = qr_mul(cg, 0.5f); cb = qr_mul(cb, 0.5f);
            const float r2 = __shfl_xor_sync(0xFFFFFFFFu, cr, d);
            const float g2 = __shfl_xor_sync(0xFFFFFFFFu, cg, d);
            const float b2 = __shfl_xor_sync(0xFFFFFFFFu, cb, d);
            if ((lane & d) == 0)
            {
                cr = qr_add(cr, r2); cg = qr_add(cg, g2); cb = qr_add(cb, b2);
            }
            else
            {
                cr = qr_add(r2, cr); cg = qr_add(g2, cg); cb = qr_add(b2, cb);
            }
        }
        const int pxo = x + (lane >> fsaa);
        if ((lane & ((1 << fsaa) - 1)) == 0 && pxo < h.x_res)
        {
            p.frame[(size_t)y * p.stride + pxo] = qr_pack(h, cr, cg, cb);
        }
    }
}

/*
 * Device-side tiling (qr_tiling.cuh), two launches per uploaded scene: the
 * tile rectangle of every leaf of the camera list, then one list per tile.
 */
__global__ void qr_tile_rect_kernel(uint8_t *img)
{
    const qr_blob_header *h = (const qr_blob_header *)img;
    const uint32_t n = (uint32_t)h->pad3[2];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const qr_kleaf *leaf = (const qr_kleaf *)(img + (uint32_t)h->pad3[1]);
    const qr_bound *bnd = (const qr_bound *)(img + h->off_bounds);
    qr_tile_rect_t *rect = (qr_tile_rect_t *)(img + (uint32_t)h->pad3[3]);
    rect[i] = qr_tile_rect(*h, bnd[leaf[i].bound]);
}

__global__ void qr_tile_list_kernel(uint8_t *img)
{
    const qr_blob_header *h = (const qr_blob_header *)img;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint32_t)h->n_tiles) return;
    const qr_kleaf *leaf = (const qr_kleaf *)(img + (uint32_t)h->pad3[1]);
    const qr_tile_rect_t *rect = (const qr_tile_rect_t *)(img + (uint32_t)h->pad3[3]);
    const uint32_t cap = (uint32_t)h->pad0[1];
    const uint32_t first = ((uint32_t)h->pad0[0] - h->off_elem) / (uint32_t)sizeof(qr_kelem) + t * cap;
    qr_kelem *out = (qr_kelem *)(img + h->off_elem) + first;
    qr_tile_list_build(leaf, (uint32_t)h->pad3[2], rect, (int32_t)(t % (uint32_t)h->tls_row),
                       (int32_t)(t / (uint32_t)h->tls_row), out);
    ((int32_t *)(img + h->off_tiles))[t] = (int32_t)first;
}

/* a rank without rows of its own still owes its completion signal */
__global__ void qr_add_kernel(unsigned int *flag)
{
    atomicAdd_system(flag, 1u);
}

/* rank 0's stream waits here until *flag has reached "target" (wrap-around safe) */
__global__ void qr_wait_kernel(unsigned int *flag, unsigned int target)
{
    volatile unsigned int *f = flag;
    while ((int)(*f - target) < 0)
    {
        __nanosleep(64);
    }
    __threadfence_system();
}

/*
 * FP32 pipe ceiling probe: 8 independent chains per thread of x = x * a + b
 * issued as separate FMUL and FADD (never FFMA), all in registers.
 */
__global__ void __launch_bounds__(256)
qr_fp32_peak_kernel(float *out, float a, float b, int iters)
{
    float x0 = threadIdx.x, x1 = x0 + 1.0f, x2 = x0 + 2.0f, x3 = x0 + 3.0f;
    float x4 = x0 + 4.0f, x5 = x0 + 5.0f, x6 = x0 + 6.0f, x7 = x0 + 7.0f;
#pragma unroll 4
    for (int i = 0; i < iters; i++)
    {
        x0 = __fadd_rn(__fmul_rn(x0, a), b); x1 = __fadd_rn(__fmul_rn(x1, a), b);
        x2 = __fadd_rn(__fmul_rn(x2, a), b); x3 = __fadd_rn(__fmul_rn(x3, a), b);
        x4 = __fadd_rn(__fmul_rn(x4, a), b); x5 = __fadd_rn(__fmul_rn(x5, a), b);
        x6 = __fadd_rn(__fmul_rn(x6, a), b); x7 = __fadd_rn(__fmul_rn(x7, a), b);
    }
    const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;       /* keep the chains alive */
}

/* ------------------------------------------------------------ host side --- */

typedef void (*qr_kernel_fn)(const qr_launch);

static qr_kernel_fn qr_kernel_of(bool staged, int shape)
{
    switch (shape)
    {
        case 1:  return staged ? qr_render_kernel<true, 512, 1> : qr_render_kernel<false, 512, 1>;
        case 2:  return staged ? qr_render_kernel<true, 640, 1> : qr_render_kernel<false, 640, 1>;
        case 3:  return staged ? qr_render_kernel<true, 768, 1> : qr_render_kernel<false, 768, 1>;
        case 4:  return staged ? qr_render_kernel<true, 384, 1> : qr_render_kernel<false, 384, 1>;
        case 5:  return staged ? qr_render_kernel<true, 128, 4> : qr_render_kernel<false, 128, 4>;
        case 6:  return staged ? qr_render_kernel<true, 896, 1> : qr_render_kernel<false, 896, 1>;
        case 7:  return staged ? qr_render_kernel<true, 1024, 1> : qr_render_kernel<false, 1024, 1>;
        default: return staged ? qr_render_kernel<true, 256, 2> : qr_render_kernel<false, 256, 2>;
    }
}

#define QR_MAX_DEV 16
#define QR_NOTIFY_SLOTS 16  /* completion counters behind the library's framebuffer */
#define QR_COPY_THREADS 16  /* upper bound; QR_B200_COPY_THREADS (default 4): one thread moves ~11 GB/s,
                               a 1080p frame in 0.75 ms, four in 0.25 */
#define QR_MAX_CHUNKS 8     /* qr_render to a host frame: render / D2H pipeline depth */

struct qr_dev
{
    int             id;
    cudaStream_t    stream;
    cudaStream_t    copy;                           /* D2H of finished chunks (GPU 0) */
    cudaStream_t    up;                             /* H2D of the next scene in pipelined mode */
    cudaEvent_t     ev0, ev1, done;
    cudaEvent_t     chunk_ev[QR_MAX_CHUNKS], copy_ev[QR_MAX_CHUNKS];
    /* scene image: two slots, so that in pipelined mode the scene of frame
     * N + 1 is packed and copied while the kernels of frame N still read theirs */
    uint8_t        *blob_d[2];  size_t blob_cap[2];
    uint8_t        *blob_h[2];  size_t blob_hcap[2];    /* pinned staging (dev 0 only) */
    uint32_t       *frame_pd[2]; size_t frame_pdcap[2]; /* device frames of pipelined mode (dev 0 only) */
    uint32_t       *frame_p[2]; size_t frame_pcap[2];   /* their pinned staging, for callers' pageable frames */
    cudaEvent_t     pipe_ev[2];                         /* frame of slot k rendered (dev 0 only) */
    cudaEvent_t     fetch_ev[2];                        /* ... and delivered to the host */
    uint32_t       *frame_d;    size_t frame_cap;   /* bytes */
    uint32_t       *frame_h;    size_t frame_hcap;  /* pinned (dev 0 only) */
    float          *t_d;        size_t t_cap;
    unsigned int   *queue_d;                        /* [0] work-item counter, [1] finished warps */
    unsigned int    queue_base;                     /* where the counter stands (see qr_launch) */
    unsigned long long *rays_d;
    /* path tracer: the blob as flattened, the seed and colour planes of the frame */
    uint8_t        *raw_d;      size_t raw_cap;
    uint32_t       *pt_seed;    float *pt_col[3];   size_t pt_slots;
    bool            pt_stack;   /* the stack limit of the recursive packet tracer is set */
    int             sm_count;
    int             ctas_per_sm;
    int             smem_optin;
    bool            peer_ok;    /* can store straight into GPU 0's framebuffer */
    bool            timed;
};

/* rows [ya, yb) of a frame in pinned staging, to be copied to the caller once "ev" has passed */
struct qr_copy_job
{
    uint32_t       *dst;
    const uint32_t *src;
    int             st, dstride, x_res, ya, yb;
    cudaEvent_t     ev;
};

struct qr_ctx
{
    int             ndev;
    qr_dev          dev[QR_MAX_DEV];
    qr_blob_header  hdr;
    bool            have_scene;
    bool            pt_scene;       /* the scene asks for the path tracer (QR_BLOB_PT) */
    float           pt_count;       /* frames accumulated since qr_pt_reset (inf_PTS_C) */
    float           pt_o, pt_u;     /* 1 / pt_count and 1 - that, for the frame in progress */
    uint32_t        stage_bytes;
    uint64_t        launches;
    uint64_t        rays[4];
    int             shape;          /* index into g_shapes */
    int             shape_fixed;    /* QR_B200_SHAPE given: no automatic choice per frame size */
    int             chunks;         /* qr_render(host frame) pipeline depth on one GPU, 0 = automatic */
    int             pin_frames;     /* QR_B200_PIN_FRAME=1: page-lock the caller's framebuffer on first use */
    int             zerocopy;       /* store pixels straight into a page-locked host frame (QR_B200_ZEROCOPY=0: off) */
    void           *pinned[4];      /* framebuffers registered that way ... */
    size_t          pinned_bytes[4];/* ... and how much of them */
    unsigned        pinned_age[4];  /* last use (the oldest entry is evicted) */
    unsigned        pin_clock;
    bool            device_tiling;  /* the current scene's tile lists were built on the device */
    int             pipelined;      /* qr_pipeline(ctx, 1): scenes alternate between two slots */
    int             slot;           /* slot of the current scene */
    bool            pending[2];     /* a frame begun in this slot has not been collected */
    int             fetching[2];    /* 0 not yet, 1 DMA into the caller's frame, 2 DMA into pinned staging */
    uint32_t       *fetch_dst[2];   /* where qr_render_fetch was told to put it */
    int             fetch_stride[2];
    bool            fetch_posted[2];
    std::thread     helper[QR_COPY_THREADS];    /* staging -> caller's pageable frame copies */
    qr_copy_job     job[2];
    int             nhelper;
    std::mutex      hmtx;
    std::condition_variable hcv;
    int             hjob[2];        /* per slot: 0 none, 1 queued, 2 done, 3 failed */
    int             hleft[2];       /* parts of the job still being copied */
    bool            htaken[2][QR_COPY_THREADS];
    bool            hfail[2];
    bool            hquit;
    qr_blob_header  pend_hdr[2];    /* its geometry */
    cudaFuncAttributes fattr;
    qr_kpacker      packer;
    char            err[512];
};

static thread_local char g_init_err[512] = "";

/* CPU affinity of the process at load time (helper threads inherit it) */
static cpu_set_t g_load_affinity;
static const bool g_load_affinity_ok = sched_getaffinity(0, sizeof(g_load_affinity), &g_load_affinity) == 0;

static int qr_fail(qr_ctx *ctx, int code, const char *fmt, ...)
{
    char *dst = ctx != NULL ? ctx->err : g_init_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define QR_CUDA(ctx, call)                                                   \
    do                                                                       \
    {                                                                        \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess)                                               \
        {                                                                    \
            return qr_fail(ctx, QR_E_CUDA, "%s failed: %s (%s:%d)", #call,   \
                           cudaGetErrorString(e_), __FILE__, __LINE__);      \
        }                                                                    \
    }                                                                        \
    while (0)

extern "C" const char *qr_last_error(const qr_ctx *ctx)
{
    return ctx != NULL ? ctx->err : g_init_err;
}

extern "C" int qr_init(const int *devices, int ndev, qr_ctx **out)
{
    if (out == NULL || ndev < 0 || ndev > QR_MAX_DEV)
    {
        return qr_fail(NULL, QR_E_ARG, "qr_init: bad arguments");
    }
    *out = NULL;

    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
    {
        return qr_fail(NULL, QR_E_NODEV, "qr_init: no CUDA device (%s); this "
                       "backend has no CPU fallback", cudaGetErrorString(e));
    }

    qr_ctx *ctx = new (std::nothrow) qr_ctx();
    if (ctx == NULL)
    {
        return qr_fail(NULL, QR_E_ARG, "qr_init: out of memory");
    }

    int cur = 0;
    cudaGetDevice(&cur);
    ctx->ndev = ndev == 0 ? 1 : ndev;

    for (int i = 0; i < ctx->ndev; i++)
    {
        qr_dev &d = ctx->dev[i];
        d.id = devices != NULL ? devices[i] : (ndev == 0 ? cur : i);
        if (d.id < 0 || d.id >= count)
        {
            int rc = qr_fail(NULL, QR_E_NODEV, "qr_init: device %d not present (%d visible)", d.id, count);
            delete ctx;
            return rc;
        }
    }

    for (int i = 0; i < ctx->ndev; i++)
    {
        qr_dev &d = ctx->dev[i];
        cudaDeviceProp prop;
        if ((e = cudaSetDevice(d.id)) != cudaSuccess
        ||  (e = cudaGetDeviceProperties(&prop, d.id)) != cudaSuccess
        ||  (e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking)) != cudaSuccess
        ||  (e = cudaStreamCreateWithFlags(&d.copy, cudaStreamNonBlocking)) != cudaSuccess
        ||  (e = cudaStreamCreateWithFlags(&d.up, cudaStreamNonBlocking)) != cudaSuccess
        ||  (e = cudaEventCreate(&d.ev0)) != cudaSuccess
        ||  (e = cudaEventCreate(&d.ev1)) != cudaSuccess
        ||  (e = cudaEventCreateWithFlags(&d.done, cudaEventDisableTiming)) != cudaSuccess
        ||  (e = cudaMalloc(&d.queue_d, 2 * sizeof(unsigned int))) != cudaSuccess
        ||  (e = cudaMemset(d.queue_d, 0, 2 * sizeof(unsigned int))) != cudaSuccess
        ||  (e = cudaMalloc(&d.rays_d, 4 * sizeof(unsigned long long))) != cudaSuccess
        ||  (e = cudaMemset(d.rays_d, 0, 4 * sizeof(unsigned long long))) != cudaSuccess)
        {
            int rc = qr_fail(NULL, QR_E_CUDA, "qr_init: device %d setup failed: %s", d.id, cudaGetErrorString(e));
            qr_shutdown(ctx);
            return rc;
        }
        if (prop.major < 10)
        {
            int rc = qr_fail(NULL, QR_E_NODEV, "qr_init: device %d is sm_%d%d; this library is built for sm_100a only",
                             d.id, prop.major, prop.minor);
            qr_shutdown(ctx);
            return rc;
        }
        for (int k = 0; k < QR_MAX_CHUNKS; k++)
        {
            if ((e = cudaEventCreateWithFlags(&d.chunk_ev[k], cudaEventDisableTiming)) != cudaSuccess
            ||  (e = cudaEventCreateWithFlags(&d.copy_ev[k], cudaEventDisableTiming)) != cudaSuccess)
            {
                int rc = qr_fail(NULL, QR_E_CUDA, "qr_init: device %d event setup failed: %s", d.id, cudaGetErrorString(e));
                qr_shutdown(ctx);
                return rc;
            }
        }
        d.sm_count = prop.multiProcessorCount;
        d.smem_optin = (int)prop.sharedMemPerBlockOptin;
        d.ctas_per_sm = 1;
    }

    /* peer access for the framebuffer gather to GPU 0 */
    for (int i = 1; i < ctx->ndev; i++)
    {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, ctx->dev[i].id, ctx->dev[0].id);
        if (can)
        {
            cudaSetDevice(ctx->dev[i].id);
            e = cudaDeviceEnablePeerAccess(ctx->dev[0].id, 0);
            ctx->dev[i].peer_ok = (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled);
            cudaGetLastError();
        }
    }

    {
        const char *env = getenv("QR_B200_PIN_FRAME");
        ctx->pin_frames = env != NULL && env[0] == '1';
        env = getenv("QR_B200_ZEROCOPY");
        ctx->zerocopy = !(env != NULL && env[0] == '0');     /* on unless QR_B200_ZEROCOPY=0 */
    }
    ctx->shape = QR_DEFAULT_SHAPE;
    ctx->chunks = 0;            /* 0 = automatic */
    {
        const char *env = getenv("QR_B200_CHUNKS");
        if (env != NULL && env[0] >= '1' && env[0] <= '0' + QR_MAX_CHUNKS && env[1] == 0)
        {
            ctx->chunks = env[0] - '0';
        }
    }
    {
        const char *env = getenv("QR_B200_SHAPE");
        if (env != NULL && env[0] >= '0' && env[0] < '0' + QR_N_SHAPES && env[1] == 0)
        {
            ctx->shape = env[0] - '0';
            ctx->shape_fixed = 1;
        }
    }
    cudaSetDevice(ctx->dev[0].id);
    e = cudaFuncGetAttributes(&ctx->fattr, (const void *)qr_kernel_of(true, ctx->shape));
    if (e != cudaSuccess)
    {
        int rc = qr_fail(NULL, QR_E_CUDA, "qr_init: kernel image not loadable on device %d: %s",
                         ctx->dev[0].id, cudaGetErrorString(e));
        qr_shutdown(ctx);
        return rc;
    }

    *out = ctx;
    return QR_OK;
}

extern "C" void qr_shutdown(qr_ctx *ctx)
{
    if (ctx == NULL)
    {
        return;
    }
    if (ctx->nhelper > 0)
    {
        {
            std::lock_guard<std::mutex> lk(ctx->hmtx);
            ctx->hquit = true;
            ctx->hcv.notify_all();
        }
        for (int k = 0; k < ctx->nhelper; k++) ctx->helper[k].join();
        ctx->nhelper = 0;
    }
    for (int k = 0; k < 4; k++)
    {
        if (ctx->pinned[k] != NULL) cudaHostUnregister(ctx->pinned[k]);
    }
    for (int i = 0; i < ctx->ndev; i++)
    {
        qr_dev &d = ctx->dev[i];
        cudaSetDevice(d.id);
        if (d.stream)  { cudaStreamSynchronize(d.stream); cudaStreamDestroy(d.stream); }
        if (d.copy)    { cudaStreamSynchronize(d.copy); cudaStreamDestroy(d.copy); }
        if (d.up)      { cudaStreamSynchronize(d.up); cudaStreamDestroy(d.up); }
        for (int k = 0; k < QR_MAX_CHUNKS; k++)
        {
            if (d.chunk_ev[k]) cudaEventDestroy(d.chunk_ev[k]);
            if (d.copy_ev[k])  cudaEventDestroy(d.copy_ev[k]);
        }
        if (d.ev0)     cudaEventDestroy(d.ev0);
        if (d.ev1)     cudaEventDestroy(d.ev1);
        if (d.done)    cudaEventDestroy(d.done);
        for (int k = 0; k < 2; k++)
        {
            if (d.blob_d[k])  cudaFree(d.blob_d[k]);
            if (d.blob_h[k])  cudaFreeHost(d.blob_h[k]);
            if (d.frame_p[k]) cudaFreeHost(d.frame_p[k]);
            if (d.frame_pd[k]) cudaFree(d.frame_pd[k]);
            if (d.pipe_ev[k]) cudaEventDestroy(d.pipe_ev[k]);
            if (d.fetch_ev[k]) cudaEventDestroy(d.fetch_ev[k]);
        }
        if (d.frame_d) cudaFree(d.frame_d);
        if (d.frame_h) cudaFreeHost(d.frame_h);
        if (d.t_d)     cudaFree(d.t_d);
        if (d.queue_d) cudaFree(d.queue_d);
        if (d.raw_d) cudaFree(d.raw_d);
        if (d.pt_seed) cudaFree(d.pt_seed);
        if (d.rays_d)  cudaFree(d.rays_d);
    }
    delete ctx;
}

static int qr_check_blob(qr_ctx *ctx, const void *blob, size_t bytes)
{
    const qr_blob_header *h = (const qr_blob_header *)blob;
    if (blob == NULL || bytes < sizeof(qr_blob_header))
    {
        return qr_fail(ctx, QR_E_BLOB, "scene blob too small");
    }
    if (h->magic != QR_BLOB_MAGIC || h->version != QR_BLOB_VERSION)
    {
        return qr_fail(ctx, QR_E_BLOB, "scene blob magic/version mismatch");
    }
    if (h->total_bytes > bytes || (h->total_bytes & 15) != 0)
    {
        return qr_fail(ctx, QR_E_BLOB, "scene blob truncated");
    }
    if (h->fsaa < 0 || h->fsaa > 2 || h->x_res <= 0 || h->y_res <= 0
    ||  h->tile_w <= 0 || h->tile_h <= 0 || h->depth < 0 || h->depth > QR_STACK_DEPTH
    ||  h->tls_row * h->tile_w < h->x_res || h->tls_col * h->tile_h < h->y_res
    ||  h->n_tiles != h->tls_row * h->tls_col)
    {
        return qr_fail(ctx, QR_E_BLOB, "scene blob header out of range");
    }
    const uint64_t total = h->total_bytes;
    if ((uint64_t)h->off_surf   + (uint64_t)h->n_surf   * sizeof(qr_surface)  > total
    ||  (uint64_t)h->off_mat    + (uint64_t)h->n_mat    * sizeof(qr_material) > total
    ||  (uint64_t)h->off_lgt    + (uint64_t)h->n_lgt    * sizeof(qr_light)    > total
    ||  (uint64_t)h->off_elem   + (uint64_t)h->n_elem   * sizeof(qr_elem)     > total
    ||  (uint64_t)h->off_tiles  + (uint64_t)h->n_tiles  * sizeof(int32_t)     > total
    ||  (uint64_t)h->off_texels + (uint64_t)h->n_texels * sizeof(uint32_t)    > total
    ||  h->n_surf < 0 || h->n_mat < 0 || h->n_lgt < 0 || h->n_elem < 0 || h->n_texels < 0)
    {
        return qr_fail(ctx, QR_E_BLOB, "scene blob section out of bounds");
    }
    return QR_OK;
}

static int qr_grow(qr_ctx *ctx, void **ptr, size_t *cap, size_t need, bool pinned)
{
    if (*cap >= need)
    {
        return QR_OK;
    }
    if (*ptr != NULL)
    {
        if (pinned) cudaFreeHost(*ptr); else cudaFree(*ptr);
        *ptr = NULL;
        *cap = 0;
    }
    size_t n = need + need / 4 + 4096;
    if (pinned)
    {
        QR_CUDA(ctx, cudaMallocHost(ptr, n));
    }
    else
    {
        /* The walk's element cursor is ONE 32-bit word (qr_core.cuh, qr_ecur):
         * no device buffer may straddle a 4 GB boundary.  cudaMalloc hardly
         * ever returns one that does; if so, allocate again while it is held. */
        void *held[4] = { NULL, NULL, NULL, NULL };
        int nheld = 0;
        cudaError_t err = cudaSuccess;
        for (;;)
        {
            err = cudaMalloc(ptr, n);
            if (err != cudaSuccess) break;
            const unsigned long long a = (unsigned long long)*ptr;
            if ((a >> 32) == ((a + n - 1) >> 32)) break;
            if (nheld == 4 || n >= (1ull << 32))
            {
                cudaFree(*ptr);
                *ptr = NULL;
                err = cudaErrorMemoryAllocation;
                break;
            }
            held[nheld++] = *ptr;
            *ptr = NULL;
        }
        for (int i = 0; i < nheld; i++) cudaFree(held[i]);
        QR_CUDA(ctx, err);
    }
    *cap = n;
    return QR_OK;
}

extern "C" int qr_scene_upload(qr_ctx *ctx, const void *blob, size_t bytes)
{
    if (ctx == NULL)
    {
        return QR_E_ARG;
    }
    int rc = qr_check_blob(ctx, blob, bytes);
    if (rc != QR_OK)
    {
        return rc;
    }
    const qr_blob_header *h = (const qr_blob_header *)blob;

    /* the blob is not copied verbatim: it is compiled into the packed kscene
     * image (qr_kscene.h) on the way into the pinned staging buffer */
    qr_kpacker &pk = ctx->packer;
    if (pk.plan(blob) != 0)
    {
        return qr_fail(ctx, QR_E_BLOB, "scene blob: malformed or not well nested list");
    }
    const size_t n = pk.bytes();

    /* from here on the slot, the header and the launch shape change step by
     * step: the context has no scene until every step has succeeded */
    ctx->have_scene = false;

    qr_dev &d0 = ctx->dev[0];
    QR_CUDA(ctx, cudaSetDevice(d0.id));
    if (ctx->pipelined)
    {
        /* the other slot: free once the frame begun in it two uploads ago is
         * complete (its kernels read the slot's device image, and the H2D
         * copies of the pinned staging precede them on the same streams) */
        ctx->slot ^= 1;
        if (d0.pipe_ev[ctx->slot] != NULL)
        {
            QR_CUDA(ctx, cudaEventSynchronize(d0.pipe_ev[ctx->slot]));
        }
    }
    else
    {
        /* the previous frame's H2D copies read the staging buffer */
        for (int i = 0; i < ctx->ndev; i++)
        {
            QR_CUDA(ctx, cudaStreamSynchronize(ctx->dev[i].stream));
        }
    }
    const int sl = ctx->slot;
    rc = qr_grow(ctx, (void **)&d0.blob_h[sl], &d0.blob_hcap[sl], n, true);
    if (rc != QR_OK)
    {
        return rc;
    }
    pk.write(d0.blob_h[sl]);
    const qr_blob_header *kh = (const qr_blob_header *)d0.blob_h[sl];

    for (int i = 0; i < ctx->ndev; i++)
    {
        qr_dev &d = ctx->dev[i];
        QR_CUDA(ctx, cudaSetDevice(d.id));
        rc = qr_grow(ctx, (void **)&d.blob_d[sl], &d.blob_cap[sl], pk.device_bytes(), false);
        if (rc != QR_OK)
        {
            return rc;
        }
        if (ctx->pipelined)
        {
            /* the slot is free (waited for above): copy on the copy stream, beside
             * the kernels of the frame in flight; the next kernel waits for it */
            QR_CUDA(ctx, cudaMemcpyAsync(d.blob_d[sl], d0.blob_h[sl], n, cudaMemcpyHostToDevice, d.up));
            QR_CUDA(ctx, cudaEventRecord(d.done, d.up));
            QR_CUDA(ctx, cudaStreamWaitEvent(d.stream, d.done, 0));
        }
        else
        {
            QR_CUDA(ctx, cudaMemcpyAsync(d.blob_d[sl], d0.blob_h[sl], n, cudaMemcpyHostToDevice, d.stream));
        }
        if (pk.device_tiling())
        {
            /* the engine left the tiling to us: tile rectangles of the camera
             * list's leaves, then one list per tile, ahead of the frame's kernels */
            const unsigned int nl = (unsigned int)kh->pad3[2], nt = (unsigned int)kh->n_tiles;
            qr_tile_rect_kernel<<<(nl + 127) / 128, 128, 0, d.stream>>>(d.blob_d[sl]);
            qr_tile_list_kernel<<<(nt + 127) / 128, 128, 0, d.stream>>>(d.blob_d[sl]);
            QR_CUDA(ctx, cudaGetLastError());
        }
    }
    ctx->device_tiling = pk.device_tiling();

    ctx->hdr = *h;

    /*
     * Launch shape by the amount of work of a full-frame launch.  The walk is
     * latency-bound (dependent FP chains, list elements through L1/L2), so the
     * big frames want every warp slot of the SM even at 64 registers with a
     * few spills (1080p 4xAA, RooT default scene: 1.87 ms at 1024 threads,
     * 2.06 ms at 640; 1080p without AA, 65 k items: 0.53 vs 0.57 ms); with few
     * work items per warp the smaller CTA's cleaner code and finer tail win
     * (800 x 480 4xAA, 48 k items: 0.18 vs 0.22 ms on demo01).
     */
    if (!ctx->shape_fixed)
    {
        const int bh = 8 >> h->fsaa;
        const uint64_t items = (uint64_t)h->tls_col * h->tls_row
                             * (uint64_t)((h->tile_h + bh - 1) / bh) * (uint64_t)((h->tile_w + 3) / 4);
        const int want = items >= (uint64_t)QR_BIG_FRAME_ITEMS * ctx->ndev ? QR_BIG_SHAPE : QR_DEFAULT_SHAPE;
        if (want != ctx->shape)
        {
            ctx->shape = want;
            QR_CUDA(ctx, cudaFuncGetAttributes(&ctx->fattr, (const void *)qr_kernel_of(true, ctx->shape)));
        }
    }

    /* shared-memory staging of the kscene prefix */
    const uint32_t prefix = kh->off_elem;
    const int slots = g_shapes[ctx->shape].threads * QR_SC_QUADS * 16;     /* per-thread scratch quads */
    const int budget = ctx->dev[0].smem_optin - (int)ctx->fattr.sharedSizeBytes - 1024 - slots;
    /* QR_B200_NOSTAGE=1 forces the variant that reads the scene through L1/L2
     * (what scenes too large for shared memory get) -- for testing */
    static const bool nostage = getenv("QR_B200_NOSTAGE") != NULL && getenv("QR_B200_NOSTAGE")[0] == '1';
    if (!nostage && (prefix & 15) == 0 && prefix < (1u << 20) && (int)prefix <= budget)
    {
        ctx->stage_bytes = prefix;
    }
    else
    {
        ctx->stage_bytes = 0;
    }

    for (int i = 0; i < ctx->ndev; i++)
    {
        qr_dev &d = ctx->dev[i];
        int nb = 0;
        QR_CUDA(ctx, cudaSetDevice(d.id));
        const void *fn = (const void *)qr_kernel_of(ctx->stage_bytes != 0, ctx->shape);
        QR_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)ctx->stage_bytes + slots));
        QR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn,
                     g_shapes[ctx->shape].threads, ctx->stage_bytes + slots));
        if (nb < 1)
        {
            return qr_fail(ctx, QR_E_CUDA, "kernel does not fit on device %d", d.id);
        }
        d.ctas_per_sm = nb;
    }
    ctx->pt_scene = (h->flags & QR_BLOB_PT) != 0;
    if (ctx->pt_scene)
    {
        /* the packet tracer reads the blob as it is; its seed and colour planes
         * are laid out for packets of 32 >> fsaa pixels (qr_pt.cuh) */
        if (ctx->pipelined || (h->x_row % (32 >> h->fsaa)) != 0 || h->x_row < h->x_res)
        {
            return qr_fail(ctx, QR_E_STATE, "path-traced scene: needs x_row %% %d == 0 and no pipelined mode",
                           32 >> h->fsaa);
        }
        for (int i = 0; i < ctx->ndev; i++)
        {
            qr_dev &d = ctx->dev[i];
            QR_CUDA(ctx, cudaSetDevice(d.id));
            rc = qr_grow(ctx, (void **)&d.raw_d, &d.raw_cap, h->total_bytes, false);
            if (rc != QR_OK)
            {
                return rc;
            }
            QR_CUDA(ctx, cudaMemcpyAsync(d.raw_d, blob, h->total_bytes, cudaMemcpyHostToDevice, d.stream));
        }
    }
    ctx->have_scene = true;
    return QR_OK;
}

extern "C" int qr_pt_reset(qr_ctx *ctx, const uint32_t *pseed, size_t n_slots)
{
    if (ctx == NULL || n_slots == 0)
    {
        return QR_E_ARG;
    }
    uint32_t *gen = NULL;
    if (pseed == NULL)
    {
        /* rt_Scene::reset_pseed, engine.cpp:3651-3685: a 48-bit LCG fills the plane */
        gen = (uint32_t *)malloc(n_slots * sizeof(uint32_t));
        if (gen == NULL)
        {
            return qr_fail(ctx, QR_E_CUDA, "qr_pt_reset: out of host memory");
        }
        unsigned long long seed = 1;
        for (size_t k = 0; k < n_slots; k++)
        {
            seed = (seed * 25214903917ull + 11ull) & 0x0000FFFFFFFFFFFFull;
            gen[k] = (uint32_t)seed;
        }
        pseed = gen;
    }
    int rc = QR_OK;
    for (int i = 0; i < ctx->ndev && rc == QR_OK; i++)
    {
        qr_dev &d = ctx->dev[i];
        cudaError_t e = cudaSetDevice(d.id);
        if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
        if (e == cudaSuccess && d.pt_slots != n_slots)
        {
            if (d.pt_seed != NULL) cudaFree(d.pt_seed);
            d.pt_seed = NULL;
            d.pt_slots = 0;
            /* one block: seeds, then the three colour planes */
            e = cudaMalloc((void **)&d.pt_seed, n_slots * 4 * sizeof(uint32_t));
            if (e == cudaSuccess)
            {
                d.pt_slots = n_slots;
                for (int k = 0; k < 3; k++) d.pt_col[k] = (float *)(d.pt_seed + (size_t)(k + 1) * n_slots);
            }
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(d.pt_seed, pseed, n_slots * sizeof(uint32_t), cudaMemcpyHostToDevice, d.stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(d.pt_col[0], 0, n_slots * 3 * sizeof(float), d.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
        if (e != cudaSuccess)
        {
            rc = qr_fail(ctx, QR_E_CUDA, "qr_pt_reset: %s", cudaGetErrorString(e));
        }
    }
    free(gen);
    ctx->pt_count = 0.0f;
    return rc;
}

extern "C" int qr_pt_frames(qr_ctx *ctx)
{
    return ctx == NULL ? 0 : (int)ctx->pt_count;
}

extern "C" int qr_pt_fetch(qr_ctx *ctx, uint32_t *pseed, float *ptr_r, float *ptr_g, float *ptr_b, size_t n_slots)
{
    if (ctx == NULL)
    {
        return QR_E_ARG;
    }
    qr_dev &d = ctx->dev[0];
    if (d.pt_seed == NULL || n_slots > d.pt_slots)
    {
        return qr_fail(ctx, QR_E_STATE, "qr_pt_fetch: no path-tracer state of that size");
    }
    QR_CUDA(ctx, cudaSetDevice(d.id));
    QR_CUDA(ctx, cudaStreamSynchronize(d.stream));
    float *dst[3] = { ptr_r, ptr_g, ptr_b };
    if (pseed != NULL) QR_CUDA(ctx, cudaMemcpy(pseed, d.pt_seed, n_slots * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; k++)
    {
        if (dst[k] != NULL) QR_CUDA(ctx, cudaMemcpy(dst[k], d.pt_col[k], n_slots * sizeof(float), cudaMemcpyDeviceToHost));
    }
    return QR_OK;
}

/* a path-traced frame begins: one more sample per pixel sample (tracer.cpp:1112-1124) */
static int qr_pt_frame(qr_ctx *ctx, const char *who)
{
    if (!ctx->pt_scene)
    {
        return QR_OK;
    }
    const qr_blob_header &h = ctx->hdr;
    const size_t need = ((size_t)h.x_row * (size_t)h.y_res) << h.fsaa;
    for (int i = 0; i < ctx->ndev; i++)
    {
        if (ctx->dev[i].pt_seed == NULL || ctx->dev[i].pt_slots < need)
        {
            return qr_fail(ctx, QR_E_STATE, "%s: the scene asks for the path tracer, qr_pt_reset first "
                           "(%zu slots needed)", who, need);
        }
    }
    ctx->pt_count = ctx->pt_count + 1.0f;
    ctx->pt_o = 1.0f / ctx->pt_count;
    ctx->pt_u = 1.0f - ctx->pt_o;
    return QR_OK;
}

/* the path tracer's launch of tile rows ty0, ty0 + step, ... on GPU i */
static int qr_launch_rows_pt(qr_ctx *ctx, int i, uint32_t *frame_dev, int stride,
                             int ty0, int step, int n, bool first, unsigned int *notify)
{
    qr_dev &d = ctx->dev[i];
    if (!d.pt_stack)
    {
        /* walk -> material -> walk recursion, up to 13 contexts deep */
        QR_CUDA(ctx, cudaDeviceSetLimit(cudaLimitStackSize, QR_PT_STACK));
        /* no shared memory in this kernel: all of the SM's array to L1, where the context stacks live */
        QR_CUDA(ctx, cudaFuncSetAttribute((const void *)qr_pt_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 0));
        d.pt_stack = true;
    }
    qr_pt_launch p;
    p.blob = d.raw_d;
    p.frame = frame_dev;
    p.stride = stride;
    p.ty0 = ty0;
    p.ty_step = step;
    p.n_trows = n;
    p.pseed = d.pt_seed;
    p.ptr_r = d.pt_col[0]; p.ptr_g = d.pt_col[1]; p.ptr_b = d.pt_col[2];
    p.pts_o = ctx->pt_o;
    p.pts_u = ctx->pt_u;
    /* the path tracer's own counter: word 1 of the queue block is free while no notify launch runs */
    p.queue = d.queue_d + 1;
    QR_CUDA(ctx, cudaMemsetAsync(p.queue, 0, sizeof(unsigned int), d.stream));
    if (first)
    {
        QR_CUDA(ctx, cudaEventRecord(d.ev0, d.stream));
    }
    const int ppp = 32 >> ctx->hdr.fsaa;
    const unsigned int packets = (unsigned int)n * (unsigned int)ctx->hdr.tile_h
                               * (unsigned int)((ctx->hdr.x_res + ppp - 1) / ppp);
    /* QR_B200_PT_CTAS=<n>: resident CTAs per SM the grid is sized for (tuning) */
    static const int pt_ctas = getenv("QR_B200_PT_CTAS") != NULL && atoi(getenv("QR_B200_PT_CTAS")) > 0
                             ? atoi(getenv("QR_B200_PT_CTAS")) : 8;
    unsigned int grid = (unsigned int)d.sm_count * (unsigned int)pt_ctas;
    const unsigned int need = (packets + QR_PT_THREADS / 32 - 1) / (QR_PT_THREADS / 32);
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    qr_pt_kernel<<<grid, QR_PT_THREADS, 0, d.stream>>>(p);
    QR_CUDA(ctx, cudaGetLastError());
    if (notify != NULL)
    {
        qr_add_kernel<<<1, 1, 0, d.stream>>>(notify);
        QR_CUDA(ctx, cudaGetLastError());
    }
    QR_CUDA(ctx, cudaEventRecord(d.ev1, d.stream));
    d.timed = true;
    ctx->launches++;
    return QR_OK;
}

/* tile rows ty0, ty0 + step, ... (n of them) on GPU i into frame_dev */
static int qr_launch_rows(qr_ctx *ctx, int i, uint32_t *frame_dev, int stride,
                          int ty0, int step, int n, float *t_out, bool first = true,
                          unsigned int *notify = NULL)
{
    qr_dev &d = ctx->dev[i];
    QR_CUDA(ctx, cudaSetDevice(d.id));
    d.timed = false;
    if (n <= 0)
    {
        return QR_OK;
    }
    if (ctx->pt_scene)
    {
        if (t_out != NULL)
        {
            return qr_fail(ctx, QR_E_STATE, "no dump mode for a path-traced scene");
        }
        return qr_launch_rows_pt(ctx, i, frame_dev, stride, ty0, step, n, first, notify);
    }

    qr_launch p;
    p.blob = d.blob_d[ctx->slot];
    p.frame = frame_dev;
    p.stride = stride;
    p.ty0 = ty0;
    p.ty_step = step;
    p.n_trows = n;
    p.stage_bytes = ctx->stage_bytes;
    p.queue = d.queue_d;
    p.rays = d.rays_d;
    p.t_out = t_out;
    p.notify = notify;
    p.done = d.queue_d + 1;

    const int bh = 8 >> ctx->hdr.fsaa;
    const unsigned int n_items = (unsigned int)n * ctx->hdr.tls_row
                               * (unsigned int)((ctx->hdr.tile_h + bh - 1) / bh)
                               * (unsigned int)((ctx->hdr.tile_w + 3) / 4);
    unsigned int grid = (unsigned int)(d.sm_count * d.ctas_per_sm);
    const int threads = g_shapes[ctx->shape].threads;
    const unsigned int warps = (unsigned int)threads / 32u;
    const unsigned int need = (n_items + warps - 1) / warps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;

    /* the work-item counter is not reset between launches: every warp of a
     * launch draws until it is handed an item past the end, exactly once, so
     * the counter ends at base + items + warps */
    if (d.queue_base > 0x7F000000u)
    {
        QR_CUDA(ctx, cudaMemsetAsync(d.queue_d, 0, sizeof(unsigned int), d.stream));
        d.queue_base = 0;
    }
    p.queue_base = d.queue_base;
    p.n_warps = grid * warps;
    d.queue_base += n_items + p.n_warps;
    if (first)
    {
        QR_CUDA(ctx, cudaEventRecord(d.ev0, d.stream));
    }
#if defined(QR_CHECKED)
    {
        const qr_blob_header *kh = (const qr_blob_header *)ctx->dev[0].blob_h[ctx->slot];
        qr_check_limits_t lim;
        lim.elems_lo = (unsigned long long)(p.blob + kh->off_elem);
        lim.elems_hi = lim.elems_lo + (unsigned long long)kh->n_elem * sizeof(qr_kelem);
        if (kh->pad0[0] != 0)
        {
            /* device-built tile lists: behind the uploaded image */
            lim.elems_hi = (unsigned long long)p.blob + (uint32_t)kh->pad0[0]
                         + (unsigned long long)kh->n_tiles * (uint32_t)kh->pad0[1] * sizeof(qr_kelem);
        }
        lim.surf_bytes = (uint32_t)kh->n_surf * QR_KSURF_QUADS * 16u;
        lim.n_tiles = (uint32_t)kh->n_tiles;
        QR_CUDA(ctx, cudaMemcpyToSymbolAsync(qr_check_limits, &lim, sizeof(lim), 0, cudaMemcpyHostToDevice, d.stream));
    }
#endif
#if defined(QR_ITEMLOG)
    qr_itemlog_t *il_dev = NULL;
    if (getenv("QR_B200_ITEM_LOG") != NULL)
    {
        QR_CUDA(ctx, cudaMalloc((void **)&il_dev, (size_t)n_items * sizeof(qr_itemlog_t)));
        QR_CUDA(ctx, cudaMemsetAsync(il_dev, 0, (size_t)n_items * sizeof(qr_itemlog_t), d.stream));
    }
    QR_CUDA(ctx, cudaMemcpyToSymbolAsync(qr_itemlog_ptr, &il_dev, sizeof(il_dev), 0, cudaMemcpyHostToDevice, d.stream));
#endif
    void *args[] = { (void *)&p };
    QR_CUDA(ctx, cudaLaunchKernel((const void *)qr_kernel_of(ctx->stage_bytes != 0, ctx->shape),
                                  dim3(grid), dim3(threads), args,
                                  ctx->stage_bytes + (size_t)threads * QR_SC_QUADS * 16u, d.stream));
    QR_CUDA(ctx, cudaGetLastError());
    QR_CUDA(ctx, cudaEventRecord(d.ev1, d.stream));
#if defined(QR_ITEMLOG)
    if (il_dev != NULL)
    {
        /* every launch overwrites the file: the last one of a run is kept */
        qr_itemlog_t *il = (qr_itemlog_t *)malloc((size_t)n_items * sizeof(qr_itemlog_t));
        QR_CUDA(ctx, cudaMemcpyAsync(il, il_dev, (size_t)n_items * sizeof(qr_itemlog_t), cudaMemcpyDeviceToHost, d.stream));
        QR_CUDA(ctx, cudaStreamSynchronize(d.stream));
        FILE *f = fopen(getenv("QR_B200_ITEM_LOG"), "wb");
        if (f != NULL)
        {
            fwrite(il, sizeof(qr_itemlog_t), n_items, f);
            fclose(f);
        }
        free(il);
        cudaFree(il_dev);
    }
#endif
    d.timed = true;
    ctx->launches++;
    {
        /* primary samples of these rows: known without asking the device */
        uint64_t rows = 0;
        for (int k = 0; k < n; k++)
        {
            int ya = (ty0 + k * step) * ctx->hdr.tile_h, yb = ya + ctx->hdr.tile_h;
            if (yb > ctx->hdr.y_res) yb = ctx->hdr.y_res;
            if (yb > ya) rows += (uint64_t)(yb - ya);
        }
        ctx->rays[0] += (rows * (uint64_t)ctx->hdr.x_res) << ctx->hdr.fsaa;
    }
    return QR_OK;
}

static int qr_collect_rays(qr_ctx *ctx)
{
    for (int i = 0; i < ctx->ndev; i++)
    {
        qr_dev &d = ctx->dev[i];
        unsigned long long r[4];
        QR_CUDA(ctx, cudaSetDevice(d.id));
        QR_CUDA(ctx, cudaMemcpyAsync(r, d.rays_d, sizeof(r), cudaMemcpyDeviceToHost, d.stream));
        QR_CUDA(ctx, cudaMemsetAsync(d.rays_d, 0, sizeof(r), d.stream));
        QR_CUDA(ctx, cudaStreamSynchronize(d.stream));
        for (int k = 0; k < 4; k++) ctx->rays[k] += r[k];
    }
    return QR_OK;
}

/* is the whole of [frame, frame + bytes) page-locked (so the copy engine or
 * the kernel can write it)?  Both ends are asked: a buffer that grew at the
 * same address is only partly registered */
static bool qr_frame_pinned(const void *frame, size_t bytes)
{
    cudaPointerAttributes at;
    bool yes = cudaPointerGetAttributes(&at, frame) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (yes && bytes > 1)
    {
        yes = cudaPointerGetAttributes(&at, (const uint8_t *)frame + bytes - 1) == cudaSuccess
           && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
    }
    return yes;
}

/*
 * QR_B200_PIN_FRAME=1: page-lock the application's framebuffer (RooT's XShm
 * image, core_test's frame) so the copy engine or the kernel writes it
 * directly.  The registry remembers how much of each buffer it registered: a
 * frame that grew, moved or changed its stride is registered again, the
 * oldest entry makes room.  Returns whether [frame, frame + bytes) is
 * page-locked now.
 */
static bool qr_pin_frame(qr_ctx *ctx, void *frame, size_t bytes)
{
    int slot = -1;
    ctx->pin_clock++;
    for (int k = 0; k < 4; k++)
    {
        if (ctx->pinned[k] == frame)
        {
            if (ctx->pinned_bytes[k] >= bytes)
            {
                ctx->pinned_age[k] = ctx->pin_clock;
                return qr_frame_pinned(frame, bytes);
            }
            cudaHostUnregister(ctx->pinned[k]);         /* grew at the same address */
            cudaGetLastError();
            ctx->pinned[k] = NULL;
            slot = k;
        }
    }
    /* an entry that overlaps the new range is stale (its buffer was freed) */
    for (int k = 0; k < 4; k++)
    {
        if (ctx->pinned[k] != NULL)
        {
            const uint8_t *a = (const uint8_t *)ctx->pinned[k], *b = (const uint8_t *)frame;
            if (a < b + bytes && b < a + ctx->pinned_bytes[k])
            {
                cudaHostUnregister(ctx->pinned[k]);
                cudaGetLastError();
                ctx->pinned[k] = NULL;
                if (slot < 0) slot = k;
            }
        }
    }
    for (int k = 0; k < 4 && slot < 0; k++)
    {
        if (ctx->pinned[k] == NULL) slot = k;
    }
    if (slot < 0)
    {
        slot = 0;
        for (int k = 1; k < 4; k++)
        {
            if ((int)(ctx->pinned_age[k] - ctx->pinned_age[slot]) < 0) slot = k;
        }
        cudaHostUnregister(ctx->pinned[slot]);
        cudaGetLastError();
        ctx->pinned[slot] = NULL;
    }
    if (cudaHostRegister(frame, bytes, cudaHostRegisterDefault) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    ctx->pinned[slot] = frame;
    ctx->pinned_bytes[slot] = bytes;
    ctx->pinned_age[slot] = ctx->pin_clock;
    return true;
}

/*
 * Helper threads: copy a frame (or a chunk of rows) that arrived in pinned
 * staging into the caller's pageable frame -- the engine's own frame, RooT's
 * XShm image.  One thread moves ~11 GB/s, i.e. a 1080p frame in 0.75 ms; four
 * of them 0.25 ms, and in pipelined mode they do it while the caller flattens
 * and uploads the next scene.  Two job slots.
 */
static void qr_copy_rows(const qr_copy_job &j, int part = 0, int parts = 1)
{
    const size_t wbytes = (size_t)j.x_res * sizeof(uint32_t);
    const int n = j.yb - j.ya;
    const int ya = j.ya + (int)((long long)n * part / parts), yb = j.ya + (int)((long long)n * (part + 1) / parts);
    if (j.st == j.dstride && j.st == j.x_res)
    {
        memcpy(j.dst + (size_t)ya * j.st, j.src + (size_t)ya * j.dstride, wbytes * (size_t)(yb - ya));
        return;
    }
    for (int y = ya; y < yb; y++)
    {
        memcpy(j.dst + (ptrdiff_t)y * j.st, j.src + (size_t)y * j.dstride, wbytes);
    }
}

static void qr_helper_main(qr_ctx *ctx, int k)
{
    {
        /* the thread that starts the helpers is typically a worker pinned to
         * one core (root/RooT_linux.cpp:681-699); they must not share that
         * core with it */
        /* ... nor escape the CPU set the process was given (taskset, cgroup
         * pinning of a comparison run): the mask captured when the library
         * was loaded, before any worker pinned itself */
        if (g_load_affinity_ok) sched_setaffinity(0, sizeof(g_load_affinity), &g_load_affinity);
    }
    cudaSetDevice(ctx->dev[0].id);
    std::unique_lock<std::mutex> lk(ctx->hmtx);
    for (;;)
    {
        int t = -1;
        ctx->hcv.wait(lk, [ctx, k, &t]
        {
            if (ctx->hquit) return true;
            for (int s = 0; s < 2; s++)
            {
                if (ctx->hjob[s] == 1 && !ctx->htaken[s][k]) { t = s; return true; }
            }
            return false;
        });
        if (ctx->hquit)
        {
            return;
        }
        ctx->htaken[t][k] = true;
        const qr_copy_job j = ctx->job[t];
        lk.unlock();
        /* every helper copies its band of the rows once they are in staging */
        const cudaError_t e = cudaEventSynchronize(j.ev);
        if (e == cudaSuccess)
        {
            qr_copy_rows(j, k, ctx->nhelper);
        }
        lk.lock();
        if (e != cudaSuccess) ctx->hfail[t] = true;
        if (--ctx->hleft[t] == 0)
        {
            ctx->hjob[t] = ctx->hfail[t] ? 3 : 2;
            ctx->hfail[t] = false;
            memset(ctx->htaken[t], 0, sizeof(ctx->htaken[t]));
            ctx->hcv.notify_all();
        }
    }
}

static void qr_helpers_start(qr_ctx *ctx)
{
    if (ctx->nhelper > 0)
    {
        return;
    }
    ctx->hquit = false;
    ctx->hjob[0] = ctx->hjob[1] = 0;
    memset(ctx->htaken, 0, sizeof(ctx->htaken));
    int n = 4;
    {
        const char *env = getenv("QR_B200_COPY_THREADS");
        if (env != NULL && atoi(env) >= 1 && atoi(env) <= QR_COPY_THREADS)
        {
            n = atoi(env);
        }
    }
    ctx->nhelper = n;           /* before the threads start: they divide the rows by it */
    for (int k = 0; k < n; k++)
    {
        ctx->helper[k] = std::thread(qr_helper_main, ctx, k);
    }
}

/* hand job "j" to the helpers in slot t; false when the slot is busy */
static bool qr_helpers_post(qr_ctx *ctx, int t, const qr_copy_job &j)
{
    if (ctx->nhelper == 0)
    {
        return false;
    }
    std::lock_guard<std::mutex> lk(ctx->hmtx);
    if (ctx->hjob[t] != 0)
    {
        return false;
    }
    ctx->job[t] = j;
    ctx->hjob[t] = 1;
    ctx->hleft[t] = ctx->nhelper;
    ctx->hcv.notify_all();
    return true;
}

/* wait for slot t: 1 copied, 0 nothing was posted, -1 failed */
static int qr_helpers_wait(qr_ctx *ctx, int t)
{
    std::unique_lock<std::mutex> lk(ctx->hmtx);
    if (ctx->hjob[t] == 0)
    {
        return 0;
    }
    ctx->hcv.wait(lk, [ctx, t] { return ctx->hjob[t] >= 2; });
    const int r = ctx->hjob[t] == 2 ? 1 : -1;
    ctx->hjob[t] = 0;
    return r;
}

static int qr_frame_ensure(qr_ctx *ctx)
{
    const qr_blob_header &h = ctx->hdr;
    const int stride = h.x_row >= h.x_res ? h.x_row : h.x_res;
    const size_t fbytes = (size_t)stride * h.y_res * sizeof(uint32_t);
    qr_dev &d0 = ctx->dev[0];
    QR_CUDA(ctx, cudaSetDevice(d0.id));
    /* the frame is followed by QR_NOTIFY_SLOTS words of completion counters
     * (qr_frame_notify_slot): zero when the buffer is (re)allocated */
    const size_t need = fbytes + QR_NOTIFY_SLOTS * sizeof(uint32_t);
    const bool fresh = d0.frame_cap < need;
    int rc = qr_grow(ctx, (void **)&d0.frame_d, &d0.frame_cap, need, false);
    if (rc == QR_OK && fresh)
    {
        QR_CUDA(ctx, cudaMemsetAsync((uint8_t *)d0.frame_d + fbytes, 0, QR_NOTIFY_SLOTS * sizeof(uint32_t), d0.stream));
    }
    return rc;
}

extern "C" int qr_render_device(qr_ctx *ctx, uint32_t *frame_dev, int stride, int y0, int y1)
{
    if (ctx == NULL)
    {
        return QR_E_ARG;
    }
    if (!ctx->have_scene)
    {
        return qr_fail(ctx, QR_E_STATE, "qr_render_device: no scene uploaded");
    }
    const qr_blob_header &h = ctx->hdr;
    if (frame_dev == NULL || stride < h.x_res || y0 < 0 || y1 > h.y_res || y0 > y1
    ||  (y0 % h.tile_h) != 0)
    {
        return qr_fail(ctx, QR_E_ARG, "qr_render_device: bad arguments");
    }
    {
        const int prc = qr_pt_frame(ctx, "qr_render_device");
        if (prc != QR_OK) return prc;
    }
    const int ty0 = y0 / h.tile_h;
    const int ty1 = (y1 + h.tile_h - 1) / h.tile_h;
    for (int i = 1; i < ctx->ndev; i++) ctx->dev[i].timed = false;
    return qr_launch_rows(ctx, 0, frame_dev, stride, ty0, 1, ty1 - ty0, NULL);
}

extern "C" int qr_render_rows(qr_ctx *ctx, uint32_t *frame_dev, int stride, int tile_row0, int tile_row_step)
{
    if (ctx == NULL)
    {
        return QR_E_ARG;
    }
    if (!ctx->have_scene)
    {
        return qr_fail(ctx, QR_E_STATE, "qr_render_rows: no scene uploaded");
    }
    const qr_blob_header &h = ctx->hdr;
    if (frame_dev == NULL || stride < h.x_res || tile_row0 < 0 || tile_row_step < 1)
    {
        return qr_fail(ctx, QR_E_ARG, "qr_render_rows: bad arguments");
    }
    return qr_render_rows_notify(ctx, frame_dev, stride, tile_row0, tile_row_step, NULL);
}

extern "C" int qr_render_rows_notify(qr_ctx *ctx, uint32_t *frame_dev, int stride, int tile_row0,
                                     int tile_row_step, uint32_t *notify_dev)
{
    if (ctx == NULL)
    {
        return QR_E_ARG;
    }
    if (!ctx->have_scene)
    {
        return qr_fail(ctx, QR_E_STATE, "qr_render_rows: no scene uploaded");
    }
    const qr_blob_header &h = ctx->hdr;
    if (frame_dev == NULL || stride < h.x_res || tile_row0 < 0 || tile_row_step < 1)
    {
        return qr_fail(ctx, QR_E_ARG, "qr_render_rows: bad arguments");
    }
    const int n = tile_row0 < h.tls_col ? (h.tls_col - tile_row0 + tile_row_step - 1) / tile_row_step : 0;
    for (int i = 1; i < ctx->ndev; i++) ctx->dev[i].timed = false;
    {
        const int prc = qr_pt_frame(ctx, "qr_render_rows");
        if (prc != QR_OK) return prc;
    }
    int rc = qr_launch_rows(ctx, 0, frame_dev, stride, tile_row0, tile_row_step, n, NULL, true,
                            (unsigned int *)notify_dev);
    if (rc == QR_OK && n <= 0 && notify_dev != NULL)
    {
        /* nothing to render for this rank: the signal is still owed */
        qr_dev &d0 = ctx->dev[0];
        qr_add_kernel<<<1, 1, 0, d0.stream>>>((unsigned int *)notify_dev);
        QR_CUDA(ctx, cudaGetLastError());
    }
    if (rc == QR_OK && ctx->pipelined)
    {
        /* pipelined mode: the next upload into this scene slot waits for the
         * kernels that read it */
        qr_dev &d0 = ctx->dev[0];
        if (d0.pipe_ev[ctx->slot] != NULL)
        {
            QR_CUDA(ctx, cudaEventRecord(d0.pipe_ev[ctx->slot], d0.stream));
        }
    }
    return rc;
}

extern "C" int qr_wait_notify(qr_ctx *ctx, uint32_t *notify_dev, uint32_t target)
{
    if (ctx == NULL || notify_dev == NULL)
    {
        return QR_E_ARG;
    }
    qr_dev &d0 = ctx->dev[0];
    QR_CUDA(ctx, cudaSetDevice(d0.id));
    qr_wait_kernel<<<1, 1, 0, d0.stream>>>((unsigned int *)notify_dev, (unsigned int)target);
    QR_CUDA(ctx, cudaGetLastError());
    return QR_OK;
}

#if defined(QR_CHECKED)
/* QR_B200_CHECK_SELFTEST=1: every counter is made to fire once (plumbing test) */
__global__ void qr_check_selftest_kernel()
{
    for (int k = 0; k < 7; k++) QR_CHECK(threadIdx.x > 0, k);
}
#endif

extern "C" int qr_check_counters(qr_ctx *ctx, uint32_t counters[8])
{
    if (ctx == NULL || counters == NULL)
    {
        return QR_E_ARG;
    }
#if defined(QR_CHECKED)
    if (getenv("QR_B200_CHECK_SELFTEST") != NULL)
    {
        QR_CUDA(ctx, cudaSetDevice(ctx->dev[0].id));
        qr_check_selftest_kernel<<<1, 1, 0, ctx->dev[0].stream>>>();
        QR_CUDA(ctx, cudaGetLastError());
    }
    unsigned int c[QR_CHECK_COUNTERS];
    const unsigned int zero[QR_CHECK_COUNTERS] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    for (int i = 0; i < ctx->ndev; i++)
    {
        QR_CUDA(ctx, cudaSetDevice(ctx->dev[i].id));
        QR_CUDA(ctx, cudaStreamSynchronize(ctx->dev[i].stream));
        QR_CUDA(ctx, cudaMemcpyFromSymbol(c, qr_check_count, sizeof(c)));
        QR_CUDA(ctx, cudaMemcpyToSymbol(qr_check_count, zero, sizeof(zero)));
        for (int k = 0; k < 8; k++) counters[k] = (i == 0 ? 0u : counters[k]) + c[k];
    }
    return QR_OK;
#else
    return qr_fail(ctx, QR_E_STATE, "qr_check_counters: this is not the checked build (make checked)");
#endif
}

extern "C" int qr_host_register(qr_ctx *ctx, void *host, size_t bytes, uint32_t **dev_ptr)
{
    if (ctx == NULL || host == NULL || bytes == 0 || dev_ptr == NULL)
    {
        return QR_E_ARG;
    }
    QR_CUDA(ctx, cudaSetDevice(ctx->dev[0].id));
    QR_CUDA(ctx, cudaHostRegister(host, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    void *d = NULL;
    cudaError_t e = cudaHostGetDevicePointer(&d, host, 0);
    if (e != cudaSuccess)
    {
        cudaHostUnregister(host);
        QR_CUDA(ctx, e);
    }
    *dev_ptr = (uint32_t *)d;
    return QR_OK;
}

extern "C" int qr_host_unregister(qr_ctx *ctx, void *host)
{
    if (ctx == NULL || host == NULL)
    {
        return QR_E_ARG;
    }
    QR_CUDA(ctx, cudaSetDevice(ctx->dev[0].id));
    QR_CUDA(ctx, cudaHostUnregister(host));
    return QR_OK;
}

extern "C" int qr_frame_notify_slot(qr_ctx *ctx, uint32_t *frame_dev, int index, uint32_t **slot_dev)
{
    if (ctx == NULL || frame_dev == NULL || slot_dev == NULL || index < 0 || index >= QR_NOTIFY_SLOTS)
    {
        return QR_E_ARG;
    }
    if (!ctx->have_scene)
    {
        return qr_fail(ctx, QR_E_STATE, "qr_frame_notify_slot: no scene uploaded");
    }
    const qr_blob_header &h = ctx->hdr;
    const int stride = h.x_row >= h.x_res ? h.x_row : h.x_res;
    *slot_dev = frame_dev + (size_t)stride * h.y_res + (size_t)index;
    return QR_OK;
}

/*
 * One frame on all GPUs of the context.  Tile rows are dealt round-robin
 * (GPU i renders rows i, i + ndev, ...: neighbouring rows cost about the same,
 * so the GPUs finish together).  GPU 0 owns the framebuffer; a peer with P2P
 * access stores its pixels straight into it over NVLink from the kernel's
 * epilogue -- the gather costs no extra pass -- otherwise it renders into a
 * local buffer of the same geometry and the rows are copied over afterwards.
 */
static int qr_render_all(qr_ctx *ctx, float *t_out_dev)
{
    const qr_blob_header &h = ctx->hdr;
    const int stride = h.x_row >= h.x_res ? h.x_row : h.x_res;
    const size_t fbytes = (size_t)stride * h.y_res * sizeof(uint32_t);
    const int ndev = t_out_dev != NULL ? 1 : ctx->ndev;
    int rc;

    rc = qr_frame_ensure(ctx);
    if (rc != QR_OK)
    {
        return rc;
    }
    qr_dev &d0 = ctx->dev[0];
    for (int i = 1; i < ndev; i++)
    {
        qr_dev &d = ctx->dev[i];
        if (!d.peer_ok)
        {
            QR_CUDA(ctx, cudaSetDevice(d.id));
            rc = qr_grow(ctx, (void **)&d.frame_d, &d.frame_cap, fbytes, false);
            if (rc != QR_OK)
            {
                return rc;
            }
        }
    }
    for (int i = ndev; i < ctx->ndev; i++) ctx->dev[i].timed = false;

    for (int i = 0; i < ndev; i++)
    {
        qr_dev &d = ctx->dev[i];
        const int n = i < h.tls_col ? (h.tls_col - i + ndev - 1) / ndev : 0;
        uint32_t *dst = (i == 0 || d.peer_ok) ? d0.frame_d : d.frame_d;
        rc = qr_launch_rows(ctx, i, dst, stride, i, ndev, n, t_out_dev);
        if (rc != QR_OK)
        {
            return rc;
        }
    }

    /* GPU 0's stream continues when every peer's rows have landed */
    for (int i = 1; i < ndev; i++)
    {
        qr_dev &d = ctx->dev[i];
        QR_CUDA(ctx, cudaSetDevice(d.id));
        if (!d.peer_ok)
        {
            for (int ty = i; ty < h.tls_col; ty += ndev)
            {
                int ya = ty * h.tile_h, yb = ya + h.tile_h;
                if (yb > h.y_res) yb = h.y_res;
                const size_t off = (size_t)ya * stride;
                const size_t n = (size_t)(yb - ya) * stride * sizeof(uint32_t);
                QR_CUDA(ctx, cudaMemcpyPeerAsync(d0.frame_d + off, d0.id, d.frame_d + off, d.id, n, d.stream));
            }
        }
        QR_CUDA(ctx, cudaEventRecord(d.done, d.stream));
        QR_CUDA(ctx, cudaSetDevice(d0.id));
        QR_CUDA(ctx, cudaStreamWaitEvent(d0.stream, d.done, 0));
    }
    return QR_OK;
}

extern "C" int qr_render(qr_ctx *ctx, uint32_t *frame, int stride)
{
    if (ctx == NULL)
    {
        return QR_E_ARG;
    }
    if (!ctx->have_scene)
    {
        return qr_fail(ctx, QR_E_STATE, "qr_render: no scene uploaded");
    }
    const qr_blob_header &h = ctx->hdr;
    if (frame != NULL && (stride < h.x_res && -stride < h.x_res))
    {
        return qr_fail(ctx, QR_E_ARG, "qr_render: stride smaller than x_res");
    }
    {
        const int prc = qr_pt_frame(ctx, "qr_render");
        if (prc != QR_OK) return prc;
    }
    if (frame == NULL)
    {
        return qr_render_all(ctx, NULL);
    }

    qr_dev &d0 = ctx->dev[0];
    const int dstride = h.x_row >= h.x_res ? h.x_row : h.x_res;
    const size_t fbytes = (size_t)dstride * h.y_res * sizeof(uint32_t);
    int rc;
    QR_CUDA(ctx, cudaSetDevice(d0.id));

    /* a caller's frame in page-locked memory (XShm segments are not, but a
     * registered or cudaMallocHost'ed buffer is) takes the D2H directly;
     * anything else, and bottom-up frames, go through pinned staging */
    bool direct = false;
    if (stride > 0)
    {
        const size_t ubytes = ((size_t)stride * (h.y_res - 1) + h.x_res) * sizeof(uint32_t);
        direct = qr_frame_pinned(frame, ubytes);
        if (!direct && ctx->pin_frames)
        {
            direct = qr_pin_frame(ctx, frame, ubytes);
        }
    }
    if (!direct)
    {
        rc = qr_grow(ctx, (void **)&d0.frame_h, &d0.frame_hcap, fbytes, true);
        if (rc != QR_OK)
        {
            return rc;
        }
    }
    uint32_t *dst = direct ? frame : d0.frame_h;
    const size_t dpitch = (size_t)(direct ? stride : dstride) * sizeof(uint32_t);
    const size_t spitch = (size_t)dstride * sizeof(uint32_t);
    const size_t wbytes = (size_t)h.x_res * sizeof(uint32_t);

    /* chunks of tile rows: the D2H (and the host copy out of staging) of
     * chunk k overlaps the rendering of chunk k + 1.  Every chunk boundary
     * drains the GPU once, which costs more than the 0.2 ms D2H of a 1080p
     * frame it could hide (measured: tools/e2e_breakdown.py), so a
     * page-locked frame is rendered in one piece; with staging, two chunks
     * hide half of the host-side copy. */
    int nch = ctx->ndev == 1 ? (ctx->chunks > 0 ? ctx->chunks : (direct ? 1 : 2)) : 1;
    if (nch > h.tls_col) nch = h.tls_col;
    if (nch < 1) nch = 1;

    /* (QR_B200_ZEROCOPY=0 turns this off) a page-locked frame is device-addressable (UVA), so
     * the kernel can store its pixels straight into host memory; the PCIe
     * writes then overlap the rendering and there is no D2H pass at all */
    if (ctx->ndev == 1 && direct && ctx->zerocopy)
    {
        uint32_t *dev_view = NULL;
        if (cudaHostGetDevicePointer((void **)&dev_view, frame, 0) == cudaSuccess && dev_view != NULL)
        {
            rc = qr_launch_rows(ctx, 0, dev_view, stride, 0, 1, h.tls_col, NULL);
            if (rc != QR_OK)
            {
                return rc;
            }
            QR_CUDA(ctx, cudaStreamSynchronize(d0.stream));
            return QR_OK;
        }
        cudaGetLastError();
    }

    if (ctx->ndev == 1)
    {
        rc = qr_frame_ensure(ctx);
        if (rc != QR_OK)
        {
            return rc;
        }
    }
    for (int k = 0; k < nch; k++)
    {
        const int ta = (int)((long long)h.tls_col * k / nch);
        const int tb = (int)((long long)h.tls_col * (k + 1) / nch);
        int ya = ta * h.tile_h, yb = tb * h.tile_h;
        if (yb > h.y_res) yb = h.y_res;
        if (ctx->ndev == 1)
        {
            rc = qr_launch_rows(ctx, 0, d0.frame_d, dstride, ta, 1, tb - ta, NULL, k == 0);
        }
        else
        {
            rc = qr_render_all(ctx, NULL);
        }
        if (rc != QR_OK)
        {
            return rc;
        }
        QR_CUDA(ctx, cudaEventRecord(d0.chunk_ev[k], d0.stream));
        QR_CUDA(ctx, cudaStreamWaitEvent(d0.copy, d0.chunk_ev[k], 0));
        QR_CUDA(ctx, cudaMemcpy2DAsync((uint8_t *)dst + (size_t)ya * dpitch, dpitch,
                                       (const uint8_t *)d0.frame_d + (size_t)ya * spitch, spitch,
                                       wbytes, (size_t)(yb - ya), cudaMemcpyDeviceToHost, d0.copy));
        QR_CUDA(ctx, cudaEventRecord(d0.copy_ev[k], d0.copy));
    }
    /* out of staging into the caller's pageable frame: by the helper threads,
     * chunk by chunk as the copies land (two job slots), else on this thread */
    bool posted[QR_MAX_CHUNKS];
    for (int k = 0; k < nch; k++) posted[k] = false;
    if (!direct && nch <= 2)
    {
        qr_helpers_start(ctx);
        for (int k = 0; k < nch; k++)
        {
            const int ta = (int)((long long)h.tls_col * k / nch);
            const int tb = (int)((long long)h.tls_col * (k + 1) / nch);
            qr_copy_job j;
            j.dst = frame; j.src = d0.frame_h; j.st = stride; j.dstride = dstride; j.x_res = h.x_res;
            j.ya = ta * h.tile_h; j.yb = tb * h.tile_h;
            if (j.yb > h.y_res) j.yb = h.y_res;
            j.ev = d0.copy_ev[k];
            posted[k] = qr_helpers_post(ctx, k, j);
        }
    }
    for (int k = 0; k < nch; k++)
    {
        if (posted[k])
        {
            if (qr_helpers_wait(ctx, k) < 0)
            {
                return qr_fail(ctx, QR_E_CUDA, "qr_render: frame transfer failed");
            }
            continue;
        }
        QR_CUDA(ctx, cudaEventSynchronize(d0.copy_ev[k]));
        if (!direct)
        {
            const int ta = (int)((long long)h.tls_col * k / nch);
            const int tb = (int)((long long)h.tls_col * (k + 1) / nch);
            int ya = ta * h.tile_h, yb = tb * h.tile_h;
            if (yb > h.y_res) yb = h.y_res;
            for (int y = ya; y < yb; y++)
            {
                memcpy(frame + (ptrdiff_t)y * stride, d0.frame_h + (size_t)y * dstride, wbytes);
            }
        }
    }
    return QR_OK;
}

extern "C" int qr_pipeline(qr_ctx *ctx, int on)
{
    if (ctx == NULL)
    {
        return QR_E_ARG;
    }
    int rc = qr_sync(ctx);
    if (rc != QR_OK)
    {
        return rc;
    }
    if (ctx->nhelper > 0)
    {
        /* a frame on its way to a caller's buffer arrives before the mode changes */
        std::unique_lock<std::mutex> lk(ctx->hmtx);
        ctx->hcv.wait(lk, [ctx] { return ctx->hjob[0] != 1 && ctx->hjob[1] != 1; });
        ctx->hjob[0] = ctx->hjob[1] = 0;
    }
    qr_dev &d0 = ctx->dev[0];
    QR_CUDA(ctx, cudaSetDevice(d0.id));
    for (int k = 0; k < 2; k++)
    {
        ctx->pending[k] = false;
        ctx->fetching[k] = 0;
        if (on && d0.pipe_ev[k] == NULL)
        {
            QR_CUDA(ctx, cudaEventCreateWithFlags(&d0.pipe_ev[k], cudaEventDisableTiming));
            QR_CUDA(ctx, cudaEventCreateWithFlags(&d0.fetch_ev[k], cudaEventDisableTiming));
        }
    }
    if ((on != 0) != (ctx->pipelined != 0))
    {
        /* the current scene lives in the slot it was uploaded to; after a mode
         * change the next upload decides */
        ctx->have_scene = false;
        ctx->slot = 0;
    }
    ctx->pipelined = on != 0;
    if (on)
    {
        qr_helpers_start(ctx);
    }
    return QR_OK;
}

extern "C" int qr_render_begin(qr_ctx *ctx, int *ticket)
{
    if (ctx == NULL || ticket == NULL)
    {
        return QR_E_ARG;
    }
    if (!ctx->pipelined || !ctx->have_scene)
    {
        return qr_fail(ctx, QR_E_STATE, "qr_render_begin: needs qr_pipeline(ctx, 1) and an uploaded scene");
    }
    const int sl = ctx->slot;
    if (ctx->pending[sl])
    {
        return qr_fail(ctx, QR_E_STATE, "qr_render_begin: the frame begun in this slot was not collected");
    }
    const qr_blob_header &h = ctx->hdr;
    const int dstride = h.x_row >= h.x_res ? h.x_row : h.x_res;
    const size_t fbytes = (size_t)dstride * h.y_res * sizeof(uint32_t);
    qr_dev &d0 = ctx->dev[0];
    int rc;
    QR_CUDA(ctx, cudaSetDevice(d0.id));

    /* the frame is rendered into device memory of its slot; it leaves the GPU
     * when it is fetched (qr_render_fetch / qr_render_end), by DMA straight
     * into a page-locked caller frame or through the slot's pinned staging */
    rc = qr_grow(ctx, (void **)&d0.frame_pd[sl], &d0.frame_pdcap[sl], fbytes, false);
    if (rc != QR_OK)
    {
        return rc;
    }
    if (ctx->ndev == 1)
    {
        rc = qr_launch_rows(ctx, 0, d0.frame_pd[sl], dstride, 0, 1, h.tls_col, NULL);
    }
    else
    {
        uint32_t *keep = d0.frame_d;
        const size_t keep_cap = d0.frame_cap;
        d0.frame_d = d0.frame_pd[sl];           /* qr_render_all gathers into GPU 0's "frame_d" */
        d0.frame_cap = d0.frame_pdcap[sl];
        rc = qr_render_all(ctx, NULL);
        d0.frame_d = keep;
        d0.frame_cap = keep_cap;
    }
    if (rc != QR_OK)
    {
        return rc;
    }
    QR_CUDA(ctx, cudaSetDevice(d0.id));
    QR_CUDA(ctx, cudaEventRecord(d0.pipe_ev[sl], d0.stream));
    ctx->pending[sl] = true;
    ctx->fetching[sl] = 0;
    ctx->pend_hdr[sl] = h;
    *ticket = sl;
    return QR_OK;
}

/*
 * Start moving the frame of "ticket" to the caller (after its kernels): a
 * page-locked frame with a positive stride is written by the copy engine
 * directly, anything else goes through the slot's pinned staging and is copied
 * by qr_render_end.  Returns at once; the caller overlaps host work.
 */
extern "C" int qr_render_fetch(qr_ctx *ctx, int ticket, uint32_t *frame, int stride)
{
    if (ctx == NULL || ticket < 0 || ticket > 1 || frame == NULL)
    {
        return QR_E_ARG;
    }
    if (!ctx->pending[ticket] || ctx->fetching[ticket] != 0)
    {
        return qr_fail(ctx, QR_E_STATE, "qr_render_fetch: no frame in flight for this ticket, or already being fetched");
    }
    const qr_blob_header &h = ctx->pend_hdr[ticket];
    if (stride < h.x_res && -stride < h.x_res)
    {
        return qr_fail(ctx, QR_E_ARG, "qr_render_fetch: stride smaller than x_res");
    }
    const int dstride = h.x_row >= h.x_res ? h.x_row : h.x_res;
    const size_t fbytes = (size_t)dstride * h.y_res * sizeof(uint32_t);
    const size_t wbytes = (size_t)h.x_res * sizeof(uint32_t);
    qr_dev &d0 = ctx->dev[0];
    QR_CUDA(ctx, cudaSetDevice(d0.id));
    QR_CUDA(ctx, cudaStreamWaitEvent(d0.copy, d0.pipe_ev[ticket], 0));

    const size_t ubytes = stride > 0 ? ((size_t)stride * (h.y_res - 1) + h.x_res) * sizeof(uint32_t) : 0;
    bool direct = stride > 0 && qr_frame_pinned(frame, ubytes);
    if (!direct && stride > 0 && ctx->pin_frames)
    {
        /* QR_B200_PIN_FRAME=1: let the copy engine write the application's frame */
        direct = qr_pin_frame(ctx, frame, ubytes);
    }
    if (direct)
    {
        QR_CUDA(ctx, cudaMemcpy2DAsync(frame, (size_t)stride * sizeof(uint32_t),
                                       d0.frame_pd[ticket], (size_t)dstride * sizeof(uint32_t),
                                       wbytes, (size_t)h.y_res, cudaMemcpyDeviceToHost, d0.copy));
        ctx->fetching[ticket] = 1;
    }
    else
    {
        int rc = qr_grow(ctx, (void **)&d0.frame_p[ticket], &d0.frame_pcap[ticket], fbytes, true);
        if (rc != QR_OK)
        {
            return rc;
        }
        QR_CUDA(ctx, cudaMemcpyAsync(d0.frame_p[ticket], d0.frame_pd[ticket], fbytes, cudaMemcpyDeviceToHost, d0.copy));
        ctx->fetching[ticket] = 2;
    }
    QR_CUDA(ctx, cudaEventRecord(d0.fetch_ev[ticket], d0.copy));
    ctx->fetch_dst[ticket] = frame;
    ctx->fetch_stride[ticket] = stride;
    if (ctx->fetching[ticket] == 2)
    {
        /* the helper threads take the frame from staging to the caller */
        qr_copy_job j;
        j.dst = frame; j.src = d0.frame_p[ticket]; j.st = stride; j.dstride = dstride;
        j.x_res = h.x_res; j.ya = 0; j.yb = h.y_res; j.ev = d0.fetch_ev[ticket];
        ctx->fetch_posted[ticket] = qr_helpers_post(ctx, ticket, j);
    }
    return QR_OK;
}

extern "C" int qr_render_end(qr_ctx *ctx, int ticket, uint32_t *frame, int stride)
{
    if (ctx == NULL || ticket < 0 || ticket > 1)
    {
        return QR_E_ARG;
    }
    if (!ctx->pending[ticket])
    {
        return qr_fail(ctx, QR_E_STATE, "qr_render_end: no frame in flight for this ticket");
    }
    qr_dev &d0 = ctx->dev[0];
    QR_CUDA(ctx, cudaSetDevice(d0.id));
    if (ctx->fetching[ticket] == 0)
    {
        if (frame == NULL)
        {
            /* wait and drop */
            QR_CUDA(ctx, cudaEventSynchronize(d0.pipe_ev[ticket]));
            ctx->pending[ticket] = false;
            return QR_OK;
        }
        int rc = qr_render_fetch(ctx, ticket, frame, stride);
        if (rc != QR_OK)
        {
            return rc;
        }
    }
    else
    if (frame != NULL && (frame != ctx->fetch_dst[ticket] || stride != ctx->fetch_stride[ticket]))
    {
        return qr_fail(ctx, QR_E_ARG, "qr_render_end: the frame is being fetched into another buffer");
    }
    bool copied = false;
    if (ctx->fetching[ticket] == 2 && ctx->fetch_posted[ticket])
    {
        ctx->fetch_posted[ticket] = false;
        if (qr_helpers_wait(ctx, ticket) < 0)
        {
            return qr_fail(ctx, QR_E_CUDA, "qr_render_end: frame transfer failed");
        }
        copied = true;
    }
    if (!copied)
    {
        QR_CUDA(ctx, cudaEventSynchronize(d0.fetch_ev[ticket]));
        if (ctx->fetching[ticket] == 2)
        {
            const qr_blob_header &h = ctx->pend_hdr[ticket];
            qr_copy_job j;
            j.dst = ctx->fetch_dst[ticket]; j.src = d0.frame_p[ticket]; j.st = ctx->fetch_stride[ticket];
            j.dstride = h.x_row >= h.x_res ? h.x_row : h.x_res;
            j.x_res = h.x_res; j.ya = 0; j.yb = h.y_res; j.ev = NULL;
            qr_copy_rows(j);
        }
    }
    ctx->pending[ticket] = false;
    ctx->fetching[ticket] = 0;
    return QR_OK;
}

extern "C" int qr_sync(qr_ctx *ctx)
{
    if (ctx == NULL)
    {
        return QR_E_ARG;
    }
    for (int i = ctx->ndev - 1; i >= 0; i--)
    {
        QR_CUDA(ctx, cudaSetDevice(ctx->dev[i].id));
        QR_CUDA(ctx, cudaStreamSynchronize(ctx->dev[i].stream));
    }
    return QR_OK;
}

extern "C" int qr_frame_device(qr_ctx *ctx, const uint32_t **frame_dev, int *stride)
{
    if (ctx == NULL || frame_dev == NULL || stride == NULL)
    {
        return QR_E_ARG;
    }
    if (!ctx->have_scene)
    {
        return qr_fail(ctx, QR_E_STATE, "qr_frame_device: no scene uploaded");
    }
    int rc = qr_frame_ensure(ctx);
    if (rc != QR_OK)
    {
        return rc;
    }
    *frame_dev = ctx->dev[0].frame_d;
    *stride = ctx->hdr.x_row >= ctx->hdr.x_res ? ctx->hdr.x_row : ctx->hdr.x_res;
    return QR_OK;
}

extern "C" int qr_frame_ipc_export(qr_ctx *ctx, void *handle64)
{
    if (ctx == NULL || handle64 == NULL)
    {
        return QR_E_ARG;
    }
    if (!ctx->have_scene)
    {
        return qr_fail(ctx, QR_E_STATE, "qr_frame_ipc_export: no scene uploaded");
    }
    int rc = qr_frame_ensure(ctx);
    if (rc != QR_OK)
    {
        return rc;
    }
    cudaIpcMemHandle_t hnd;
    QR_CUDA(ctx, cudaSetDevice(ctx->dev[0].id));
    QR_CUDA(ctx, cudaIpcGetMemHandle(&hnd, ctx->dev[0].frame_d));
    static_assert(sizeof(hnd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &hnd, sizeof(hnd));
    return QR_OK;
}

extern "C" int qr_frame_ipc_open(qr_ctx *ctx, const void *handle64, uint32_t **frame_dev)
{
    if (ctx == NULL || handle64 == NULL || frame_dev == NULL)
    {
        return QR_E_ARG;
    }
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, handle64, sizeof(hnd));
    void *ptr = NULL;
    QR_CUDA(ctx, cudaSetDevice(ctx->dev[0].id));
    QR_CUDA(ctx, cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
    *frame_dev = (uint32_t *)ptr;
    return QR_OK;
}

extern "C" int qr_frame_ipc_close(qr_ctx *ctx, uint32_t *frame_dev)
{
    if (ctx == NULL || frame_dev == NULL)
    {
        return QR_E_ARG;
    }
    QR_CUDA(ctx, cudaSetDevice(ctx->dev[0].id));
    QR_CUDA(ctx, cudaStreamSynchronize(ctx->dev[0].stream));
    QR_CUDA(ctx, cudaIpcCloseMemHandle(frame_dev));
    return QR_OK;
}

extern "C" int qr_dump_hits(qr_ctx *ctx, float *t_out)
{
    if (ctx == NULL || t_out == NULL)
    {
        return QR_E_ARG;
    }
    if (!ctx->have_scene)
    {
        return qr_fail(ctx, QR_E_STATE, "qr_dump_hits: no scene uploaded");
    }
    if (ctx->pt_scene)
    {
        return qr_fail(ctx, QR_E_STATE, "qr_dump_hits: no dump mode for a path-traced scene");
    }
    const qr_blob_header &h = ctx->hdr;
    const size_t n = ((size_t)h.x_res * h.y_res) << h.fsaa;
    qr_dev &d0 = ctx->dev[0];
    QR_CUDA(ctx, cudaSetDevice(d0.id));
    int rc = qr_grow(ctx, (void **)&d0.t_d, &d0.t_cap, n * sizeof(float), false);
    if (rc != QR_OK)
    {
        return rc;
    }
    rc = qr_render_all(ctx, d0.t_d);
    if (rc != QR_OK)
    {
        return rc;
    }
    QR_CUDA(ctx, cudaMemcpyAsync(t_out, d0.t_d, n * sizeof(float), cudaMemcpyDeviceToHost, d0.stream));
    QR_CUDA(ctx, cudaStreamSynchronize(d0.stream));
    return QR_OK;
}

extern "C" int qr_ray_counts(qr_ctx *ctx, uint64_t counts[4])
{
    if (ctx == NULL || counts == NULL)
    {
        return QR_E_ARG;
    }
    int rc = qr_collect_rays(ctx);
    if (rc != QR_OK)
    {
        return rc;
    }
    for (int k = 0; k < 4; k++)
    {
        counts[k] = ctx->rays[k];
        ctx->rays[k] = 0;
    }
    return QR_OK;
}

extern "C" int qr_last_render_ms(qr_ctx *ctx, float *ms)
{
    if (ctx == NULL || ms == NULL)
    {
        return QR_E_ARG;
    }
    float best = 0.0f;
    for (int i = 0; i < ctx->ndev; i++)
    {
        qr_dev &d = ctx->dev[i];
        if (!d.timed)
        {
            continue;
        }
        float t = 0.0f;
        QR_CUDA(ctx, cudaSetDevice(d.id));
        QR_CUDA(ctx, cudaEventSynchronize(d.ev1));
        QR_CUDA(ctx, cudaEventElapsedTime(&t, d.ev0, d.ev1));
        if (t > best) best = t;
    }
    *ms = best;
    return QR_OK;
}

extern "C" void *qr_stream(qr_ctx *ctx, int index)
{
    if (ctx == NULL || index < 0 || index >= ctx->ndev)
    {
        return NULL;
    }
    return (void *)ctx->dev[index].stream;
}

extern "C" uint64_t qr_launch_count(const qr_ctx *ctx)
{
    return ctx != NULL ? ctx->launches : 0;
}

extern "C" int qr_fp32_peak(qr_ctx *ctx, double *tera_ops)
{
    if (ctx == NULL || tera_ops == NULL)
    {
        return QR_E_ARG;
    }
    qr_dev &d = ctx->dev[0];
    QR_CUDA(ctx, cudaSetDevice(d.id));
    const int iters = 8192, grid = d.sm_count * 8;
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++)
    {
        float ms = 0.0f;
        QR_CUDA(ctx, cudaEventRecord(d.ev0, d.stream));
        qr_fp32_peak_kernel<<<grid, 256, 0, d.stream>>>((float *)d.rays_d, 0.999f, 0.001f, iters);
        QR_CUDA(ctx, cudaGetLastError());
        QR_CUDA(ctx, cudaEventRecord(d.ev1, d.stream));
        QR_CUDA(ctx, cudaEventSynchronize(d.ev1));
        QR_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev0, d.ev1));
        ctx->launches++;
        const double ops = (double)grid * 256.0 * iters * 16.0;
        const double t = ops / (ms * 1e-3) / 1e12;
        if (rep > 0 && t > best) best = t;
    }
    d.timed = false;
    *tera_ops = best;
    return QR_OK;
}

extern "C" int qr_kernel_query(qr_ctx *ctx, qr_kernel_info *info)
{
    if (ctx == NULL || info == NULL)
    {
        return QR_E_ARG;
    }
    info->sm_count = ctx->dev[0].sm_count;
    info->threads_per_cta = g_shapes[ctx->shape].threads;
    info->ctas_per_sm = ctx->dev[0].ctas_per_sm;
    info->regs_per_thread = ctx->fattr.numRegs;
    info->local_bytes_per_thread = (int)ctx->fattr.localSizeBytes;
    info->smem_static_bytes = (int)ctx->fattr.sharedSizeBytes;
    info->smem_dynamic_bytes = (int)ctx->stage_bytes + g_shapes[ctx->shape].threads * QR_SC_QUADS * 16;
    info->scene_in_smem = ctx->stage_bytes != 0;
    info->device_tiling = ctx->have_scene && ctx->device_tiling ? 1 : 0;
    return QR_OK;
}
