/*
 * qr_scene_blob.h -- the flattened, position-independent scene blob.
 *
 * This is the data format that crosses the C ABI (include/quadray_b200.h):
 * one contiguous little-endian buffer of 32-bit words that holds everything
 * the reference's render0 reads through raw pointers from rt_SIMD_INFOX
 * (core/tracer/tracer.h:150-407) -- camera, context level 0, surfaces,
 * materials, lights, texels, list elements and the tile heads -- with every
 * pointer replaced by an index and every SIMD-broadcast field stored once.
 * SURVEY.md appendix B lists which fields render0 consumes; this header is
 * their de-broadcast layout.
 *
 * Producers: quadray-engine_b200/host/qr_flatten.cpp (walks the engine's
 * pointer graph inside rt_Platform::render0).  Consumers: the CUDA kernels
 * (quadray-engine_b200/csrc) and, for tests only, oracle/render0_oracle.c.
 *
 * All indices are int32; QR_NIL (-1) is the NULL pointer.
 */
#ifndef QR_SCENE_BLOB_H
#define QR_SCENE_BLOB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QR_BLOB_MAGIC    0x31425251u   /* "QRB1" */
#define QR_BLOB_VERSION  3u
#define QR_NIL           (-1)

/* context stack depth for secondary rays, core/tracer/tracer.h:46 */
#define QR_STACK_DEPTH   10

/* material property bits, core/tracer/tracer.h:61-72 */
#define QR_PROP_LIGHT    0x00000010
#define QR_PROP_METAL    0x00000020
#define QR_PROP_GAMMA    0x00000040
#define QR_PROP_FRESNEL  0x00000080
#define QR_PROP_NORMAL   0x00000100
#define QR_PROP_OPAQUE   0x00000200
#define QR_PROP_TRANSP   0x00000400
#define QR_PROP_TEXTURE  0x00000800
#define QR_PROP_REFLECT  0x00001000
#define QR_PROP_REFRACT  0x00002000
#define QR_PROP_DIFFUSE  0x00004000
#define QR_PROP_SPECULAR 0x00008000

/* qr_blob_header::flags */
#define QR_BLOB_PT          0x100u

/* context flags, core/tracer/tracer.cpp:504-512 */
#define QR_FLAG_SIDE_OUTER 0
#define QR_FLAG_SIDE_INNER 1
#define QR_FLAG_SIDE       1
#define QR_FLAG_PASS_BACK  0
#define QR_FLAG_PASS_THRU  2
#define QR_FLAG_PASS       2
#define QR_FLAG_SHAD       4

/*
 * Blob header.  Scalars first, then the section table (byte offsets from the
 * start of the blob, all multiples of 16).
 */
typedef struct qr_blob_header
{
    uint32_t magic;
    uint32_t version;
    uint32_t total_bytes;
    uint32_t flags;           /* QR_BLOB_PT: the engine asked for the path tracer (rt_SIMD_INFOX::pt_on) */

    /* rt_SIMD_INFOX externals, tracer.h:154-216 (engine.cpp:842-857) */
    int32_t  x_res;           /* frm_w */
    int32_t  y_res;           /* frm_h */
    int32_t  x_row;           /* frm_row, framebuffer stride in pixels */
    int32_t  fsaa;            /* 0 none, 1 2X, 2 4X (engine.h:52-55) */
    int32_t  depth;           /* inf_DEPTH at entry (RT_STACK_DEPTH) */
    int32_t  tile_w;
    int32_t  tile_h;
    int32_t  tls_row;         /* tiles per row */
    int32_t  tls_col;         /* tile rows */
    int32_t  lst_head;        /* inf_LST (camera list), elem index */
    int32_t  pad0[2];

    /* rt_SIMD_CONTEXT level 0 inputs, engine.cpp:3588-3596 */
    uint32_t ctx_flags;       /* param[1]: RT_PROP_GAMMA or 0 */
    float    t_min;           /* cam->pov */
    float    org[3];
    uint32_t off_bounds;      /* optional section (0 = absent): qr_bound per surface */
    int32_t  n_bounds;        /* = n_surf when present */
    int32_t  pad1[1];

    /* rt_SIMD_CAMERA, engine.cpp:3559-3584 */
    float    cam_t_max;       /* +inf */
    float    dir[3];
    float    hor[3];
    float    ver[3];
    float    hor_a[4];        /* per-lane AA addends, period 4 */
    float    ver_a[4];
    float    amb[3];          /* accumulated ambient colour (cam col_rgb) */
    float    cam_clamp;       /* 255.0 */
    uint32_t cam_cmask;       /* 255 */
    int32_t  pad2[1];

    /* section table */
    int32_t  n_surf;   uint32_t off_surf;
    int32_t  n_mat;    uint32_t off_mat;
    int32_t  n_lgt;    uint32_t off_lgt;
    int32_t  n_elem;   uint32_t off_elem;
    int32_t  n_tiles;  uint32_t off_tiles;
    int32_t  n_texels; uint32_t off_texels;
    int32_t  pad3[4];
} qr_blob_header;

/*
 * Surface record = rt_SIMD_SURFACE (tracer.h:821-969) de-broadcast.
 * 64 words.
 */
typedef struct qr_surface
{
    float    pos[3];          /* pos_x/y/z */
    float    d_eps;           /* root sorting thresholds, object.h:41-42 */
    float    min[3];          /* axis min clippers */
    float    t_eps;
    float    max[3];          /* axis max clippers */
    int32_t  minmax_t;        /* bit a: min_t[a] != 0, bit 3+a: max_t[a] != 0 */
    float    tci[3];          /* transform coeffs, row i */
    int32_t  conic;           /* msc_p[1]: 0 none, 1 cone-like, 2 mask out J */
    float    tcj[3];
    int32_t  trnode;          /* msc_p[3]: trnode's surface index */
    float    tck[3];
    int32_t  clip_head;       /* msc_p[2]: custom clippers list, elem index */
    float    sci[4];          /* geometry scaling coeffs x,y,z,w */
    float    scj[3];
    uint32_t c_def;           /* clipping accum default (all ones) */
    int32_t  a_map[4];        /* I,J,K: field index 0..5 (x,y,z,i,j,k);
                                 L: 0 none, 1 scale, 2 rotate, 3 both */
    int32_t  a_sgn[4];        /* I,J,K: 1 = flip sign; L: 0 or 3 (field shift) */
    int32_t  srf_t[4];        /* solver, material redirect, clipper, obj tag */
    int32_t  mat[2];          /* mat_p[0], mat_p[2]: outer/inner material */
    int32_t  props[2];        /* mat_p[1], mat_p[3]: outer/inner props */
    int32_t  lst_lgt[2];      /* lst_p[0], lst_p[2]: outer/inner light lists */
    int32_t  lst_srf[2];      /* lst_p[1], lst_p[3]: outer/inner rfl/rfr lists */
    int32_t  pad[12];
} qr_surface;

#define QR_SURF_WORDS 64

/* Material record = rt_SIMD_MATERIAL (tracer.h:979-1078).  32 words. */
typedef struct qr_material
{
    float    xscal, yscal, xoffs, yoffs;
    uint32_t xmask, ymask, yshft;
    int32_t  tex;             /* tex_p[0]: first texel, index into texel pool */
    int32_t  t_map[2];        /* 0 = tex_u, 1 = tex_v */
    float    l_dff, l_spc;
    uint32_t l_pow;           /* fixed-point 28.4 */
    float    c_rfl, c_trn, c_rfr, rfr_2, c_rcp, ext_2;
    float    clamp;           /* 255.0 */
    uint32_t cmask;           /* 255 */
    float    col[3];          /* col_r/g/b: self-emission of a light-emitting surface, path tracer
                                 only (object.cpp:1332-1372); zero in blobs written before it existed */
    int32_t  pad[8];
} qr_material;

#define QR_MAT_WORDS 32

/* Light record = rt_SIMD_LIGHT (tracer.h:765-811).  16 words. */
typedef struct qr_light
{
    float    t_max;           /* 1.0 */
    float    pos[3];
    float    col[3];          /* already scaled by lum[1] */
    float    a_qdr, a_lnr, a_cnt;
    int32_t  pad[6];
} qr_light;

#define QR_LGT_WORDS 16

/*
 * List element = rt_ELEM (tracer.h:127-141).  4 words.
 *   surface lists (tiles, inf_LST, lst_p[1/3], shadow lists):
 *       simd = surface index; data_i = elm.data & 3 (0 trnode/plain, 1 bvnode);
 *       data_p = elem index of (elm.data & ~3): the last leaf of the array
 *   light lists (lst_p[0/2]):
 *       simd = light index; data_p = head of the light's shadow surface list
 *   clip lists (msc_p[2]):
 *       simd = clipper surface index or QR_NIL for an accum marker;
 *       data_i = clip side / accum marker (+-1) for non-array clippers;
 *       data_p = trnode's last element when the clipper is an array (tag < 0)
 */
typedef struct qr_elem
{
    int32_t  data_i;
    int32_t  data_p;
    int32_t  simd;
    int32_t  next;
} qr_elem;

#define QR_ELEM_WORDS 4

/*
 * Bounding-box vertices of a surface in world space = rt_BOUND::verts of the
 * element's "temp" (core/engine/rtgeom.h:254-299), what rt_SceneThread::stile
 * projects onto the tile buffer (core/engine/engine.cpp:1956-2128).  Present
 * when the engine left the tiling to the backend (RT_OPTS_TILING off: every
 * tile head is the camera list): the device then culls the list per tile
 * itself.  n = 0: unbounded (covers every tile, engine.cpp:2097-2105);
 * n = -1: surface not in the camera list.
 */
typedef struct qr_bound
{
    int32_t  n;
    float    v[8][3];
    int32_t  pad[3];
} qr_bound;

#define QR_BOUND_WORDS 28

#ifdef __cplusplus
}
#endif

#ifdef __cplusplus
static_assert(sizeof(qr_blob_header) == 256, "qr_blob_header must be 64 words");
static_assert(sizeof(qr_surface)  == QR_SURF_WORDS * 4, "qr_surface size");
static_assert(sizeof(qr_material) == QR_MAT_WORDS  * 4, "qr_material size");
static_assert(sizeof(qr_light)    == QR_LGT_WORDS  * 4, "qr_light size");
static_assert(sizeof(qr_elem)     == QR_ELEM_WORDS * 4, "qr_elem size");
static_assert(sizeof(qr_bound)    == QR_BOUND_WORDS * 4, "qr_bound size");
#else
_Static_assert(sizeof(qr_blob_header) == 256, "qr_blob_header must be 64 words");
_Static_assert(sizeof(qr_surface)  == QR_SURF_WORDS * 4, "qr_surface size");
_Static_assert(sizeof(qr_material) == QR_MAT_WORDS  * 4, "qr_material size");
_Static_assert(sizeof(qr_light)    == QR_LGT_WORDS  * 4, "qr_light size");
_Static_assert(sizeof(qr_elem)     == QR_ELEM_WORDS * 4, "qr_elem size");
_Static_assert(sizeof(qr_bound)    == QR_BOUND_WORDS * 4, "qr_bound size");
#endif

#endif /* QR_SCENE_BLOB_H */
