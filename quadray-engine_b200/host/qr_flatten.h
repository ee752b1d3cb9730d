/*
 * qr_flatten.h -- host-side flattener: engine pointer graph -> scene blob.
 *
 * Sits inside the replacement rt_Platform::render0 (tracer_b200.cpp).  The
 * engine hands render0 an rt_SIMD_INFOX whose ctx/cam/tiles pointers lead to
 * every rt_ELEM, rt_SIMD_SURFACE, rt_SIMD_MATERIAL, rt_SIMD_LIGHT and texture
 * the frame needs (core/engine/engine.cpp:3600-3627); the lists live in a
 * per-frame pool that is released right after render (engine.cpp:3317-3323),
 * so the graph has to be copied out during the call.  The result is the
 * index-based blob of include/qr_scene_blob.h.
 */
#ifndef QR_FLATTEN_H
#define QR_FLATTEN_H

#include <stdint.h>
#include <stddef.h>
#include <vector>

#include "qr_scene_blob.h"

struct rt_SIMD_INFOX;
struct rt_SIMD_SURFACE;
struct rt_SIMD_MATERIAL;
struct rt_SIMD_LIGHT;
struct rt_ELEM;

/*
 * Pointer -> index map for one frame: open addressing, no allocation once
 * warm, cleared in O(1) by bumping a generation stamp (the flattener runs
 * every frame inside render0, on ~20 k list elements for the demo scenes).
 */
class qr_PtrMap
{
    public:

    qr_PtrMap() : gen(1), used(0) {}

    void clear()
    {
        gen++;
        used = 0;
        if (gen == 0)                           /* stamp wrapped: really clear */
        {
            stamp.assign(stamp.size(), 0);
            gen = 1;
        }
    }

    /* index stored for "p", or -1 */
    int32_t find(const void *p) const
    {
        if (keys.empty()) return -1;
        const size_t mask = keys.size() - 1;
        for (size_t i = hash(p) & mask; ; i = (i + 1) & mask)
        {
            if (stamp[i] != gen) return -1;
            if (keys[i] == p) return vals[i];
        }
    }

    /* store "v" for "p" unless present; returns the stored index */
    int32_t insert(const void *p, int32_t v)
    {
        if ((used + 1) * 2 > keys.size()) grow();
        const size_t mask = keys.size() - 1;
        for (size_t i = hash(p) & mask; ; i = (i + 1) & mask)
        {
            if (stamp[i] != gen)
            {
                stamp[i] = gen; keys[i] = p; vals[i] = v;
                used++;
                return v;
            }
            if (keys[i] == p) return vals[i];
        }
    }

    private:

    static size_t hash(const void *p)
    {
        /* the engine bump-allocates its records, so neighbours in memory are
         * neighbours in a list: keep runs of 16 records (512 B) adjacent in the
         * table, and scatter the runs with a multiplicative hash -- heap chunks
         * differ in high address bits only and would otherwise pile up on the
         * same slots (seen with the 100 k-quadric scenes: 20 M elements) */
        const uintptr_t u = (uintptr_t)p >> 5;
        const uint64_t run = (uint64_t)(u >> 4) * 0x9E3779B97F4A7C15ull;
        return (size_t)(((run >> 24) << 4) | (u & 15));
    }

    void grow()
    {
        std::vector<const void *> ok; std::vector<int32_t> ov; std::vector<uint32_t> os;
        ok.swap(keys); ov.swap(vals); os.swap(stamp);
        const size_t n = ok.empty() ? 1024 : ok.size() * 2;
        keys.assign(n, (const void *)0); vals.assign(n, 0); stamp.assign(n, 0);
        const uint32_t g = gen;
        used = 0;
        for (size_t i = 0; i < ok.size(); i++)
        {
            if (os[i] == g) insert(ok[i], ov[i]);
        }
    }

    std::vector<const void *> keys;
    std::vector<int32_t>      vals;
    std::vector<uint32_t>     stamp;
    uint32_t                  gen;
    size_t                    used;
};

class qr_Flattener
{
    public:

    /* Build the blob for one frame; the returned buffer stays valid until
     * the next call.  Throws rt_Exception on malformed input. */
    const uint8_t  *build(const rt_SIMD_INFOX *s_inf, size_t *bytes);

    private:

    enum ListKind { LIST_SURF = 0, LIST_LIGHT = 1, LIST_CLIP = 2 };

    int32_t     list_head(const rt_ELEM *head, ListKind kind);
    int32_t     surface(const rt_SIMD_SURFACE *s);
    int32_t     material(const rt_SIMD_MATERIAL *m);
    int32_t     light(const rt_SIMD_LIGHT *l);
    void        drain();

    qr_PtrMap   elem_idx, surf_idx, mat_idx, lgt_idx, tex_idx;
    std::vector<const rt_ELEM *>            elem_src;
    std::vector<ListKind>                   elem_kind;
    std::vector<const rt_SIMD_SURFACE *>    surf_src;
    std::vector<uint8_t>                    surf_array;     /* srf_t[3] < 0 per indexed surface */

    std::vector<qr_elem>        elems;
    std::vector<qr_surface>     surfs;
    std::vector<qr_material>    mats;
    std::vector<qr_light>       lgts;
    std::vector<uint32_t>       texels;
    std::vector<int32_t>        tiles;
    std::vector<qr_bound>       bounds;         /* per surface, when the engine left the tiling to us */
    std::vector<uint8_t>        blob;
    size_t                      surf_done;
};

#endif /* QR_FLATTEN_H */
