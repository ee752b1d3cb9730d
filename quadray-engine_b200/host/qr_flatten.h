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
#include <unordered_map>

#include "qr_scene_blob.h"

struct rt_SIMD_INFOX;
struct rt_SIMD_SURFACE;
struct rt_SIMD_MATERIAL;
struct rt_SIMD_LIGHT;
struct rt_ELEM;

class qr_Flattener
{
    public:

    /* Build the blob for one frame; the returned buffer stays valid until
     * the next call.  Throws rt_Exception on malformed input. */
    const uint8_t  *build(const rt_SIMD_INFOX *s_inf, size_t *bytes);

    private:

    enum ListKind { LIST_SURF = 0, LIST_LIGHT = 1, LIST_CLIP = 2 };

    struct PendingList { const rt_ELEM *head; ListKind kind; };

    int32_t     list_head(const rt_ELEM *head, ListKind kind);
    int32_t     surface(const rt_SIMD_SURFACE *s);
    int32_t     material(const rt_SIMD_MATERIAL *m);
    int32_t     light(const rt_SIMD_LIGHT *l);
    void        drain();

    std::unordered_map<const void *, int32_t> elem_idx, surf_idx, mat_idx,
                                              lgt_idx, tex_idx;
    std::vector<const rt_ELEM *>            elem_src;
    std::vector<ListKind>                   elem_kind;
    std::vector<const rt_SIMD_SURFACE *>    surf_src;
    std::vector<PendingList>                pending;

    std::vector<qr_elem>        elems;
    std::vector<qr_surface>     surfs;
    std::vector<qr_material>    mats;
    std::vector<qr_light>       lgts;
    std::vector<uint32_t>       texels;
    std::vector<int32_t>        tiles;
    std::vector<uint8_t>        blob;
    size_t                      surf_done;
};

#endif /* QR_FLATTEN_H */
