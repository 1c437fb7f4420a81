// Host-side BVH flattener: pointer-based scene graph -> linear arrays for the device (flat_scene.h).
//
// What it preserves from the reference, so that closest-hit results stay bit-compatible:
//  * dfs_id — the order in which the reference's recursive traversal (hittable_list.cpp:5-19 in insertion
//    order, bvh.cpp:44-52 left then right, box.cpp:17-19 six sides, wrappers hittable.cpp:3-12/57-85) tests
//    the leaves. Every reference leaf accepts t == t_max, so among equal-t candidates the one tested LAST
//    wins; the device applies "larger dfs_id wins ties" under its own traversal order.
//  * culling — a reference bvh_node whose box has zero thickness on some axis can never be entered
//    (aabb.h:25 rejects t_max <= t_min), so every leaf below it is unreachable; such leaves keep their
//    dfs_id but are dropped from the device arrays.
//  * transform chains — translate / rotate_y wrappers are recorded per primitive (outermost first) and
//    replayed on the device with the reference's exact arithmetic, including the face-normal re-orientation
//    each wrapper performs on the way out.
//  * media — constant_medium becomes one primitive whose boundary sub-graph is flattened into a private
//    primitive range (scanned linearly, in dfs order, exactly like a hittable_list).
// What it replaces: the reference's trees. One binned-SAH binary BVH is built over the WORLD-space boxes
// of all reachable world primitives.
#ifndef ARTB_HOST_FLATTEN_H
#define ARTB_HOST_FLATTEN_H

#include "scene.h"
#include "../device/flat_scene.h"

#include <vector>

namespace artb {

struct FlatScene {
    std::vector<FlatPrim> prims;          // [0,n_global) global prims, [n_global,n_world) BVH leaf order, then boundary prims
    std::vector<FlatNode8> nodes;        // [0] = root (present iff n_world_prims > n_global_prims)
    std::vector<FlatXformOp> xform_ops;
    std::vector<FlatXform> xforms;        // [0] = identity
    std::vector<FlatMaterial> materials;
    std::vector<FlatTexture> textures;
    std::vector<FlatImage> images;
    std::vector<unsigned char> texels;
    std::vector<FlatPerlin> perlins;
    std::vector<FlatRefNode> ref_nodes;   // the reference's bvh_nodes (tie-breaking only)
    uint32_t n_world_prims = 0;
    uint32_t n_global_prims = 0;
    uint32_t n_global_media = 0;          // the last n_global_media globals are the scene's constant_media
    uint32_t n_media = 0;
    uint32_t n_leaves_total = 0;          // dfs ids handed out (reachable or not)
    uint32_t n_culled = 0;
    double background[3] = {0, 0, 0};
    double world_lo[3] = {0, 0, 0}, world_hi[3] = {0, 0, 0};
    int bvh_depth = 0;                    // levels of 8-wide nodes
    int device_bvh_leaf_prims = 1;        // leaf size for the device builder when `nodes` is empty (FlattenOptions::device_bvh)
};

struct FlattenOptions {
    int max_leaf_prims = 1;               // <= ARTB_MAX_LEAF_PRIMS; 1: every primitive's own box is tested before its exact test
    double global_area_fraction = 0.25;   // prims whose box half-area exceeds this share of the scene's go to the global list
    int max_global_prims = 8;
    int sah_bins = 16;
    double traversal_cost = 1.0;
    double intersect_cost = 1.5;
    bool device_bvh = false;              // skip the host SAH build: leave `nodes` empty, Renderer::upload builds the tree on the GPU
};

FlatScene flatten(const hittable_list& world, const color& background, const FlattenOptions& opt = FlattenOptions());

} // namespace artb

#endif
