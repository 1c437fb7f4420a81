// flat_scene.h — the flattened scene as it lives in HBM. Plain-old-data shared by the host flattener
// (g++) and the sm_100a kernels (nvcc). See DESIGN.md "Data layout in HBM". The record types themselves are PUBLIC:
// include/artb_flat.h defines them (a foreign host fills them and calls artb_upload_flat); this header gives them the
// names the library's own code uses and adds the device-side views.
//
// The reference's pointer graph (hittable_list / bvh_node / translate / rotate_y / box / primitives /
// constant_medium, shared_ptr<material>, shared_ptr<texture>) becomes:
//   prims[]      one fixed-size record per leaf primitive: first the few "global" primitives whose box spans a
//                large part of the scene (tested for every ray before the traversal: no divergence, and an
//                early t_max), then the BVH leaves in leaf order, then the media boundaries
//   nodes[]      one 8-wide BVH over ALL world primitives in WORLD space (instanced primitives keep their
//                object-space geometry and a transform-chain id; their world boxes bound the transformed
//                geometry). The closest hit does not depend on the acceleration structure, so this
//                single tree replaces the reference's list-of-BVHs-of-instances and is bit-compatible
//                with it as long as the tie-break (dfs_id) and culling rules below are kept.
//   xforms[]/xform_ops[]  translate / rotate_y chains, outermost first
//   materials[] / textures[] / images[] + texel atlas / perlins[]
#ifndef ARTB_FLAT_SCENE_H
#define ARTB_FLAT_SCENE_H

#include <stdint.h>

#include "artb_flat.h"       // include/: the record definitions are public (a foreign host fills them, artb_upload_flat)

#if defined(__CUDACC__)
#define ARTB_ALIGN(n) __align__(n)
#else
#define ARTB_ALIGN(n) alignas(n)
#endif

namespace artb {

// the public records under the names the library's C++ / CUDA code uses
using FlatPrim = ::artb_flat_prim;
using FlatNode8 = ::artb_flat_node8;
using FlatRefNode = ::artb_flat_ref_node;
using FlatXformOp = ::artb_flat_xform_op;
using FlatXform = ::artb_flat_xform;
using FlatMaterial = ::artb_flat_material;
using FlatTexture = ::artb_flat_texture;
using FlatImage = ::artb_flat_image;
using FlatPerlin = ::artb_flat_perlin;

enum : uint32_t {
    PRIM_SPHERE = ARTB_PRIM_SPHERE, PRIM_MSPHERE = ARTB_PRIM_MSPHERE, PRIM_RECT_XY = ARTB_PRIM_RECT_XY,
    PRIM_RECT_XZ = ARTB_PRIM_RECT_XZ, PRIM_RECT_YZ = ARTB_PRIM_RECT_YZ, PRIM_TRIANGLE = ARTB_PRIM_TRIANGLE,
    PRIM_MEDIUM = ARTB_PRIM_MEDIUM
};
enum : uint32_t { XFORM_TRANSLATE = ARTB_XFORM_TRANSLATE, XFORM_ROTATE_Y = ARTB_XFORM_ROTATE_Y };
enum : uint32_t { MAT_LAMBERTIAN = ARTB_MAT_LAMBERTIAN, MAT_METAL = ARTB_MAT_METAL, MAT_DIELECTRIC = ARTB_MAT_DIELECTRIC,
                  MAT_DIFFUSE_LIGHT = ARTB_MAT_DIFFUSE_LIGHT, MAT_ISOTROPIC = ARTB_MAT_ISOTROPIC, MAT_COUNT = ARTB_MAT_COUNT };
enum : uint32_t { TEX_SOLID = ARTB_TEX_SOLID, TEX_CHECKER = ARTB_TEX_CHECKER, TEX_NOISE = ARTB_TEX_NOISE, TEX_IMAGE = ARTB_TEX_IMAGE,
                  TEX_BARY_IMAGE = ARTB_TEX_BARY_IMAGE, TEX_BARY = ARTB_TEX_BARY };

// device-side view (all pointers are device pointers)
// Code path a primitive takes in the double-precision test: k_pairs sorts each tile of pairs by it so that the lanes of
// a warp run the same test (sphere / moving sphere / rect / triangle, each without or with a transform chain).
#define ARTB_PRIM_CLASSES 8
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint32_t prim_class_of(uint32_t type, uint32_t xform)
{
    const uint32_t path = type == PRIM_SPHERE ? 0u : type == PRIM_MSPHERE ? 1u : type == PRIM_TRIANGLE ? 3u : 2u;
    return path * 2u + (xform ? 1u : 0u);
}

struct SceneView {
    const FlatPrim* prims;
    const FlatRefNode* ref_nodes;     // reference bvh_nodes for exact tie-breaking, or nullptr (ties then go to the larger dfs_id)
    const uint8_t* prim_class;        // per primitive: prim_class_of(type, xform) (device only; built by Renderer::upload)
    const uint8_t* prim_mat_type;     // per primitive: type of its material = the shade queue a hit goes to (device only)
    const FlatNode8* nodes;
    const FlatXformOp* xform_ops;
    const FlatXform* xforms;
    const FlatMaterial* materials;
    const FlatTexture* textures;
    const FlatImage* images;
    const unsigned char* texels;
    const FlatPerlin* perlins;
    uint32_t n_prims, n_world_prims, n_global_prims, n_nodes, n_media, n_xforms, n_materials, n_textures;
    uint32_t n_global_media;          // media among the global primitives (tested after the traversal)
    uint32_t one_bits;                // 0x3f800000 as a RUNTIME value: the node test wants it in a register next to immediate
                                      // PRMT selectors; as a literal ptxas re-materialises it before every PRMT (rt_intersect.cuh)
    double background[3];
    // This view in memory the out-of-line device functions can point at (tie_replaces, medium_interval_general,
    // replay_ties): the device copy in global memory (Renderer::sync_view), the object itself on the host. They take this
    // pointer and their other arguments by value — a C++ reference to a kernel parameter would make every thread of the
    // kernel copy the whole parameter block into local memory at start-up (a 256-480 byte stack frame per thread).
    const SceneView* self;
};

struct CameraParams {     // camera.h:49-56 + engine.h:62-63
    double origin[3], lower_left_corner[3], horizontal[3], vertical[3], u[3], v[3], w[3];
    double lens_radius, time0, time1;
};

// closest-hit record reported by artb_trace_batch and consumed by the shade kernels
struct ARTB_ALIGN(8) HitRecordOut {
    double t, p[3], normal[3], u, v;
    int32_t hit, prim_id, front_face, material_type;
};

// what artb_shade_batch reports per (ray, hit): material::emitted + material::scatter of the hit (material.h:10-135)
struct ARTB_ALIGN(8) ShadeRecordOut {
    double emitted[3], attenuation[3];
    double origin[3], direction[3], time;     // the scattered ray (valid when scattered == 1)
    int32_t scattered;                        // scatter()'s return value; -1: the hit id does not name a primitive
    int32_t material_type, front_face, pad;
};

} // namespace artb

#endif
