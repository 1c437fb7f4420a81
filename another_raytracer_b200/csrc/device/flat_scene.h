// flat_scene.h — the flattened scene as it lives in HBM. Plain-old-data shared by the host flattener
// (g++) and the sm_100a kernels (nvcc). See DESIGN.md "Data layout in HBM".
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

#if defined(__CUDACC__)
#define ARTB_ALIGN(n) __align__(n)
#else
#define ARTB_ALIGN(n) alignas(n)
#endif

namespace artb {

enum : uint32_t {
    PRIM_SPHERE = 0,    // g = cx cy cz r                      sphere.h:39-65
    PRIM_MSPHERE = 1,   // g = c0(3) c1(3) t0 t1 r             moving_sphere.h:29-58
    PRIM_RECT_XY = 2,   // g = x0 x1 y0 y1 k                   aarect.cpp:3-19
    PRIM_RECT_XZ = 3,   // g = x0 x1 z0 z1 k                   aarect.cpp:21-37
    PRIM_RECT_YZ = 4,   // g = y0 y1 z0 z1 k                   aarect.cpp:39-55
    PRIM_TRIANGLE = 5,  // g = p1(3) p2(3) p3(3)               triangle.h:22-88
    PRIM_MEDIUM = 6     // g = neg_inv_density, first boundary prim, boundary prim count, medium slot
                        //                                      constant_medium.h:37-82
};

struct ARTB_ALIGN(16) FlatPrim {      // 96 B
    double g[10];         // every type but PRIM_MEDIUM: g[9] = 1 + index of the innermost reference bvh_node that encloses
                          // the primitive (FlatRefNode, below), 0 = none
    uint32_t type;
    uint32_t material;
    uint32_t xform;       // 0 = identity
    uint32_t dfs_id;      // position in the reference's depth-first test order: reported id + tie-break
};

// 8-wide BVH node with boxes quantised to 8 bits per plane on a per-node grid (the compressed wide BVH of Ylitie,
// Karras & Laine, HPG 2017): 80 bytes = five 16-byte loads decide up to eight descents, and the whole tree of the
// reference's scenes (a few hundred nodes) fits in one SM's shared memory, where k_extend stages it.
//   child box plane (axis a)  =  p[a] + q * 2^(e[a]-127),  q in 0..255; lo planes rounded DOWN, hi planes UP, so the
//   quantised box CONTAINS the double-precision world box. The node test only has to be conservative (it may admit
//   a ray that misses, never reject one that hits): every candidate primitive is re-tested with the reference's
//   double-precision formula, so results do not depend on the tree at all — only the amount of work does.
//   slot s (0..7): meta[s] = 0                          empty
//                  meta[s] = 0b001'00000 | (24 + s)     internal child; its node index is child_base + (number of
//                                                       internal slots below s), imask has bit s set
//                  meta[s] = unary(count) << 5 | offset  leaf of `count` (1..3) primitives prims[prim_base + offset ..]
//   Children are placed in slots so that a ray of octant o (bit 2: dx<0, bit 1: dy<0, bit 0: dz<0) meets them roughly
//   front to back when it visits slots in the order of decreasing (slot XOR (7-o)).
struct ARTB_ALIGN(16) FlatNode8 {     // 80 B
    float p[3];
    uint8_t e[3];
    uint8_t imask;
    uint32_t child_base;
    uint32_t prim_base;
    uint8_t meta[8];
    uint8_t qlo[3][8];
    uint8_t qhi[3][8];
};
#define ARTB_MAX_LEAF_PRIMS 3

// One bvh_node of the REFERENCE's own trees (bvh.h:22-25), kept only to break exact ties the way the reference does:
// bvh_node::hit tests its box with the current closest t as t_max, and aabb::hit rejects t_max <= t_min (aabb.h:25,
// strict), so a later leaf whose enclosing boxes are entered exactly AT the closest t so far is never visited and
// cannot overwrite an equal-t hit (rt_intersect.cuh: tie_later_wins). Coincident faces of adjacent boxes do this.
struct ARTB_ALIGN(16) FlatRefNode {   // 64 B
    double lo[3], hi[3];  // bvh_node::box, in the space of its transform chain
    int32_t parent;       // enclosing reference bvh_node (through lists and wrappers), -1 = none
    uint32_t xform;       // transform chain in force at the node
    int32_t depth;        // number of enclosing reference bvh_nodes
    uint32_t pad;
};

enum : uint32_t { XFORM_TRANSLATE = 0, XFORM_ROTATE_Y = 1 };

struct ARTB_ALIGN(16) FlatXformOp {   // 32 B
    double a, b, c;       // translate: offset xyz          hittable.cpp:3-12
                          // rotate_y:  sin_theta, cos_theta hittable.cpp:57-85
    uint32_t type;
    uint32_t pad;
};

struct FlatXform { uint32_t first_op, n_ops; };   // ops are applied first_op .. first_op+n_ops-1 going IN

enum : uint32_t { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_DIFFUSE_LIGHT = 3, MAT_ISOTROPIC = 4,
                  MAT_COUNT = 5 };

struct ARTB_ALIGN(16) FlatMaterial {  // 48 B
    double albedo[3];     // metal only
    double param;         // metal: fuzz; dielectric: ir
    uint32_t type;
    uint32_t texture;     // lambertian / diffuse_light / isotropic
    uint32_t pad[2];
};

enum : uint32_t { TEX_SOLID = 0, TEX_CHECKER = 1, TEX_NOISE = 2, TEX_IMAGE = 3, TEX_BARY_IMAGE = 4, TEX_BARY = 5 };

struct ARTB_ALIGN(16) FlatTexture {   // 96 B
    double p[9];          // solid: rgb | noise: scale | bary_image: ua va ub vb uc vc | bary: a(3) b(3) c(3)
    uint32_t type;
    uint32_t a, b;        // checker: even, odd texture | noise: perlin index | image / bary_image: image index
    uint32_t pad[3];
};

struct FlatImage { uint64_t offset; int32_t width, height, bytes_per_pixel, pad; };   // width==0: "no data" => cyan

struct FlatPerlin {       // perlin.h:56-61
    double ranvec[256][3];
    int32_t perm_x[256], perm_y[256], perm_z[256];
};

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

} // namespace artb

#endif
