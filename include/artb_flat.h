/* artb_flat.h — the flattened scene, record by record: what `artb_flat_scene` (artb.h) points at.
 *
 * A host that owns its own scene representation (the reference's pointer graph, a DCC exporter, a test) fills these
 * plain-old-data arrays and hands them to artb_upload_flat(). They are exactly what the sm_100a kernels read from HBM
 * (the library's C++ side aliases them: csrc/device/flat_scene.h), so sizes and offsets are part of the ABI and are
 * static_assert'ed below. A reference-side host does not have to write any of this by hand: include/
 * artb_reference_adapter.hpp serialises the reference's graph and artb_scene_from_graph() + artb_scene_get_flat()
 * produce the arrays.
 *
 * The reference's pointer graph (hittable_list / bvh_node / translate / rotate_y / box / primitives / constant_medium,
 * shared_ptr<material>, shared_ptr<texture>) becomes:
 *   prims[]     one record per leaf primitive: first the few "global" primitives whose box spans a large part of the scene
 *               (tested for every ray; ALL constant_media are global and come last among them), then the BVH leaves in
 *               leaf order [n_global_prims, n_world_prims), then the boundary primitives of the media
 *   nodes[]     one 8-wide BVH over the non-global world primitives in WORLD space — optional: with n_nodes == 0 the
 *               library builds the tree on the GPU (bvh_build.cuh) and permutes the leaves itself
 *   xforms[] / xform_ops[]   translate / rotate_y chains, outermost first; xform 0 = identity
 *   materials[] / textures[] / images[] + texel atlas / perlins[]
 *   ref_nodes[] the reference's own bvh_node boxes, only used to break exact-t ties the way bvh_node::hit does (optional)
 */
#ifndef ARTB_FLAT_H
#define ARTB_FLAT_H

#include <stdint.h>

#if defined(__cplusplus)
#define ARTB_FLAT_ALIGN16 alignas(16)
#else
#define ARTB_FLAT_ALIGN16 _Alignas(16)
#endif

enum artb_prim_type {
    ARTB_PRIM_SPHERE = 0,    /* g = cx cy cz r                      reference: src/primitives/sphere.h:39-65 */
    ARTB_PRIM_MSPHERE = 1,   /* g = c0(3) c1(3) t0 t1 r             moving_sphere.h:29-58 */
    ARTB_PRIM_RECT_XY = 2,   /* g = x0 x1 y0 y1 k                   aarect.cpp:3-19 */
    ARTB_PRIM_RECT_XZ = 3,   /* g = x0 x1 z0 z1 k                   aarect.cpp:21-37 */
    ARTB_PRIM_RECT_YZ = 4,   /* g = y0 y1 z0 z1 k                   aarect.cpp:39-55 */
    ARTB_PRIM_TRIANGLE = 5,  /* g = p1(3) p2(3) p3(3)               triangle.h:22-88 */
    ARTB_PRIM_MEDIUM = 6     /* g = neg_inv_density, first boundary prim, boundary prim count, medium slot,
                                world box lo(3) hi(3)                constant_medium.h:37-82 */
};

typedef struct artb_flat_prim {       /* 96 B */
    ARTB_FLAT_ALIGN16 double g[10];   /* every type but MEDIUM: g[9] = 1 + index of the innermost reference bvh_node that encloses
                                         the primitive (artb_flat_ref_node), 0 = none */
    uint32_t type;                    /* artb_prim_type */
    uint32_t material;
    uint32_t xform;                   /* 0 = identity */
    uint32_t dfs_id;                  /* position in the reference's depth-first test order: the id artb_trace_batch reports,
                                         and the tie-break (hittable_list.cpp:11-15: the leaf tested last wins an equal t) */
} artb_flat_prim;

/* 8-wide BVH node, boxes quantised to 8 bits per plane on a per-node power-of-two grid (Ylitie, Karras & Laine, HPG 2017):
 *   child box plane (axis a) = p[a] + q * 2^(e[a]-127), q in 0..255; lo planes rounded DOWN, hi planes UP, so the quantised
 *   box CONTAINS the double-precision world box (the node test only has to be conservative: every candidate primitive is
 *   re-tested with the reference's double-precision formula).
 *   slot s (0..7): meta[s] = 0                          empty
 *                  meta[s] = 0b001'00000 | (24 + s)     internal child; node index = child_base + (internal slots below s),
 *                                                       imask has bit s set; child_base must be > the node's own index
 *                  meta[s] = unary(count) << 5 | offset  leaf of `count` (1..3) primitives prims[prim_base + offset ..]       */
typedef struct artb_flat_node8 {      /* 80 B */
    ARTB_FLAT_ALIGN16 float p[3];
    uint8_t e[3];
    uint8_t imask;
    uint32_t child_base;
    uint32_t prim_base;
    uint8_t meta[8];
    uint8_t qlo[3][8];
    uint8_t qhi[3][8];
} artb_flat_node8;
#define ARTB_MAX_LEAF_PRIMS 3
#define ARTB_MAX_BVH8_DEPTH 14          /* levels of 8-wide nodes the device traversal stack is sized for */

/* One bvh_node of the REFERENCE's own trees (bvh.h:22-25): bvh_node::hit tests its box with the current closest t as t_max
 * and aabb::hit rejects t_max <= t_min (aabb.h:25), so a later leaf whose enclosing boxes are entered exactly AT the closest
 * t so far is never visited and cannot overwrite an equal-t hit. Kept only to replay that. */
typedef struct artb_flat_ref_node {   /* 64 B */
    ARTB_FLAT_ALIGN16 double lo[3];
    double hi[3];                     /* bvh_node::box, in the space of its transform chain */
    int32_t parent;                   /* enclosing reference bvh_node (through lists and wrappers), -1 = none; < own index */
    uint32_t xform;                   /* transform chain in force at the node */
    int32_t depth;                    /* number of enclosing reference bvh_nodes */
    uint32_t pad;
} artb_flat_ref_node;

enum artb_xform_type { ARTB_XFORM_TRANSLATE = 0, ARTB_XFORM_ROTATE_Y = 1 };

typedef struct artb_flat_xform_op {   /* 32 B */
    ARTB_FLAT_ALIGN16 double a;       /* translate: offset xyz           hittable.cpp:3-12  */
    double b, c;                      /* rotate_y:  a = sin, b = cos     hittable.cpp:57-85 */
    uint32_t type;                    /* artb_xform_type */
    uint32_t pad;
} artb_flat_xform_op;

typedef struct artb_flat_xform { uint32_t first_op, n_ops; } artb_flat_xform;   /* ops first_op .. first_op+n_ops-1 going IN; n_ops <= 8 */

enum artb_material_type { ARTB_MAT_LAMBERTIAN = 0, ARTB_MAT_METAL = 1, ARTB_MAT_DIELECTRIC = 2, ARTB_MAT_DIFFUSE_LIGHT = 3,
                          ARTB_MAT_ISOTROPIC = 4, ARTB_MAT_COUNT = 5 };

typedef struct artb_flat_material {   /* 48 B */
    ARTB_FLAT_ALIGN16 double albedo[3];   /* metal only */
    double param;                     /* metal: fuzz; dielectric: ir */
    uint32_t type;                    /* artb_material_type */
    uint32_t texture;                 /* lambertian / diffuse_light / isotropic */
    uint32_t pad[2];
} artb_flat_material;

enum artb_texture_type { ARTB_TEX_SOLID = 0, ARTB_TEX_CHECKER = 1, ARTB_TEX_NOISE = 2, ARTB_TEX_IMAGE = 3, ARTB_TEX_BARY_IMAGE = 4,
                         ARTB_TEX_BARY = 5 };

typedef struct artb_flat_texture {    /* 96 B */
    ARTB_FLAT_ALIGN16 double p[9];    /* solid: rgb | noise: scale | bary_image: ua va ub vb uc vc | bary: a(3) b(3) c(3) */
    uint32_t type;                    /* artb_texture_type */
    uint32_t a, b;                    /* checker: even, odd texture (both > own index is NOT required, cycles are rejected) |
                                         noise: perlin index | image / bary_image: image index */
    uint32_t pad[3];
} artb_flat_texture;

typedef struct artb_flat_image { uint64_t offset; int32_t width, height, bytes_per_pixel, pad; } artb_flat_image;   /* width==0: cyan */

typedef struct artb_flat_perlin {     /* perlin.h:56-61 */
    double ranvec[256][3];
    int32_t perm_x[256], perm_y[256], perm_z[256];
} artb_flat_perlin;

#if defined(__cplusplus)
static_assert(sizeof(artb_flat_prim) == 96 && sizeof(artb_flat_node8) == 80 && sizeof(artb_flat_ref_node) == 64 &&
              sizeof(artb_flat_xform_op) == 32 && sizeof(artb_flat_xform) == 8 && sizeof(artb_flat_material) == 48 &&
              sizeof(artb_flat_texture) == 96 && sizeof(artb_flat_image) == 24 && sizeof(artb_flat_perlin) == 9216,
              "flat record sizes are part of the ABI");
#else
_Static_assert(sizeof(artb_flat_prim) == 96 && sizeof(artb_flat_node8) == 80 && sizeof(artb_flat_ref_node) == 64 &&
               sizeof(artb_flat_xform_op) == 32 && sizeof(artb_flat_xform) == 8 && sizeof(artb_flat_material) == 48 &&
               sizeof(artb_flat_texture) == 96 && sizeof(artb_flat_image) == 24 && sizeof(artb_flat_perlin) == 9216,
               "flat record sizes are part of the ABI");
#endif

#endif
