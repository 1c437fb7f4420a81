/* artb_tokens.h — token codes of the scene-graph dump stream.
 *
 * artb_scene_dump_graph() (include/artb.h) serialises a host scene graph in depth-first pre-order
 * as a flat array of doubles. The oracle glue (oracle/ref_glue/ref_tool.cpp) produces the same stream
 * from the UNMODIFIED reference's scene graph, so "our host scene construction == the reference's"
 * is a plain array comparison (tests/test_host_scene.py).
 *
 * Grammar (every item is a double; ints are stored exactly):
 *   LIST n  child*n                         reference: src/engine/hittable_list.h:9-24
 *   BVH  min(3) max(3) same  left [right]   reference: src/primitives/bvh.h:9-26 (same=1 when left==right)
 *   TRANSLATE off(3)  child                 reference: src/engine/hittable.h:31-44
 *   ROTATE_Y sin cos hasbox min(3) max(3)  child   reference: src/engine/hittable.h:46-64
 *   BOX min(3) max(3)  LIST(6 rects)        reference: src/primitives/box.h:9-24
 *   SPHERE c(3) r  material                 reference: src/primitives/sphere.h:7-22
 *   MSPHERE c0(3) c1(3) t0 t1 r  material   reference: src/primitives/moving_sphere.h:7-27
 *   XY|XZ|YZ a0 a1 b0 b1 k  material        reference: src/primitives/aarect.h:8-66
 *   TRI p1(3) p2(3) p3(3)  material         reference: src/primitives/triangle.h:7-20
 *   MEDIUM neg_inv_density  material  child(boundary)   reference: src/engine/constant_medium.h:10-35
 * material:
 *   LAMBERTIAN texture | METAL rgb fuzz | DIELECTRIC ir | DIFFUSE_LIGHT texture | ISOTROPIC texture
 * texture:
 *   SOLID rgb | CHECKER even odd | NOISE scale hash_lo hash_hi | IMAGE w h bpp hash_lo hash_hi
 *   | BARY_IMAGE ua va ub vb uc vc IMAGE... | BARY a(3) b(3) c(3)
 * hash = FNV-1a/64 over the raw table bytes, split in two 32-bit halves.
 */
#ifndef ARTB_TOKENS_H
#define ARTB_TOKENS_H

enum artb_graph_token {
    ARTB_T_LIST = 100, ARTB_T_BVH = 101, ARTB_T_TRANSLATE = 102, ARTB_T_ROTATE_Y = 103, ARTB_T_BOX = 104,
    ARTB_T_SPHERE = 110, ARTB_T_MSPHERE = 111, ARTB_T_XY = 112, ARTB_T_XZ = 113, ARTB_T_YZ = 114,
    ARTB_T_TRI = 115, ARTB_T_MEDIUM = 116,
    ARTB_M_LAMBERTIAN = 200, ARTB_M_METAL = 201, ARTB_M_DIELECTRIC = 202, ARTB_M_DIFFUSE_LIGHT = 203,
    ARTB_M_ISOTROPIC = 204,
    ARTB_X_SOLID = 300, ARTB_X_CHECKER = 301, ARTB_X_NOISE = 302, ARTB_X_IMAGE = 303,
    ARTB_X_BARY_IMAGE = 304, ARTB_X_BARY = 305
};

#endif
