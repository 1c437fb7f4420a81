// rt_intersect.cuh — closest-hit query over the flattened scene (the "extend" stage).
//
// Replaces, in one iterative traversal of an 8-wide compressed BVH: hittable_list::hit (hittable_list.cpp:5-19),
// bvh_node::hit + aabb::hit (bvh.cpp:44-52, aabb.h:16-29), box::hit (box.cpp:17-19), translate::hit and
// rotate_y::hit (hittable.cpp:3-12, 57-85), sphere / moving_sphere / x?_rect / triangle ::hit and
// constant_medium::hit (constant_medium.h:37-82).
//
// Equivalence argument (DESIGN.md "Closest hit"): every reference leaf computes a candidate t that does
// not depend on t_max and accepts it iff t_min <= t <= t_max (inclusive), and containers only narrow
// t_max to the best t so far. Hence the reference returns the minimum-t candidate, ties going to the leaf
// tested last in its depth-first order. Here leaves may be visited in any order: a candidate replaces the
// current best iff t < best_t or (t == best_t and dfs_id > best_dfs). Primitive formulas are the
// reference's, evaluated in double in the primitive's object space (transform chains replayed with the
// reference's arithmetic), so t is the same number up to FMA contraction. Node tests are FP32 and only
// conservative: they decide how much work is done, never the result.
#ifndef ARTB_RT_INTERSECT_CUH
#define ARTB_RT_INTERSECT_CUH

#include "flat_scene.h"
#include "rt_math.cuh"

namespace artb {

#define ARTB_STACK_DEPTH 32          // uint2 entries; at most 2 per level of 8-wide nodes
// ARTB_MAX_BVH8_DEPTH (14 levels): include/artb_flat.h — the limit is part of the ABI

// optional work counters (hostsim / instrumented builds): node visits, primitive tests, medium tests
struct TraversalStats { unsigned long long nodes, prims, media; };
#if defined(ARTB_TRAVERSAL_STATS)
#define ARTB_STAT(s, field) do { if (s) (s)->field++; } while (0)
#else
#define ARTB_STAT(s, field) do { } while (0)
#endif

// ---- transform chains ---------------------------------------------------------------------------
ARTB_HD void xform_op_in(const FlatXformOp& op, DRay& r)
{
    if (op.type == XFORM_TRANSLATE) {
        // translate::hit — hittable.cpp:4
        r.o = r.o - mk3(op.a, op.b, op.c);
    } else {
        // rotate_y::hit — hittable.cpp:61-65 (a = sin, b = cos)
        const double ox = op.b * r.o.x - op.a * r.o.z;
        const double oz = op.a * r.o.x + op.b * r.o.z;
        const double dx = op.b * r.d.x - op.a * r.d.z;
        const double dz = op.a * r.d.x + op.b * r.d.z;
        r.o.x = ox; r.o.z = oz; r.d.x = dx; r.d.z = dz;
    }
}

ARTB_HD DRay to_object_space(const SceneView& s, uint32_t xform, const DRay& world)
{
    DRay r = world;
    if (xform) {
        const FlatXform xf = s.xforms[xform];
        for (uint32_t i = 0; i < xf.n_ops; ++i) xform_op_in(s.xform_ops[xf.first_op + i], r);
    }
    return r;
}

// ---- primitive candidates: t only -----------------------------------------------------------------
ARTB_HD bool sphere_t(d3 center, double radius, const DRay& r, double t_min, double t_max, double& t)
{
    // sphere.h:40-57
    const d3 oc = r.o - center;
    const double a = length_squared(r.d);
    const double half_b = dot(oc, r.d);
    const double c = length_squared(oc) - radius * radius;
    const double discriminant = half_b * half_b - a * c;
    if (discriminant < 0) return false;
    const double sqrtd = sqrt(discriminant);
    double root = (-half_b - sqrtd) / a;
    if (root < t_min || t_max < root) {
        root = (-half_b + sqrtd) / a;
        if (root < t_min || t_max < root) return false;
    }
    t = root;
    return true;
}

ARTB_HD d3 msphere_center(const FlatPrim& p, double time)
{
    // moving_sphere.h:29-31
    const d3 c0 = ld3(p.g), c1 = ld3(p.g + 3);
    return c0 + ((time - p.g[6]) / (p.g[7] - p.g[6])) * (c1 - c0);
}

ARTB_HD bool rect_t(uint32_t type, const double* g, const DRay& r, double t_min, double t_max, double& t)
{
    // aarect.cpp:3-55; g = a0 a1 b0 b1 k
    double ok, dk, oa, da, ob, db;
    if (type == PRIM_RECT_XY)      { ok = r.o.z; dk = r.d.z; oa = r.o.x; da = r.d.x; ob = r.o.y; db = r.d.y; }
    else if (type == PRIM_RECT_XZ) { ok = r.o.y; dk = r.d.y; oa = r.o.x; da = r.d.x; ob = r.o.z; db = r.d.z; }
    else                           { ok = r.o.x; dk = r.d.x; oa = r.o.y; da = r.d.y; ob = r.o.z; db = r.d.z; }
    const double tt = (g[4] - ok) / dk;
    if (tt < t_min || tt > t_max) return false;
    const double a = oa + tt * da;
    const double b = ob + tt * db;
    if (a < g[0] || a > g[1] || b < g[2] || b > g[3]) return false;
    t = tt;
    return true;
}

// FP32 screen for rect_t: true only when the double-precision test above CERTAINLY returns false, so skipping the test
// cannot change a result. k_finish runs it over the global rectangles of a scene (the five walls and the light of a
// Cornell box) for all lanes at once and then hands each lane only its own surviving candidates: the exact tests ran
// one rectangle after the other with the 5 lanes of 32 that happened to hit it (profiles/r02_cornell_kernels.txt).
// Every float quantity carries an error bound that covers the roundings of the inputs (2^-24 each), of 1/d (2^-23) and of
// the operations, with a factor 2-4 to spare; a NaN or an infinity anywhere makes every comparison false, i.e. "not
// certain", and the exact test decides.
struct RectScreen { float k, a0, a1, b0, b1; uint32_t axes; };     // axes: k | a << 2 | b << 4; a0, b0 rounded down, a1, b1 up

ARTB_HD bool rect_screen_make(const FlatPrim& p, RectScreen& q)      // false: not an untransformed rectangle
{
    if (p.xform != 0u || (p.type != PRIM_RECT_XY && p.type != PRIM_RECT_XZ && p.type != PRIM_RECT_YZ)) return false;
    q.axes = p.type == PRIM_RECT_XY ? (2u | 0u << 2 | 1u << 4) : p.type == PRIM_RECT_XZ ? (1u | 0u << 2 | 2u << 4) : (0u | 1u << 2 | 2u << 4);
    q.a0 = d2f_rd(p.g[0]); q.a1 = d2f_ru(p.g[1]); q.b0 = d2f_rd(p.g[2]); q.b1 = d2f_ru(p.g[3]);
    q.k = float(p.g[4]);
    return true;
}

// o, d: the ray rounded to float; inv = fl(1 / d); t_min_f <= t_min and t_max_f >= t_max of the exact test
ARTB_HD bool rect_certain_miss_axes(const RectScreen& q, float ok, float ik, float oa, float da, float ob, float db, float t_min_f, float t_max_f)
{
    const float e21 = 4.76837158203125e-07f;        // 2^-21
    const float tt = (q.k - ok) * ik;
    const float et = fabsf(ik) * (fabsf(q.k) + fabsf(ok)) * e21 + fabsf(tt) * e21 + 1e-30f;
    if (tt + et < t_min_f || tt - et > t_max_f) return true;
    const float pa = tt * da, pb = tt * db;
    const float a = pa + oa, b = pb + ob;
    const float ea = et * fabsf(da) * 1.01f + (fabsf(oa) + fabsf(pa)) * e21 + 1e-30f;
    const float eb = et * fabsf(db) * 1.01f + (fabsf(ob) + fabsf(pb)) * e21 + 1e-30f;
    return a + ea < q.a0 || a - ea > q.a1 || b + eb < q.b0 || b - eb > q.b1;
}

ARTB_HD bool rect_certain_miss(const RectScreen& q, const float o[3], const float d[3], const float inv[3], float t_min_f, float t_max_f)
{
    // (the plane's axis is the same for every lane of a warp: a uniform branch, and no dynamically indexed arrays)
    const uint32_t ak = q.axes & 3u;
    if (ak == 2u) return rect_certain_miss_axes(q, o[2], inv[2], o[0], d[0], o[1], d[1], t_min_f, t_max_f);
    if (ak == 1u) return rect_certain_miss_axes(q, o[1], inv[1], o[0], d[0], o[2], d[2], t_min_f, t_max_f);
    return rect_certain_miss_axes(q, o[0], inv[0], o[1], d[1], o[2], d[2], t_min_f, t_max_f);
}

ARTB_HD bool triangle_t(const double* g, const DRay& r, double t_min, double t_max, double& t)
{
    // triangle.h:28-76
    const d3 pt1 = ld3(g), pt2 = ld3(g + 3), pt3 = ld3(g + 6);
    const d3 n = cross(pt2 - pt1, pt3 - pt1);
    const double ndotd = dot(n, r.d);
    if (fabs(ndotd) < DBL_EPSILON) return false;
    const double d = -dot(n, pt1);
    const double tt = -(dot(n, r.o) + d) / ndotd;
    if (tt < t_min || t_max < tt) return false;
    const d3 p = r.o + tt * r.d;
    if (dot(n, cross(pt2 - pt1, p - pt1)) < 0) return false;
    if (dot(n, cross(pt3 - pt2, p - pt2)) < 0) return false;
    if (dot(n, cross(pt1 - pt3, p - pt3)) < 0) return false;
    t = tt;
    return true;
}

ARTB_HD bool surface_t(const FlatPrim& p, const DRay& r, double t_min, double t_max, double& t)
{
    switch (p.type) {
        case PRIM_SPHERE:   return sphere_t(ld3(p.g), p.g[3], r, t_min, t_max, t);
        case PRIM_MSPHERE:  return sphere_t(msphere_center(p, r.tm), p.g[8], r, t_min, t_max, t);
        case PRIM_TRIANGLE: return triangle_t(p.g, r, t_min, t_max, t);
        default:            return rect_t(p.type, p.g, r, t_min, t_max, t);
    }
}

// closest boundary hit of a medium: hittable::hit of the boundary sub-graph, scanned like a list
ARTB_HD bool boundary_t(const SceneView& s, uint32_t first, uint32_t count, const DRay& world,
                        double t_min, double t_max, double& t)
{
    bool any = false;
    double closest = t_max;
    uint32_t cached = 0xffffffffu;
    DRay r = world;
    for (uint32_t i = first; i < first + count; ++i) {
        const FlatPrim& p = s.prims[i];
        if (p.xform != cached) { r = to_object_space(s, p.xform, world); cached = p.xform; }
        double tt;
        if (surface_t(p, r, t_min, closest, tt)) { any = true; closest = tt; }
    }
    t = closest;
    return any;
}

// Source of the one random number a constant_medium consumes per query (constant_medium.h:63):
// an explicit table (parity hook of artb_trace_batch) or a Philox stream keyed by (path, bounce, medium).
struct MediumXi {
    const double* row;       // n_media values for this ray, or nullptr
    uint32_t seed, stream, path_lo, path_hi, bounce;
    ARTB_HD double get(uint32_t slot) const
    {
        if (row) return row[slot];
        Philox g;
        g.init(seed, stream, path_lo, path_hi, (bounce << 8) | slot, 0x6d656469u /* 'medi' */);
        return g.next_double();
    }
};

// The two boundary queries of a medium (constant_medium.h:44-47): rec1 = boundary.hit(-inf, inf),
// rec2 = boundary.hit(rec1.t + 0.0001, inf).
ARTB_HD_NOINLINE bool medium_interval_general(const SceneView* scene, uint32_t first, uint32_t count, DRay world,
                                              double* t1, double* t2)
{
    const SceneView& s = *scene;
    if (!boundary_t(s, first, count, world, -INFINITY, INFINITY, *t1)) return false;
    return boundary_t(s, first, count, world, *t1 + 0.0001, INFINITY, *t2);
}

ARTB_HD bool medium_interval(const SceneView& s, uint32_t first, uint32_t count, const DRay& world, double& t1, double& t2)
{
    const FlatPrim& b = s.prims[first];
    if (count == 1 && b.type == PRIM_SPHERE) {
        // One sphere: both queries evaluate the same two roots (sphere.h:40-57), so compute them once. The first
        // query has no bounds and returns the smaller root; the second one rejects it (it is below its t_min) and
        // accepts the larger root iff that is >= t_min. Same expressions, hence the same numbers as two calls.
        const DRay r = to_object_space(s, b.xform, world);
        const d3 oc = r.o - ld3(b.g);
        const double a = length_squared(r.d);
        const double half_b = dot(oc, r.d);
        const double c = length_squared(oc) - b.g[3] * b.g[3];
        const double discriminant = half_b * half_b - a * c;
        if (discriminant < 0) return false;
        const double sqrtd = sqrt(discriminant);
        const double r1 = (-half_b - sqrtd) / a;
        const double r2 = (-half_b + sqrtd) / a;
        const double t_min2 = r1 + 0.0001;
        if (r1 < t_min2 && fabs(r1) <= DBL_MAX) {      // always, except for degenerate rays: those take the general path
            if (r2 < t_min2) return false;
            t1 = r1; t2 = r2;
            return true;
        }
    }
    return medium_interval_general(s.self, first, count, world, &t1, &t2);
}

// log(u) behind a call: the compiler must not start the (long) double-precision logarithm before the float test
// below has decided that it is needed
ARTB_HD_NOINLINE double medium_log(double u) { return log(u); }

ARTB_HD bool medium_t(const SceneView& s, const FlatPrim& p, const DRay& world, double t_min, double t_max,
                      const MediumXi& xi, double& t)
{
    // constant_medium.h:44-69
    const uint32_t first = uint32_t(p.g[1]), count = uint32_t(p.g[2]);
    const double u = xi.get(uint32_t(p.g[3]));
    // explicit-xi table: NaN marks "the reference drew nothing here", i.e. its interval was already empty
    // because a closer surface had been found first — then this medium cannot be the closest hit either
    if (!(u == u)) return false;
    // Most queries end with "scattered beyond the interval". Decide the clear cases with a float logarithm:
    // |log(u) - __logf(float(u))| <= 1e-6 (1 + |log u|) (2^-21.4 absolute / 3 ulp, plus the rounding of u), so a
    // lower bound of hit_distance that already exceeds the interval is a certain miss; everything else takes the
    // reference's double-precision expression below.
    const float lf = fast_logf(float(u));
    const double hit_distance_lo = -p.g[0] * double(-lf - 1e-6f * (1.f - lf));
    const double ray_length = length(world.d);
    // ... and before the boundary is even intersected: the clamped interval lies within [max(t_min, 0), t_max], and
    // subtraction and multiplication round monotonically, so (t_max - max(t_min, 0)) * |d| bounds it from above
    if (hit_distance_lo > (t_max - (t_min > 0 ? t_min : 0.0)) * ray_length) return false;
    double t1, t2;
    if (!medium_interval(s, first, count, world, t1, t2)) return false;
    if (t1 < t_min) t1 = t_min;
    if (t2 > t_max) t2 = t_max;
    if (t1 >= t2) return false;
    if (t1 < 0) t1 = 0;
    const double distance_inside_boundary = (t2 - t1) * ray_length;
    if (hit_distance_lo > distance_inside_boundary) return false;
    const double hit_distance = p.g[0] * medium_log(u);
    if (hit_distance > distance_inside_boundary) return false;
    t = t1 + hit_distance / ray_length;
    return true;
}

// ---- exact ties ---------------------------------------------------------------------------------------------------
// aabb::hit as the reference evaluates it (aabb.h:16-29): 1/d in double, strict rejection when the interval collapses
ARTB_HD bool ref_aabb_hit(const FlatRefNode& n, const DRay& r, double t_min, double t_max)
{
    const double o[3] = { r.o.x, r.o.y, r.o.z }, d[3] = { r.d.x, r.d.y, r.d.z };
    for (int a = 0; a < 3; ++a) {
        const double invD = 1.0 / d[a];
        double t0 = (n.lo[a] - o[a]) * invD;
        double t1 = (n.hi[a] - o[a]) * invD;
        if (invD < 0.0) { const double tmp = t0; t0 = t1; t1 = tmp; }
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max <= t_min) return false;
    }
    return true;
}

ARTB_HD int prim_ref_node(const FlatPrim& p) { return p.type == PRIM_MEDIUM ? -1 : int(p.g[9]) - 1; }

// Two leaves report the SAME t. In the reference the one tested later overwrites the earlier one (every leaf accepts
// t == t_max) — if it is tested at all: by then t_max is that very t, and every bvh_node between the two leaves' common
// ancestor and the later leaf first asks its box with it (bvh.cpp:45, aabb.h:25: t_max <= t_min rejects). Replays those
// box tests with the reference's arithmetic. Rare (coincident faces of adjacent boxes), hence not inlined.
ARTB_HD bool tie_later_wins(const SceneView& s, const DRay& world, double t_min, double t, uint32_t prim_earlier,
                                     uint32_t prim_later)
{
    if (!s.ref_nodes) return true;
    int a = prim_ref_node(s.prims[prim_earlier]), b = prim_ref_node(s.prims[prim_later]);
    // lowest common ancestor of the two chains of reference nodes (-1: none, the leaves meet in a plain list)
    int x = a, y = b;
    while (x != y) {
        const int dx = x < 0 ? -1 : s.ref_nodes[x].depth, dy = y < 0 ? -1 : s.ref_nodes[y].depth;
        if (dx >= dy) x = s.ref_nodes[x].parent;
        else y = s.ref_nodes[y].parent;
    }
    for (int n = b; n != x; n = s.ref_nodes[n].parent) {
        const FlatRefNode& node = s.ref_nodes[n];
        if (!ref_aabb_hit(node, to_object_space(s, node.xform, world), t_min, t)) return false;
    }
    return true;
}

// The merge rule, i.e. what hittable_list::hit / bvh_node::hit do by narrowing t_max (hittable_list.cpp:11-15,
// bvh.cpp:48-49): a candidate replaces the best iff it is closer; at equal t the leaf that comes later in the
// reference's depth-first order wins, provided the reference would have visited it (tie_later_wins).
ARTB_HD_NOINLINE bool tie_replaces(const SceneView* scene, DRay world, double t_min, double t, uint32_t prim, uint32_t dfs,
                                   uint32_t best_prim, uint32_t best_dfs)
{
    const SceneView& s = *scene;
    if (dfs > best_dfs) return tie_later_wins(s, world, t_min, t, best_prim, prim);
    return !tie_later_wins(s, world, t_min, t, prim, best_prim);
}

ARTB_HD bool hit_replaces(const SceneView& s, const DRay& world, double t_min, double t, uint32_t prim, uint32_t dfs, double best_t,
                          int32_t best_prim, uint32_t best_dfs)
{
    if (t < best_t || best_prim < 0) return t <= best_t;
    if (t != best_t) return false;
    return tie_replaces(s.self, world, t_min, t, prim, dfs, uint32_t(best_prim), best_dfs);  // rare, out of line
}

// ---- closest hit ----------------------------------------------------------------------------------
struct ClosestHit { double t; int32_t prim; uint32_t dfs; };

ARTB_HD void test_surface(const SceneView& s, uint32_t i, const DRay& world, double t_min, ClosestHit& best)
{
    const FlatPrim& p = s.prims[i];
    if (p.type == PRIM_MEDIUM) return;          // media never sit in the tree (flatten.cpp: they are all global)
    double t;
    bool ok;
    if (p.xform) ok = surface_t(p, to_object_space(s, p.xform, world), t_min, best.t, t);
    else ok = surface_t(p, world, t_min, best.t, t);
    if (ok && hit_replaces(s, world, t_min, t, i, p.dfs_id, best.t, best.prim, best.dfs)) {
        best.t = t; best.prim = int32_t(i); best.dfs = p.dfs_id;
    }
}

// Per-ray constants of the FP32 box tests. inv = fl(1/d), magnitude clamped to 1e20 so that no inf/NaN ever appears
// (a clamped slab interval is the true one scaled towards 0: every box here is at least 2e-9 thick around its
// primitive, so the clamp cannot cut off a reachable t). c = fl(o) * inv.
struct BoxRay {
    float inv[3];
    float c[3];
    uint32_t octinv4;        // (7 - octant) in every byte; octant bit 2: dx < 0, bit 1: dy < 0, bit 0: dz < 0
};

ARTB_HD BoxRay make_box_ray(const DRay& r)
{
    BoxRay b;
    const double o[3] = { r.o.x, r.o.y, r.o.z }, d[3] = { r.d.x, r.d.y, r.d.z };
    uint32_t oct = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double id = 1.0 / d[a];
        id = fmin(fmax(id, -1e20), 1e20);
        if (!(id == id)) id = 1e20;
        b.inv[a] = float(id);
        b.c[a] = float(o[a]) * b.inv[a];
        if (b.inv[a] < 0.f) oct |= 4u >> a;
    }
    b.octinv4 = (7u - oct) * 0x01010101u;
    return b;
}

// The ray as the FP32 tree walk sees it: origin and direction rounded to float (32 bytes, written next to the
// double-precision ray by whoever creates it).
struct ARTB_ALIGN(16) TravRay { float o[3], d[3], tm, pad; };

ARTB_HD TravRay make_trav_ray(const DRay& r)
{
    TravRay t;
    t.o[0] = float(r.o.x); t.o[1] = float(r.o.y); t.o[2] = float(r.o.z);
    t.d[0] = float(r.d.x); t.d[1] = float(r.d.y); t.d[2] = float(r.d.z);
    t.tm = float(r.tm); t.pad = 0.f;
    return t;
}

// Same constants from the float ray: inv = fl(1/fl(d)) carries the two roundings the error budget of node8_hits
// allows for, c = fl(o) * inv as above.
ARTB_HD BoxRay make_box_ray(const TravRay& t)
{
    BoxRay b;
    uint32_t oct = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float id = rcp_rn(t.d[a]);
        id = fminf(fmaxf(id, -1e20f), 1e20f);
        if (!(id == id)) id = 1e20f;
        b.inv[a] = id;
        b.c[a] = t.o[a] * id;
        if (id < 0.f) oct |= 4u >> a;
    }
    b.octinv4 = (7u - oct) * 0x01010101u;
    return b;
}

// Conservative slab test against an explicit float box (used for the media's world boxes). Error budget as in
// node8_hits below: planes are exact floats, so the only errors are the roundings of inv, c and the FMA.
ARTB_HD bool slab_box(const float* lo, const float* hi, const BoxRay& r, float t_min, float t_best)
{
    float tn = t_min, tf = t_best;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const bool pos = r.inv[a] >= 0.f;
        const float n = fmaf(pos ? lo[a] : hi[a], r.inv[a], -r.c[a]);
        const float f = fmaf(pos ? hi[a] : lo[a], r.inv[a], -r.c[a]);
        const float pad = (fabsf(r.c[a]) + fmaxf(fabsf(n), fabsf(f))) * 4.76837158203125e-07f + 1e-30f;   // 2^-21
        tn = fmaxf(tn, n - pad);
        tf = fminf(tf, f + pad);
    }
    return tn <= tf;
}

// One 8-wide node against one ray. Returns the hit mask in the layout the traversal stack uses: bits 24..31 =
// internal children, placed at 24 + (slot XOR (7-octant)) so that the highest set bit is the child to enter first;
// bits 0..23 = primitives of the leaf children that were hit, at their offsets from prim_base.
//
// Plane parameter on axis a for quantised coordinate q:  t = (p + q*s - o) / d.  Computed as fmaf(v, A, B) with
//   v = 1 + q*2^-15   (one PRMT drops the byte q into the mantissa of 1.0f — exact, no int->float conversion),
//   A = 2^15 * s * inv                (exact: powers of two),
//   B = fmaf(p, inv, -c) - A          (two roundings).
// Error budget (eps = 2^-24): inv carries (1+eps) twice (1/d in double, then to float), c = fl(fl(o)*inv) carries
// |c|*2 eps, B carries (|B| + |A|)*eps + |B|*eps, the final FMA |t|*eps, and |t| <= |B| + 2|A| for every plane of the
// node. All of it is below pad = (|c| + |B| + 2|A|) * 2^-21, which is SUBTRACTED from near planes and ADDED to far
// planes: the interval only grows. 2|A|*2^-21 is 1/32 of a quantisation cell, so the loosening is invisible next to
// the quantisation itself.
// child J of a group of four: byte J of each plane word -> mantissa of 1.0f (selector bytes: 00 | q_J | 80 | 3f)
template <int J>
ARTB_HD void node8_child(uint32_t nx4, uint32_t ny4, uint32_t nz4, uint32_t fx4, uint32_t fy4, uint32_t fz4, const float* A,
                         const float* Bn, const float* Bf, float t_min, float t_best, uint32_t child_bits4, uint32_t bit_index4,
                         uint32_t one, uint32_t& hitmask, float& tn_min)
{
    constexpr uint32_t kSel = 0x7604u | (uint32_t(J) << 4);
    const float tnx = fmaf(as_float(byte_perm_sel<kSel>(nx4, one)), A[0], Bn[0]);
    const float tny = fmaf(as_float(byte_perm_sel<kSel>(ny4, one)), A[1], Bn[1]);
    const float tnz = fmaf(as_float(byte_perm_sel<kSel>(nz4, one)), A[2], Bn[2]);
    const float tfx = fmaf(as_float(byte_perm_sel<kSel>(fx4, one)), A[0], Bf[0]);
    const float tfy = fmaf(as_float(byte_perm_sel<kSel>(fy4, one)), A[1], Bf[1]);
    const float tfz = fmaf(as_float(byte_perm_sel<kSel>(fz4, one)), A[2], Bf[2]);
    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, t_min));
    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, t_best));
    if (tn <= tf) {
        const uint32_t bits = (child_bits4 >> (8 * J)) & 0xffu;
        const uint32_t index = (bit_index4 >> (8 * J)) & 0xffu;
        hitmask |= bits << index;
        tn_min = fminf(tn_min, tn);
    }
}

// tn_min (in/out): lowered to the smallest near-plane parameter among the children that were hit — a conservative lower
// bound (the pads only widen the interval) of the t of anything the ray can hit inside them.
ARTB_HD uint32_t node8_hits(const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3, const uint4 n4, const BoxRay& r,
                            float t_min, float t_best, uint32_t one /* SceneView::one_bits */, float& tn_min)
{
    float A[3], Bn[3], Bf[3];
    const float p[3] = { as_float(n0.x), as_float(n0.y), as_float(n0.z) };
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float s15 = as_float((((n0.w >> (8 * a)) & 0xffu) + 15u) << 23);      // 2^15 * cell
        A[a] = s15 * r.inv[a];
        const float B = fmaf(p[a], r.inv[a], -r.c[a]) - A[a];
        const float pad = (fabsf(r.c[a]) + fabsf(B) + 2.f * fabsf(A[a])) * 4.76837158203125e-07f + 1e-30f;
        Bn[a] = B - pad;
        Bf[a] = B + pad;
    }
    // near planes: lo where the ray travels in +axis, hi otherwise (n2 = qlo.x | qlo.y, n3 = qlo.z | qhi.x, n4 = qhi.y | qhi.z)
    const bool px = r.inv[0] >= 0.f, py = r.inv[1] >= 0.f, pz = r.inv[2] >= 0.f;
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t meta4 = half ? n1.w : n1.z;
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;             // 24..31 have bits 3 and 4 set
        const uint32_t inner_mask4 = sign_extend_bytes(is_inner4 << 3);
        const uint32_t bit_index4 = (meta4 ^ (r.octinv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t qlox = half ? n2.y : n2.x, qloy = half ? n2.w : n2.z, qloz = half ? n3.y : n3.x;
        const uint32_t qhix = half ? n3.w : n3.z, qhiy = half ? n4.y : n4.x, qhiz = half ? n4.w : n4.z;
        const uint32_t nx4 = px ? qlox : qhix, fx4 = px ? qhix : qlox;
        const uint32_t ny4 = py ? qloy : qhiy, fy4 = py ? qhiy : qloy;
        const uint32_t nz4 = pz ? qloz : qhiz, fz4 = pz ? qhiz : qloz;
        node8_child<0>(nx4, ny4, nz4, fx4, fy4, fz4, A, Bn, Bf, t_min, t_best, child_bits4, bit_index4, one, hitmask, tn_min);
        node8_child<1>(nx4, ny4, nz4, fx4, fy4, fz4, A, Bn, Bf, t_min, t_best, child_bits4, bit_index4, one, hitmask, tn_min);
        node8_child<2>(nx4, ny4, nz4, fx4, fy4, fz4, A, Bn, Bf, t_min, t_best, child_bits4, bit_index4, one, hitmask, tn_min);
        node8_child<3>(nx4, ny4, nz4, fx4, fy4, fz4, A, Bn, Bf, t_min, t_best, child_bits4, bit_index4, one, hitmask, tn_min);
    }
    return hitmask;
}

// Traversal state of one lane: a node group {x = index of the first internal child, y = pending internal-child bits
// (24..31) | imask (0..7)} and a primitive group {x = prim_base, y = pending primitive bits (0..23)}; both kinds share
// the lane's stack. `alive` drops when the stack runs empty.
struct Traversal {
    uint2 ng, pg;
    int sp;
    float t_best_f;
    bool alive;
    ClosestHit best;
};

ARTB_HD void trav_begin(Traversal& tv, double t_max, bool has_tree)
{
    tv.ng = make_uint2(0u, has_tree ? 0x80000000u : 0u);      // the root as the only child of a virtual group
    tv.pg = make_uint2(0u, 0u);
    tv.sp = 0;
    tv.best.t = t_max; tv.best.prim = -1; tv.best.dfs = 0;
    tv.t_best_f = d2f_ru(t_max);
    tv.alive = true;
}

// One round of the WARP-SYNCHRONOUS traversal ("while-while" with a vote): the lanes named in `wm` either all take
// a node step or all take a primitive step, so FP32 box tests run next to box tests and the double-precision
// primitive tests next to primitive tests. A lane that has found candidate primitives waits (its node work is
// postponed, which keeps the pruning by best.t sharp); the primitive step is taken as soon as
// ARTB_PRIM_VOTE_NUM/DEN of the live lanes wait for one, or nobody can take a node step. Returns the ballot of the
// lanes that still have work. On the host (hostsim) a warp is one lane and the vote is trivially unanimous.
// NodePtr: const FlatNode8* in global memory, or the copy k_extend staged in shared memory.
#ifndef ARTB_PRIM_VOTE_NUM
#define ARTB_PRIM_VOTE_NUM 1
#define ARTB_PRIM_VOTE_DEN 2
#endif

template <class NodePtr, class Stack>
ARTB_HD uint32_t trav_round(const SceneView& s, NodePtr nodes, uint32_t wm, Traversal& tv, Stack& stack, const DRay& ray,
                            const BoxRay& br, double t_min, float t_min_f, TraversalStats* stats)
{
    // (A lane that has nothing pending pops its stack — or learns that its walk has ended — at the END of a round, below:
    // a caller that refills lanes between rounds, k_walk, then hands such a lane its next ray BEFORE the following round
    // instead of letting it idle through it. trav_begin leaves the root pending, so the first round needs no pop.)
    const uint32_t m_alive = warp_ballot(wm, tv.alive);
    if (m_alive == 0u) return 0u;
    const bool want_prims = tv.alive && tv.pg.y != 0u;
    const uint32_t m_prims = warp_ballot(wm, want_prims);
    if (ARTB_PRIM_VOTE_DEN * pop_count(m_prims) >= ARTB_PRIM_VOTE_NUM * pop_count(m_alive)) {
        // every lane of the warp takes this branch (the vote is uniform): one candidate per lane and round, with a
        // re-convergence point after each, so that the many exits of the primitive tests meet again
        bool more = want_prims;
        while (warp_ballot(wm, more) != 0u) {
            if (more) {
                const int bit = find_msb(tv.pg.y);
                tv.pg.y &= ~(1u << bit);
                ARTB_STAT(stats, prims);
                test_surface(s, tv.pg.x + uint32_t(bit), ray, t_min, tv.best);
                more = tv.pg.y != 0u;
            }
            warp_sync_mask(wm);
        }
        if (want_prims) tv.t_best_f = d2f_ru(tv.best.t);
    } else if (tv.alive && !want_prims && tv.ng.y > 0x00ffffffu) {
        // ng holds at least one internal child (always, except for a scene without a tree): enter the one the ray meets first
        const uint32_t hits = tv.ng.y;
        const int bit = find_msb(hits);
        tv.ng.y &= ~(1u << bit);
        if (tv.ng.y > 0x00ffffffu && tv.sp < ARTB_STACK_DEPTH) stack[tv.sp++] = tv.ng;
        const uint32_t slot = uint32_t(bit - 24) ^ (br.octinv4 & 0xffu);
        const uint32_t rel = uint32_t(pop_count(hits & ~(0xffffffffu << slot) & 0xffu));
        const uint4* q = reinterpret_cast<const uint4*>(nodes + (tv.ng.x + rel));
        const uint4 n0 = ld_node16(q), n1 = ld_node16(q + 1), n2 = ld_node16(q + 2), n3 = ld_node16(q + 3), n4 = ld_node16(q + 4);
        ARTB_STAT(stats, nodes);
        float tn_unused = 0.f;
        const uint32_t hitmask = node8_hits(n0, n1, n2, n3, n4, br, t_min_f, tv.t_best_f, s.one_bits, tn_unused);
        tv.ng.x = n1.x;
        tv.ng.y = (hitmask & 0xff000000u) | (n0.w >> 24);
        tv.pg.x = n1.y;
        tv.pg.y = hitmask & 0x00ffffffu;
        if (tv.pg.y) {            // start fetching the first candidate while other lanes keep walking
            const char* pf = reinterpret_cast<const char*>(s.prims + (tv.pg.x + uint32_t(find_msb(tv.pg.y))));
            prefetch_l1(pf);
            prefetch_l1(pf + 64);
        }
    }
    if (tv.alive && tv.pg.y == 0u && tv.ng.y <= 0x00ffffffu) {
        if (tv.sp == 0) tv.alive = false;
        else {
            tv.ng = stack[--tv.sp];
            if (tv.ng.y <= 0x00ffffffu) { tv.pg = tv.ng; tv.ng = make_uint2(0u, 0u); }
        }
    }
    return warp_ballot(wm, tv.alive);
}

// global surfaces: tested for every ray before the tree (same work for every lane, and an early t_max)
ARTB_HD void test_global_surfaces(const SceneView& s, const DRay& ray, double t_min, Traversal& tv, TraversalStats* stats)
{
    const uint32_t n_surf = s.n_global_prims - s.n_global_media;
    for (uint32_t i = 0; i < n_surf; ++i) {
        ARTB_STAT(stats, prims);
        test_surface(s, i, ray, t_min, tv.best);
    }
    tv.t_best_f = d2f_ru(tv.best.t);
}

// Media last: their candidate t does not depend on the order either, but with the final t_max the
// interval is short, a medium whose box starts beyond the closest surface is skipped after six FMAs, and most of
// the others leave through the float-logarithm exit of medium_t.
ARTB_HD void test_global_media(const SceneView& s, const DRay& ray, const BoxRay& br, double t_min, float t_min_f,
                               const MediumXi& xi, ClosestHit& best, TraversalStats* stats)
{
    for (uint32_t i = s.n_global_prims - s.n_global_media; i < s.n_global_prims; ++i) {
        const FlatPrim& p = s.prims[i];
        const float lo[3] = { d2f_rd(p.g[4]), d2f_rd(p.g[5]), d2f_rd(p.g[6]) };
        const float hi[3] = { d2f_ru(p.g[7]), d2f_ru(p.g[8]), d2f_ru(p.g[9]) };
        if (!slab_box(lo, hi, br, t_min_f, d2f_ru(best.t))) continue;
        ARTB_STAT(stats, media);
        double t;
        if (medium_t(s, p, ray, t_min, best.t, xi, t) && (t < best.t || best.prim < 0 || p.dfs_id > best.dfs)) {
            best.t = t; best.prim = int32_t(i); best.dfs = p.dfs_id;
        }
    }
}

// ---- two-phase closest hit: FP32 candidate collection, then double-precision confirmation -------------------------
// Phase 1 (k_traverse) walks the tree with the float ray only and emits a (ray, primitive) PAIR for every primitive
// whose quantised leaf box the ray enters within [t_min, t_max]: no primitive data, no doubles, no pruning by hits —
// so the walk is a pure FP32 / shared-memory loop with few registers, and since nothing a lane finds changes what it
// visits next, lanes can be refilled one by one. (Measured on the reference's scenes: 1.1-1.3x the node visits of the
// pruned walk.) Phase 2 (k_pairs) runs the reference's double-precision test on exactly those pairs, one pair per
// thread, and merges the results per ray. The box tests are conservative, so every primitive the reference could
// report is among the candidates, and the closest-hit rule (minimum t, ties to the larger depth-first id) does not
// depend on the order: the result is the same as the one-phase walk's.
struct Collector { uint2 ng; int sp; };

ARTB_HD void collect_begin(Collector& c, bool has_tree)
{
    c.ng = make_uint2(0u, has_tree ? 0x80000000u : 0u);
    c.sp = 0;
}

// One node visit, in two parts so that a warp can re-converge between them (k_traverse):
//   collect_pop   false when the walk has ended (nothing pending and the stack is empty)
//   collect_node  enters the nearest pending child; the primitives of its leaf children that were hit are
//                 prims[prim_base + b] for every set bit b of prim_mask (possibly none)
// Stack: anything indexable that yields uint2& — a lane-private array (uint2*), or k_traverse's shared-memory stack
template <class Stack>
ARTB_HD bool collect_pop(Collector& c, const Stack& stack)
{
    if (c.ng.y <= 0x00ffffffu) {
        if (c.sp == 0) return false;
        c.ng = stack[--c.sp];
    }
    return true;
}

template <class NodePtr, class Stack>
ARTB_HD void collect_node(NodePtr nodes, const BoxRay& br, float t_min_f, float t_max_f, Collector& c, Stack& stack,
                          uint32_t& prim_base, uint32_t& prim_mask, float& tn_min, uint32_t one_bits, TraversalStats* stats,
                          int stack_capacity = ARTB_STACK_DEPTH)
{
    tn_min = INFINITY;
    const uint32_t hits = c.ng.y;
    const int bit = find_msb(hits);
    c.ng.y &= ~(1u << bit);
    if (c.ng.y > 0x00ffffffu && c.sp < stack_capacity) stack[c.sp++] = c.ng;
    const uint32_t slot = uint32_t(bit - 24) ^ (br.octinv4 & 0xffu);
    const uint32_t rel = uint32_t(pop_count(hits & ~(0xffffffffu << slot) & 0xffu));
    const uint4* q = reinterpret_cast<const uint4*>(nodes + (c.ng.x + rel));
    const uint4 n0 = ld_node16(q), n1 = ld_node16(q + 1), n2 = ld_node16(q + 2), n3 = ld_node16(q + 3), n4 = ld_node16(q + 4);
    ARTB_STAT(stats, nodes);
    const uint32_t hitmask = node8_hits(n0, n1, n2, n3, n4, br, t_min_f, t_max_f, one_bits, tn_min);
    c.ng.x = n1.x;
    c.ng.y = (hitmask & 0xff000000u) | (n0.w >> 24);
    prim_base = n1.y;
    prim_mask = hitmask & 0x00ffffffu;
}

template <class NodePtr>
ARTB_HD bool collect_step(NodePtr nodes, const BoxRay& br, float t_min_f, float t_max_f, Collector& c, uint2* stack,
                          uint32_t& prim_base, uint32_t& prim_mask, uint32_t one_bits, TraversalStats* stats)
{
    prim_mask = 0u;
    if (!collect_pop(c, stack)) return false;
    float tn_min;
    collect_node(nodes, br, t_min_f, t_max_f, c, stack, prim_base, prim_mask, tn_min, one_bits, stats);
    return true;
}

// one ray per lane from start to end (k_trace_batch, hostsim); k_extend drives the same pieces itself so that it
// can hand a fresh ray to a lane as soon as its old one is finished
template <class NodePtr>
ARTB_HD ClosestHit closest_hit(const SceneView& s, NodePtr nodes, const DRay& ray, double t_min, double t_max, const MediumXi& xi,
                               TraversalStats* stats = nullptr)
{
    const uint32_t wm = warp_active_mask();
    const BoxRay br = make_box_ray(ray);
    const float t_min_f = d2f_rd(t_min);
    Traversal tv;
    trav_begin(tv, t_max, s.n_world_prims > s.n_global_prims);
    test_global_surfaces(s, ray, t_min, tv, stats);
    uint2 stack[ARTB_STACK_DEPTH];
    while (trav_round(s, nodes, wm, tv, stack, ray, br, t_min, t_min_f, stats) != 0u) { }
    test_global_media(s, ray, br, t_min, t_min_f, xi, tv.best, stats);
    return tv.best;
}

// both phases for one ray (hostsim; the kernels run the phases in separate launches): global surfaces, the pairs,
// then the media — the order k_finish / closest_hit use
#define ARTB_HOSTSIM_MAX_PAIRS 4096
template <class NodePtr>
ARTB_HD ClosestHit closest_hit_two_phase(const SceneView& s, NodePtr nodes, const DRay& ray, double t_min, double t_max,
                                         const MediumXi& xi, TraversalStats* stats = nullptr)
{
    const BoxRay br = make_box_ray(make_trav_ray(ray));
    Collector c;
    collect_begin(c, s.n_world_prims > s.n_global_prims);
    uint2 stack[ARTB_STACK_DEPTH];
    uint32_t ids[ARTB_HOSTSIM_MAX_PAIRS];
    uint32_t count = 0, base = 0, mask = 0;
    while (collect_step(nodes, br, d2f_rd(t_min), d2f_ru(t_max), c, stack, base, mask, s.one_bits, stats)) {
        while (mask) {
            const int b = find_msb(mask);
            mask &= ~(1u << b);
            if (count == ARTB_HOSTSIM_MAX_PAIRS) return closest_hit(s, nodes, ray, t_min, t_max, xi, stats);
            ids[count++] = base + uint32_t(b);
        }
    }
    ClosestHit best;
    best.t = t_max; best.prim = -1; best.dfs = 0;
    for (uint32_t k = 0; k < count; ++k) {
        ARTB_STAT(stats, prims);
        const FlatPrim& p = s.prims[ids[k]];
        if (p.type == PRIM_MEDIUM) continue;
        double t;
        const bool ok = p.xform ? surface_t(p, to_object_space(s, p.xform, ray), t_min, t_max, t) : surface_t(p, ray, t_min, t_max, t);
        if (ok && hit_replaces(s, ray, t_min, t, ids[k], p.dfs_id, best.t, best.prim, best.dfs)) { best.t = t; best.prim = int32_t(ids[k]); best.dfs = p.dfs_id; }
    }
    const uint32_t n_surf = s.n_global_prims - s.n_global_media;
    for (uint32_t i = 0; i < n_surf; ++i) {
        ARTB_STAT(stats, prims);
        test_surface(s, i, ray, t_min, best);
    }
    if (s.n_global_media) test_global_media(s, ray, br, t_min, d2f_rd(t_min), xi, best, stats);
    return best;
}

} // namespace artb

#endif
