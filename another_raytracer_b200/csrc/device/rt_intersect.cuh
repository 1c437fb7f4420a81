// rt_intersect.cuh — closest-hit query over the flattened scene (the "extend" stage).
//
// Replaces, in one iterative short-stack traversal: hittable_list::hit (hittable_list.cpp:5-19),
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
// reference's arithmetic), so t is the same number up to FMA contraction.
#ifndef ARTB_RT_INTERSECT_CUH
#define ARTB_RT_INTERSECT_CUH

#include "flat_scene.h"
#include "rt_math.cuh"

namespace artb {

#define ARTB_STACK_DEPTH 64

// ---- transform chains ---------------------------------------------------------------------------
__device__ __forceinline__ void xform_op_in(const FlatXformOp& op, DRay& r)
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

__device__ __forceinline__ DRay to_object_space(const SceneView& s, uint32_t xform, const DRay& world)
{
    DRay r = world;
    if (xform) {
        const FlatXform xf = s.xforms[xform];
        for (uint32_t i = 0; i < xf.n_ops; ++i) xform_op_in(s.xform_ops[xf.first_op + i], r);
    }
    return r;
}

// ---- primitive candidates: t only -----------------------------------------------------------------
__device__ __forceinline__ bool sphere_t(d3 center, double radius, const DRay& r, double t_min, double t_max, double& t)
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

__device__ __forceinline__ d3 msphere_center(const FlatPrim& p, double time)
{
    // moving_sphere.h:29-31
    const d3 c0 = ld3(p.g), c1 = ld3(p.g + 3);
    return c0 + ((time - p.g[6]) / (p.g[7] - p.g[6])) * (c1 - c0);
}

__device__ __forceinline__ bool rect_t(uint32_t type, const double* g, const DRay& r, double t_min, double t_max, double& t)
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

__device__ __forceinline__ bool triangle_t(const double* g, const DRay& r, double t_min, double t_max, double& t)
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

__device__ __forceinline__ bool surface_t(const FlatPrim& p, const DRay& r, double t_min, double t_max, double& t)
{
    switch (p.type) {
        case PRIM_SPHERE:   return sphere_t(ld3(p.g), p.g[3], r, t_min, t_max, t);
        case PRIM_MSPHERE:  return sphere_t(msphere_center(p, r.tm), p.g[8], r, t_min, t_max, t);
        case PRIM_TRIANGLE: return triangle_t(p.g, r, t_min, t_max, t);
        default:            return rect_t(p.type, p.g, r, t_min, t_max, t);
    }
}

// closest boundary hit of a medium: hittable::hit of the boundary sub-graph, scanned like a list
__device__ inline bool boundary_t(const SceneView& s, uint32_t first, uint32_t count, const DRay& world,
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
    __device__ __forceinline__ double get(uint32_t slot) const
    {
        if (row) return row[slot];
        Philox g;
        g.init(seed, stream, path_lo, path_hi, (bounce << 8) | slot, 0x6d656469u /* 'medi' */);
        return g.next_double();
    }
};

// The two boundary queries of a medium (constant_medium.h:44-47): rec1 = boundary.hit(-inf, inf),
// rec2 = boundary.hit(rec1.t + 0.0001, inf).
__device__ __noinline__ bool medium_interval_general(const SceneView& s, uint32_t first, uint32_t count, const DRay& world,
                                                     double& t1, double& t2)
{
    if (!boundary_t(s, first, count, world, -INFINITY, INFINITY, t1)) return false;
    return boundary_t(s, first, count, world, t1 + 0.0001, INFINITY, t2);
}

__device__ __forceinline__ bool medium_interval(const SceneView& s, uint32_t first, uint32_t count, const DRay& world,
                                                double& t1, double& t2)
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
    return medium_interval_general(s, first, count, world, t1, t2);
}

__device__ __forceinline__ bool medium_t(const SceneView& s, const FlatPrim& p, const DRay& world, double t_min, double t_max,
                                         const MediumXi& xi, double& t)
{
    // constant_medium.h:44-69
    const uint32_t first = uint32_t(p.g[1]), count = uint32_t(p.g[2]);
    double t1, t2;
    if (!medium_interval(s, first, count, world, t1, t2)) return false;
    if (t1 < t_min) t1 = t_min;
    if (t2 > t_max) t2 = t_max;
    if (t1 >= t2) return false;
    if (t1 < 0) t1 = 0;
    const double ray_length = length(world.d);
    const double distance_inside_boundary = (t2 - t1) * ray_length;
    const double u = xi.get(uint32_t(p.g[3]));
    // explicit-xi table: NaN marks "the reference drew nothing here", i.e. its interval was already empty
    // because a closer surface had been found first — then this medium cannot be the closest hit either
    if (!(u == u)) return false;
    // Most queries end here with "scattered beyond the interval". Decide the clear cases with a float logarithm:
    // |log(u) - __logf(float(u))| <= 1e-6 (1 + |log u|) (2^-21.4 absolute / 3 ulp, plus the rounding of u), so a
    // lower bound of hit_distance that already exceeds the interval is a certain miss; everything else takes the
    // reference's double-precision expression below.
    const float lf = __logf(float(u));
    const double hit_distance_lo = -p.g[0] * double(-lf - 1e-6f * (1.f - lf));
    if (hit_distance_lo > distance_inside_boundary) return false;
    const double hit_distance = p.g[0] * log(u);
    if (hit_distance > distance_inside_boundary) return false;
    t = t1 + hit_distance / ray_length;
    return true;
}

// ---- BVH traversal ------------------------------------------------------------------------------
struct ClosestHit { double t; int32_t prim; uint32_t dfs; };

// surfaces of one leaf (or of the global list); media never sit in leaves (flatten.cpp: they are all global)
__device__ __forceinline__ void test_prims(const SceneView& s, uint32_t first, uint32_t count, const DRay& world,
                                           double t_min, ClosestHit& best, uint32_t& cached_xform, DRay& cached_ray)
{
    for (uint32_t i = first; i < first + count; ++i) {
        const FlatPrim& p = s.prims[i];
        if (p.type == PRIM_MEDIUM) continue;
        if (p.xform != cached_xform) { cached_ray = to_object_space(s, p.xform, world); cached_xform = p.xform; }
        double t;
        if (surface_t(p, cached_ray, t_min, best.t, t) && (t < best.t || best.prim < 0 || p.dfs_id > best.dfs)) {
            best.t = t; best.prim = int32_t(i); best.dfs = p.dfs_id;
        }
    }
}

// Conservative FP32 slab test. The ray is held as inv = fl(1/d) (magnitude clamped to 1e30, so no inf/NaN ever
// appears) and, per axis, c = fl(o)*inv with an absolute pad |c|*2^-22 that covers the rounding of the origin to
// float, of c itself and of the fused multiply-add; the near plane subtracts c+pad, the far plane c-pad. A final
// relative widening by 2^-21 covers the rounding of inv and of the node planes (which the flattener already rounds
// OUTWARD). The interval can only grow, never shrink: a ray that hits the double-precision box always passes, and
// since every candidate primitive is then tested in double with the reference's formula, results do not depend on
// this test at all — only the amount of work does. Inclusive comparison, so equal-t candidates are still visited.
struct SlabRay {
    float ix, iy, iz;        // 1/d
    float nx, ny, nz;        // c + pad  (subtracted on the near plane)
    float fx, fy, fz;        // c - pad  (subtracted on the far plane)
    bool px, py, pz;         // direction sign >= 0: near plane is lo
};

__device__ __forceinline__ void slab_axis(double o, double d, float& inv, float& cn, float& cf, bool& pos)
{
    double id = 1.0 / d;
    id = fmin(fmax(id, -1e30), 1e30);
    if (!(id == id)) id = 1e30;
    inv = float(id);
    pos = inv >= 0.f;
    const float c = float(o) * inv;
    const float pad = fabsf(c) * 2.384185791015625e-07f + 1e-30f;   // 2^-22
    cn = c + pad;
    cf = c - pad;
}

__device__ __forceinline__ SlabRay make_slab_ray(const DRay& r)
{
    SlabRay s;
    slab_axis(r.o.x, r.d.x, s.ix, s.nx, s.fx, s.px);
    slab_axis(r.o.y, r.d.y, s.iy, s.ny, s.fy, s.py);
    slab_axis(r.o.z, r.d.z, s.iz, s.nz, s.fz, s.pz);
    return s;
}

__device__ __forceinline__ bool slab(const float* lo, const float* hi, const SlabRay& r, float t_min, float t_best, float& enter)
{
    const float nx = fmaf(r.px ? lo[0] : hi[0], r.ix, -r.nx), fx = fmaf(r.px ? hi[0] : lo[0], r.ix, -r.fx);
    const float ny = fmaf(r.py ? lo[1] : hi[1], r.iy, -r.ny), fy = fmaf(r.py ? hi[1] : lo[1], r.iy, -r.fy);
    const float nz = fmaf(r.pz ? lo[2] : hi[2], r.iz, -r.nz), fz = fmaf(r.pz ? hi[2] : lo[2], r.iz, -r.fz);
    float tn = fmaxf(fmaxf(nx, ny), fmaxf(nz, t_min));
    float tf = fminf(fminf(fx, fy), fminf(fz, t_best));
    tn -= fabsf(tn) * 4.76837158203125e-07f;    // 2^-21
    tf += fabsf(tf) * 4.76837158203125e-07f;
    enter = tn;
    return tn <= tf;
}

// "while-while" traversal (Aila & Laine 2009): an inner loop that only walks internal nodes until this lane holds
// a leaf, then an inner loop that only intersects primitives. The lanes of a warp re-converge between the two, so
// node tests run next to node tests and primitive tests next to primitive tests.
__device__ __forceinline__ void traverse_bvh(const SceneView& s, const DRay& ray, const SlabRay& sr, double t_min, float t_min_f,
                                             ClosestHit& best, uint32_t& cached_xform, DRay& cached_ray)
{
    float t_best_f = __double2float_ru(best.t);
    int32_t stack[ARTB_STACK_DEPTH];
    int sp = 0;
    int32_t cur = s.root;            // >= 0 internal node, < 0 leaf code
    const int32_t kDone = 0x7fffffff;
    while (cur != kDone) {
        // ---- walk internal nodes until a leaf is in hand
        while (cur >= 0 && cur != kDone) {
            const float4* q = reinterpret_cast<const float4*>(s.nodes + cur);
            const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
            const float lo0[3] = { a.x, a.y, a.z }, hi0[3] = { a.w, b.x, b.y };
            const float lo1[3] = { b.z, b.w, c.x }, hi1[3] = { c.y, c.z, c.w };
            const int32_t c0 = __float_as_int(d.x), c1 = __float_as_int(d.y);
            float e0, e1;
            const bool h0 = slab(lo0, hi0, sr, t_min_f, t_best_f, e0);
            const bool h1 = slab(lo1, hi1, sr, t_min_f, t_best_f, e1);
            if (h0 && h1) {
                const bool first0 = e0 <= e1;
                cur = first0 ? c0 : c1;
                if (sp < ARTB_STACK_DEPTH) stack[sp++] = first0 ? c1 : c0;
            } else if (h0) cur = c0;
            else if (h1) cur = c1;
            else cur = sp ? stack[--sp] : kDone;
        }
        // ---- intersect the leaf
        if (cur != kDone) {
            const uint32_t code = uint32_t(~cur);
            test_prims(s, code >> 3, (code & 7u) + 1u, ray, t_min, best, cached_xform, cached_ray);
            t_best_f = __double2float_ru(best.t);
            cur = sp ? stack[--sp] : kDone;
        }
    }
}

__device__ inline ClosestHit closest_hit(const SceneView& s, const DRay& ray, double t_min, double t_max, const MediumXi& xi)
{
    ClosestHit best;
    best.t = t_max; best.prim = -1; best.dfs = 0;
    uint32_t cached_xform = 0;
    DRay cached_ray = ray;

    const uint32_t n_surf = s.n_global_prims - s.n_global_media;
    const SlabRay sr = make_slab_ray(ray);
    const float t_min_f = __double2float_rd(t_min);
    // global surfaces first: same work for every lane, and an early t_max for the traversal
    if (n_surf) test_prims(s, 0, n_surf, ray, t_min, best, cached_xform, cached_ray);
    if (s.n_world_prims > s.n_global_prims) traverse_bvh(s, ray, sr, t_min, t_min_f, best, cached_xform, cached_ray);
    // Media last: their candidate t does not depend on the order either, but with the final t_max the
    // interval is short, a medium whose box starts beyond the closest surface is skipped after six FMAs, and most of
    // the others leave through the float-logarithm exit of medium_t.
    for (uint32_t i = n_surf; i < s.n_global_prims; ++i) {
        const FlatPrim& p = s.prims[i];
        const float lo[3] = { __double2float_rd(p.g[4]), __double2float_rd(p.g[5]), __double2float_rd(p.g[6]) };
        const float hi[3] = { __double2float_ru(p.g[7]), __double2float_ru(p.g[8]), __double2float_ru(p.g[9]) };
        float enter;
        if (!slab(lo, hi, sr, t_min_f, __double2float_ru(best.t), enter)) continue;
        double t;
        if (medium_t(s, p, ray, t_min, best.t, xi, t) && (t < best.t || best.prim < 0 || p.dfs_id > best.dfs)) {
            best.t = t; best.prim = int32_t(i); best.dfs = p.dfs_id;
        }
    }
    return best;
}

} // namespace artb

#endif
