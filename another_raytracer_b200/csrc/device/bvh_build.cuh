// bvh_build.cuh — builds the 8-wide quantised BVH (flat_scene.h: FlatNode8) ON THE DEVICE.
//
// Replaces the constructor of bvh_node (bvh.cpp:3-42: random axis, std::sort, median split, and a full copy of the
// object vector per node, i.e. O(N^2)) for the tree the kernels traverse. The closest hit does not depend on the
// acceleration structure (DESIGN.md "Closest hit"): ties and ids live in FlatPrim::dfs_id, every candidate is re-tested
// in double precision, so any tree whose boxes CONTAIN their primitives gives the reference's answers. That is what
// tests/test_gpu_bvh_build.py checks (same ids, bit-identical t with the host-built and the device-built tree).
//
// Pipeline (all on one stream, no host data except a few counters):
//   k_bvh_boxes    world-space box of every primitive, in double with the reference's formulas (sphere.h:67-72,
//                  moving_sphere.h:60-69, aarect.h:16-21/36-41/56-61, triangle.h:90-95, the eight rotated corners of
//                  hittable.cpp:38-52), rounded OUTWARD to float; scene bounds by atomics
//   k_bvh_morton   63-bit Morton code of the box centre
//   cub radix sort (keys: codes, values: primitive indices)
//   k_bvh_karras   binary radix tree (Karras, HPG 2012): every internal node finds its range and split by itself
//   k_bvh_refit    bottom-up boxes; the second thread to arrive at a node does the union
//   k_bvh_collapse level by level: a wide node adopts a binary node and opens the child with the largest surface area
//                  until it holds eight; children are placed in octant order, boxes quantised to the node's 8-bit grid
//                  (lo planes rounded down, hi planes up), leaf primitives laid out next to each other
//   k_bvh_permute  primitives into leaf order
#ifndef ARTB_BVH_BUILD_CUH
#define ARTB_BVH_BUILD_CUH

#include "flat_scene.h"
#include "renderer.h"
#include "rt_math.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace artb {

namespace bvhb {

struct Arrays {
    float4* box_lo; float4* box_hi;          // per primitive (build order)
    float4* node_lo; float4* node_hi;        // per binary node: internal 0..n-2, leaf (n-1)+k
    int* left; int* right; int* parent;      // children / parent of binary nodes (parent of the root: -1)
    int* first; int* last;                   // sorted-position range of every internal node
    unsigned int* flag;                      // arrival counters of the refit
    unsigned long long* keys; unsigned long long* keys_sorted;
    uint32_t* idx; uint32_t* idx_sorted;
    int* bounds;                             // 6 ordered-int floats: min xyz, max xyz of the box centres
    uint32_t* perm;                          // leaf order -> build-order primitive index
    uint2* queue[2];                         // {wide node, binary node}
    uint32_t* counters;                      // [0] wide nodes, [1] leaf prims laid out, [2] next-level queue length
};

__device__ __forceinline__ int float_to_ordered(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ inline void grow(double lo[3], double hi[3], double x, double y, double z)
{
    lo[0] = fmin(lo[0], x); hi[0] = fmax(hi[0], x);
    lo[1] = fmin(lo[1], y); hi[1] = fmax(hi[1], y);
    lo[2] = fmin(lo[2], z); hi[2] = fmax(hi[2], z);
}

__global__ void k_bvh_boxes(const FlatPrim* prims, uint32_t n, const FlatXformOp* ops, const FlatXform* xforms, Arrays A)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const FlatPrim& p = prims[i];
    double lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    switch (p.type) {
        case PRIM_SPHERE: {
            const double r = fabs(p.g[3]);
            grow(lo, hi, p.g[0] - r, p.g[1] - r, p.g[2] - r);
            grow(lo, hi, p.g[0] + r, p.g[1] + r, p.g[2] + r);
            break;
        }
        case PRIM_MSPHERE: {
            const double r = fabs(p.g[8]);
            for (int k = 0; k < 2; ++k) {
                grow(lo, hi, p.g[3 * k] - r, p.g[3 * k + 1] - r, p.g[3 * k + 2] - r);
                grow(lo, hi, p.g[3 * k] + r, p.g[3 * k + 1] + r, p.g[3 * k + 2] + r);
            }
            break;
        }
        case PRIM_RECT_XY: grow(lo, hi, p.g[0], p.g[2], p.g[4] - 0.0001); grow(lo, hi, p.g[1], p.g[3], p.g[4] + 0.0001); break;
        case PRIM_RECT_XZ: grow(lo, hi, p.g[0], p.g[4] - 0.0001, p.g[2]); grow(lo, hi, p.g[1], p.g[4] + 0.0001, p.g[3]); break;
        case PRIM_RECT_YZ: grow(lo, hi, p.g[4] - 0.0001, p.g[0], p.g[2]); grow(lo, hi, p.g[4] + 0.0001, p.g[1], p.g[3]); break;
        case PRIM_TRIANGLE:
            for (int k = 0; k < 3; ++k) grow(lo, hi, p.g[3 * k], p.g[3 * k + 1], p.g[3 * k + 2]);
            break;
        default: lo[0] = lo[1] = lo[2] = hi[0] = hi[1] = hi[2] = 0; break;      // media are never tree primitives
    }
    if (p.xform) {
        // the eight corners through the chain, innermost op first on the way out (hittable.cpp:38-52, :14-23)
        const FlatXform xf = xforms[p.xform];
        double wlo[3] = { INFINITY, INFINITY, INFINITY }, whi[3] = { -INFINITY, -INFINITY, -INFINITY };
        for (int c = 0; c < 8; ++c) {
            double q[3] = { (c & 1) ? hi[0] : lo[0], (c & 2) ? hi[1] : lo[1], (c & 4) ? hi[2] : lo[2] };
            for (uint32_t k = xf.n_ops; k-- > 0;) {
                const FlatXformOp& op = ops[xf.first_op + k];
                if (op.type == XFORM_ROTATE_Y) {
                    const double x = op.b * q[0] + op.a * q[2];
                    const double z = -op.a * q[0] + op.b * q[2];
                    q[0] = x; q[2] = z;
                } else {
                    q[0] += op.a; q[1] += op.b; q[2] += op.c;
                }
            }
            grow(wlo, whi, q[0], q[1], q[2]);
        }
        for (int a = 0; a < 3; ++a) { lo[a] = wlo[a]; hi[a] = whi[a]; }
    }
    float flo[3], fhi[3];
    for (int a = 0; a < 3; ++a) {   // conservative slack, then outward to float: candidates are re-tested exactly anyway
        const double pad = 1e-9 * (1.0 + fmax(fabs(lo[a]), fabs(hi[a])));
        flo[a] = __double2float_rd(lo[a] - pad);
        fhi[a] = __double2float_ru(hi[a] + pad);
    }
    A.box_lo[i] = make_float4(flo[0], flo[1], flo[2], 0.f);
    A.box_hi[i] = make_float4(fhi[0], fhi[1], fhi[2], 0.f);
    for (int a = 0; a < 3; ++a) {
        const float c = 0.5f * flo[a] + 0.5f * fhi[a];
        atomicMin(A.bounds + a, float_to_ordered(c));
        atomicMax(A.bounds + 3 + a, float_to_ordered(c));
    }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x)
{
    x &= 0x1fffffull;
    x = (x | (x << 32)) & 0x1f00000000ffffull;
    x = (x | (x << 16)) & 0x1f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

__global__ void k_bvh_morton(uint32_t n, Arrays A)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long code = 0;
    const float4 lo = A.box_lo[i], hi = A.box_hi[i];
    const float c[3] = { 0.5f * lo.x + 0.5f * hi.x, 0.5f * lo.y + 0.5f * hi.y, 0.5f * lo.z + 0.5f * hi.z };
    for (int a = 0; a < 3; ++a) {
        const float mn = ordered_to_float(A.bounds[a]), mx = ordered_to_float(A.bounds[3 + a]);
        const float ext = mx - mn;
        float u = ext > 0.f ? (c[a] - mn) / ext : 0.f;
        u = fminf(fmaxf(u, 0.f), 1.f);
        const unsigned long long q = (unsigned long long)fminf(u * 2097152.f, 2097151.f);
        code |= spread21(q) << (2 - a);
    }
    A.keys[i] = code;
    A.idx[i] = i;
}

// length of the common prefix of the keys at sorted positions i and j; equal keys are told apart by the position
__device__ __forceinline__ int delta(const unsigned long long* keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

__global__ void k_bvh_karras(int n, Arrays A)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const unsigned long long* keys = A.keys_sorted;
    const int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int div = 2, t = (l + div - 1) / div; ; div *= 2, t = (l + div - 1) / div) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t <= 1) break;
    }
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const int left = lo == gamma ? (n - 1) + gamma : gamma;
    const int right = hi == gamma + 1 ? (n - 1) + gamma + 1 : gamma + 1;
    A.left[i] = left; A.right[i] = right;
    A.first[i] = lo; A.last[i] = hi;
    A.parent[left] = i; A.parent[right] = i;
    if (i == 0) A.parent[0] = -1;
}

__global__ void k_bvh_refit(int n, Arrays A)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t prim = A.idx_sorted[k];
    int node = (n - 1) + k;
    A.node_lo[node] = A.box_lo[prim];
    A.node_hi[node] = A.box_hi[prim];
    __threadfence();
    node = A.parent[node];
    while (node >= 0) {
        if (atomicAdd(A.flag + node, 1u) == 0u) return;      // the sibling subtree is not done yet: its thread continues
        __threadfence();
        const int l = A.left[node], r = A.right[node];
        const float4 a = __ldcg(A.node_lo + l), b = __ldcg(A.node_lo + r);
        const float4 c = __ldcg(A.node_hi + l), d = __ldcg(A.node_hi + r);
        A.node_lo[node] = make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), 0.f);
        A.node_hi[node] = make_float4(fmaxf(c.x, d.x), fmaxf(c.y, d.y), fmaxf(c.z, d.z), 0.f);
        __threadfence();
        node = A.parent[node];
    }
}

__device__ __forceinline__ int subtree_count(const Arrays& A, int n, int node) { return node >= n - 1 ? 1 : A.last[node] - A.first[node] + 1; }
__device__ __forceinline__ int subtree_first(const Arrays& A, int n, int node) { return node >= n - 1 ? node - (n - 1) : A.first[node]; }

__device__ __forceinline__ float half_area(float4 lo, float4 hi)
{
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

// one thread = one wide node of the current level
__global__ void k_bvh_collapse(int n, int n_jobs, int level_in, int max_leaf, Arrays A, FlatNode8* nodes, uint32_t node_capacity,
                               int* error)
{
    const int job = blockIdx.x * blockDim.x + threadIdx.x;
    if (job >= n_jobs) return;
    const uint2 me = A.queue[level_in][job];
    const int root = int(me.y);
    int kids[8];
    int nk = 0;
    if (root >= n - 1 || subtree_count(A, n, root) <= max_leaf) kids[nk++] = root;       // a tree that is a single leaf
    else { kids[nk++] = A.left[root]; kids[nk++] = A.right[root]; }
    while (nk < 8) {        // open the child with the largest surface area
        int best = -1;
        float best_area = -1.f;
        for (int k = 0; k < nk; ++k) {
            const int c = kids[k];
            if (c >= n - 1 || subtree_count(A, n, c) <= max_leaf) continue;
            const float area = half_area(A.node_lo[c], A.node_hi[c]);
            if (area > best_area) { best_area = area; best = k; }
        }
        if (best < 0) break;
        const int c = kids[best];
        kids[best] = A.left[c];
        kids[nk++] = A.right[c];
    }
    // slots: a ray of octant o visits slot o first, so slot s wants the child that is nearest along
    // dir_s = (s&4 ? -1 : +1, s&2 ? -1 : +1, s&1 ? -1 : +1): greedy assignment on that cost (as host/flatten.cpp)
    const float4 blo = A.node_lo[root], bhi = A.node_hi[root];
    const float centre[3] = { 0.5f * blo.x + 0.5f * bhi.x, 0.5f * blo.y + 0.5f * bhi.y, 0.5f * blo.z + 0.5f * bhi.z };
    int kid_in[8];
    bool placed[8];
    for (int k = 0; k < 8; ++k) { kid_in[k] = -1; placed[k] = false; }
    for (int round = 0; round < nk; ++round) {
        float best = INFINITY;
        int bk = -1, bs = -1;
        for (int k = 0; k < nk; ++k) {
            if (placed[k]) continue;
            const float4 lo = A.node_lo[kids[k]], hi = A.node_hi[kids[k]];
            const float rel[3] = { 0.5f * lo.x + 0.5f * hi.x - centre[0], 0.5f * lo.y + 0.5f * hi.y - centre[1],
                                   0.5f * lo.z + 0.5f * hi.z - centre[2] };
            for (int sl = 0; sl < 8; ++sl) {
                if (kid_in[sl] >= 0) continue;
                float cost = 0.f;
                for (int a = 0; a < 3; ++a) cost += ((sl >> (2 - a)) & 1) ? -rel[a] : rel[a];
                if (cost < best) { best = cost; bk = k; bs = sl; }
            }
        }
        placed[bk] = true; kid_in[bs] = bk;
    }
    // quantisation grid: origin = the node's lo corner, cell = smallest power of two that keeps every plane of every
    // child within 255 cells
    FlatNode8 node;
    memset(&node, 0, sizeof(node));
    const float plo[3] = { blo.x, blo.y, blo.z }, phi[3] = { bhi.x, bhi.y, bhi.z };
    double cell[3];
    for (int a = 0; a < 3; ++a) {
        node.p[a] = plo[a];
        const double extent = double(phi[a]) - double(plo[a]);
        int e = 0;
        frexp(extent / 255.0, &e);
        if (e < -100) e = -100;
        while (ceil(extent / ldexp(1.0, e)) > 255.0) ++e;
        if (e > 100) { *error = 1; return; }
        node.e[a] = uint8_t(e + 127);
        cell[a] = ldexp(1.0, e);
    }
    int n_internal = 0, n_leaf_prims = 0;
    for (int sl = 0; sl < 8; ++sl) {
        if (kid_in[sl] < 0) continue;
        const int c = kids[kid_in[sl]];
        if (c < n - 1 && subtree_count(A, n, c) > max_leaf) n_internal++;
        else n_leaf_prims += subtree_count(A, n, c);
    }
    const uint32_t child_base = n_internal ? atomicAdd(A.counters + 0, uint32_t(n_internal)) : 0u;
    const uint32_t prim_base = n_leaf_prims ? atomicAdd(A.counters + 1, uint32_t(n_leaf_prims)) : 0u;
    if (child_base + uint32_t(n_internal) > node_capacity) { *error = 2; return; }
    const uint32_t queue_base = n_internal ? atomicAdd(A.counters + 2, uint32_t(n_internal)) : 0u;
    node.child_base = child_base;
    node.prim_base = prim_base;                    // relative to the first tree primitive; the caller adds that offset
    uint32_t prim_offset = 0, rank = 0;
    for (int sl = 0; sl < 8; ++sl) {
        if (kid_in[sl] < 0) continue;
        const int c = kids[kid_in[sl]];
        const float4 lo4 = A.node_lo[c], hi4 = A.node_hi[c];
        const float clo[3] = { lo4.x, lo4.y, lo4.z }, chi[3] = { hi4.x, hi4.y, hi4.z };
        for (int a = 0; a < 3; ++a) {
            double lo = floor((double(clo[a]) - double(plo[a])) / cell[a]);
            double hi = ceil((double(chi[a]) - double(plo[a])) / cell[a]);
            while (double(plo[a]) + lo * cell[a] > double(clo[a])) lo -= 1;
            while (double(plo[a]) + hi * cell[a] < double(chi[a])) hi += 1;
            if (lo < 0) lo = 0;                     // cannot happen: a child lies inside its parent's box
            if (hi > 255) { *error = 3; return; }
            node.qlo[a][sl] = uint8_t(lo);
            node.qhi[a][sl] = uint8_t(hi);
        }
        const int cnt = subtree_count(A, n, c);
        if (c < n - 1 && cnt > max_leaf) {
            node.imask = uint8_t(node.imask | (1u << sl));
            node.meta[sl] = uint8_t((1u << 5) | (24u + uint32_t(sl)));
            A.queue[level_in ^ 1][queue_base + rank] = make_uint2(child_base + rank, uint32_t(c));
            rank++;
        } else {
            node.meta[sl] = uint8_t((((1u << cnt) - 1u) << 5) | prim_offset);
            const int first = subtree_first(A, n, c);
            for (int q = 0; q < cnt; ++q) A.perm[prim_base + prim_offset + uint32_t(q)] = A.idx_sorted[first + q];
            prim_offset += uint32_t(cnt);
        }
    }
    nodes[me.x] = node;
}

__global__ void k_bvh_permute(const FlatPrim* src, FlatPrim* dst, const uint32_t* perm, uint32_t n)
{
    // 96-byte records as six 16-byte pieces: consecutive threads move consecutive pieces
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = t / 6u, piece = t % 6u;
    if (i >= n) return;
    reinterpret_cast<uint4*>(dst + i)[piece] = reinterpret_cast<const uint4*>(src + perm[i])[piece];
}

__global__ void k_bvh_offset_prim_base(FlatNode8* nodes, uint32_t n_nodes, uint32_t offset)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_nodes) nodes[i].prim_base += offset;
}

__global__ void k_bvh_init(Arrays A, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A.flag[i] = 0u;
    if (i == 0) {
        for (int a = 0; a < 3; ++a) { A.bounds[a] = 0x7fffffff; A.bounds[3 + a] = int(0x80000000u); }
        A.counters[0] = 1u; A.counters[1] = 0u; A.counters[2] = 0u;      // node 0 = the root
        A.queue[0][0] = make_uint2(0u, 0u);
    }
}

__global__ void k_bvh_next_level(Arrays A) { A.counters[2] = 0u; }

} // namespace bvhb

} // namespace artb

#endif
