// kernels.cu — the sm_100a kernels of the wavefront path tracer and the host code that launches them.
//
// Pipeline (one "wave" = every live path advances by one bounce), replacing the recursion of
// engine::_ray_color (engine.h:447-466) and the pixel/sample loops of engine::_run_single (engine.h:58-94):
//
//   k_shade      one warp = 32 entries of ONE queue: 32 hits of one material type (hit record, texture, emitted,
//                scatter) or 32 free slots that start the next (pixel, sample): jitter + camera::get_ray. Paths
//                that end (light, absorbed, depth 50) add their radiance and free the slot; scattered and fresh
//                rays are queued for the closest-hit stage
//   k_traverse   FP32 walk of the 8-wide tree (in shared memory), emits (ray, primitive) pairs        } the closest
//   k_pairs      one double-precision primitive test per thread, merged by a 128-bit compare-and-swap  } hit of every
//   k_finish     global surfaces, media, exact ties; misses add the background and free the slot,     } queued ray
//                hits are binned BY MATERIAL TYPE into the shade queues (warp-aggregated appends)
//   (k_extend    the pruned one-phase walk: overflow rays, and scenes whose tree is too big for shared memory)
//   k_resolve    write_color (color.h:6-22): mean, gamma 2, clamp, 8-bit
//
// Path state lives in HBM as ONE 128-byte record per pool slot (ray 64 B | bookkeeping 32 B | hit 16 B | ties 16 B).
// Five launches per wave on one stream; no kernel waits on the host and the queue counters are reset inside the
// kernels (see WaveCounters).
#include "renderer.h"
#include "rt_shade.cuh"
#include "bvh_build.cuh"
#include "adaptive_squares.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace artb {

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
    throw std::runtime_error(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

// Path state: ONE 128-byte line per pool slot — ray (64 B) | aux (32 B) | hit (16 B) | ties (16 B) — so that a kernel that
// needs the ray, the path's bookkeeping and its hit record touches exactly one aligned line (four sectors) per path.
// (As three separate arrays the 16- and 32-byte records dragged their neighbours' sectors along on every random
// access: ncu showed 320 B of DRAM reads per ray in k_finish for 116 B of payload.)
#ifndef ARTB_PATH_STRIDE
#define ARTB_PATH_STRIDE 128          // (256 / 512: probes of the address-translation reach — same lines touched, 2-4x the pages)
#endif
template <class T>
struct Strided {
    char* base; uint32_t stride;
    __host__ __device__ __forceinline__ T* at(uint32_t i) const { return reinterpret_cast<T*>(base + size_t(i) * stride); }
};
struct __align__(16) PathRay { double o[3], d[3], tm, pad; };                       // 64 B
struct __align__(16) PathAux { float thr[3]; uint32_t pixel, sample, rng_block, depth, pad; };   // 32 B
struct __align__(16) PathHit { double t; int32_t prim; uint32_t dfs; };             // 16 B: the best hit so far
struct __align__(16) PathTie { uint32_t n; uint32_t prim[3]; };                     // 16 B: other candidates with the SAME t (merge_hit)

// Queues written by k_extend (one per material type + the free-slot list) and their counters. Counters are
// double-buffered by wave parity so that each kernel can reset, from one thread, the counters that NEITHER it nor
// any concurrently running kernel touches — no separate housekeeping launches:
//   k_shade(cur)   consumes n_out[cur][*], appends to n_extend[cur^1] and n_out[cur^1][Q_REGEN];
//                  resets n_extend[cur], n_out[cur^1][0..MAT_COUNT) and the chunk cursor
//   k_extend(cur)  consumes n_extend[cur], appends to n_out[cur][*];
//                  resets n_out[cur^1][Q_REGEN] (after adding it to `issued`) and counts the rays
#define Q_REGEN MAT_COUNT
#define Q_OUT_COUNT (MAT_COUNT + 1)

struct WaveCounters {
    uint32_t n_extend[2];
    uint32_t n_out[2][Q_OUT_COUNT];
    uint32_t cursor[3];               // chunk cursors of k_traverse, (unused) and k_extend
    uint32_t n_overflow;              // rays whose candidate list overflowed in k_traverse (traced by k_extend)
    uint32_t n_list;                  // pixels in the current round's list
    uint32_t n_next_list;             // adaptive: pixels still unconverged
    unsigned long long issued, total, rays, finished, overflow_total;
};

#define ARTB_MAX_REGIONS 1024
#ifndef ARTB_PAIRS_PER_RAY
#define ARTB_PAIRS_PER_RAY 8         // capacity of the pair buffer per pool slot (mean over a wave; measured 3-4)
#endif
struct WaveBuffers {
    Strided<PathRay> rays;            // views into the pool of 128-byte path records
    Strided<PathAux> aux;
    Strided<PathHit> hits;
    Strided<PathTie> ties;
    uint32_t* q_extend[2];
    uint32_t* q_regen[2];
    uint32_t* q_shade;                // MAT_COUNT queues of `pool` entries, back to back
    uint32_t* q_overflow;
    uint2* pairs;                     // (ray, primitive) pairs, k_traverse -> k_pairs: pair_capacity entries
    uint32_t* region_count;           // ARTB_MAX_REGIONS counters
    size_t pair_capacity;
    WaveCounters* c;
    uint32_t pool;
};

struct RoundParams {
    CameraParams cam;
    int width, height, max_depth;
    uint32_t seed, stream;
    uint32_t first_sample;            // per-pixel index of the first sample of this round
    uint32_t batch;                   // samples per listed pixel in this round
    const uint32_t* pixel_list;       // nullptr: identity over width*height
    float* accum;                     // W*H*3 sums
    uint32_t* count;                  // per-pixel finished samples (adaptive) or nullptr
    float* lumsq;                     // per-pixel sum of squared sample luminance (adaptive) or nullptr
};

// ---- small device helpers -------------------------------------------------------------------------
__device__ __forceinline__ void queue_push(uint32_t* q, uint32_t* count, bool pred, uint32_t value, int lane)
{
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(count, uint32_t(__popc(m)));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) q[base + __popc(m & ((1u << lane) - 1u))] = value;
}

// Append to one of SEVERAL queues with a single round trip to the counters: lanes that target the same queue
// (same key) elect a leader, and all leaders issue their atomicAdd in the same instruction. key < 0: no append.
// Returns the lane's position in its queue.
__device__ __forceinline__ uint32_t multi_queue_reserve(int key, uint32_t* counter, int lane)
{
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (key >= 0 && lane == leader) base = atomicAdd(counter, uint32_t(__popc(peers)));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + uint32_t(__popc(peers & ((1u << lane) - 1u)));
}

// The same append split in two, so that the round trip of the atomic is not waited for: `reserve` elects the leaders
// and issues their atomicAdd; `commit` — called one loop iteration later — reads the result and writes the entry.
struct PendingAppend {
    uint32_t* queue;        // nullptr: nothing pending
    uint32_t value, rank, base_raw;
    int leader;
};

__device__ __forceinline__ void append_commit(PendingAppend& p)
{
    const uint32_t base = __shfl_sync(0xffffffffu, p.base_raw, p.leader);
    if (p.queue) p.queue[base + p.rank] = p.value;
    p.queue = nullptr;
}

__device__ __forceinline__ void append_reserve(PendingAppend& p, int key, uint32_t* counter, uint32_t* queue, uint32_t value, int lane)
{
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    p.leader = __ffs(peers) - 1;
    p.rank = uint32_t(__popc(peers & ((1u << lane) - 1u)));
    p.base_raw = 0;
    if (key >= 0 && lane == p.leader) p.base_raw = atomicAdd(counter, uint32_t(__popc(peers)));
    p.queue = key >= 0 ? queue : nullptr;
    p.value = value;
}

__device__ __forceinline__ DRay load_ray(const PathRay* ray)
{
    const double2* p = reinterpret_cast<const double2*>(ray);
    const double2 a = p[0], b = p[1], c = p[2], d = p[3];
    DRay r;
    r.o = mk3(a.x, a.y, b.x);
    r.d = mk3(b.y, c.x, c.y);
    r.tm = d.x;
    return r;
}

__device__ __forceinline__ void store_ray(PathRay* ray, const DRay& r)
{
    double2* p = reinterpret_cast<double2*>(ray);
    p[0] = make_double2(r.o.x, r.o.y);
    p[1] = make_double2(r.o.z, r.d.x);
    p[2] = make_double2(r.d.y, r.d.z);
    p[3] = make_double2(r.tm, 0.0);
}

__device__ __forceinline__ void load_aux(const PathAux* aux, PathAux& a)
{
    *reinterpret_cast<uint4*>(&a) = *reinterpret_cast<const uint4*>(aux);
    *(reinterpret_cast<uint4*>(&a) + 1) = *(reinterpret_cast<const uint4*>(aux) + 1);
}

__device__ __forceinline__ void store_aux(PathAux* aux, const PathAux& a)
{
    *reinterpret_cast<uint4*>(aux) = *reinterpret_cast<const uint4*>(&a);
    *(reinterpret_cast<uint4*>(aux) + 1) = *(reinterpret_cast<const uint4*>(&a) + 1);
}

__device__ __forceinline__ void add_radiance(const RoundParams& rp, uint32_t pixel, float r, float g, float b)
{
    float* a = rp.accum + size_t(pixel) * 3;
    if (r != 0.f) atomicAdd(a + 0, r);
    if (g != 0.f) atomicAdd(a + 1, g);
    if (b != 0.f) atomicAdd(a + 2, b);
    if (rp.count) {
        atomicAdd(rp.count + pixel, 1u);
        const float y = 0.2126f * r + 0.7152f * g + 0.0722f * b;
        if (y != 0.f) atomicAdd(rp.lumsq + pixel, y * y);
    }
}

// ---- kernels ---------------------------------------------------------------------------------------
__global__ void k_init_pool(WaveBuffers w)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.pool) w.q_regen[0][i] = i;
    if (i == 0) {
        WaveCounters* c = w.c;
        c->n_extend[0] = c->n_extend[1] = 0;
        for (int k = 0; k < 2; ++k)
            for (int m = 0; m < Q_OUT_COUNT; ++m) c->n_out[k][m] = 0;
        c->n_out[0][Q_REGEN] = w.pool;
        c->cursor[0] = c->cursor[1] = c->cursor[2] = 0; c->n_overflow = 0; c->n_list = 0; c->n_next_list = 0;
        c->issued = c->total = c->rays = c->finished = c->overflow_total = 0;
    }
}

// n_list_fixed >= 0: use it; otherwise take the list length the adaptive update left in n_next_list
__global__ void k_round_begin(WaveCounters* c, long long n_list_fixed, uint32_t batch)
{
    const uint32_t n = n_list_fixed >= 0 ? uint32_t(n_list_fixed) : c->n_next_list;
    c->n_list = n;
    c->n_next_list = 0;
    c->issued = 0;
    c->total = (unsigned long long)n * batch;
}

// =====================================================================================================
// The closest-hit stage ("extend"), as three kernels + an overflow pass (rt_intersect.cuh "two-phase closest hit"):
//
//   k_traverse   FP32 only: walks the 8-wide tree (staged in shared memory when it fits) with the float image of
//                the ray and appends a (ray, primitive) PAIR for every primitive whose leaf box the ray enters.
//                Nothing a lane finds changes what it visits next, so a finished lane is REFILLED with the next
//                queued ray while its neighbours keep walking: the warp stays full although per-ray node counts
//                vary 1..40. Pairs go to one region of the pair buffer per block (a shared-memory counter hands out
//                positions, one atomic per warp and node step), so no global counter is contended.
//   k_pairs      one thread per PAIR: the reference's double-precision test of that primitive against that ray; a
//                hit is merged into the ray's 16-byte record {t, prim, dfs} with a 128-bit compare-and-swap under
//                the reference's rule (smaller t; equal t: later in the reference's depth-first order). A ray with
//                12 candidates and a ray with one keep the same number of lanes busy.
//   k_finish     one thread per ray: the global surfaces, then the media (closest_hit() has the reasons for this
//                order), then misses add the background and free the slot while hits are binned BY MATERIAL TYPE
//                into the shade queues (warp-aggregated appends).
//   k_extend     the one-phase walk (closest_hit) for rays whose pairs did not fit their block's region, and for
//                scenes whose tree is too big for shared memory (there the pruned walk visits far fewer nodes).
//
// The same kernels serve artb_trace_batch (kBatch: explicit rays in, HitRecordOut out), so the parity tests
// exercise the code the renderer runs.
// =====================================================================================================
// A (ray, primitive) pair in 8 bytes: ray id 24 bits | primitive index 24 bits | 16 bits of the FP32 parameter at
// which the ray enters the primitive's leaf box, truncated (i.e. rounded DOWN: it stays a lower bound of the hit's t).
// k_pairs drops a pair whose bound already exceeds the ray's best hit: nothing is loaded or tested for it.
#define ARTB_PAIR_MAX_ID (1u << 24)
__device__ __forceinline__ uint2 pack_pair(uint32_t ray, uint32_t prim, float tn)
{
    uint32_t tb = __float_as_uint(tn) >> 16;          // bf16 by truncation: towards zero = down for tn >= 0
    if (!(tn >= 0.f)) tb = 0xff80u;                   // negative or NaN: -inf, never culls
    return make_uint2(ray | (prim << 24), (prim >> 8) | (tb << 16));
}
__device__ __forceinline__ void unpack_pair(uint2 p, uint32_t& ray, uint32_t& prim, float& tn)
{
    ray = p.x & 0xffffffu;
    prim = (p.x >> 24) | ((p.y & 0xffffu) << 8);
    tn = __uint_as_float(p.y & 0xffff0000u);
}

struct PairBuf {
    uint2* pairs;                     // n_regions * region_cap packed pairs (pack_pair); x == ~0: void
    uint32_t* region_count;           // entries used in every region (written by k_traverse)
    uint32_t region_cap, n_regions;
};

struct TraceIO {
    const uint32_t* n_ptr;            // number of rays of this launch (device memory)
    const uint32_t* list;             // ray ids (path slots); nullptr: 0..n-1
    Strided<PathTie> ties;            // one record per ray id: candidates whose t equals the record's (exact ties)
    Strided<PathHit> hits;            // one record per ray id: the best hit so far; prim == ARTB_HIT_OVERFLOW: some of the
                                      // ray's pairs did not fit, the overflow pass traces it
    PairBuf pb;
    uint32_t* cursor_traverse;
    uint32_t* q_overflow;
    uint32_t* n_overflow;
    double t_min, t_max;
    Strided<PathRay> rays;            // wave mode: the path pool
    Strided<PathAux> aux;
    const double* batch_rays;         // batch mode: n * 7 doubles
    const double* batch_xi;           // batch mode: n * n_media explicit medium draws, or nullptr
    HitRecordOut* batch_out;
};

template <bool kBatch>
__device__ __forceinline__ DRay trace_io_ray(const TraceIO& io, uint32_t id)
{
    if (kBatch) {
        const double* p = io.batch_rays + size_t(id) * 7;
        DRay r;
        r.o = ld3(p); r.d = ld3(p + 3); r.tm = p[6];
        return r;
    }
    return load_ray(io.rays.at(id));
}

__device__ __forceinline__ PathHit load_hit(const PathHit* h)
{
    const double2 v = *reinterpret_cast<const double2*>(h);
    PathHit r;
    *reinterpret_cast<double2*>(&r) = v;
    return r;
}

#define ARTB_HIT_OVERFLOW (-2)
__device__ __forceinline__ void mark_overflow(PathHit* h) { h->prim = ARTB_HIT_OVERFLOW; }

__device__ __forceinline__ void store_hit(PathHit* h, double t, int32_t prim, uint32_t dfs)
{
    PathHit v; v.t = t; v.prim = prim; v.dfs = dfs;
    *reinterpret_cast<double2*>(h) = *reinterpret_cast<const double2*>(&v);
}

// Merge a candidate into a ray's record: 16-byte compare-and-swap loop (ATOMG.CAS.128). Records only ever improve,
// so a candidate that does not beat a (possibly stale) copy cannot beat the current value either — which is why the
// caller may read that copy long before the candidate is known (peek_hit next to the primitive and ray loads).
__device__ __forceinline__ void peek_hit(const PathHit* h, unsigned long long& lo, unsigned long long& hi)
{
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(h) : "memory");
}

__device__ __forceinline__ bool cas_hit(PathHit* h, unsigned long long& cur_lo, unsigned long long& cur_hi, unsigned long long new_lo,
                                        unsigned long long new_hi)
{
    unsigned long long old_lo, old_hi;
    asm volatile("{\n\t.reg .b128 c, s, d;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 s, {%4, %5};\n\t"
                 "atom.global.relaxed.gpu.cas.b128 d, [%6], c, s;\n\tmov.b128 {%0, %1}, d;\n\t}"
                 : "=l"(old_lo), "=l"(old_hi) : "l"(cur_lo), "l"(cur_hi), "l"(new_lo), "l"(new_hi), "l"(h) : "memory");
    const bool done = old_lo == cur_lo && old_hi == cur_hi;
    cur_lo = old_lo; cur_hi = old_hi;
    return done;
}

// Merge a candidate into a ray's record (k_pairs). An exact tie (the candidate's t EQUALS the record's) is not decided
// here: the reference's rule for it needs the boxes of its own bvh_nodes and the world-space ray (rt_intersect.cuh:
// hit_replaces), which would cost this kernel a call and a stack frame for something that happens a few times per
// million rays. The record keeps the later leaf and gets ARTB_HIT_TIED set, the other leaf goes to the ray's tie list;
// k_finish — which has the ray — replays the tied leaves in the reference's order.
#define ARTB_HIT_TIED 0x80000000u     // bit 31 of PathHit::dfs
template <bool kBatch>
__device__ __forceinline__ void merge_hit(const TraceIO& io, uint32_t ray_id, unsigned long long cur_lo, unsigned long long cur_hi,
                                          double t, uint32_t prim, uint32_t dfs)
{
    PathHit* h = io.hits.at(ray_id);
    for (;;) {
        unsigned long long new_lo = (unsigned long long)__double_as_longlong(t);
        unsigned long long new_hi = (unsigned long long)prim | ((unsigned long long)dfs << 32);
        const double cur_t = __longlong_as_double((long long)cur_lo);
        const int32_t cur_prim = int32_t(uint32_t(cur_hi));
        if (cur_prim == ARTB_HIT_OVERFLOW) return;
        uint32_t other = 0xffffffffu;                             // the leaf that does NOT stay in the record on a tie
        if (cur_prim >= 0) {
            if (t > cur_t) return;
            if (t == cur_t) {
                const uint32_t cur_dfs = uint32_t(cur_hi >> 32);
                if ((cur_dfs & ~ARTB_HIT_TIED) >= dfs) { new_lo = cur_lo; new_hi = cur_hi; other = prim; }
                else other = uint32_t(cur_prim);
                new_hi |= (unsigned long long)ARTB_HIT_TIED << 32;
            }
        } else if (!(t <= cur_t)) return;
        if (cas_hit(h, cur_lo, cur_hi, new_lo, new_hi)) {
            if (other != 0xffffffffu) {
                PathTie* tie = io.ties.at(ray_id);
                const uint32_t k = atomicAdd(&tie->n, 1u);
                if (k < 3u) tie->prim[k] = other;
            }
            return;
        }
    }
}

#ifndef ARTB_TRAVERSE_THREADS
#define ARTB_TRAVERSE_THREADS 1024
#endif
#ifndef ARTB_TRAVERSE_BLOCKS
#define ARTB_TRAVERSE_BLOCKS 1        // resident blocks per SM (registers: 65536 / threads / blocks)
#endif
#ifndef ARTB_TRAVERSE_GRID_PER_SM
#define ARTB_TRAVERSE_GRID_PER_SM ARTB_TRAVERSE_BLOCKS   // blocks launched per SM (< ARTB_TRAVERSE_BLOCKS leaves room for other streams)
#endif
#ifndef ARTB_EMIT_LANE_ATOMICS
#define ARTB_EMIT_LANE_ATOMICS 1      // measured: -7 % on k_traverse against a warp scan + one atomic per warp
#endif
#ifndef ARTB_TRAVERSE_CLAIM
#define ARTB_TRAVERSE_CLAIM 32        // rays a warp claims from the queue per refill (<= 32, the width of its feed)
#endif
#ifndef ARTB_TRAVERSE_POP_LAST
#define ARTB_TRAVERSE_POP_LAST 1      // 1: a lane learns that its walk has ended at the end of an iteration (refilled before the next node test)
#endif
extern __shared__ uint4 smem_nodes[];
struct WarpFeed { float inv[3][32], c[3][32]; uint32_t octinv4[32], id[32]; };      // 1 KB per warp

// The traversal stack of k_traverse. The collector pushes at most one entry per level of 8-wide nodes, and the trees of
// the reference's scenes are 2-8 levels deep, so the stack fits in SHARED memory next to the node array ([entry][thread]:
// a warp's accesses to one entry are 32 consecutive words pairs, no bank conflicts). As a lane-private array it lived in
// local memory: an LDL / STL round trip through L1 on every pop and push, 7 % of the kernel's stall samples.
#ifndef ARTB_SMEM_STACK
#define ARTB_SMEM_STACK 0             // entries (0 = lane-private local-memory stack). 8 was measured: the same speed
#endif
template <int kEntries>
struct TraverseStack {
    uint2 local[kEntries ? 1 : ARTB_STACK_DEPTH];
    uint2* shared;
    __device__ __forceinline__ uint2& operator[](int i)
    {
        if (kEntries) return shared[uint32_t(i) * ARTB_TRAVERSE_THREADS];
        return local[i];
    }
    __device__ __forceinline__ const uint2& operator[](int i) const
    {
        if (kEntries) return shared[uint32_t(i) * ARTB_TRAVERSE_THREADS];
        return local[i];
    }
};

// kShared: every block first stages the whole 8-wide node array (a few hundred 80-byte nodes for the reference's
// scenes) in shared memory with 16-byte loads and traverses it there; one block per SM, so one copy per SM.
// Otherwise nodes are read through L1/L2.
template <bool kShared, bool kBatch, int kSmemStack>
__global__ void __launch_bounds__(kShared ? ARTB_TRAVERSE_THREADS : 256, kShared ? ARTB_TRAVERSE_BLOCKS : 3)
k_traverse(TraceIO io, SceneView s, WaveCounters* c, int cur)
{
    static_assert(!kSmemStack || kShared, "the shared-memory stack comes with the shared-memory node array");
    __shared__ uint32_t s_pairs;      // pairs this block has appended to its region
    __shared__ WarpFeed feeds[(kShared ? ARTB_TRAVERSE_THREADS : 256) / 32];
    const uint32_t n = *io.n_ptr;
    if (!kBatch && blockIdx.x == 0 && threadIdx.x == 0) {
        // the free list k_shade(cur^1) just consumed: account the samples it started, then clear it
        const unsigned long long left = c->total - c->issued;
        const unsigned long long nr = c->n_out[cur ^ 1][Q_REGEN];
        c->issued += nr < left ? nr : left;
        c->n_out[cur ^ 1][Q_REGEN] = 0;
        c->rays += n;
    }
    if (n == 0) {
        if (threadIdx.x == 0) io.pb.region_count[blockIdx.x] = 0;
        return;
    }
    if (threadIdx.x == 0) s_pairs = 0;
    const FlatNode8* nodes = s.nodes;
    TraverseStack<kSmemStack> stack;
    stack.shared = nullptr;
    if (kShared) {
        const uint4* src = reinterpret_cast<const uint4*>(s.nodes);
        for (uint32_t i = threadIdx.x; i < s.n_nodes * 5u; i += blockDim.x) smem_nodes[i] = __ldg(src + i);
        nodes = reinterpret_cast<const FlatNode8*>(smem_nodes);
        if (kSmemStack) stack.shared = reinterpret_cast<uint2*>(smem_nodes + (s.n_nodes ? s.n_nodes * 5u : 1u)) + threadIdx.x;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool has_tree = s.n_world_prims > s.n_global_prims;
    const float t_min_f = d2f_rd(io.t_min), t_max_f = d2f_ru(io.t_max);
    const uint32_t cap = io.pb.region_cap;
    uint2* region = io.pb.pairs + size_t(blockIdx.x) * cap;
    BoxRay br;
    Collector col;
    uint32_t id = 0;
    bool active = false, spilled = false;
    bool more = true;                 // warp-uniform: the queue may still hold rays
    collect_begin(col, false);
    br.inv[0] = br.inv[1] = br.inv[2] = 1.f; br.c[0] = br.c[1] = br.c[2] = 0.f; br.octinv4 = 0u;
    // The warp's FEED: 32 rays fetched, converted to their FP32 box-test constants and parked in shared memory by ALL
    // lanes together (full lane efficiency for the per-ray setup); a lane whose ray ends pops the next one with a
    // handful of shared loads, so lanes are refilled one by one, at once, while their neighbours keep walking.
    WarpFeed& feed = feeds[threadIdx.x >> 5];
    uint32_t feed_count = 0;          // warp-uniform: entries not handed out yet (taken from the top)
    for (;;) {
        uint32_t m_active = __ballot_sync(0xffffffffu, active);
        if (m_active != 0xffffffffu) {
            if (feed_count == 0u && more) {
                // (claiming 64 or 128 rays per atomic measured 1-4 % slower: the tail of the kernel gets less even)
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(io.cursor_traverse, uint32_t(ARTB_TRAVERSE_CLAIM));
                base = __shfl_sync(0xffffffffu, base, 0);
                const uint32_t got = base < n ? (n - base < uint32_t(ARTB_TRAVERSE_CLAIM) ? n - base : uint32_t(ARTB_TRAVERSE_CLAIM)) : 0u;
                if (base + uint32_t(ARTB_TRAVERSE_CLAIM) >= n) more = false;
                if (uint32_t(lane) < got) {
                    const uint32_t idx = base + uint32_t(lane);
                    const uint32_t rid = io.list ? io.list[idx] : idx;
                    const BoxRay b = make_box_ray(make_trav_ray(trace_io_ray<kBatch>(io, rid)));
                    feed.inv[0][lane] = b.inv[0]; feed.inv[1][lane] = b.inv[1]; feed.inv[2][lane] = b.inv[2];
                    feed.c[0][lane] = b.c[0]; feed.c[1][lane] = b.c[1]; feed.c[2][lane] = b.c[2];
                    feed.octinv4[lane] = b.octinv4;
                    feed.id[lane] = rid;
                    store_hit(io.hits.at(rid), io.t_max, -1, 0u);
                    io.ties.at(rid)->n = 0u;
                }
                __syncwarp();
                feed_count = got;
            }
            if (feed_count) {
                const uint32_t idle = ~m_active;
                const uint32_t n_idle = uint32_t(__popc(idle));
                const uint32_t take = n_idle < feed_count ? n_idle : feed_count;
                const uint32_t rank = uint32_t(__popc(idle & lt_mask));
                if (!active && rank < take) {
                    const uint32_t e = feed_count - 1u - rank;
                    br.inv[0] = feed.inv[0][e]; br.inv[1] = feed.inv[1][e]; br.inv[2] = feed.inv[2][e];
                    br.c[0] = feed.c[0][e]; br.c[1] = feed.c[1][e]; br.c[2] = feed.c[2][e];
                    br.octinv4 = feed.octinv4[e];
                    id = feed.id[e];
                    collect_begin(col, has_tree);
                    active = true;
                    spilled = false;
                }
                feed_count -= take;
                __syncwarp();
                m_active = __ballot_sync(0xffffffffu, active);
            }
            if (m_active == 0u) {
                if (feed_count == 0u && !more) break;
                continue;
            }
        }
        uint32_t prim_base = 0, prim_mask = 0;
        float tn_min = 0.f;
#if !ARTB_TRAVERSE_POP_LAST
        if (active && !collect_pop(col, stack)) {
            if (spilled) mark_overflow(io.hits.at(id));
            active = false;
        }
        warp_sync();
#endif
        // (ARTB_TRAVERSE_POP_LAST: every active lane arrives here with a child to enter — the pop that finds the walk ended
        // happens at the END of the previous iteration, so the lane is refilled before this node test instead of idling
        // through it: one wasted node-test slot per ray less)
        if (active && (!ARTB_TRAVERSE_POP_LAST || col.ng.y > 0x00ffffffu))
            collect_node(nodes, br, t_min_f, t_max_f, col, stack, prim_base, prim_mask, tn_min, s.one_bits, nullptr,
                         kSmemStack ? kSmemStack : ARTB_STACK_DEPTH);
        warp_sync();
        // append this step's pairs
        const uint32_t cnt = uint32_t(__popc(prim_mask));
#if ARTB_EMIT_LANE_ATOMICS
        // one shared-memory atomic per lane that found something (a third of the lanes per step)
        if (cnt) {
            uint32_t pos = atomicAdd(&s_pairs, cnt);
#else
        // positions from a warp-wide scan + one shared-memory atomic per warp
        if (__any_sync(0xffffffffu, cnt != 0u)) {
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            uint32_t wbase = 0;
            if (lane == 31) wbase = atomicAdd(&s_pairs, incl);
            wbase = __shfl_sync(0xffffffffu, wbase, 31);
            uint32_t pos = wbase + incl - cnt;
#endif
            if (pos + cnt <= cap) {
                const uint32_t tn_bits = pack_pair(0u, 0u, tn_min).y;       // the same bound for every pair of this step
                while (prim_mask) {
                    const int b = find_msb(prim_mask);
                    prim_mask &= ~(1u << b);
                    const uint32_t prim = prim_base + uint32_t(b);
                    region[pos++] = make_uint2(id | (prim << 24), (prim >> 8) | tn_bits);
                }
            } else if (cnt) {
                spilled = true;                   // region full: void what is left of it, the overflow pass traces the ray
                for (uint32_t k = 0; k < cnt && pos + k < cap; ++k) region[pos + k] = make_uint2(0xffffffffu, 0u);
            }
        }
        warp_sync();
#if ARTB_TRAVERSE_POP_LAST
        if (active && !collect_pop(col, stack)) {     // nothing pending and the stack is empty: this ray is done
            if (spilled) mark_overflow(io.hits.at(id));
            active = false;
        }
        warp_sync();
#endif
    }
    __syncthreads();
    if (threadIdx.x == 0) io.pb.region_count[blockIdx.x] = s_pairs < cap ? s_pairs : cap;
}

// k_walk — the closest TREE hit for scenes whose node array does not fit in shared memory (a million triangles): the
// pruned, vote-based walk of closest_hit() (trav_round: FP32 box tests, double-precision primitive tests, pruning by the
// best hit so far — on big trees pruning saves more than the two-phase split gains), but with the per-lane refill of
// k_traverse: a lane whose ray ends stores its 16-byte result and pops the next ray from the warp's feed, so the warp
// stays full. The global surfaces, the media and the binning are left to k_finish, exactly as after k_pairs.
struct WalkFeed { double ray[7][32]; float inv[3][32], c[3][32]; uint32_t octinv4[32], id[32]; };      // 2.75 KB per warp
#define ARTB_WALK_THREADS 128
#ifndef ARTB_WALK_BLOCKS
#define ARTB_WALK_BLOCKS 5             // 102 registers: measured +5.7 % over 4 blocks (128), 3 and 6 are slower
#endif
#ifndef ARTB_WALK_COOP
#define ARTB_WALK_COOP 1               // the primitive step of k_walk is shared by the whole warp (walk_round)
#endif
#ifndef ARTB_WALK_COOP_MIN
#define ARTB_WALK_COOP_MIN 24          // pending (ray, primitive) candidates in the warp that start a primitive step
#endif
#ifndef ARTB_WALK_COOP_ITEMS
#define ARTB_WALK_COOP_ITEMS 256
#endif

// The warp's share of shared memory for the cooperative primitive step: the double-precision ray of every lane (node steps
// only need the float BoxRay, so the lanes do not keep it in registers) and the list of pending (lane, primitive) candidates.
struct WalkCoop { double ray[7][32]; uint32_t item[ARTB_WALK_COOP_ITEMS]; };                            // 2.75 KB per warp

__device__ __forceinline__ DRay walk_coop_ray(const WalkCoop& co, uint32_t lane)
{
    DRay r;
    r.o = mk3(co.ray[0][lane], co.ray[1][lane], co.ray[2][lane]);
    r.d = mk3(co.ray[3][lane], co.ray[4][lane], co.ray[5][lane]);
    r.tm = co.ray[6][lane];
    return r;
}

// trav_round for k_walk (full warps, rays in shared memory). The node step is trav_round's. The PRIMITIVE step is taken by
// the warp as a whole: in trav_round every lane tests its own candidates one after the other, so the double-precision
// tests ran with the 9 lanes of 32 that happened to have a candidate left (profiles/r02_lines_k_walk.txt: more than half of
// the kernel's samples). Here the pending candidates of all lanes are listed in shared memory and dealt out 32 at a time —
// lane j tests candidate j against the ray of the lane that owns it, then the hits (few) are handed back to their owners
// one by one through shuffles and merged with the same rule (hit_replaces, which does not depend on the order). A step
// starts when ARTB_WALK_COOP_MIN candidates wait or no lane can take a node step.
__device__ __forceinline__ void walk_round(const SceneView& s, Traversal& tv, uint2* stack, const BoxRay& br, WalkCoop& co,
                                           double t_min, float t_min_f)
{
    const uint32_t full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t m_alive = __ballot_sync(full, tv.alive);
    if (m_alive == 0u) return;
    const bool want_prims = tv.alive && tv.pg.y != 0u;
    const uint32_t pend = want_prims ? tv.pg.y : 0u;
    const uint32_t cnt = uint32_t(__popc(pend));
    const uint32_t m_prims = __ballot_sync(full, want_prims);
    const uint32_t total = __reduce_add_sync(full, cnt);
    if (total >= uint32_t(ARTB_WALK_COOP_MIN) || (m_prims == m_alive && total != 0u)) {
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(full, incl, d);
            if (int(lane) >= d) incl += v;
        }
        uint32_t k = incl - cnt;
        uint32_t left = pend;
        while (left != 0u && k < uint32_t(ARTB_WALK_COOP_ITEMS)) {
            const int bit = find_msb(left);
            left &= ~(1u << bit);
            co.item[k++] = (lane << 27) | (tv.pg.x + uint32_t(bit));
        }
        if (want_prims) tv.pg.y = left;                   // candidates beyond the list wait for the next step
        __syncwarp();
        const uint32_t n_items = total < uint32_t(ARTB_WALK_COOP_ITEMS) ? total : uint32_t(ARTB_WALK_COOP_ITEMS);
        for (uint32_t base = 0; base < n_items; base += 32u) {
            const uint32_t j = base + lane;
            const bool valid = j < n_items;
            const uint32_t item = valid ? co.item[j] : (lane << 27);
            const uint32_t owner = item >> 27;
            ClosestHit cand;
            cand.t = __shfl_sync(full, tv.best.t, int(owner));
            cand.prim = -1; cand.dfs = 0;
            if (valid) test_surface(s, item & 0x07ffffffu, walk_coop_ray(co, owner), t_min, cand);
            uint32_t m_hit = __ballot_sync(full, valid && cand.prim >= 0);
            while (m_hit != 0u) {
                const int src = __ffs(int(m_hit)) - 1;
                m_hit &= m_hit - 1u;
                const uint32_t to = __shfl_sync(full, owner, src);
                const double ht = __shfl_sync(full, cand.t, src);
                const int32_t hp = __shfl_sync(full, cand.prim, src);
                const uint32_t hd = __shfl_sync(full, cand.dfs, src);
                if (lane == to) {
                    bool take;
                    if (ht < tv.best.t || tv.best.prim < 0) take = ht <= tv.best.t;
                    else if (ht != tv.best.t) take = false;
                    else take = tie_replaces(s.self, walk_coop_ray(co, lane), t_min, ht, uint32_t(hp), hd, uint32_t(tv.best.prim), tv.best.dfs);
                    if (take) { tv.best.t = ht; tv.best.prim = hp; tv.best.dfs = hd; }
                }
            }
        }
        __syncwarp();
        if (want_prims) tv.t_best_f = d2f_ru(tv.best.t);
    } else if (tv.alive && !want_prims && tv.ng.y > 0x00ffffffu) {
        const uint32_t hits = tv.ng.y;
        const int bit = find_msb(hits);
        tv.ng.y &= ~(1u << bit);
        if (tv.ng.y > 0x00ffffffu && tv.sp < ARTB_STACK_DEPTH) stack[tv.sp++] = tv.ng;
        const uint32_t slot = uint32_t(bit - 24) ^ (br.octinv4 & 0xffu);
        const uint32_t rel = uint32_t(__popc(hits & ~(0xffffffffu << slot) & 0xffu));
        const uint4* q = reinterpret_cast<const uint4*>(s.nodes + (tv.ng.x + rel));
        const uint4 n0 = ld_node16(q), n1 = ld_node16(q + 1), n2 = ld_node16(q + 2), n3 = ld_node16(q + 3), n4 = ld_node16(q + 4);
        float tn_unused = 0.f;
        const uint32_t hitmask = node8_hits(n0, n1, n2, n3, n4, br, t_min_f, tv.t_best_f, s.one_bits, tn_unused);
        tv.ng.x = n1.x;
        tv.ng.y = (hitmask & 0xff000000u) | (n0.w >> 24);
        tv.pg.x = n1.y;
        tv.pg.y = hitmask & 0x00ffffffu;
        if (tv.pg.y) {
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
}
template <bool kBatch>
__global__ void __launch_bounds__(ARTB_WALK_THREADS, ARTB_WALK_BLOCKS)
k_walk(TraceIO io, SceneView s, WaveCounters* c, int cur)
{
    __shared__ WalkFeed feeds[ARTB_WALK_THREADS / 32];
#if ARTB_WALK_COOP
    __shared__ WalkCoop coops[ARTB_WALK_THREADS / 32];
    WalkCoop& coop = coops[threadIdx.x >> 5];
#endif
    const uint32_t n = *io.n_ptr;
    if (!kBatch && blockIdx.x == 0 && threadIdx.x == 0) {
        // the free list k_shade(cur^1) just consumed: account the samples it started, then clear it
        const unsigned long long left = c->total - c->issued;
        const unsigned long long nr = c->n_out[cur ^ 1][Q_REGEN];
        c->issued += nr < left ? nr : left;
        c->n_out[cur ^ 1][Q_REGEN] = 0;
        c->rays += n;
    }
    if (n == 0) return;
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool has_tree = s.n_world_prims > s.n_global_prims;
    const float t_min_f = d2f_rd(io.t_min);
    WalkFeed& feed = feeds[threadIdx.x >> 5];
    uint32_t feed_count = 0;          // warp-uniform
    bool more = true;                 // warp-uniform
    uint2 stack[ARTB_STACK_DEPTH];      // (in shared memory, [entry][thread]: measured 2 % slower)
    Traversal tv;
    trav_begin(tv, io.t_max, false);
    tv.alive = false;
#if ARTB_WALK_COOP
    BoxRay br;
    br.inv[0] = br.inv[1] = br.inv[2] = 1.f; br.c[0] = br.c[1] = br.c[2] = 0.f; br.octinv4 = 0x07070707u;
#else
    DRay ray;
    ray.o = mk3(0, 0, 0); ray.d = mk3(1, 0, 0); ray.tm = 0;
    BoxRay br = make_box_ray(ray);
#endif
    uint32_t id = 0;
    bool has = false;                 // this lane holds a ray whose result is not stored yet
    for (;;) {
        uint32_t m_alive = __ballot_sync(0xffffffffu, tv.alive);
        if (m_alive != 0xffffffffu) {
            if (has && !tv.alive) {
                store_hit(io.hits.at(id), tv.best.t, tv.best.prim, tv.best.dfs);
                has = false;
            }
            if (feed_count == 0u && more) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(io.cursor_traverse, 32u);
                base = __shfl_sync(0xffffffffu, base, 0);
                const uint32_t got = base < n ? (n - base < 32u ? n - base : 32u) : 0u;
                if (uint32_t(lane) < got) {
                    const uint32_t idx = base + uint32_t(lane);
                    const uint32_t rid = io.list ? io.list[idx] : idx;
                    const DRay r = trace_io_ray<kBatch>(io, rid);
                    const BoxRay b = make_box_ray(r);
                    feed.ray[0][lane] = r.o.x; feed.ray[1][lane] = r.o.y; feed.ray[2][lane] = r.o.z;
                    feed.ray[3][lane] = r.d.x; feed.ray[4][lane] = r.d.y; feed.ray[5][lane] = r.d.z; feed.ray[6][lane] = r.tm;
                    feed.inv[0][lane] = b.inv[0]; feed.inv[1][lane] = b.inv[1]; feed.inv[2][lane] = b.inv[2];
                    feed.c[0][lane] = b.c[0]; feed.c[1][lane] = b.c[1]; feed.c[2][lane] = b.c[2];
                    feed.octinv4[lane] = b.octinv4;
                    feed.id[lane] = rid;
                }
                __syncwarp();
                feed_count = got;
                if (base + 32u >= n) more = false;
            }
            if (feed_count) {
                const uint32_t idle = ~m_alive;
                const uint32_t n_idle = uint32_t(__popc(idle));
                const uint32_t take = n_idle < feed_count ? n_idle : feed_count;
                const uint32_t rank = uint32_t(__popc(idle & lt_mask));
                if (!tv.alive && rank < take) {
                    const uint32_t e = feed_count - 1u - rank;
#if ARTB_WALK_COOP
#pragma unroll
                    for (int k = 0; k < 7; ++k) coop.ray[k][lane] = feed.ray[k][e];
#else
                    ray.o = mk3(feed.ray[0][e], feed.ray[1][e], feed.ray[2][e]);
                    ray.d = mk3(feed.ray[3][e], feed.ray[4][e], feed.ray[5][e]);
                    ray.tm = feed.ray[6][e];
#endif
                    br.inv[0] = feed.inv[0][e]; br.inv[1] = feed.inv[1][e]; br.inv[2] = feed.inv[2][e];
                    br.c[0] = feed.c[0][e]; br.c[1] = feed.c[1][e]; br.c[2] = feed.c[2][e];
                    br.octinv4 = feed.octinv4[e];
                    id = feed.id[e];
                    trav_begin(tv, io.t_max, has_tree);
                    has = true;
                }
                feed_count -= take;
                __syncwarp();
                m_alive = __ballot_sync(0xffffffffu, tv.alive);
            }
            if (m_alive == 0u) {
                if (feed_count == 0u && !more) break;
                continue;
            }
        }
#if ARTB_WALK_COOP
        walk_round(s, tv, stack, br, coop, io.t_min, t_min_f);
#else
        trav_round(s, s.nodes, 0xffffffffu, tv, stack, ray, br, io.t_min, t_min_f, nullptr);
#endif
    }
}

#ifndef ARTB_PAIRS_THREADS
#define ARTB_PAIRS_THREADS 64         // small blocks: the three barriers of the tile sort stall fewer warps (measured -4 %)
#endif
#ifndef ARTB_PAIRS_BLOCKS
#define ARTB_PAIRS_BLOCKS 16
#endif
#ifndef ARTB_PAIRS_PREFETCH
#define ARTB_PAIRS_PREFETCH 0           // measured: prefetching ray / primitive lines during the sort costs 10 % (L1 thrash)
#endif
#define ARTB_PAIRS_PER_THREAD 4       // tile = ARTB_PAIRS_THREADS * 4 pairs, sorted by primitive class in shared memory (8: -17 %)
template <bool kBatch>
__global__ void __launch_bounds__(ARTB_PAIRS_THREADS, ARTB_PAIRS_BLOCKS)
k_pairs(TraceIO io, SceneView s)
{
    constexpr uint32_t kTile = ARTB_PAIRS_THREADS * ARTB_PAIRS_PER_THREAD;
    __shared__ uint2 tile[kTile];
    __shared__ uint32_t class_count[ARTB_PRIM_CLASSES + 1], class_start[ARTB_PRIM_CLASSES + 1];
    const uint32_t nr = io.pb.n_regions;
    const uint32_t r = blockIdx.x % nr, j = blockIdx.x / nr, J = gridDim.x / nr;
    if (j >= J) return;
    const uint32_t cnt = io.pb.region_count[r];
    const uint2* region = io.pb.pairs + size_t(r) * io.pb.region_cap;
    // (Tiles are dealt to the region's blocks round-robin. One contiguous share per block would process a ray's pairs
    // more strictly in emission order and cull more of them, but measured 12 % slower: the shares' live work is uneven.)
    for (uint32_t t0 = j * kTile; t0 < cnt; t0 += J * kTile) {
        // 1. load the tile, rank every pair within its class (void, culled and media pairs: class 8, dropped)
        if (threadIdx.x <= ARTB_PRIM_CLASSES) class_count[threadIdx.x] = 0;
        __syncthreads();
        uint2 pr[ARTB_PAIRS_PER_THREAD];
        uint32_t cls_rank[ARTB_PAIRS_PER_THREAD];
#pragma unroll
        for (int k = 0; k < ARTB_PAIRS_PER_THREAD; ++k) {
            const uint32_t i = t0 + uint32_t(k) * ARTB_PAIRS_THREADS + threadIdx.x;
            pr[k] = i < cnt ? region[i] : make_uint2(0xffffffffu, 0u);
        }
        // hit-aware culling: the ray enters this primitive's leaf box no earlier than tn; if its record already holds a
        // closer hit (merged by pairs processed before — nearer leaves are emitted, hence processed, earlier) the pair
        // cannot win and nothing is loaded or tested for it. The record loads also warm the rays' lines; all of a
        // thread's loads are issued before the first one is looked at.
        unsigned long long rec_lo[ARTB_PAIRS_PER_THREAD], rec_hi[ARTB_PAIRS_PER_THREAD];
#pragma unroll
        for (int k = 0; k < ARTB_PAIRS_PER_THREAD; ++k) {
            rec_lo[k] = 0; rec_hi[k] = 0;
            if (pr[k].x != 0xffffffffu) peek_hit(io.hits.at(pr[k].x & 0xffffffu), rec_lo[k], rec_hi[k]);
        }
#pragma unroll
        for (int k = 0; k < ARTB_PAIRS_PER_THREAD; ++k) {
            uint32_t cls = ARTB_PRIM_CLASSES;
            if (pr[k].x != 0xffffffffu) {
                uint32_t ray, prim;
                float tn;
                unpack_pair(pr[k], ray, prim, tn);
                const int32_t cur_prim = int32_t(uint32_t(rec_hi[k]));
                const bool culled = cur_prim == ARTB_HIT_OVERFLOW || (cur_prim >= 0 && double(tn) > __longlong_as_double((long long)rec_lo[k]));
                if (!culled) cls = s.prim_class[prim];
            }
            cls_rank[k] = (cls << 24) | atomicAdd(&class_count[cls], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t acc = 0;
            for (int q = 0; q <= ARTB_PRIM_CLASSES; ++q) { class_start[q] = acc; acc += class_count[q]; }
        }
        __syncthreads();
        // 2. scatter into class order
#pragma unroll
        for (int k = 0; k < ARTB_PAIRS_PER_THREAD; ++k) {
            const uint32_t cls = cls_rank[k] >> 24;
            if (cls < ARTB_PRIM_CLASSES) tile[class_start[cls] + (cls_rank[k] & 0xffffffu)] = pr[k];
        }
        __syncthreads();
        const uint32_t n_live = class_start[ARTB_PRIM_CLASSES];
        // 3. one pair per thread, warps (almost) uniform in class. Warp-uniform trip count and explicit re-convergence
        // points: the primitive tests have many exits, and without them the compiler clones the merge into every exit.
        for (uint32_t base = threadIdx.x & ~31u; base < n_live; base += ARTB_PAIRS_THREADS) {
            const uint32_t i = base + (threadIdx.x & 31u);
            bool ok = false;
            double t = 0.0;
            uint32_t dfs = 0;
            uint32_t ray = 0, prim = 0;
            unsigned long long cur_lo = 0, cur_hi = 0;
            if (i < n_live) {
                float tn;
                unpack_pair(tile[i], ray, prim, tn);
                peek_hit(io.hits.at(ray), cur_lo, cur_hi);
                const FlatPrim& p = s.prims[prim];
                dfs = p.dfs_id;
                DRay q = trace_io_ray<kBatch>(io, ray);
                if (p.xform) q = to_object_space(s, p.xform, q);
                ok = surface_t(p, q, io.t_min, io.t_max, t);
            }
            warp_sync();
            if (ok && t == t) merge_hit<kBatch>(io, ray, cur_lo, cur_hi, t, prim, dfs);
            warp_sync();
        }
        __syncthreads();
    }
}

// Several leaves report the same t: test them again one by one in the reference's depth-first order, as
// hittable_list::hit / bvh_node::hit would (test_surface applies the exact rule, hit_replaces). Rare: out of line.
__device__ __noinline__ ClosestHit replay_ties(const SceneView* scene, const PathTie* tie, uint32_t n_tied, uint32_t record_prim, DRay r,
                                               double t_min, double t_max)
{
    // (scene through its copy in global memory, the ray by value: a reference to a kernel parameter would make every
    // thread of the calling kernel copy the parameter block to local memory)
    const SceneView& s = *scene;
    uint32_t cand[4];
    for (uint32_t k = 0; k < n_tied; ++k) cand[k] = tie->prim[k];
    cand[n_tied] = record_prim;
    for (uint32_t a = 1; a <= n_tied; ++a)
        for (uint32_t b = a; b > 0 && s.prims[cand[b - 1]].dfs_id > s.prims[cand[b]].dfs_id; --b) {
            const uint32_t tmp = cand[b]; cand[b] = cand[b - 1]; cand[b - 1] = tmp;
        }
    ClosestHit best;
    best.t = t_max; best.prim = -1; best.dfs = 0;
    for (uint32_t k = 0; k <= n_tied; ++k) test_surface(s, cand[k], r, t_min, best);
    return best;
}

// Queue appends of SEVERAL loop iterations with one atomic per queue: a warp parks the (queue, slot) of each of its last
// kBatch iterations in shared memory and reserves the space for all of them at once — one returning atomicAdd per queue
// that has entries, instead of one per queue and iteration. Every warp of the GPU appends to the same 2-6 counters; their
// L2 atomic unit serialises same-address updates (tools/atomic_probe.cu), so the NUMBER of atomics, not their latency, is
// what the append costs. keys[j][t] < 0: no entry.
template <int kBatch, int kThreads, class QueueOf>
__device__ __forceinline__ void flush_batched_appends(int (*keys)[kThreads], uint32_t (*values)[kThreads], int n_iter, int n_queues,
                                                      uint32_t* counters, QueueOf queue_of, int lane)
{
    const uint32_t lt = (1u << lane) - 1u;
    int key[kBatch];
    uint32_t val[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
        key[j] = j < n_iter ? keys[j][threadIdx.x] : -1;
        val[j] = j < n_iter ? values[j][threadIdx.x] : 0u;
    }
    for (int q = 0; q < n_queues; ++q) {
        uint32_t m[kBatch], total = 0;
#pragma unroll
        for (int j = 0; j < kBatch; ++j) { m[j] = __ballot_sync(0xffffffffu, key[j] == q); total += uint32_t(__popc(m[j])); }
        if (total == 0u) continue;                     // warp-uniform
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(counters + q, total);
        base = __shfl_sync(0xffffffffu, base, 0);
        uint32_t* dst = queue_of(q);
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            if (key[j] == q) dst[base + uint32_t(__popc(m[j] & lt))] = val[j];
            base += uint32_t(__popc(m[j]));
        }
    }
}

#ifndef ARTB_FINISH_SCREEN
#define ARTB_FINISH_SCREEN 1          // global rectangles are screened in float before the exact tests (rect_certain_miss)
#endif
#ifndef ARTB_FINISH_BATCH
#define ARTB_FINISH_BATCH 16          // iterations per queue reservation in k_finish. On the final scene and the Cornell box 1, 4, 8 and
                                      // 16 measure the same (500.6 / 493.5 / 499.6 / 500.8 Msamples/s); on the mesh scenes, where half of a
                                      // wave's rays end in the sky (free-slot queue) and the other half on ONE material, every warp hits
                                      // the same two counters every iteration and the wait for the append position was the top stall of
                                      // the kernel (15 % of its samples, profiles/r02_lines_k_finish_cow.txt): k_finish 0.254 -> 0.167 ms
                                      // (8) -> 0.145 ms (16) on the cow, 0.242 -> 0.147 -> 0.134 ms on the dino (+17.6 % / +26.7 %
                                      // samples/s in all)
#endif
#ifndef ARTB_FINISH_DEFER
#define ARTB_FINISH_DEFER 0           // deferred queue appends cost k_finish more in spills than the hidden latency gains
#endif
#ifndef ARTB_FINISH_THREADS
#define ARTB_FINISH_THREADS 128
#endif
#ifndef ARTB_FINISH_BLOCKS
#define ARTB_FINISH_BLOCKS 3             // 168 registers, no spills: measured 3 % faster than 4 blocks of 128 registers with spills
#endif
// kSurf / kMedia: whether the scene HAS global surfaces / global media. The kernel is compiled once per combination so that
// a scene without media (Cornell box, the sphere scenes) or without global surfaces (final scene, the meshes) does not carry
// the code it never runs (Cornell box: -23 % on this kernel), and the variant with neither reads only the 16-byte hit
// records. Register budgets measured per variant (profiles/r02_finish_variants.txt): spills cost every variant more than the
// extra resident warps give back — 3 blocks (168 registers) where media are tested, 4 blocks (128, no spills) elsewhere.
#ifndef ARTB_FINISH_BLOCKS_LIGHT
#define ARTB_FINISH_BLOCKS_LIGHT 4
#endif
#ifndef ARTB_FINISH_BLOCKS_BIN
#define ARTB_FINISH_BLOCKS_BIN 4
#endif
template <bool kBatch, bool kSurf, bool kMedia>
__global__ void __launch_bounds__(ARTB_FINISH_THREADS, kBatch || kMedia ? ARTB_FINISH_BLOCKS : kSurf ? ARTB_FINISH_BLOCKS_LIGHT : ARTB_FINISH_BLOCKS_BIN)
k_finish(TraceIO io, SceneView s, WaveBuffers w, RoundParams rp, int cur)
{
    const uint32_t n = *io.n_ptr;
    const int lane = threadIdx.x & 31;
    WaveCounters* c = w.c;
    const uint32_t n_surf = s.n_global_prims - s.n_global_media;
    const uint32_t stride = gridDim.x * blockDim.x;
#if ARTB_FINISH_SCREEN
    // float screens of the first 32 global surfaces (rect_certain_miss), built once per block
    __shared__ RectScreen s_screen[kSurf ? 32 : 1];
    __shared__ uint32_t s_screen_mask;
    if (kSurf) {
        if (threadIdx.x == 0) s_screen_mask = 0u;
        __syncthreads();
        if (threadIdx.x < 32u && threadIdx.x < n_surf) {
            RectScreen q;
            if (rect_screen_make(s.prims[threadIdx.x], q)) { s_screen[threadIdx.x] = q; atomicOr(&s_screen_mask, 1u << threadIdx.x); }
        }
        __syncthreads();
    }
#endif
    // The ray id of the NEXT iteration is loaded one iteration ahead and its records are prefetched while this
    // iteration computes; all records of this iteration are requested together, before anything waits on them.
    uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u;
    uint32_t id = 0;
    if (base + uint32_t(lane) < n) id = io.list ? io.list[base + lane] : base + uint32_t(lane);
    PendingAppend pend;
    pend.queue = nullptr; pend.value = pend.rank = pend.base_raw = 0; pend.leader = 0;
#if ARTB_FINISH_BATCH > 1
    __shared__ int s_keys[ARTB_FINISH_BATCH][ARTB_FINISH_THREADS];
    __shared__ uint32_t s_ids[ARTB_FINISH_BATCH][ARTB_FINISH_THREADS];
    int n_parked = 0;                  // warp-uniform: iterations parked since the last flush
    auto queue_of = [&](int q) -> uint32_t* { return q == Q_REGEN ? w.q_regen[cur] : w.q_shade + size_t(q) * w.pool; };
#endif
    for (; base < n; base += stride) {
        const uint32_t idx = base + uint32_t(lane);
        const bool valid = idx < n;
        const uint32_t next_idx = idx + stride;
        uint32_t next_id = 0;
        if (next_idx < n) next_id = io.list ? io.list[next_idx] : next_idx;
        PathHit ph;
        ph.t = 0; ph.prim = -1; ph.dfs = 0;
        DRay r;
        r.o = mk3(0, 0, 0); r.d = mk3(1, 0, 0); r.tm = 0;
        PathAux a;
        a.thr[0] = a.thr[1] = a.thr[2] = 0.f; a.pixel = a.sample = a.rng_block = a.depth = a.pad = 0;
        // A scene without global surfaces and media (or with its media tested by k_pairs) needs only the 16-byte hit
        // record here; the ray is read for the rare exact tie, the bookkeeping record for a miss. Everything else reads
        // all three at once.
        constexpr bool kNeedRay = kSurf || kMedia || kBatch;
        if (valid) {
            ph = load_hit(io.hits.at(id));
            if (kNeedRay) r = trace_io_ray<kBatch>(io, id);
            if (!kBatch && kMedia) load_aux(io.aux.at(id), a);
        }
        if (next_idx < n) {
            prefetch_l1(io.hits.at(next_id));
            if (kBatch) prefetch_l1(io.batch_rays + size_t(next_id) * 7);
            else if (kNeedRay) { prefetch_l1(io.rays.at(next_id)); prefetch_l1(io.aux.at(next_id)); }
        }
        // flagged by k_traverse (pairs did not fit), or more leaves tied than the list holds: the one-phase pass traces the ray
        const bool tied = valid && ph.prim >= 0 && (ph.dfs & ARTB_HIT_TIED) != 0u;
        if (!kNeedRay && tied) r = trace_io_ray<kBatch>(io, id);
        if (!kBatch && !kMedia && valid && ph.prim < 0 && ph.prim != ARTB_HIT_OVERFLOW) load_aux(io.aux.at(id), a);     // a miss (no global surface can turn it into a hit)
        uint32_t n_tied = 0;
        if (tied) n_tied = io.ties.at(id)->n;
        const bool overflow = ph.prim == ARTB_HIT_OVERFLOW || n_tied > 3u;
        queue_push(io.q_overflow, io.n_overflow, overflow, id, lane);
        const bool live = valid && !overflow;
        int key = -1;
        if (live) {
            ClosestHit best;
            best.t = ph.t; best.prim = ph.prim; best.dfs = ph.dfs & ~ARTB_HIT_TIED;
            if (tied) best = replay_ties(s.self, io.ties.at(id), n_tied, uint32_t(ph.prim), r, io.t_min, io.t_max);
#if ARTB_FINISH_SCREEN
            if (kSurf) {
                // every lane screens all rectangles in float, then takes only its own candidates through the exact tests
                const uint32_t n_first = n_surf < 32u ? n_surf : 32u;
                uint32_t cand = n_first == 32u ? 0xffffffffu : (1u << n_first) - 1u;
                if (s_screen_mask) {
                    const float of[3] = { float(r.o.x), float(r.o.y), float(r.o.z) }, df[3] = { float(r.d.x), float(r.d.y), float(r.d.z) };
                    const float inv[3] = { 1.0f / df[0], 1.0f / df[1], 1.0f / df[2] };
                    const float t_lo = d2f_rd(io.t_min), t_hi = d2f_ru(best.t);
                    for (uint32_t m = s_screen_mask; m != 0u; m &= m - 1u) {
                        const int i = __ffs(int(m)) - 1;
                        if (rect_certain_miss(s_screen[i], of, df, inv, t_lo, t_hi)) cand &= ~(1u << i);
                    }
                }
                while (cand != 0u) {
                    const int i = __ffs(int(cand)) - 1;
                    cand &= cand - 1u;
                    test_surface(s, uint32_t(i), r, io.t_min, best);
                }
                for (uint32_t i = 32u; i < n_surf; ++i) test_surface(s, i, r, io.t_min, best);
            }
#else
            if (kSurf) for (uint32_t i = 0; i < n_surf; ++i) test_surface(s, i, r, io.t_min, best);
#endif
            if (kMedia && s.n_global_media) {
                MediumXi xi;
                if (kBatch) {
                    xi.row = io.batch_xi ? io.batch_xi + size_t(id) * s.n_media : nullptr;
                    xi.seed = 0; xi.stream = 0; xi.path_lo = id; xi.path_hi = 0; xi.bounce = 0;
                } else {
                    xi.row = nullptr; xi.seed = rp.seed; xi.stream = rp.stream;
                    xi.path_lo = a.pixel; xi.path_hi = a.sample; xi.bounce = a.depth;
                }
                const BoxRay br = make_box_ray(make_trav_ray(r));
                test_global_media(s, r, br, io.t_min, d2f_rd(io.t_min), xi, best, nullptr);
            }
            if (kBatch) {
                HitRecordOut o;
                if (best.prim < 0) {
                    o.hit = 0; o.prim_id = -1; o.front_face = 0; o.material_type = -1;
                    o.t = o.u = o.v = 0; o.p[0] = o.p[1] = o.p[2] = 0; o.normal[0] = o.normal[1] = o.normal[2] = 0;
                } else {
                    const HitInfo hi = make_hit_info(s, best.prim, r, best.t);
                    o.hit = 1; o.prim_id = int32_t(best.dfs); o.front_face = hi.front_face ? 1 : 0;
                    o.material_type = int32_t(s.materials[hi.material].type);
                    o.t = hi.t; o.u = hi.u; o.v = hi.v;
                    o.p[0] = hi.p.x; o.p[1] = hi.p.y; o.p[2] = hi.p.z;
                    o.normal[0] = hi.normal.x; o.normal[1] = hi.normal.y; o.normal[2] = hi.normal.z;
                }
                io.batch_out[id] = o;
            } else if (best.prim < 0) {
                key = Q_REGEN;
                add_radiance(rp, a.pixel, a.thr[0] * float(s.background[0]), a.thr[1] * float(s.background[1]),
                             a.thr[2] * float(s.background[2]));
            } else {
                store_hit(w.hits.at(id), best.t, best.prim, best.dfs);
                key = int(s.prim_mat_type[best.prim]);
            }
        }
        if (!kBatch) {
#if ARTB_FINISH_BATCH > 1
            s_keys[n_parked][threadIdx.x] = key;
            s_ids[n_parked][threadIdx.x] = id;
            if (++n_parked == ARTB_FINISH_BATCH) {
                flush_batched_appends<ARTB_FINISH_BATCH, ARTB_FINISH_THREADS>(s_keys, s_ids, n_parked, Q_OUT_COUNT, &c->n_out[cur][0], queue_of, lane);
                n_parked = 0;
            }
#elif ARTB_FINISH_DEFER
            append_commit(pend);             // the previous iteration's entries: their atomics have long returned
            append_reserve(pend, key, &c->n_out[cur][key < 0 ? 0 : key],
                           key == Q_REGEN ? w.q_regen[cur] : w.q_shade + size_t(key < 0 ? 0 : key) * w.pool, id, lane);
#else
            const uint32_t pos = multi_queue_reserve(key, &c->n_out[cur][key < 0 ? 0 : key], lane);
            if (key >= 0) {
                uint32_t* q = key == Q_REGEN ? w.q_regen[cur] : w.q_shade + size_t(key) * w.pool;
                q[pos] = id;
            }
#endif
        }
        id = next_id;
    }
#if ARTB_FINISH_BATCH > 1
    if (!kBatch && n_parked) flush_batched_appends<ARTB_FINISH_BATCH, ARTB_FINISH_THREADS>(s_keys, s_ids, n_parked, Q_OUT_COUNT, &c->n_out[cur][0], queue_of, lane);
#endif
    if (!kBatch) append_commit(pend);
}

// k_bin — what is left of k_finish for a scene without global surfaces and without media (the 1 M-triangle height field, the
// sphere scenes): read the 16-byte hit record of every ray, add the background for a miss, bin the hits by material. There
// is nothing to compute, only records to fetch, so the kernel is kept to 32 registers — 64 resident warps per SM instead
// of the 16 of k_finish — and reads nothing but the hit record unless the ray missed (its bookkeeping record) or carries an
// exact tie (rare: resolved out of line, which is also where the ray itself is read).
__device__ __noinline__ void bin_resolve_tie(const SceneView* scene, PathHit* hit, const PathTie* tie, const PathRay* ray, double t_min, double t_max)
{
    const uint32_t n_tied = tie->n;
    if (n_tied > 3u) { hit->prim = ARTB_HIT_OVERFLOW; return; }              // more ties than the list holds: the one-phase pass
    const PathHit ph = load_hit(hit);
    const ClosestHit best = replay_ties(scene, tie, n_tied, uint32_t(ph.prim), load_ray(ray), t_min, t_max);
    store_hit(hit, best.t, best.prim, best.dfs);
}

#ifndef ARTB_BIN_THREADS
#define ARTB_BIN_THREADS 256
#endif
#ifndef ARTB_BIN_BLOCKS
#define ARTB_BIN_BLOCKS 8
#endif
#ifndef ARTB_BIN_BATCH
#define ARTB_BIN_BATCH 8              // iterations per queue reservation (as ARTB_FINISH_BATCH): height field +7.5 %; 16 spills at the 32 registers of k_bin
#endif
__global__ void __launch_bounds__(ARTB_BIN_THREADS, ARTB_BIN_BLOCKS)
k_bin(TraceIO io, SceneView s, WaveBuffers w, RoundParams rp, int cur)
{
    const uint32_t n = *io.n_ptr;
    const int lane = threadIdx.x & 31;
    WaveCounters* c = w.c;
    const uint32_t stride = gridDim.x * blockDim.x;
#if ARTB_BIN_BATCH > 1
    __shared__ int s_keys[ARTB_BIN_BATCH][ARTB_BIN_THREADS];
    __shared__ uint32_t s_ids[ARTB_BIN_BATCH][ARTB_BIN_THREADS];
    int n_parked = 0;                  // warp-uniform: iterations parked since the last flush
    auto queue_of = [&](int q) -> uint32_t* { return q == Q_REGEN ? w.q_regen[cur] : w.q_shade + size_t(q) * w.pool; };
#endif
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += stride) {
        const uint32_t idx = base + uint32_t(lane);
        const bool valid = idx < n;
        uint32_t id = 0;
        PathHit ph;
        ph.t = 0; ph.prim = -1; ph.dfs = 0;
        if (valid) {
            id = io.list ? io.list[idx] : idx;
            ph = load_hit(io.hits.at(id));
            if (ph.prim >= 0 && (ph.dfs & ARTB_HIT_TIED) != 0u) {
                bin_resolve_tie(s.self, io.hits.at(id), io.ties.at(id), io.rays.at(id), io.t_min, io.t_max);
                ph = load_hit(io.hits.at(id));
            }
        }
        const bool overflow = valid && ph.prim == ARTB_HIT_OVERFLOW;
        queue_push(io.q_overflow, io.n_overflow, overflow, id, lane);
        int key = -1;
        if (valid && !overflow) {
            if (ph.prim < 0) {
                key = Q_REGEN;
                PathAux a;
                load_aux(io.aux.at(id), a);
                add_radiance(rp, a.pixel, a.thr[0] * float(s.background[0]), a.thr[1] * float(s.background[1]),
                             a.thr[2] * float(s.background[2]));
            } else {
                key = int(s.prim_mat_type[ph.prim]);
            }
        }
#if ARTB_BIN_BATCH > 1
        s_keys[n_parked][threadIdx.x] = key;
        s_ids[n_parked][threadIdx.x] = id;
        if (++n_parked == ARTB_BIN_BATCH) {
            flush_batched_appends<ARTB_BIN_BATCH, ARTB_BIN_THREADS>(s_keys, s_ids, n_parked, Q_OUT_COUNT, &c->n_out[cur][0], queue_of, lane);
            n_parked = 0;
        }
#else
        const uint32_t pos = multi_queue_reserve(key, &c->n_out[cur][key < 0 ? 0 : key], lane);
        if (key >= 0) {
            uint32_t* q = key == Q_REGEN ? w.q_regen[cur] : w.q_shade + size_t(key) * w.pool;
            q[pos] = id;
        }
#endif
    }
#if ARTB_BIN_BATCH > 1
    if (n_parked) flush_batched_appends<ARTB_BIN_BATCH, ARTB_BIN_THREADS>(s_keys, s_ids, n_parked, Q_OUT_COUNT, &c->n_out[cur][0], queue_of, lane);
#endif
}

// The one-phase walk: closest_hit() per ray, results appended to the same queues. all = 0: the rays k_finish found
// flagged (their pairs did not fit); all = 1: every ray of q_extend[cur] (scenes whose tree does not fit in shared
// memory: there the pruning by the hits found so far saves more than the two-phase split gains). Chunks of 32 rays
// are handed out by an atomic cursor.
#ifndef ARTB_EXTEND_THREADS
#define ARTB_EXTEND_THREADS 128
#endif
__global__ void __launch_bounds__(ARTB_EXTEND_THREADS, 4)
k_extend(WaveBuffers w, SceneView s, RoundParams rp, int cur, int all)
{
    WaveCounters* c = w.c;
    const uint32_t n = all ? c->n_extend[cur] : c->n_overflow;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (all) {
            const unsigned long long left = c->total - c->issued;
            const unsigned long long nr = c->n_out[cur ^ 1][Q_REGEN];
            c->issued += nr < left ? nr : left;
            c->n_out[cur ^ 1][Q_REGEN] = 0;
            c->rays += n;
        } else {
            c->overflow_total += n;
        }
    }
    if (n == 0) return;
    const uint32_t* list = all ? w.q_extend[cur] : w.q_overflow;
    const int lane = threadIdx.x & 31;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&c->cursor[2], 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const bool valid = base + lane < n;
        const uint32_t slot = valid ? list[base + lane] : 0u;
        int key = -1;
        if (valid) {
            const DRay r = load_ray(w.rays.at(slot));
            PathAux a;
            load_aux(w.aux.at(slot), a);
            MediumXi xi;
            xi.row = nullptr; xi.seed = rp.seed; xi.stream = rp.stream;
            xi.path_lo = a.pixel; xi.path_hi = a.sample; xi.bounce = a.depth;
            const ClosestHit h = closest_hit(s, s.nodes, r, 0.001, INFINITY, xi);
            if (h.prim < 0) {
                key = Q_REGEN;
                add_radiance(rp, a.pixel, a.thr[0] * float(s.background[0]), a.thr[1] * float(s.background[1]),
                             a.thr[2] * float(s.background[2]));
            } else {
                store_hit(w.hits.at(slot), h.t, h.prim, h.dfs);
                key = int(s.materials[s.prims[h.prim].material].type);
            }
        }
        const uint32_t pos = multi_queue_reserve(key, &c->n_out[cur][key < 0 ? 0 : key], lane);
        if (key >= 0) {
            uint32_t* q = key == Q_REGEN ? w.q_regen[cur] : w.q_shade + size_t(key) * w.pool;
            q[pos] = slot;
        }
    }
}

// One warp = 32 entries of ONE queue: the hits of one material type (hit record, texture, emitted, scatter), or 32
// free slots that start the next (pixel, sample) of the round: _stochastic_sample (engine.h:62-63) + camera::get_ray.
// Scattered and fresh rays go to q_extend[cur^1]; slots whose path ended (light, absorbed, depth limit) or that found
// no sample left go to the free list q_regen[cur^1].
#ifndef ARTB_SHADE_BLOCKS
#define ARTB_SHADE_BLOCKS 4
#endif
#ifndef ARTB_SHADE_BATCH
#define ARTB_SHADE_BATCH 4            // iterations per queue reservation (1: the deferred per-iteration append)
#endif
__global__ void __launch_bounds__(128, ARTB_SHADE_BLOCKS) k_shade(WaveBuffers w, SceneView s, RoundParams rp, int cur)
{
    WaveCounters* c = w.c;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c->n_extend[cur] = 0;
#pragma unroll
        for (int m = 0; m < MAT_COUNT; ++m) c->n_out[cur ^ 1][m] = 0;
        c->cursor[0] = c->cursor[1] = c->cursor[2] = 0;
        c->n_overflow = 0;
    }
    uint32_t start[Q_OUT_COUNT + 1], qn[Q_OUT_COUNT];
    start[0] = 0;
#pragma unroll
    for (int m = 0; m < Q_OUT_COUNT; ++m) { qn[m] = c->n_out[cur][m]; start[m + 1] = start[m] + ((qn[m] + 31u) >> 5); }
    const unsigned long long issued = c->issued, total = c->total;
    const uint32_t n_list = c->n_list;
    const int lane = threadIdx.x & 31;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    // which queue a warp index belongs to, the lane's entry in it and the slot stored there
    auto locate = [&](uint32_t wi, int& m, uint32_t& i, bool& valid) -> uint32_t {
        m = 0;
        uint32_t first = 0, count = qn[0];
#pragma unroll
        for (int k = 1; k < Q_OUT_COUNT; ++k) if (wi >= start[k]) { m = k; first = start[k]; count = qn[k]; }
        i = (wi - first) * 32u + uint32_t(lane);
        valid = wi < start[Q_OUT_COUNT] && i < count;
        if (!valid) return 0u;
        return m == Q_REGEN ? w.q_regen[cur][i] : w.q_shade[size_t(m) * w.pool + i];
    };
    // software pipeline: the slot of the NEXT iteration is loaded one iteration ahead and its records are prefetched
    // while this iteration shades (two iterations deep measured the same)
    uint32_t wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int m = 0; uint32_t i = 0; bool valid = false;
    uint32_t slot = locate(wi, m, i, valid);
    PendingAppend pend;
    pend.queue = nullptr; pend.value = pend.rank = pend.base_raw = 0; pend.leader = 0;
#if ARTB_SHADE_BATCH > 1
    // appends of ARTB_SHADE_BATCH iterations are reserved with ONE atomic per queue (flush_batched_appends): queue 0 = rays
    // for the next wave (n_extend), queue 1 = free slots (n_out[.][Q_REGEN]); the two counters are 48 bytes apart
    __shared__ int s_keys[ARTB_SHADE_BATCH][128];
    __shared__ uint32_t s_slots[ARTB_SHADE_BATCH][128];
    int n_parked = 0;
    uint32_t* const ctr2 = &c->n_extend[cur ^ 1];
    const int regen_off = int(&c->n_out[cur ^ 1][Q_REGEN] - ctr2);
    auto flush = [&]() {
        const uint32_t lt = (1u << lane) - 1u;
        int key[ARTB_SHADE_BATCH]; uint32_t val[ARTB_SHADE_BATCH];
#pragma unroll
        for (int j = 0; j < ARTB_SHADE_BATCH; ++j) {
            key[j] = j < n_parked ? s_keys[j][threadIdx.x] : -1;
            val[j] = j < n_parked ? s_slots[j][threadIdx.x] : 0u;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            uint32_t mk[ARTB_SHADE_BATCH], total = 0;
#pragma unroll
            for (int j = 0; j < ARTB_SHADE_BATCH; ++j) { mk[j] = __ballot_sync(0xffffffffu, key[j] == q); total += uint32_t(__popc(mk[j])); }
            if (total == 0u) continue;
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(ctr2 + (q ? regen_off : 0), total);
            base = __shfl_sync(0xffffffffu, base, 0);
            uint32_t* dst = q ? w.q_regen[cur ^ 1] : w.q_extend[cur ^ 1];
#pragma unroll
            for (int j = 0; j < ARTB_SHADE_BATCH; ++j) {
                if (key[j] == q) dst[base + uint32_t(__popc(mk[j] & lt))] = val[j];
                base += uint32_t(__popc(mk[j]));
            }
        }
        n_parked = 0;
    };
#endif
    for (; wi < start[Q_OUT_COUNT]; wi += n_warps) {
        int next_m = 0; uint32_t next_i = 0; bool next_valid = false;
        const uint32_t next_slot = locate(wi + n_warps, next_m, next_i, next_valid);
        int key = -1;                       // 0: a ray for the next wave, 1: slot is free
        if (m == Q_REGEN) {
            if (valid) {
                const unsigned long long sidx = issued + i;
                key = 1;
                if (sidx < total) {
                    key = 0;
                    const uint32_t li = uint32_t(sidx % n_list);
                    const uint32_t k = uint32_t(sidx / n_list);
                    const uint32_t pixel = rp.pixel_list ? rp.pixel_list[li] : li;
                    const uint32_t sample = rp.first_sample + k;
                    const int px = int(pixel % uint32_t(rp.width)), py = int(pixel / uint32_t(rp.width));
                    Philox g;
                    g.init(rp.seed, rp.stream, pixel, sample, 0, 0);
                    // engine.h:62-63 — row 0 is the top of the image, v is flipped
                    const double u = (px + g.next_double()) / (rp.width - 1);
                    const double v = ((rp.height - 1 - py) + g.next_double()) / (rp.height - 1);
                    const DRay r = camera_ray(rp.cam, u, v, g);
                    store_ray(w.rays.at(slot), r);
                    PathAux a;
                    a.thr[0] = a.thr[1] = a.thr[2] = 1.f;
                    a.pixel = pixel; a.sample = sample; a.rng_block = g.c2; a.depth = 0; a.pad = 0;
                    store_aux(w.aux.at(slot), a);
                }
            }
        } else if (valid) {
            const DRay r = load_ray(w.rays.at(slot));
            PathAux a;
            load_aux(w.aux.at(slot), a);
            const PathHit ph = load_hit(w.hits.at(slot));
            if (next_valid && next_m != Q_REGEN) {
                prefetch_l1(w.rays.at(next_slot)); prefetch_l1(w.aux.at(next_slot)); prefetch_l1(w.hits.at(next_slot));
            }
            const HitInfo h = make_hit_info(s, ph.prim, r, ph.t);
            const FlatMaterial& mat = s.materials[h.material];
            key = 1;
            if (mat.type == MAT_DIFFUSE_LIGHT) {
                const d3 e = material_emitted(s, mat, h);
                add_radiance(rp, a.pixel, a.thr[0] * float(e.x), a.thr[1] * float(e.y), a.thr[2] * float(e.z));
            } else {
                Philox g;
                g.init(rp.seed, rp.stream, a.pixel, a.sample, a.rng_block, 0);
                d3 att; DRay sc;
                const bool ok = material_scatter(s, mat, r, h, g, att, sc);
                if (ok && a.depth + 1u < uint32_t(rp.max_depth)) {
                    key = 0;
                    store_ray(w.rays.at(slot), sc);
                    a.thr[0] *= float(att.x); a.thr[1] *= float(att.y); a.thr[2] *= float(att.z);
                    a.depth += 1; a.rng_block = g.c2;
                    store_aux(w.aux.at(slot), a);
                } else {
                    add_radiance(rp, a.pixel, 0.f, 0.f, 0.f);   // absorbed / depth limit: a black sample (engine.h:451-452, 462-463)
                }
            }
        }
#if ARTB_SHADE_BATCH > 1
        s_keys[n_parked][threadIdx.x] = key;
        s_slots[n_parked][threadIdx.x] = slot;
        if (++n_parked == ARTB_SHADE_BATCH) flush();
#else
        append_commit(pend);                 // the previous iteration's entries: their atomics have long returned
        append_reserve(pend, key, key == 1 ? &c->n_out[cur ^ 1][Q_REGEN] : &c->n_extend[cur ^ 1],
                       key == 1 ? w.q_regen[cur ^ 1] : w.q_extend[cur ^ 1], slot, lane);
#endif
        m = next_m; i = next_i; valid = next_valid; slot = next_slot;
    }
#if ARTB_SHADE_BATCH > 1
    if (n_parked) flush();
#endif
    append_commit(pend);
}

// adaptive variance mask: keep the pixels of the current list whose luminance mean is not yet settled.
// The per-sample variance is POOLED over the 3x3 neighbourhood: a pixel whose own first samples all came back
// black next to neighbours that caught rare bright paths is not declared converged on its own zero variance.
__global__ void k_adaptive_update(WaveCounters* c, const uint32_t* list, uint32_t* next_list, const float* accum,
                                  const uint32_t* count, const float* lumsq, int width, int height, uint32_t min_spp,
                                  uint32_t max_spp, float threshold)
{
    const uint32_t n = c->n_list;
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += stride) {
        const uint32_t i = base + lane;
        bool keep = false;
        uint32_t pixel = 0;
        if (i < n) {
            pixel = list ? list[i] : i;
            const uint32_t cnt = count[pixel];
            if (cnt < max_spp) {
                keep = true;
                if (cnt >= min_spp && cnt > 1) {
                    const int px = int(pixel % uint32_t(width)), py = int(pixel / uint32_t(width));
                    float sum = 0.f, sumsq = 0.f, nn = 0.f;
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int x = px + dx, y = py + dy;
                            if (x < 0 || y < 0 || x >= width || y >= height) continue;
                            const uint32_t q = uint32_t(y * width + x);
                            const float cq = float(count[q]);
                            if (cq == 0.f) continue;
                            const float* a = accum + size_t(q) * 3;
                            sum += 0.2126f * a[0] + 0.7152f * a[1] + 0.0722f * a[2];
                            sumsq += lumsq[q];
                            nn += cq;
                        }
                    const float pooled_mean = sum / nn;
                    const float var = fmaxf(sumsq / nn - pooled_mean * pooled_mean, 0.f) * nn / fmaxf(nn - 1.f, 1.f);
                    const float* a = accum + size_t(pixel) * 3;
                    const float mean = (0.2126f * a[0] + 0.7152f * a[1] + 0.0722f * a[2]) / float(cnt);
                    const float se = sqrtf(var / float(cnt));
                    keep = se > threshold * (fmaxf(mean, pooled_mean) + 0.05f);
                }
            }
        }
        queue_push(next_list, &c->n_next_list, keep, pixel, lane);
    }
}

// Partition B (tile split): the pixels of the tiles t = ty*tiles_x + tx with t % parts == part, tile by tile (row-major
// inside a tile). One block per owned tile; offsets[k] = pixels of the owned tiles before the k-th one (a few thousand
// numbers, prefix-summed on the host once per (size, tiling, part) and cached with the list).
__global__ void k_tile_list(uint32_t* list, const uint32_t* offsets, int width, int height, int ts, int parts, int part)
{
    const int tiles_x = (width + ts - 1) / ts;
    const int t = part + int(blockIdx.x) * parts;
    const int ty = t / tiles_x, tx = t % tiles_x;
    const int w = (width - tx * ts < ts ? width - tx * ts : ts), h = (height - ty * ts < ts ? height - ty * ts : ts);
    const uint32_t offset = offsets[blockIdx.x];
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
        const int x = tx * ts + i % w, y = ty * ts + i / w;
        list[offset + uint32_t(i)] = uint32_t(y) * uint32_t(width) + uint32_t(x);
    }
}

// fixed-spp rendering: every listed pixel received exactly `spp` samples (the count buffer of include/artb.h)
__global__ void k_count_fixed(const uint32_t* list, uint32_t n, uint32_t spp, uint32_t* count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) count[list ? list[i] : i] += spp;
}

__global__ void k_resolve(const float* accum, const uint32_t* count, int n_pixels, int fixed_spp, uint8_t* rgb)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    const double n = fixed_spp > 0 ? double(fixed_spp) : double(count[i] > 0 ? count[i] : 1u);
    const double scale = 1.0 / n;                          // color.h:12-21
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double v = sqrt(scale * double(accum[size_t(i) * 3 + k]));
        v = v < 0.0 ? 0.0 : (v > 0.999 ? 0.999 : v);
        rgb[size_t(i) * 3 + k] = uint8_t(256 * v);
    }
}

// one-phase walk over explicit rays: the overflow pass of the batch path (list = the flagged ray indices, their number
// read from device memory: nothing waits on the host) and the path of scenes whose ids do not fit pack_pair
__global__ void __launch_bounds__(128) k_trace_batch(SceneView s, const double* rays, const uint32_t* list, uint64_t n_fixed,
                                                     const uint32_t* n_ptr, unsigned long long* n_total, double t_min, double t_max,
                                                     const double* xi, HitRecordOut* out)
{
    const uint64_t n = n_ptr ? uint64_t(*n_ptr) : n_fixed;
    if (n_total && blockIdx.x == 0 && threadIdx.x == 0) *n_total += n;
    for (uint64_t j = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < n; j += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t i = list ? list[j] : j;
        DRay r;
        r.o = ld3(rays + i * 7); r.d = ld3(rays + i * 7 + 3); r.tm = rays[i * 7 + 6];
        MediumXi mx;
        mx.row = xi ? xi + i * s.n_media : nullptr;
        mx.seed = 0; mx.stream = 0; mx.path_lo = uint32_t(i); mx.path_hi = uint32_t(i >> 32); mx.bounce = 0;
        const ClosestHit h = closest_hit(s, s.nodes, r, t_min, t_max, mx);
        HitRecordOut o;
        if (h.prim < 0) {
            o.hit = 0; o.prim_id = -1; o.front_face = 0; o.material_type = -1;
            o.t = o.u = o.v = 0; o.p[0] = o.p[1] = o.p[2] = 0; o.normal[0] = o.normal[1] = o.normal[2] = 0;
        } else {
            const HitInfo hi = make_hit_info(s, h.prim, r, h.t);
            o.hit = 1; o.prim_id = int32_t(h.dfs); o.front_face = hi.front_face ? 1 : 0;
            o.material_type = int32_t(s.materials[hi.material].type);
            o.t = hi.t; o.u = hi.u; o.v = hi.v;
            o.p[0] = hi.p.x; o.p[1] = hi.p.y; o.p[2] = hi.p.z;
            o.normal[0] = hi.normal.x; o.normal[1] = hi.normal.y; o.normal[2] = hi.normal.z;
        }
        out[i] = o;
    }
}

// ---- shade-stage parity hooks (artb_shade_batch / artb_camera_rays): the device functions k_shade runs — make_hit_info,
// material_emitted, material_scatter, texture_value, camera_ray — on explicit inputs with explicit random numbers
__global__ void k_dfs_map(const FlatPrim* prims, uint32_t n_world, int32_t* map, uint32_t map_size)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_world && prims[i].dfs_id < map_size) map[prims[i].dfs_id] = int32_t(i);
}

__global__ void __launch_bounds__(128) k_shade_batch(SceneView s, const double* rays, const int32_t* prim_id, const double* t,
                                                     const double* xi, const int32_t* dfs_map, uint32_t map_size, uint64_t n,
                                                     ShadeRecordOut* out)
{
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ShadeRecordOut o;
    memset(&o, 0, sizeof(o));
    const int32_t id = prim_id[i];
    const int32_t prim = id >= 0 && uint32_t(id) < map_size ? dfs_map[id] : -1;
    if (prim < 0) {
        o.scattered = -1; o.material_type = -1;
        out[i] = o;
        return;
    }
    DRay r;
    r.o = ld3(rays + i * 7); r.d = ld3(rays + i * 7 + 3); r.tm = rays[i * 7 + 6];
    const HitInfo h = make_hit_info(s, prim, r, t[i]);
    const FlatMaterial& mat = s.materials[h.material];
    const d3 e = material_emitted(s, mat, h);
    XiTable g{ xi + i * 4 };
    d3 att = mk3(0, 0, 0);
    DRay sc;
    sc.o = mk3(0, 0, 0); sc.d = mk3(0, 0, 0); sc.tm = 0;
    const bool ok = material_scatter(s, mat, r, h, g, att, sc);
    o.emitted[0] = e.x; o.emitted[1] = e.y; o.emitted[2] = e.z;
    o.scattered = ok ? 1 : 0;
    o.material_type = int32_t(mat.type);
    o.front_face = h.front_face ? 1 : 0;
    if (mat.type != MAT_DIFFUSE_LIGHT) {          // diffuse_light::scatter returns before touching its outputs
        o.attenuation[0] = att.x; o.attenuation[1] = att.y; o.attenuation[2] = att.z;
        o.origin[0] = sc.o.x; o.origin[1] = sc.o.y; o.origin[2] = sc.o.z;
        o.direction[0] = sc.d.x; o.direction[1] = sc.d.y; o.direction[2] = sc.d.z;
        o.time = sc.tm;
    }
    out[i] = o;
}

__global__ void k_camera_rays(CameraParams cam, const double* st, const double* xi, uint64_t n, double* rays_out)
{
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    XiTable g{ xi + i * 3 };
    const DRay r = camera_ray(cam, st[i * 2], st[i * 2 + 1], g);
    double* o = rays_out + i * 7;
    o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.d.x; o[4] = r.d.y; o[5] = r.d.z; o[6] = r.tm;
}

// counters of one chunk of artb_trace_batch: {n, traverse cursor, -, overflow count}
__global__ void k_batch_counters(uint32_t* ctr, uint32_t n)
{
    ctr[0] = n; ctr[1] = 0u; ctr[2] = 0u; ctr[3] = 0u;
}

// =====================================================================================================
// host side
// =====================================================================================================
struct Renderer::Impl {
    // scene arrays
    std::vector<void*> scene_allocs;
    // wave buffers
    WaveBuffers wb{};
    std::vector<void*> pool_allocs;
    WaveCounters* counters_host = nullptr;      // pinned: two snapshots (counters_host[0..1]), polled one batch of waves behind
    cudaEvent_t poll_ev[2] = { nullptr, nullptr };
    int sm_count = 148;
    size_t max_shared_nodes = 112 << 10;        // node array up to this size is staged in shared memory by k_traverse
    size_t max_traverse_smem = 112 << 10;       // dynamic shared memory k_traverse may ask for (nodes + traversal stack)
    int bvh_depth = ARTB_STACK_DEPTH;           // levels of 8-wide nodes of the uploaded tree (upload / build_bvh_device)
    SceneView* view_dev = nullptr;              // copy of the current SceneView in global memory (TraceIO::scene_dev)
    // device scratch of render_host (grow-only: cudaMalloc costs tens of milliseconds next to a render)
    float* host_accum = nullptr; uint32_t* host_count = nullptr; uint8_t* host_rgb = nullptr;
    size_t host_pixels = 0;
    // adaptive scratch
    uint32_t* list[2] = { nullptr, nullptr };
    uint32_t* tile_list = nullptr;
    int tile_key[5] = { 0, 0, 0, 0, -1 };         // width, height, tile size, parts, part of the list held in tile_list
    size_t tile_count = 0;
    uint32_t* tile_offsets = nullptr;
    size_t tile_offsets_capacity = 0;
    float* lumsq = nullptr;
    size_t list_capacity = 0;
    // event pool for per-kernel timing
    std::vector<cudaEvent_t> ev;
    // scratch of trace_batch_device (grow-only: the hook is also what "Mrays/s on fixed batches" is measured with)
    struct BatchScratch {
        std::vector<void*> bag;
        PathHit* hits = nullptr; PathTie* ties = nullptr; uint2* pairs = nullptr; uint32_t* region_count = nullptr;
        uint32_t* overflow = nullptr; uint32_t* ctr = nullptr; unsigned long long* overflow_total = nullptr;
        uint64_t cap = 0; size_t pair_capacity = 0;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
    } batch;

    // launch helpers shared by the renderer and the batch path. The tree walk runs in shared memory when the node
    // array fits; its grid decides how many regions the pair buffer is cut into.
    bool nodes_fit_shared(const SceneView& view) const
    {
        const size_t node_bytes = size_t(view.n_nodes) * sizeof(FlatNode8);
        return node_bytes <= max_shared_nodes;
    }
    void set_pair_buffer(const SceneView& view, TraceIO& io, uint2* pairs, size_t capacity, uint32_t* region_count) const
    {
        io.pb.pairs = pairs; io.pb.region_count = region_count;
        io.pb.n_regions = nodes_fit_shared(view) ? uint32_t(sm_count) * ARTB_TRAVERSE_GRID_PER_SM : uint32_t(sm_count) * 6u;
        io.pb.region_cap = uint32_t(std::min<size_t>(capacity / io.pb.n_regions, 0x7fffffffu));
    }
    template <bool kBatch>
    void launch_finish(const SceneView& view, const TraceIO& io, const WaveBuffers& w, const RoundParams& rp, int cur, cudaStream_t st)
    {
        const bool surf = view.n_global_prims > view.n_global_media, media = view.n_global_media > 0;
        const unsigned heavy = unsigned(sm_count) * ARTB_FINISH_BLOCKS;
        const unsigned light = unsigned(sm_count) * (kBatch ? ARTB_FINISH_BLOCKS : ARTB_FINISH_BLOCKS_LIGHT);
        const unsigned bin = unsigned(sm_count) * (kBatch ? ARTB_FINISH_BLOCKS : ARTB_FINISH_BLOCKS_BIN);
        if (surf && media) k_finish<kBatch, true, true><<<heavy, ARTB_FINISH_THREADS, 0, st>>>(io, view, w, rp, cur);
        else if (surf) k_finish<kBatch, true, false><<<light, ARTB_FINISH_THREADS, 0, st>>>(io, view, w, rp, cur);
        else if (media) k_finish<kBatch, false, true><<<heavy, ARTB_FINISH_THREADS, 0, st>>>(io, view, w, rp, cur);
        else if (kBatch) k_finish<kBatch, false, false><<<bin, ARTB_FINISH_THREADS, 0, st>>>(io, view, w, rp, cur);
        else k_bin<<<unsigned(sm_count) * ARTB_BIN_BLOCKS, ARTB_BIN_THREADS, 0, st>>>(io, view, w, rp, cur);
    }
    // two_phase: k_traverse + k_pairs + k_finish; otherwise the pruned walk k_walk + k_finish (trees beyond shared memory).
    // ev (optional): three events recorded after the tree kernel, after k_pairs (at once in walk mode) and after k_finish.
    template <bool kBatch>
    void launch_trace(const SceneView& view, const TraceIO& io, const WaveBuffers& w, const RoundParams& rp, int cur, cudaStream_t st,
                      cudaEvent_t* ev, bool two_phase)
    {
        if (!two_phase) {
            k_walk<kBatch><<<sm_count * ARTB_WALK_BLOCKS, ARTB_WALK_THREADS, 0, st>>>(io, view, w.c, cur);
            if (ev) { CUDA_CHECK(cudaEventRecord(ev[0], st)); CUDA_CHECK(cudaEventRecord(ev[1], st)); }
            launch_finish<kBatch>(view, io, w, rp, cur, st);
            if (ev) CUDA_CHECK(cudaEventRecord(ev[2], st));
            CUDA_CHECK(cudaGetLastError());
            return;
        }
        const size_t node_bytes = std::max<size_t>(size_t(view.n_nodes) * sizeof(FlatNode8), 16);
        const size_t stack_bytes = size_t(ARTB_SMEM_STACK) * ARTB_TRAVERSE_THREADS * sizeof(uint2);
        if (nodes_fit_shared(view) && ARTB_SMEM_STACK > 0 && bvh_depth <= ARTB_SMEM_STACK && node_bytes + stack_bytes <= max_traverse_smem) {
            k_traverse<true, kBatch, ARTB_SMEM_STACK><<<io.pb.n_regions, ARTB_TRAVERSE_THREADS, node_bytes + stack_bytes, st>>>(io, view, w.c, cur);
        } else if (nodes_fit_shared(view)) {
            // (the opt-in shared-memory limit of k_traverse is raised once, in Renderer::init: the attribute belongs to the
            // function, not to this object — a per-object cache let a second context with a smaller tree LOWER it)
            k_traverse<true, kBatch, 0><<<io.pb.n_regions, ARTB_TRAVERSE_THREADS, node_bytes, st>>>(io, view, w.c, cur);
        } else {
            k_traverse<false, kBatch, 0><<<io.pb.n_regions, 256, 0, st>>>(io, view, w.c, cur);
        }
        if (ev) CUDA_CHECK(cudaEventRecord(ev[0], st));
        const uint32_t J = std::max(1u, uint32_t(sm_count) * ARTB_PAIRS_BLOCKS / io.pb.n_regions);
        k_pairs<kBatch><<<io.pb.n_regions * J, ARTB_PAIRS_THREADS, 0, st>>>(io, view);
        if (ev) CUDA_CHECK(cudaEventRecord(ev[1], st));
        launch_finish<kBatch>(view, io, w, rp, cur, st);
        if (ev) CUDA_CHECK(cudaEventRecord(ev[2], st));
        CUDA_CHECK(cudaGetLastError());
    }
};

template <class T>
static T* dev_alloc(std::vector<void*>& bag, size_t n)
{
    void* p = nullptr;
    CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    bag.push_back(p);
    return static_cast<T*>(p);
}

template <class T>
static const T* dev_upload(std::vector<void*>& bag, const std::vector<T>& v, cudaStream_t st, uint64_t& bytes)
{
    T* p = dev_alloc<T>(bag, v.size());
    if (!v.empty()) CUDA_CHECK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    bytes += v.size() * sizeof(T);
    return p;
}

Renderer::Renderer(int device, uint32_t pool_paths) : impl_(new Impl), device_(device)
{
    // a constructor that throws does not run the destructor: release what was acquired so far here
    try {
        init(pool_paths);
    } catch (...) {
        release();
        throw;
    }
}

void Renderer::init(uint32_t pool_paths)
{
    const int device = device_;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) throw std::runtime_error("no CUDA device: this backend has no CPU fallback");
    if (device < 0 || device >= n_dev) throw std::runtime_error("CUDA device ordinal out of range");
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    impl_->sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("ARTB_MAX_SHARED_NODES")) impl_->max_shared_nodes = size_t(atoll(e));   // tuning knob (bytes)
    impl_->max_shared_nodes = std::min<size_t>(impl_->max_shared_nodes, size_t(prop.sharedMemPerBlockOptin) - (48u << 10));
    // the tree walk stages up to max_shared_nodes bytes of nodes in dynamic shared memory: permit that once per device
    // static shared memory of the kernel: the warp feeds (1 KB per warp) + a few words
    impl_->max_traverse_smem = size_t(prop.sharedMemPerBlockOptin) - (ARTB_TRAVERSE_THREADS / 32) * sizeof(WarpFeed) - 1024;
    CUDA_CHECK(cudaFuncSetAttribute(k_traverse<true, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(impl_->max_shared_nodes)));
    CUDA_CHECK(cudaFuncSetAttribute(k_traverse<true, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(impl_->max_shared_nodes)));
#if ARTB_SMEM_STACK > 0
    CUDA_CHECK(cudaFuncSetAttribute(k_traverse<true, false, ARTB_SMEM_STACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(impl_->max_traverse_smem)));
    CUDA_CHECK(cudaFuncSetAttribute(k_traverse<true, true, ARTB_SMEM_STACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(impl_->max_traverse_smem)));
#endif
    if (const char* e = getenv("ARTB_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, size_t(atoll(e)));
    cudaStream_t st;
    CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    stream_ = st;

    if (pool_paths == 0) pool_paths = 1u << 21;
    pool_paths = (pool_paths + 31u) & ~31u;
    WaveBuffers& w = impl_->wb;
    w.pool = pool_paths;
    char* recs = dev_alloc<char>(impl_->pool_allocs, size_t(pool_paths) * ARTB_PATH_STRIDE);
    w.rays = { recs, ARTB_PATH_STRIDE };
    w.aux = { recs + sizeof(PathRay), ARTB_PATH_STRIDE };
    w.hits = { recs + sizeof(PathRay) + sizeof(PathAux), ARTB_PATH_STRIDE };
    w.ties = { recs + sizeof(PathRay) + sizeof(PathAux) + sizeof(PathHit), ARTB_PATH_STRIDE };
    for (int k = 0; k < 2; ++k) {
        w.q_extend[k] = dev_alloc<uint32_t>(impl_->pool_allocs, pool_paths);
        w.q_regen[k] = dev_alloc<uint32_t>(impl_->pool_allocs, pool_paths);
    }
    w.q_shade = dev_alloc<uint32_t>(impl_->pool_allocs, size_t(pool_paths) * MAT_COUNT);
    w.q_overflow = dev_alloc<uint32_t>(impl_->pool_allocs, pool_paths);
    w.pair_capacity = size_t(pool_paths) * ARTB_PAIRS_PER_RAY;
    if (const char* e = getenv("ARTB_PAIR_CAPACITY")) w.pair_capacity = std::max<size_t>(size_t(atoll(e)), 1);   // test knob: forces the overflow pass
    w.pairs = dev_alloc<uint2>(impl_->pool_allocs, w.pair_capacity);
    w.region_count = dev_alloc<uint32_t>(impl_->pool_allocs, ARTB_MAX_REGIONS);
    w.c = dev_alloc<WaveCounters>(impl_->pool_allocs, 1);
    CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&impl_->counters_host), 2 * sizeof(WaveCounters)));
    for (int k = 0; k < 2; ++k) CUDA_CHECK(cudaEventCreateWithFlags(&impl_->poll_ev[k], cudaEventDisableTiming));
}

Renderer::~Renderer() { release(); }

void Renderer::release()
{
    if (!impl_) return;
    cudaSetDevice(device_);
    cudaDeviceSynchronize();
    for (void* p : impl_->scene_allocs) cudaFree(p);
    for (void* p : impl_->pool_allocs) cudaFree(p);
    if (impl_->list[0]) cudaFree(impl_->list[0]);
    if (impl_->list[1]) cudaFree(impl_->list[1]);
    if (impl_->lumsq) cudaFree(impl_->lumsq);
    if (impl_->host_accum) cudaFree(impl_->host_accum);
    if (impl_->host_count) cudaFree(impl_->host_count);
    if (impl_->host_rgb) cudaFree(impl_->host_rgb);
    if (impl_->counters_host) cudaFreeHost(impl_->counters_host);
    for (int k = 0; k < 2; ++k) if (impl_->poll_ev[k]) cudaEventDestroy(impl_->poll_ev[k]);
    for (void* p : impl_->batch.bag) cudaFree(p);
    if (impl_->batch.e0) cudaEventDestroy(impl_->batch.e0);
    if (impl_->batch.e1) cudaEventDestroy(impl_->batch.e1);
    for (cudaEvent_t e : impl_->ev) cudaEventDestroy(e);
    if (impl_->tile_list) cudaFree(impl_->tile_list);
    if (impl_->tile_offsets) cudaFree(impl_->tile_offsets);
    if (impl_->view_dev) cudaFree(impl_->view_dev);
    if (stream_) cudaStreamDestroy(static_cast<cudaStream_t>(stream_));
    stream_ = nullptr;
    delete impl_;
    impl_ = nullptr;
}

// the copy of view_ the out-of-line device functions read (TraceIO::scene_dev)
void Renderer::sync_view()
{
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (!impl_->view_dev) CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&impl_->view_dev), sizeof(SceneView)));
    view_.self = impl_->view_dev;
    CUDA_CHECK(cudaMemcpyAsync(impl_->view_dev, &view_, sizeof(SceneView), cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
}

void Renderer::upload(const FlatScene& fs)
{
    CUDA_CHECK(cudaSetDevice(device_));
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    // everything that can be rejected is rejected BEFORE the previous scene is touched ...
    if (fs.bvh_depth > ARTB_MAX_BVH8_DEPTH)
        throw std::runtime_error("scene BVH is deeper than the device traversal stack");
    CUDA_CHECK(cudaStreamSynchronize(st));
    // ... and from here on there is no scene until the new one is complete: a failing allocation / copy / device build
    // must not leave has_scene_ pointing at freed arrays
    has_scene_ = false;
    view_ = SceneView{};
    scene_bytes_ = 0;
    for (void* p : impl_->scene_allocs) cudaFree(p);
    impl_->scene_allocs.clear();
    last_build_ = BvhBuildStats{};
    uint64_t bytes = 0;
    SceneView v{};
    v.prims = dev_upload(impl_->scene_allocs, fs.prims, st, bytes);
    v.ref_nodes = fs.ref_nodes.empty() ? nullptr : dev_upload(impl_->scene_allocs, fs.ref_nodes, st, bytes);
    std::vector<uint8_t> prim_class(fs.prims.size());
    for (size_t i = 0; i < fs.prims.size(); ++i) prim_class[i] = uint8_t(prim_class_of(fs.prims[i].type, fs.prims[i].xform));
    v.prim_class = dev_upload(impl_->scene_allocs, prim_class, st, bytes);
    std::vector<uint8_t> prim_mat_type(fs.prims.size());
    for (size_t i = 0; i < fs.prims.size(); ++i)
        prim_mat_type[i] = fs.prims[i].material < fs.materials.size() ? uint8_t(fs.materials[fs.prims[i].material].type) : uint8_t(0);
    v.prim_mat_type = dev_upload(impl_->scene_allocs, prim_mat_type, st, bytes);
    v.nodes = dev_upload(impl_->scene_allocs, fs.nodes, st, bytes);
    v.xform_ops = dev_upload(impl_->scene_allocs, fs.xform_ops, st, bytes);
    v.xforms = dev_upload(impl_->scene_allocs, fs.xforms, st, bytes);
    v.materials = dev_upload(impl_->scene_allocs, fs.materials, st, bytes);
    v.textures = dev_upload(impl_->scene_allocs, fs.textures, st, bytes);
    v.images = dev_upload(impl_->scene_allocs, fs.images, st, bytes);
    v.texels = dev_upload(impl_->scene_allocs, fs.texels, st, bytes);
    v.perlins = dev_upload(impl_->scene_allocs, fs.perlins, st, bytes);
    v.n_prims = uint32_t(fs.prims.size());
    v.n_world_prims = fs.n_world_prims;
    v.n_nodes = uint32_t(fs.nodes.size());
    v.n_media = fs.n_media;
    v.n_xforms = uint32_t(fs.xforms.size());
    v.n_materials = uint32_t(fs.materials.size());
    v.n_textures = uint32_t(fs.textures.size());
    v.n_global_prims = fs.n_global_prims;
    v.n_global_media = fs.n_global_media;
    v.one_bits = 0x3f800000u;
    for (int k = 0; k < 3; ++k) v.background[k] = fs.background[k];
    CUDA_CHECK(cudaStreamSynchronize(st));
    view_ = v;
    scene_bytes_ = bytes;
    has_scene_ = true;
    sync_view();
    impl_->bvh_depth = fs.bvh_depth;
    max_dfs_id_ = 0;
    for (uint32_t i = 0; i < fs.n_world_prims; ++i) max_dfs_id_ = std::max(max_dfs_id_, fs.prims[i].dfs_id);
    // a flat scene that arrives without a tree (FlattenOptions::device_bvh, or a foreign host that has none) gets
    // one built here, on the device
    if (fs.nodes.empty() && fs.n_world_prims > fs.n_global_prims) {
        try {
            build_bvh_device(fs.device_bvh_leaf_prims, &last_build_);
        } catch (...) {
            has_scene_ = false;          // primitives without a tree cannot be rendered
            throw;
        }
    }
}

float Renderer::trace_batch_device(const double* rays_dev, uint64_t n, double t_min, double t_max, const double* xi_dev,
                                   HitRecordOut* out_dev)
{
    if (!has_scene_) throw std::runtime_error("trace_batch: no scene uploaded");
    CUDA_CHECK(cudaSetDevice(device_));
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (n == 0) return 0.f;
    Impl& I = *impl_;
    Impl::BatchScratch& B = I.batch;
    const uint64_t kChunk = 1u << 22;                         // rays per pass
    const uint64_t cap = std::min<uint64_t>(n, kChunk);
    size_t pair_capacity = std::max<size_t>(size_t(cap) * ARTB_PAIRS_PER_RAY, size_t(ARTB_MAX_REGIONS) * 8192);
    if (const char* e = getenv("ARTB_PAIR_CAPACITY")) pair_capacity = std::max<size_t>(size_t(atoll(e)), 1);
    if (!B.e0) { CUDA_CHECK(cudaEventCreate(&B.e0)); CUDA_CHECK(cudaEventCreate(&B.e1)); }
    if (B.cap < cap || B.pair_capacity != pair_capacity) {
        CUDA_CHECK(cudaStreamSynchronize(st));
        for (void* p : B.bag) cudaFree(p);
        B.bag.clear();
        const uint64_t want = std::max(cap, B.cap);
        B.cap = 0; B.pair_capacity = 0;
        B.hits = dev_alloc<PathHit>(B.bag, want);
        B.ties = dev_alloc<PathTie>(B.bag, want);
        B.pairs = dev_alloc<uint2>(B.bag, pair_capacity);
        B.region_count = dev_alloc<uint32_t>(B.bag, ARTB_MAX_REGIONS);
        B.overflow = dev_alloc<uint32_t>(B.bag, want);
        B.ctr = dev_alloc<uint32_t>(B.bag, 4);
        B.overflow_total = dev_alloc<unsigned long long>(B.bag, 1);
        B.cap = want; B.pair_capacity = pair_capacity;
    }
    // the pair kernels when the tree fits in shared memory, else the pruned walk — the choice render_device makes
    bool two_phase = I.nodes_fit_shared(view_);
    if (const char* e = getenv("ARTB_TWO_PHASE")) two_phase = atoi(e) != 0;
    const bool ids_fit = view_.n_prims < ARTB_PAIR_MAX_ID;     // pack_pair holds 24-bit ids
    if (!ids_fit) two_phase = false;
    last_trace_overflow_ = 0;
    CUDA_CHECK(cudaMemsetAsync(B.overflow_total, 0, sizeof(unsigned long long), st));
    CUDA_CHECK(cudaEventRecord(B.e0, st));
    for (uint64_t first = 0; first < n; first += kChunk) {
        const uint32_t m = uint32_t(std::min<uint64_t>(kChunk, n - first));
        k_batch_counters<<<1, 1, 0, st>>>(B.ctr, m);
        TraceIO io{};
        io.n_ptr = B.ctr; io.list = nullptr; io.hits = { reinterpret_cast<char*>(B.hits), uint32_t(sizeof(PathHit)) };
        io.ties = { reinterpret_cast<char*>(B.ties), uint32_t(sizeof(PathTie)) };
        I.set_pair_buffer(view_, io, B.pairs, pair_capacity, B.region_count);
        io.cursor_traverse = B.ctr + 1; io.q_overflow = B.overflow; io.n_overflow = B.ctr + 3;
        io.t_min = t_min; io.t_max = t_max;
        io.batch_rays = rays_dev + first * 7;
        io.batch_xi = xi_dev ? xi_dev + first * view_.n_media : nullptr;
        io.batch_out = out_dev + first;
        I.launch_trace<true>(view_, io, I.wb, RoundParams{}, 0, st, nullptr, two_phase);
        // rays k_finish flagged (pairs did not fit, or more exact ties than the list holds): the one-phase walk
        k_trace_batch<<<I.sm_count * 8, 128, 0, st>>>(view_, io.batch_rays, B.overflow, 0, B.ctr + 3, B.overflow_total, t_min, t_max,
                                                      io.batch_xi, io.batch_out);
        CUDA_CHECK(cudaGetLastError());
    }
    CUDA_CHECK(cudaEventRecord(B.e1, st));
    unsigned long long n_over = 0;
    CUDA_CHECK(cudaMemcpyAsync(&n_over, B.overflow_total, sizeof(n_over), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    last_trace_overflow_ = n_over;
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, B.e0, B.e1));
    return ms;
}

void Renderer::trace_batch(const double* rays_host, uint64_t n, double t_min, double t_max, const double* xi_host,
                           HitRecordOut* out_host)
{
    if (!has_scene_) throw std::runtime_error("trace_batch: no scene uploaded");
    if (n == 0) return;
    CUDA_CHECK(cudaSetDevice(device_));
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    std::vector<void*> bag;
    try {
        double* d_rays = dev_alloc<double>(bag, n * 7);
        HitRecordOut* d_out = dev_alloc<HitRecordOut>(bag, n);
        double* d_xi = nullptr;
        CUDA_CHECK(cudaMemcpyAsync(d_rays, rays_host, n * 7 * sizeof(double), cudaMemcpyHostToDevice, st));
        if (xi_host && view_.n_media) {
            d_xi = dev_alloc<double>(bag, n * view_.n_media);
            CUDA_CHECK(cudaMemcpyAsync(d_xi, xi_host, n * view_.n_media * sizeof(double), cudaMemcpyHostToDevice, st));
        }
        trace_batch_device(d_rays, n, t_min, t_max, d_xi, d_out);
        CUDA_CHECK(cudaMemcpyAsync(out_host, d_out, n * sizeof(HitRecordOut), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
    } catch (...) {
        for (void* p : bag) cudaFree(p);
        throw;
    }
    for (void* p : bag) cudaFree(p);
}

void Renderer::shade_batch(const double* rays_host, const int32_t* prim_id_host, const double* t_host, const double* xi_host,
                           uint64_t n, ShadeRecordOut* out_host)
{
    if (!has_scene_) throw std::runtime_error("shade_batch: no scene uploaded");
    if (n == 0) return;
    CUDA_CHECK(cudaSetDevice(device_));
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    std::vector<void*> bag;
    try {
        double* d_rays = dev_alloc<double>(bag, n * 7);
        int32_t* d_id = dev_alloc<int32_t>(bag, n);
        double* d_t = dev_alloc<double>(bag, n);
        double* d_xi = dev_alloc<double>(bag, n * 4);
        ShadeRecordOut* d_out = dev_alloc<ShadeRecordOut>(bag, n);
        const uint32_t map_size = max_dfs_id_ + 1u;
        int32_t* d_map = dev_alloc<int32_t>(bag, map_size);
        CUDA_CHECK(cudaMemcpyAsync(d_rays, rays_host, n * 7 * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaMemcpyAsync(d_id, prim_id_host, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaMemcpyAsync(d_t, t_host, n * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaMemcpyAsync(d_xi, xi_host, n * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaMemsetAsync(d_map, 0xff, size_t(map_size) * sizeof(int32_t), st));
        k_dfs_map<<<(view_.n_world_prims + 255) / 256, 256, 0, st>>>(view_.prims, view_.n_world_prims, d_map, map_size);
        k_shade_batch<<<unsigned((n + 127) / 128), 128, 0, st>>>(view_, d_rays, d_id, d_t, d_xi, d_map, map_size, n, d_out);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaMemcpyAsync(out_host, d_out, n * sizeof(ShadeRecordOut), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
    } catch (...) {
        for (void* p : bag) cudaFree(p);
        throw;
    }
    for (void* p : bag) cudaFree(p);
}

void Renderer::camera_rays(const CameraParams& cam, const double* st_host, const double* xi_host, uint64_t n, double* rays_out_host)
{
    if (n == 0) return;
    CUDA_CHECK(cudaSetDevice(device_));
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    std::vector<void*> bag;
    try {
        double* d_st = dev_alloc<double>(bag, n * 2);
        double* d_xi = dev_alloc<double>(bag, n * 3);
        double* d_out = dev_alloc<double>(bag, n * 7);
        CUDA_CHECK(cudaMemcpyAsync(d_st, st_host, n * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_CHECK(cudaMemcpyAsync(d_xi, xi_host, n * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
        k_camera_rays<<<unsigned((n + 127) / 128), 128, 0, st>>>(cam, d_st, d_xi, n, d_out);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaMemcpyAsync(rays_out_host, d_out, n * 7 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
    } catch (...) {
        for (void* p : bag) cudaFree(p);
        throw;
    }
    for (void* p : bag) cudaFree(p);
}

void Renderer::resolve_device(const float* accum_dev, const uint32_t* count_dev, int n_pixels, int fixed_spp, uint8_t* rgb_dev)
{
    CUDA_CHECK(cudaSetDevice(device_));
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    k_resolve<<<(n_pixels + 255) / 256, 256, 0, st>>>(accum_dev, count_dev, n_pixels, fixed_spp, rgb_dev);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(st));
}

// ---- device BVH build (bvh_build.cuh) ------------------------------------------------------------------------------
__global__ void k_prim_tables(const FlatPrim* prims, const FlatMaterial* materials, uint32_t n_materials, uint32_t n, uint8_t* prim_class,
                              uint8_t* prim_mat_type)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const FlatPrim& p = prims[i];
    prim_class[i] = uint8_t(prim_class_of(p.type, p.xform));
    prim_mat_type[i] = p.material < n_materials ? uint8_t(materials[p.material].type) : uint8_t(0);
}

void Renderer::build_bvh_device(int max_leaf_prims, BvhBuildStats* stats)
{
    if (!has_scene_) throw std::runtime_error("build_bvh_device: no scene uploaded");
    if (max_leaf_prims < 1) max_leaf_prims = 1;
    if (max_leaf_prims > ARTB_MAX_LEAF_PRIMS) throw std::runtime_error("build_bvh_device: max_leaf_prims exceeds ARTB_MAX_LEAF_PRIMS");
    CUDA_CHECK(cudaSetDevice(device_));
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    CUDA_CHECK(cudaStreamSynchronize(st));
    BvhBuildStats S;
    const uint32_t first = view_.n_global_prims;
    const uint32_t n = view_.n_world_prims - view_.n_global_prims;
    S.n_prims = n;
    if (n == 0) { if (stats) *stats = S; return; }
    using namespace bvhb;
    std::vector<void*> bag;
    FlatNode8* nodes_tmp = nullptr;
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
    try {
        Arrays A{};
        const size_t n_bin = 2 * size_t(n);
        A.box_lo = dev_alloc<float4>(bag, n); A.box_hi = dev_alloc<float4>(bag, n);
        A.node_lo = dev_alloc<float4>(bag, n_bin); A.node_hi = dev_alloc<float4>(bag, n_bin);
        A.left = dev_alloc<int>(bag, n); A.right = dev_alloc<int>(bag, n); A.parent = dev_alloc<int>(bag, n_bin);
        A.first = dev_alloc<int>(bag, n); A.last = dev_alloc<int>(bag, n);
        A.flag = dev_alloc<unsigned int>(bag, n);
        A.keys = dev_alloc<unsigned long long>(bag, n); A.keys_sorted = dev_alloc<unsigned long long>(bag, n);
        A.idx = dev_alloc<uint32_t>(bag, n); A.idx_sorted = dev_alloc<uint32_t>(bag, n);
        A.bounds = dev_alloc<int>(bag, 6);
        A.perm = dev_alloc<uint32_t>(bag, n);
        A.queue[0] = dev_alloc<uint2>(bag, n); A.queue[1] = dev_alloc<uint2>(bag, n);
        A.counters = dev_alloc<uint32_t>(bag, 4);
        const uint32_t node_capacity = n;                       // wide nodes <= binary internal nodes + 1
        nodes_tmp = dev_alloc<FlatNode8>(bag, node_capacity);
        FlatPrim* prims_tmp = dev_alloc<FlatPrim>(bag, n);
        int* d_error = dev_alloc<int>(bag, 1);
        size_t sort_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, A.keys, A.keys_sorted, A.idx, A.idx_sorted, int(n), 0, 63, st);
        void* sort_tmp = dev_alloc<char>(bag, sort_bytes);
        FlatPrim* prims = const_cast<FlatPrim*>(view_.prims) + first;
        const unsigned B = 256, G = (n + B - 1) / B;

        CUDA_CHECK(cudaEventRecord(e0, st));
        CUDA_CHECK(cudaMemsetAsync(d_error, 0, sizeof(int), st));
        k_bvh_init<<<G, B, 0, st>>>(A, n);
        k_bvh_boxes<<<G, B, 0, st>>>(prims, n, view_.xform_ops, view_.xforms, A);
        k_bvh_morton<<<G, B, 0, st>>>(n, A);
        CUDA_CHECK(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, A.keys, A.keys_sorted, A.idx, A.idx_sorted, int(n), 0, 63, st));
        if (n > 1) k_bvh_karras<<<G, B, 0, st>>>(int(n), A);
        else CUDA_CHECK(cudaMemsetAsync(A.parent, 0xff, sizeof(int), st));      // a single leaf: node (n-1)+0 = 0, no parent
        k_bvh_refit<<<G, B, 0, st>>>(int(n), A);
        // level by level; the host only reads the length of the next level's queue
        uint32_t n_jobs = 1;
        int level = 0;
        while (n_jobs) {
            if (++S.depth > ARTB_MAX_BVH8_DEPTH) throw std::runtime_error("build_bvh_device: tree deeper than the traversal stack allows");
            k_bvh_next_level<<<1, 1, 0, st>>>(A);
            k_bvh_collapse<<<(n_jobs + 127) / 128, 128, 0, st>>>(int(n), int(n_jobs), level, max_leaf_prims, A, nodes_tmp, node_capacity, d_error);
            uint32_t ctr[4];
            CUDA_CHECK(cudaMemcpyAsync(ctr, A.counters, sizeof(ctr), cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(cudaStreamSynchronize(st));
            n_jobs = ctr[2];
            S.n_nodes = ctr[0];
            level ^= 1;
        }
        int err = 0;
        CUDA_CHECK(cudaMemcpyAsync(&err, d_error, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        if (err) throw std::runtime_error("build_bvh_device: quantisation / capacity error " + std::to_string(err));
        // primitives into leaf order, then the per-primitive tables and the final node array
        k_bvh_permute<<<(n * 6 + B - 1) / B, B, 0, st>>>(prims, prims_tmp, A.perm, n);
        CUDA_CHECK(cudaMemcpyAsync(prims, prims_tmp, size_t(n) * sizeof(FlatPrim), cudaMemcpyDeviceToDevice, st));
        k_bvh_offset_prim_base<<<(S.n_nodes + B - 1) / B, B, 0, st>>>(nodes_tmp, S.n_nodes, first);
        FlatNode8* nodes = dev_alloc<FlatNode8>(impl_->scene_allocs, S.n_nodes);
        CUDA_CHECK(cudaMemcpyAsync(nodes, nodes_tmp, size_t(S.n_nodes) * sizeof(FlatNode8), cudaMemcpyDeviceToDevice, st));
        k_prim_tables<<<(view_.n_prims + B - 1) / B, B, 0, st>>>(view_.prims, view_.materials, view_.n_materials, view_.n_prims,
                                                                 const_cast<uint8_t*>(view_.prim_class), const_cast<uint8_t*>(view_.prim_mat_type));
        CUDA_CHECK(cudaEventRecord(e1, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        CUDA_CHECK(cudaGetLastError());
        float ms = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        S.total_ms = ms;
        view_.nodes = nodes;                 // the previous node array stays in scene_allocs until the next upload
        view_.n_nodes = S.n_nodes;
        impl_->bvh_depth = S.depth;
        sync_view();
    } catch (...) {
        for (void* p : bag) cudaFree(p);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        throw;
    }
    for (void* p : bag) cudaFree(p);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (stats) *stats = S;
}

void Renderer::render_device(const CameraParams& cam, const RenderParams& p, float* accum_dev, uint32_t* count_dev,
                             RenderStats* stats, int profile_every)
{
    if (!has_scene_) throw std::runtime_error("render: no scene uploaded");
    if (p.width < 2 || p.height < 2 || p.samples_per_pixel < 1 || p.max_depth < 1)
        throw std::runtime_error("render: bad image size / spp / depth");
    if (p.adaptive == 2) throw std::runtime_error("render_device: the adaptive-squares mode produces 8-bit pixels, use render_host");
    if (p.adaptive && !count_dev) throw std::runtime_error("render: adaptive mode needs a count buffer");
    CUDA_CHECK(cudaSetDevice(device_));
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    Impl& I = *impl_;
    WaveBuffers& w = I.wb;
    const size_t n_pixels = size_t(p.width) * size_t(p.height);

    // scratch for pixel lists
    if (I.list_capacity < n_pixels) {
        for (int k = 0; k < 2; ++k) { if (I.list[k]) cudaFree(I.list[k]); I.list[k] = nullptr; }
        if (I.tile_list) { cudaFree(I.tile_list); I.tile_list = nullptr; }
        I.tile_key[4] = -1;
        if (I.lumsq) { cudaFree(I.lumsq); I.lumsq = nullptr; }
        for (int k = 0; k < 2; ++k) CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&I.list[k]), n_pixels * sizeof(uint32_t)));
        CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&I.tile_list), n_pixels * sizeof(uint32_t)));
        CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&I.lumsq), n_pixels * sizeof(float)));
        I.list_capacity = n_pixels;
    }

    // partition B: the pixels of this rank's tiles
    const uint32_t* base_list = nullptr;
    size_t n_local = n_pixels;
    if (p.pixel_list_dev) {            // internal: an explicit list of pixels (adaptive squares)
        base_list = p.pixel_list_dev;
        n_local = p.n_pixel_list;
    } else if (p.tile_parts > 1) {
        // built on the device and kept: the same (size, tiling, part) comes back on every step of a multi-GPU render
        const int ts = std::max(p.tile_size, 1);
        const int key[5] = { p.width, p.height, ts, p.tile_parts, p.tile_part };
        if (std::memcmp(key, I.tile_key, sizeof(key)) != 0) {
            const int tiles_x = (p.width + ts - 1) / ts, tiles_y = (p.height + ts - 1) / ts;
            const long long n_tiles = (long long)tiles_x * tiles_y;
            const long long n_owned = p.tile_part < n_tiles ? (n_tiles - 1 - p.tile_part) / p.tile_parts + 1 : 0;
            std::vector<uint32_t> offsets(size_t(std::max<long long>(n_owned, 1)));
            size_t count = 0;               // pixels of the owned tiles (per-tile arithmetic, no per-pixel loop)
            for (long long k = 0; k < n_owned; ++k) {
                const long long t = p.tile_part + k * p.tile_parts;
                const int tx = int(t % tiles_x), ty = int(t / tiles_x);
                offsets[size_t(k)] = uint32_t(count);
                count += size_t(std::min(ts, p.width - tx * ts)) * size_t(std::min(ts, p.height - ty * ts));
            }
            if (n_owned) {
                if (I.tile_offsets_capacity < size_t(n_owned)) {
                    if (I.tile_offsets) { cudaFree(I.tile_offsets); I.tile_offsets = nullptr; I.tile_offsets_capacity = 0; }
                    CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&I.tile_offsets), size_t(n_owned) * sizeof(uint32_t)));
                    I.tile_offsets_capacity = size_t(n_owned);
                }
                CUDA_CHECK(cudaMemcpyAsync(I.tile_offsets, offsets.data(), size_t(n_owned) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
                k_tile_list<<<unsigned(n_owned), 256, 0, st>>>(I.tile_list, I.tile_offsets, p.width, p.height, ts, p.tile_parts, p.tile_part);
                CUDA_CHECK(cudaGetLastError());
                CUDA_CHECK(cudaStreamSynchronize(st));      // `offsets` lives on this frame (once per key, not per call)
            }
            std::memcpy(I.tile_key, key, sizeof(key));
            I.tile_count = count;
        }
        n_local = I.tile_count;
        base_list = I.tile_list;
    }

    if (!p.accumulate) {
        CUDA_CHECK(cudaMemsetAsync(accum_dev, 0, n_pixels * 3 * sizeof(float), st));
        if (count_dev) CUDA_CHECK(cudaMemsetAsync(count_dev, 0, n_pixels * sizeof(uint32_t), st));
    }
    if (p.adaptive) CUDA_CHECK(cudaMemsetAsync(I.lumsq, 0, n_pixels * sizeof(float), st));

    RoundParams rp{};
    rp.cam = cam;
    rp.width = p.width; rp.height = p.height; rp.max_depth = p.max_depth;
    rp.seed = p.seed; rp.stream = p.stream;
    rp.accum = accum_dev;
    rp.count = p.adaptive ? count_dev : nullptr;
    rp.lumsq = p.adaptive ? I.lumsq : nullptr;

    const int shade_blocks = I.sm_count * 8;
    // two kernels per phase when the tree fits in shared memory, else the pruned one-phase walk (see k_extend)
    bool two_phase = I.nodes_fit_shared(view_);
    if (const char* e = getenv("ARTB_TWO_PHASE")) two_phase = atoi(e) != 0;
    if (w.pool >= ARTB_PAIR_MAX_ID || view_.n_prims >= ARTB_PAIR_MAX_ID) two_phase = false;     // pack_pair: 24-bit ids
    bool use_walk = true;               // big trees: k_walk + k_finish; ARTB_WALK=0 falls back to the chunked k_extend
    if (const char* e = getenv("ARTB_WALK")) use_walk = atoi(e) != 0;
    RenderStats S;
    struct EventPair {            // destroyed on every exit, including the CUDA_CHECK throws below
        cudaEvent_t a = nullptr, b = nullptr;
        ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } evp;
    CUDA_CHECK(cudaEventCreate(&evp.a)); CUDA_CHECK(cudaEventCreate(&evp.b));
    const cudaEvent_t ev_begin = evp.a, ev_end = evp.b;
    // per-wave events around the kernels of the closest-hit stage (roofline denominator), drained whenever the pool
    // fills up: [5k] before k_traverse, [5k+1] after it, [5k+2] after k_pairs, [5k+3] after k_finish, [5k+4] after k_extend
    const size_t kPairs = 1024;
    if (profile_every > 0 && I.ev.size() < kPairs * 5) {
        I.ev.resize(kPairs * 5);
        for (auto& e : I.ev) CUDA_CHECK(cudaEventCreate(&e));
    }
    size_t pairs_used = 0;
    auto drain = [&]() {
        CUDA_CHECK(cudaStreamSynchronize(st));
        for (size_t k = 0; k < pairs_used; ++k) {
            float ms = 0.f;
            CUDA_CHECK(cudaEventElapsedTime(&ms, I.ev[5 * k], I.ev[5 * k + 4]));
            S.extend_ms += ms;
            {   // k_traverse (or k_walk) | k_pairs | k_finish; all three events are recorded in every mode
                CUDA_CHECK(cudaEventElapsedTime(&ms, I.ev[5 * k], I.ev[5 * k + 1]));
                S.traverse_ms += ms;
                CUDA_CHECK(cudaEventElapsedTime(&ms, I.ev[5 * k + 1], I.ev[5 * k + 2]));
                S.pairs_ms += ms;
                CUDA_CHECK(cudaEventElapsedTime(&ms, I.ev[5 * k + 2], I.ev[5 * k + 3]));
                S.finish_ms += ms;
            }
        }
        S.extend_launches += pairs_used;
        pairs_used = 0;
    };

    CUDA_CHECK(cudaEventRecord(ev_begin, st));
    k_init_pool<<<(w.pool + 255) / 256, 256, 0, st>>>(w);
    S.launches++;

    // One wave = k_shade(cur) [shade the hits of the previous wave + start new paths in free slots] followed by
    // k_extend(cur^1) [closest hits of everything k_shade queued]. The host only looks at the counters every few waves.
    int cur = 0;
    auto run_round = [&](long long n_list_fixed, const uint32_t* list, uint32_t first_sample, uint32_t batch) {
        rp.pixel_list = list;
        rp.first_sample = first_sample;
        rp.batch = batch;
        k_round_begin<<<1, 1, 0, st>>>(w.c, n_list_fixed, batch);
        S.launches++;
        // The host looks at the counters ONE BATCH BEHIND: after queueing batch b it copies the counters (as they will be
        // once b has run) and immediately queues batch b+1, then waits for the copy of batch b. The GPU therefore never
        // idles while the host wakes up, reads and launches (with several ranks on one box that wake-up is not prompt);
        // the price is at most one batch of empty waves at the end of a round (a few launches that find nothing to do).
        int check_every = 4, snap = 0;
        bool have_prev = false;
        auto queue_batch = [&](int n_waves) {
            for (int k = 0; k < n_waves; ++k) {
                k_shade<<<shade_blocks, 128, 0, st>>>(w, view_, rp, cur);
                cur ^= 1;
                // profile = N: every N-th wave carries the events (five records per wave cost 1.7 % of a render when every
                // wave has them; the per-kernel averages of a systematic sample are the same)
                const bool profile = profile_every > 0 && S.waves % (unsigned long long)profile_every == 0;
                if (profile) {
                    if (pairs_used == kPairs) drain();
                    CUDA_CHECK(cudaEventRecord(I.ev[5 * pairs_used], st));
                }
                if (two_phase || use_walk) {
                    TraceIO io{};
                    io.n_ptr = &w.c->n_extend[cur]; io.list = w.q_extend[cur]; io.hits = w.hits; io.ties = w.ties;
                    I.set_pair_buffer(view_, io, w.pairs, w.pair_capacity, w.region_count);
                    io.cursor_traverse = &w.c->cursor[0];
                    io.q_overflow = w.q_overflow; io.n_overflow = &w.c->n_overflow;
                    io.t_min = 0.001; io.t_max = INFINITY;
                    io.rays = w.rays; io.aux = w.aux;
                                I.launch_trace<false>(view_, io, w, rp, cur, st, profile ? &I.ev[5 * pairs_used + 1] : nullptr, two_phase);
                    // overflow pass: only the pair kernels can flag rays (full regions, more exact ties than the list holds)
                    if (two_phase) k_extend<<<I.sm_count * 4, ARTB_EXTEND_THREADS, 0, st>>>(w, view_, rp, cur, 0);
                    S.launches += two_phase ? 5 : 3;
                } else {
                    if (profile) for (int e = 1; e <= 3; ++e) CUDA_CHECK(cudaEventRecord(I.ev[5 * pairs_used + e], st));
                    k_extend<<<I.sm_count * 4, ARTB_EXTEND_THREADS, 0, st>>>(w, view_, rp, cur, 1);
                    S.launches += 2;
                }
                if (profile) { CUDA_CHECK(cudaEventRecord(I.ev[5 * pairs_used + 4], st)); pairs_used++; }
                S.waves++;
            }
            CUDA_CHECK(cudaMemcpyAsync(&I.counters_host[snap], w.c, sizeof(WaveCounters), cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(cudaEventRecord(I.poll_ev[snap], st));
            CUDA_CHECK(cudaGetLastError());
        };
        // a snapshot says "done" when every sample of the round was issued and no hit waits to be shaded; `parity` is the
        // value `cur` had when the snapshot was queued (the shade queues are double-buffered by it)
        auto done = [&](int which, int parity) {
            CUDA_CHECK(cudaEventSynchronize(I.poll_ev[which]));
            const WaveCounters& c = I.counters_host[which];
            uint32_t pending = 0;
            for (int m = 0; m < MAT_COUNT; ++m) pending += c.n_out[parity][m];
            return c.issued >= c.total && pending == 0;
        };
        int prev_parity = 0;
        while (true) {
            queue_batch(check_every);
            const int this_snap = snap, this_parity = cur;
            snap ^= 1;
            if (have_prev && done(this_snap ^ 1, prev_parity)) {
                // the batch just queued runs on an empty pool; wait for it so that the round really is over
                CUDA_CHECK(cudaEventSynchronize(I.poll_ev[this_snap]));
                I.counters_host[0] = I.counters_host[this_snap];
                break;
            }
            have_prev = true;
            prev_parity = this_parity;
            if (check_every < 8) check_every *= 2;        // polling costs nothing now; a short batch bounds the empty tail
        }
    };

    if (!p.adaptive) {
        run_round((long long)n_local, base_list, uint32_t(p.sample_begin), uint32_t(p.samples_per_pixel));
        S.samples = (unsigned long long)n_local * (unsigned long long)p.samples_per_pixel;
        if (count_dev && n_local) {
            k_count_fixed<<<unsigned((n_local + 255) / 256), 256, 0, st>>>(base_list, uint32_t(n_local), uint32_t(p.samples_per_pixel), count_dev);
            S.launches++;
        }
    } else {
        // round 0: every local pixel gets min_spp samples; then only the unconverged ones are re-queued
        const uint32_t max_spp = uint32_t(p.samples_per_pixel);
        uint32_t first = uint32_t(p.sample_begin);
        uint32_t batch = uint32_t(std::min(p.adaptive_min_spp, p.samples_per_pixel));
        const uint32_t* list = base_list;
        long long n_fixed = (long long)n_local;
        int flip = 0;
        unsigned long long issued_total = 0;
        while (true) {
            run_round(n_fixed, list, first, batch);
            issued_total += I.counters_host->total;
            first += batch;
            if (first - uint32_t(p.sample_begin) >= max_spp) break;
            k_adaptive_update<<<I.sm_count * 4, 256, 0, st>>>(w.c, list, I.list[flip], accum_dev, count_dev, I.lumsq, p.width, p.height,
                                                              uint32_t(p.adaptive_min_spp), max_spp, float(p.adaptive_threshold));
            S.launches++;
            CUDA_CHECK(cudaMemcpyAsync(I.counters_host, w.c, sizeof(WaveCounters), cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(cudaStreamSynchronize(st));
            const uint32_t remaining = I.counters_host->n_next_list;
            if (remaining == 0) break;
            list = I.list[flip];
            flip ^= 1;
            n_fixed = -1;
            batch = uint32_t(std::min<long long>(p.adaptive_batch_spp, (long long)max_spp - (long long)(first - uint32_t(p.sample_begin))));
            S.pixels_active = remaining;
        }
        S.samples = issued_total;
        S.pixels_converged = n_local - S.pixels_active;
    }

    CUDA_CHECK(cudaEventRecord(ev_end, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (profile_every > 0) drain();
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, ev_begin, ev_end));
    S.elapsed_ms = ms;
    CUDA_CHECK(cudaMemcpy(I.counters_host, w.c, sizeof(WaveCounters), cudaMemcpyDeviceToHost));
    S.rays = I.counters_host->rays;
    S.overflow_rays = I.counters_host->overflow_total;
    if (stats) *stats = S;
}

// joins whatever was started, also when starting the next thread throws (a joinable std::thread must not be destroyed)
struct ThreadGroup {
    std::vector<std::thread> threads;
    ~ThreadGroup() { join(); }
    void join() { for (auto& t : threads) if (t.joinable()) t.join(); }
};

// ---- several GPUs in one process: the exchange -----------------------------------------------------------------------
// dst += the frames of up to 7 peers, read in place over NVLink (peer access) — the exchange of engine_mode::
// parallel_images (engine.h:424-440: the four workers' float frames summed pixel by pixel) fused with nothing else because
// nothing else follows it but write_color. 16-byte loads; the tail is handled scalar.
#define ARTB_MAX_GROUP 8
struct PeerFrames { const float* accum[ARTB_MAX_GROUP - 1]; const uint32_t* count[ARTB_MAX_GROUP - 1]; int n; };

__global__ void k_sum_peer_frames(float* accum, uint32_t* count, PeerFrames pf, size_t n_floats, size_t n_counts)
{
    const size_t stride = size_t(gridDim.x) * blockDim.x, tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t n4 = n_floats / 4;
    for (size_t i = tid; i < n4; i += stride) {
        float4 a = reinterpret_cast<float4*>(accum)[i];
        for (int k = 0; k < pf.n; ++k) {
            const float4 b = reinterpret_cast<const float4*>(pf.accum[k])[i];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        reinterpret_cast<float4*>(accum)[i] = a;
    }
    for (size_t i = n4 * 4 + tid; i < n_floats; i += stride) {
        float a = accum[i];
        for (int k = 0; k < pf.n; ++k) a += pf.accum[k][i];
        accum[i] = a;
    }
    if (count)
        for (size_t i = tid; i < n_counts; i += stride) {
            uint32_t c = count[i];
            for (int k = 0; k < pf.n; ++k) c += pf.count[k][i];
            count[i] = c;
        }
}

// Frames of the peers as device `self` can read them: in place where peer access exists (or the peer is the same GPU),
// else staged on `self` (copies queued on s0, buffers appended to `staged` for the caller to free).
static PeerFrames reachable_frames(int self, cudaStream_t s0, Renderer* const* peers, int n_peers, float* const* accum,
                                   uint32_t* const* count, size_t n_pixels, bool with_count, std::vector<void*>& staged)
{
    PeerFrames pf{};
    pf.n = n_peers;
    for (int k = 0; k < n_peers; ++k) {
        const int pd = peers[k]->device();
        int can = pd == self ? 1 : 0;
        if (pd != self) CUDA_CHECK(cudaDeviceCanAccessPeer(&can, self, pd));
        if (can) {
            if (pd != self) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(pd, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_CHECK(e);
                cudaGetLastError();
            }
            pf.accum[k] = accum[k]; pf.count[k] = with_count ? count[k] : nullptr;
        } else {
            float* a = dev_alloc<float>(staged, n_pixels * 3);
            CUDA_CHECK(cudaMemcpyPeerAsync(a, self, accum[k], pd, n_pixels * 3 * sizeof(float), s0));
            pf.accum[k] = a; pf.count[k] = nullptr;
            if (with_count) {
                uint32_t* c = dev_alloc<uint32_t>(staged, n_pixels);
                CUDA_CHECK(cudaMemcpyPeerAsync(c, self, count[k], pd, n_pixels * sizeof(uint32_t), s0));
                pf.count[k] = c;
            }
        }
    }
    return pf;
}

// engine_mode::adaptive of the reference (engine.h:151-333), level by level (adaptive_squares.cuh)
void Renderer::render_adaptive_squares(const CameraParams& cam, const RenderParams& p, uint8_t* rgb_dev, float* accum_dev,
                                       uint32_t* count_dev, RenderStats* stats, Renderer* const* peers, int n_peers)
{
    if (n_peers < 0 || n_peers > ARTB_MAX_GROUP - 1) throw std::runtime_error("render: at most 8 devices per group");
    if (!has_scene_) throw std::runtime_error("render: no scene uploaded");
    constexpr int kBig = 12;
    if (p.width % kBig != 0 || p.height % kBig != 0 || p.width > 65535 || p.height > 65535)   // engine.h:182-183
        throw std::logic_error("for adaptive strategy image size should perfectly fit big square size for now!!");
    CUDA_CHECK(cudaSetDevice(device_));
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    using namespace asq;
    const size_t n_pixels = size_t(p.width) * size_t(p.height);
    const uint32_t n_big = uint32_t((p.width / kBig) * (p.height / kBig));
    std::vector<void*> bag;
    RenderStats S;
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
    try {
        int* work = dev_alloc<int>(bag, n_pixels * 3);
        Lists L{};
        L.squares[0] = dev_alloc<uint32_t>(bag, size_t(n_big) * 16);
        L.squares[1] = dev_alloc<uint32_t>(bag, size_t(n_big) * 16);
        L.pixels = dev_alloc<uint32_t>(bag, n_pixels);
        L.counters = dev_alloc<uint32_t>(bag, 2);
        CUDA_CHECK(cudaEventRecord(e0, st));
        CUDA_CHECK(cudaMemsetAsync(accum_dev, 0, n_pixels * 3 * sizeof(float), st));
        k_sq_init<<<unsigned((n_pixels + 255) / 256), 256, 0, st>>>(work, count_dev, int(n_pixels));
        k_sq_level0<<<(n_big + 127) / 128, 128, 0, st>>>(L, p.width, p.height, kBig);
        uint32_t n_sq = n_big, n_px = n_big * 4;
        int level = 0;
        RenderParams q = p;
        q.adaptive = 0; q.accumulate = 1; q.tile_parts = 1; q.tile_part = 0;
        for (int size = kBig; ; size /= 2) {
            // trace the pixels this level asked for, at full spp
            if (n_px && n_peers > 0) {
                // this level's pixels, split evenly over the group: every device traces its slice at full spp (the peers
                // into their own zeroed frames), then the peers' frames are added to this one over peer memory
                CUDA_CHECK(cudaStreamSynchronize(st));                     // the list is complete
                const int n = n_peers + 1;
                std::vector<RenderStats> rs(static_cast<size_t>(n));
                std::vector<std::string> errors(static_cast<size_t>(n));
                std::vector<float*> pacc(static_cast<size_t>(n_peers), nullptr);
                std::vector<uint32_t*> pcnt(static_cast<size_t>(n_peers), nullptr);
                ThreadGroup workers;
                for (int k = 0; k < n; ++k) {
                    workers.threads.emplace_back([&, k]() {
                        try {
                            const uint32_t b = uint32_t(uint64_t(n_px) * uint64_t(k) / uint64_t(n));
                            const uint32_t e = uint32_t(uint64_t(n_px) * uint64_t(k + 1) / uint64_t(n));
                            RenderParams qq = q;
                            qq.n_pixel_list = e - b;
                            if (k == 0) {
                                qq.pixel_list_dev = L.pixels + b;
                                if (e > b) render_device(cam, qq, accum_dev, nullptr, &rs[0], 0);
                                return;
                            }
                            Renderer* pr = peers[k - 1];
                            pr->ensure_frame(n_pixels, &pacc[size_t(k - 1)], &pcnt[size_t(k - 1)], nullptr);
                            CUDA_CHECK(cudaSetDevice(pr->device()));
                            // the slice of the list travels into the peer's (otherwise unused) count frame
                            uint32_t* list_k = pcnt[size_t(k - 1)];
                            if (e > b) CUDA_CHECK(cudaMemcpyPeer(list_k, pr->device(), L.pixels + b, device_, size_t(e - b) * sizeof(uint32_t)));
                            qq.pixel_list_dev = list_k;
                            qq.accumulate = 0;
                            if (e > b) pr->render_device(cam, qq, pacc[size_t(k - 1)], nullptr, &rs[size_t(k)], 0);
                            else CUDA_CHECK(cudaMemset(pacc[size_t(k - 1)], 0, n_pixels * 3 * sizeof(float)));
                        } catch (const std::exception& ex) {
                            errors[size_t(k)] = ex.what();
                        } catch (...) {
                            errors[size_t(k)] = "unknown error";
                        }
                    });
                }
                workers.join();
                for (int k = 0; k < n; ++k)
                    if (!errors[size_t(k)].empty()) throw std::runtime_error("adaptive squares, device share " + std::to_string(k) + ": " + errors[size_t(k)]);
                CUDA_CHECK(cudaSetDevice(device_));
                std::vector<void*> staged;
                try {
                    const PeerFrames pf = reachable_frames(device_, st, peers, n_peers, pacc.data(), pcnt.data(), n_pixels, false, staged);
                    k_sum_peer_frames<<<impl_->sm_count * 4, 256, 0, st>>>(accum_dev, nullptr, pf, n_pixels * 3, 0);
                    CUDA_CHECK(cudaGetLastError());
                    CUDA_CHECK(cudaStreamSynchronize(st));
                } catch (...) {
                    for (void* x : staged) cudaFree(x);
                    throw;
                }
                for (void* x : staged) cudaFree(x);
                for (const RenderStats& r : rs) { S.samples += r.samples; S.rays += r.rays; S.waves = std::max(S.waves, r.waves); S.launches += r.launches; }
                S.launches++;
                k_sq_commit<<<(n_px + 255) / 256, 256, 0, st>>>(L.pixels, n_px, accum_dev, p.samples_per_pixel, work, count_dev);
            } else if (n_px) {
                q.pixel_list_dev = L.pixels; q.n_pixel_list = n_px;
                RenderStats r;
                render_device(cam, q, accum_dev, nullptr, &r, 0);
                S.samples += r.samples; S.rays += r.rays; S.waves += r.waves; S.launches += r.launches;
                k_sq_commit<<<(n_px + 255) / 256, 256, 0, st>>>(L.pixels, n_px, accum_dev, p.samples_per_pixel, work, count_dev);
            }
            if (size < 3 || n_sq == 0) break;
            CUDA_CHECK(cudaMemsetAsync(L.counters, 0, 2 * sizeof(uint32_t), st));
            k_sq_decide<<<(n_sq + 127) / 128, 128, 0, st>>>(L, level, n_sq, size, work, p.width);
            uint32_t ctr[2];
            CUDA_CHECK(cudaMemcpyAsync(ctr, L.counters, sizeof(ctr), cudaMemcpyDeviceToHost, st));
            CUDA_CHECK(cudaStreamSynchronize(st));
            n_sq = ctr[0]; n_px = ctr[1];
            level ^= 1;
            S.launches += 2;
            if (size == 3) { n_sq = 0; }        // the 3x3 squares only ask for pixels; one more pass traces them
        }
        k_sq_final<<<unsigned((n_pixels * 3 + 255) / 256), 256, 0, st>>>(work, int(n_pixels * 3), rgb_dev);
        CUDA_CHECK(cudaEventRecord(e1, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        CUDA_CHECK(cudaGetLastError());
        float ms = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        S.elapsed_ms = ms;
        S.pixels_active = S.samples / uint64_t(std::max(p.samples_per_pixel, 1));     // pixels traced
        S.pixels_converged = n_pixels - S.pixels_active;                              // pixels interpolated
    } catch (...) {
        for (void* q : bag) cudaFree(q);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        throw;
    }
    for (void* q : bag) cudaFree(q);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (stats) *stats = S;
}

void Renderer::ensure_frame(size_t n_pixels, float** accum, uint32_t** count, uint8_t** rgb)
{
    CUDA_CHECK(cudaSetDevice(device_));
    Impl& I = *impl_;
    if (I.host_pixels < n_pixels) {
        if (I.host_accum) cudaFree(I.host_accum);
        if (I.host_count) cudaFree(I.host_count);
        if (I.host_rgb) cudaFree(I.host_rgb);
        I.host_accum = nullptr; I.host_count = nullptr; I.host_rgb = nullptr; I.host_pixels = 0;
        CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&I.host_accum), n_pixels * 3 * sizeof(float)));
        CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&I.host_count), n_pixels * sizeof(uint32_t)));
        CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&I.host_rgb), n_pixels * 3));
        I.host_pixels = n_pixels;
    }
    if (accum) *accum = I.host_accum;
    if (count) *count = I.host_count;
    if (rgb) *rgb = I.host_rgb;
}

void Renderer::render_host(const CameraParams& cam, const RenderParams& p, uint8_t* rgb_out, float* accum_out,
                           uint32_t* count_out, RenderStats* stats)
{
    const size_t n_pixels = size_t(p.width) * size_t(p.height);
    float* accum; uint32_t* count; uint8_t* rgb;
    ensure_frame(n_pixels, &accum, &count, &rgb);
    RenderParams q = p;
    q.accumulate = 0;
    if (p.adaptive == 2) {
        render_adaptive_squares(cam, q, rgb, accum, count, stats);
    } else {
        render_device(cam, q, accum, count, stats, 0);
        resolve_device(accum, count, int(n_pixels), p.adaptive ? 0 : p.samples_per_pixel, rgb);
    }
    if (rgb_out) CUDA_CHECK(cudaMemcpy(rgb_out, rgb, n_pixels * 3, cudaMemcpyDeviceToHost));
    if (accum_out) CUDA_CHECK(cudaMemcpy(accum_out, accum, n_pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    if (count_out) CUDA_CHECK(cudaMemcpy(count_out, count, n_pixels * sizeof(uint32_t), cudaMemcpyDeviceToHost));
}

// ---- several GPUs in one process: the render ----------------------------------------------------------------------------
void Renderer::render_host_multi(Renderer* const* peers, int n_peers, int partition, const CameraParams& cam, const RenderParams& p,
                                 uint8_t* rgb_out, float* accum_out, uint32_t* count_out, RenderStats* stats, RenderStats* per_device,
                                 double* reduce_ms)
{
    if (n_peers < 0 || n_peers > ARTB_MAX_GROUP - 1) throw std::runtime_error("render_multi: at most 8 devices per group");
    if (p.adaptive == 2) {
        // the reference's own adaptive strategy: the levels are serial, the pixels of a level are split over the devices
        const size_t px = size_t(p.width) * size_t(p.height);
        float* a; uint32_t* c; uint8_t* rgb8;
        ensure_frame(px, &a, &c, &rgb8);
        RenderParams q = p;
        q.accumulate = 0;
        RenderStats S;
        render_adaptive_squares(cam, q, rgb8, a, c, &S, peers, n_peers);
        if (rgb_out) CUDA_CHECK(cudaMemcpy(rgb_out, rgb8, px * 3, cudaMemcpyDeviceToHost));
        if (accum_out) CUDA_CHECK(cudaMemcpy(accum_out, a, px * 3 * sizeof(float), cudaMemcpyDeviceToHost));
        if (count_out) CUDA_CHECK(cudaMemcpy(count_out, c, px * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (stats) *stats = S;
        if (per_device) per_device[0] = S;
        if (reduce_ms) *reduce_ms = 0;
        return;
    }
    const int n = n_peers + 1;
    const size_t n_pixels = size_t(p.width) * size_t(p.height);
    std::vector<Renderer*> group{ this };
    for (int k = 0; k < n_peers; ++k) group.push_back(peers[k]);
    const size_t un = size_t(n);
    std::vector<float*> accum(un, nullptr);
    std::vector<uint32_t*> count(un, nullptr);
    uint8_t* rgb = nullptr;
    for (int k = 0; k < n; ++k) group[size_t(k)]->ensure_frame(n_pixels, &accum[size_t(k)], &count[size_t(k)], k == 0 ? &rgb : nullptr);

    // every device renders its share from its own host thread (a context's calls are blocking)
    std::vector<RenderStats> st(un);
    std::vector<std::string> errors(un);
    ThreadGroup workers;
    for (int k = 0; k < n; ++k) {
        workers.threads.emplace_back([&, k]() {
            try {
                RenderParams q = p;
                q.accumulate = 0;
                if (partition == 1) {                  // tiles: the whole sample range of this device's interleaved tiles
                    q.tile_parts = n; q.tile_part = k;
                } else {                               // samples: contiguous disjoint index ranges, shares differ by at most one
                    const int base = p.samples_per_pixel / n, rest = p.samples_per_pixel % n;
                    q.samples_per_pixel = base + (k < rest ? 1 : 0);
                    q.sample_begin = p.sample_begin + k * base + std::min(k, rest);
                    q.tile_parts = 1; q.tile_part = 0;
                    if (p.adaptive) q.stream = p.stream + uint32_t(k);
                }
                if (q.samples_per_pixel < 1) {         // more devices than samples: this one has nothing to do
                    CUDA_CHECK(cudaSetDevice(group[size_t(k)]->device()));
                    CUDA_CHECK(cudaMemset(accum[size_t(k)], 0, n_pixels * 3 * sizeof(float)));
                    CUDA_CHECK(cudaMemset(count[size_t(k)], 0, n_pixels * sizeof(uint32_t)));
                    return;
                }
                group[size_t(k)]->render_device(cam, q, accum[size_t(k)], count[size_t(k)], &st[size_t(k)], 0);
            } catch (const std::exception& e) {
                errors[size_t(k)] = e.what();
            } catch (...) {
                errors[size_t(k)] = "unknown error";
            }
        });
    }
    workers.join();
    for (int k = 0; k < n; ++k)
        if (!errors[size_t(k)].empty()) throw std::runtime_error("render_multi: device " + std::to_string(group[size_t(k)]->device()) + ": " + errors[size_t(k)]);

    // the exchange: one kernel on this device reads the peers' frames where they lie
    CUDA_CHECK(cudaSetDevice(device_));
    cudaStream_t s0 = static_cast<cudaStream_t>(stream_);
    Impl& I = *impl_;
    PeerFrames pf{};
    std::vector<void*> staged;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float ms = 0.f;
    try {
        pf = reachable_frames(device_, s0, peers, n_peers, accum.data() + 1, count.data() + 1, n_pixels, true, staged);
        CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
        CUDA_CHECK(cudaEventRecord(e0, s0));
        if (n_peers) k_sum_peer_frames<<<I.sm_count * 4, 256, 0, s0>>>(accum[0], count[0], pf, n_pixels * 3, n_pixels);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaEventRecord(e1, s0));
        CUDA_CHECK(cudaStreamSynchronize(s0));
        CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    } catch (...) {
        for (void* q : staged) cudaFree(q);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        throw;
    }
    for (void* q : staged) cudaFree(q);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (reduce_ms) *reduce_ms = ms;

    // tiles: every pixel got the full spp from exactly one device; samples: the shares add up to spp — either way the
    // summed frame holds spp samples per pixel in fixed mode, and `count` in the adaptive mode
    resolve_device(accum[0], count[0], int(n_pixels), p.adaptive ? 0 : p.samples_per_pixel, rgb);
    if (rgb_out) CUDA_CHECK(cudaMemcpy(rgb_out, rgb, n_pixels * 3, cudaMemcpyDeviceToHost));
    if (accum_out) CUDA_CHECK(cudaMemcpy(accum_out, accum[0], n_pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    if (count_out) CUDA_CHECK(cudaMemcpy(count_out, count[0], n_pixels * sizeof(uint32_t), cudaMemcpyDeviceToHost));

    RenderStats S;
    for (int k = 0; k < n; ++k) {
        const RenderStats& r = st[size_t(k)];
        S.elapsed_ms = std::max(S.elapsed_ms, r.elapsed_ms);          // the devices run side by side
        S.samples += r.samples; S.rays += r.rays; S.waves += r.waves; S.launches += r.launches;
        S.overflow_rays += r.overflow_rays; S.pixels_active += r.pixels_active; S.pixels_converged += r.pixels_converged;
        if (per_device) per_device[k] = r;
    }
    S.elapsed_ms += ms;
    S.launches += n_peers ? 1 : 0;
    if (stats) *stats = S;
}

} // namespace artb
