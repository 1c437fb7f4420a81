// hostsim.cpp — TEST INFRASTRUCTURE. Compiles the device traversal source (csrc/device/rt_intersect.cuh) with g++ and
// replays it ray by ray on the host, so that the 8-wide tree the flattener emits and the traversal that walks it can
// be checked against the golden ray batches (tests/golden/trace_*.npz) on machines without a GPU, and so that node
// visits / primitive tests per ray can be counted when tuning the builder. Only tests/ load this library; the
// product (libartb.so) has no CPU path and never links it.
#define ARTB_TRAVERSAL_STATS 1
#include "../../another_raytracer_b200/csrc/device/rt_intersect.cuh"
#include "../../include/artb.h"

using namespace artb;

extern "C" {

struct hostsim_hit { double t; int32_t hit; int32_t prim_id; int32_t prim_index; int32_t pad; };

// closest hits of n rays (7 doubles each: origin, direction, time) over a flattened scene held in HOST memory.
// xi: n * n_media explicit medium draws or NULL (Philox keyed by the ray index, as k_trace_batch does).
// stats_out: 3 counters (node visits, primitive tests, medium tests) summed over the batch, or NULL.
// per_ray: n * 3 counters (the same three, per ray) or NULL.
static int g_two_phase = 1;
void hostsim_set_two_phase(int on) { g_two_phase = on; }

int hostsim_trace(const artb_flat_scene* f, const double* rays, uint64_t n, double t_min, double t_max, const double* xi,
                  hostsim_hit* out, unsigned long long* stats_out, uint32_t* per_ray)
{
    SceneView s{};
    s.prims = static_cast<const FlatPrim*>(f->prims);
    s.nodes = static_cast<const FlatNode8*>(f->nodes);
    s.xform_ops = static_cast<const FlatXformOp*>(f->xform_ops);
    s.xforms = static_cast<const FlatXform*>(f->xforms);
    s.materials = static_cast<const FlatMaterial*>(f->materials);
    s.textures = static_cast<const FlatTexture*>(f->textures);
    s.images = static_cast<const FlatImage*>(f->images);
    s.texels = f->texels;
    s.perlins = static_cast<const FlatPerlin*>(f->perlins);
    s.ref_nodes = f->n_ref_nodes ? static_cast<const FlatRefNode*>(f->ref_nodes) : nullptr;
    s.n_prims = f->n_prims; s.n_world_prims = f->n_world_prims; s.n_global_prims = f->n_global_prims;
    s.n_nodes = f->n_nodes; s.n_media = f->n_media; s.n_xforms = f->n_xforms;
    s.n_materials = f->n_materials; s.n_textures = f->n_textures;
    s.n_global_media = 0;
    s.one_bits = 0x3f800000u;
    for (uint32_t i = 0; i < s.n_global_prims; ++i) s.n_global_media += s.prims[i].type == PRIM_MEDIUM ? 1u : 0u;
    for (int k = 0; k < 3; ++k) s.background[k] = f->background[k];
    s.self = &s;                      // the out-of-line helpers of rt_intersect.cuh reach the view through this pointer

    TraversalStats st{};
    for (uint64_t i = 0; i < n; ++i) {
        DRay r;
        r.o = ld3(rays + i * 7); r.d = ld3(rays + i * 7 + 3); r.tm = rays[i * 7 + 6];
        MediumXi mx;
        mx.row = xi ? xi + i * s.n_media : nullptr;
        mx.seed = 0; mx.stream = 0; mx.path_lo = uint32_t(i); mx.path_hi = uint32_t(i >> 32); mx.bounce = 0;
        const TraversalStats before = st;
        const ClosestHit h = g_two_phase ? closest_hit_two_phase(s, s.nodes, r, t_min, t_max, mx, &st)
                                         : closest_hit(s, s.nodes, r, t_min, t_max, mx, &st);
        if (per_ray) {
            per_ray[i * 3 + 0] = uint32_t(st.nodes - before.nodes);
            per_ray[i * 3 + 1] = uint32_t(st.prims - before.prims);
            per_ray[i * 3 + 2] = uint32_t(st.media - before.media);
        }
        out[i].t = h.prim < 0 ? 0.0 : h.t;
        out[i].hit = h.prim < 0 ? 0 : 1;
        out[i].prim_id = h.prim < 0 ? -1 : int32_t(h.dfs);
        out[i].prim_index = h.prim;
        out[i].pad = 0;
    }
    if (stats_out) { stats_out[0] = st.nodes; stats_out[1] = st.prims; stats_out[2] = st.media; }
    return 0;
}

// The float screen k_finish puts in front of the exact rectangle test (rect_certain_miss), next to the exact test itself:
// out[i] bit 0 = rect_t() hits within (t_min, t_max[i]], bit 1 = the screen calls ray i a certain miss. The two bits must
// never be set together. type: ARTB_PRIM_RECT_XY / XZ / YZ; g = a0 a1 b0 b1 k; t_max: one value per ray.
int hostsim_rect_screen(int type, const double* g, const double* rays, uint64_t n, double t_min, const double* t_max, uint8_t* out)
{
    FlatPrim p{};
    p.type = uint32_t(type);
    for (int k = 0; k < 5; ++k) p.g[k] = g[k];
    RectScreen q;
    if (!rect_screen_make(p, q)) return 1;
    for (uint64_t i = 0; i < n; ++i) {
        DRay r;
        r.o = mk3(rays[i * 7 + 0], rays[i * 7 + 1], rays[i * 7 + 2]);
        r.d = mk3(rays[i * 7 + 3], rays[i * 7 + 4], rays[i * 7 + 5]);
        r.tm = rays[i * 7 + 6];
        double t = 0;
        const bool hit = rect_t(p.type, p.g, r, t_min, t_max[i], t);
        const float of[3] = { float(r.o.x), float(r.o.y), float(r.o.z) }, df[3] = { float(r.d.x), float(r.d.y), float(r.d.z) };
        const float inv[3] = { 1.0f / df[0], 1.0f / df[1], 1.0f / df[2] };
        const bool miss = rect_certain_miss(q, of, df, inv, d2f_rd(t_min), d2f_ru(t_max[i]));
        out[i] = uint8_t((hit ? 1 : 0) | (miss ? 2 : 0));
    }
    return 0;
}

}
