// C ABI of the backend (include/artb.h): exception firewall + argument checking around the C++ host
// mirror (csrc/host) and the CUDA renderer (csrc/device).
#include "artb.h"

#include "../device/renderer.h"
#include "../host/flatten.h"
#include "../host/graph_dump.h"
#include "../host/imageio.h"
#include "../host/scene_manager.h"

#include <cmath>
#include <climits>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime_api.h>

using namespace artb;

struct artb_scene {
    scene host;
    FlatScene flat;
    double rng_next = 0;
    uint64_t rng_draws = 0;
};

struct artb_ctx {
    std::unique_ptr<Renderer> renderer;                       // devices[0]: every call but artb_upload_* / artb_render uses only this one
    std::vector<std::unique_ptr<Renderer>> peers;             // devices[1..] of a multi-device context
    int partition = ARTB_PARTITION_SAMPLES;
    std::string last_error;
    uint64_t last_upload_bytes = 0;
};

static_assert(sizeof(artb_hit) == sizeof(HitRecordOut), "artb_hit must match the device record");
static_assert(sizeof(artb_shade_record) == sizeof(ShadeRecordOut), "artb_shade_record must match the device record");
static_assert(sizeof(FlatPrim) == 96 && sizeof(FlatNode8) == 80 && sizeof(FlatXformOp) == 32 && sizeof(FlatXform) == 8 &&
              sizeof(FlatMaterial) == 48 && sizeof(FlatTexture) == 96 && sizeof(FlatImage) == 24 && sizeof(FlatPerlin) == 9216,
              "flat record sizes are part of the ABI");

namespace {

thread_local std::string g_last_error;
std::mutex g_build_mutex;   // the host RNG is process-global, like the reference's

int fail(artb_ctx* ctx, int code, const std::string& msg)
{
    g_last_error = msg;
    if (ctx) ctx->last_error = msg;
    return code;
}

template <class F>
int guarded(artb_ctx* ctx, F&& f)
{
    try {
        return f();
    } catch (const std::bad_alloc&) {
        return fail(ctx, ARTB_E_CUDA, "out of host memory");
    } catch (const std::logic_error& e) {
        return fail(ctx, ARTB_E_SCENE, e.what());
    } catch (const std::exception& e) {
        return fail(ctx, ARTB_E_CUDA, e.what());
    } catch (...) {
        return fail(ctx, ARTB_E_CUDA, "unknown error");
    }
}

CameraParams make_camera(const artb_camera& c)
{
    const camera cam(point3(c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]), point3(c.lookat[0], c.lookat[1], c.lookat[2]),
                     vec3(c.vup[0], c.vup[1], c.vup[2]), c.vfov, c.aspect_ratio, c.aperture, c.focus_dist, c.time0, c.time1);
    CameraParams p;
    for (int k = 0; k < 3; ++k) {
        p.origin[k] = cam.origin[k]; p.lower_left_corner[k] = cam.lower_left_corner[k];
        p.horizontal[k] = cam.horizontal[k]; p.vertical[k] = cam.vertical[k];
        p.u[k] = cam.u[k]; p.v[k] = cam.v[k]; p.w[k] = cam.w[k];
    }
    p.lens_radius = cam.lens_radius; p.time0 = cam.time0; p.time1 = cam.time1;
    return p;
}

int check_render_args(artb_ctx* ctx, const artb_camera* cam, const artb_render_params* p)
{
    if (!ctx || !cam || !p) return fail(ctx, ARTB_E_INVALID, "null argument");
    if (!ctx->renderer->has_scene()) {
        // engine::run on an empty world: "Invalid input scene!" and -1 (engine.h:32-36)
        return fail(ctx, ARTB_E_EMPTY_SCENE, "Invalid input scene!");
    }
    if (p->width < 2 || p->height < 2) return fail(ctx, ARTB_E_INVALID, "image must be at least 2x2");
    // pixel and component indices are 32-bit on the device (PathAux::pixel, the resolve kernels)
    if (uint64_t(p->width) * uint64_t(p->height) * 3u > uint64_t(INT_MAX)) return fail(ctx, ARTB_E_INVALID, "image too large: width*height*3 must fit in 31 bits");
    if (p->samples_per_pixel < 1) return fail(ctx, ARTB_E_INVALID, "samples_per_pixel must be >= 1");
    if (p->max_depth < 0) return fail(ctx, ARTB_E_INVALID, "max_depth must be >= 0");
    if (p->tile_parts > 1 && (p->tile_part < 0 || p->tile_part >= p->tile_parts)) return fail(ctx, ARTB_E_INVALID, "tile_part out of range");
    return ARTB_OK;
}

RenderParams to_params(const artb_render_params& p)
{
    RenderParams r;
    r.width = p.width; r.height = p.height;
    r.samples_per_pixel = p.samples_per_pixel;
    r.max_depth = p.max_depth > 0 ? p.max_depth : 50;
    r.seed = p.seed; r.stream = p.stream;
    r.sample_begin = p.sample_begin;
    r.tile_size = p.tile_size > 0 ? p.tile_size : 32;
    r.tile_parts = p.tile_parts > 1 ? p.tile_parts : 1;
    r.tile_part = p.tile_part;
    r.adaptive = p.mode == ARTB_MODE_VARIANCE_MASK ? 1 : (p.mode == ARTB_MODE_ADAPTIVE_SQUARES ? 2 : 0);
    r.adaptive_min_spp = p.adaptive_min_spp > 0 ? p.adaptive_min_spp : 16;
    r.adaptive_batch_spp = p.adaptive_batch_spp > 0 ? p.adaptive_batch_spp : 8;
    r.adaptive_threshold = p.adaptive_threshold > 0 ? p.adaptive_threshold : 0.02;
    r.accumulate = p.accumulate;
    return r;
}

void to_stats(const RenderStats& s, artb_stats* out)
{
    if (!out) return;
    std::memset(out, 0, sizeof(*out));
    out->elapsed_ms = s.elapsed_ms;
    out->samples = s.samples; out->rays = s.rays; out->waves = s.waves; out->launches = s.launches;
    out->extend_ms = s.extend_ms; out->extend_launches = s.extend_launches;
    out->pixels_active = s.pixels_active; out->pixels_converged = s.pixels_converged;
    out->traverse_ms = s.traverse_ms; out->pairs_ms = s.pairs_ms; out->finish_ms = s.finish_ms; out->overflow_rays = s.overflow_rays;
}

} // namespace

extern "C" {

int artb_abi_version(void) { return ARTB_ABI_VERSION; }

int artb_create(const artb_config* cfg, artb_ctx** out)
{
    if (!out) return fail(nullptr, ARTB_E_INVALID, "null argument");
    *out = nullptr;
    try {
        auto ctx = std::make_unique<artb_ctx>();
        const int n = cfg ? cfg->n_devices : 0;
        if (n < 0 || n > ARTB_MAX_DEVICES) return fail(nullptr, ARTB_E_INVALID, "n_devices out of range");
        if (cfg && cfg->partition != ARTB_PARTITION_SAMPLES && cfg->partition != ARTB_PARTITION_TILES)
            return fail(nullptr, ARTB_E_INVALID, "unknown partition");
        ctx->renderer = std::make_unique<Renderer>(cfg ? (n >= 1 ? cfg->devices[0] : cfg->device) : 0, cfg ? cfg->pool_paths : 0);
        for (int k = 1; k < n; ++k) ctx->peers.push_back(std::make_unique<Renderer>(cfg->devices[k], cfg->pool_paths));
        if (cfg) ctx->partition = cfg->partition;
        *out = ctx.release();
        return ARTB_OK;
    } catch (const std::exception& e) {
        return fail(nullptr, ARTB_E_NO_DEVICE, e.what());
    } catch (...) {
        return fail(nullptr, ARTB_E_NO_DEVICE, "unknown error");
    }
}

void artb_destroy(artb_ctx* ctx) { delete ctx; }

int artb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int artb_context_devices(const artb_ctx* ctx) { return ctx ? 1 + int(ctx->peers.size()) : 0; }

const char* artb_last_error(const artb_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_last_error.c_str(); }

int artb_scene_build(int32_t alias, const artb_scene_options* opt, artb_scene** out)
{
    if (!out) return fail(nullptr, ARTB_E_INVALID, "null argument");
    *out = nullptr;
    return guarded(nullptr, [&]() -> int {
        std::lock_guard<std::mutex> lock(g_build_mutex);
        if (opt && opt->assets_root) set_assets_root(opt->assets_root);
        reset_global_rng();
        scene_manager mgr;
        if (opt) {
            mgr.mesh_as_soup = opt->mesh_as_soup != 0;
            if (opt->heightfield_side > 0) mgr.heightfield_side = opt->heightfield_side;
        }
        auto s = std::make_unique<artb_scene>();
        s->host = mgr.build(static_cast<scene_alias>(alias));
        s->rng_draws = global_rng().draws();
        s->rng_next = random_double();
        FlattenOptions fo;
        if (opt && opt->max_leaf_prims > 0) fo.max_leaf_prims = opt->max_leaf_prims;
        if (opt && opt->device_bvh) fo.device_bvh = true;
        s->flat = flatten(s->host.objects, s->host.background, fo);
        *out = s.release();
        return ARTB_OK;
    });
}

int artb_scene_from_graph(const double* tokens, int64_t n_tokens, const artb_graph_blob* blobs, int32_t n_blobs,
                          const double background[3], const artb_scene_options* opt, artb_scene** out)
{
    if (!out) return fail(nullptr, ARTB_E_INVALID, "null argument");
    *out = nullptr;
    if (!tokens || n_tokens <= 0 || !background || n_blobs < 0 || (n_blobs && !blobs)) return fail(nullptr, ARTB_E_INVALID, "null argument");
    static_assert(sizeof(artb_graph_blob) == sizeof(GraphBlob), "blob tables are passed through as they are");
    return guarded(nullptr, [&]() -> int {
        auto s = std::make_unique<artb_scene>();
        s->host.objects = build_graph(tokens, size_t(n_tokens), reinterpret_cast<const GraphBlob*>(blobs), size_t(n_blobs));
        s->host.background = color(background[0], background[1], background[2]);
        s->host.lookfrom = point3(0, 0, 0); s->host.lookat = point3(0, 0, 0); s->host.vfov = 0; s->host.aperture = 0;
        FlattenOptions fo;
        if (opt && opt->max_leaf_prims > 0) fo.max_leaf_prims = opt->max_leaf_prims;
        if (opt && opt->device_bvh) fo.device_bvh = true;
        s->flat = flatten(s->host.objects, s->host.background, fo);
        *out = s.release();
        return ARTB_OK;
    });
}

void artb_scene_free(artb_scene* s) { delete s; }

int artb_scene_get_info(const artb_scene* s, artb_scene_info* out)
{
    if (!s || !out) return fail(nullptr, ARTB_E_INVALID, "null argument");
    std::memset(out, 0, sizeof(*out));
    for (int k = 0; k < 3; ++k) {
        out->lookfrom[k] = s->host.lookfrom[k]; out->lookat[k] = s->host.lookat[k]; out->background[k] = s->host.background[k];
        out->world_lo[k] = s->flat.world_lo[k]; out->world_hi[k] = s->flat.world_hi[k];
    }
    out->vfov = s->host.vfov; out->aperture = s->host.aperture;
    out->rng_next = s->rng_next; out->rng_draws = s->rng_draws;
    const FlatScene& f = s->flat;
    out->n_prims = uint32_t(f.prims.size()); out->n_world_prims = f.n_world_prims; out->n_nodes = uint32_t(f.nodes.size());
    out->n_media = f.n_media; out->n_materials = uint32_t(f.materials.size()); out->n_textures = uint32_t(f.textures.size());
    out->n_images = uint32_t(f.images.size()); out->n_xforms = uint32_t(f.xforms.size());
    out->n_leaves_total = f.n_leaves_total; out->n_culled = f.n_culled; out->bvh_depth = uint32_t(f.bvh_depth);
    out->top_level_objects = uint32_t(s->host.objects.objects.size());
    out->flat_bytes = f.prims.size() * sizeof(FlatPrim) + f.nodes.size() * sizeof(FlatNode8) + f.xform_ops.size() * sizeof(FlatXformOp) +
                      f.xforms.size() * sizeof(FlatXform) + f.materials.size() * sizeof(FlatMaterial) +
                      f.textures.size() * sizeof(FlatTexture) + f.images.size() * sizeof(FlatImage) + f.texels.size() +
                      f.perlins.size() * sizeof(FlatPerlin);
    return ARTB_OK;
}

int64_t artb_scene_dump_graph(const artb_scene* s, double* out, int64_t capacity)
{
    if (!s) return fail(nullptr, ARTB_E_INVALID, "null argument");
    try {
        const std::vector<double> d = dump_graph(s->host.objects);
        if (out && capacity > 0)
            std::memcpy(out, d.data(), size_t(std::min<int64_t>(capacity, int64_t(d.size()))) * sizeof(double));
        return int64_t(d.size());
    } catch (const std::exception& e) {
        return fail(nullptr, ARTB_E_SCENE, e.what());
    }
}

int artb_scene_get_flat(const artb_scene* s, artb_flat_scene* out)
{
    if (!s || !out) return fail(nullptr, ARTB_E_INVALID, "null argument");
    const FlatScene& f = s->flat;
    std::memset(out, 0, sizeof(*out));
    out->prims = f.prims.data(); out->n_prims = uint32_t(f.prims.size()); out->n_world_prims = f.n_world_prims;
    out->nodes = f.nodes.data(); out->n_nodes = uint32_t(f.nodes.size());
    out->reserved0 = 0; out->n_global_prims = f.n_global_prims; out->bvh_depth = f.bvh_depth;
    out->xform_ops = f.xform_ops.data(); out->n_xform_ops = uint32_t(f.xform_ops.size());
    out->xforms = f.xforms.data(); out->n_xforms = uint32_t(f.xforms.size());
    out->materials = f.materials.data(); out->n_materials = uint32_t(f.materials.size());
    out->textures = f.textures.data(); out->n_textures = uint32_t(f.textures.size());
    out->images = f.images.data(); out->n_images = uint32_t(f.images.size());
    out->texels = f.texels.data(); out->n_texel_bytes = f.texels.size();
    out->perlins = f.perlins.data(); out->n_perlins = uint32_t(f.perlins.size());
    out->n_media = f.n_media;
    for (int k = 0; k < 3; ++k) out->background[k] = f.background[k];
    out->ref_nodes = f.ref_nodes.data(); out->n_ref_nodes = uint32_t(f.ref_nodes.size());
    return ARTB_OK;
}

int artb_upload_scene(artb_ctx* ctx, const artb_scene* s)
{
    if (!ctx || !s) return fail(ctx, ARTB_E_INVALID, "null argument");
    if (s->host.objects.objects.empty()) return fail(ctx, ARTB_E_EMPTY_SCENE, "Invalid input scene!");
    return guarded(ctx, [&]() -> int {
        ctx->renderer->upload(s->flat);
        for (auto& r : ctx->peers) r->upload(s->flat);
        ctx->last_upload_bytes = ctx->renderer->scene_bytes() * (1 + ctx->peers.size());
        return ARTB_OK;
    });
}

int artb_upload_flat(artb_ctx* ctx, const artb_flat_scene* f)
{
    if (!ctx || !f) return fail(ctx, ARTB_E_INVALID, "null argument");
    if (f->n_world_prims == 0) return fail(ctx, ARTB_E_EMPTY_SCENE, "Invalid input scene!");
    if (f->n_world_prims > f->n_prims || f->n_global_prims > f->n_world_prims)
        return fail(ctx, ARTB_E_INVALID, "inconsistent flat scene");
    return guarded(ctx, [&]() -> int {
        FlatScene fs;
        auto copy = [](auto& dst, const void* src, size_t n) {
            using T = typename std::remove_reference<decltype(dst)>::type::value_type;
            if (n && !src) throw std::logic_error("flat scene: null array with non-zero count");
            dst.assign(static_cast<const T*>(src), static_cast<const T*>(src) + n);
        };
        copy(fs.prims, f->prims, f->n_prims);
        copy(fs.nodes, f->nodes, f->n_nodes);
        copy(fs.xform_ops, f->xform_ops, f->n_xform_ops);
        copy(fs.xforms, f->xforms, f->n_xforms);
        copy(fs.materials, f->materials, f->n_materials);
        copy(fs.textures, f->textures, f->n_textures);
        copy(fs.images, f->images, f->n_images);
        copy(fs.texels, f->texels, f->n_texel_bytes);
        copy(fs.perlins, f->perlins, f->n_perlins);
        copy(fs.ref_nodes, f->ref_nodes, f->n_ref_nodes);
        fs.n_world_prims = f->n_world_prims; fs.n_media = f->n_media;
        fs.n_global_prims = f->n_global_prims; fs.bvh_depth = f->bvh_depth;
        // every constant_medium must sit at the END of the global range (the device tests media after the traversal)
        for (uint32_t i = 0; i < fs.n_world_prims; ++i) {
            const bool is_medium = fs.prims[i].type == PRIM_MEDIUM;
            if (is_medium && i >= fs.n_global_prims) throw std::logic_error("flat scene: a constant_medium outside the global range");
            if (is_medium) fs.n_global_media++;
            else if (fs.n_global_media && i < fs.n_global_prims) throw std::logic_error("flat scene: global media must come last");
        }
        for (int k = 0; k < 3; ++k) fs.background[k] = f->background[k];
        // bounds-check indices so that a malformed scene cannot make the kernels read out of range
        for (const FlatPrim& p : fs.prims) {
            if (p.material >= fs.materials.size() || p.xform >= std::max<size_t>(fs.xforms.size(), 1) || p.type > PRIM_MEDIUM)
                throw std::logic_error("flat scene: primitive refers to a missing material / transform");
            if (p.type == PRIM_MEDIUM) {
                // first boundary primitive, boundary count, medium slot: stored as doubles, used as integers on the device — a
                // NaN, a negative or a fractional value would turn into an arbitrary index there
                auto whole = [](double v) { return std::isfinite(v) && v >= 0 && v <= 4294967295.0 && v == std::floor(v); };
                if (!whole(p.g[1]) || !whole(p.g[2]) || !whole(p.g[3]))
                    throw std::logic_error("flat scene: medium boundary range / slot must be non-negative integers");
                if (p.g[1] < fs.n_world_prims || p.g[2] < 1 || p.g[1] + p.g[2] > double(fs.prims.size()))
                    throw std::logic_error("flat scene: medium boundary range out of bounds");
                if (p.g[3] >= double(fs.n_media)) throw std::logic_error("flat scene: medium slot >= n_media");
                for (uint32_t b = uint32_t(p.g[1]); b < uint32_t(p.g[1] + p.g[2]); ++b)
                    if (fs.prims[b].type == PRIM_MEDIUM) throw std::logic_error("flat scene: a medium inside a medium boundary");
            }
        }
        for (size_t i = 0; i < fs.ref_nodes.size(); ++i) {
            const FlatRefNode& rn = fs.ref_nodes[i];
            if (rn.parent < -1 || rn.parent >= int32_t(i) || rn.xform >= fs.xforms.size() ||
                rn.depth != (rn.parent < 0 ? 0 : fs.ref_nodes[size_t(rn.parent)].depth + 1))
                throw std::logic_error("flat scene: bad reference bvh node");
        }
        for (uint32_t i = 0; i < fs.n_world_prims; ++i)
            if (fs.prims[i].type != PRIM_MEDIUM && !(fs.prims[i].g[9] >= 0 && fs.prims[i].g[9] <= double(fs.ref_nodes.size())))
                throw std::logic_error("flat scene: primitive refers to a missing reference bvh node (g[9])");
        // BVH primitives without nodes: allowed — Renderer::upload then builds the tree on the device (bvh_build.cuh)
        if (fs.bvh_depth < 0 || fs.bvh_depth > ARTB_MAX_BVH8_DEPTH) throw std::logic_error("flat scene: bvh_depth out of range");
        if (fs.n_media > 255) throw std::logic_error("flat scene: more than 255 media");      // MediumXi keys (bounce << 8) | slot
        for (size_t ni = 0; ni < fs.nodes.size(); ++ni) {
            const FlatNode8& n = fs.nodes[ni];
            uint32_t inner = 0;
            for (int sl = 0; sl < 8; ++sl) {
                const uint32_t m = n.meta[sl];
                if (m == 0) { if (n.imask & (1u << sl)) throw std::logic_error("flat scene: imask marks an empty slot"); continue; }
                if ((m & 0x18u) == 0x18u) {          // 24 + slot: internal child
                    if (m != ((1u << 5) | (24u + uint32_t(sl))) || !(n.imask & (1u << sl))) throw std::logic_error("flat scene: bad internal child");
                    ++inner;
                } else {
                    const uint32_t bits = m >> 5, off = m & 31u;
                    const uint32_t count = bits == 1 ? 1 : bits == 3 ? 2 : bits == 7 ? 3 : 0;
                    if (!count || (n.imask & (1u << sl)) || off + count > 24 || n.prim_base < fs.n_global_prims ||
                        size_t(n.prim_base) + off + count > fs.n_world_prims)
                        throw std::logic_error("flat scene: BVH leaf out of bounds");
                }
                for (int a = 0; a < 3; ++a)
                    if (n.e[a] < 1 || n.e[a] > 254) throw std::logic_error("flat scene: bad quantisation exponent");
            }
            if (inner && size_t(n.child_base) + inner > fs.nodes.size()) throw std::logic_error("flat scene: BVH child out of bounds");
            // children strictly after their parent: the node array is a tree laid out forwards, no walk can cycle
            if (inner && size_t(n.child_base) <= ni) throw std::logic_error("flat scene: BVH child_base must be greater than the node's own index");
        }
        for (const FlatXform& x : fs.xforms)
            if (size_t(x.first_op) + x.n_ops > fs.xform_ops.size() || x.n_ops > 8) throw std::logic_error("flat scene: bad transform chain");
        for (const FlatMaterial& m : fs.materials)
            if (m.type >= MAT_COUNT || ((m.type == MAT_LAMBERTIAN || m.type == MAT_DIFFUSE_LIGHT || m.type == MAT_ISOTROPIC) && m.texture >= fs.textures.size()))
                throw std::logic_error("flat scene: bad material");
        for (const FlatImage& im : fs.images) {
            if (im.width == 0) continue;                                 // "no data": rendered cyan, texels never read
            if (im.width < 0 || im.height <= 0 || im.bytes_per_pixel < 3 || im.bytes_per_pixel > 4)
                throw std::logic_error("flat scene: bad image dimensions");
            const uint64_t bytes = uint64_t(im.width) * uint64_t(im.height) * uint64_t(im.bytes_per_pixel);
            if (im.offset > fs.texels.size() || bytes > fs.texels.size() - im.offset)
                throw std::logic_error("flat scene: image extends past the texel atlas");
        }
        for (const FlatTexture& t : fs.textures) {
            if (t.type > TEX_BARY) throw std::logic_error("flat scene: bad texture type");
            if (t.type == TEX_CHECKER && (t.a >= fs.textures.size() || t.b >= fs.textures.size()))
                throw std::logic_error("flat scene: checker texture refers to a missing texture");
            if (t.type == TEX_NOISE && t.a >= fs.perlins.size()) throw std::logic_error("flat scene: noise texture refers to a missing Perlin table");
            if ((t.type == TEX_IMAGE || t.type == TEX_BARY_IMAGE) && t.a >= fs.images.size())
                throw std::logic_error("flat scene: image texture refers to a missing image");
        }
        {   // checker textures nest (texture.h:31-50): the nesting must be a DAG of depth <= 16 (texture_value's loop bound)
            std::vector<int> depth(fs.textures.size(), -1);          // -1 unknown, -2 on the current path
            std::function<int(uint32_t)> walk = [&](uint32_t i) -> int {
                if (depth[i] == -2) throw std::logic_error("flat scene: checker textures form a cycle");
                if (depth[i] >= 0) return depth[i];
                int d = 1;
                if (fs.textures[i].type == TEX_CHECKER) {
                    depth[i] = -2;
                    d = 1 + std::max(walk(fs.textures[i].a), walk(fs.textures[i].b));
                }
                if (d > 16) throw std::logic_error("flat scene: checker textures nested deeper than 16");
                return depth[i] = d;
            };
            for (uint32_t i = 0; i < fs.textures.size(); ++i) walk(i);
        }
        for (const FlatPerlin& pn : fs.perlins)
            for (int i = 0; i < 256; ++i)
                if (uint32_t(pn.perm_x[i]) > 255u || uint32_t(pn.perm_y[i]) > 255u || uint32_t(pn.perm_z[i]) > 255u)
                    throw std::logic_error("flat scene: Perlin permutation entry out of range");
        ctx->renderer->upload(fs);
        for (auto& r : ctx->peers) r->upload(fs);
        ctx->last_upload_bytes = ctx->renderer->scene_bytes() * (1 + ctx->peers.size());
        return ARTB_OK;
    });
}

int artb_trace_batch(artb_ctx* ctx, const double* rays, uint64_t n, double t_min, double t_max,
                     const double* medium_xi, artb_hit* hits_out)
{
    if (!ctx || (n && (!rays || !hits_out))) return fail(ctx, ARTB_E_INVALID, "null argument");
    if (!ctx->renderer->has_scene()) return fail(ctx, ARTB_E_EMPTY_SCENE, "Invalid input scene!");
    return guarded(ctx, [&]() -> int {
        ctx->renderer->trace_batch(rays, n, t_min, t_max, medium_xi, reinterpret_cast<HitRecordOut*>(hits_out));
        return ARTB_OK;
    });
}

int artb_trace_batch_device(artb_ctx* ctx, const double* rays_dev, uint64_t n, double t_min, double t_max,
                            const double* medium_xi_dev, artb_hit* hits_dev, float* kernel_ms)
{
    if (!ctx || (n && (!rays_dev || !hits_dev))) return fail(ctx, ARTB_E_INVALID, "null argument");
    if (!ctx->renderer->has_scene()) return fail(ctx, ARTB_E_EMPTY_SCENE, "Invalid input scene!");
    return guarded(ctx, [&]() -> int {
        const float ms = ctx->renderer->trace_batch_device(rays_dev, n, t_min, t_max, medium_xi_dev,
                                                           reinterpret_cast<HitRecordOut*>(hits_dev));
        if (kernel_ms) *kernel_ms = ms;
        return ARTB_OK;
    });
}

int artb_shade_batch(artb_ctx* ctx, const double* rays, const int32_t* prim_id, const double* t, const double* xi,
                     uint64_t n, artb_shade_record* out)
{
    if (!ctx || (n && (!rays || !prim_id || !t || !xi || !out))) return fail(ctx, ARTB_E_INVALID, "null argument");
    if (!ctx->renderer->has_scene()) return fail(ctx, ARTB_E_EMPTY_SCENE, "Invalid input scene!");
    return guarded(ctx, [&]() -> int {
        ctx->renderer->shade_batch(rays, prim_id, t, xi, n, reinterpret_cast<ShadeRecordOut*>(out));
        return ARTB_OK;
    });
}

int artb_camera_rays(artb_ctx* ctx, const artb_camera* cam, const double* st, const double* xi, uint64_t n, double* rays_out)
{
    if (!ctx || !cam || (n && (!st || !xi || !rays_out))) return fail(ctx, ARTB_E_INVALID, "null argument");
    return guarded(ctx, [&]() -> int {
        ctx->renderer->camera_rays(make_camera(*cam), st, xi, n, rays_out);
        return ARTB_OK;
    });
}

static void to_bvh_stats(const BvhBuildStats& b, artb_bvh_stats* out)
{
    if (!out) return;
    out->build_ms = b.total_ms; out->n_prims = b.n_prims; out->n_nodes = b.n_nodes; out->depth = b.depth; out->reserved = 0;
}

int artb_build_bvh_device(artb_ctx* ctx, int32_t max_leaf_prims, artb_bvh_stats* stats)
{
    if (!ctx) return fail(ctx, ARTB_E_INVALID, "null argument");
    if (!ctx->renderer->has_scene()) return fail(ctx, ARTB_E_EMPTY_SCENE, "Invalid input scene!");
    return guarded(ctx, [&]() -> int {
        BvhBuildStats b;
        ctx->renderer->build_bvh_device(max_leaf_prims > 0 ? max_leaf_prims : 1, &b);
        to_bvh_stats(b, stats);
        return ARTB_OK;
    });
}

int artb_last_bvh_build(const artb_ctx* ctx, artb_bvh_stats* stats)
{
    if (!ctx || !stats) return ARTB_E_INVALID;
    to_bvh_stats(ctx->renderer->last_build(), stats);
    return ARTB_OK;
}

uint64_t artb_trace_overflow_rays(const artb_ctx* ctx)
{
    return ctx && ctx->renderer ? ctx->renderer->last_trace_overflow() : 0;
}

int artb_render(artb_ctx* ctx, const artb_camera* cam, const artb_render_params* p, uint8_t* rgb8_out,
                float* accum_out, uint32_t* count_out, artb_stats* stats)
{
    const int rc = check_render_args(ctx, cam, p);
    if (rc != ARTB_OK) return rc;
    if (!rgb8_out) return fail(ctx, ARTB_E_INVALID, "null output image");
    if (p->mode == ARTB_MODE_ADAPTIVE_SQUARES && (p->width % 12 != 0 || p->height % 12 != 0))      // engine.h:182-183
        return fail(ctx, ARTB_E_ADAPTIVE_SIZE, "for adaptive strategy image size should perfectly fit big square size for now!!");
    return guarded(ctx, [&]() -> int {
        RenderStats s;
        const size_t n_dev = 1 + ctx->peers.size();
        std::vector<RenderStats> per(n_dev);
        double reduce_ms = 0;
        if (n_dev > 1) {
            std::vector<Renderer*> peers;
            for (auto& r : ctx->peers) peers.push_back(r.get());
            ctx->renderer->render_host_multi(peers.data(), int(peers.size()), ctx->partition, make_camera(*cam), to_params(*p), rgb8_out,
                                             accum_out, count_out, &s, per.data(), &reduce_ms);
        } else {
            ctx->renderer->render_host(make_camera(*cam), to_params(*p), rgb8_out, accum_out, count_out, &s);
            per[0] = s;
        }
        to_stats(s, stats);
        if (stats) {
            stats->n_devices = int32_t(n_dev);
            stats->reduce_ms = reduce_ms;
            for (size_t k = 0; k < n_dev && k < ARTB_MAX_DEVICES; ++k) { stats->device_ms[k] = per[k].elapsed_ms; stats->device_samples[k] = per[k].samples; }
            const uint64_t px = uint64_t(p->width) * uint64_t(p->height);
            stats->h2d_bytes = sizeof(CameraParams) + sizeof(RenderParams);
            stats->d2h_bytes = px * 3 + (accum_out ? px * 12 : 0) + (count_out ? px * 4 : 0);
        }
        return ARTB_OK;
    });
}

int artb_render_device(artb_ctx* ctx, const artb_camera* cam, const artb_render_params* p, float* accum_dev,
                       uint32_t* count_dev, artb_stats* stats)
{
    const int rc = check_render_args(ctx, cam, p);
    if (rc != ARTB_OK) return rc;
    if (!accum_dev) return fail(ctx, ARTB_E_INVALID, "null accumulation buffer");
    if (p->mode == ARTB_MODE_VARIANCE_MASK && !count_dev) return fail(ctx, ARTB_E_INVALID, "variance-mask mode needs a count buffer");
    return guarded(ctx, [&]() -> int {
        RenderStats s;
        ctx->renderer->render_device(make_camera(*cam), to_params(*p), accum_dev, count_dev, &s, p->profile > 0 ? p->profile : 0);
        to_stats(s, stats);
        return ARTB_OK;
    });
}

int artb_resolve_device(artb_ctx* ctx, const float* accum_dev, const uint32_t* count_dev, int32_t n_pixels,
                        int32_t fixed_spp, uint8_t* rgb8_dev)
{
    if (!ctx || !accum_dev || !rgb8_dev || n_pixels <= 0 || (fixed_spp <= 0 && !count_dev))
        return fail(ctx, ARTB_E_INVALID, "bad argument");
    return guarded(ctx, [&]() -> int {
        ctx->renderer->resolve_device(accum_dev, count_dev, n_pixels, fixed_spp, rgb8_dev);
        return ARTB_OK;
    });
}

void* artb_stream(artb_ctx* ctx) { return ctx ? ctx->renderer->stream() : nullptr; }

int artb_save_png(const char* path, int32_t width, int32_t height, int32_t channels, const uint8_t* data)
{
    if (!path || !data) return fail(nullptr, ARTB_E_INVALID, "null argument");
    return imageio::save_image(path, width, height, channels, data) ? ARTB_OK : fail(nullptr, ARTB_E_INVALID, "cannot write image");
}

void artb_host_rng_reset(uint32_t seed) { reset_global_rng(seed); }
double artb_host_rng_next(void) { return random_double(); }

} // extern "C"
