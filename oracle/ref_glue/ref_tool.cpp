// TEST INFRASTRUCTURE (oracle) — never linked into or called by the product.
//
// ref_tool: drives the UNMODIFIED reference (blackccpie/another_raytracer, sources compiled where they
// lie under /root/reference) to produce ground truth for the hot path:
//   rng N                         first N values of the reference's global RNG (tracer_utils.h:27-31)
//   info ALIAS                    scene statistics + next RNG value after scene_manager::build
//   graph ALIAS OUT               pre-order dump of the built scene graph (include/artb_tokens.h)
//   export ALIAS OUT              the same stream + the payloads it names (Perlin tables, texels) in the container
//                                 format of include/artb_reference_adapter.hpp: input of artb_scene_from_graph
//   trace ALIAS W H STRIDE SPP MAX OUT   every closest-hit query (engine.h:455) made while path tracing
//                                 a sparse pixel grid: ray, hit record, winning leaf id, media xi
//   render ALIAS W H SPP OUT [BURN]      single-mode render with runtime W/H/spp (restates the three
//                                 loops of engine.h:58-94 on top of the reference's own camera::get_ray,
//                                 hittable::hit, material::scatter, write_color); raw RGB8 out
//   decode IN OUT                 imageio::load_image (stb) -> "w h bpp\n" + raw texels
//   uv                            get_sphere_uv known answers (sphere.h:28-30)
// Built with -fno-access-control so that private members (sphere::center, image_texture::data, ...)
// can be read without editing any reference file. This TU includes scene_manager.cpp directly because
// sphere.h / moving_sphere.h / triangle.h / constant_medium.h define non-inline member functions.
//
// Scene aliases 1..9 are the reference's (scene_manager.h:16-27). 10 = cow.obj and 11 = dino.obj,
// assembled here exactly like scene_manager::_mesh_scene (scene_manager.cpp:236-258) with the camera
// presets the reference keeps as comments (scene_manager.cpp:337-342).

#include "scene_manager.cpp"   // the reference TU (resolved through -I<reference>/src)

#include "camera.h"
#include "color.h"

#include "ref_scenes.h"

#include "artb_tokens.h"
#define ARTB_REF_ADAPTER_NO_LIBRARY            // ref_tool does not link libartb.so: it only writes files
#include "artb_reference_adapter.hpp"

#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <random>
#include <set>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------
// graph dump / export: include/artb_reference_adapter.hpp — the SAME walker a reference maintainer would use to hand the
// reference's graph to the backend; `graph` writes its token stream (pinned by tests/golden/scene_info.json),
// `export` the container file (tokens + Perlin / texel payloads) that artb_scene_from_graph consumes
// ------------------------------------------------------------------------------------------------
using artb_ref::fnv1a;
using dumper = artb_ref::exporter;

// ------------------------------------------------------------------------------------------------
// leaf tagging: each surface leaf and each constant_medium is wrapped by a forwarding
// decorator that records its DFS id on success. Because every successful leaf hit overwrites the
// record, the LAST successful tagged hit of a closest-hit query is the winner.
// Ids are handed out in DFS pre-order; media boundaries are walked (ids reserved, so numbering
// matches the product's flattener) but not wrapped.
// ------------------------------------------------------------------------------------------------
thread_local int g_last_id = -1;

struct tagged final : public hittable
{
    tagged(std::shared_ptr<hittable> i, int id_) : inner(std::move(i)), id(id_) {}
    bool hit(const ray& r, double t_min, double t_max, hit_record& rec) const override
    {
        if (!inner->hit(r, t_min, t_max, rec)) return false;
        g_last_id = id;
        return true;
    }
    bool bounding_box(double t0, double t1, aabb& b) const override { return inner->bounding_box(t0, t1, b); }
    std::shared_ptr<hittable> inner;
    int id;
};

// A shadow copy of the reference's global generator, resynchronised around every constant_medium::hit
// so that the xi that call consumed (constant_medium.h:63) can be recorded.
struct shadow_rng
{
    std::uniform_real_distribution<double> distribution{0.0, 1.0};
    std::mt19937 generator;
    double next() { return distribution(generator); }
    void sync()   // consumes one value of the real stream and fast-forwards the shadow to just after it
    {
        const double x = random_double();
        for (long guard = 0; guard < 2000000000L; ++guard) if (next() == x) return;
        throw std::logic_error("shadow rng lost sync");
    }
};
shadow_rng g_shadow;
std::vector<double>* g_xi_row = nullptr;   // per-query xi slots, one per medium (NaN = no draw)

struct medium_probe final : public hittable
{
    medium_probe(std::shared_ptr<hittable> i, int id_, int slot_) : inner(std::move(i)), id(id_), slot(slot_) {}
    bool hit(const ray& r, double t_min, double t_max, hit_record& rec) const override
    {
        g_shadow.sync();
        const bool ok = inner->hit(r, t_min, t_max, rec);
        const double after = random_double();
        const double s1 = g_shadow.next();
        if (s1 != after) {                       // the medium consumed exactly one value: s1
            const double s2 = g_shadow.next();
            if (s2 != after) throw std::logic_error("medium consumed more than one random value");
            if (g_xi_row) (*g_xi_row)[size_t(slot)] = s1;
        }
        if (ok) g_last_id = id;
        return ok;
    }
    bool bounding_box(double t0, double t1, aabb& b) const override { return inner->bounding_box(t0, t1, b); }
    std::shared_ptr<hittable> inner;
    int id, slot;
};

struct tagger
{
    int next_id = 0;
    int n_media = 0;
    bool probe_media = true;

    void skip(const std::shared_ptr<hittable>& h)   // reserve ids inside a medium boundary
    {
        if (auto l = std::dynamic_pointer_cast<hittable_list>(h)) { for (auto& o : l->objects) skip(o); }
        else if (auto b = std::dynamic_pointer_cast<bvh_node>(h)) {
            skip(b->left); if (b->left.get() != b->right.get()) skip(b->right);
        }
        else if (auto t = std::dynamic_pointer_cast<translate>(h)) skip(t->ptr);
        else if (auto r = std::dynamic_pointer_cast<rotate_y>(h)) skip(r->ptr);
        else if (auto bx = std::dynamic_pointer_cast<box>(h)) { for (auto& o : bx->sides.objects) skip(o); }
        else if (auto cm = std::dynamic_pointer_cast<constant_medium>(h)) { next_id++; skip(cm->boundary); }
        else next_id++;
    }

    void wrap(std::shared_ptr<hittable>& h)
    {
        if (auto l = std::dynamic_pointer_cast<hittable_list>(h)) { for (auto& o : l->objects) wrap(o); }
        else if (auto b = std::dynamic_pointer_cast<bvh_node>(h)) {
            if (b->left.get() == b->right.get()) { wrap(b->left); b->right = b->left; }
            else { wrap(b->left); wrap(b->right); }
        }
        else if (auto t = std::dynamic_pointer_cast<translate>(h)) wrap(t->ptr);
        else if (auto r = std::dynamic_pointer_cast<rotate_y>(h)) wrap(r->ptr);
        else if (auto bx = std::dynamic_pointer_cast<box>(h)) { for (auto& o : bx->sides.objects) wrap(o); }
        else if (auto cm = std::dynamic_pointer_cast<constant_medium>(h)) {
            const int id = next_id++;
            skip(cm->boundary);
            if (probe_media) h = std::make_shared<medium_probe>(h, id, n_media);
            else             h = std::make_shared<tagged>(h, id);
            n_media++;
        }
        else { h = std::make_shared<tagged>(h, next_id++); }
    }
};

int material_code(const std::shared_ptr<material>& m)
{
    if (std::dynamic_pointer_cast<lambertian>(m)) return ARTB_M_LAMBERTIAN;
    if (std::dynamic_pointer_cast<metal>(m)) return ARTB_M_METAL;
    if (std::dynamic_pointer_cast<dielectric>(m)) return ARTB_M_DIELECTRIC;
    if (std::dynamic_pointer_cast<diffuse_light>(m)) return ARTB_M_DIFFUSE_LIGHT;
    if (std::dynamic_pointer_cast<isotropic>(m)) return ARTB_M_ISOTROPIC;
    return -1;
}

// ------------------------------------------------------------------------------------------------
// trace recorder
// ------------------------------------------------------------------------------------------------
struct trace_log
{
    size_t max_rays = 0;
    int n_media = 0;
    std::vector<double> rays;     // 7 per query
    std::vector<double> tmax;     // 1 per query (always +inf here, kept for the format)
    std::vector<std::int32_t> hit, prim, ff, mat, bounce;
    std::vector<double> t, p, n, uv;   // 1,3,3,2
    std::vector<double> xi;            // n_media per query
    bool full() const { return hit.size() >= max_rays; }
};

const double nan_v = std::numeric_limits<double>::quiet_NaN();

// restates engine::_ray_color (engine.h:447-466) with a recorder around the closest-hit query
color ray_color_rec(const ray& r, const color& background, const hittable& world, int depth, trace_log* log)
{
    hit_record rec;
    if (depth <= 0)
        return color(0,0,0);

    std::vector<double> xi_row;
    if (log) { xi_row.assign(size_t(log->n_media), nan_v); g_xi_row = &xi_row; }
    g_last_id = -1;
    // u,v are left untouched by moving_sphere / constant_medium (Q8); give them a defined value
    rec.u = 0; rec.v = 0; rec.front_face = false;
    const bool ok = world.hit(r, 0.001, infinity, rec);
    g_xi_row = nullptr;
    if (log && !log->full()) {
        for (int k = 0; k < 3; ++k) log->rays.push_back(r.orig[k]);
        for (int k = 0; k < 3; ++k) log->rays.push_back(r.dir[k]);
        log->rays.push_back(r.tm);
        log->tmax.push_back(infinity);
        log->hit.push_back(ok ? 1 : 0);
        log->prim.push_back(ok ? g_last_id : -1);
        log->ff.push_back(ok ? (rec.front_face ? 1 : 0) : 0);
        log->mat.push_back(ok ? material_code(rec.mat_ptr) : -1);
        log->bounce.push_back(tracer_constants::max_depth - depth);
        log->t.push_back(ok ? rec.t : nan_v);
        for (int k = 0; k < 3; ++k) log->p.push_back(ok ? rec.p[k] : nan_v);
        for (int k = 0; k < 3; ++k) log->n.push_back(ok ? rec.normal[k] : nan_v);
        log->uv.push_back(ok ? rec.u : nan_v);
        log->uv.push_back(ok ? rec.v : nan_v);
        for (double x : xi_row) log->xi.push_back(x);
    }
    if (!ok)
        return background;

    ray scattered;
    color attenuation;
    color emitted = rec.mat_ptr->emitted(rec.u, rec.v, rec.p);

    if (!rec.mat_ptr->scatter(r, rec, attenuation, scattered))
        return emitted;

    return emitted + attenuation * ray_color_rec(scattered, background, world, depth-1, log);
}

template<typename T>
void write_vec(std::ofstream& f, const std::vector<T>& v)
{
    f.write(reinterpret_cast<const char*>(v.data()), std::streamsize(v.size() * sizeof(T)));
}

int cmd_rng(int n)
{
    for (int i = 0; i < n; ++i) std::printf("%.17g\n", random_double());
    return 0;
}

int cmd_info(int alias)
{
    scene s = build_scene(alias);
    const double next = random_double();
    dumper d;
    d.list(s.objects, 0);
    aabb bb; const bool hasbb = s.objects.bounding_box(0, 1, bb);
    std::printf("{\"alias\": %d, \"rng_next\": %.17g, \"top_level\": %zu, \"max_depth\": %d, "
                "\"bvh_nodes\": %d, \"bvh_single_child\": %d, \"stream_len\": %zu, \"stream_hash\": \"%016" PRIx64 "\"",
                alias, next, s.objects.objects.size(), d.max_depth, d.bvh_nodes, d.bvh_single, d.out.size(),
                fnv1a(d.out.data(), d.out.size() * sizeof(double)));
    if (hasbb)
        std::printf(", \"bbox\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g]", bb.min().x(), bb.min().y(), bb.min().z(),
                    bb.max().x(), bb.max().y(), bb.max().z());
    std::printf(", \"lookfrom\": [%.17g, %.17g, %.17g], \"lookat\": [%.17g, %.17g, %.17g], \"vfov\": %.17g, "
                "\"aperture\": %.17g, \"background\": [%.17g, %.17g, %.17g]",
                s.lookfrom.x(), s.lookfrom.y(), s.lookfrom.z(), s.lookat.x(), s.lookat.y(), s.lookat.z(), s.vfov,
                s.aperture, s.background.x(), s.background.y(), s.background.z());
    std::printf(", \"counts\": {");
    bool first = true;
    for (auto& kv : d.counts) { std::printf("%s\"%s\": %d", first ? "" : ", ", kv.first.c_str(), kv.second); first = false; }
    std::printf("}}\n");
    return 0;
}

int cmd_graph(int alias, const char* out)
{
    scene s = build_scene(alias);
    dumper d;
    d.list(s.objects, 0);
    std::ofstream f(out, std::ios::binary);
    write_vec(f, d.out);
    return f.good() ? 0 : 1;
}

int cmd_trace(int alias, int W, int H, int stride, int spp, size_t max_rays, const char* out)
{
    scene s = build_scene(alias);
    camera cam = make_camera(s, W, H);
    tagger tg;
    std::shared_ptr<hittable> root = std::make_shared<hittable_list>(s.objects);
    tg.wrap(root);
    hittable_list& world = *std::dynamic_pointer_cast<hittable_list>(root);

    trace_log log;
    log.max_rays = max_rays;
    log.n_media = tg.n_media;
    for (int j = stride / 2; j < H && !log.full(); j += stride) {
        for (int i = stride / 2; i < W && !log.full(); i += stride) {
            for (int sidx = 0; sidx < spp; ++sidx) {
                auto u = (i + random_double()) / (W-1);
                auto v = ((H-1-j) + random_double()) / (H-1);
                ray r = cam.get_ray(u, v);
                ray_color_rec(r, s.background, world, tracer_constants::max_depth, &log);
            }
        }
    }
    std::ofstream f(out, std::ios::binary);
    const std::int64_t header[8] = { 0x54545241 /* 'ARTT' */, 1, std::int64_t(log.hit.size()), log.n_media,
                                     tg.next_id, alias, W, H };
    f.write(reinterpret_cast<const char*>(header), sizeof(header));
    write_vec(f, log.rays); write_vec(f, log.tmax);
    write_vec(f, log.hit); write_vec(f, log.prim); write_vec(f, log.ff); write_vec(f, log.mat); write_vec(f, log.bounce);
    write_vec(f, log.t); write_vec(f, log.p); write_vec(f, log.n); write_vec(f, log.uv); write_vec(f, log.xi);
    std::printf("{\"rays\": %zu, \"n_media\": %d, \"n_leaves\": %d}\n", log.hit.size(), log.n_media, tg.next_id);
    return f.good() ? 0 : 1;
}

int cmd_render(int alias, int W, int H, int spp, const char* out, long burn)
{
    scene s = build_scene(alias);
    for (long i = 0; i < burn; ++i) random_double();   // shift the render stream, scene unchanged
    camera cam = make_camera(s, W, H);
    const hittable_list& world = s.objects;
    std::vector<std::uint8_t> img(size_t(W) * size_t(H) * 3);
    std::vector<double> lin(size_t(W) * size_t(H) * 3);
    const auto start = std::chrono::steady_clock::now();
    for (int j = 0; j < H; ++j) {                       // engine.h:78-89
        size_t offset = size_t(3) * size_t(j) * size_t(W);
        for (int i = 0; i < W; ++i) {
            color pixel_color(0, 0, 0);                 // engine.h:58-68
            for (int sidx = 0; sidx < spp; ++sidx) {
                auto u = (i + random_double()) / (W-1);
                auto v = ((H-1-j) + random_double()) / (H-1);
                ray r = cam.get_ray(u, v);
                pixel_color += ray_color_rec(r, s.background, world, tracer_constants::max_depth, nullptr);
            }
            write_color(img.data() + offset, pixel_color, spp);
            for (int k = 0; k < 3; ++k) lin[offset + size_t(k)] = pixel_color[k] / spp;
            offset += 3;
        }
    }
    const auto end = std::chrono::steady_clock::now();
    const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(end - start).count();
    std::ofstream f(out, std::ios::binary);
    write_vec(f, img);
    std::ofstream fl(std::string(out) + ".lin", std::ios::binary);
    write_vec(fl, lin);
    std::printf("{\"alias\": %d, \"w\": %d, \"h\": %d, \"spp\": %d, \"elapsed_ms\": %lld}\n", alias, W, H, spp, (long long)ms);
    return f.good() ? 0 : 1;
}

int cmd_export(int alias, const char* out)
{
    scene s = build_scene(alias);
    dumper d;
    d.list(s.objects);
    if (!d.write_file(out, s.background)) { std::fprintf(stderr, "cannot write %s\n", out); return 1; }
    std::printf("{\"alias\": %d, \"tokens\": %zu, \"blobs\": %zu, \"lookfrom\": [%.17g, %.17g, %.17g], \"lookat\": [%.17g, %.17g, %.17g], "
                "\"vfov\": %.17g, \"aperture\": %.17g}\n", alias, d.out.size(), d.blobs.size(), s.lookfrom.x(), s.lookfrom.y(), s.lookfrom.z(),
                s.lookat.x(), s.lookat.y(), s.lookat.z(), s.vfov, s.aperture);
    return 0;
}

int cmd_decode(const char* in, const char* out)
{
    int w = 0, h = 0, bpp = 0;
    auto data = imageio::load_image(in, w, h, bpp);
    if (!data) { std::fprintf(stderr, "cannot decode %s\n", in); return 1; }
    std::ofstream f(out, std::ios::binary);
    f << w << " " << h << " " << bpp << "\n";
    f.write(reinterpret_cast<const char*>(data.get()), std::streamsize(size_t(w) * size_t(h) * size_t(bpp)));
    std::printf("{\"w\": %d, \"h\": %d, \"bpp\": %d, \"hash\": \"%016" PRIx64 "\"}\n", w, h, bpp,
                fnv1a(data.get(), size_t(w) * size_t(h) * size_t(bpp)));
    return 0;
}

int cmd_uv()
{
    const double dirs[6][3] = { {1,0,0}, {-1,0,0}, {0,1,0}, {0,-1,0}, {0,0,1}, {0,0,-1} };
    for (auto& d : dirs) {
        double u, v;
        sphere::get_sphere_uv(point3(d[0], d[1], d[2]), u, v);
        std::printf("%.17g %.17g %.17g %.17g %.17g\n", d[0], d[1], d[2], u, v);
    }
    return 0;
}

} // namespace

int main(int argc, char** argv) try
{
    if (argc < 2) { std::fprintf(stderr, "usage: ref_tool rng|info|graph|trace|render|decode|uv ...\n"); return 2; }
    const std::string cmd = argv[1];
    if (cmd == "rng" && argc >= 3)     return cmd_rng(std::atoi(argv[2]));
    if (cmd == "info" && argc >= 3)    return cmd_info(std::atoi(argv[2]));
    if (cmd == "graph" && argc >= 4)   return cmd_graph(std::atoi(argv[2]), argv[3]);
    if (cmd == "export" && argc >= 4)  return cmd_export(std::atoi(argv[2]), argv[3]);
    if (cmd == "trace" && argc >= 9)   return cmd_trace(std::atoi(argv[2]), std::atoi(argv[3]), std::atoi(argv[4]),
                                                        std::atoi(argv[5]), std::atoi(argv[6]),
                                                        size_t(std::atoll(argv[7])), argv[8]);
    if (cmd == "render" && argc >= 7)  return cmd_render(std::atoi(argv[2]), std::atoi(argv[3]), std::atoi(argv[4]),
                                                         std::atoi(argv[5]), argv[6], argc >= 8 ? std::atol(argv[7]) : 0);
    if (cmd == "decode" && argc >= 4)  return cmd_decode(argv[2], argv[3]);
    if (cmd == "uv")                   return cmd_uv();
    std::fprintf(stderr, "bad arguments\n");
    return 2;
}
catch (const std::exception& e)
{
    std::fprintf(stderr, "ref_tool: %s\n", e.what());
    return 1;
}
