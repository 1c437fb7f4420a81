// artb_reference_adapter.hpp — the glue a maintainer of blackccpie/another_raytracer adds to hand the reference's OWN scene
// graph to the B200 backend. Header-only; include it AFTER the reference's headers (it names the reference's classes:
// hittable_list, bvh_node, translate, rotate_y, box, sphere, moving_sphere, xy/xz/yz_rect, triangle, constant_medium,
// the materials and the textures — reference: src/engine/hittable_list.h:9-24, src/primitives/bvh.h:9-26,
// src/engine/hittable.h:31-64, src/primitives/*.h, src/engine/constant_medium.h:10-35, src/rendering/material.h:10-135,
// src/rendering/texture.h:11-154, src/rendering/perlin.h:8-97).
//
//     scene world = scene_manager().build(alias);                       // the reference builds its scene as always
//     artb_scene* s = artb_ref::make_scene(world.objects, world.background);
//     artb_upload_scene(ctx, s);                                        // or artb_scene_get_flat + artb_upload_flat
//
// What it does: one depth-first walk that writes the graph as the token stream of artb_tokens.h (doubles only; every
// constructor argument and every DERIVED member the reference computed — bvh_node boxes, rotate_y's sin / cos / bbox —
// is copied, nothing is recomputed) plus a table of the binary payloads the stream names by FNV-1a hash (Perlin tables,
// decoded image texels). artb_scene_from_graph() rebuilds the graph inside the library and flattens it.
//
// Access: a few members it reads are private in the reference (the textures' fields, perlin's tables, image_texture's
// pixels, dielectric::ir, sphere::center ...). Either add `friend struct artb_ref::exporter;` to those classes or compile
// this one translation unit with -fno-access-control (what oracle/Makefile does for oracle/ref_glue/ref_tool.cpp, which
// uses this header to export the unmodified reference's scenes for the tests).
#ifndef ARTB_REFERENCE_ADAPTER_HPP
#define ARTB_REFERENCE_ADAPTER_HPP

#include "artb_tokens.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef ARTB_REF_ADAPTER_NO_LIBRARY
#include "artb.h"
#endif

namespace artb_ref {

inline std::uint64_t fnv1a(const void* data, size_t n, std::uint64_t h = 1469598103934665603ULL)
{
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ULL; }
    return h;
}

struct blob { std::uint64_t hash; std::vector<unsigned char> bytes; };

struct exporter
{
    std::vector<double> out;                  // the token stream
    std::vector<blob> blobs;                  // payloads named by hash in the stream (each once)
    // diagnostics
    std::map<std::string,int> counts;
    int max_depth = 0;
    int bvh_nodes = 0, bvh_single = 0;

    void v3(const vec3& v) { out.push_back(v.x()); out.push_back(v.y()); out.push_back(v.z()); }
    void hash(std::uint64_t h) { out.push_back(double(h & 0xffffffffULL)); out.push_back(double(h >> 32)); }
    void keep(std::uint64_t h, const void* a, size_t na, const void* b = nullptr, size_t nb = 0)
    {
        for (const blob& x : blobs) if (x.hash == h) return;
        blob x;
        x.hash = h;
        x.bytes.assign(static_cast<const unsigned char*>(a), static_cast<const unsigned char*>(a) + na);
        if (b) x.bytes.insert(x.bytes.end(), static_cast<const unsigned char*>(b), static_cast<const unsigned char*>(b) + nb);
        blobs.push_back(std::move(x));
    }

    void tex(const std::shared_ptr<texture>& t)
    {
        if (auto s = std::dynamic_pointer_cast<solid_color>(t)) {
            out.push_back(ARTB_X_SOLID); v3(s->color_value);
        } else if (auto c = std::dynamic_pointer_cast<checker_texture>(t)) {
            out.push_back(ARTB_X_CHECKER); tex(c->even); tex(c->odd);
        } else if (auto n = std::dynamic_pointer_cast<noise_texture>(t)) {
            out.push_back(ARTB_X_NOISE); out.push_back(n->scale);
            // perlin.h:56-61: 256 gradient vectors + three permutations; one payload of 256*24 + 3*256*4 = 9216 bytes
            static_assert(sizeof(vec3) == 24 && sizeof(int) == 4, "payload layout");
            std::vector<unsigned char> pl(9216);
            std::memcpy(pl.data(), n->noise.ranvec.data(), 6144);
            std::memcpy(pl.data() + 6144, n->noise.perm_x.data(), 1024);
            std::memcpy(pl.data() + 7168, n->noise.perm_y.data(), 1024);
            std::memcpy(pl.data() + 8192, n->noise.perm_z.data(), 1024);
            const std::uint64_t h = fnv1a(pl.data(), pl.size());
            keep(h, pl.data(), pl.size());
            hash(h);
        } else if (auto i = std::dynamic_pointer_cast<image_texture>(t)) {
            image(i);
        } else if (auto b = std::dynamic_pointer_cast<barycentric_image_texture>(t)) {
            out.push_back(ARTB_X_BARY_IMAGE);
            out.push_back(b->texcoord_a.first); out.push_back(b->texcoord_a.second);
            out.push_back(b->texcoord_b.first); out.push_back(b->texcoord_b.second);
            out.push_back(b->texcoord_c.first); out.push_back(b->texcoord_c.second);
            image(b->tex);
        } else if (auto bc = std::dynamic_pointer_cast<barycentric_texture>(t)) {
            out.push_back(ARTB_X_BARY); v3(bc->color_a); v3(bc->color_b); v3(bc->color_c);
        } else throw std::logic_error("artb_ref: unknown texture type");
    }
    std::map<const image_texture*, std::uint64_t> image_hashes;   // one image is shared by many triangles
    void image(const std::shared_ptr<image_texture>& i)
    {
        out.push_back(ARTB_X_IMAGE);
        out.push_back(i->width); out.push_back(i->height); out.push_back(i->bytes_per_pixel);
        auto it = image_hashes.find(i.get());
        if (it == image_hashes.end()) {
            const size_t n = size_t(i->width) * size_t(i->height) * size_t(i->bytes_per_pixel);
            const std::uint64_t h = i->data ? fnv1a(i->data.get(), n) : 0;       // no data: the texture renders cyan (texture.h:88-90)
            if (i->data) keep(h, i->data.get(), n);
            it = image_hashes.emplace(i.get(), h).first;
        }
        hash(it->second);
    }
    void mat(const std::shared_ptr<material>& m)
    {
        if (auto l = std::dynamic_pointer_cast<lambertian>(m)) {
            out.push_back(ARTB_M_LAMBERTIAN); tex(l->albedo);
        } else if (auto me = std::dynamic_pointer_cast<metal>(m)) {
            out.push_back(ARTB_M_METAL); v3(me->albedo); out.push_back(me->fuzz);
        } else if (auto d = std::dynamic_pointer_cast<dielectric>(m)) {
            out.push_back(ARTB_M_DIELECTRIC); out.push_back(d->ir);
        } else if (auto dl = std::dynamic_pointer_cast<diffuse_light>(m)) {
            out.push_back(ARTB_M_DIFFUSE_LIGHT); tex(dl->emit);
        } else if (auto is = std::dynamic_pointer_cast<isotropic>(m)) {
            out.push_back(ARTB_M_ISOTROPIC); tex(is->albedo);
        } else throw std::logic_error("artb_ref: unknown material type");
    }
    void list(const hittable_list& l, int depth = 0)
    {
        out.push_back(ARTB_T_LIST); out.push_back(double(l.objects.size()));
        for (const auto& o : l.objects) node(o, depth + 1);
    }
    void node(const std::shared_ptr<hittable>& h, int depth)
    {
        max_depth = std::max(max_depth, depth);
        if (auto l = std::dynamic_pointer_cast<hittable_list>(h)) {
            counts["list"]++; list(*l, depth);
        } else if (auto b = std::dynamic_pointer_cast<bvh_node>(h)) {
            counts["bvh_node"]++; bvh_nodes++;
            const bool same = b->left.get() == b->right.get();        // bvh.cpp:15-16: a span of one puts the object on both sides
            if (same) bvh_single++;
            out.push_back(ARTB_T_BVH); v3(b->box.min()); v3(b->box.max()); out.push_back(same ? 1 : 0);
            node(b->left, depth + 1);
            if (!same) node(b->right, depth + 1);
        } else if (auto t = std::dynamic_pointer_cast<translate>(h)) {
            counts["translate"]++;
            out.push_back(ARTB_T_TRANSLATE); v3(t->offset); node(t->ptr, depth + 1);
        } else if (auto r = std::dynamic_pointer_cast<rotate_y>(h)) {
            counts["rotate_y"]++;
            out.push_back(ARTB_T_ROTATE_Y); out.push_back(r->sin_theta); out.push_back(r->cos_theta);
            out.push_back(r->hasbox ? 1 : 0); v3(r->bbox.min()); v3(r->bbox.max());
            node(r->ptr, depth + 1);
        } else if (auto bx = std::dynamic_pointer_cast<box>(h)) {
            counts["box"]++;
            out.push_back(ARTB_T_BOX); v3(bx->box_min); v3(bx->box_max); list(bx->sides, depth + 1);
        } else if (auto s = std::dynamic_pointer_cast<sphere>(h)) {
            counts["sphere"]++;
            out.push_back(ARTB_T_SPHERE); v3(s->center); out.push_back(s->radius); mat(s->mat_ptr);
        } else if (auto ms = std::dynamic_pointer_cast<moving_sphere>(h)) {
            counts["moving_sphere"]++;
            out.push_back(ARTB_T_MSPHERE); v3(ms->center0); v3(ms->center1);
            out.push_back(ms->time0); out.push_back(ms->time1); out.push_back(ms->radius); mat(ms->mat_ptr);
        } else if (auto xy = std::dynamic_pointer_cast<xy_rect>(h)) {
            counts["xy_rect"]++;
            out.push_back(ARTB_T_XY);
            for (double d : { xy->x0, xy->x1, xy->y0, xy->y1, xy->k }) out.push_back(d);
            mat(xy->mp);
        } else if (auto xz = std::dynamic_pointer_cast<xz_rect>(h)) {
            counts["xz_rect"]++;
            out.push_back(ARTB_T_XZ);
            for (double d : { xz->x0, xz->x1, xz->z0, xz->z1, xz->k }) out.push_back(d);
            mat(xz->mp);
        } else if (auto yz = std::dynamic_pointer_cast<yz_rect>(h)) {
            counts["yz_rect"]++;
            out.push_back(ARTB_T_YZ);
            for (double d : { yz->y0, yz->y1, yz->z0, yz->z1, yz->k }) out.push_back(d);
            mat(yz->mp);
        } else if (auto tr = std::dynamic_pointer_cast<triangle>(h)) {
            counts["triangle"]++;
            out.push_back(ARTB_T_TRI); v3(tr->pt1); v3(tr->pt2); v3(tr->pt3); mat(tr->mat_ptr);
        } else if (auto cm = std::dynamic_pointer_cast<constant_medium>(h)) {
            counts["constant_medium"]++;
            out.push_back(ARTB_T_MEDIUM); out.push_back(cm->neg_inv_density); mat(cm->phase_function);
            node(cm->boundary, depth + 1);
        } else throw std::logic_error("artb_ref: unknown hittable type");
    }

    // container file: int64 header {'ARTG', 1, n_tokens, n_blobs, 0, 0} | background(3 doubles) | tokens | per blob: hash,
    // byte count (uint64 each), bytes padded to a multiple of 8
    bool write_file(const char* path, const color& background) const
    {
        std::FILE* f = std::fopen(path, "wb");
        if (!f) return false;
        const std::int64_t header[6] = { 0x47545241, 1, std::int64_t(out.size()), std::int64_t(blobs.size()), 0, 0 };
        const double bg[3] = { background.x(), background.y(), background.z() };
        bool ok = std::fwrite(header, sizeof(header), 1, f) == 1 && std::fwrite(bg, sizeof(bg), 1, f) == 1 &&
                  (out.empty() || std::fwrite(out.data(), sizeof(double), out.size(), f) == out.size());
        for (const blob& b : blobs) {
            const std::uint64_t hd[2] = { b.hash, std::uint64_t(b.bytes.size()) };
            const unsigned char pad[8] = { 0 };
            ok = ok && std::fwrite(hd, sizeof(hd), 1, f) == 1 && std::fwrite(b.bytes.data(), 1, b.bytes.size(), f) == b.bytes.size();
            const size_t rest = (8 - b.bytes.size() % 8) % 8;
            if (rest) ok = ok && std::fwrite(pad, 1, rest, f) == rest;
        }
        return std::fclose(f) == 0 && ok;
    }
};

#ifndef ARTB_REF_ADAPTER_NO_LIBRARY
// the reference's graph -> a library scene (needs libartb.so). Returns nullptr on failure (artb_last_error(nullptr) says why).
inline artb_scene* make_scene(const hittable_list& world, const color& background, const artb_scene_options* opt = nullptr)
{
    exporter e;
    e.list(world);
    std::vector<artb_graph_blob> table;
    for (const blob& b : e.blobs) table.push_back({ b.hash, b.bytes.data(), std::uint64_t(b.bytes.size()) });
    const double bg[3] = { background.x(), background.y(), background.z() };
    artb_scene* s = nullptr;
    if (artb_scene_from_graph(e.out.data(), std::int64_t(e.out.size()), table.data(), std::int32_t(table.size()), bg, opt, &s) != ARTB_OK)
        return nullptr;
    return s;
}
#endif

} // namespace artb_ref

#endif
