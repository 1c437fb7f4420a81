#include "graph_dump.h"

#include "artb_tokens.h"

#include <map>
#include <stdexcept>
#include <unordered_map>

namespace artb {

std::uint64_t fnv1a64(const void* data, size_t n, std::uint64_t h)
{
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ULL; }
    return h;
}

namespace {

struct writer {
    std::vector<double> out;

    void put(double d) { out.push_back(d); }
    void put(const vec3& v) { put(v.x()); put(v.y()); put(v.z()); }
    void put_hash(std::uint64_t h) { put(double(h & 0xffffffffULL)); put(double(h >> 32)); }

    std::unordered_map<const image_texture*, std::uint64_t> image_hashes;   // one image is shared by many triangles

    void image(const image_texture& i)
    {
        put(ARTB_X_IMAGE); put(i.width); put(i.height); put(i.bytes_per_pixel);
        auto it = image_hashes.find(&i);
        if (it == image_hashes.end())
            it = image_hashes.emplace(&i, i.data.empty() ? 0 : fnv1a64(i.data.data(), i.data.size())).first;
        put_hash(it->second);
    }
    void tex(const texture& t)
    {
        switch (t.kind()) {
            case texture_kind::solid: put(ARTB_X_SOLID); put(static_cast<const solid_color&>(t).color_value); break;
            case texture_kind::checker: {
                const auto& c = static_cast<const checker_texture&>(t);
                put(ARTB_X_CHECKER); tex(*c.even); tex(*c.odd); break;
            }
            case texture_kind::noise: {
                const auto& n = static_cast<const noise_texture&>(t);
                put(ARTB_X_NOISE); put(n.scale);
                std::uint64_t h = fnv1a64(n.noise.ranvec.data(), sizeof(vec3) * 256);
                h = fnv1a64(n.noise.perm_x.data(), sizeof(int) * 256, h);
                h = fnv1a64(n.noise.perm_y.data(), sizeof(int) * 256, h);
                h = fnv1a64(n.noise.perm_z.data(), sizeof(int) * 256, h);
                put_hash(h); break;
            }
            case texture_kind::image: image(static_cast<const image_texture&>(t)); break;
            case texture_kind::barycentric_image: {
                const auto& b = static_cast<const barycentric_image_texture&>(t);
                put(ARTB_X_BARY_IMAGE);
                put(b.texcoord_a.first); put(b.texcoord_a.second);
                put(b.texcoord_b.first); put(b.texcoord_b.second);
                put(b.texcoord_c.first); put(b.texcoord_c.second);
                image(*b.tex); break;
            }
            case texture_kind::barycentric: {
                const auto& b = static_cast<const barycentric_texture&>(t);
                put(ARTB_X_BARY); put(b.color_a); put(b.color_b); put(b.color_c); break;
            }
        }
    }
    void mat(const material& m)
    {
        switch (m.kind()) {
            case material_kind::lambertian: put(ARTB_M_LAMBERTIAN); tex(*static_cast<const lambertian&>(m).albedo); break;
            case material_kind::metal: {
                const auto& me = static_cast<const metal&>(m);
                put(ARTB_M_METAL); put(me.albedo); put(me.fuzz); break;
            }
            case material_kind::dielectric: put(ARTB_M_DIELECTRIC); put(static_cast<const dielectric&>(m).ir); break;
            case material_kind::diffuse_light: put(ARTB_M_DIFFUSE_LIGHT); tex(*static_cast<const diffuse_light&>(m).emit); break;
            case material_kind::isotropic: put(ARTB_M_ISOTROPIC); tex(*static_cast<const isotropic&>(m).albedo); break;
        }
    }
    void list(const hittable_list& l)
    {
        put(ARTB_T_LIST); put(double(l.objects.size()));
        for (const auto& o : l.objects) node(*o);
    }
    void node(const hittable& h)
    {
        switch (h.kind()) {
            case hittable_kind::list: list(static_cast<const hittable_list&>(h)); break;
            case hittable_kind::bvh: {
                const auto& b = static_cast<const bvh_node&>(h);
                const bool same = b.left.get() == b.right.get();
                put(ARTB_T_BVH); put(b.box.min()); put(b.box.max()); put(same ? 1 : 0);
                node(*b.left);
                if (!same) node(*b.right);
                break;
            }
            case hittable_kind::translate: {
                const auto& t = static_cast<const translate&>(h);
                put(ARTB_T_TRANSLATE); put(t.offset); node(*t.ptr); break;
            }
            case hittable_kind::rotate_y: {
                const auto& r = static_cast<const rotate_y&>(h);
                put(ARTB_T_ROTATE_Y); put(r.sin_theta); put(r.cos_theta); put(r.hasbox ? 1 : 0);
                put(r.bbox.min()); put(r.bbox.max()); node(*r.ptr); break;
            }
            case hittable_kind::box: {
                const auto& b = static_cast<const box&>(h);
                put(ARTB_T_BOX); put(b.box_min); put(b.box_max); list(b.sides); break;
            }
            case hittable_kind::sphere: {
                const auto& s = static_cast<const sphere&>(h);
                put(ARTB_T_SPHERE); put(s.center); put(s.radius); mat(*s.mat_ptr); break;
            }
            case hittable_kind::moving_sphere: {
                const auto& s = static_cast<const moving_sphere&>(h);
                put(ARTB_T_MSPHERE); put(s.center0); put(s.center1); put(s.time0); put(s.time1); put(s.radius);
                mat(*s.mat_ptr); break;
            }
            case hittable_kind::xy_rect: {
                const auto& r = static_cast<const xy_rect&>(h);
                put(ARTB_T_XY); put(r.x0); put(r.x1); put(r.y0); put(r.y1); put(r.k); mat(*r.mp); break;
            }
            case hittable_kind::xz_rect: {
                const auto& r = static_cast<const xz_rect&>(h);
                put(ARTB_T_XZ); put(r.x0); put(r.x1); put(r.z0); put(r.z1); put(r.k); mat(*r.mp); break;
            }
            case hittable_kind::yz_rect: {
                const auto& r = static_cast<const yz_rect&>(h);
                put(ARTB_T_YZ); put(r.y0); put(r.y1); put(r.z0); put(r.z1); put(r.k); mat(*r.mp); break;
            }
            case hittable_kind::triangle: {
                const auto& t = static_cast<const triangle&>(h);
                put(ARTB_T_TRI); put(t.pt1); put(t.pt2); put(t.pt3); mat(*t.mat_ptr); break;
            }
            case hittable_kind::constant_medium: {
                const auto& m = static_cast<const constant_medium&>(h);
                put(ARTB_T_MEDIUM); put(m.neg_inv_density); mat(*m.phase_function); node(*m.boundary); break;
            }
            case hittable_kind::triangle_soup: {
                // serialised as the list of triangles it stands for
                const auto& s = static_cast<const triangle_soup&>(h);
                put(ARTB_T_LIST); put(double(s.faces.size()));
                for (size_t f = 0; f < s.faces.size(); ++f) {
                    put(ARTB_T_TRI);
                    for (int k = 0; k < 3; ++k) put(s.vertices[size_t(s.faces[f][k])]);
                    mat(*s.materials[s.materials.size() == 1 ? 0 : f]);
                }
                break;
            }
        }
    }
};

} // namespace

std::vector<double> dump_graph(const hittable_list& world)
{
    writer w;
    w.list(world);
    return std::move(w.out);
}

namespace {

struct reader {
    const double* t;
    size_t n, pos = 0;
    const GraphBlob* blobs;
    size_t n_blobs;
    // equal token sequences -> one shared object
    std::map<std::vector<double>, std::shared_ptr<material>> materials;
    std::map<std::vector<double>, std::shared_ptr<texture>> textures;
    std::unordered_map<std::uint64_t, std::shared_ptr<image_texture>> images;
    int depth = 0;

    double get()
    {
        if (pos >= n) throw std::logic_error("graph import: token stream ends early");
        return t[pos++];
    }
    int code()
    {
        const double d = get();
        if (!(d >= 0 && d < 1000) || d != double(int(d))) throw std::logic_error("graph import: bad token");
        return int(d);
    }
    vec3 v3() { const double x = get(), y = get(), z = get(); return vec3(x, y, z); }
    std::uint64_t hash()
    {
        const double lo = get(), hi = get();
        if (!(lo >= 0 && lo < 4294967296.0 && hi >= 0 && hi < 4294967296.0)) throw std::logic_error("graph import: bad hash");
        return std::uint64_t(lo) | (std::uint64_t(hi) << 32);
    }
    const GraphBlob& blob(std::uint64_t h, std::uint64_t bytes)
    {
        for (size_t i = 0; i < n_blobs; ++i)
            if (blobs[i].hash == h) {
                if (blobs[i].bytes != bytes || !blobs[i].data) throw std::logic_error("graph import: blob has the wrong size");
                if (fnv1a64(blobs[i].data, size_t(bytes)) != h) throw std::logic_error("graph import: blob does not match its hash");
                return blobs[i];
            }
        throw std::logic_error("graph import: missing blob (Perlin tables / image texels)");
    }
    std::shared_ptr<image_texture> image()
    {
        if (code() != ARTB_X_IMAGE) throw std::logic_error("graph import: image expected");
        const int w = int(get()), h = int(get()), bpp = int(get());
        const std::uint64_t hh = hash();
        auto it = images.find(hh ^ (std::uint64_t(w) << 40) ^ (std::uint64_t(h) << 20) ^ std::uint64_t(bpp));
        if (it != images.end()) return it->second;
        std::shared_ptr<image_texture> im;
        if (hh == 0 || w <= 0 || h <= 0 || bpp <= 0) {
            im = std::make_shared<image_texture>(std::vector<unsigned char>(), 0, 0, 0);       // "no data": renders cyan
        } else {
            if (bpp < 3 || bpp > 4 || w > 65536 || h > 65536) throw std::logic_error("graph import: unsupported image format");
            const GraphBlob& b = blob(hh, std::uint64_t(w) * std::uint64_t(h) * std::uint64_t(bpp));
            const unsigned char* px = static_cast<const unsigned char*>(b.data);
            im = std::make_shared<image_texture>(std::vector<unsigned char>(px, px + b.bytes), w, h, bpp);
        }
        images.emplace(hh ^ (std::uint64_t(w) << 40) ^ (std::uint64_t(h) << 20) ^ std::uint64_t(bpp), im);
        return im;
    }
    std::shared_ptr<texture> tex()
    {
        const size_t start = pos;
        if (++depth > 64) throw std::logic_error("graph import: textures nested too deep");
        std::shared_ptr<texture> out;
        const int c = code();
        switch (c) {
            case ARTB_X_SOLID: out = std::make_shared<solid_color>(v3()); break;
            case ARTB_X_CHECKER: { auto e = tex(); auto o = tex(); out = std::make_shared<checker_texture>(e, o); break; }
            case ARTB_X_NOISE: {
                const double scale = get();
                const GraphBlob& b = blob(hash(), 9216);
                const double* rv = static_cast<const double*>(b.data);
                const int* perm = reinterpret_cast<const int*>(static_cast<const char*>(b.data) + 256 * 3 * sizeof(double));
                for (int i = 0; i < 768; ++i) if (perm[i] < 0 || perm[i] > 255) throw std::logic_error("graph import: bad Perlin permutation");
                out = std::make_shared<noise_texture>(scale, perlin(rv, perm, perm + 256, perm + 512));
                break;
            }
            case ARTB_X_IMAGE: --pos; out = image(); break;
            case ARTB_X_BARY_IMAGE: {
                const double ua = get(), va = get(), ub = get(), vb = get(), uc = get(), vc = get();
                out = std::make_shared<barycentric_image_texture>(std::make_pair(ua, va), std::make_pair(ub, vb), std::make_pair(uc, vc), image());
                break;
            }
            case ARTB_X_BARY: { const vec3 a = v3(), b = v3(), c3 = v3(); out = std::make_shared<barycentric_texture>(a, b, c3); break; }
            default: throw std::logic_error("graph import: unknown texture token");
        }
        --depth;
        auto ins = textures.emplace(std::vector<double>(t + start, t + pos), out);
        return ins.first->second;
    }
    std::shared_ptr<material> mat()
    {
        const size_t start = pos;
        std::shared_ptr<material> out;
        switch (code()) {
            case ARTB_M_LAMBERTIAN: out = std::make_shared<lambertian>(tex()); break;
            case ARTB_M_METAL: { const vec3 a = v3(); const double f = get(); out = std::make_shared<metal>(a, f); break; }
            case ARTB_M_DIELECTRIC: out = std::make_shared<dielectric>(get()); break;
            case ARTB_M_DIFFUSE_LIGHT: out = std::make_shared<diffuse_light>(tex()); break;
            case ARTB_M_ISOTROPIC: out = std::make_shared<isotropic>(tex()); break;
            default: throw std::logic_error("graph import: unknown material token");
        }
        auto ins = materials.emplace(std::vector<double>(t + start, t + pos), out);
        return ins.first->second;
    }
    hittable_list list()
    {
        if (code() != ARTB_T_LIST) throw std::logic_error("graph import: list expected");
        const double cnt = get();
        if (!(cnt >= 0 && cnt <= double(n))) throw std::logic_error("graph import: bad list length");
        hittable_list l;
        l.objects.reserve(size_t(cnt));
        for (size_t i = 0; i < size_t(cnt); ++i) l.add(node());
        return l;
    }
    std::shared_ptr<hittable> node()
    {
        if (++depth > 4096) throw std::logic_error("graph import: graph nested too deep");
        std::shared_ptr<hittable> out;
        const int c = code();
        switch (c) {
            case ARTB_T_LIST: --pos; out = std::make_shared<hittable_list>(list()); break;
            case ARTB_T_BVH: {
                const vec3 lo = v3(), hi = v3();
                const bool same = get() != 0;
                auto l = node();
                auto r = same ? l : node();
                out = std::make_shared<bvh_node>(l, r, aabb(lo, hi));
                break;
            }
            case ARTB_T_TRANSLATE: { const vec3 off = v3(); out = std::make_shared<translate>(node(), off); break; }
            case ARTB_T_ROTATE_Y: {
                const double sn = get(), cs = get();
                const bool has = get() != 0;
                const vec3 lo = v3(), hi = v3();
                out = std::make_shared<rotate_y>(node(), sn, cs, has, aabb(lo, hi));
                break;
            }
            case ARTB_T_BOX: { const vec3 lo = v3(), hi = v3(); out = std::make_shared<box>(lo, hi, list()); break; }
            case ARTB_T_SPHERE: { const vec3 cn = v3(); const double r = get(); out = std::make_shared<sphere>(cn, r, mat()); break; }
            case ARTB_T_MSPHERE: {
                const vec3 c0 = v3(), c1 = v3();
                const double t0 = get(), t1 = get(), r = get();
                out = std::make_shared<moving_sphere>(c0, c1, t0, t1, r, mat());
                break;
            }
            case ARTB_T_XY: case ARTB_T_XZ: case ARTB_T_YZ: {
                const double a0 = get(), a1 = get(), b0 = get(), b1 = get(), k = get();
                auto m = mat();
                if (c == ARTB_T_XY) out = std::make_shared<xy_rect>(a0, a1, b0, b1, k, m);
                else if (c == ARTB_T_XZ) out = std::make_shared<xz_rect>(a0, a1, b0, b1, k, m);
                else out = std::make_shared<yz_rect>(a0, a1, b0, b1, k, m);
                break;
            }
            case ARTB_T_TRI: { const vec3 a = v3(), b = v3(), c3 = v3(); out = std::make_shared<triangle>(a, b, c3, mat()); break; }
            case ARTB_T_MEDIUM: {
                const double nid = get();
                auto phase = mat();
                if (phase->kind() != material_kind::isotropic) throw std::logic_error("graph import: medium without an isotropic phase function");
                auto cm = std::make_shared<constant_medium>(node(), 1.0, static_cast<const isotropic&>(*phase).albedo);
                cm->neg_inv_density = nid;
                cm->phase_function = phase;
                out = cm;
                break;
            }
            default: throw std::logic_error("graph import: unknown hittable token");
        }
        --depth;
        return out;
    }
};

} // namespace

hittable_list build_graph(const double* tokens, size_t n_tokens, const GraphBlob* blobs, size_t n_blobs)
{
    if (!tokens || n_tokens < 2) throw std::logic_error("graph import: empty token stream");
    reader r{ tokens, n_tokens, 0, blobs, n_blobs };
    hittable_list world = r.list();
    if (r.pos != n_tokens) throw std::logic_error("graph import: trailing tokens");
    return world;
}

} // namespace artb
