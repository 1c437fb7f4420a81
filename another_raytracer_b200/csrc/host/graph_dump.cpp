#include "graph_dump.h"

#include "artb_tokens.h"

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

} // namespace artb
