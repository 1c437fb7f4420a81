// TEST INFRASTRUCTURE (oracle) — CPU restatement of the reference's hot path. Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this; the product (libartb.so) never does.
//
// Parity status: PINNED. tests/test_oracle_port.py checks this restatement against fixtures produced by the
// UNMODIFIED reference (oracle/_ref/ref_tool, tools/make_golden.py): single-mode renders are bit-identical (RGB8 and
// the float64 per-pixel means) for scene aliases 1-11, and every recorded closest-hit query (ray -> leaf id, t, p,
// normal, u, v, front_face, material, xi consumed by media) is reproduced exactly.
//
// It walks the SAME host scene graph the product flattens (another_raytracer_b200/csrc/host/scene.h — data only, no
// intersection code there) with the reference's recursive algorithm and the reference's global RNG stream
// (artb::global_rng, an MT19937 restatement), consuming random numbers in the reference's order — including GCC's
// right-to-left evaluation of constructor arguments. Each function cites the reference lines it follows.
#include "scene_manager.h"

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

using namespace artb;

#define ORACLE_API extern "C" __attribute__((visibility("default")))

namespace {

struct hit_record {                       // hittable.h:9-23
    point3 p;
    vec3 normal;
    const material* mat_ptr = nullptr;
    double t = 0, u = 0, v = 0;
    bool front_face = false;

    void set_face_normal(const ray& r, const vec3& outward_normal)
    {
        front_face = dot(r.direction(), outward_normal) < 0;
        normal = front_face ? outward_normal : -outward_normal;
    }
};

struct query_ctx {
    const std::unordered_map<const hittable*, int>* leaf_ids = nullptr;   // a triangle_soup maps to the id of its first face
    int last_leaf = -1;                   // id of the last world leaf that reported a hit
    int boundary_depth = 0;               // > 0 while a medium queries its boundary
    const double* xi_row = nullptr;       // explicit xi per medium slot (nullptr: draw from the global stream)
    double* xi_used = nullptr;            // records the xi each medium consumed (NaN when none)
    const std::unordered_map<const hittable*, int>* medium_slots = nullptr;
};

// ---- vec3 sampling helpers: vec3.h:117-154 --------------------------------------------------------
vec3 random_in_unit_sphere()
{
    while (true) {
        const vec3 p = vec3::random(-1, 1);
        if (p.length_squared() >= 1) continue;
        return p;
    }
}
vec3 random_unit_vector() { return unit_vector(random_in_unit_sphere()); }
vec3 random_in_unit_disk()
{
    while (true) {
        const double y = random_double(-1, 1);      // arguments are evaluated right to left: y draws first
        const double x = random_double(-1, 1);
        const vec3 p(x, y, 0);
        if (p.length_squared() >= 1) continue;
        return p;
    }
}
vec3 reflect(const vec3& v, const vec3& n) { return v - 2 * dot(v, n) * n; }
vec3 refract(const vec3& uv, const vec3& n, double etai_over_etat)
{
    const double cos_theta = std::min(dot(-uv, n), 1.0);
    const vec3 r_out_perp = etai_over_etat * (uv + cos_theta * n);
    const vec3 r_out_parallel = -std::sqrt(std::abs(1.0 - r_out_perp.length_squared())) * n;
    return r_out_perp + r_out_parallel;
}

// ---- aabb::hit: aabb.h:16-29 ------------------------------------------------------------------------
bool aabb_hit(const aabb& b, const ray& r, double t_min, double t_max)
{
    for (int a = 0; a < 3; a++) {
        const double invD = 1.0f / r.direction()[a];
        double t0 = (b.min()[a] - r.origin()[a]) * invD;
        double t1 = (b.max()[a] - r.origin()[a]) * invD;
        if (invD < 0.0f) std::swap(t0, t1);
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max <= t_min) return false;
    }
    return true;
}

void get_sphere_uv(const point3& p, double& u, double& v)   // sphere.h:24-37
{
    const double theta = std::acos(-p.y());
    const double phi = std::atan2(-p.z(), p.x()) + pi;
    u = phi / (2 * pi);
    v = theta / pi;
}

bool hit_any(const hittable& h, const ray& r, double t_min, double t_max, hit_record& rec, query_ctx& q);

void note_leaf(const hittable& h, query_ctx& q)
{
    if (q.boundary_depth == 0 && q.leaf_ids) {
        auto it = q.leaf_ids->find(&h);
        if (it != q.leaf_ids->end()) q.last_leaf = it->second;
    }
}

bool hit_sphere_at(const point3& center, double radius, const ray& r, double t_min, double t_max, hit_record& rec, bool with_uv)
{
    // sphere.h:39-65 / moving_sphere.h:33-58
    const vec3 oc = r.origin() - center;
    const double a = r.direction().length_squared();
    const double half_b = dot(oc, r.direction());
    const double c = oc.length_squared() - radius * radius;
    const double discriminant = half_b * half_b - a * c;
    if (discriminant < 0) return false;
    const double sqrtd = std::sqrt(discriminant);
    double root = (-half_b - sqrtd) / a;
    if (root < t_min || t_max < root) {
        root = (-half_b + sqrtd) / a;
        if (root < t_min || t_max < root) return false;
    }
    rec.t = root;
    rec.p = r.at(rec.t);
    const vec3 outward_normal = (rec.p - center) / radius;
    rec.set_face_normal(r, outward_normal);
    if (with_uv) get_sphere_uv(outward_normal, rec.u, rec.v);
    return true;
}

template <int KA, int KB, int KK>
bool hit_rect(double a0, double a1, double b0, double b1, double k, const ray& r, double t_min, double t_max, hit_record& rec)
{
    // aarect.cpp:3-55 (KA, KB = in-plane axes, KK = normal axis)
    const double t = (k - r.origin()[KK]) / r.direction()[KK];
    if (t < t_min || t > t_max) return false;
    const double a = r.origin()[KA] + t * r.direction()[KA];
    const double b = r.origin()[KB] + t * r.direction()[KB];
    if (a < a0 || a > a1 || b < b0 || b > b1) return false;
    rec.u = (a - a0) / (a1 - a0);
    rec.v = (b - b0) / (b1 - b0);
    rec.t = t;
    vec3 outward_normal(0, 0, 0);
    outward_normal[KK] = 1;
    rec.set_face_normal(r, outward_normal);
    rec.p = r.at(t);
    return true;
}

bool hit_triangle(const triangle& tr, const ray& r, double t_min, double t_max, hit_record& rec)
{
    // triangle.h:22-88
    const vec3 v1v2 = tr.pt2 - tr.pt1;
    const vec3 v1v3 = tr.pt3 - tr.pt1;
    const vec3 outward_normal = cross(v1v2, v1v3);
    const double normal_dot_ray_direction = dot(outward_normal, r.direction());
    if (std::fabs(normal_dot_ray_direction) < std::numeric_limits<double>::epsilon()) return false;
    const double d = -dot(outward_normal, tr.pt1);
    const double t = -(dot(outward_normal, r.origin()) + d) / normal_dot_ray_direction;
    if (t < t_min || t_max < t) return false;
    const vec3 p = r.origin() + t * r.direction();
    double u, v;
    vec3 c;
    const vec3 edge1 = tr.pt2 - tr.pt1;
    const vec3 vpt1 = p - tr.pt1;
    c = cross(edge1, vpt1);
    if (dot(outward_normal, c) < 0) return false;
    const vec3 edge2 = tr.pt3 - tr.pt2;
    const vec3 vpt2 = p - tr.pt2;
    c = cross(edge2, vpt2);
    if ((u = dot(outward_normal, c)) < 0) return false;
    const vec3 edge3 = tr.pt1 - tr.pt3;
    const vec3 vpt3 = p - tr.pt3;
    c = cross(edge3, vpt3);
    if ((v = dot(outward_normal, c)) < 0) return false;
    rec.t = t;
    rec.p = p;
    rec.set_face_normal(r, outward_normal);
    rec.u = u / outward_normal.length_squared();
    rec.v = v / outward_normal.length_squared();
    rec.mat_ptr = tr.mat_ptr.get();
    return true;
}

bool hit_list(const hittable_list& l, const ray& r, double t_min, double t_max, hit_record& rec, query_ctx& q)
{
    // hittable_list.cpp:5-19 (temp_rec is uninitialised upstream; zero here — only unused fields differ)
    hit_record temp_rec;
    bool hit_anything = false;
    double closest_so_far = t_max;
    for (const auto& object : l.objects) {
        if (hit_any(*object, r, t_min, closest_so_far, temp_rec, q)) {
            hit_anything = true;
            closest_so_far = temp_rec.t;
            rec = temp_rec;
        }
    }
    return hit_anything;
}

bool hit_medium(const constant_medium& m, const ray& r, double t_min, double t_max, hit_record& rec, query_ctx& q)
{
    // constant_medium.h:37-82
    hit_record rec1, rec2;
    q.boundary_depth++;
    const bool h1 = hit_any(*m.boundary, r, -infinity, infinity, rec1, q);
    const bool h2 = h1 && hit_any(*m.boundary, r, rec1.t + 0.0001, infinity, rec2, q);
    q.boundary_depth--;
    if (!h1 || !h2) return false;
    if (rec1.t < t_min) rec1.t = t_min;
    if (rec2.t > t_max) rec2.t = t_max;
    if (rec1.t >= rec2.t) return false;
    if (rec1.t < 0) rec1.t = 0;
    const double ray_length = r.direction().length();
    const double distance_inside_boundary = (rec2.t - rec1.t) * ray_length;
    double xi;
    int slot = -1;
    if (q.medium_slots) { auto it = q.medium_slots->find(&m); if (it != q.medium_slots->end()) slot = it->second; }
    if (q.xi_row && slot >= 0) {
        xi = q.xi_row[slot];
        if (!(xi == xi)) return false;   // the reference drew nothing for this medium on this ray
    } else {
        xi = random_double();
    }
    if (q.xi_used && slot >= 0) q.xi_used[slot] = xi;
    const double hit_distance = m.neg_inv_density * std::log(xi);
    if (hit_distance > distance_inside_boundary) return false;
    rec.t = rec1.t + hit_distance / ray_length;
    rec.p = r.at(rec.t);
    rec.normal = vec3(1, 0, 0);
    rec.front_face = true;
    rec.mat_ptr = m.phase_function.get();
    return true;
}

bool hit_any(const hittable& h, const ray& r, double t_min, double t_max, hit_record& rec, query_ctx& q)
{
    switch (h.kind()) {
        case hittable_kind::list:
            return hit_list(static_cast<const hittable_list&>(h), r, t_min, t_max, rec, q);
        case hittable_kind::bvh: {
            // bvh.cpp:44-52
            const auto& n = static_cast<const bvh_node&>(h);
            if (!aabb_hit(n.box, r, t_min, t_max)) return false;
            const bool hit_left = hit_any(*n.left, r, t_min, t_max, rec, q);
            const bool hit_right = hit_any(*n.right, r, t_min, hit_left ? rec.t : t_max, rec, q);
            return hit_left || hit_right;
        }
        case hittable_kind::translate: {
            // hittable.cpp:3-12
            const auto& t = static_cast<const translate&>(h);
            const ray moved_r(r.origin() - t.offset, r.direction(), r.time());
            if (!hit_any(*t.ptr, moved_r, t_min, t_max, rec, q)) return false;
            rec.p += t.offset;
            rec.set_face_normal(moved_r, rec.normal);
            return true;
        }
        case hittable_kind::rotate_y: {
            // hittable.cpp:57-85
            const auto& ro = static_cast<const rotate_y&>(h);
            point3 origin = r.origin();
            vec3 direction = r.direction();
            origin[0] = ro.cos_theta * r.origin()[0] - ro.sin_theta * r.origin()[2];
            origin[2] = ro.sin_theta * r.origin()[0] + ro.cos_theta * r.origin()[2];
            direction[0] = ro.cos_theta * r.direction()[0] - ro.sin_theta * r.direction()[2];
            direction[2] = ro.sin_theta * r.direction()[0] + ro.cos_theta * r.direction()[2];
            const ray rotated_r(origin, direction, r.time());
            if (!hit_any(*ro.ptr, rotated_r, t_min, t_max, rec, q)) return false;
            point3 p = rec.p;
            vec3 normal = rec.normal;
            p[0] = ro.cos_theta * rec.p[0] + ro.sin_theta * rec.p[2];
            p[2] = -ro.sin_theta * rec.p[0] + ro.cos_theta * rec.p[2];
            normal[0] = ro.cos_theta * rec.normal[0] + ro.sin_theta * rec.normal[2];
            normal[2] = -ro.sin_theta * rec.normal[0] + ro.cos_theta * rec.normal[2];
            rec.p = p;
            rec.set_face_normal(rotated_r, normal);
            return true;
        }
        case hittable_kind::box:
            return hit_list(static_cast<const box&>(h).sides, r, t_min, t_max, rec, q);   // box.cpp:17-19
        case hittable_kind::sphere: {
            const auto& s = static_cast<const sphere&>(h);
            if (!hit_sphere_at(s.center, s.radius, r, t_min, t_max, rec, true)) return false;
            rec.mat_ptr = s.mat_ptr.get();
            note_leaf(h, q);
            return true;
        }
        case hittable_kind::moving_sphere: {
            const auto& s = static_cast<const moving_sphere&>(h);
            if (!hit_sphere_at(s.center(r.time()), s.radius, r, t_min, t_max, rec, false)) return false;
            rec.mat_ptr = s.mat_ptr.get();
            note_leaf(h, q);
            return true;
        }
        case hittable_kind::xy_rect: {
            const auto& a = static_cast<const xy_rect&>(h);
            if (!hit_rect<0, 1, 2>(a.x0, a.x1, a.y0, a.y1, a.k, r, t_min, t_max, rec)) return false;
            rec.mat_ptr = a.mp.get();
            note_leaf(h, q);
            return true;
        }
        case hittable_kind::xz_rect: {
            const auto& a = static_cast<const xz_rect&>(h);
            if (!hit_rect<0, 2, 1>(a.x0, a.x1, a.z0, a.z1, a.k, r, t_min, t_max, rec)) return false;
            rec.mat_ptr = a.mp.get();
            note_leaf(h, q);
            return true;
        }
        case hittable_kind::yz_rect: {
            const auto& a = static_cast<const yz_rect&>(h);
            if (!hit_rect<1, 2, 0>(a.y0, a.y1, a.z0, a.z1, a.k, r, t_min, t_max, rec)) return false;
            rec.mat_ptr = a.mp.get();
            note_leaf(h, q);
            return true;
        }
        case hittable_kind::triangle: {
            if (!hit_triangle(static_cast<const triangle&>(h), r, t_min, t_max, rec)) return false;
            note_leaf(h, q);
            return true;
        }
        case hittable_kind::constant_medium: {
            if (!hit_medium(static_cast<const constant_medium&>(h), r, t_min, t_max, rec, q)) return false;
            note_leaf(h, q);
            return true;
        }
        case hittable_kind::triangle_soup: {
            // Not a reference class: it stands for the hittable_list of `triangle` objects with the same vertices and
            // materials (that is how graph_dump serialises it), so it is intersected as hittable_list::hit would
            // (hittable_list.cpp:5-19): every face in order, t_max narrowed to the closest so far, later faces win ties.
            const auto& sp = static_cast<const triangle_soup&>(h);
            hit_record temp_rec;
            bool hit_anything = false;
            double closest_so_far = t_max;
            int base = -1;
            if (q.boundary_depth == 0 && q.leaf_ids) { auto it = q.leaf_ids->find(&h); if (it != q.leaf_ids->end()) base = it->second; }
            triangle tr(point3(0, 0, 0), point3(0, 0, 0), point3(0, 0, 0), nullptr);
            for (size_t f = 0; f < sp.faces.size(); ++f) {
                tr.pt1 = sp.vertices[size_t(sp.faces[f][0])];
                tr.pt2 = sp.vertices[size_t(sp.faces[f][1])];
                tr.pt3 = sp.vertices[size_t(sp.faces[f][2])];
                if (hit_triangle(tr, r, t_min, closest_so_far, temp_rec)) {
                    hit_anything = true;
                    closest_so_far = temp_rec.t;
                    rec = temp_rec;
                    rec.mat_ptr = sp.materials[sp.materials.size() == 1 ? 0 : f].get();
                    if (base >= 0) q.last_leaf = base + int(f);
                }
            }
            return hit_anything;
        }
    }
    return false;
}

// ---- textures: texture.h:16-154, perlin.h:21-96 ---------------------------------------------------
double perlin_noise(const perlin& pn, const point3& p)
{
    const double u = p.x() - std::floor(p.x());
    const double v = p.y() - std::floor(p.y());
    const double w = p.z() - std::floor(p.z());
    const int i = static_cast<int>(std::floor(p.x()));
    const int j = static_cast<int>(std::floor(p.y()));
    const int k = static_cast<int>(std::floor(p.z()));
    vec3 c[2][2][2];
    for (int di = 0; di < 2; di++)
        for (int dj = 0; dj < 2; dj++)
            for (int dk = 0; dk < 2; dk++)
                c[di][dj][dk] = pn.ranvec[size_t(pn.perm_x[size_t((i + di) & 255)] ^ pn.perm_y[size_t((j + dj) & 255)] ^
                                                 pn.perm_z[size_t((k + dk) & 255)])];
    const double uu = u * u * (3 - 2 * u);
    const double vv = v * v * (3 - 2 * v);
    const double ww = w * w * (3 - 2 * w);
    double accum = 0.0;
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++)
            for (int d = 0; d < 2; d++) {
                const vec3 weight_v(u - a, v - b, w - d);
                accum += (a * uu + (1 - a) * (1 - uu)) * (b * vv + (1 - b) * (1 - vv)) * (d * ww + (1 - d) * (1 - ww)) *
                         dot(c[a][b][d], weight_v);
            }
    return accum;
}

color image_value(const image_texture& im, double u, double v)
{
    if (im.data.empty()) return color(0, 1, 1);
    u = clamp(u, 0.0, 1.0);
    v = 1.0 - clamp(v, 0.0, 1.0);
    int i = static_cast<int>(u * im.width);
    int j = static_cast<int>(v * im.height);
    if (i >= im.width) i = im.width - 1;
    if (j >= im.height) j = im.height - 1;
    const double color_scale = 1.0 / 255.0;
    const unsigned char* pixel = im.data.data() + j * (im.bytes_per_pixel * im.width) + i * im.bytes_per_pixel;
    return color(color_scale * pixel[0], color_scale * pixel[1], color_scale * pixel[2]);
}

color texture_value(const texture& t, double u, double v, const point3& p)
{
    switch (t.kind()) {
        case texture_kind::solid: return static_cast<const solid_color&>(t).color_value;
        case texture_kind::checker: {
            const auto& c = static_cast<const checker_texture&>(t);
            const double sines = std::sin(10 * p.x()) * std::sin(10 * p.y()) * std::sin(10 * p.z());
            return sines < 0 ? texture_value(*c.odd, u, v, p) : texture_value(*c.even, u, v, p);
        }
        case texture_kind::noise: {
            const auto& n = static_cast<const noise_texture&>(t);
            return color(1, 1, 1) * 0.5 * (1.0 + perlin_noise(n.noise, n.scale * p));
        }
        case texture_kind::image: return image_value(static_cast<const image_texture&>(t), u, v);
        case texture_kind::barycentric_image: {
            const auto& b = static_cast<const barycentric_image_texture&>(t);
            return image_value(*b.tex, u * b.texcoord_a.first + v * b.texcoord_b.first + (1 - u - v) * b.texcoord_c.first,
                               u * b.texcoord_a.second + v * b.texcoord_b.second + (1 - u - v) * b.texcoord_c.second);
        }
        case texture_kind::barycentric: {
            const auto& b = static_cast<const barycentric_texture&>(t);
            return u * b.color_a + v * b.color_b + (1.0 - u - v) * b.color_c;
        }
    }
    return color(0, 0, 0);
}

// ---- materials: material.h:10-135 -------------------------------------------------------------------
color emitted(const material& m, double u, double v, const point3& p)
{
    if (m.kind() == material_kind::diffuse_light) return texture_value(*static_cast<const diffuse_light&>(m).emit, u, v, p);
    return color(0, 0, 0);
}

// Where scatter() and get_ray() take their random points from. ref_sampler: the reference's rejection loops on the
// global MT19937 stream (vec3.h:117-143) — what every pinned test and every render uses. xi_sampler: explicit numbers
// through the closed forms the DEVICE uses (rt_shade.cuh: same distributions), for the value-by-value shade parity
// hooks oracle_shade / oracle_camera_rays: both sides then evaluate the same expressions on the same xi.
struct ref_sampler {
    vec3 unit_vector_on_sphere() { return random_unit_vector(); }
    vec3 in_unit_sphere() { return random_in_unit_sphere(); }
    vec3 in_unit_disk() { return random_in_unit_disk(); }
    double uniform() { return random_double(); }
    double uniform(double a, double b) { return random_double(a, b); }
};

struct xi_sampler {
    const double* xi;
    double next() { return *xi++; }
    vec3 unit_vector_on_sphere()
    {
        const double z = 1.0 - 2.0 * next();
        const double rxy = std::sqrt(std::max(0.0, 1.0 - z * z));
        const double phi = 2.0 * next();                       // in units of pi, like sincospi on the device
        return vec3(rxy * std::cos(pi * phi), rxy * std::sin(pi * phi), z);
    }
    vec3 in_unit_sphere()
    {
        const vec3 dir = unit_vector_on_sphere();
        return std::cbrt(next()) * dir;
    }
    vec3 in_unit_disk()
    {
        const double r = std::sqrt(next());
        const double phi = 2.0 * next();
        return vec3(r * std::cos(pi * phi), r * std::sin(pi * phi), 0);
    }
    double uniform() { return next(); }
    double uniform(double a, double b) { return a + (b - a) * next(); }      // tracer_utils.h:33-35
};

template <class Sampler>
bool scatter_with(const material& m, const ray& r_in, const hit_record& rec, color& attenuation, ray& scattered, Sampler& smp)
{
    switch (m.kind()) {
        case material_kind::lambertian: {
            vec3 scatter_direction = rec.normal + smp.unit_vector_on_sphere();
            if (scatter_direction.near_zero()) scatter_direction = rec.normal;
            scattered = ray(rec.p, scatter_direction, r_in.time());
            attenuation = texture_value(*static_cast<const lambertian&>(m).albedo, rec.u, rec.v, rec.p);
            return true;
        }
        case material_kind::metal: {
            const auto& me = static_cast<const metal&>(m);
            const vec3 reflected = reflect(unit_vector(r_in.direction()), rec.normal);
            scattered = ray(rec.p, reflected + me.fuzz * smp.in_unit_sphere(), r_in.time());
            attenuation = me.albedo;
            return dot(scattered.direction(), rec.normal) > 0;
        }
        case material_kind::dielectric: {
            const double ir = static_cast<const dielectric&>(m).ir;
            attenuation = color(1.0, 1.0, 1.0);
            const double refraction_ratio = rec.front_face ? (1.0 / ir) : ir;
            const vec3 unit_direction = unit_vector(r_in.direction());
            const double cos_theta = std::min(dot(-unit_direction, rec.normal), 1.0);
            const double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
            const bool cannot_refract = refraction_ratio * sin_theta > 1.0;
            vec3 direction;
            auto reflectance = [](double cosine, double ref_idx) {
                double r0 = (1 - ref_idx) / (1 + ref_idx);
                r0 = r0 * r0;
                return r0 + (1 - r0) * std::pow((1 - cosine), 5);
            };
            if (cannot_refract || reflectance(cos_theta, refraction_ratio) > smp.uniform())
                direction = reflect(unit_direction, rec.normal);
            else
                direction = refract(unit_direction, rec.normal, refraction_ratio);
            scattered = ray(rec.p, direction, r_in.time());
            return true;
        }
        case material_kind::diffuse_light: return false;
        case material_kind::isotropic: {
            scattered = ray(rec.p, smp.in_unit_sphere(), r_in.time());
            attenuation = texture_value(*static_cast<const isotropic&>(m).albedo, rec.u, rec.v, rec.p);
            return true;
        }
    }
    return false;
}

bool scatter(const material& m, const ray& r_in, const hit_record& rec, color& attenuation, ray& scattered)
{
    ref_sampler smp;
    return scatter_with(m, r_in, rec, attenuation, scattered, smp);
}

// camera::get_ray: camera.h:38-47
template <class Sampler>
ray get_ray_with(const camera& c, double s, double t, Sampler& smp)
{
    const vec3 rd = c.lens_radius * smp.in_unit_disk();
    const vec3 offset = c.u * rd.x() + c.v * rd.y();
    const double tm = smp.uniform(c.time0, c.time1);
    return ray(c.origin + offset, c.lower_left_corner + s * c.horizontal + t * c.vertical - c.origin - offset, tm);
}

ray get_ray(const camera& c, double s, double t)
{
    ref_sampler smp;
    return get_ray_with(c, s, t, smp);
}

// optional log of every closest-hit query of a render (origin, direction, time, remaining depth): lets the tests and
// tools/ replay realistic ray sets (all bounces of whole paths) through the device traversal
double* g_ray_log = nullptr;
std::uint64_t g_ray_log_cap = 0, g_ray_log_n = 0;

// engine::_ray_color: engine.h:447-466
color ray_color(const ray& r, const color& background, const hittable& world, int depth, query_ctx& q, std::uint64_t* rays)
{
    hit_record rec;
    if (depth <= 0) return color(0, 0, 0);
    if (rays) ++*rays;
    if (g_ray_log && g_ray_log_n < g_ray_log_cap) {
        double* e = g_ray_log + 8 * g_ray_log_n++;
        for (int k = 0; k < 3; ++k) { e[k] = r.origin()[k]; e[3 + k] = r.direction()[k]; }
        e[6] = r.time(); e[7] = double(depth);
    }
    if (!hit_any(world, r, 0.001, infinity, rec, q)) return background;
    ray scattered;
    color attenuation;
    const color emit = emitted(*rec.mat_ptr, rec.u, rec.v, rec.p);
    if (!scatter(*rec.mat_ptr, r, rec, attenuation, scattered)) return emit;
    return emit + attenuation * ray_color(scattered, background, world, depth - 1, q, rays);
}

int material_code(const material* m)
{
    switch (m->kind()) {
        case material_kind::lambertian: return 0;
        case material_kind::metal: return 1;
        case material_kind::dielectric: return 2;
        case material_kind::diffuse_light: return 3;
        case material_kind::isotropic: return 4;
    }
    return -1;
}

struct oracle_scene {
    scene s;
    std::unordered_map<const hittable*, int> leaf_ids, medium_slots;
    int n_leaves = 0, n_media = 0;
    double rng_next = 0;

    void number(const hittable& h, bool in_boundary)
    {
        switch (h.kind()) {
            case hittable_kind::list: for (auto& o : static_cast<const hittable_list&>(h).objects) number(*o, in_boundary); break;
            case hittable_kind::bvh: {
                const auto& n = static_cast<const bvh_node&>(h);
                number(*n.left, in_boundary);
                if (n.left.get() != n.right.get()) number(*n.right, in_boundary);
                break;
            }
            case hittable_kind::translate: number(*static_cast<const translate&>(h).ptr, in_boundary); break;
            case hittable_kind::rotate_y: number(*static_cast<const rotate_y&>(h).ptr, in_boundary); break;
            case hittable_kind::box: for (auto& o : static_cast<const box&>(h).sides.objects) number(*o, in_boundary); break;
            case hittable_kind::constant_medium: {
                if (!in_boundary) { leaf_ids[&h] = n_leaves; medium_slots[&h] = n_media++; }
                n_leaves++;
                number(*static_cast<const constant_medium&>(h).boundary, true);
                break;
            }
            case hittable_kind::triangle_soup:        // one id per face, in face order (what the flattener hands out)
                if (!in_boundary) leaf_ids[&h] = n_leaves;
                n_leaves += int(static_cast<const triangle_soup&>(h).faces.size());
                break;
            default:
                if (!in_boundary) leaf_ids[&h] = n_leaves;
                n_leaves++;
        }
    }
};

} // namespace

struct oracle_hit {
    double t, p[3], normal[3], u, v;
    std::int32_t hit, prim_id, front_face, material_type;
};

// mesh_as_soup / heightfield_side: the scene_manager options of include/artb.h's artb_scene_options (0 = default)
ORACLE_API void* oracle_scene_build_opt(int alias, const char* assets_root, int mesh_as_soup, int heightfield_side)
{
    try {
        if (assets_root) set_assets_root(assets_root);
        reset_global_rng();
        auto* o = new oracle_scene;
        scene_manager mgr;
        mgr.mesh_as_soup = mesh_as_soup != 0;
        if (heightfield_side > 0) mgr.heightfield_side = heightfield_side;
        o->s = mgr.build(static_cast<scene_alias>(alias));
        o->number(o->s.objects, false);
        return o;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "oracle_scene_build: %s\n", e.what());
        return nullptr;
    }
}

ORACLE_API void* oracle_scene_build(int alias, const char* assets_root) { return oracle_scene_build_opt(alias, assets_root, 0, 0); }

ORACLE_API void oracle_scene_free(void* h) { delete static_cast<oracle_scene*>(h); }
ORACLE_API int oracle_scene_n_media(void* h) { return static_cast<oracle_scene*>(h)->n_media; }
ORACLE_API int oracle_scene_n_leaves(void* h) { return static_cast<oracle_scene*>(h)->n_leaves; }
ORACLE_API void oracle_ray_log_begin(double* buf, std::uint64_t capacity) { g_ray_log = buf; g_ray_log_cap = capacity; g_ray_log_n = 0; }
ORACLE_API std::uint64_t oracle_ray_log_end() { g_ray_log = nullptr; return g_ray_log_n; }
ORACLE_API double oracle_rng_next() { return random_double(); }
ORACLE_API void oracle_rng_burn(long n) { for (long i = 0; i < n; ++i) random_double(); }
// Re-seed the render stream AFTER the scene is built (the scene itself always comes from the default seed, like the
// reference's). The reference has no seed API; independent sample sets for noise estimates need streams that do not
// overlap, which offsets into one stream cannot give once a render consumes more draws than the offset.
ORACLE_API void oracle_rng_reseed(unsigned seed) { reset_global_rng(seed); }

// closest hit for explicit rays; xi: n x n_media table (NULL: media draw from the global stream, in list order,
// exactly like the reference); xi_used (optional): receives the xi each medium consumed
ORACLE_API int oracle_trace(void* h, const double* rays, std::uint64_t n, double t_min, double t_max, const double* xi,
                            double* xi_used, oracle_hit* out)
{
    try {
        auto* o = static_cast<oracle_scene*>(h);
        const double nan_v = std::numeric_limits<double>::quiet_NaN();
        for (std::uint64_t i = 0; i < n; ++i) {
            const double* r7 = rays + i * 7;
            const ray r(point3(r7[0], r7[1], r7[2]), vec3(r7[3], r7[4], r7[5]), r7[6]);
            query_ctx q;
            q.leaf_ids = &o->leaf_ids;
            q.medium_slots = &o->medium_slots;
            q.xi_row = xi ? xi + i * std::uint64_t(o->n_media) : nullptr;
            if (xi_used) {
                q.xi_used = xi_used + i * std::uint64_t(o->n_media);
                for (int k = 0; k < o->n_media; ++k) q.xi_used[k] = nan_v;
            }
            hit_record rec;
            const bool ok = hit_any(o->s.objects, r, t_min, t_max, rec, q);
            oracle_hit& w = out[i];
            std::memset(&w, 0, sizeof(w));
            w.hit = ok ? 1 : 0;
            w.prim_id = ok ? q.last_leaf : -1;
            w.material_type = ok ? material_code(rec.mat_ptr) : -1;
            if (ok) {
                w.t = rec.t; w.u = rec.u; w.v = rec.v; w.front_face = rec.front_face ? 1 : 0;
                for (int k = 0; k < 3; ++k) { w.p[k] = rec.p[k]; w.normal[k] = rec.normal[k]; }
            }
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "oracle_trace: %s\n", e.what());
        return -1;
    }
}

// Shade-stage parity hook (the counterpart of artb_shade_batch): closest hit of every ray with the reference's
// algorithm (medium_xi as in oracle_trace), then material::emitted and material::scatter of that hit with the random
// numbers xi4[i*4..] fed through the device's closed-form samplers (xi_sampler above).
struct oracle_shade_record {
    double emitted[3], attenuation[3], origin[3], direction[3], time;
    std::int32_t scattered, material_type, front_face, prim_id;      // prim_id = the hit's leaf id (-1: miss, record is zero)
};

ORACLE_API int oracle_shade(void* h, const double* rays, std::uint64_t n, const double* medium_xi, const double* xi4,
                            oracle_hit* hits_out, oracle_shade_record* out)
{
    try {
        auto* o = static_cast<oracle_scene*>(h);
        for (std::uint64_t i = 0; i < n; ++i) {
            const double* r7 = rays + i * 7;
            const ray r(point3(r7[0], r7[1], r7[2]), vec3(r7[3], r7[4], r7[5]), r7[6]);
            query_ctx q;
            q.leaf_ids = &o->leaf_ids;
            q.medium_slots = &o->medium_slots;
            q.xi_row = medium_xi ? medium_xi + i * std::uint64_t(o->n_media) : nullptr;
            hit_record rec;
            const bool ok = hit_any(o->s.objects, r, 0.001, infinity, rec, q);
            oracle_shade_record& w = out[i];
            std::memset(&w, 0, sizeof(w));
            w.prim_id = ok ? q.last_leaf : -1;
            w.material_type = -1;
            if (hits_out) {
                oracle_hit& hh = hits_out[i];
                std::memset(&hh, 0, sizeof(hh));
                hh.hit = ok ? 1 : 0; hh.prim_id = w.prim_id; hh.material_type = ok ? material_code(rec.mat_ptr) : -1;
                if (ok) {
                    hh.t = rec.t; hh.u = rec.u; hh.v = rec.v; hh.front_face = rec.front_face ? 1 : 0;
                    for (int k = 0; k < 3; ++k) { hh.p[k] = rec.p[k]; hh.normal[k] = rec.normal[k]; }
                }
            }
            if (!ok) continue;
            w.material_type = material_code(rec.mat_ptr);
            w.front_face = rec.front_face ? 1 : 0;
            const color e = emitted(*rec.mat_ptr, rec.u, rec.v, rec.p);
            xi_sampler smp{ xi4 + i * 4 };
            color att(0, 0, 0);
            ray sc;
            const bool s_ok = scatter_with(*rec.mat_ptr, r, rec, att, sc, smp);
            w.scattered = s_ok ? 1 : 0;
            for (int k = 0; k < 3; ++k) w.emitted[k] = e[k];
            if (rec.mat_ptr->kind() != material_kind::diffuse_light) {
                for (int k = 0; k < 3; ++k) { w.attenuation[k] = att[k]; w.origin[k] = sc.origin()[k]; w.direction[k] = sc.direction()[k]; }
                w.time = sc.time();
            }
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "oracle_shade: %s\n", e.what());
        return -1;
    }
}

// camera::camera + camera::get_ray (camera.h:8-47) for explicit (s, t) and xi3 = (disk radius, disk angle, time)
ORACLE_API int oracle_camera_rays(const double* cam9, double vfov, double aspect, double aperture, double focus_dist, double time0,
                                  double time1, const double* st, const double* xi3, std::uint64_t n, double* rays_out)
{
    try {
        const camera cam(point3(cam9[0], cam9[1], cam9[2]), point3(cam9[3], cam9[4], cam9[5]), vec3(cam9[6], cam9[7], cam9[8]), vfov,
                         aspect, aperture, focus_dist, time0, time1);
        for (std::uint64_t i = 0; i < n; ++i) {
            xi_sampler smp{ xi3 + i * 3 };
            const ray r = get_ray_with(cam, st[i * 2], st[i * 2 + 1], smp);
            double* w = rays_out + i * 7;
            for (int k = 0; k < 3; ++k) { w[k] = r.origin()[k]; w[3 + k] = r.direction()[k]; }
            w[6] = r.time();
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "oracle_camera_rays: %s\n", e.what());
        return -1;
    }
}

// engine::_run_single (engine.h:70-94) with runtime W/H/spp. Rows [row_begin,row_end) only (the whole image when
// 0,H): lets bench.py time a bounded sample. Returns elapsed ms; rays_out counts closest-hit queries.
ORACLE_API double oracle_render(void* h, int W, int H, int spp, int max_depth, int row_begin, int row_end,
                                std::uint8_t* rgb_out, double* lin_out, std::uint64_t* rays_out)
{
    try {
        auto* o = static_cast<oracle_scene*>(h);
        const camera cam(o->s.lookfrom, o->s.lookat, vec3(0, 1, 0), o->s.vfov, double(W) / double(H), o->s.aperture, 10.0, 0.0, 1.0);
        query_ctx q;
        std::uint64_t rays = 0;
        const auto start = std::chrono::steady_clock::now();
        for (int j = row_begin; j < row_end; ++j) {
            for (int i = 0; i < W; ++i) {
                color pixel_color(0, 0, 0);
                for (int s = 0; s < spp; ++s) {
                    const double u = (i + random_double()) / (W - 1);
                    const double v = ((H - 1 - j) + random_double()) / (H - 1);
                    const ray r = get_ray(cam, u, v);
                    pixel_color += ray_color(r, o->s.background, o->s.objects, max_depth, q, &rays);
                }
                const size_t off = (size_t(j) * size_t(W) + size_t(i)) * 3;
                if (rgb_out) {
                    // write_color: color.h:6-22
                    const double scale = 1.0 / spp;
                    for (int k = 0; k < 3; ++k)
                        rgb_out[off + size_t(k)] = static_cast<std::uint8_t>(256 * clamp(std::sqrt(scale * pixel_color[k]), 0.0, 0.999));
                }
                if (lin_out) for (int k = 0; k < 3; ++k) lin_out[off + size_t(k)] = pixel_color[k] / spp;
            }
        }
        const auto end = std::chrono::steady_clock::now();
        if (rays_out) *rays_out = rays;
        return std::chrono::duration<double, std::milli>(end - start).count();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "oracle_render: %s\n", e.what());
        return -1.0;
    }
}
