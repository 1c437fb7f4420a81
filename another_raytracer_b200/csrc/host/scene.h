// Host scene-graph classes of the B200 path-tracing backend.
//
// These keep the constructor signatures and public member names of the reference's L2 layer, so that
// scene-building code (e.g. the reference's scene_manager.cpp) ports by changing the namespace only.
// Unlike the reference they carry NO intersection or shading code: the host graph exists to be
// (a) built with the same RNG consumption as the reference, (b) flattened (flatten.h) into the SoA
// arrays the sm_100a kernels traverse. Each class cites the reference declaration it stands in for.
//
//   textures   solid_color, checker_texture, noise_texture(+perlin), image_texture,
//              barycentric_texture, barycentric_image_texture        reference: src/rendering/texture.h:11-154,
//                                                                               src/rendering/perlin.h:8-97
//   materials  lambertian, metal, dielectric, diffuse_light, isotropic  reference: src/rendering/material.h:10-135
//   hittables  hittable_list, bvh_node, translate, rotate_y, box, sphere, moving_sphere,
//              xy_rect/xz_rect/yz_rect, triangle, constant_medium    reference: src/engine/hittable*.h,
//                                                                    src/engine/constant_medium.h, src/primitives/*.h
//   camera                                                           reference: src/engine/camera.h:6-57
//   triangle_soup (new: a whole indexed mesh as ONE host object, for 1 M-triangle scenes — SURVEY 8f rank 2)
#ifndef ARTB_HOST_SCENE_H
#define ARTB_HOST_SCENE_H

#include "core.h"

#include <array>
#include <string>
#include <utility>
#include <vector>

namespace artb {

// ------------------------------------------------------------------------------------------------
// textures
// ------------------------------------------------------------------------------------------------
enum class texture_kind { solid, checker, noise, image, barycentric, barycentric_image };

class texture {
public:
    virtual ~texture() = default;
    virtual texture_kind kind() const = 0;
};

class solid_color final : public texture {
public:
    solid_color(color c) : color_value(c) {}
    solid_color(double r, double g, double b) : color_value(r, g, b) {}
    texture_kind kind() const override { return texture_kind::solid; }
    color color_value;
};

class checker_texture final : public texture {
public:
    checker_texture(std::shared_ptr<texture> _even, std::shared_ptr<texture> _odd) : even(std::move(_even)), odd(std::move(_odd)) {}
    checker_texture(color c1, color c2) : even(std::make_shared<solid_color>(c1)), odd(std::make_shared<solid_color>(c2)) {}
    texture_kind kind() const override { return texture_kind::checker; }
    std::shared_ptr<texture> even, odd;
};

// Perlin gradient tables. Construction draws 3*256 doubles for the gradients, then 3*255 for the three
// Fisher-Yates permutations, from the global stream — reference: perlin.h:10-19, 63-77.
class perlin {
public:
    static constexpr int point_count = 256;
    perlin();
    // tables taken as they are (a graph imported from a foreign host: no draws from the global stream)
    perlin(const double* ranvec_xyz, const int* px, const int* py, const int* pz);
    std::array<vec3, point_count> ranvec;
    std::array<int, point_count> perm_x, perm_y, perm_z;
private:
    static void generate_perm(int* p);
};

class noise_texture final : public texture {
public:
    noise_texture(double sc) : scale(sc) {}
    noise_texture(double sc, const perlin& tables) : noise(tables), scale(sc) {}
    texture_kind kind() const override { return texture_kind::noise; }
    perlin noise;
    double scale;
};

// Decoded 8-bit image; nearest-texel lookup happens on the device. A file that cannot be read leaves
// data empty, which renders as solid cyan exactly like the reference (texture.h:88-90).
class image_texture final : public texture {
public:
    explicit image_texture(const std::string& filename);
    image_texture(std::vector<unsigned char> texels, int w, int h, int bpp)
        : data(std::move(texels)), width(w), height(h), bytes_per_pixel(bpp) {}
    texture_kind kind() const override { return texture_kind::image; }
    std::vector<unsigned char> data;
    int width = 0, height = 0, bytes_per_pixel = 0;
};

class barycentric_texture final : public texture {
public:
    barycentric_texture(color a, color b, color c) : color_a(a), color_b(b), color_c(c) {}
    texture_kind kind() const override { return texture_kind::barycentric; }
    color color_a, color_b, color_c;
};

class barycentric_image_texture final : public texture {
public:
    using uv = std::pair<double, double>;
    barycentric_image_texture(uv a, uv b, uv c, std::shared_ptr<image_texture> _tex)
        : texcoord_a(a), texcoord_b(b), texcoord_c(c), tex(std::move(_tex)) {}
    texture_kind kind() const override { return texture_kind::barycentric_image; }
    uv texcoord_a, texcoord_b, texcoord_c;
    std::shared_ptr<image_texture> tex;
};

// ------------------------------------------------------------------------------------------------
// materials
// ------------------------------------------------------------------------------------------------
enum class material_kind { lambertian, metal, dielectric, diffuse_light, isotropic };

class material {
public:
    virtual ~material() = default;
    virtual material_kind kind() const = 0;
};

class lambertian final : public material {
public:
    explicit lambertian(const color& a) : albedo(std::make_shared<solid_color>(a)) {}
    explicit lambertian(std::shared_ptr<texture> a) : albedo(std::move(a)) {}
    material_kind kind() const override { return material_kind::lambertian; }
    std::shared_ptr<texture> albedo;
};

class metal final : public material {
public:
    metal(const color& a, double f) : albedo(a), fuzz(f < 1. ? f : 1.) {}
    material_kind kind() const override { return material_kind::metal; }
    color albedo;
    double fuzz;
};

class dielectric final : public material {
public:
    explicit dielectric(double index_of_refraction) : ir(index_of_refraction) {}
    material_kind kind() const override { return material_kind::dielectric; }
    double ir;
};

class diffuse_light final : public material {
public:
    explicit diffuse_light(std::shared_ptr<texture> a) : emit(std::move(a)) {}
    explicit diffuse_light(color c) : emit(std::make_shared<solid_color>(c)) {}
    material_kind kind() const override { return material_kind::diffuse_light; }
    std::shared_ptr<texture> emit;
};

class isotropic final : public material {
public:
    explicit isotropic(color c) : albedo(std::make_shared<solid_color>(c)) {}
    explicit isotropic(std::shared_ptr<texture> a) : albedo(std::move(a)) {}
    material_kind kind() const override { return material_kind::isotropic; }
    std::shared_ptr<texture> albedo;
};

// ------------------------------------------------------------------------------------------------
// hittables
// ------------------------------------------------------------------------------------------------
enum class hittable_kind {
    list, bvh, translate, rotate_y, box, sphere, moving_sphere, xy_rect, xz_rect, yz_rect, triangle,
    constant_medium, triangle_soup
};

class hittable {
public:
    virtual ~hittable() = default;
    virtual hittable_kind kind() const = 0;
    // same contract as the reference (hittable.h:28): false when the object has no box
    virtual bool bounding_box(double time0, double time1, aabb& output_box) const = 0;
};

class hittable_list : public hittable {
public:
    hittable_list() = default;
    explicit hittable_list(std::shared_ptr<hittable> object) { add(std::move(object)); }

    bool empty() const { return objects.empty(); }
    size_t size() const { return objects.size(); }
    void clear() { objects.clear(); }
    void add(std::shared_ptr<hittable> object) { objects.push_back(std::move(object)); }

    hittable_kind kind() const override { return hittable_kind::list; }
    bool bounding_box(double time0, double time1, aabb& output_box) const override;

    std::vector<std::shared_ptr<hittable>> objects;
};

// Same tree as the reference builds (bvh.cpp:3-42): one random_int(0,2) axis per node drawn in
// pre-order, libstdc++ std::sort on the box minimum at time (0,0), median split, span 1 => left==right,
// span 2 => ordered pair. The device does NOT traverse this tree; it is kept because (a) building it
// consumes the global RNG, which later scene content depends on, (b) its shape decides the reference's
// tie-break order and its zero-thickness-box culling, which the flattener reproduces.
// The reference copies the whole object vector at every node (O(N^2)); sorting one shared array in
// place visits the same sub-ranges with the same contents, so the result is identical in O(N log^2 N).
class bvh_node final : public hittable {
public:
    bvh_node(const hittable_list& list, double time0, double time1);
    bvh_node(std::vector<std::shared_ptr<hittable>>& objects, size_t start, size_t end, double time0, double time1);
    // an existing node taken as it is (imported graph): no sort, no draw from the global stream
    bvh_node(std::shared_ptr<hittable> l, std::shared_ptr<hittable> r, const aabb& b) : left(std::move(l)), right(std::move(r)), box(b) {}

    hittable_kind kind() const override { return hittable_kind::bvh; }
    bool bounding_box(double, double, aabb& output_box) const override { output_box = box; return true; }

    std::shared_ptr<hittable> left, right;
    aabb box;
private:
    void init(std::vector<std::shared_ptr<hittable>>& objects, size_t start, size_t end, double time0, double time1);
};

class translate final : public hittable {
public:
    translate(std::shared_ptr<hittable> p, const vec3& displacement) : ptr(std::move(p)), offset(displacement) {}
    hittable_kind kind() const override { return hittable_kind::translate; }
    bool bounding_box(double time0, double time1, aabb& output_box) const override;
    std::shared_ptr<hittable> ptr;
    vec3 offset;
};

class rotate_y final : public hittable {
public:
    rotate_y(std::shared_ptr<hittable> p, double angle);
    rotate_y(std::shared_ptr<hittable> p, double sin_t, double cos_t, bool has, const aabb& b)     // imported graph
        : ptr(std::move(p)), sin_theta(sin_t), cos_theta(cos_t), hasbox(has), bbox(b) {}
    hittable_kind kind() const override { return hittable_kind::rotate_y; }
    bool bounding_box(double, double, aabb& output_box) const override { output_box = bbox; return hasbox; }
    std::shared_ptr<hittable> ptr;
    double sin_theta, cos_theta;
    bool hasbox;
    aabb bbox;
};

class sphere final : public hittable {
public:
    sphere(point3 cen, double r, std::shared_ptr<material> m) : center(cen), radius(r), mat_ptr(std::move(m)) {}
    hittable_kind kind() const override { return hittable_kind::sphere; }
    bool bounding_box(double time0, double time1, aabb& output_box) const override;
    point3 center;
    double radius;
    std::shared_ptr<material> mat_ptr;
};

class moving_sphere final : public hittable {
public:
    moving_sphere(point3 cen0, point3 cen1, double _time0, double _time1, double r, std::shared_ptr<material> m)
        : center0(cen0), center1(cen1), time0(_time0), time1(_time1), radius(r), mat_ptr(std::move(m)) {}
    hittable_kind kind() const override { return hittable_kind::moving_sphere; }
    bool bounding_box(double _time0, double _time1, aabb& output_box) const override;
    point3 center(double time) const;
    point3 center0, center1;
    double time0, time1;
    double radius;
    std::shared_ptr<material> mat_ptr;
};

// The three axis-aligned rectangle classes. Member names follow the reference (aarect.h:8-66).
class xy_rect final : public hittable {
public:
    xy_rect(double _x0, double _x1, double _y0, double _y1, double _k, std::shared_ptr<material> mat)
        : x0(_x0), x1(_x1), y0(_y0), y1(_y1), k(_k), mp(std::move(mat)) {}
    hittable_kind kind() const override { return hittable_kind::xy_rect; }
    bool bounding_box(double, double, aabb& b) const override
    { b = aabb(point3(x0, y0, k - 0.0001), point3(x1, y1, k + 0.0001)); return true; }
    double x0, x1, y0, y1, k;
    std::shared_ptr<material> mp;
};

class xz_rect final : public hittable {
public:
    xz_rect(double _x0, double _x1, double _z0, double _z1, double _k, std::shared_ptr<material> mat)
        : x0(_x0), x1(_x1), z0(_z0), z1(_z1), k(_k), mp(std::move(mat)) {}
    hittable_kind kind() const override { return hittable_kind::xz_rect; }
    bool bounding_box(double, double, aabb& b) const override
    { b = aabb(point3(x0, k - 0.0001, z0), point3(x1, k + 0.0001, z1)); return true; }
    double x0, x1, z0, z1, k;
    std::shared_ptr<material> mp;
};

class yz_rect final : public hittable {
public:
    yz_rect(double _y0, double _y1, double _z0, double _z1, double _k, std::shared_ptr<material> mat)
        : y0(_y0), y1(_y1), z0(_z0), z1(_z1), k(_k), mp(std::move(mat)) {}
    hittable_kind kind() const override { return hittable_kind::yz_rect; }
    bool bounding_box(double, double, aabb& b) const override
    { b = aabb(point3(k - 0.0001, y0, z0), point3(k + 0.0001, y1, z1)); return true; }
    double y0, y1, z0, z1, k;
    std::shared_ptr<material> mp;
};

// Six rectangles in the reference's order: +z, -z, +y, -y, +x, -x faces (box.cpp:3-15).
class box final : public hittable {
public:
    box(const point3& p0, const point3& p1, std::shared_ptr<material> ptr);
    box(const point3& p0, const point3& p1, hittable_list six_sides) : box_min(p0), box_max(p1), sides(std::move(six_sides)) {}   // imported graph
    hittable_kind kind() const override { return hittable_kind::box; }
    bool bounding_box(double, double, aabb& b) const override { b = aabb(box_min, box_max); return true; }
    point3 box_min, box_max;
    hittable_list sides;
};

class triangle final : public hittable {
public:
    triangle(point3 _pt1, point3 _pt2, point3 _pt3, std::shared_ptr<material> m)
        : pt1(_pt1), pt2(_pt2), pt3(_pt3), mat_ptr(std::move(m)) {}
    hittable_kind kind() const override { return hittable_kind::triangle; }
    bool bounding_box(double, double, aabb& b) const override
    { b = aabb(min(pt1, min(pt2, pt3)), max(pt1, max(pt2, pt3))); return true; }   // exact, NOT padded (triangle.h:90-95)
    point3 pt1, pt2, pt3;
    std::shared_ptr<material> mat_ptr;
};

class constant_medium final : public hittable {
public:
    constant_medium(std::shared_ptr<hittable> b, double d, std::shared_ptr<texture> a)
        : boundary(std::move(b)), neg_inv_density(-1 / d), phase_function(std::make_shared<isotropic>(std::move(a))) {}
    constant_medium(std::shared_ptr<hittable> b, double d, color c)
        : boundary(std::move(b)), neg_inv_density(-1 / d), phase_function(std::make_shared<isotropic>(c)) {}
    hittable_kind kind() const override { return hittable_kind::constant_medium; }
    bool bounding_box(double time0, double time1, aabb& b) const override { return boundary->bounding_box(time0, time1, b); }
    std::shared_ptr<hittable> boundary;
    double neg_inv_density;
    std::shared_ptr<material> phase_function;
};

// NEW (not in the reference): an indexed triangle mesh held as one object. It flattens to the same
// per-triangle device records as N `triangle` objects would, without N shared_ptr allocations and
// without the reference-shaped bvh_node above it (the device BVH is built over the flat triangles
// anyway). Triangle semantics on the device are those of triangle.h:22-88.
class triangle_soup final : public hittable {
public:
    hittable_kind kind() const override { return hittable_kind::triangle_soup; }
    bool bounding_box(double, double, aabb& b) const override;
    std::vector<vec3> vertices;
    std::vector<std::array<int, 3>> faces;
    std::vector<std::shared_ptr<material>> materials;   // size 1 (shared) or one per face
};

// ------------------------------------------------------------------------------------------------
// camera — reference: src/engine/camera.h:8-36. get_ray runs on the device (k_raygen).
// ------------------------------------------------------------------------------------------------
class camera {
public:
    camera(point3 lookfrom, point3 lookat, vec3 vup, double vfov, double aspect_ratio, double aperture,
           double focus_dist, double _time0 = 0, double _time1 = 0);
    point3 origin, lower_left_corner;
    vec3 horizontal, vertical;
    vec3 u, v, w;
    double lens_radius;
    double time0, time1;
};

} // namespace artb

#endif
