// Host scene-graph construction: bounding boxes, the reference-shaped BVH build, Perlin tables, camera.
// Everything here must produce bit-identical doubles and consume the global RNG exactly as the
// reference does, because later scene content (and the oracle comparison) depends on it.
#include "scene.h"
#include "imageio.h"

#include <iostream>
#include <stdexcept>

namespace artb {

rng_state& global_rng()
{
    static rng_state g;
    return g;
}

void reset_global_rng(std::uint32_t seed) { global_rng().reseed(seed); }

// ---- textures ----------------------------------------------------------------------------------
perlin::perlin(const double* ranvec_xyz, const int* px, const int* py, const int* pz)
{
    for (int i = 0; i < point_count; ++i) {
        ranvec[size_t(i)] = vec3(ranvec_xyz[3 * i], ranvec_xyz[3 * i + 1], ranvec_xyz[3 * i + 2]);
        perm_x[size_t(i)] = px[i]; perm_y[size_t(i)] = py[i]; perm_z[size_t(i)] = pz[i];
    }
}

perlin::perlin()
{
    // reference: perlin.h:10-19 — gradients first, then the x, y, z permutations
    for (auto& v : ranvec) v = unit_vector(vec3::random(-1, 1));
    generate_perm(perm_x.data());
    generate_perm(perm_y.data());
    generate_perm(perm_z.data());
}

void perlin::generate_perm(int* p)
{
    // reference: perlin.h:63-77 — identity then Fisher-Yates from the top with random_int(0,i)
    for (int i = 0; i < point_count; ++i) p[i] = i;
    for (int i = point_count - 1; i > 0; --i) {
        const int target = random_int(0, i);
        std::swap(p[i], p[target]);
    }
}

image_texture::image_texture(const std::string& filename)
{
    // reference: texture.h:70-85
    if (!imageio::load_image(filename, data, width, height, bytes_per_pixel)) {
        std::cerr << "ERROR: Could not load texture image file '" << filename << "'.\n";
        data.clear();
        width = height = bytes_per_pixel = 0;
    }
}

// ---- hittables ---------------------------------------------------------------------------------
bool hittable_list::bounding_box(double time0, double time1, aabb& output_box) const
{
    // reference: hittable_list.cpp:21-34
    if (objects.empty()) return false;
    aabb temp_box;
    bool first_box = true;
    for (const auto& object : objects) {
        if (!object->bounding_box(time0, time1, temp_box)) return false;
        output_box = first_box ? temp_box : surrounding_box(output_box, temp_box);
        first_box = false;
    }
    return true;
}

namespace {
bool box_less(const std::shared_ptr<hittable>& a, const std::shared_ptr<hittable>& b, int axis)
{
    // reference: bvh.h:28-36 — compares box minima at time (0,0)
    aabb box_a, box_b;
    if (!a->bounding_box(0, 0, box_a) || !b->bounding_box(0, 0, box_b))
        std::cerr << "No bounding box in bvh_node constructor.\n";
    return box_a.min().e[axis] < box_b.min().e[axis];
}
} // namespace

bvh_node::bvh_node(const hittable_list& list, double time0, double time1)
{
    std::vector<std::shared_ptr<hittable>> objects = list.objects;   // the one modifiable copy
    init(objects, 0, objects.size(), time0, time1);
}

bvh_node::bvh_node(std::vector<std::shared_ptr<hittable>>& objects, size_t start, size_t end, double time0, double time1)
{
    init(objects, start, end, time0, time1);
}

void bvh_node::init(std::vector<std::shared_ptr<hittable>>& objects, size_t start, size_t end, double time0, double time1)
{
    // reference: bvh.cpp:3-42
    const int axis = random_int(0, 2);
    auto comparator = [axis](const std::shared_ptr<hittable>& a, const std::shared_ptr<hittable>& b) {
        return box_less(a, b, axis);
    };
    const size_t object_span = end - start;

    if (object_span == 1) {
        left = right = objects[start];
    } else if (object_span == 2) {
        if (comparator(objects[start], objects[start + 1])) {
            left = objects[start];
            right = objects[start + 1];
        } else {
            left = objects[start + 1];
            right = objects[start];
        }
    } else {
        std::sort(objects.begin() + std::ptrdiff_t(start), objects.begin() + std::ptrdiff_t(end), comparator);
        const size_t mid = start + object_span / 2;
        left = std::make_shared<bvh_node>(objects, start, mid, time0, time1);
        right = std::make_shared<bvh_node>(objects, mid, end, time0, time1);
    }

    aabb box_left, box_right;
    if (!left->bounding_box(time0, time1, box_left) || !right->bounding_box(time0, time1, box_right))
        std::cerr << "No bounding box in bvh_node constructor.\n";
    box = surrounding_box(box_left, box_right);
}

bool translate::bounding_box(double time0, double time1, aabb& output_box) const
{
    if (!ptr->bounding_box(time0, time1, output_box)) return false;
    output_box = aabb(output_box.min() + offset, output_box.max() + offset);
    return true;
}

rotate_y::rotate_y(std::shared_ptr<hittable> p, double angle) : ptr(std::move(p))
{
    // reference: hittable.cpp:25-55 — box of the 8 rotated corners of the child's box over times (0,1)
    const double radians = degrees_to_radians(angle);
    sin_theta = std::sin(radians);
    cos_theta = std::cos(radians);
    hasbox = ptr->bounding_box(0, 1, bbox);

    point3 lo(infinity, infinity, infinity), hi(-infinity, -infinity, -infinity);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) {
                const double x = i * bbox.max().x() + (1 - i) * bbox.min().x();
                const double y = j * bbox.max().y() + (1 - j) * bbox.min().y();
                const double z = k * bbox.max().z() + (1 - k) * bbox.min().z();
                const double newx = cos_theta * x + sin_theta * z;
                const double newz = -sin_theta * x + cos_theta * z;
                const vec3 tester(newx, y, newz);
                for (int c = 0; c < 3; ++c) {
                    lo[c] = std::fmin(lo[c], tester[c]);
                    hi[c] = std::fmax(hi[c], tester[c]);
                }
            }
    bbox = aabb(lo, hi);
}

bool sphere::bounding_box(double, double, aabb& output_box) const
{
    output_box = aabb(center - vec3(radius, radius, radius), center + vec3(radius, radius, radius));
    return true;
}

point3 moving_sphere::center(double time) const
{
    // reference: moving_sphere.h:29-31
    return center0 + ((time - time0) / (time1 - time0)) * (center1 - center0);
}

bool moving_sphere::bounding_box(double _time0, double _time1, aabb& output_box) const
{
    const vec3 r(radius, radius, radius);
    const aabb box0(center(_time0) - r, center(_time0) + r);
    const aabb box1(center(_time1) - r, center(_time1) + r);
    output_box = surrounding_box(box0, box1);
    return true;
}

box::box(const point3& p0, const point3& p1, std::shared_ptr<material> ptr) : box_min(p0), box_max(p1)
{
    sides.add(std::make_shared<xy_rect>(p0.x(), p1.x(), p0.y(), p1.y(), p1.z(), ptr));
    sides.add(std::make_shared<xy_rect>(p0.x(), p1.x(), p0.y(), p1.y(), p0.z(), ptr));
    sides.add(std::make_shared<xz_rect>(p0.x(), p1.x(), p0.z(), p1.z(), p1.y(), ptr));
    sides.add(std::make_shared<xz_rect>(p0.x(), p1.x(), p0.z(), p1.z(), p0.y(), ptr));
    sides.add(std::make_shared<yz_rect>(p0.y(), p1.y(), p0.z(), p1.z(), p1.x(), ptr));
    sides.add(std::make_shared<yz_rect>(p0.y(), p1.y(), p0.z(), p1.z(), p0.x(), ptr));
}

bool triangle_soup::bounding_box(double, double, aabb& b) const
{
    if (faces.empty()) return false;
    point3 lo(infinity, infinity, infinity), hi(-infinity, -infinity, -infinity);
    for (const auto& f : faces)
        for (int k = 0; k < 3; ++k) {
            lo = min(lo, vertices[size_t(f[k])]);
            hi = max(hi, vertices[size_t(f[k])]);
        }
    b = aabb(lo, hi);
    return true;
}

// ---- camera ------------------------------------------------------------------------------------
camera::camera(point3 lookfrom, point3 lookat, vec3 vup, double vfov, double aspect_ratio, double aperture,
               double focus_dist, double _time0, double _time1)
{
    // reference: camera.h:19-35
    const double theta = degrees_to_radians(vfov);
    const double h = std::tan(theta / 2);
    const double viewport_height = 2.0 * h;
    const double viewport_width = aspect_ratio * viewport_height;

    w = unit_vector(lookfrom - lookat);
    u = unit_vector(cross(vup, w));
    v = cross(w, u);

    origin = lookfrom;
    horizontal = focus_dist * viewport_width * u;
    vertical = focus_dist * viewport_height * v;
    lower_left_corner = origin - horizontal / 2 - vertical / 2 - focus_dist * w;

    lens_radius = aperture / 2;
    time0 = _time0;
    time1 = _time1;
}

} // namespace artb
