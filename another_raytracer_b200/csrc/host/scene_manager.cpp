// Scene builders. Object order and every draw from the global RNG follow the reference
// (src/scene_manager.cpp) so that graphs — and therefore images — can be compared one to one.
// Where the reference relies on GCC's right-to-left evaluation of constructor arguments, the order is
// written out explicitly (see core.h, vec3::random).
#include "scene_manager.h"
#include "mesh.h"

#include <cstdlib>
#include <stdexcept>

namespace artb {

namespace {
std::string g_assets_root;

template <class T, class... A>
std::shared_ptr<T> mk(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }
} // namespace

void set_assets_root(const std::string& path) { g_assets_root = path; }

std::string assets_root()
{
    if (!g_assets_root.empty()) return g_assets_root;
    if (const char* e = std::getenv("ARTB_ASSETS")) return e;
    return "/root/repo/assets";
}

// reference: scene_manager.cpp:13-64
static hittable_list build_random_spheres()
{
    hittable_list objects;

    auto ground = mk<checker_texture>(color(0.2, 0.3, 0.1), color(0.9, 0.9, 0.9));
    objects.add(mk<sphere>(point3(0, -1000, 0), 1000, mk<lambertian>(ground)));

    for (int a = -11; a < 11; a++) {
        for (int b = -11; b < 11; b++) {
            const double choose_mat = random_double();
            const double jz = random_double();      // z is drawn before x (right-to-left argument evaluation)
            const double jx = random_double();
            const point3 center(a + 0.9 * jx, 0.2, b + 0.9 * jz);

            if ((center - point3(4, 0.2, 0)).length() > 0.9) {
                if (choose_mat < 0.8) {
                    // diffuse: a static AND a moving sphere at the same spot
                    const color albedo = color::random() * color::random();
                    auto m = mk<lambertian>(albedo);
                    objects.add(mk<sphere>(center, 0.2, m));
                    const point3 center2 = center + vec3(0, random_double(0, .5), 0);
                    objects.add(mk<moving_sphere>(center, center2, 0.0, 1.0, 0.2, m));
                } else if (choose_mat < 0.95) {
                    const color albedo = color::random(0.5, 1);
                    const double fuzz = random_double(0, 0.5);
                    objects.add(mk<sphere>(center, 0.2, mk<metal>(albedo, fuzz)));
                } else {
                    objects.add(mk<sphere>(center, 0.2, mk<dielectric>(1.5)));
                }
            }
        }
    }

    objects.add(mk<sphere>(point3(0, 1, 0), 1.0, mk<dielectric>(1.5)));
    objects.add(mk<sphere>(point3(-4, 1, 0), 1.0, mk<lambertian>(color(0.4, 0.2, 0.1))));
    objects.add(mk<sphere>(point3(4, 1, 0), 1.0, mk<metal>(color(0.7, 0.6, 0.5), 0.0)));

    hittable_list world;
    world.add(mk<bvh_node>(objects, 0, 1));
    return world;
}

// reference: scene_manager.cpp:66-76
static hittable_list build_two_spheres()
{
    hittable_list objects;
    auto checker = mk<checker_texture>(color(0.2, 0.3, 0.1), color(0.9, 0.9, 0.9));
    objects.add(mk<sphere>(point3(0, -10, 0), 10, mk<lambertian>(checker)));
    objects.add(mk<sphere>(point3(0, 10, 0), 10, mk<lambertian>(checker)));
    return objects;
}

// reference: scene_manager.cpp:78-87
static hittable_list build_two_perlin_spheres()
{
    hittable_list objects;
    auto pertext = mk<noise_texture>(4);
    objects.add(mk<sphere>(point3(0, -1000, 0), 1000, mk<lambertian>(pertext)));
    objects.add(mk<sphere>(point3(0, 2, 0), 2, mk<lambertian>(pertext)));
    return objects;
}

// reference: scene_manager.cpp:89-96
static hittable_list build_earth()
{
    auto earth_texture = mk<image_texture>(assets_root() + "/textures/earthmap.png");
    auto globe = mk<sphere>(point3(0, 0, 0), 2, mk<lambertian>(earth_texture));
    return hittable_list{globe};
}

// reference: scene_manager.cpp:98-110
static hittable_list build_simple_light()
{
    hittable_list objects;
    auto pertext = mk<noise_texture>(4);
    objects.add(mk<sphere>(point3(0, -1000, 0), 1000, mk<lambertian>(pertext)));
    objects.add(mk<sphere>(point3(0, 2, 0), 2, mk<lambertian>(pertext)));
    objects.add(mk<xy_rect>(3, 5, 1, 3, -2, mk<diffuse_light>(color(4, 4, 4))));
    return objects;
}

namespace {
struct cornell_walls { std::shared_ptr<material> red, white, green; };

std::shared_ptr<hittable> placed_box(const point3& extent, double degrees, const vec3& where, std::shared_ptr<material> m)
{
    std::shared_ptr<hittable> b = mk<box>(point3(0, 0, 0), extent, std::move(m));
    b = mk<rotate_y>(b, degrees);
    return mk<translate>(b, where);
}
} // namespace

// reference: scene_manager.cpp:112-139
static hittable_list build_cornell_box()
{
    hittable_list objects;
    auto red = mk<lambertian>(color(.65, .05, .05));
    auto white = mk<lambertian>(color(.73, .73, .73));
    auto green = mk<lambertian>(color(.12, .45, .15));
    auto light = mk<diffuse_light>(color(15, 15, 15));

    objects.add(mk<yz_rect>(0, 555, 0, 555, 555, green));
    objects.add(mk<yz_rect>(0, 555, 0, 555, 0, red));
    objects.add(mk<xz_rect>(213, 343, 227, 332, 554, light));
    objects.add(mk<xz_rect>(0, 555, 0, 555, 0, white));
    objects.add(mk<xz_rect>(0, 555, 0, 555, 555, white));
    objects.add(mk<xy_rect>(0, 555, 0, 555, 555, white));

    objects.add(placed_box(point3(165, 330, 165), 15, vec3(265, 0, 295), white));
    objects.add(placed_box(point3(165, 165, 165), -18, vec3(130, 0, 65), white));
    return objects;
}

// reference: scene_manager.cpp:141-169 (note the swapped floor/ceiling order w.r.t. _cornell_box)
static hittable_list build_cornell_smoke()
{
    hittable_list objects;
    auto red = mk<lambertian>(color(.65, .05, .05));
    auto white = mk<lambertian>(color(.73, .73, .73));
    auto green = mk<lambertian>(color(.12, .45, .15));
    auto light = mk<diffuse_light>(color(7, 7, 7));

    objects.add(mk<yz_rect>(0, 555, 0, 555, 555, green));
    objects.add(mk<yz_rect>(0, 555, 0, 555, 0, red));
    objects.add(mk<xz_rect>(113, 443, 127, 432, 554, light));
    objects.add(mk<xz_rect>(0, 555, 0, 555, 555, white));
    objects.add(mk<xz_rect>(0, 555, 0, 555, 0, white));
    objects.add(mk<xy_rect>(0, 555, 0, 555, 555, white));

    auto box1 = placed_box(point3(165, 330, 165), 15, vec3(265, 0, 295), white);
    auto box2 = placed_box(point3(165, 165, 165), -18, vec3(130, 0, 65), white);
    objects.add(mk<constant_medium>(box1, 0.01, color(0, 0, 0)));
    objects.add(mk<constant_medium>(box2, 0.01, color(1, 1, 1)));
    return objects;
}

// reference: scene_manager.cpp:171-234
static hittable_list build_next_week_final()
{
    hittable_list boxes1;
    auto ground = mk<lambertian>(color(0.48, 0.83, 0.53));

    const int boxes_per_side = 20;
    for (int i = 0; i < boxes_per_side; i++) {
        for (int j = 0; j < boxes_per_side; j++) {
            const double w = 100.0;
            const double x0 = -1000.0 + i * w;
            const double z0 = -1000.0 + j * w;
            const double y0 = 0.0;
            const double x1 = x0 + w;
            const double y1 = random_double(1, 101);
            const double z1 = z0 + w;
            boxes1.add(mk<box>(point3(x0, y0, z0), point3(x1, y1, z1), ground));
        }
    }

    hittable_list objects;
    objects.add(mk<bvh_node>(boxes1, 0, 1));

    objects.add(mk<xz_rect>(123, 423, 147, 412, 554, mk<diffuse_light>(color(7, 7, 7))));

    const point3 center1(400, 400, 200);
    const point3 center2 = center1 + vec3(30, 0, 0);
    objects.add(mk<moving_sphere>(center1, center2, 0, 1, 50, mk<lambertian>(color(0.7, 0.3, 0.1))));

    objects.add(mk<sphere>(point3(260, 150, 45), 50, mk<dielectric>(1.5)));
    objects.add(mk<sphere>(point3(0, 150, 145), 50, mk<metal>(color(0.8, 0.8, 0.9), 1.0)));

    // glass ball filled with blue smoke: the same sphere object is a surface AND a medium boundary
    auto boundary = mk<sphere>(point3(360, 150, 145), 70, mk<dielectric>(1.5));
    objects.add(boundary);
    objects.add(mk<constant_medium>(boundary, 0.2, color(0.2, 0.4, 0.9)));
    // thin mist over everything
    boundary = mk<sphere>(point3(0, 0, 0), 5000, mk<dielectric>(1.5));
    objects.add(mk<constant_medium>(boundary, .0001, color(1, 1, 1)));

    auto emat = mk<lambertian>(mk<image_texture>(assets_root() + "/textures/earthmap.png"));
    objects.add(mk<sphere>(point3(400, 200, 400), 100, emat));
    auto pertext = mk<noise_texture>(0.1);
    objects.add(mk<sphere>(point3(220, 280, 300), 80, mk<lambertian>(pertext)));

    hittable_list boxes2;
    auto white = mk<lambertian>(color(.73, .73, .73));
    const int ns = 1000;
    for (int j = 0; j < ns; j++)
        boxes2.add(mk<sphere>(point3::random(0, 165), 10, white));

    objects.add(mk<translate>(mk<rotate_y>(mk<bvh_node>(boxes2, 0.0, 1.0), 15), vec3(-100, 270, 395)));
    return objects;
}

// reference: scene_manager.cpp:236-258. all_metal: BASELINE configs[3]'s metal variant — the mesh is built exactly as the
// reference builds it (per-triangle lambertians, random colours drawn), then every triangle gets the one metal.
static hittable_list build_mesh_scene(const std::string& obj_path, bool mesh_as_soup, bool all_metal = false)
{
    mesh m;
    if (!m.parse(obj_path)) throw std::logic_error("cannot parse input obj file!");

    const auto steel = all_metal ? mk<metal>(color(0.8, 0.8, 0.9), 0.1) : std::shared_ptr<metal>();
    hittable_list world;
    if (mesh_as_soup) {
        auto soup = m.build_soup();
        if (all_metal) soup->materials.assign(1, steel);
        world.add(soup);
    } else {
        hittable_list triangles = m.build();
        if (all_metal)
            for (auto& o : triangles.objects) static_cast<triangle&>(*o).mat_ptr = steel;
        world.add(mk<bvh_node>(triangles, 0.0, 1.0));
    }
    world.add(mk<xz_rect>(123, 423, 147, 412, 554, mk<diffuse_light>(color(7, 7, 7))));
    auto boundary = mk<sphere>(point3(0, 0, 0), 5000, mk<dielectric>(1.5));
    world.add(mk<constant_medium>(boundary, .0001, color(1, 1, 1)));
    return world;
}

// NEW — BASELINE.json config 5: a closed-form height field, no RNG involved.
// side x side vertices on [-12,12]^2, y = sum of three sine products; 2*(side-1)^2 triangles.
static hittable_list build_heightfield(int heightfield_side)
{
    const int n = heightfield_side;
    if (n < 2) throw std::logic_error("heightfield_side must be >= 2");
    auto soup = mk<triangle_soup>();
    soup->vertices.reserve(size_t(n) * size_t(n));
    const double extent = 12.0;
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < n; ++i) {
            const double x = -extent + 2 * extent * i / (n - 1);
            const double z = -extent + 2 * extent * j / (n - 1);
            const double y = 0.9 * std::sin(0.7 * x) * std::cos(0.5 * z)
                           + 0.35 * std::sin(2.3 * x + 0.4) * std::sin(1.9 * z)
                           + 0.08 * std::sin(7.1 * x) * std::cos(6.3 * z + 1.0);
            soup->vertices.emplace_back(x, y, z);
        }
    }
    soup->faces.reserve(size_t(2) * size_t(n - 1) * size_t(n - 1));
    for (int j = 0; j + 1 < n; ++j) {
        for (int i = 0; i + 1 < n; ++i) {
            const int v00 = j * n + i, v10 = v00 + 1, v01 = v00 + n, v11 = v01 + 1;
            soup->faces.push_back({ v00, v01, v10 });     // wound so that the geometric normal points up (+y)
            soup->faces.push_back({ v10, v01, v11 });
        }
    }
    soup->materials.push_back(mk<lambertian>(color(0.55, 0.6, 0.45)));

    hittable_list world;
    world.add(soup);
    world.add(mk<xz_rect>(-4, 4, -4, 4, 9, mk<diffuse_light>(color(6, 6, 5))));
    world.add(mk<sphere>(point3(0, 1.6, 0), 1.0, mk<metal>(color(0.8, 0.8, 0.9), 0.05)));
    return world;
}

// reference: scene_manager.cpp:260-355
scene scene_manager::build(scene_alias alias)
{
    scene world;
    const color sky(0.70, 0.80, 1.00);

    switch (alias) {
        case scene_alias::random:
            world.objects = build_random_spheres();
            world.background = sky;
            world.lookfrom = point3(13, 2, 3);
            world.lookat = point3(0, 0, 0);
            world.vfov = 20.0;
            world.aperture = 0.1;
            break;
        case scene_alias::two_spheres:
            world.objects = build_two_spheres();
            world.background = sky;
            world.lookfrom = point3(13, 2, 3);
            world.lookat = point3(0, 0, 0);
            world.vfov = 20.0;
            break;
        case scene_alias::two_perlin_spheres:
            world.objects = build_two_perlin_spheres();
            world.background = sky;
            world.lookfrom = point3(13, 2, 3);
            world.lookat = point3(0, 0, 0);
            world.vfov = 20.0;
            break;
        case scene_alias::earth:
            world.objects = build_earth();
            world.background = sky;
            world.lookfrom = point3(13, 2, 3);
            world.lookat = point3(0, 0, 0);
            world.vfov = 20.0;
            break;
        case scene_alias::simple_light:
            world.objects = build_simple_light();
            world.background = color(0, 0, 0);
            world.lookfrom = point3(26, 3, 6);
            world.lookat = point3(0, 2, 0);
            world.vfov = 20.0;
            break;
        case scene_alias::cornell_box:
            world.objects = build_cornell_box();
            world.background = color(0, 0, 0);
            world.lookfrom = point3(278, 278, -800);
            world.lookat = point3(278, 278, 0);
            world.vfov = 40.0;
            break;
        case scene_alias::cornell_smoke:
            world.objects = build_cornell_smoke();
            world.background = color(0, 0, 0);
            world.lookfrom = point3(278, 278, -800);
            world.lookat = point3(278, 278, 0);
            world.vfov = 40.0;
            break;
        case scene_alias::final:
            world.objects = build_next_week_final();
            world.background = color(0, 0, 0);
            world.lookfrom = point3(478, 278, -600);
            world.lookat = point3(278, 278, 0);
            world.vfov = 40.0;
            break;
        case scene_alias::mesh:
            world.objects = build_mesh_scene(assets_root() + "/models/capsule/capsule.obj", mesh_as_soup);
            world.background = sky;
            world.lookfrom = point3(2, 2, 1);
            world.lookat = point3(0, 0, 0);
            world.vfov = 75.0;
            break;
        case scene_alias::cow:
            world.objects = build_mesh_scene(assets_root() + "/models/cow.obj", mesh_as_soup);
            world.background = sky;
            world.lookfrom = point3(4, 2, 6);
            world.lookat = point3(2, 0, 0);
            world.vfov = 75.0;
            break;
        case scene_alias::dino:
            world.objects = build_mesh_scene(assets_root() + "/models/dino.obj", mesh_as_soup);
            world.background = sky;
            world.lookfrom = point3(0, 15, 25);
            world.lookat = point3(0, 10, 0);
            world.vfov = 75.0;
            break;
        case scene_alias::cow_metal:
            world.objects = build_mesh_scene(assets_root() + "/models/cow.obj", mesh_as_soup, true);
            world.background = sky;
            world.lookfrom = point3(4, 2, 6);
            world.lookat = point3(2, 0, 0);
            world.vfov = 75.0;
            break;
        case scene_alias::dino_metal:
            world.objects = build_mesh_scene(assets_root() + "/models/dino.obj", mesh_as_soup, true);
            world.background = sky;
            world.lookfrom = point3(0, 15, 25);
            world.lookat = point3(0, 10, 0);
            world.vfov = 75.0;
            break;
        case scene_alias::heightfield:
            world.objects = build_heightfield(heightfield_side);
            world.background = sky;
            world.lookfrom = point3(0, 7, 19);
            world.lookat = point3(0, 0.3, 0);
            world.vfov = 40.0;
            break;
        default:
            throw std::logic_error("unkwnown scene requested!");
    }
    return world;
}

} // namespace artb
