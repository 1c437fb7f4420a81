// TEST INFRASTRUCTURE (oracle). Scene/camera construction shared by ref_tool.cpp and ref_engine.cpp.
// Must be included AFTER the reference's scene_manager.cpp in the same TU.
#ifndef ARTB_REF_SCENES_H
#define ARTB_REF_SCENES_H

#include "camera.h"

namespace {

// ------------------------------------------------------------------------------------------------
// scene construction (aliases 10/11 mirror _mesh_scene with another .obj)
// ------------------------------------------------------------------------------------------------
hittable_list build_obj_scene(const std::string& path, bool all_metal = false)
{
    mesh m;
    if (!m.parse(path)) throw std::logic_error("cannot parse input obj file!");
    hittable_list world;
    auto triangles = m.build();
    if (all_metal) {      // aliases 13/14: BASELINE configs[3] "under Lambertian/metal" — the reference's triangles, one metal
        auto steel = std::make_shared<metal>(color(0.8, 0.8, 0.9), 0.1);
        for (auto& o : triangles.objects) std::static_pointer_cast<triangle>(o)->mat_ptr = steel;
    }
    world.add(std::make_shared<bvh_node>(triangles, 0.0, 1.0));
    auto light = std::make_shared<diffuse_light>(color(7, 7, 7));
    world.add(std::make_shared<xz_rect>(123, 423, 147, 412, 554, light));
    auto boundary = std::make_shared<sphere>(point3(0, 0, 0), 5000, std::make_shared<dielectric>(1.5));
    world.add(std::make_shared<constant_medium>(boundary, .0001, color(1,1,1)));
    return world;
}

scene build_scene(int alias)
{
    if (alias >= 1 && alias <= 9) {
        scene_manager mgr;
        return mgr.build(static_cast<scene_alias>(alias));
    }
    scene world;
    world.background = color(0.70, 0.80, 1.00);
    world.vfov = 75.0;
    if (alias == 10) {
        world.objects = build_obj_scene(ressources::cow_obj_path);
        world.lookfrom = point3(4,2,6);
        world.lookat = point3(2,0,0);
    } else if (alias == 11 || alias == 14) {
        world.objects = build_obj_scene(ressources::dino_obj_path, alias == 14);
        world.lookfrom = point3(0,15,25);
        world.lookat = point3(0,10,0);
    } else if (alias == 13) {
        world.objects = build_obj_scene(ressources::cow_obj_path, true);
        world.lookfrom = point3(4,2,6);
        world.lookat = point3(2,0,0);
    } else {
        throw std::logic_error("unknown scene alias");
    }
    return world;
}

camera make_camera(const scene& s, int W, int H)
{
    // src/main.cpp:33-35
    vec3 vup(0,1,0);
    auto dist_to_focus = 10.0;
    return camera(s.lookfrom, s.lookat, vup, s.vfov, double(W) / double(H), s.aperture, dist_to_focus, 0.0, 1.0);
}

} // namespace

#endif
