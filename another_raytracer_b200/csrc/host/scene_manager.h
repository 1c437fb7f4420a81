// Scene builders of the host side.
//
// Same entry point as the reference: scene_manager::build(scene_alias) -> scene
// (reference: src/scene_manager.h:6-44, src/scene_manager.cpp:13-355). Aliases 1..9 are the reference's
// scenes, rebuilt here with the same object order and the same draws from the global RNG, so the
// resulting graphs are identical to the reference's down to the last bit (tests/test_scene_parity.py).
// Added aliases (BASELINE.json configs 4 and 5):
//   cow = 10, dino = 11   the reference's _mesh_scene assembly (scene_manager.cpp:236-258) around
//                         models/cow.obj / models/dino.obj with the camera presets the reference keeps as
//                         comments (scene_manager.cpp:337-342)
//   heightfield = 12      synthetic closed-form 708x708-vertex height field = 999 698 triangles,
//                         one triangle_soup, a light and a sky background
#ifndef ARTB_HOST_SCENE_MANAGER_H
#define ARTB_HOST_SCENE_MANAGER_H

#include "scene.h"

#include <string>

namespace artb {

struct scene {
    point3 lookfrom;
    point3 lookat;
    double vfov = 40.;
    double aperture = 0.;
    color background{0, 0, 0};
    hittable_list objects;
};

enum class scene_alias {
    random = 1,
    two_spheres = 2,
    two_perlin_spheres = 3,
    earth = 4,
    simple_light = 5,
    cornell_box = 6,
    cornell_smoke = 7,
    final = 8,
    mesh = 9,
    cow = 10,
    dino = 11,
    heightfield = 12
};

// asset root: $ARTB_ASSETS or <repo>/assets (set_assets_root overrides both)
std::string assets_root();
void set_assets_root(const std::string& path);

class scene_manager {
public:
    scene build(scene_alias alias);

    // NEW: options the reference hard-codes
    bool mesh_as_soup = false;        // cow/dino/mesh: emit one triangle_soup instead of per-triangle objects
    int heightfield_side = 708;       // vertices per side of alias 12
private:
    hittable_list _random_scene();
    hittable_list _two_spheres();
    hittable_list _two_perlin_spheres();
    hittable_list _earth();
    hittable_list _simple_light();
    hittable_list _cornell_box();
    hittable_list _cornell_smoke();
    hittable_list _final_scene();
    hittable_list _mesh_scene(const std::string& obj_path);
    hittable_list _heightfield_scene();
};

} // namespace artb

#endif
