// Scene builders of the host side.
//
// Same entry point as the reference: scene_manager::build(scene_alias) -> scene
// (reference: src/scene_manager.h:6-44, src/scene_manager.cpp:13-355). Aliases 1..9 are the reference's
// scenes, rebuilt here with the same object order and the same draws from the global RNG, so the
// resulting graphs are identical to the reference's down to the last bit (tests/test_host_scene.py).
// Added aliases (BASELINE.json configs 4 and 5):
//   cow = 10, dino = 11   the reference's _mesh_scene assembly (scene_manager.cpp:236-258) around
//                         models/cow.obj / models/dino.obj with the camera presets the reference keeps as
//                         comments (scene_manager.cpp:337-342)
//   cow_metal = 13, dino_metal = 14   the same two scenes, every triangle under metal(0.8, 0.8, 0.9; fuzz 0.1): the
//                         per-triangle lambertians are still created (and draw their random colours) before being
//                         replaced, so the global RNG advances exactly as in aliases 10 / 11
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
    heightfield = 12,
    cow_metal = 13,       // BASELINE.json configs[3] "under Lambertian/metal": the same meshes with every triangle's
    dino_metal = 14       // material replaced by one metal (unnormalised triangle normals reach metal::scatter)
};

// asset root: $ARTB_ASSETS or <repo>/assets (set_assets_root overrides both)
std::string assets_root();
void set_assets_root(const std::string& path);

// scene_manager::build is the reference's entry point (scene_manager.h:29-44); the per-scene builders are file-local
// functions of scene_manager.cpp.
class scene_manager {
public:
    // options the reference hard-codes
    bool mesh_as_soup = false;        // aliases 9-11: emit one triangle_soup instead of per-triangle objects
    int heightfield_side = 708;       // alias 12: vertices per side

    scene build(scene_alias alias);
};

} // namespace artb

#endif
