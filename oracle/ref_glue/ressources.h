// TEST INFRASTRUCTURE (oracle). Stand-in for the header CMake generates from the reference's
// src/ressources.h.in:4-17 (configure_file). Same namespace and variable names; the asset root is this
// repository's assets/ directory (override with ARTB_ASSETS). The JPEG textures of the reference are
// shipped here as PNGs holding the texels exactly as the reference's own decoder (stb_image) produced
// them, see tools/make_assets.py.
#ifndef RESSOURCE_H
#define RESSOURCE_H

#include <cstdlib>
#include <string>

namespace ressources {

inline std::string assets_root()
{
    const char* e = std::getenv("ARTB_ASSETS");
    return std::string(e ? e : "/root/repo/assets");
}

const std::string models_path = assets_root() + "/models/";

const std::string cow_obj_path = models_path + "cow.obj";
const std::string dino_obj_path = models_path + "dino.obj";
const std::string capsule_obj_path = models_path + "capsule/capsule.obj";

const std::string textures_path = assets_root() + "/textures/";

#ifdef ARTB_REF_ORIGINAL_JPEG
const std::string earthmap_texture = textures_path + "earthmap.jpg";
#else
const std::string earthmap_texture = textures_path + "earthmap.png";
#endif

} // namespace ressources

#endif
