// Wavefront .obj / .mtl loading for the host side.
//
// Stands in for the reference's mesh class (src/primitives/mesh.h:29-148), which delegates parsing and
// triangulation to the third-party rapidobj 1.0.1 (vendored upstream, not part of this repository).
// This file restates the parts of that pipeline that decide WHICH triangles come out and in WHAT order,
// because the order decides the per-triangle random colours (mesh.h:132-138) and the BVH:
//   * faces are emitted in file order; groups/objects do not reorder anything;
//   * a quad (0,1,2,3) is split along its shorter diagonal, compared in float:
//     |p0-p2|^2 < |p1-p3|^2 -> (0,1,2)(0,2,3) else (0,1,3)(1,2,3)       rapidobj.hpp Triangulate, quad branch
//   * a polygon with >4 vertices is projected on its dominant plane (largest float shoelace area) and
//     ear-clipped the way mapbox/earcut 2.2.4 does for a hole-free ring of <80 points, then each output
//     triangle has its first two indices swapped                         rapidobj.hpp Triangulate, n-gon branch
//   * vertex positions and texcoords are parsed as float and widened to double (mesh.h:77-90).
// tests/test_host_scene.py compares the resulting triangle streams of cow/dino/capsule with the
// reference's own.
#ifndef ARTB_HOST_MESH_H
#define ARTB_HOST_MESH_H

#include "scene.h"

#include <map>
#include <string>
#include <vector>

namespace artb {

struct obj_index { int position_index = -1, texcoord_index = -1, normal_index = -1; };

struct obj_material {
    std::string name;
    float ambient[3] = {0, 0, 0};
    float diffuse[3] = {0, 0, 0};
    std::string diffuse_texname;
};

struct obj_data {
    std::vector<float> positions;   // xyz
    std::vector<float> texcoords;   // uv
    std::vector<obj_index> indices; // 3 per triangle after triangulate()
    std::vector<int> material_ids;  // 1 per triangle
    std::vector<obj_material> materials;
    std::string error;
};

class mesh {
public:
    bool parse(const std::string& mesh_path);           // reference: mesh.h:31-65
    hittable_list build();                               // reference: mesh.h:67-145 (one triangle object per face)
    // NEW: whole mesh as a single triangle_soup — with one shared material, or with the per-face
    // materials build() would give the triangle objects (same colours, same map_Kd image textures, same RNG draws)
    std::shared_ptr<triangle_soup> build_soup(std::shared_ptr<material> mat) const;
    std::shared_ptr<triangle_soup> build_soup();
    size_t num_triangles() const { return parse_data.indices.size() / 3; }

    obj_data parse_data;
private:
    // material of face i as mesh.h:104-138 chooses it; `maps` caches the decoded map_Kd images (material_map_handler)
    std::shared_ptr<material> face_material(size_t i, std::map<std::string, std::shared_ptr<image_texture>>& maps);
    std::string model_work_path;
};

namespace detail {
// exposed for tests: ear clipping of one ring; returns vertex positions within the ring, 3 per triangle
std::vector<unsigned> earcut_ring(const std::vector<std::array<float, 2>>& ring);
} // namespace detail

} // namespace artb

#endif
