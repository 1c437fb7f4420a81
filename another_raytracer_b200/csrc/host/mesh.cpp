#include "mesh.h"

#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>

namespace artb {

namespace detail {

// ---- ear clipping, after mapbox/earcut.hpp v2.2.4 (as vendored inside rapidobj 1.0.1) -------------
// Only what a single ring of < 80 points needs: no holes, no z-order hashing, and the first two
// passes of the ear loop (plain, then after dropping collinear/duplicate points).
namespace {
struct node { unsigned i; double x, y; int prev, next; };

struct clipper {
    std::vector<node> n;
    std::vector<unsigned> out;

    double area(int p, int q, int r) const
    { return (n[q].y - n[p].y) * (n[r].x - n[q].x) - (n[q].x - n[p].x) * (n[r].y - n[q].y); }
    bool same(int a, int b) const { return n[a].x == n[b].x && n[a].y == n[b].y; }
    static bool in_tri(double ax, double ay, double bx, double by, double cx, double cy, double px, double py)
    {
        return (cx - px) * (ay - py) >= (ax - px) * (cy - py) &&
               (ax - px) * (by - py) >= (bx - px) * (ay - py) &&
               (bx - px) * (cy - py) >= (cx - px) * (by - py);
    }
    void unlink(int p) { n[n[p].next].prev = n[p].prev; n[n[p].prev].next = n[p].next; }
    int insert(unsigned i, double x, double y, int last)
    {
        const int p = int(n.size());
        n.push_back({ i, x, y, p, p });
        if (last >= 0) {
            n[p].next = n[last].next;
            n[p].prev = last;
            n[n[last].next].prev = p;
            n[last].next = p;
        }
        return p;
    }
    bool is_ear(int ear) const
    {
        const int a = n[ear].prev, b = ear, c = n[ear].next;
        if (area(a, b, c) >= 0) return false;
        for (int p = n[n[ear].next].next; p != n[ear].prev; p = n[p].next)
            if (in_tri(n[a].x, n[a].y, n[b].x, n[b].y, n[c].x, n[c].y, n[p].x, n[p].y) &&
                area(n[p].prev, p, n[p].next) >= 0)
                return false;
        return true;
    }
    int filter(int start)
    {
        int end = start, p = start;
        bool again;
        do {
            again = false;
            if (same(p, n[p].next) || area(n[p].prev, p, n[p].next) == 0) {
                unlink(p);
                p = end = n[p].prev;
                if (p == n[p].next) break;
                again = true;
            } else {
                p = n[p].next;
            }
        } while (again || p != end);
        return end;
    }
    void clip(int ear, int pass)
    {
        int stop = ear;
        while (n[ear].prev != n[ear].next) {
            const int prev = n[ear].prev, next = n[ear].next;
            if (is_ear(ear)) {
                out.push_back(n[prev].i); out.push_back(n[ear].i); out.push_back(n[next].i);
                unlink(ear);
                ear = n[next].next;
                stop = n[next].next;
                continue;
            }
            ear = next;
            if (ear == stop) {
                if (pass == 0) clip(filter(ear), 1);
                else throw std::runtime_error("obj: polygon needs earcut passes this loader does not implement");
                break;
            }
        }
    }
};
} // namespace

std::vector<unsigned> earcut_ring(const std::vector<std::array<float, 2>>& ring)
{
    clipper c;
    const size_t len = ring.size();
    if (len < 3) return {};
    double sum = 0;
    for (size_t i = 0, j = len - 1; i < len; j = i++) {
        const double p10 = ring[i][0], p11 = ring[i][1], p20 = ring[j][0], p21 = ring[j][1];
        sum += (p20 - p10) * (p11 + p21);
    }
    int last = -1;
    if (sum > 0) { for (size_t i = 0; i < len; ++i) last = c.insert(unsigned(i), ring[i][0], ring[i][1], last); }
    else         { for (size_t i = len; i-- > 0;) last = c.insert(unsigned(i), ring[i][0], ring[i][1], last); }
    if (last >= 0 && c.same(last, c.n[last].next)) { c.unlink(last); last = c.n[last].next; }
    if (last < 0 || c.n[last].prev == c.n[last].next) return {};
    c.clip(last, 0);
    return c.out;
}

} // namespace detail

namespace {

float polygon_area(const float* x, const float* y, size_t size)
{
    float area = 0.0f;
    for (size_t i = 1; i != size; ++i) {
        const float avg_height = (y[i - 1] + y[i]) / 2;
        const float width = x[i] - x[i - 1];
        area += width * avg_height;
    }
    const float avg_height = (y[0] + y[size - 1]) / 2;
    const float width = x[0] - x[size - 1];
    area += width * avg_height;
    return std::abs(area);
}

const char* skip_ws(const char* p) { while (*p == ' ' || *p == '\t') ++p; return p; }

bool parse_mtl(const std::string& path, std::vector<obj_material>& out)
{
    std::ifstream f(path);
    if (!f) return false;
    std::string line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        const char* p = skip_ws(line.c_str());
        if (!std::strncmp(p, "newmtl", 6) && (p[6] == ' ' || p[6] == '\t')) {
            out.emplace_back();
            out.back().name = skip_ws(p + 6);
        } else if (out.empty()) {
            continue;
        } else if (!std::strncmp(p, "Ka", 2) && (p[2] == ' ' || p[2] == '\t')) {
            char* e = const_cast<char*>(p + 2);
            for (int k = 0; k < 3; ++k) out.back().ambient[k] = std::strtof(e, &e);
        } else if (!std::strncmp(p, "Kd", 2) && (p[2] == ' ' || p[2] == '\t')) {
            char* e = const_cast<char*>(p + 2);
            for (int k = 0; k < 3; ++k) out.back().diffuse[k] = std::strtof(e, &e);
        } else if (!std::strncmp(p, "map_Kd", 6) && (p[6] == ' ' || p[6] == '\t')) {
            std::string name = skip_ws(p + 6);
            while (!name.empty() && (name.back() == ' ' || name.back() == '\t')) name.pop_back();
            out.back().diffuse_texname = name;
        }
    }
    return true;
}

// one polygonal face -> triangles appended to data.indices / data.material_ids
void triangulate_face(obj_data& data, const std::vector<obj_index>& face, int material_id)
{
    const size_t nv = face.size();
    auto pos = [&](const obj_index& ix, int c) { return data.positions[size_t(3 * ix.position_index + c)]; };
    auto emit = [&](const obj_index& a, const obj_index& b, const obj_index& c) {
        data.indices.push_back(a); data.indices.push_back(b); data.indices.push_back(c);
        data.material_ids.push_back(material_id);
    };
    if (nv == 3) {
        emit(face[0], face[1], face[2]);
    } else if (nv == 4) {
        const float e02x = pos(face[0], 0) - pos(face[2], 0), e02y = pos(face[0], 1) - pos(face[2], 1), e02z = pos(face[0], 2) - pos(face[2], 2);
        const float e13x = pos(face[1], 0) - pos(face[3], 0), e13y = pos(face[1], 1) - pos(face[3], 1), e13z = pos(face[1], 2) - pos(face[3], 2);
        const float d02 = e02x * e02x + e02y * e02y + e02z * e02z;
        const float d13 = e13x * e13x + e13y * e13y + e13z * e13z;
        const bool d02_is_less = d02 < d13;
        emit(face[0], face[1], d02_is_less ? face[2] : face[3]);
        emit(d02_is_less ? face[0] : face[1], face[2], face[3]);
    } else {
        std::vector<float> xs(nv), ys(nv), zs(nv);
        for (size_t k = 0; k < nv; ++k) { xs[k] = pos(face[k], 0); ys[k] = pos(face[k], 1); zs[k] = pos(face[k], 2); }
        const float area_x = polygon_area(ys.data(), zs.data(), nv);
        const float area_y = polygon_area(xs.data(), zs.data(), nv);
        const float area_z = polygon_area(xs.data(), ys.data(), nv);
        if (FLT_MIN > std::max({ area_x, area_y, area_z })) throw std::runtime_error("obj: degenerate polygon");
        int plane;   // 0 = X, 1 = Y, 2 = Z
        if (area_x > area_y) plane = area_x > area_z ? 0 : 2;
        else                 plane = area_y > area_z ? 1 : 2;
        std::vector<std::array<float, 2>> ring(nv);
        for (size_t k = 0; k < nv; ++k) {
            if (plane == 0)      ring[k] = { ys[k], zs[k] };
            else if (plane == 1) ring[k] = { xs[k], zs[k] };
            else                 ring[k] = { xs[k], ys[k] };
        }
        std::vector<unsigned> tri = detail::earcut_ring(ring);
        if (tri.empty() || tri.size() % 3 != 0) throw std::runtime_error("obj: polygon triangulation failed");
        for (size_t k = 0; k < tri.size(); k += 3) std::swap(tri[k], tri[k + 1]);
        for (size_t k = 0; k < tri.size(); k += 3) emit(face[tri[k]], face[tri[k + 1]], face[tri[k + 2]]);
    }
}

} // namespace

bool mesh::parse(const std::string& mesh_path)
{
    parse_data = obj_data();
    std::ifstream f(mesh_path, std::ios::binary);
    if (!f) {
        parse_data.error = "cannot open file";
        std::cerr << "cannot parse mesh file (error : " << parse_data.error << ")" << std::endl;
        return false;
    }
    const size_t slash = mesh_path.find_last_of('/');
    model_work_path = slash == std::string::npos ? std::string() : mesh_path.substr(0, slash);

    std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::map<std::string, int> material_index;
    int current_material = -1;
    size_t num_normals = 0;
    std::vector<obj_index> face;

    try {
        const char* p = text.c_str();
        const char* end = p + text.size();
        size_t line_num = 0;
        while (p < end) {
            ++line_num;
            const char* eol = static_cast<const char*>(std::memchr(p, '\n', size_t(end - p)));
            if (!eol) eol = end;
            const char* q = skip_ws(p);
            if (q[0] == 'v' && (q[1] == ' ' || q[1] == '\t')) {
                char* e = const_cast<char*>(q + 1);
                for (int k = 0; k < 3; ++k) parse_data.positions.push_back(std::strtof(e, &e));
            } else if (q[0] == 'v' && q[1] == 't' && (q[2] == ' ' || q[2] == '\t')) {
                char* e = const_cast<char*>(q + 2);
                for (int k = 0; k < 2; ++k) parse_data.texcoords.push_back(std::strtof(e, &e));
            } else if (q[0] == 'v' && q[1] == 'n' && (q[2] == ' ' || q[2] == '\t')) {
                ++num_normals;
            } else if (q[0] == 'f' && (q[1] == ' ' || q[1] == '\t')) {
                face.clear();
                const char* s = q + 1;
                while (true) {
                    s = skip_ws(s);
                    if (s >= eol || *s == '\r' || *s == '\n' || *s == '#') break;
                    obj_index ix;
                    char* e = nullptr;
                    long v = std::strtol(s, &e, 10);
                    if (e == s) throw std::runtime_error("bad face element on line " + std::to_string(line_num));
                    ix.position_index = int(v > 0 ? v - 1 : long(parse_data.positions.size() / 3) + v);
                    s = e;
                    if (*s == '/') {
                        ++s;
                        if (*s != '/') {
                            v = std::strtol(s, &e, 10);
                            if (e != s) ix.texcoord_index = int(v > 0 ? v - 1 : long(parse_data.texcoords.size() / 2) + v);
                            s = e;
                        }
                        if (*s == '/') {
                            ++s;
                            v = std::strtol(s, &e, 10);
                            if (e != s) ix.normal_index = int(v > 0 ? v - 1 : long(num_normals) + v);
                            s = e;
                        }
                    }
                    if (ix.position_index < 0 || size_t(ix.position_index) >= parse_data.positions.size() / 3)
                        throw std::runtime_error("face index out of range on line " + std::to_string(line_num));
                    face.push_back(ix);
                }
                if (face.size() < 3) throw std::runtime_error("face with fewer than 3 vertices on line " + std::to_string(line_num));
                triangulate_face(parse_data, face, current_material);
            } else if (!std::strncmp(q, "usemtl", 6) && (q[6] == ' ' || q[6] == '\t')) {
                std::string name(skip_ws(q + 6), eol);
                while (!name.empty() && (name.back() == '\r' || name.back() == ' ' || name.back() == '\t')) name.pop_back();
                auto it = material_index.find(name);
                current_material = it == material_index.end() ? -1 : it->second;
            } else if (!std::strncmp(q, "mtllib", 6) && (q[6] == ' ' || q[6] == '\t')) {
                std::string name(skip_ws(q + 6), eol);
                while (!name.empty() && (name.back() == '\r' || name.back() == ' ' || name.back() == '\t')) name.pop_back();
                const size_t before = parse_data.materials.size();
                if (!parse_mtl(model_work_path.empty() ? name : model_work_path + "/" + name, parse_data.materials))
                    throw std::runtime_error("cannot open material library " + name);
                for (size_t i = before; i < parse_data.materials.size(); ++i)
                    material_index[parse_data.materials[i].name] = int(i);
            }
            p = eol + 1;
        }
    } catch (const std::exception& e) {
        parse_data.error = e.what();
        std::cerr << "cannot parse mesh file (error : " << parse_data.error << ")" << std::endl;
        return false;
    }
    return true;
}

std::shared_ptr<material> mesh::face_material(size_t i, std::map<std::string, std::shared_ptr<image_texture>>& material_maps)
{
    const std::vector<obj_index>& indices = parse_data.indices;
    auto texcoord = [this](int index) -> std::pair<double, double> {
        const std::vector<float>& t = parse_data.texcoords;
        return { t[size_t(2 * index) + 0], t[size_t(2 * index) + 1] };
    };
    if (parse_data.materials.empty()) return std::make_shared<lambertian>(color::random());      // mesh.h:132-138
    const int mid = parse_data.material_ids[i];
    if (mid < 0) throw std::logic_error("obj: face without material in a file that declares materials");
    const obj_material& m = parse_data.materials[size_t(mid)];
    if (!m.diffuse_texname.empty()) {
        auto it = material_maps.find(m.diffuse_texname);
        if (it == material_maps.end())
            it = material_maps.emplace(m.diffuse_texname,
                                       std::make_shared<image_texture>(model_work_path + "/" + m.diffuse_texname)).first;
        auto tex = std::make_shared<barycentric_image_texture>(
            texcoord(indices[3 * i + 0].texcoord_index),
            texcoord(indices[3 * i + 1].texcoord_index),
            texcoord(indices[3 * i + 2].texcoord_index), it->second);
        return std::make_shared<lambertian>(tex);
    }
    // reference: mesh.h:128 — float sums widened to double
    return std::make_shared<lambertian>(color(m.ambient[0] + m.diffuse[0], m.ambient[1] + m.diffuse[1], m.ambient[2] + m.diffuse[2]));
}

hittable_list mesh::build()
{
    hittable_list triangles;
    std::map<std::string, std::shared_ptr<image_texture>> material_maps;   // reference: material_map_handler, mesh.h:9-27

    auto vertex = [this](int index) {
        const std::vector<float>& p = parse_data.positions;
        return vec3(p[size_t(3 * index) + 0], p[size_t(3 * index) + 1], p[size_t(3 * index) + 2]);
    };

    const std::vector<obj_index>& indices = parse_data.indices;
    for (size_t i = 0; i < indices.size() / 3; ++i) {
        const vec3 a = vertex(indices[3 * i + 0].position_index);
        const vec3 b = vertex(indices[3 * i + 1].position_index);
        const vec3 c = vertex(indices[3 * i + 2].position_index);
        triangles.add(std::make_shared<triangle>(a, b, c, face_material(i, material_maps)));
    }
    return triangles;
}

// Direct path: the same triangles and the same per-face materials as build(), held as ONE object. Textured faces keep
// their barycentric_image_texture; the flattener puts every distinct map_Kd image into the device texel atlas once.
std::shared_ptr<triangle_soup> mesh::build_soup()
{
    std::map<std::string, std::shared_ptr<image_texture>> material_maps;
    auto soup = build_soup(std::shared_ptr<material>());
    soup->materials.clear();
    soup->materials.reserve(soup->faces.size());
    for (size_t i = 0; i < soup->faces.size(); ++i) soup->materials.push_back(face_material(i, material_maps));
    return soup;
}

std::shared_ptr<triangle_soup> mesh::build_soup(std::shared_ptr<material> mat) const
{
    auto soup = std::make_shared<triangle_soup>();
    const std::vector<float>& p = parse_data.positions;
    soup->vertices.reserve(p.size() / 3);
    for (size_t i = 0; i + 2 < p.size(); i += 3) soup->vertices.emplace_back(p[i], p[i + 1], p[i + 2]);
    soup->faces.reserve(parse_data.indices.size() / 3);
    for (size_t i = 0; i + 2 < parse_data.indices.size(); i += 3)
        soup->faces.push_back({ parse_data.indices[i].position_index, parse_data.indices[i + 1].position_index,
                                parse_data.indices[i + 2].position_index });
    soup->materials.push_back(std::move(mat));
    return soup;
}

} // namespace artb
