#include "flatten.h"

#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <unordered_map>

namespace artb {

namespace {

struct Box3 {
    double lo[3] = { infinity, infinity, infinity };
    double hi[3] = { -infinity, -infinity, -infinity };
    void grow(const double p[3]) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
    void grow(const Box3& b) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    double half_area() const
    {
        const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
    bool valid() const { return lo[0] <= hi[0]; }
};

struct Item { FlatPrim prim; Box3 box; };

class Flattener {
public:
    explicit Flattener(const FlattenOptions& o) : opt(o)
    {
        out.xforms.push_back({ 0, 0 });   // identity
    }

    FlatScene run(const hittable_list& world, const color& background)
    {
        walk_list(world, false, world_items);
        for (int k = 0; k < 3; ++k) out.background[k] = background[k];
        out.n_leaves_total = next_id;
        build_bvh();
        // boundary prims follow the world prims; patch the media's ranges
        const uint32_t base = uint32_t(out.prims.size());
        for (auto& p : out.prims)
            if (p.type == PRIM_MEDIUM) p.g[1] += double(base);
        for (const auto& it : boundary_items) out.prims.push_back(it.prim);
        return std::move(out);
    }

private:
    FlattenOptions opt;
    FlatScene out;
    std::vector<Item> world_items, boundary_items;
    uint32_t next_id = 0;
    int cur_ref = -1;                       // innermost reference bvh_node being walked
    std::vector<FlatXformOp> chain;
    int chain_id = 0;                       // -1: current chain not materialised yet
    std::unordered_map<const void*, uint32_t> mat_ids, tex_ids, img_ids, perlin_ids;

    // ---- transform chains ------------------------------------------------------------------------
    uint32_t current_xform()
    {
        if (chain.empty()) return 0;
        if (chain_id < 0) {
            chain_id = int(out.xforms.size());
            out.xforms.push_back({ uint32_t(out.xform_ops.size()), uint32_t(chain.size()) });
            out.xform_ops.insert(out.xform_ops.end(), chain.begin(), chain.end());
        }
        return uint32_t(chain_id);
    }

    // object-space box -> world-space box through the current chain (innermost op first on the way out)
    Box3 to_world(const aabb& b) const
    {
        Box3 w;
        for (int i = 0; i < 8; ++i) {
            double p[3] = { (i & 1) ? b.max().x() : b.min().x(), (i & 2) ? b.max().y() : b.min().y(),
                            (i & 4) ? b.max().z() : b.min().z() };
            for (size_t k = chain.size(); k-- > 0;) {
                const FlatXformOp& op = chain[k];
                if (op.type == XFORM_ROTATE_Y) {
                    const double x = op.b * p[0] + op.a * p[2];
                    const double z = -op.a * p[0] + op.b * p[2];
                    p[0] = x; p[2] = z;
                } else {
                    p[0] += op.a; p[1] += op.b; p[2] += op.c;
                }
            }
            w.grow(p);
        }
        for (int k = 0; k < 3; ++k) {   // conservative slack: candidates are re-tested exactly on the device
            const double pad = 1e-9 * (1.0 + std::max(std::fabs(w.lo[k]), std::fabs(w.hi[k])));
            w.lo[k] -= pad; w.hi[k] += pad;
        }
        return w;
    }

    // ---- textures / materials --------------------------------------------------------------------
    uint32_t image_id(const std::shared_ptr<image_texture>& i)
    {
        auto it = img_ids.find(i.get());
        if (it != img_ids.end()) return it->second;
        FlatImage fi{};
        while (out.texels.size() % 16) out.texels.push_back(0);
        fi.offset = out.texels.size();
        fi.width = i->data.empty() ? 0 : i->width;
        fi.height = i->data.empty() ? 0 : i->height;
        fi.bytes_per_pixel = i->bytes_per_pixel;
        out.texels.insert(out.texels.end(), i->data.begin(), i->data.end());
        const uint32_t id = uint32_t(out.images.size());
        out.images.push_back(fi);
        img_ids[i.get()] = id;
        return id;
    }

    uint32_t texture_id(const std::shared_ptr<texture>& t)
    {
        auto it = tex_ids.find(t.get());
        if (it != tex_ids.end()) return it->second;
        FlatTexture ft{};
        switch (t->kind()) {
            case texture_kind::solid: {
                const auto& s = static_cast<const solid_color&>(*t);
                ft.type = TEX_SOLID;
                for (int k = 0; k < 3; ++k) ft.p[k] = s.color_value[k];
                break;
            }
            case texture_kind::checker: {
                const auto& c = static_cast<const checker_texture&>(*t);
                ft.type = TEX_CHECKER;
                ft.a = texture_id(c.even);
                ft.b = texture_id(c.odd);
                break;
            }
            case texture_kind::noise: {
                const auto& n = static_cast<const noise_texture&>(*t);
                ft.type = TEX_NOISE;
                ft.p[0] = n.scale;
                FlatPerlin fp;
                for (int i = 0; i < 256; ++i) {
                    for (int k = 0; k < 3; ++k) fp.ranvec[i][k] = n.noise.ranvec[size_t(i)][k];
                    fp.perm_x[i] = n.noise.perm_x[size_t(i)];
                    fp.perm_y[i] = n.noise.perm_y[size_t(i)];
                    fp.perm_z[i] = n.noise.perm_z[size_t(i)];
                }
                ft.a = uint32_t(out.perlins.size());
                out.perlins.push_back(fp);
                break;
            }
            case texture_kind::image: {
                ft.type = TEX_IMAGE;
                ft.a = image_id(std::static_pointer_cast<image_texture>(t));
                break;
            }
            case texture_kind::barycentric_image: {
                const auto& b = static_cast<const barycentric_image_texture&>(*t);
                ft.type = TEX_BARY_IMAGE;
                ft.p[0] = b.texcoord_a.first; ft.p[1] = b.texcoord_a.second;
                ft.p[2] = b.texcoord_b.first; ft.p[3] = b.texcoord_b.second;
                ft.p[4] = b.texcoord_c.first; ft.p[5] = b.texcoord_c.second;
                ft.a = image_id(b.tex);
                break;
            }
            case texture_kind::barycentric: {
                const auto& b = static_cast<const barycentric_texture&>(*t);
                ft.type = TEX_BARY;
                for (int k = 0; k < 3; ++k) { ft.p[k] = b.color_a[k]; ft.p[3 + k] = b.color_b[k]; ft.p[6 + k] = b.color_c[k]; }
                break;
            }
        }
        const uint32_t id = uint32_t(out.textures.size());
        out.textures.push_back(ft);
        tex_ids[t.get()] = id;
        return id;
    }

    uint32_t material_id(const std::shared_ptr<material>& m)
    {
        auto it = mat_ids.find(m.get());
        if (it != mat_ids.end()) return it->second;
        FlatMaterial fm{};
        switch (m->kind()) {
            case material_kind::lambertian:
                fm.type = MAT_LAMBERTIAN; fm.texture = texture_id(static_cast<const lambertian&>(*m).albedo); break;
            case material_kind::metal: {
                const auto& me = static_cast<const metal&>(*m);
                fm.type = MAT_METAL; fm.param = me.fuzz;
                for (int k = 0; k < 3; ++k) fm.albedo[k] = me.albedo[k];
                break;
            }
            case material_kind::dielectric:
                fm.type = MAT_DIELECTRIC; fm.param = static_cast<const dielectric&>(*m).ir; break;
            case material_kind::diffuse_light:
                fm.type = MAT_DIFFUSE_LIGHT; fm.texture = texture_id(static_cast<const diffuse_light&>(*m).emit); break;
            case material_kind::isotropic:
                fm.type = MAT_ISOTROPIC; fm.texture = texture_id(static_cast<const isotropic&>(*m).albedo); break;
        }
        const uint32_t id = uint32_t(out.materials.size());
        out.materials.push_back(fm);
        mat_ids[m.get()] = id;
        return id;
    }

    // ---- graph walk ------------------------------------------------------------------------------
    void emit(std::vector<Item>& dst, bool culled, uint32_t type, const std::shared_ptr<material>& m,
              const aabb& object_box, std::initializer_list<double> g)
    {
        const uint32_t id = next_id++;
        if (culled) { out.n_culled++; return; }
        Item it;
        std::memset(&it.prim, 0, sizeof(it.prim));
        int k = 0;
        for (double d : g) it.prim.g[k++] = d;
        it.prim.type = type;
        it.prim.material = material_id(m);
        it.prim.xform = current_xform();
        it.prim.dfs_id = id;
        it.prim.g[9] = double(cur_ref + 1);      // innermost enclosing reference bvh_node (no primitive type uses g[9])
        it.box = to_world(object_box);
        dst.push_back(it);
    }

    void walk_list(const hittable_list& l, bool culled, std::vector<Item>& dst)
    {
        for (const auto& o : l.objects) walk(*o, culled, dst);
    }

    void walk(const hittable& h, bool culled, std::vector<Item>& dst)
    {
        aabb b;
        switch (h.kind()) {
            case hittable_kind::list: walk_list(static_cast<const hittable_list&>(h), culled, dst); break;
            case hittable_kind::bvh: {
                const auto& n = static_cast<const bvh_node&>(h);
                bool flat = false;
                for (int k = 0; k < 3; ++k) flat = flat || !(n.box.min()[k] < n.box.max()[k]);
                // remember the reference's own node: its box decides exact ties (FlatRefNode)
                FlatRefNode rn{};
                for (int k = 0; k < 3; ++k) { rn.lo[k] = n.box.min()[k]; rn.hi[k] = n.box.max()[k]; }
                rn.parent = cur_ref;
                rn.xform = current_xform();
                rn.depth = cur_ref < 0 ? 0 : out.ref_nodes[size_t(cur_ref)].depth + 1;
                const int saved_ref = cur_ref;
                cur_ref = int(out.ref_nodes.size());
                out.ref_nodes.push_back(rn);
                walk(*n.left, culled || flat, dst);
                if (n.left.get() != n.right.get()) walk(*n.right, culled || flat, dst);
                cur_ref = saved_ref;
                break;
            }
            case hittable_kind::translate: {
                const auto& t = static_cast<const translate&>(h);
                push_op({ t.offset.x(), t.offset.y(), t.offset.z(), XFORM_TRANSLATE, 0 }, *t.ptr, culled, dst);
                break;
            }
            case hittable_kind::rotate_y: {
                const auto& r = static_cast<const rotate_y&>(h);
                push_op({ r.sin_theta, r.cos_theta, 0.0, XFORM_ROTATE_Y, 0 }, *r.ptr, culled, dst);
                break;
            }
            case hittable_kind::box: walk_list(static_cast<const box&>(h).sides, culled, dst); break;
            case hittable_kind::sphere: {
                const auto& s = static_cast<const sphere&>(h);
                s.bounding_box(0, 1, b);
                emit(dst, culled, PRIM_SPHERE, s.mat_ptr, b, { s.center.x(), s.center.y(), s.center.z(), s.radius });
                break;
            }
            case hittable_kind::moving_sphere: {
                const auto& s = static_cast<const moving_sphere&>(h);
                s.bounding_box(0, 1, b);
                emit(dst, culled, PRIM_MSPHERE, s.mat_ptr, b,
                     { s.center0.x(), s.center0.y(), s.center0.z(), s.center1.x(), s.center1.y(), s.center1.z(),
                       s.time0, s.time1, s.radius });
                break;
            }
            case hittable_kind::xy_rect: {
                const auto& r = static_cast<const xy_rect&>(h);
                r.bounding_box(0, 1, b);
                emit(dst, culled, PRIM_RECT_XY, r.mp, b, { r.x0, r.x1, r.y0, r.y1, r.k });
                break;
            }
            case hittable_kind::xz_rect: {
                const auto& r = static_cast<const xz_rect&>(h);
                r.bounding_box(0, 1, b);
                emit(dst, culled, PRIM_RECT_XZ, r.mp, b, { r.x0, r.x1, r.z0, r.z1, r.k });
                break;
            }
            case hittable_kind::yz_rect: {
                const auto& r = static_cast<const yz_rect&>(h);
                r.bounding_box(0, 1, b);
                emit(dst, culled, PRIM_RECT_YZ, r.mp, b, { r.y0, r.y1, r.z0, r.z1, r.k });
                break;
            }
            case hittable_kind::triangle: {
                const auto& t = static_cast<const triangle&>(h);
                t.bounding_box(0, 1, b);
                emit(dst, culled, PRIM_TRIANGLE, t.mat_ptr, b,
                     { t.pt1.x(), t.pt1.y(), t.pt1.z(), t.pt2.x(), t.pt2.y(), t.pt2.z(), t.pt3.x(), t.pt3.y(), t.pt3.z() });
                break;
            }
            case hittable_kind::triangle_soup: {
                const auto& s = static_cast<const triangle_soup&>(h);
                if (s.materials.size() != 1 && s.materials.size() != s.faces.size())
                    throw std::logic_error("triangle_soup: need one material or one per face");
                for (size_t f = 0; f < s.faces.size(); ++f) {
                    const vec3& p1 = s.vertices[size_t(s.faces[f][0])];
                    const vec3& p2 = s.vertices[size_t(s.faces[f][1])];
                    const vec3& p3 = s.vertices[size_t(s.faces[f][2])];
                    emit(dst, culled, PRIM_TRIANGLE, s.materials[s.materials.size() == 1 ? 0 : f],
                         aabb(min(p1, min(p2, p3)), max(p1, max(p2, p3))),
                         { p1.x(), p1.y(), p1.z(), p2.x(), p2.y(), p2.z(), p3.x(), p3.y(), p3.z() });
                }
                break;
            }
            case hittable_kind::constant_medium: {
                const auto& m = static_cast<const constant_medium&>(h);
                if (&dst == &boundary_items)
                    throw std::logic_error("flatten: a constant_medium inside a medium boundary is not supported");
                const uint32_t id = next_id++;
                const size_t first = boundary_items.size();
                walk(*m.boundary, false, boundary_items);   // ids continue; boundary is queried directly (constant_medium.h:44-47)
                const size_t count = boundary_items.size() - first;
                if (culled) { out.n_culled++; break; }
                Item it;
                std::memset(&it.prim, 0, sizeof(it.prim));
                it.prim.g[0] = m.neg_inv_density;
                it.prim.g[1] = double(first);           // rebased once the world prim count is known
                it.prim.g[2] = double(count);
                it.prim.g[3] = double(out.n_media++);
                it.prim.type = PRIM_MEDIUM;
                it.prim.material = material_id(m.phase_function);
                it.prim.xform = current_xform();
                it.prim.dfs_id = id;
                for (size_t k = first; k < first + count; ++k) it.box.grow(boundary_items[k].box);
                if (!it.box.valid()) break;              // empty boundary: can never be hit
                dst.push_back(it);
                break;
            }
        }
    }

    void push_op(const FlatXformOp& op, const hittable& child, bool culled, std::vector<Item>& dst)
    {
        if (chain.size() >= 8) throw std::logic_error("flatten: transform chains deeper than 8 are not supported");
        const int saved = chain_id;
        chain.push_back(op);
        chain_id = -1;
        walk(child, culled, dst);
        chain.pop_back();
        chain_id = saved;
    }

    // ---- binned-SAH BVH over world boxes ---------------------------------------------------------
    struct TNode { Box3 box; int left = -1, right = -1; int first = 0, count = 0; };
    std::vector<TNode> tnodes;
    std::vector<int> order;

    int build_range(int first, int count, int depth)
    {
        out.bvh_depth = std::max(out.bvh_depth, depth);
        const int me = int(tnodes.size());
        tnodes.emplace_back();
        Box3 box, cbox;
        for (int i = first; i < first + count; ++i) {
            const Box3& b = world_items[size_t(order[size_t(i)])].box;
            box.grow(b);
            const double c[3] = { 0.5 * (b.lo[0] + b.hi[0]), 0.5 * (b.lo[1] + b.hi[1]), 0.5 * (b.lo[2] + b.hi[2]) };
            cbox.grow(c);
        }
        tnodes[size_t(me)].box = box;
        tnodes[size_t(me)].first = first;
        tnodes[size_t(me)].count = count;
        if (count <= 1) return me;

        const int nb = opt.sah_bins;
        double best_cost = infinity;
        int best_axis = -1, best_split = -1;
        for (int axis = 0; axis < 3; ++axis) {
            const double extent = cbox.hi[axis] - cbox.lo[axis];
            if (!(extent > 0)) continue;
            std::vector<Box3> bb((size_t(nb)));
            std::vector<int> bc(size_t(nb), 0);
            const double scale = nb / extent;
            for (int i = first; i < first + count; ++i) {
                const Box3& b = world_items[size_t(order[size_t(i)])].box;
                int bin = int((0.5 * (b.lo[axis] + b.hi[axis]) - cbox.lo[axis]) * scale);
                bin = std::min(std::max(bin, 0), nb - 1);
                bb[size_t(bin)].grow(b);
                bc[size_t(bin)]++;
            }
            std::vector<double> right_area(size_t(nb), 0.0);
            std::vector<int> right_count(size_t(nb), 0);
            Box3 acc; int cnt = 0;
            for (int i = nb - 1; i > 0; --i) {
                acc.grow(bb[size_t(i)]); cnt += bc[size_t(i)];
                right_area[size_t(i)] = cnt ? acc.half_area() : 0.0;
                right_count[size_t(i)] = cnt;
            }
            acc = Box3(); cnt = 0;
            for (int i = 0; i < nb - 1; ++i) {
                acc.grow(bb[size_t(i)]); cnt += bc[size_t(i)];
                if (cnt == 0 || right_count[size_t(i + 1)] == 0) continue;
                const double cost = acc.half_area() * cnt + right_area[size_t(i + 1)] * right_count[size_t(i + 1)];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = i; }
            }
        }

        int mid = -1;
        if (best_axis >= 0) {
            const double leaf_cost = opt.intersect_cost * count;
            const double split_cost = opt.traversal_cost + opt.intersect_cost * best_cost / std::max(box.half_area(), 1e-300);
            if (count <= opt.max_leaf_prims && leaf_cost <= split_cost) return me;   // stay a leaf
            const double extent = cbox.hi[best_axis] - cbox.lo[best_axis];
            const double scale = nb / extent;
            auto begin = order.begin() + first, end = begin + count;
            auto it = std::partition(begin, end, [&](int idx) {
                const Box3& b = world_items[size_t(idx)].box;
                int bin = int((0.5 * (b.lo[best_axis] + b.hi[best_axis]) - cbox.lo[best_axis]) * scale);
                bin = std::min(std::max(bin, 0), nb - 1);
                return bin <= best_split;
            });
            mid = int(it - order.begin());
        }
        if (mid <= first || mid >= first + count) {
            if (count <= opt.max_leaf_prims) return me;
            mid = first + count / 2;   // coincident centroids: split by index
        }
        const int l = build_range(first, mid - first, depth + 1);
        const int r = build_range(mid, first + count - mid, depth + 1);
        tnodes[size_t(me)].left = l;
        tnodes[size_t(me)].right = r;
        return me;
    }

    static float round_down(double x)
    {
        float f = float(x);
        if (double(f) > x) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
        return f;
    }
    static float round_up(double x)
    {
        float f = float(x);
        if (double(f) < x) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
        return f;
    }

    // ---- which binary subtrees become the children of one 8-wide node ---------------------------------------------
    // Dynamic programme of Ylitie, Karras & Laine (HPG 2017, sec. 4.1) over the binary SAH tree: wide_cost[n][i] is the
    // least expected number of wide-node visits (sum of half-areas of the wide nodes created) when the subtree of
    // binary node n is represented by AT MOST i roots, each root being a wide node or a leaf. A wide node at n costs
    // area(n) plus the best forest of <= 8 roots below it; fewer roots can always stand in for more. Leaves are the
    // binary tree's leaves (their primitives are not regrouped), so their cost is a constant and left out.
    std::vector<std::array<double, 9>> wide_cost;
    std::vector<std::array<int8_t, 9>> wide_split;      // 0: one root here (wide node or leaf), k > 0: k roots from the left child

    void distribute(int n, int roots, double& cost, int& split) const
    {
        const TNode& t = tnodes[size_t(n)];
        cost = infinity; split = 1;
        for (int k = 1; k < roots; ++k) {
            const double c = wide_cost[size_t(t.left)][size_t(k)] + wide_cost[size_t(t.right)][size_t(roots - k)];
            if (c < cost) { cost = c; split = k; }
        }
    }

    void build_wide_cost()
    {
        wide_cost.assign(tnodes.size(), {});
        wide_split.assign(tnodes.size(), {});
        // children have larger indices than their parents (build_range appends depth first): go backwards
        for (int n = int(tnodes.size()) - 1; n >= 0; --n) {
            const TNode& t = tnodes[size_t(n)];
            auto& c = wide_cost[size_t(n)];
            auto& sp = wide_split[size_t(n)];
            if (t.left < 0) {
                for (int i = 0; i <= 8; ++i) { c[size_t(i)] = 0.0; sp[size_t(i)] = 0; }
                continue;
            }
            double dc; int ds;
            distribute(n, 8, dc, ds);
            c[1] = t.box.half_area() + dc;
            sp[1] = 0;
            for (int i = 2; i <= 8; ++i) {
                distribute(n, i, dc, ds);
                if (dc < c[size_t(i - 1)]) { c[size_t(i)] = dc; sp[size_t(i)] = int8_t(ds); }
                else { c[size_t(i)] = c[size_t(i - 1)]; sp[size_t(i)] = sp[size_t(i - 1)]; }
            }
        }
    }

    // the roots the table chose for subtree n with a budget of `roots`; open_here: n itself is the wide node being
    // filled, so it must be opened rather than referenced
    void collect_children(int n, int roots, bool open_here, std::vector<int>& out) const
    {
        const TNode& t = tnodes[size_t(n)];
        if (t.left < 0) { out.push_back(n); return; }
        int split;
        if (open_here) { double c; distribute(n, roots, c, split); }
        else {
            split = wide_split[size_t(n)][size_t(roots)];
            if (split == 0) { out.push_back(n); return; }         // a wide node of its own
            // the entry may have been inherited from a smaller budget: find the budget it was computed for
            int r = roots;
            while (r > 2 && wide_cost[size_t(n)][size_t(r)] == wide_cost[size_t(n)][size_t(r - 1)] &&
                   wide_split[size_t(n)][size_t(r - 1)] == split) --r;
            roots = r;
        }
        collect_children(t.left, split, false, out);
        collect_children(t.right, roots - split, false, out);
    }

    // ---- collapse the binary tree into 8-wide compressed nodes (flat_scene.h: FlatNode8) -----------------------
    void emit_bvh8(int root_t)
    {
        struct Job { int tnode; uint32_t node; int depth; };
        std::vector<Job> queue;
        out.nodes.emplace_back();
        queue.push_back({ root_t, 0u, 1 });
        for (size_t qi = 0; qi < queue.size(); ++qi) {
            const Job job = queue[qi];
            out.bvh_depth = std::max(out.bvh_depth, job.depth);
            // children: the forest of at most eight subtrees the cost table picked for this binary node
            std::vector<int> kids;
            const TNode& me = tnodes[size_t(job.tnode)];
            if (me.left < 0) kids.push_back(job.tnode);
            else collect_children(job.tnode, 8, true, kids);
            // slots: a ray of octant o visits slot o first, so slot s wants the child that is nearest along
            // dir_s = (s&4 ? -1 : +1, s&2 ? -1 : +1, s&1 ? -1 : +1). Greedy assignment on that cost.
            double centre[3];
            for (int a = 0; a < 3; ++a) centre[a] = 0.5 * (me.box.lo[a] + me.box.hi[a]);
            int slot_of[8], kid_in[8];
            for (int k = 0; k < 8; ++k) { slot_of[k] = -1; kid_in[k] = -1; }
            for (size_t round = 0; round < kids.size(); ++round) {
                double best = infinity; int bk = -1, bs = -1;
                for (size_t k = 0; k < kids.size(); ++k) {
                    if (slot_of[k] >= 0) continue;
                    const Box3& cb = tnodes[size_t(kids[k])].box;
                    for (int sl = 0; sl < 8; ++sl) {
                        if (kid_in[sl] >= 0) continue;
                        double cost = 0;
                        for (int a = 0; a < 3; ++a) {
                            const double dir = (sl >> (2 - a)) & 1 ? -1.0 : 1.0;
                            cost += dir * (0.5 * (cb.lo[a] + cb.hi[a]) - centre[a]);
                        }
                        if (cost < best) { best = cost; bk = int(k); bs = sl; }
                    }
                }
                slot_of[bk] = bs; kid_in[bs] = bk;
            }
            // quantisation grid: origin = box.lo rounded down to float, cell = smallest power of two that keeps
            // every plane of every child within 255 cells
            FlatNode8 n;
            std::memset(&n, 0, sizeof(n));
            double cell[3];
            for (int a = 0; a < 3; ++a) {
                n.p[a] = round_down(me.box.lo[a]);
                int e = 0;
                const double extent = me.box.hi[a] - double(n.p[a]);
                std::frexp(extent / 255.0, &e);                    // extent/255 <= 2^e
                if (e < -100) e = -100;
                while (std::ceil(extent / std::ldexp(1.0, e)) > 255.0) ++e;
                if (e > 100) throw std::logic_error("flatten: scene extent out of range for the quantised BVH");
                n.e[a] = uint8_t(e + 127);
                cell[a] = std::ldexp(1.0, e);
            }
            n.child_base = uint32_t(out.nodes.size());
            n.prim_base = uint32_t(out.prims.size());
            uint32_t prim_offset = 0;
            for (int sl = 0; sl < 8; ++sl) {
                if (kid_in[sl] < 0) continue;
                const int t = kids[size_t(kid_in[sl])];
                const TNode& c = tnodes[size_t(t)];
                for (int a = 0; a < 3; ++a) {
                    double lo = std::floor((c.box.lo[a] - double(n.p[a])) / cell[a]);
                    double hi = std::ceil((c.box.hi[a] - double(n.p[a])) / cell[a]);
                    while (double(n.p[a]) + lo * cell[a] > c.box.lo[a]) lo -= 1;     // exact in double: p float, cell 2^e, small integers
                    while (double(n.p[a]) + hi * cell[a] < c.box.hi[a]) hi += 1;
                    if (lo < 0 || hi > 255) throw std::logic_error("flatten: quantised BVH box out of range");
                    n.qlo[a][sl] = uint8_t(lo);
                    n.qhi[a][sl] = uint8_t(hi);
                }
                if (c.left >= 0) {
                    n.imask = uint8_t(n.imask | (1u << sl));
                    n.meta[sl] = uint8_t((1u << 5) | (24u + uint32_t(sl)));
                    queue.push_back({ t, uint32_t(out.nodes.size()), job.depth + 1 });
                    out.nodes.emplace_back();
                } else {
                    if (c.count < 1 || c.count > ARTB_MAX_LEAF_PRIMS) throw std::logic_error("flatten: bad leaf size");
                    n.meta[sl] = uint8_t((((1u << c.count) - 1u) << 5) | prim_offset);
                    for (int i = c.first; i < c.first + c.count; ++i) out.prims.push_back(world_items[size_t(order[size_t(i)])].prim);
                    prim_offset += uint32_t(c.count);
                }
            }
            out.nodes[job.node] = n;
        }
    }

    void build_bvh()
    {
        // 1. global primitives, tested for every ray outside the tree: every constant_medium (its test is two boundary
        //    queries + a random draw: the device runs it once per ray, AFTER the traversal, when the final t_max makes
        //    the interval short; g[4..9] keep the medium's world box for a cheap pre-test), and surfaces whose boxes span
        //    a large share of the scene (no divergence, and an early t_max for the traversal)
        std::vector<Item> globals, media, rest;
        FlattenOptions opt = this->opt;
        if (const char* e = std::getenv("ARTB_MAX_GLOBAL_PRIMS")) opt.max_global_prims = std::atoi(e);   // tuning knob
        for (const Item& it : world_items) {
            if (it.prim.type == PRIM_MEDIUM) {
                Item m = it;
                for (int k = 0; k < 3; ++k) { m.prim.g[4 + k] = it.box.lo[k]; m.prim.g[7 + k] = it.box.hi[k]; }
                media.push_back(m);
            } else rest.push_back(it);
        }
        for (int pass = 0; pass < 4 && int(globals.size()) < opt.max_global_prims; ++pass) {
            Box3 all;
            for (const Item& it : rest) all.grow(it.box);
            if (rest.size() <= 1 || !(all.half_area() > 0)) break;
            std::vector<Item> keep;
            bool moved = false;
            for (const Item& it : rest) {
                if (int(globals.size()) < opt.max_global_prims && it.box.half_area() >= opt.global_area_fraction * all.half_area()) {
                    globals.push_back(it); moved = true;
                } else keep.push_back(it);
            }
            rest.swap(keep);
            if (!moved) break;
        }
        if (rest.size() + globals.size() <= size_t(opt.max_global_prims)) {   // tiny scene: a plain list
            globals.insert(globals.end(), rest.begin(), rest.end());
            rest.clear();
        }
        // globals keep the reference's test order (dfs), which makes early best-t pruning deterministic; media last
        std::sort(globals.begin(), globals.end(), [](const Item& a, const Item& b) { return a.prim.dfs_id < b.prim.dfs_id; });
        std::sort(media.begin(), media.end(), [](const Item& a, const Item& b) { return a.prim.dfs_id < b.prim.dfs_id; });
        out.n_global_media = uint32_t(media.size());
        globals.insert(globals.end(), media.begin(), media.end());
        out.n_global_prims = uint32_t(globals.size());
        out.n_world_prims = uint32_t(world_items.size());
        Box3 whole;
        for (const Item& it : world_items) whole.grow(it.box);
        for (int k = 0; k < 3; ++k) { out.world_lo[k] = whole.valid() ? whole.lo[k] : 0; out.world_hi[k] = whole.valid() ? whole.hi[k] : 0; }
        out.prims.reserve(world_items.size() + boundary_items.size());
        for (const Item& it : globals) out.prims.push_back(it.prim);

        // 2. binned-SAH binary tree over the rest, collapsed into 8-wide nodes
        world_items.swap(rest);
        const int n = int(world_items.size());
        if (n == 0) return;
        if (opt.max_leaf_prims > ARTB_MAX_LEAF_PRIMS) throw std::logic_error("flatten: max_leaf_prims exceeds ARTB_MAX_LEAF_PRIMS");
        if (opt.device_bvh) {
            // no host tree: the primitives go out in walk order and Renderer::upload builds the tree on the GPU
            for (const Item& it : world_items) out.prims.push_back(it.prim);
            out.device_bvh_leaf_prims = std::max(opt.max_leaf_prims, 1);
            out.bvh_depth = 0;
            return;
        }
        order.resize(size_t(n));
        for (int i = 0; i < n; ++i) order[size_t(i)] = i;
        tnodes.reserve(size_t(2 * n));
        const int root = build_range(0, n, 0);
        out.bvh_depth = 0;
        build_wide_cost();
        emit_bvh8(root);
    }
};

} // namespace

FlatScene flatten(const hittable_list& world, const color& background, const FlattenOptions& opt)
{
    Flattener f(opt);
    return f.run(world, background);
}

} // namespace artb
