// rt_shade.cuh — hit-record reconstruction, textures and material scattering (the "shade" stage).
//
// Replaces: hit_record / set_face_normal (hittable.h:9-23), the record-filling halves of every
// primitive's hit() and of translate / rotate_y (hittable.cpp:8-11, 71-84), texture::value
// (texture.h:16-154, perlin.h:21-96), material::emitted / scatter (material.h:10-135) and the vec3
// sampling helpers (vec3.h:117-154).
//
// Sampling: the reference draws points in the unit ball / disk by rejection (vec3.h:117-143). The device
// draws from the SAME distributions in closed form (no divergent loops):
//   uniform on the sphere  z = 1-2*xi1, phi = 2*pi*xi2         == unit_vector(random_in_unit_sphere())
//   uniform in the ball    direction above, radius cbrt(xi3)    == random_in_unit_sphere()
//   uniform in the disk    r = sqrt(xi1), phi = 2*pi*xi2        == random_in_unit_disk()
#ifndef ARTB_RT_SHADE_CUH
#define ARTB_RT_SHADE_CUH

#include "rt_intersect.cuh"

namespace artb {

#define ARTB_PI 3.1415926535897932385

struct HitInfo {
    d3 p, normal;
    double t, u, v;
    bool front_face;
    uint32_t material;
};

__device__ __forceinline__ void set_face_normal(HitInfo& h, d3 ray_dir, d3 outward_normal)
{
    // hittable.h:18-22
    h.front_face = dot(ray_dir, outward_normal) < 0;
    h.normal = h.front_face ? outward_normal : -outward_normal;
}

__device__ __forceinline__ void sphere_uv(d3 p, double& u, double& v)
{
    // sphere.h:24-37
    const double theta = acos(-p.y);
    const double phi = atan2(-p.z, p.x) + ARTB_PI;
    u = phi / (2 * ARTB_PI);
    v = theta / ARTB_PI;
}

// Fill the record for primitive `prim` hit at parameter t by world-space ray `world`.
__device__ inline HitInfo make_hit_info(const SceneView& s, int32_t prim, const DRay& world, double t)
{
    const FlatPrim& p = s.prims[prim];
    HitInfo h;
    h.t = t; h.u = 0; h.v = 0;          // moving_sphere / constant_medium leave u,v untouched (Q8); they are unused there
    h.material = p.material;

    // the ray in object space, and its DIRECTION at every level of the transform chain (the wrappers re-orient the
    // normal against their own inner ray on the way out). Untransformed primitives — most hits — touch no local memory.
    DRay r = world;
    d3 inner_dir[8];
    uint32_t n_ops = 0, first_op = 0;
    if (p.xform) {
        const FlatXform xf = s.xforms[p.xform];
        n_ops = xf.n_ops; first_op = xf.first_op;
        for (uint32_t i = 0; i < n_ops; ++i) { xform_op_in(s.xform_ops[first_op + i], r); inner_dir[i] = r.d; }
    }

    h.p = ray_at(r, t);
    switch (p.type) {
        case PRIM_SPHERE: {
            const d3 outward = (h.p - ld3(p.g)) / p.g[3];          // sphere.h:59-62
            set_face_normal(h, r.d, outward);
            sphere_uv(outward, h.u, h.v);
            break;
        }
        case PRIM_MSPHERE: {
            const d3 outward = (h.p - msphere_center(p, r.tm)) / p.g[8];   // moving_sphere.h:53-54
            set_face_normal(h, r.d, outward);
            break;
        }
        case PRIM_RECT_XY: case PRIM_RECT_XZ: case PRIM_RECT_YZ: {
            // aarect.cpp:8-17 — u,v from the in-plane coordinates, computed as o + t*d per component
            double a, b; d3 outward;
            if (p.type == PRIM_RECT_XY)      { a = r.o.x + t * r.d.x; b = r.o.y + t * r.d.y; outward = mk3(0, 0, 1); }
            else if (p.type == PRIM_RECT_XZ) { a = r.o.x + t * r.d.x; b = r.o.z + t * r.d.z; outward = mk3(0, 1, 0); }
            else                             { a = r.o.y + t * r.d.y; b = r.o.z + t * r.d.z; outward = mk3(1, 0, 0); }
            h.u = (a - p.g[0]) / (p.g[1] - p.g[0]);
            h.v = (b - p.g[2]) / (p.g[3] - p.g[2]);
            set_face_normal(h, r.d, outward);
            break;
        }
        case PRIM_TRIANGLE: {
            // triangle.h:28-86 — UNNORMALISED normal, barycentrics of p1 and p2
            const d3 pt1 = ld3(p.g), pt2 = ld3(p.g + 3), pt3 = ld3(p.g + 6);
            const d3 n = cross(pt2 - pt1, pt3 - pt1);
            const double uu = dot(n, cross(pt3 - pt2, h.p - pt2));
            const double vv = dot(n, cross(pt1 - pt3, h.p - pt3));
            set_face_normal(h, r.d, n);
            h.u = uu / length_squared(n);
            h.v = vv / length_squared(n);
            break;
        }
        default: {
            // constant_medium.h:78-80
            h.normal = mk3(1, 0, 0);
            h.front_face = true;
            break;
        }
    }

    // unwind the wrappers, innermost first (hittable.cpp:8-11 and 71-84): each one maps p (and, for
    // rotate_y, the normal) back to its parent space and re-orients the normal against ITS OWN inner ray
    for (uint32_t i = n_ops; i-- > 0;) {
        const FlatXformOp& op = s.xform_ops[first_op + i];
        if (op.type == XFORM_TRANSLATE) {
            h.p = h.p + mk3(op.a, op.b, op.c);
            set_face_normal(h, inner_dir[i], h.normal);
        } else {
            const double px = op.b * h.p.x + op.a * h.p.z;
            const double pz = -op.a * h.p.x + op.b * h.p.z;
            const double nx = op.b * h.normal.x + op.a * h.normal.z;
            const double nz = -op.a * h.normal.x + op.b * h.normal.z;
            h.p.x = px; h.p.z = pz;
            set_face_normal(h, inner_dir[i], mk3(nx, h.normal.y, nz));
        }
    }
    return h;
}

// ---- textures -------------------------------------------------------------------------------------
__device__ inline double perlin_noise(const FlatPerlin& pn, d3 p)
{
    // perlin.h:21-40, 79-96
    const double fx = floor(p.x), fy = floor(p.y), fz = floor(p.z);
    const double u = p.x - fx, v = p.y - fy, w = p.z - fz;
    const int i = int(fx), j = int(fy), k = int(fz);
    const double uu = u * u * (3 - 2 * u), vv = v * v * (3 - 2 * v), ww = w * w * (3 - 2 * w);
    double accum = 0.0;
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
        for (int dj = 0; dj < 2; ++dj)
#pragma unroll
            for (int dk = 0; dk < 2; ++dk) {
                const int idx = pn.perm_x[(i + di) & 255] ^ pn.perm_y[(j + dj) & 255] ^ pn.perm_z[(k + dk) & 255];
                const d3 c = ld3(pn.ranvec[idx]);
                const d3 weight_v = mk3(u - di, v - dj, w - dk);
                accum += (di * uu + (1 - di) * (1 - uu)) * (dj * vv + (1 - dj) * (1 - vv)) * (dk * ww + (1 - dk) * (1 - ww)) *
                         dot(c, weight_v);
            }
    return accum;
}

__device__ inline d3 image_value(const SceneView& s, uint32_t image, double u, double v)
{
    // texture.h:86-111
    const FlatImage im = s.images[image];
    if (im.width == 0) return mk3(0, 1, 1);
    u = fmin(fmax(u, 0.0), 1.0);
    v = 1.0 - fmin(fmax(v, 0.0), 1.0);
    int i = int(u * im.width);
    int j = int(v * im.height);
    if (i >= im.width) i = im.width - 1;
    if (j >= im.height) j = im.height - 1;
    const double color_scale = 1.0 / 255.0;
    const unsigned char* px = s.texels + im.offset + size_t(j) * size_t(im.width * im.bytes_per_pixel) + size_t(i * im.bytes_per_pixel);
    return mk3(color_scale * px[0], color_scale * px[1], color_scale * px[2]);
}

__device__ inline d3 texture_value(const SceneView& s, uint32_t tex, double u, double v, d3 p)
{
    // checker textures nest (texture.h:31-50); resolve them iteratively
    for (int guard = 0; guard < 16; ++guard) {
        const FlatTexture& t = s.textures[tex];
        switch (t.type) {
            case TEX_SOLID: return ld3(t.p);
            case TEX_CHECKER: {
                const double sines = sin(10 * p.x) * sin(10 * p.y) * sin(10 * p.z);
                tex = sines < 0 ? t.b : t.a;
                break;
            }
            case TEX_NOISE: {
                const double n = perlin_noise(s.perlins[t.a], t.p[0] * p);           // texture.h:57-58
                return mk3(1, 1, 1) * 0.5 * (1.0 + n);
            }
            case TEX_IMAGE: return image_value(s, t.a, u, v);
            case TEX_BARY_IMAGE: {
                const double w = 1 - u - v;                                          // texture.h:141-147
                return image_value(s, t.a, u * t.p[0] + v * t.p[2] + w * t.p[4], u * t.p[1] + v * t.p[3] + w * t.p[5]);
            }
            default: {
                return u * ld3(t.p) + v * ld3(t.p + 3) + (1.0 - u - v) * ld3(t.p + 6);   // texture.h:125-127
            }
        }
    }
    return mk3(0, 0, 0);
}

// ---- sampling -------------------------------------------------------------------------------------
// Random source of the samplers below: the Philox stream of a path (rt_math.cuh), or an explicit table of numbers —
// the parity hooks artb_shade_batch / artb_camera_rays feed the SAME xi to the device and to the oracle.
struct XiTable {
    const double* p;
    __device__ __forceinline__ double next_double() { return *p++; }
};

template <class Rng>
__device__ __forceinline__ d3 sample_unit_sphere(Rng& g)
{
    const double z = 1.0 - 2.0 * g.next_double();
    const double rxy = sqrt(fmax(0.0, 1.0 - z * z));
    double sn, cs;
    sincospi(2.0 * g.next_double(), &sn, &cs);       // the angle in units of pi: no argument reduction by 2*pi
    return mk3(rxy * cs, rxy * sn, z);
}

template <class Rng>
__device__ __forceinline__ d3 sample_unit_ball(Rng& g)
{
    const d3 dir = sample_unit_sphere(g);
    return cbrt(g.next_double()) * dir;
}

template <class Rng>
__device__ __forceinline__ void sample_unit_disk(Rng& g, double& x, double& y)
{
    const double r = sqrt(g.next_double());
    double sn, cs;
    sincospi(2.0 * g.next_double(), &sn, &cs);
    x = r * cs; y = r * sn;
}

__device__ __forceinline__ d3 reflect(d3 v, d3 n) { return v - 2 * dot(v, n) * n; }

__device__ __forceinline__ d3 refract(d3 uv, d3 n, double etai_over_etat)
{
    // vec3.h:149-154
    const double cos_theta = fmin(dot(-uv, n), 1.0);
    const d3 r_out_perp = etai_over_etat * (uv + cos_theta * n);
    const d3 r_out_parallel = -sqrt(fabs(1.0 - length_squared(r_out_perp))) * n;
    return r_out_perp + r_out_parallel;
}

// material::emitted (material.h:12-14, 113-115)
__device__ __forceinline__ d3 material_emitted(const SceneView& s, const FlatMaterial& m, const HitInfo& h)
{
    if (m.type == MAT_DIFFUSE_LIGHT) return texture_value(s, m.texture, h.u, h.v, h.p);
    return mk3(0, 0, 0);
}

// material::scatter (material.h:25-39, 49-56, 67-87, 106-110, 125-131); false = path ends here
template <class Rng>
__device__ inline bool material_scatter(const SceneView& s, const FlatMaterial& m, const DRay& r_in, const HitInfo& h,
                                        Rng& g, d3& attenuation, DRay& scattered)
{
    scattered.o = h.p;
    scattered.tm = r_in.tm;
    switch (m.type) {
        case MAT_LAMBERTIAN: {
            d3 dir = h.normal + sample_unit_sphere(g);
            if (fabs(dir.x) < 1e-8 && fabs(dir.y) < 1e-8 && fabs(dir.z) < 1e-8) dir = h.normal;
            scattered.d = dir;
            attenuation = texture_value(s, m.texture, h.u, h.v, h.p);
            return true;
        }
        case MAT_METAL: {
            const d3 reflected = reflect(unit_vector(r_in.d), h.normal);
            scattered.d = reflected + m.param * sample_unit_ball(g);
            attenuation = ld3(m.albedo);
            return dot(scattered.d, h.normal) > 0;
        }
        case MAT_DIELECTRIC: {
            attenuation = mk3(1.0, 1.0, 1.0);
            const double refraction_ratio = h.front_face ? (1.0 / m.param) : m.param;
            const d3 unit_direction = unit_vector(r_in.d);
            const double cos_theta = fmin(dot(-unit_direction, h.normal), 1.0);
            const double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
            const bool cannot_refract = refraction_ratio * sin_theta > 1.0;
            bool do_reflect = cannot_refract;
            if (!do_reflect) {
                double r0 = (1 - refraction_ratio) / (1 + refraction_ratio);     // Schlick, material.h:93-98
                r0 = r0 * r0;
                const double x = 1 - cos_theta;
                const double reflectance = r0 + (1 - r0) * (x * x * x * x * x);
                do_reflect = reflectance > g.next_double();
            }
            scattered.d = do_reflect ? reflect(unit_direction, h.normal) : refract(unit_direction, h.normal, refraction_ratio);
            return true;
        }
        case MAT_ISOTROPIC: {
            scattered.d = sample_unit_ball(g);
            attenuation = texture_value(s, m.texture, h.u, h.v, h.p);
            return true;
        }
        default:
            return false;   // diffuse_light
    }
}

// camera::get_ray (camera.h:38-47) for image-plane coordinates (s,t)
template <class Rng>
__device__ __forceinline__ DRay camera_ray(const CameraParams& c, double sx, double ty, Rng& g)
{
    double dx, dy;
    sample_unit_disk(g, dx, dy);
    const d3 rd = mk3(c.lens_radius * dx, c.lens_radius * dy, 0);
    const d3 offset = ld3(c.u) * rd.x + ld3(c.v) * rd.y;
    DRay r;
    r.o = ld3(c.origin) + offset;
    r.d = ld3(c.lower_left_corner) + sx * ld3(c.horizontal) + ty * ld3(c.vertical) - ld3(c.origin) - offset;
    r.tm = c.time0 + (c.time1 - c.time0) * g.next_double();
    return r;
}

} // namespace artb

#endif
