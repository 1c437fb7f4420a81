// Host-side math / RNG core of the B200 path-tracing backend.
//
// Mirrors the API surface of the reference's L1 layer so scene-building code written against
// blackccpie/another_raytracer compiles unchanged inside namespace artb:
//   vec3 / point3 / color      reference: src/core/vec3.h:10-69
//   ray                        reference: src/core/ray.h:6-25
//   aabb, surrounding_box      reference: src/primitives/aabb.h:8-39   (storage + union only: the slab
//                                                                        test lives on the device)
//   random_double / random_int reference: src/utils/tracer_utils.h:27-41
//
// The host never intersects rays: there is no hit() here and no CPU fallback. What the host must do
// bit-identically to the reference is *scene construction*, which draws from one process-global,
// default-seeded Mersenne twister (tracer_utils.h:27-31). rng_state below restates MT19937 and the
// libstdc++ generate_canonical<double,53> mapping (two 32-bit draws, first draw = low word) so that
// builders consume the very same stream; tests pin it against the reference's first values
// (0.1354770042967805, 0.8350085899945795, ...).
#ifndef ARTB_HOST_CORE_H
#define ARTB_HOST_CORE_H

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>

namespace artb {

const double infinity = std::numeric_limits<double>::infinity();
const double pi = 3.1415926535897932385;

inline double degrees_to_radians(double degrees) { return degrees * pi / 180.0; }

inline double clamp(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

// ---------------------------------------------------------------------------------------------
// MT19937 (Matsumoto & Nishimura 1998), default seed 5489, + libstdc++'s canonical double mapping.
// ---------------------------------------------------------------------------------------------
class rng_state {
public:
    explicit rng_state(std::uint32_t seed = 5489u) { reseed(seed); }

    void reseed(std::uint32_t seed)
    {
        mt_[0] = seed;
        for (int i = 1; i < N; ++i)
            mt_[i] = 1812433253u * (mt_[i - 1] ^ (mt_[i - 1] >> 30)) + std::uint32_t(i);
        idx_ = N;
        draws_ = 0;
    }

    std::uint32_t next_u32()
    {
        if (idx_ >= N) refill();
        std::uint32_t y = mt_[idx_++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }

    // uniform_real_distribution<double>(0,1) as libstdc++ evaluates it: (lo + hi*2^32) / 2^64, the sum
    // rounded once to double; a result of 1.0 is nudged to the previous representable value.
    double next_double()
    {
        const double lo = double(next_u32());
        const double hi = double(next_u32());
        double r = (lo + hi * 4294967296.0) / 18446744073709551616.0;
        if (r >= 1.0) r = std::nextafter(1.0, 0.0);
        ++draws_;
        return r;
    }

    std::uint64_t draws() const { return draws_; }

private:
    static constexpr int N = 624, M = 397;
    void refill()
    {
        for (int i = 0; i < N; ++i) {
            const std::uint32_t y = (mt_[i] & 0x80000000u) | (mt_[(i + 1) % N] & 0x7fffffffu);
            mt_[i] = mt_[(i + M) % N] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        idx_ = 0;
    }
    std::uint32_t mt_[N];
    int idx_ = N;
    std::uint64_t draws_ = 0;
};

// the process-global stream every builder draws from (reference: function-local static, same lifetime)
rng_state& global_rng();
void reset_global_rng(std::uint32_t seed = 5489u);

inline double random_double() { return global_rng().next_double(); }
inline double random_double(double lo, double hi) { return lo + (hi - lo) * random_double(); }
inline int random_int(int lo, int hi) { return static_cast<int>(random_double(lo, hi + 1)); }

// ---------------------------------------------------------------------------------------------
struct vec3 {
    double e[3];

    vec3() : e{0, 0, 0} {}
    vec3(double a, double b, double c) : e{a, b, c} {}

    double x() const { return e[0]; }
    double y() const { return e[1]; }
    double z() const { return e[2]; }
    double operator[](int i) const { return e[i]; }
    double& operator[](int i) { return e[i]; }

    vec3 operator-() const { return vec3(-e[0], -e[1], -e[2]); }
    vec3& operator+=(const vec3& o) { e[0] += o.e[0]; e[1] += o.e[1]; e[2] += o.e[2]; return *this; }
    vec3& operator*=(double t) { e[0] *= t; e[1] *= t; e[2] *= t; return *this; }
    vec3& operator/=(double t) { return *this *= 1 / t; }   // multiply by reciprocal, like the reference

    double length_squared() const { return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]; }
    double length() const { return std::sqrt(length_squared()); }
    bool near_zero() const
    {
        const double s = 1e-8;
        return std::fabs(e[0]) < s && std::fabs(e[1]) < s && std::fabs(e[2]) < s;
    }

    // NB: the reference writes vec3(random_double(), random_double(), random_double()) (vec3.h:55-61).
    // Constructor arguments are unsequenced in C++ and GCC evaluates them RIGHT TO LEFT, so the
    // reference binary draws z first, then y, then x (verified against its scene dump: first small
    // sphere of the random scene has x = -11+0.9*r[2], z = -11+0.9*r[1]). The order is spelled out here
    // so that this file gives the same scenes under any compiler.
    static vec3 random()
    {
        const double c = random_double(), b = random_double(), a = random_double();
        return vec3(a, b, c);
    }
    static vec3 random(double lo, double hi)
    {
        const double c = random_double(lo, hi), b = random_double(lo, hi), a = random_double(lo, hi);
        return vec3(a, b, c);
    }
};

using point3 = vec3;
using color = vec3;

inline vec3 operator+(const vec3& u, const vec3& v) { return vec3(u.e[0] + v.e[0], u.e[1] + v.e[1], u.e[2] + v.e[2]); }
inline vec3 operator-(const vec3& u, const vec3& v) { return vec3(u.e[0] - v.e[0], u.e[1] - v.e[1], u.e[2] - v.e[2]); }
inline vec3 operator*(const vec3& u, const vec3& v) { return vec3(u.e[0] * v.e[0], u.e[1] * v.e[1], u.e[2] * v.e[2]); }
inline vec3 operator*(double t, const vec3& v) { return vec3(t * v.e[0], t * v.e[1], t * v.e[2]); }
inline vec3 operator*(const vec3& v, double t) { return t * v; }
inline vec3 operator/(const vec3& v, double t) { return (1 / t) * v; }
inline double dot(const vec3& u, const vec3& v) { return u.e[0] * v.e[0] + u.e[1] * v.e[1] + u.e[2] * v.e[2]; }
inline vec3 cross(const vec3& u, const vec3& v)
{
    return vec3(u.e[1] * v.e[2] - u.e[2] * v.e[1], u.e[2] * v.e[0] - u.e[0] * v.e[2], u.e[0] * v.e[1] - u.e[1] * v.e[0]);
}
inline vec3 unit_vector(const vec3& v) { return v / v.length(); }
inline vec3 min(const vec3& a, const vec3& b) { return vec3(std::min(a.x(), b.x()), std::min(a.y(), b.y()), std::min(a.z(), b.z())); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(std::max(a.x(), b.x()), std::max(a.y(), b.y()), std::max(a.z(), b.z())); }

// ---------------------------------------------------------------------------------------------
struct ray {
    point3 orig;
    vec3 dir;
    double tm = 0.0;

    ray() = default;
    ray(const point3& o, const vec3& d, double time = 0.0) : orig(o), dir(d), tm(time) {}
    point3 origin() const { return orig; }
    vec3 direction() const { return dir; }
    double time() const { return tm; }
    point3 at(double t) const { return orig + t * dir; }
};

struct aabb {
    point3 minimum, maximum;
    aabb() = default;
    aabb(const point3& a, const point3& b) : minimum(a), maximum(b) {}
    point3 min() const { return minimum; }
    point3 max() const { return maximum; }
};

inline aabb surrounding_box(const aabb& a, const aabb& b) { return aabb(min(a.min(), b.min()), max(a.max(), b.max())); }

} // namespace artb

#endif
