// rt_math.cuh — double-precision vector helpers and the counter-based RNG used by every kernel.
//
// Arithmetic type: the reference computes everything in IEEE double (vec3::e is double[3], vec3.h:64) and
// its scenes mix radius-1000/5000 spheres with radius-0.2 ones, so the device keeps double for all
// geometry (B200 issues FP64 at half the FP32 rate). Operation ORDER follows the reference's expressions
// so that results agree to the last few ulps (nvcc contracts a*b+c into FMA, which g++ on baseline x86-64
// does not, hence "few ulps" and not "bit-identical").
//
// RNG: Philox4x32-10 (Salmon et al., SC'11), replacing the reference's process-global mt19937
// (tracer_utils.h:27-31). key = (seed, stream), counter = (path id lo, path id hi, draw index, purpose).
#ifndef ARTB_RT_MATH_CUH
#define ARTB_RT_MATH_CUH

#include "rt_portable.cuh"
#include <stdint.h>
#include <math.h>
#include <float.h>

namespace artb {

struct d3 { double x, y, z; };

ARTB_HD d3 mk3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
ARTB_HD d3 operator+(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
ARTB_HD d3 operator-(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
ARTB_HD d3 operator-(d3 a) { return mk3(-a.x, -a.y, -a.z); }
ARTB_HD d3 operator*(d3 a, d3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
ARTB_HD d3 operator*(double t, d3 a) { return mk3(t * a.x, t * a.y, t * a.z); }
ARTB_HD d3 operator*(d3 a, double t) { return mk3(t * a.x, t * a.y, t * a.z); }
ARTB_HD d3 operator/(d3 a, double t) { const double s = 1 / t; return mk3(s * a.x, s * a.y, s * a.z); }
ARTB_HD double dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
ARTB_HD d3 cross(d3 a, d3 b)
{
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
ARTB_HD double length_squared(d3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
ARTB_HD double length(d3 a) { return sqrt(length_squared(a)); }
ARTB_HD d3 unit_vector(d3 a) { return a / length(a); }
ARTB_HD d3 ld3(const double* p) { return mk3(p[0], p[1], p[2]); }

struct DRay { d3 o, d; double tm; };

ARTB_HD d3 ray_at(const DRay& r, double t) { return r.o + t * r.d; }

// ---- Philox4x32-10 ------------------------------------------------------------------------------
struct Philox {
    uint32_t key0, key1;
    uint32_t c0, c1, c2, c3;     // c2 = 128-bit block index: the only part that advances
    double cached;               // second double of the current block
    bool has_cached;

    ARTB_HD void init(uint32_t seed, uint32_t stream, uint32_t a, uint32_t b,
                                                  uint32_t block, uint32_t purpose)
    {
        key0 = seed; key1 = stream; c0 = a; c1 = b; c2 = block; c3 = purpose; has_cached = false; cached = 0.0;
    }

    ARTB_HD static void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
    {
#if defined(__CUDA_ARCH__)
        hi = __umulhi(a, b); lo = a * b;
#else
        const uint64_t p = uint64_t(a) * uint64_t(b);
        hi = uint32_t(p >> 32); lo = uint32_t(p);
#endif
    }

    ARTB_HD static double to_unit(uint32_t a, uint32_t b)
    {
        // 53 random bits -> [0,1)
        const uint64_t bits = ((uint64_t(a) << 32) | uint64_t(b)) >> 11;
        return double(bits) * (1.0 / 9007199254740992.0);
    }

    // one Philox4x32-10 block = two doubles
    ARTB_HD void block(double& u0, double& u1)
    {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, k0 = key0, k1 = key1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(0xD2511F53u, x0, hi0, lo0);
            mulhilo(0xCD9E8D57u, x2, hi1, lo1);
            const uint32_t y0 = hi1 ^ x1 ^ k0, y1 = lo1, y2 = hi0 ^ x3 ^ k1, y3 = lo0;
            x0 = y0; x1 = y1; x2 = y2; x3 = y3;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        ++c2;
        u0 = to_unit(x0, x1);
        u1 = to_unit(x2, x3);
    }

    ARTB_HD double next_double()
    {
        if (has_cached) { has_cached = false; return cached; }
        double u0;
        block(u0, cached);
        has_cached = true;
        return u0;
    }
};

} // namespace artb

#endif
