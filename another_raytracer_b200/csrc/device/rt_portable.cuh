// rt_portable.cuh — the handful of hardware intrinsics the traversal uses, with plain C++ equivalents.
//
// The device code of rt_intersect.cuh is written against these wrappers so that the SAME source also compiles with
// g++ into tests/hostsim (TEST INFRASTRUCTURE: a lane-by-lane replay of the traversal used to check the 8-wide tree
// against the golden ray batches on machines without a GPU, and to count node visits / primitive tests). The
// product never executes the host versions: libartb.so has no CPU path.
#ifndef ARTB_RT_PORTABLE_CUH
#define ARTB_RT_PORTABLE_CUH

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define ARTB_HD __host__ __device__ __forceinline__
#define ARTB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define ARTB_HD inline
#define ARTB_HD_NOINLINE inline
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r; r.x = x; r.y = y; return r; }
#endif

namespace artb {

ARTB_HD float as_float(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

ARTB_HD float d2f_rd(double x)     // largest float <= x
{
#if defined(__CUDA_ARCH__)
    return __double2float_rd(x);
#else
    float f = float(x);
    if (double(f) > x) f = nextafterf(f, -INFINITY);
    return f;
#endif
}

ARTB_HD float d2f_ru(double x)     // smallest float >= x
{
#if defined(__CUDA_ARCH__)
    return __double2float_ru(x);
#else
    float f = float(x);
    if (double(f) < x) f = nextafterf(f, INFINITY);
    return f;
#endif
}

ARTB_HD float rcp_rn(float x)      // correctly rounded 1/x (the same value as 1.0f / x, in fewer instructions on the device)
{
#if defined(__CUDA_ARCH__)
    return __frcp_rn(x);
#else
    return 1.0f / x;
#endif
}

ARTB_HD float fast_logf(float x)
{
#if defined(__CUDA_ARCH__)
    return __logf(x);
#else
    return logf(x);
#endif
}

ARTB_HD uint4 ld_node16(const uint4* p)
{
    return *p;
}

// PRMT in its default mode: result byte i = byte (selector nibble i) of the 8-byte value {b:a}
ARTB_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;      // prmt.b32 itself: the __byte_perm intrinsic ignores bit 3 of the selector nibbles (sign replication)
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
#else
    const uint64_t v = (uint64_t(b) << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t n = (sel >> (4 * i)) & 0xfu;
        uint32_t byte = uint32_t(v >> (8 * (n & 7u))) & 0xffu;
        if (n & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
        r |= byte << (8 * i);
    }
    return r;
#endif
}

// The same with the selector as a compile-time constant (a PTX immediate, encoded in the instruction). With "r"
// operands for everything ptxas copied selector and constant into fresh registers before every PRMT: 48 extra moves per
// 8-wide node test.
template <uint32_t SEL>
ARTB_HD uint32_t byte_perm_sel(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "n"(SEL));
    return r;
#else
    return byte_perm(a, b, SEL);
#endif
}

// every byte becomes 0xff if its top bit is set, else 0x00
ARTB_HD uint32_t sign_extend_bytes(uint32_t x) { return byte_perm(x, 0u, 0xba98u); }

ARTB_HD int find_msb(uint32_t x)   // x != 0
{
#if defined(__CUDA_ARCH__)
    return 31 - __clz(int(x));
#else
    return 31 - __builtin_clz(x);
#endif
}

ARTB_HD int pop_count(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// warp votes: on the host (hostsim replays one ray at a time) a "warp" is the single calling lane
ARTB_HD uint32_t warp_active_mask()
{
#if defined(__CUDA_ARCH__)
    return __activemask();
#else
    return 1u;
#endif
}

ARTB_HD uint32_t warp_ballot(uint32_t mask, bool pred)
{
#if defined(__CUDA_ARCH__)
    return __ballot_sync(mask, pred);
#else
    (void)mask;
    return pred ? 1u : 0u;
#endif
}

// Re-convergence point for the whole warp (every lane must reach it). Without one, nvcc is free to duplicate the code
// after a divergent branch into both arms, and the two groups of lanes then run the rest of the step one after the
// other at half the lane efficiency (measured: 12 of 32 lanes in the node test before this was added).
ARTB_HD void warp_sync()
{
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
}

ARTB_HD void warp_sync_mask(uint32_t mask)
{
#if defined(__CUDA_ARCH__)
    __syncwarp(mask);
#else
    (void)mask;
#endif
}

ARTB_HD void prefetch_l1(const void* p)
{
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
#else
    (void)p;
#endif
}

ARTB_HD void prefetch_l2(const void* p)
{
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
#else
    (void)p;
#endif
}

} // namespace artb

#endif
