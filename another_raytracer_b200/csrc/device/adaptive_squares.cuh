// adaptive_squares.cuh — engine_mode::adaptive of the reference (engine.h:151-333) on the device.
//
// The reference renders the four CORNER pixels of every 12x12 square at full spp, looks at their 8-bit colours
// (_compute_corners_heuristic, engine.h:96-137: any edge of the corner quadrilateral with a squared RGB distance above 100
// asks for a finer look), and then either fills the square by bilinear interpolation of the corner colours
// (_interpolate, engine.h:139-149; integer colours, no gamma: write_color_raw) or repeats the procedure on its four
// 6x6 squares, then on their 3x3 squares, where the five remaining pixels are finally traced.
//
// Here each level is one pass: a pixel list (corners not traced yet) rendered by the wavefront kernels at full spp,
// one kernel that turns their radiance sums into the integer work image (write_color<int>, color.h:6-22), and one
// kernel per square that applies the heuristic and either interpolates or emits the next level's squares and pixels.
// Differences from the reference, on purpose: a corner that a coarser level already traced is NOT traced again (the
// reference redraws it with fresh random numbers and overwrites it — same distribution, wasted work), and pixels are
// keyed by (pixel, sample) in Philox, so a traced pixel has exactly the value it has in a full render.
// tests/test_gpu_adaptive_squares.py replays engine.h:151-333 in numpy on a full render and compares.
#ifndef ARTB_ADAPTIVE_SQUARES_CUH
#define ARTB_ADAPTIVE_SQUARES_CUH

#include <stdint.h>

namespace artb {
namespace asq {

struct Lists {
    uint32_t* squares[2];      // packed (x | y << 16) of the squares of the current / next level
    uint32_t* pixels;          // pixel indices to trace for the next level
    uint32_t* counters;        // [0] squares of the next level, [1] pixels to trace
};

__device__ __forceinline__ void write_color_int(int* out, const float* sum, int spp)
{
    // color.h:6-22 with T = int
    const double scale = 1.0 / spp;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double v = sqrt(scale * double(sum[k]));
        v = v < 0.0 ? 0.0 : (v > 0.999 ? 0.999 : v);
        out[k] = int(256 * v);
    }
}

__global__ void k_sq_init(int* work, uint32_t* count, int n_pixels)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    work[3 * i] = work[3 * i + 1] = work[3 * i + 2] = -1;       // frame_allocator::get_frame(0, -1), engine.h:173
    count[i] = 0;
}

// level 0: every 12x12 square and its four corner pixels
__global__ void k_sq_level0(Lists L, int width, int height, int size)
{
    const int sx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nx = width / size, ny = height / size;
    if (sx >= nx * ny) return;
    const int x = (sx % nx) * size, y = (sx / nx) * size;
    L.squares[0][sx] = uint32_t(x) | (uint32_t(y) << 16);
    const uint32_t base = uint32_t(sx) * 4u;
    L.pixels[base + 0] = uint32_t(y * width + x);
    L.pixels[base + 1] = uint32_t(y * width + x + size - 1);
    L.pixels[base + 2] = uint32_t((y + size - 1) * width + x);
    L.pixels[base + 3] = uint32_t((y + size - 1) * width + x + size - 1);
}

// the traced pixels of a level: radiance sums -> integer work image
__global__ void k_sq_commit(const uint32_t* pixels, uint32_t n, const float* accum, int spp, int* work, uint32_t* count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = pixels[i];
    write_color_int(work + 3 * size_t(p), accum + 3 * size_t(p), spp);
    count[p] = uint32_t(spp);
}

__device__ __forceinline__ bool corners_differ(const int* work, int width, int x, int y, int size)
{
    // _compute_corners_heuristic, engine.h:96-137 (corners 1 up-left, 2 up-right, 3 bottom-left, 4 bottom-right)
    const int* c1 = work + 3 * (y * width + x);
    const int* c2 = work + 3 * (y * width + x + size - 1);
    const int* c3 = work + 3 * ((y + size - 1) * width + x);
    const int* c4 = work + 3 * ((y + size - 1) * width + x + size - 1);
    auto d2 = [](const int* a, const int* b) {
        return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
    };
    return d2(c1, c2) > 100 || d2(c2, c4) > 100 || d2(c4, c3) > 100 || d2(c3, c1) > 100;
}

__device__ __forceinline__ void interpolate_square(int* work, int width, int x, int y, int size)
{
    // interpolate_square + _interpolate, engine.h:186-219, 139-149: colours as doubles, results truncated to int,
    // pixels that already hold a value are kept
    const int x1 = x, x2 = x + size - 1, y1 = y, y2 = y + size - 1;
    double q11[3], q12[3], q21[3], q22[3];
    for (int k = 0; k < 3; ++k) {
        q11[k] = double(work[3 * (y1 * width + x1) + k]);
        q12[k] = double(work[3 * (y2 * width + x1) + k]);
        q21[k] = double(work[3 * (y1 * width + x2) + k]);
        q22[k] = double(work[3 * (y2 * width + x2) + k]);
    }
    const int xdiff = x2 - x1, ydiff = y2 - y1;
    for (int l = 0; l < size; ++l)
        for (int k = 0; k < size; ++k) {
            int* px = work + 3 * ((y + l) * width + x + k);
            if (px[0] >= 0) continue;
            const int xx = x + k, yy = y + l;
            for (int ch = 0; ch < 3; ++ch) {
                // vec3 arithmetic of the reference: (x2-x)*Q/xdiff means (double(x2-x) * Q) * (1/xdiff) (vec3.h operator/)
                const double r1 = (double(x2 - xx) * q11[ch]) * (1.0 / xdiff) + (double(xx - x1) * q21[ch]) * (1.0 / xdiff);
                const double r2 = (double(x2 - xx) * q12[ch]) * (1.0 / xdiff) + (double(xx - x1) * q22[ch]) * (1.0 / xdiff);
                px[ch] = int((double(y2 - yy) * r1) * (1.0 / ydiff) + (double(yy - y1) * r2) * (1.0 / ydiff));
            }
        }
}

// one thread per square of the current level: interpolate it, or hand its sub-squares (and the pixels they still
// need) to the next level. size 12 and 6 split into four squares of half the size; size 3 asks for its five
// remaining pixels (engine.h:272-283).
__global__ void k_sq_decide(Lists L, int level_in, uint32_t n_squares, int size, int* work, int width)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_squares) return;
    const uint32_t packed = L.squares[level_in][i];
    const int x = int(packed & 0xffffu), y = int(packed >> 16);
    if (!corners_differ(work, width, x, y, size)) {
        interpolate_square(work, width, x, y, size);
        return;
    }
    if (size == 3) {
        const uint32_t base = atomicAdd(L.counters + 1, 5u);
        L.pixels[base + 0] = uint32_t(y * width + x + 1);
        L.pixels[base + 1] = uint32_t((y + 1) * width + x);
        L.pixels[base + 2] = uint32_t((y + 1) * width + x + 1);
        L.pixels[base + 3] = uint32_t((y + 1) * width + x + 2);
        L.pixels[base + 4] = uint32_t((y + 2) * width + x + 1);
        return;
    }
    const int half = size / 2;
    const uint32_t sbase = atomicAdd(L.counters + 0, 4u);
    uint32_t todo[16];
    uint32_t n_todo = 0;
    for (int q = 0; q < 4; ++q) {
        const int sx = x + (q & 1) * half, sy = y + (q >> 1) * half;
        L.squares[level_in ^ 1][sbase + uint32_t(q)] = uint32_t(sx) | (uint32_t(sy) << 16);
        const int cx[4] = { sx, sx + half - 1, sx, sx + half - 1 };
        const int cy[4] = { sy, sy, sy + half - 1, sy + half - 1 };
        for (int c = 0; c < 4; ++c) {
            const uint32_t p = uint32_t(cy[c] * width + cx[c]);
            if (work[3 * size_t(p)] < 0) todo[n_todo++] = p;       // corners of the parent are already there
        }
    }
    const uint32_t pbase = atomicAdd(L.counters + 1, n_todo);
    for (uint32_t k = 0; k < n_todo; ++k) L.pixels[pbase + k] = todo[k];
}

__global__ void k_sq_final(const int* work, int n_values, uint8_t* rgb)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_values) rgb[i] = uint8_t(work[i]);        // std::transform ... static_cast<std::uint8_t>, engine.h:326-327
}

} // namespace asq
} // namespace artb

#endif
