// TEST INFRASTRUCTURE (oracle). Force-included (-include cfg.h) ahead of every reference TU so that
// the reference's compile-time configuration (src/core/tracer_constants.h:4-15) can be chosen per
// binary without touching the reference sources. Uses the same include guard as the reference header,
// so the later quoted #include "tracer_constants.h" becomes a no-op.
#ifndef TRACER_CONSTANTS_H
#define TRACER_CONSTANTS_H

#ifndef CFG_W
#define CFG_W 720
#endif
#ifndef CFG_H
#define CFG_H 540
#endif
#ifndef CFG_SPP
#define CFG_SPP 100
#endif

namespace tracer_constants
{
    constexpr int image_width = CFG_W;
    constexpr int image_height = CFG_H;
    constexpr auto aspect_ratio = double(CFG_W) / double(CFG_H);
    constexpr int color_channels = 3;
    constexpr int frame_size = image_width * image_height * color_channels;
    constexpr bool stack_alloc = false;
    constexpr int samples_per_pixel = CFG_SPP;
    constexpr int max_depth = 50;
    constexpr bool progress_gui = false;
}

#endif
