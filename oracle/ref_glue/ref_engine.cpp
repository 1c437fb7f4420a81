// TEST INFRASTRUCTURE (oracle) — never linked into or called by the product.
//
// ref_engine: times the UNMODIFIED reference renderer, engine<W,H,3>::run (src/engine/engine.h:30-54),
// exactly as src/main.cpp:29-46 drives it. W, H and spp are compile-time constants of the reference
// (engine.h:18, tracer_constants.h:6-13), so one binary is built per configuration (see ../Makefile).
//   ref_engine ALIAS MODE OUT [BURN]      MODE = single | adaptive | stripes | images
// Prints one JSON line with the engine's own elapsed milliseconds (engine.h:91-93) and writes raw RGB8.
#include "scene_manager.cpp"   // the reference TU

#include "ref_scenes.h"

#include "color.h"
#include "engine.h"

#include <cstdio>
#include <fstream>
#include <vector>

namespace tc = tracer_constants;

int main(int argc, char** argv) try
{
    if (argc < 4) { std::fprintf(stderr, "usage: ref_engine ALIAS single|adaptive|stripes|images OUT [BURN]\n"); return 2; }
    const int alias = std::atoi(argv[1]);
    const std::string mode = argv[2];
    engine_mode m = engine_mode::single;
    if (mode == "adaptive") m = engine_mode::adaptive;
    else if (mode == "stripes") m = engine_mode::parallel_stripes;
    else if (mode == "images") m = engine_mode::parallel_images;

    scene world = build_scene(alias);
    const long burn = argc >= 5 ? std::atol(argv[4]) : 0;
    for (long i = 0; i < burn; ++i) random_double();
    camera cam = make_camera(world, tc::image_width, tc::image_height);

    std::vector<std::uint8_t> output_image(size_t(tc::frame_size));
    engine<tc::image_width, tc::image_height, tc::color_channels> eng(cam, m);
    eng.set_scene(world.objects, world.background);
    const int elapsed_ms = eng.run(output_image.data());

    std::ofstream f(argv[3], std::ios::binary);
    f.write(reinterpret_cast<const char*>(output_image.data()), std::streamsize(output_image.size()));
    std::printf("\n{\"alias\": %d, \"mode\": \"%s\", \"w\": %d, \"h\": %d, \"spp\": %d, \"elapsed_ms\": %d}\n",
                alias, mode.c_str(), tc::image_width, tc::image_height, tc::samples_per_pixel, elapsed_ms);
    return elapsed_ms < 0 ? 1 : 0;
}
catch (const std::exception& e)
{
    std::fprintf(stderr, "ref_engine: %s\n", e.what());
    return 1;
}
