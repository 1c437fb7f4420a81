// TEST INFRASTRUCTURE (reference-side integration proof). The reference's own main (src/main.cpp:29-46) with `engine`
// replaced by `cuda_engine` (cuda_engine.h): the UNMODIFIED reference builds the scene with its scene_manager, the adapter
// hands ITS graph to libartb.so, the B200 renders. Built by oracle/Makefile (needs /root/reference), linked against
// ../another_raytracer_b200/libartb.so; run by tests/test_gpu_reference_host.py on the GPU box.
//   ref_cuda_engine ALIAS W H SPP MODE OUT.raw [N_DEVICES]      MODE = single | adaptive | stripes | images
#include "scene_manager.cpp"   // the reference TU

#include "ref_scenes.h"

#include "color.h"             // engine.h needs write_color declared first (as in ref_engine.cpp)
#include "engine.h"            // engine_mode (engine.h:10-16)
#include "cuda_engine.h"

#include <cstdio>
#include <fstream>
#include <vector>

int main(int argc, char** argv) try
{
    if (argc < 7) { std::fprintf(stderr, "usage: ref_cuda_engine ALIAS W H SPP single|adaptive|stripes|images OUT [N_DEVICES]\n"); return 2; }
    const int alias = std::atoi(argv[1]), W = std::atoi(argv[2]), H = std::atoi(argv[3]), spp = std::atoi(argv[4]);
    const std::string mode = argv[5];
    engine_mode m = engine_mode::single;
    if (mode == "adaptive") m = engine_mode::adaptive;
    else if (mode == "stripes") m = engine_mode::parallel_stripes;
    else if (mode == "images") m = engine_mode::parallel_images;
    const int n_devices = argc >= 8 ? std::atoi(argv[7]) : 1;

    scene world = build_scene(alias);                                            // scene_manager::build of the reference
    cuda_engine eng(W, H, spp, tracer_constants::max_depth, world.lookfrom, world.lookat, vec3(0, 1, 0), world.vfov, world.aperture,
                    10.0, 0.0, 1.0, m, n_devices);                               // main.cpp:33-35
    eng.set_scene(world.objects, world.background);
    std::vector<std::uint8_t> output_image(size_t(W) * size_t(H) * 3);
    const int elapsed_ms = eng.run(output_image.data());

    std::ofstream f(argv[6], std::ios::binary);
    f.write(reinterpret_cast<const char*>(output_image.data()), std::streamsize(output_image.size()));
    const artb_stats& st = eng.last_stats();
    std::printf("\n{\"alias\": %d, \"mode\": \"%s\", \"w\": %d, \"h\": %d, \"spp\": %d, \"elapsed_ms\": %d, \"samples\": %llu, \"rays\": %llu, \"n_devices\": %d}\n",
                alias, mode.c_str(), W, H, spp, elapsed_ms, (unsigned long long)st.samples, (unsigned long long)st.rays, st.n_devices);
    return elapsed_ms < 0 ? 1 : 0;
}
catch (const std::exception& e)
{
    std::fprintf(stderr, "ref_cuda_engine: %s\n", e.what());
    return 1;
}
