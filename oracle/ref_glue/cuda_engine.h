// cuda_engine.h — the class a maintainer of blackccpie/another_raytracer adds next to `engine` (src/engine/engine.h:18-54)
// to render on the B200 backend: same three-call contract (constructor, set_scene, run), W / H / spp at run time.
// This is the binding INTEGRATION.md shows; it is compiled here against the UNMODIFIED reference (oracle/Makefile,
// target ref_cuda_engine) so that the text in INTEGRATION.md is known to build and run. It lives under oracle/ref_glue
// because it needs the reference's headers (camera.h, hittable_list.h, vec3.h), which are not part of this repository.
//
// Include after the reference's scene headers (include/artb_reference_adapter.hpp names the reference's classes).
#ifndef ARTB_REF_CUDA_ENGINE_H
#define ARTB_REF_CUDA_ENGINE_H

#include "artb.h"
#include "artb_reference_adapter.hpp"

#include <cstdint>
#include <iostream>
#include <stdexcept>

class cuda_engine
{
public:
    // camera arguments: exactly what src/main.cpp:33-35 passes to `camera`
    cuda_engine(int width, int height, int samples_per_pixel, int max_depth, point3 lookfrom, point3 lookat, vec3 vup, double vfov,
                double aperture, double focus_dist, double time0, double time1, engine_mode mode = engine_mode::single,
                int n_devices = 1)
        : W(width), H(height), SPP(samples_per_pixel), depth_(max_depth), mode_(mode)
    {
        cam_ = artb_camera{ { lookfrom.x(), lookfrom.y(), lookfrom.z() }, { lookat.x(), lookat.y(), lookat.z() },
                            { vup.x(), vup.y(), vup.z() }, vfov, double(width) / double(height), aperture, focus_dist, time0, time1 };
        artb_config cfg{};
        cfg.device = 0;
        // the reference's parallel modes use 4 worker threads (engine.h:335-445); here a worker is a GPU
        if (n_devices > 1 && (mode == engine_mode::parallel_images || mode == engine_mode::parallel_stripes)) {
            cfg.n_devices = n_devices;
            cfg.partition = mode == engine_mode::parallel_stripes ? ARTB_PARTITION_TILES : ARTB_PARTITION_SAMPLES;
            for (int k = 0; k < n_devices && k < ARTB_MAX_DEVICES; ++k) cfg.devices[k] = k;
        }
        if (artb_create(&cfg, &ctx_) != ARTB_OK) throw std::runtime_error(artb_last_error(nullptr));
    }
    ~cuda_engine()
    {
        if (scene_) artb_scene_free(scene_);
        artb_destroy(ctx_);
    }
    cuda_engine(const cuda_engine&) = delete;
    cuda_engine& operator=(const cuda_engine&) = delete;

    void set_scene(const hittable_list& world, const color& background)        // engine.h:24-28
    {
        if (scene_) { artb_scene_free(scene_); scene_ = nullptr; }
        empty_ = world.objects.empty();
        if (empty_) return;
        scene_ = artb_ref::make_scene(world, background);                       // the reference's OWN graph, serialised and flattened
        if (!scene_) throw std::runtime_error(artb_last_error(nullptr));
        if (artb_upload_scene(ctx_, scene_) != ARTB_OK) throw std::runtime_error(artb_last_error(ctx_));
    }

    int run(std::uint8_t* output_image)                                          // engine.h:30-54
    {
        if (empty_) { std::cerr << "Invalid input scene!" << std::endl; return -1; }
        artb_render_params p{};
        p.width = W; p.height = H; p.samples_per_pixel = SPP; p.max_depth = depth_;
        p.seed = 1;
        p.mode = mode_ == engine_mode::adaptive ? ARTB_MODE_ADAPTIVE_SQUARES : ARTB_MODE_FIXED;
        artb_stats st{};
        const int rc = artb_render(ctx_, &cam_, &p, output_image, nullptr, nullptr, &st);
        if (rc == ARTB_E_ADAPTIVE_SIZE) throw std::logic_error(artb_last_error(ctx_));       // engine.h:182-183
        if (rc != ARTB_OK) throw std::runtime_error(artb_last_error(ctx_));
        last_ = st;
        return static_cast<int>(st.elapsed_ms);                                  // what engine::run returns (engine.h:91-93)
    }

    const artb_stats& last_stats() const { return last_; }

private:
    int W, H, SPP, depth_;
    engine_mode mode_;
    artb_camera cam_{};
    artb_ctx* ctx_ = nullptr;
    artb_scene* scene_ = nullptr;
    artb_stats last_{};
    bool empty_ = true;
};

#endif
