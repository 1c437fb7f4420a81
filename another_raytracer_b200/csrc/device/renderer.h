// renderer.h — host-callable interface of the CUDA side (implemented in kernels.cu, compiled by nvcc).
// Everything above it (C-ABI, scene builders, flattener) is plain C++ compiled by g++.
#ifndef ARTB_RENDERER_H
#define ARTB_RENDERER_H

#include "flat_scene.h"
#include "../host/flatten.h"

#include <cstdint>
#include <string>

namespace artb {

struct RenderParams {
    int width = 0, height = 0;
    int samples_per_pixel = 1;
    int max_depth = 50;
    uint32_t seed = 1;
    uint32_t stream = 0;              // Philox key word 1: one stream per rank
    // partition A (sample split): this call renders per-pixel sample indices [sample_begin, sample_begin+spp)
    int sample_begin = 0;
    // partition B (tile split): only tiles with ((ty*tiles_x+tx) % tile_parts) == tile_part are rendered
    int tile_size = 32, tile_parts = 1, tile_part = 0;
    // adaptive per-pixel variance mask (0 = off): after min_spp samples a pixel stops once the standard
    // error of its luminance mean drops below rel_threshold * (mean + 0.05)
    int adaptive = 0;                 // 1: variance mask; 2: the reference's 12/6/3 squares (render_host only)
    int adaptive_min_spp = 16;
    int adaptive_batch_spp = 8;
    double adaptive_threshold = 0.02;
    int accumulate = 0;               // 0: accum/count buffers are zeroed first
    const uint32_t* pixel_list_dev = nullptr;   // internal: render exactly these pixels (device pointer) ...
    uint32_t n_pixel_list = 0;                  // ... this many
};

struct RenderStats {
    double elapsed_ms = 0;            // CUDA-event time of the whole render (the reference's run() return value)
    unsigned long long samples = 0;   // primary paths traced
    unsigned long long rays = 0;      // closest-hit queries (engine.h:455)
    unsigned long long waves = 0;
    unsigned long long launches = 0;  // kernels launched
    double extend_ms = 0;             // profile=N: every N-th wave timed; the closest-hit stage (k_traverse + k_pairs + k_finish + k_extend), per-wave events
    double traverse_ms = 0, pairs_ms = 0, finish_ms = 0;
    unsigned long long overflow_rays = 0;   // rays whose candidate list overflowed (traced by the one-phase pass)
    unsigned long long extend_launches = 0;
    unsigned long long pixels_active = 0, pixels_converged = 0;
};

struct BvhBuildStats {
    double total_ms = 0;              // CUDA-event time of the whole device build
    uint32_t n_prims = 0, n_nodes = 0;
    int depth = 0;                    // levels of 8-wide nodes
};

class Renderer {
public:
    Renderer(int device, uint32_t pool_paths);
    ~Renderer();
    Renderer(const Renderer&) = delete;
    Renderer& operator=(const Renderer&) = delete;

    void upload(const FlatScene& scene);
    bool has_scene() const { return has_scene_; }
    uint32_t n_media() const { return view_.n_media; }
    uint64_t scene_bytes() const { return scene_bytes_; }
    uint64_t last_trace_overflow() const { return last_trace_overflow_; }   // rays of the last trace_batch that took the one-phase pass

    // (Re)build the 8-wide BVH of the uploaded scene ON THE DEVICE (bvh_build.cuh): replaces the node array and puts the
    // tree primitives into the new leaf order. upload() calls it by itself for a flat scene that carries no nodes.
    void build_bvh_device(int max_leaf_prims, BvhBuildStats* stats);
    const BvhBuildStats& last_build() const { return last_build_; }

    // parity hook: closest hit for explicit rays. rays = n*7 doubles (o, d, time), xi = n*n_media or null.
    void trace_batch(const double* rays_host, uint64_t n, double t_min, double t_max, const double* xi_host,
                     HitRecordOut* out_host);
    // same with device-resident buffers; returns the kernel time in ms
    float trace_batch_device(const double* rays_dev, uint64_t n, double t_min, double t_max, const double* xi_dev,
                             HitRecordOut* out_dev);

    // parity hook for the shade stage: hit record + material::emitted / scatter + texture::value for explicit hits
    // (prim_id = depth-first leaf id and t as trace_batch reports them) with explicit random numbers: xi = n*4 doubles,
    // consumed in the order the device samplers draw them. Host pointers.
    void shade_batch(const double* rays_host, const int32_t* prim_id_host, const double* t_host, const double* xi_host,
                     uint64_t n, ShadeRecordOut* out_host);
    // parity hook for _stochastic_sample + camera::get_ray: st = n*2 image-plane coordinates (s, t), xi = n*3 doubles
    // (lens disk radius, lens disk angle, shutter time); rays_out = n*7 doubles. Host pointers.
    void camera_rays(const CameraParams& cam, const double* st_host, const double* xi_host, uint64_t n, double* rays_out_host);

    // Wavefront render into a device accumulation buffer (W*H*3 float sums) + per-pixel sample counts.
    void render_device(const CameraParams& cam, const RenderParams& p, float* accum_dev, uint32_t* count_dev,
                       RenderStats* stats, int profile_every);
    // engine_mode::adaptive of the reference (engine.h:151-333): corners of 12/6/3-pixel squares traced at full spp,
    // the rest interpolated. Device buffers: rgb W*H*3, accum W*H*3 sums of the traced pixels, count = spp where traced.
    // peers (optional): the pixels each level asks for are split evenly over this device and the peers; their frames are
    // summed into accum_dev over peer memory before the level's pixels are committed.
    void render_adaptive_squares(const CameraParams& cam, const RenderParams& p, uint8_t* rgb_dev, float* accum_dev,
                                 uint32_t* count_dev, RenderStats* stats, Renderer* const* peers = nullptr, int n_peers = 0);
    // write_color (color.h:6-22) over accum/count -> RGB8 (device buffers)
    void resolve_device(const float* accum_dev, const uint32_t* count_dev, int n_pixels, int fixed_spp, uint8_t* rgb_dev);

    // convenience: host-buffer render (allocates scratch, copies results back)
    void render_host(const CameraParams& cam, const RenderParams& p, uint8_t* rgb_out, float* accum_out,
                     uint32_t* count_out, RenderStats* stats);

    // ---- several GPUs in one process (artb_config::n_devices): engine_mode::parallel_images / parallel_stripes of the
    // reference (engine.h:335-445: 4 workers, a final sum of their frames) with one B200 per worker.
    // This renderer (device 0 of the group) and `peers` each render their share — partition 0: disjoint per-pixel
    // sample ranges on their own Philox streams; partition 1: interleaved tiles — from their own host thread into their own
    // frame; then ONE kernel on this device sums the peers' frames straight out of their memory over NVLink (peer loads;
    // staged copies where peer access is unavailable), resolves and copies back. per_device (optional): one RenderStats
    // per device, this one first.
    void render_host_multi(Renderer* const* peers, int n_peers, int partition, const CameraParams& cam, const RenderParams& p,
                           uint8_t* rgb_out, float* accum_out, uint32_t* count_out, RenderStats* stats, RenderStats* per_device,
                           double* reduce_ms);
    // grow-only device frame (W*H*3 float sums, W*H counts, W*H*3 bytes) of render_host / render_host_multi
    void ensure_frame(size_t n_pixels, float** accum, uint32_t** count, uint8_t** rgb);

    void* stream() const { return stream_; }
    int device() const { return device_; }

private:
    void init(uint32_t pool_paths);
    void release();
    void sync_view();
    struct Impl;
    Impl* impl_;
    int device_;
    void* stream_ = nullptr;
    bool has_scene_ = false;
    SceneView view_{};
    uint64_t scene_bytes_ = 0;
    uint64_t last_trace_overflow_ = 0;
    uint32_t max_dfs_id_ = 0;         // largest depth-first leaf id among the world primitives (shade_batch maps ids back)
    BvhBuildStats last_build_{};
};

} // namespace artb

#endif
