/* artb.h — C ABI of the B200 path-tracing backend for blackccpie/another_raytracer.
 *
 * The reference has no plugin / FFI interface; the seam this library sits behind is the three-method
 * contract of its render driver (reference: src/engine/engine.h:18-54, used by src/main.cpp:44-46):
 *
 *     engine<W,H,C> eng(cam, mode);  eng.set_scene(world.objects, world.background);  ms = eng.run(out);
 *
 * Each entry point below cites the reference interface it replaces. Conventions: plain pointers and sizes
 * only; every function returns 0 on success or a negative artb_status and never throws across the ABI;
 * artb_last_error() gives the message; buffers are caller-owned; calls are blocking; one context per host
 * thread. There is NO CPU fallback: artb_create fails with ARTB_E_NO_DEVICE when no CUDA device exists.
 */
#ifndef ARTB_H
#define ARTB_H

#include <stddef.h>
#include <stdint.h>

#include "artb_flat.h"
#include "artb_tokens.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ARTB_ABI_VERSION 2     /* 2: artb_config carries a device list; artb_shade_batch, artb_camera_rays, artb_scene_from_graph */

typedef enum artb_status {
    ARTB_OK = 0,
    ARTB_E_INVALID = -1,       /* bad argument */
    ARTB_E_NO_DEVICE = -2,     /* no CUDA device / CUDA runtime failure at start-up */
    ARTB_E_CUDA = -3,          /* CUDA error while running */
    ARTB_E_SCENE = -4,         /* scene cannot be built / flattened (unsupported graph, missing asset) */
    ARTB_E_EMPTY_SCENE = -5,   /* reference: engine::run returns -1 and prints "Invalid input scene!" (engine.h:32-36) */
    ARTB_E_ADAPTIVE_SIZE = -6  /* reference: std::logic_error when W or H is not a multiple of 12 (engine.h:182-183) */
} artb_status;

typedef struct artb_ctx artb_ctx;       /* one GPU + its path pool               */
typedef struct artb_scene artb_scene;   /* host scene graph + its flattened form */

/* ---- context ------------------------------------------------------------------------------------ */
#define ARTB_MAX_DEVICES 8
typedef enum artb_partition {
    ARTB_PARTITION_SAMPLES = 0,   /* every device renders all pixels with its share of the per-pixel samples on its own random
                                     stream: engine_mode::parallel_images (engine.h:378-445, a sum of the workers' frames) */
    ARTB_PARTITION_TILES = 1      /* interleaved tile_size x tile_size tiles, tile t on device t % n_devices: the load-balanced
                                     form of engine_mode::parallel_stripes (engine.h:335-376, one row stripe per worker) */
} artb_partition;

typedef struct artb_config {
    int32_t device;            /* CUDA device ordinal (used when n_devices <= 1) */
    uint32_t pool_paths;       /* paths in flight per device (0 = default 2 Mi) */
    /* several GPUs of one node behind ONE context: the scene is replicated by artb_upload_*, artb_render splits
       the work over the devices (one host thread each), sums the frames with one kernel that reads the peers' memory over
       NVLink, and resolves on devices[0]. Every other call (trace / shade / device variants) runs on devices[0].
       n_devices == 0 or 1: a single-device context on `device`. An ordinal may repeat (several path pools on one GPU). */
    int32_t n_devices;         /* 0..ARTB_MAX_DEVICES */
    int32_t partition;         /* artb_partition used by artb_render when n_devices > 1 */
    int32_t devices[ARTB_MAX_DEVICES];
} artb_config;

int artb_abi_version(void);
int artb_device_count(void);   /* CUDA devices visible to the process (0 without a GPU) */
int artb_create(const artb_config* cfg, artb_ctx** out);
int artb_context_devices(const artb_ctx* ctx);    /* devices behind this context (>= 1) */
void artb_destroy(artb_ctx* ctx);
/* message of the last failing call on this context (ctx may be NULL: last failure of a ctx-less call) */
const char* artb_last_error(const artb_ctx* ctx);

/* ---- host scene construction (replaces scene_manager::build, src/scene_manager.cpp:260-355) ----- */
typedef struct artb_scene_options {
    int32_t mesh_as_soup;      /* aliases 9-11: flatten the .obj as one triangle_soup (no per-triangle objects) */
    int32_t heightfield_side;  /* alias 12: vertices per side (0 = 708 -> 999 698 triangles) */
    int32_t max_leaf_prims;    /* device BVH leaf size (0 = default) */
    int32_t device_bvh;        /* 1: skip the host SAH build; artb_upload_scene builds the 8-wide tree on the GPU */
    const char* assets_root;   /* NULL: $ARTB_ASSETS or <repo>/assets */
} artb_scene_options;

typedef struct artb_scene_info {
    double lookfrom[3], lookat[3], vfov, aperture, background[3];   /* scene_manager.h:6-14 */
    double rng_next;           /* next value of the global RNG after the build (pins RNG consumption) */
    double world_lo[3], world_hi[3];
    uint64_t rng_draws;
    uint32_t n_prims, n_world_prims, n_nodes, n_media, n_materials, n_textures, n_images, n_xforms;
    uint32_t n_leaves_total, n_culled, bvh_depth, top_level_objects;
    uint64_t flat_bytes;
} artb_scene_info;

/* alias: 1..9 = the reference's scene_alias (scene_manager.h:16-27); 10 cow, 11 dino, 12 height field.
 * The global host RNG is reset to its default seed first, so every build sees the stream a fresh
 * reference process would. */
int artb_scene_build(int32_t alias, const artb_scene_options* opt, artb_scene** out);
void artb_scene_free(artb_scene* s);
int artb_scene_get_info(const artb_scene* s, artb_scene_info* out);
/* pre-order token stream of the host graph (artb_tokens.h). Returns the stream length; copies at most
 * `capacity` doubles into out (out may be NULL to query the length). */
int64_t artb_scene_dump_graph(const artb_scene* s, double* out, int64_t capacity);

/* ---- flattened scene as plain arrays (what a foreign host hands over) ----------------- */
typedef struct artb_flat_scene {         /* record types: artb_flat.h */
    const artb_flat_prim* prims;          uint32_t n_prims, n_world_prims;   /* globals, BVH leaves, then media boundaries */
    const artb_flat_node8* nodes;         uint32_t n_nodes; int32_t reserved0; uint32_t n_global_prims; int32_t bvh_depth;
                                          /* [0] = root; n_nodes == 0: the tree is built on the GPU; bvh_depth <= 14 levels */
    const artb_flat_xform_op* xform_ops;  uint32_t n_xform_ops;
    const artb_flat_xform* xforms;        uint32_t n_xforms;                 /* [0] = identity */
    const artb_flat_material* materials;  uint32_t n_materials;
    const artb_flat_texture* textures;    uint32_t n_textures;
    const artb_flat_image* images;        uint32_t n_images;
    const unsigned char* texels;          uint64_t n_texel_bytes;
    const artb_flat_perlin* perlins;      uint32_t n_perlins;
    uint32_t n_media;
    double background[3];
    const artb_flat_ref_node* ref_nodes;  uint32_t n_ref_nodes;              /* the reference's own bvh_nodes, used only
                                                                  to break exact-t ties as bvh_node::hit does (may be NULL/0:
                                                                  ties then go to the larger dfs_id); prim g[9] = 1 + index */
} artb_flat_scene;

/* A scene graph built by a FOREIGN host (the reference's own hittable_list): `tokens` is its pre-order serialisation
 * (artb_tokens.h; include/artb_reference_adapter.hpp writes it from the reference's classes, reference:
 * src/engine/hittable_list.h:9-24, src/primitives/bvh.h:9-26, ...), `blobs` the Perlin tables / image texels the stream
 * names by hash. The graph is rebuilt as it is (the reference's bvh_node shapes and boxes included: they decide culling
 * and exact-tie order) and flattened; nothing is drawn from the RNG. opt may be NULL (mesh_as_soup / heightfield_side /
 * assets_root are ignored). The result is an ordinary artb_scene: artb_upload_scene, or artb_scene_get_flat +
 * artb_upload_flat. lookfrom / lookat / vfov of artb_scene_info are zero: the caller owns the camera. */
typedef struct artb_graph_blob { uint64_t hash; const void* data; uint64_t bytes; } artb_graph_blob;
int artb_scene_from_graph(const double* tokens, int64_t n_tokens, const artb_graph_blob* blobs, int32_t n_blobs,
                          const double background[3], const artb_scene_options* opt, artb_scene** out);

/* view of the arrays owned by `s` (valid until artb_scene_free) */
int artb_scene_get_flat(const artb_scene* s, artb_flat_scene* out);

/* ---- engine::set_scene (engine.h:24-28): copy the scene to HBM ------------------------------------ */
int artb_upload_scene(artb_ctx* ctx, const artb_scene* s);
int artb_upload_flat(artb_ctx* ctx, const artb_flat_scene* flat);

/* ---- parity hook: hittable_list::hit on explicit rays (engine.h:455, hittable_list.cpp:5-19) ----- */
typedef struct artb_hit {
    double t, p[3], normal[3], u, v;     /* hit_record, hittable.h:9-16 */
    int32_t hit;                         /* 0 = miss */
    int32_t prim_id;                     /* leaf id in the reference's depth-first order, -1 on miss */
    int32_t front_face;
    int32_t material_type;               /* 0 lambertian 1 metal 2 dielectric 3 diffuse_light 4 isotropic */
} artb_hit;

/* rays: n x 7 doubles (origin, direction, time); medium_xi: n x n_media doubles, the random number each
 * constant_medium consumes for that ray (constant_medium.h:63), NaN where the reference drew none;
 * NULL = draw from the Philox stream. All pointers are HOST pointers. */
int artb_trace_batch(artb_ctx* ctx, const double* rays, uint64_t n, double t_min, double t_max,
                     const double* medium_xi, artb_hit* hits_out);
/* same on DEVICE pointers; *kernel_ms (optional) receives the CUDA-event time of the traversal kernel */
int artb_trace_batch_device(artb_ctx* ctx, const double* rays_dev, uint64_t n, double t_min, double t_max,
                            const double* medium_xi_dev, artb_hit* hits_dev, float* kernel_ms);
/* (Re)build the 8-wide BVH of the uploaded scene ON THE GPU (replaces the bvh_node constructor, bvh.cpp:3-42, for the
 * tree the kernels traverse; results do not depend on the tree). max_leaf_prims: 1..3 (0 = 1). */
typedef struct artb_bvh_stats {
    double build_ms;                    /* CUDA-event time of the device build */
    uint32_t n_prims, n_nodes;
    int32_t depth;                      /* levels of 8-wide nodes */
    int32_t reserved;
} artb_bvh_stats;
int artb_build_bvh_device(artb_ctx* ctx, int32_t max_leaf_prims, artb_bvh_stats* stats);
/* stats of the build artb_upload_scene / artb_upload_flat ran by itself (all zero if the scene came with a tree) */
int artb_last_bvh_build(const artb_ctx* ctx, artb_bvh_stats* stats);

/* how many rays of the last artb_trace_batch* call took the one-phase overflow pass instead of the pair kernels
 * (0 for the reference's scenes; the tests assert it so that parity is known to cover the kernels the renderer runs) */
uint64_t artb_trace_overflow_rays(const artb_ctx* ctx);

/* ---- parity hooks for the shade stage ------------------------------------------------------------------
 * material::emitted / material::scatter / texture::value (material.h:10-135, texture.h:16-154, perlin.h:21-96) and
 * camera::get_ray (camera.h:38-47) are deterministic once the random numbers they consume are given. These two
 * calls run the device functions the render kernels run, on explicit inputs with EXPLICIT random numbers, so the
 * shade stage can be compared with the oracle value by value (like artb_trace_batch does for the hit stage). The
 * device draws unit-sphere / unit-ball / unit-disk samples in closed form (same distributions as the reference's
 * rejection loops, vec3.h:117-143); xi are consumed in this order:
 *   lambertian  z = 1-2*xi0, phi = 2*pi*xi1                     metal / isotropic  the same + radius cbrt(xi2)
 *   dielectric  xi0 against the Schlick reflectance (only when refraction is possible)
 *   camera      lens radius sqrt(xi0), lens angle 2*pi*xi1, time = time0 + (time1-time0)*xi2                     */
typedef struct artb_shade_record {
    double emitted[3];                  /* material::emitted(u, v, p) */
    double attenuation[3];              /* scatter()'s attenuation */
    double origin[3], direction[3], time;   /* scatter()'s scattered ray */
    int32_t scattered;                  /* scatter()'s return value; -1: prim_id does not name a primitive */
    int32_t material_type, front_face, reserved;
} artb_shade_record;

/* rays: n x 7 doubles; prim_id / t: the hit as artb_trace_batch reported it; xi: n x 4 doubles. HOST pointers. */
int artb_shade_batch(artb_ctx* ctx, const double* rays, const int32_t* prim_id, const double* t, const double* xi,
                     uint64_t n, artb_shade_record* out);
struct artb_camera;
/* st: n x 2 image-plane coordinates (s, t) as engine.h:62-63 computes them; xi: n x 3; rays_out: n x 7. HOST pointers. */
int artb_camera_rays(artb_ctx* ctx, const struct artb_camera* cam, const double* st, const double* xi, uint64_t n,
                     double* rays_out);

/* ---- engine::run (engine.h:30-54) ------------------------------------------------------------------ */
typedef struct artb_camera {            /* camera::camera arguments, camera.h:8-17 */
    double lookfrom[3], lookat[3], vup[3];
    double vfov, aspect_ratio, aperture, focus_dist, time0, time1;
} artb_camera;

typedef enum artb_mode {
    ARTB_MODE_FIXED = 0,                /* engine_mode::single / parallel_*: spp samples for every pixel */
    ARTB_MODE_VARIANCE_MASK = 1,        /* north_star (3): per-pixel variance mask, unconverged pixels re-queued */
    ARTB_MODE_ADAPTIVE_SQUARES = 2      /* engine_mode::adaptive as the reference does it (engine.h:151-333): corners of
                                           12/6/3-pixel squares at full spp, the rest interpolated; W, H multiples of 12
                                           (else ARTB_E_ADAPTIVE_SIZE); artb_render only */
} artb_mode;

typedef struct artb_render_params {
    int32_t width, height;              /* engine<W,H,3> template arguments */
    int32_t samples_per_pixel;          /* tracer_constants::samples_per_pixel; cap in variance-mask mode */
    int32_t max_depth;                  /* tracer_constants::max_depth (0 = 50) */
    uint32_t seed;
    uint32_t stream;                    /* Philox stream: use the rank for sample-split rendering */
    int32_t mode;                       /* artb_mode */
    int32_t sample_begin;               /* partition A: first per-pixel sample index of this call */
    int32_t tile_size, tile_parts, tile_part;   /* partition B: interleaved tiles (tile_parts<=1: whole image) */
    int32_t adaptive_min_spp, adaptive_batch_spp;
    double adaptive_threshold;          /* relative standard error of the pixel luminance mean */
    int32_t accumulate;                 /* device variant: add to the buffers instead of clearing them */
    int32_t profile;                    /* N > 0: time the kernels of every N-th wave with CUDA events (stats.extend_ms, extend_launches = waves timed) */
} artb_render_params;

typedef struct artb_stats {
    double elapsed_ms;                  /* what engine::run returns (engine.h:91-93), measured with CUDA events */
    uint64_t samples, rays, waves, launches;
    double extend_ms; uint64_t extend_launches;
    uint64_t pixels_active, pixels_converged;
    uint64_t h2d_bytes, d2h_bytes;
    double traverse_ms, pairs_ms, finish_ms;   /* profile=N: the kernels of the closest-hit stage in the timed waves (extend_ms = whole stage) */
    uint64_t overflow_rays;             /* rays traced by the one-phase overflow pass */
    /* contexts with several devices: elapsed_ms = slowest device + the exchange; counts are summed over the devices */
    int32_t n_devices, reserved;
    double reduce_ms;                   /* the frame-sum kernel over peer memory */
    double device_ms[ARTB_MAX_DEVICES]; /* CUDA-event time of every device's render */
    uint64_t device_samples[ARTB_MAX_DEVICES];
} artb_stats;

/* rgb8_out: W*H*3 bytes, row 0 = top, RGB (the layout engine::run writes). accum_out (optional):
 * W*H*3 float per-pixel radiance SUMS; count_out (optional): W*H samples per pixel. HOST pointers. */
int artb_render(artb_ctx* ctx, const artb_camera* cam, const artb_render_params* p, uint8_t* rgb8_out,
                float* accum_out, uint32_t* count_out, artb_stats* stats);
/* DEVICE pointers; no resolve. For multi-GPU: every rank renders its share, the sums are reduced, rank 0
 * resolves. count_dev may be NULL in ARTB_MODE_FIXED. */
int artb_render_device(artb_ctx* ctx, const artb_camera* cam, const artb_render_params* p, float* accum_dev,
                       uint32_t* count_dev, artb_stats* stats);
/* write_color (color.h:6-22) on DEVICE buffers: fixed_spp > 0 divides by it, else by count_dev */
int artb_resolve_device(artb_ctx* ctx, const float* accum_dev, const uint32_t* count_dev, int32_t n_pixels,
                        int32_t fixed_spp, uint8_t* rgb8_dev);
/* the CUDA stream the context launches on (cudaStream_t) */
void* artb_stream(artb_ctx* ctx);

/* ---- imageio::save_image / load_image (imageio.h:9-10) --------------------------------------------- */
int artb_save_png(const char* path, int32_t width, int32_t height, int32_t channels, const uint8_t* data);

/* ---- host RNG (tracer_utils.h:27-41), exposed so tests can pin it -------------------------------- */
void artb_host_rng_reset(uint32_t seed);
double artb_host_rng_next(void);

#ifdef __cplusplus
}
#endif

#endif
