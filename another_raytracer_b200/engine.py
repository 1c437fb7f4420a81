"""Host-side mirror of the reference's render seam on top of the C ABI.

Reference: `engine<W,H,C>{set_scene, run}` + `engine_mode` (src/engine/engine.h:10-54), `scene_manager::build`
(src/scene_manager.h:29-44), `camera` (src/engine/camera.h:8-17). Same names, argument meaning and error behaviour:
`Engine.run` returns the elapsed milliseconds, or -1 after printing "Invalid input scene!" when no scene is set.
All work is done by libartb.so on the GPU; numpy/torch only carry buffers.
"""
from __future__ import annotations

import ctypes as C
import enum
import sys
from dataclasses import dataclass, field

import numpy as np

from . import _abi

HIT_DTYPE = np.dtype([("t", "<f8"), ("p", "<f8", 3), ("normal", "<f8", 3), ("u", "<f8"), ("v", "<f8"),
                      ("hit", "<i4"), ("prim_id", "<i4"), ("front_face", "<i4"), ("material_type", "<i4")])
assert HIT_DTYPE.itemsize == C.sizeof(_abi.Hit)
SHADE_DTYPE = np.dtype([("emitted", "<f8", 3), ("attenuation", "<f8", 3), ("origin", "<f8", 3), ("direction", "<f8", 3),
                        ("time", "<f8"), ("scattered", "<i4"), ("material_type", "<i4"), ("front_face", "<i4"), ("reserved", "<i4")])
assert SHADE_DTYPE.itemsize == C.sizeof(_abi.ShadeRecord)

MATERIAL_NAMES = ["lambertian", "metal", "dielectric", "diffuse_light", "isotropic"]


class ArtbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"artb error {code}: {message}")
        self.code = code


class SceneAlias(enum.IntEnum):
    """scene_alias of the reference (scene_manager.h:16-27) + the added mesh / synthetic scenes."""
    random = 1
    two_spheres = 2
    two_perlin_spheres = 3
    earth = 4
    simple_light = 5
    cornell_box = 6
    cornell_smoke = 7
    final = 8
    mesh = 9
    cow = 10
    dino = 11
    heightfield = 12


class EngineMode(enum.Enum):
    """engine_mode of the reference (engine.h:10-16). On the GPU `single`, `parallel_stripes` and
    `parallel_images` are the same fixed-spp wavefront render; `adaptive` is the reference's own strategy (corners of
    12/6/3-pixel squares traced, the rest interpolated: engine.h:151-333); `variance_mask` is this backend's extension
    (north_star 3: per-pixel variance mask that re-queues only unconverged pixels)."""
    single = 0
    adaptive = 1
    parallel_stripes = 2
    parallel_images = 3
    variance_mask = 4


def _check(rc: int, ctx=None):
    if rc != 0:
        msg = _abi.load().artb_last_error(ctx)
        raise ArtbError(rc, (msg or b"").decode(errors="replace"))


def read_graph_file(path):
    """(tokens, {hash: bytes}, background) of an exported scene graph: int64 header {'ARTG', 1, n_tokens, n_blobs, 0, 0},
    background (3 doubles), tokens, then per blob hash + byte count (uint64) + bytes padded to a multiple of 8"""
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw, dtype="<i8", count=6)
    if hdr[0] != 0x47545241 or hdr[1] != 1:
        raise ValueError(f"{path}: not an exported scene graph")
    n_tokens, n_blobs = int(hdr[2]), int(hdr[3])
    background = np.frombuffer(raw, dtype="<f8", count=3, offset=48).copy()
    tokens = np.frombuffer(raw, dtype="<f8", count=n_tokens, offset=72).copy()
    off = 72 + 8 * n_tokens
    blobs = {}
    for _ in range(n_blobs):
        h, nbytes = (int(x) for x in np.frombuffer(raw, dtype="<u8", count=2, offset=off))
        off += 16
        blobs[h] = raw[off:off + nbytes]
        off += (nbytes + 7) // 8 * 8
    if off != len(raw):
        raise ValueError(f"{path}: trailing bytes")
    return tokens, blobs, background


class Scene:
    """A host scene graph built by the C++ scene_manager mirror, plus its flattened arrays."""

    def __init__(self, handle):
        self._h = handle
        info = _abi.SceneInfo()
        _check(_abi.load().artb_scene_get_info(self._h, C.byref(info)))
        self.info = {name: (list(getattr(info, name)) if hasattr(getattr(info, name), "__len__") else getattr(info, name))
                     for name, _ in _abi.SceneInfo._fields_}

    @classmethod
    def build(cls, alias, mesh_as_soup: bool = False, heightfield_side: int = 0, max_leaf_prims: int = 0,
              assets_root: str | None = None, device_bvh: bool = False) -> "Scene":
        """device_bvh=True skips the host SAH build: the flat scene carries no nodes and Context.upload builds the
        8-wide tree on the GPU (csrc/device/bvh_build.cuh)."""
        lib = _abi.load()
        opt = _abi.SceneOptions(int(mesh_as_soup), int(heightfield_side), int(max_leaf_prims), int(device_bvh),
                                assets_root.encode() if assets_root else None)
        h = C.c_void_p()
        _check(lib.artb_scene_build(int(alias), C.byref(opt), C.byref(h)))
        return cls(h)

    @classmethod
    def from_graph(cls, tokens: np.ndarray, blobs: dict, background, max_leaf_prims: int = 0, device_bvh: bool = False) -> "Scene":
        """artb_scene_from_graph: a scene graph built by a foreign host (the reference's own hittable_list serialised by
        include/artb_reference_adapter.hpp). tokens: the stream of artb_tokens.h; blobs: {hash: bytes} of the Perlin
        tables / image texels the stream names."""
        lib = _abi.load()
        tokens = np.ascontiguousarray(tokens, dtype=np.float64)
        keep = [np.frombuffer(b, dtype=np.uint8) for b in blobs.values()]
        table = (_abi.GraphBlob * max(len(blobs), 1))()
        for i, (h, a) in enumerate(zip(blobs.keys(), keep)):
            table[i] = _abi.GraphBlob(int(h), a.ctypes.data, a.size)
        bg = (C.c_double * 3)(*background)
        opt = _abi.SceneOptions(0, 0, int(max_leaf_prims), int(device_bvh), None)
        h = C.c_void_p()
        _check(lib.artb_scene_from_graph(tokens.ctypes.data, tokens.size, table, len(blobs), C.byref(bg), C.byref(opt), C.byref(h)))
        return cls(h)

    @classmethod
    def from_graph_file(cls, path, **kw) -> "Scene":
        """the container file artb_ref::exporter::write_file writes (include/artb_reference_adapter.hpp)"""
        tokens, blobs, background = read_graph_file(path)
        return cls.from_graph(tokens, blobs, background, **kw)

    def dump_graph(self) -> np.ndarray:
        lib = _abi.load()
        n = lib.artb_scene_dump_graph(self._h, None, 0)
        if n < 0:
            _check(int(n))
        out = np.empty(n, dtype=np.float64)
        lib.artb_scene_dump_graph(self._h, out.ctypes.data, n)
        return out

    def flat_view(self) -> _abi.FlatSceneView:
        v = _abi.FlatSceneView()
        _check(_abi.load().artb_scene_get_flat(self._h, C.byref(v)))
        return v

    def camera(self, width: int, height: int) -> "CameraDesc":
        """The camera src/main.cpp:33-35 builds for this scene: vup (0,1,0), focus 10, shutter [0,1)."""
        i = self.info
        return CameraDesc(i["lookfrom"], i["lookat"], (0.0, 1.0, 0.0), i["vfov"], width / height, i["aperture"], 10.0, 0.0, 1.0)

    def close(self):
        if self._h:
            _abi.load().artb_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class CameraDesc:
    """Arguments of camera::camera (camera.h:8-17)."""
    lookfrom: tuple
    lookat: tuple
    vup: tuple = (0.0, 1.0, 0.0)
    vfov: float = 40.0
    aspect_ratio: float = 1.0
    aperture: float = 0.0
    focus_dist: float = 10.0
    time0: float = 0.0
    time1: float = 0.0

    def to_c(self) -> _abi.Camera:
        d3 = C.c_double * 3
        return _abi.Camera(d3(*self.lookfrom), d3(*self.lookat), d3(*self.vup), self.vfov, self.aspect_ratio,
                           self.aperture, self.focus_dist, self.time0, self.time1)


@dataclass
class RenderOptions:
    width: int
    height: int
    samples_per_pixel: int
    max_depth: int = 50
    seed: int = 1
    stream: int = 0
    mode: int = _abi.MODE_FIXED
    sample_begin: int = 0
    tile_size: int = 32
    tile_parts: int = 1
    tile_part: int = 0
    adaptive_min_spp: int = 16
    adaptive_batch_spp: int = 8
    adaptive_threshold: float = 0.02
    accumulate: bool = False
    profile: int = 0              # N > 0: the kernels of every N-th wave are timed with CUDA events (stats extend_ms, traverse_ms ...)

    def to_c(self) -> _abi.RenderParams:
        return _abi.RenderParams(self.width, self.height, self.samples_per_pixel, self.max_depth, self.seed, self.stream,
                                 self.mode, self.sample_begin, self.tile_size, self.tile_parts, self.tile_part,
                                 self.adaptive_min_spp, self.adaptive_batch_spp, self.adaptive_threshold,
                                 int(self.accumulate), int(self.profile))


def _stats_dict(s: _abi.Stats) -> dict:
    d = {name: getattr(s, name) for name, _ in _abi.Stats._fields_}
    n = max(int(s.n_devices), 0)
    d["device_ms"] = list(s.device_ms)[:n]
    d["device_samples"] = list(s.device_samples)[:n]
    return d


class Context:
    """One GPU: artb_ctx."""

    def __init__(self, device: int = 0, pool_paths: int = 0, devices: list | None = None, partition: str = "samples"):
        """devices: several GPUs of this node behind one context (artb_config::devices): uploads replicate the scene,
        render() splits the work (partition "samples" = engine_mode::parallel_images, "tiles" = parallel_stripes made
        load-balanced) and sums the frames over NVLink on devices[0]."""
        lib = _abi.load()
        self._lib = lib
        self._h = C.c_void_p()
        devs = list(devices) if devices else []
        if len(devs) > _abi.MAX_DEVICES:
            raise ValueError(f"at most {_abi.MAX_DEVICES} devices per context")
        arr = (C.c_int32 * _abi.MAX_DEVICES)(*devs)
        cfg = _abi.Config(devs[0] if devs else device, pool_paths, len(devs),
                          _abi.PARTITION_TILES if partition == "tiles" else _abi.PARTITION_SAMPLES, arr)
        _check(lib.artb_create(C.byref(cfg), C.byref(self._h)))
        self.n_media = 0
        self.n_devices = lib.artb_context_devices(self._h)

    def upload(self, scene: Scene):
        _check(self._lib.artb_upload_scene(self._h, scene._h), self._h)
        self.n_media = scene.info["n_media"]

    def upload_flat(self, view: _abi.FlatSceneView):
        _check(self._lib.artb_upload_flat(self._h, C.byref(view)), self._h)
        self.n_media = view.n_media

    def trace_batch(self, rays: np.ndarray, t_min: float = 0.001, t_max: float = float("inf"),
                    medium_xi: np.ndarray | None = None) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 7)
        n = rays.shape[0]
        out = np.zeros(n, dtype=HIT_DTYPE)
        xi_ptr = None
        if medium_xi is not None and self.n_media:
            medium_xi = np.ascontiguousarray(medium_xi, dtype=np.float64).reshape(n, self.n_media)
            xi_ptr = medium_xi.ctypes.data
        _check(self._lib.artb_trace_batch(self._h, rays.ctypes.data, n, t_min, t_max, xi_ptr, out.ctypes.data), self._h)
        return out

    def shade_batch(self, rays: np.ndarray, prim_id: np.ndarray, t: np.ndarray, xi: np.ndarray) -> np.ndarray:
        """artb_shade_batch: material::emitted / scatter of explicit hits with explicit random numbers (n x 4)"""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 7)
        n = rays.shape[0]
        prim_id = np.ascontiguousarray(prim_id, dtype=np.int32).reshape(n)
        t = np.ascontiguousarray(t, dtype=np.float64).reshape(n)
        xi = np.ascontiguousarray(xi, dtype=np.float64).reshape(n, 4)
        out = np.zeros(n, dtype=SHADE_DTYPE)
        _check(self._lib.artb_shade_batch(self._h, rays.ctypes.data, prim_id.ctypes.data, t.ctypes.data, xi.ctypes.data, n,
                                          out.ctypes.data), self._h)
        return out

    def camera_rays(self, cam: "CameraDesc", st: np.ndarray, xi: np.ndarray) -> np.ndarray:
        """artb_camera_rays: camera::get_ray for explicit image-plane coordinates (n x 2) and random numbers (n x 3)"""
        st = np.ascontiguousarray(st, dtype=np.float64).reshape(-1, 2)
        n = st.shape[0]
        xi = np.ascontiguousarray(xi, dtype=np.float64).reshape(n, 3)
        out = np.zeros((n, 7), dtype=np.float64)
        c = cam.to_c()
        _check(self._lib.artb_camera_rays(self._h, C.byref(c), st.ctypes.data, xi.ctypes.data, n, out.ctypes.data), self._h)
        return out

    def build_bvh_device(self, max_leaf_prims: int = 1) -> dict:
        """artb_build_bvh_device: rebuild the tree of the uploaded scene on the GPU; returns build_ms / n_nodes / depth"""
        st = _abi.BvhStats()
        _check(self._lib.artb_build_bvh_device(self._h, max_leaf_prims, C.byref(st)), self._h)
        return {"build_ms": st.build_ms, "n_prims": st.n_prims, "n_nodes": st.n_nodes, "depth": st.depth}

    @staticmethod
    def last_global_prims(scene: "Scene") -> int:
        """number of global primitives (tested for every ray, not in the tree) of a scene"""
        return int(scene.flat_view().n_global_prims)

    @property
    def last_bvh_build(self) -> dict:
        st = _abi.BvhStats()
        self._lib.artb_last_bvh_build(self._h, C.byref(st))
        return {"build_ms": st.build_ms, "n_prims": st.n_prims, "n_nodes": st.n_nodes, "depth": st.depth}

    @property
    def trace_overflow_rays(self) -> int:
        """rays of the last trace_batch that took the one-phase overflow pass (0: all went through the pair kernels)"""
        return int(self._lib.artb_trace_overflow_rays(self._h))

    def trace_batch_device(self, rays_ptr: int, n: int, hits_ptr: int, t_min: float = 0.001, t_max: float = float("inf"),
                           xi_ptr: int | None = None) -> float:
        ms = C.c_float(0)
        _check(self._lib.artb_trace_batch_device(self._h, rays_ptr, n, t_min, t_max, xi_ptr, hits_ptr, C.byref(ms)), self._h)
        return ms.value

    def render(self, cam: CameraDesc, opt: RenderOptions, want_accum: bool = False, want_count: bool = False):
        """artb_render: host buffers in and out. Returns (rgb8[H,W,3], stats, accum|None, count|None)."""
        rgb = np.zeros((opt.height, opt.width, 3), dtype=np.uint8)
        accum = np.zeros((opt.height, opt.width, 3), dtype=np.float32) if want_accum else None
        count = np.zeros((opt.height, opt.width), dtype=np.uint32) if want_count else None
        st = _abi.Stats()
        c, p = cam.to_c(), opt.to_c()
        _check(self._lib.artb_render(self._h, C.byref(c), C.byref(p), rgb.ctypes.data,
                                     accum.ctypes.data if want_accum else None,
                                     count.ctypes.data if want_count else None, C.byref(st)), self._h)
        return rgb, _stats_dict(st), accum, count

    def render_device(self, cam: CameraDesc, opt: RenderOptions, accum_ptr: int, count_ptr: int | None = None) -> dict:
        st = _abi.Stats()
        c, p = cam.to_c(), opt.to_c()
        _check(self._lib.artb_render_device(self._h, C.byref(c), C.byref(p), accum_ptr, count_ptr, C.byref(st)), self._h)
        return _stats_dict(st)

    def resolve_device(self, accum_ptr: int, count_ptr: int | None, n_pixels: int, fixed_spp: int, rgb_ptr: int):
        _check(self._lib.artb_resolve_device(self._h, accum_ptr, count_ptr, n_pixels, fixed_spp, rgb_ptr), self._h)

    @property
    def stream(self) -> int:
        return self._lib.artb_stream(self._h) or 0

    def close(self):
        if self._h:
            self._lib.artb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """engine<W,H,3> of the reference (engine.h:18-54), runtime-sized.

        eng = Engine(cam, EngineMode.single, width=800, height=800, samples_per_pixel=100)
        eng.set_scene(scene)
        ms = eng.run(output_image)       # output_image: uint8 array of W*H*3, row 0 = top
    """

    def __init__(self, cam: CameraDesc, mode: EngineMode = EngineMode.single, *, width: int, height: int,
                 samples_per_pixel: int = 100, max_depth: int = 50, seed: int = 1, device: int = 0,
                 context: Context | None = None, devices: list | None = None):
        """devices: the reference's parallel modes use 4 worker threads (engine.h:335-445); here a worker is a GPU.
        EngineMode.parallel_images splits the samples over `devices`, parallel_stripes the image (interleaved tiles);
        without `devices` every mode renders on the one GPU `device`."""
        self.cam = cam
        self.mode = mode
        self.width, self.height = width, height
        self.samples_per_pixel, self.max_depth, self.seed = samples_per_pixel, max_depth, seed
        self._ctx = context
        self._device = device
        self._devices = list(devices) if devices else None
        self._scene = None
        self.last_stats: dict = {}

    def set_scene(self, scene: Scene | None):
        self._scene = scene
        if scene is not None and scene.info["top_level_objects"] > 0:
            if self._ctx is None:
                if self._devices and self.mode in (EngineMode.parallel_images, EngineMode.parallel_stripes):
                    self._ctx = Context(devices=self._devices,
                                        partition="tiles" if self.mode is EngineMode.parallel_stripes else "samples")
                else:
                    self._ctx = Context(self._devices[0] if self._devices else self._device)
            self._ctx.upload(scene)

    def run(self, output_image: np.ndarray) -> int:
        if self._scene is None or self._scene.info["top_level_objects"] == 0 or self._ctx is None:
            print("Invalid input scene!", file=sys.stderr)      # engine.h:32-36
            return -1
        if self.mode is EngineMode.adaptive and (self.width % 12 or self.height % 12):
            # engine.h:182-183 — kept so that callers relying on the reference's contract see the same failure
            raise ValueError("for adaptive strategy image size should perfectly fit big square size for now!!")
        mode = {EngineMode.adaptive: _abi.MODE_ADAPTIVE_SQUARES, EngineMode.variance_mask: _abi.MODE_VARIANCE_MASK}.get(self.mode, _abi.MODE_FIXED)
        opt = RenderOptions(self.width, self.height, self.samples_per_pixel, self.max_depth, self.seed, mode=mode)
        rgb, stats, _, _ = self._ctx.render(self.cam, opt)
        np.copyto(np.asarray(output_image).reshape(self.height, self.width, 3), rgb)
        self.last_stats = stats
        return int(round(stats["elapsed_ms"]))
