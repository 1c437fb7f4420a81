"""Multi-GPU partitioning of a render: one process per GPU, scene replicated, ONE exchange at the end.

The reference itself splits work both ways on its 4-thread pool: row stripes (engine.h:335-376) and a sample split with
a final sum of four float frames (`_run_parallel_images`, engine.h:378-445 — integer spp/4 per worker, summed at
:424-440). Here:

  partition A  "samples": every rank renders all pixels with its share of the per-pixel sample indices on its own
               Philox stream; exchange = SUM-reduce of the W*H*3 float radiance sums (+ W*H counts in variance-mask
               mode) to rank 0, which resolves (divide, gamma 2, clamp, 8 bit — color.h:6-22).
  partition B  "tiles": interleaved tile_size x tile_size tiles, tile t belongs to rank t % world; frames are zero
               outside a rank's tiles, so the same SUM-reduce assembles the image (disjoint support).

Unlike the reference's spp/4 (which drops samples when spp % 4 != 0), shares differ by at most one sample and always
add up to spp. Nothing else crosses GPUs. The collective is torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU
tests); payloads are tiny (800x800x3 f32 = 7.7 MB).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Share:
    rank: int
    world: int
    mode: str               # "samples" | "tiles"
    sample_begin: int       # first per-pixel sample index rendered by this rank
    samples: int            # samples per pixel rendered by this rank
    tile_parts: int
    tile_part: int
    stream: int             # Philox stream id


def sample_share(spp: int, rank: int, world: int, base: int = 0) -> Share:
    """Split `spp` per-pixel samples as evenly as possible; ranks get contiguous, disjoint index ranges."""
    if not (0 <= rank < world) or spp < 0:
        raise ValueError("bad rank/world/spp")
    q, r = divmod(spp, world)
    mine = q + (1 if rank < r else 0)
    begin = base + rank * q + min(rank, r)
    return Share(rank, world, "samples", begin, mine, 1, 0, 0)


def tile_share(spp: int, rank: int, world: int, base: int = 0) -> Share:
    """Every rank renders all `spp` samples of its own interleaved tiles."""
    if not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return Share(rank, world, "tiles", base, spp, world, rank, 0)


def owned_pixel_mask(width: int, height: int, tile_size: int, share: Share):
    """Boolean [H,W] mask of the pixels `share` renders (all True for the sample split). Mirrors the tile rule of
    Renderer::render_device (csrc/device/kernels.cu): tile index (y/ts)*tiles_x + (x/ts), owner = index % parts."""
    import numpy as np
    if share.mode == "samples" or share.tile_parts <= 1:
        return np.ones((height, width), dtype=bool)
    ts = max(tile_size, 1)
    tiles_x = (width + ts - 1) // ts
    ys, xs = np.mgrid[0:height, 0:width]
    return ((ys // ts) * tiles_x + (xs // ts)) % share.tile_parts == share.tile_part


def reduce_frames(accum, count=None, dst: int = 0):
    """SUM-reduce the per-rank radiance sums (and sample counts) to `dst`. No-op without an initialised process group
    or with world size 1. `accum` / `count` are torch tensors (CUDA for NCCL, CPU for gloo); count may be None."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    if count is not None:
        # uint32 is not reducible on every backend: counts travel as int32/int64 views
        dist.reduce(count, dst=dst, op=dist.ReduceOp.SUM)


def resolve_host(accum, spp_or_count):
    """write_color (color.h:6-22) on the host, for tests: uint8 = 256 * clamp(sqrt(sum / n), 0, 0.999)."""
    import numpy as np
    a = np.asarray(accum, dtype=np.float64)
    n = np.asarray(spp_or_count, dtype=np.float64)
    if n.ndim == 2:
        n = np.maximum(n, 1.0)[..., None]
    v = np.sqrt(a / n)
    return (256.0 * np.clip(v, 0.0, 0.999)).astype(np.uint8)
