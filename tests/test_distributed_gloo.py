"""Multi-GPU host logic on CPU: world_size 2 over gloo (the N>1 path of bench.py minus the CUDA render).

Each rank fills the frame it WOULD have rendered (its tiles / its sample share) with a deterministic synthetic radiance
field, the frames are SUM-reduced to rank 0 exactly as bench.py does, and rank 0 resolves. The result must equal the
single-process frame."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

sys.path.insert(0, str(ROOT))
from another_raytracer_b200 import distributed as D  # noqa: E402

W, H, SPP, TS = 96, 64, 37, 32


def field(w, h):
    ys, xs = np.mgrid[0:h, 0:w]
    return np.stack([0.2 + 0.5 * xs / w, 0.1 + 0.7 * ys / h, 0.05 + 0.0 * xs], axis=-1)   # per-sample radiance


def _worker(rank, world, port, mode, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    share = D.sample_share(SPP, rank, world) if mode == "samples" else D.tile_share(SPP, rank, world)
    mask = D.owned_pixel_mask(W, H, TS, share)
    accum = torch.from_numpy(field(W, H) * share.samples * mask[..., None]).to(torch.float32)
    count = torch.from_numpy((mask * share.samples).astype(np.int64))
    D.reduce_frames(accum, count)
    if rank == 0:
        np.save(os.path.join(out_dir, f"accum_{mode}.npy"), accum.numpy())
        np.save(os.path.join(out_dir, f"count_{mode}.npy"), count.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode", ["samples", "tiles"])
def test_two_rank_reduce_assembles_the_frame(tmp_path, mode):
    mp.spawn(_worker, args=(2, _free_port(), mode, str(tmp_path)), nprocs=2, join=True)
    accum = np.load(tmp_path / f"accum_{mode}.npy")
    count = np.load(tmp_path / f"count_{mode}.npy")
    assert np.all(count == SPP)                                   # every pixel got exactly spp samples
    assert np.allclose(accum, field(W, H) * SPP, rtol=1e-6)
    img = D.resolve_host(accum, SPP)
    assert np.array_equal(img, D.resolve_host(field(W, H) * SPP, np.full((H, W), SPP)))


def test_sample_shares_partition_exactly():
    for spp in (1, 7, 10, 256, 10000):
        for world in (1, 2, 3, 4, 8):
            shares = [D.sample_share(spp, r, world) for r in range(world)]
            assert sum(s.samples for s in shares) == spp          # unlike the reference's spp/4 (engine.h:411-414)
            pos = 0
            for s in shares:
                assert s.sample_begin == pos
                pos += s.samples
            assert max(s.samples for s in shares) - min(s.samples for s in shares) <= 1


def test_tile_masks_are_disjoint_and_cover():
    for world in (2, 3, 4, 8):
        masks = [D.owned_pixel_mask(100, 70, 32, D.tile_share(4, r, world)) for r in range(world)]
        assert np.all(np.sum(masks, axis=0) == 1)
    with pytest.raises(ValueError):
        D.sample_share(4, 2, 2)


def test_resolve_host_matches_write_color():
    """color.h:6-22: sqrt(sum/n), clamp to [0,0.999], *256, truncate"""
    acc = np.array([[[0.0, 4.0, 1.0], [0.25 * 4, 100.0, 0.0001]]])
    out = D.resolve_host(acc, 4)
    assert out.tolist() == [[[0, 255, 128], [128, 255, 1]]]
