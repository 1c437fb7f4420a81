#!/usr/bin/env python3
"""Converged oracle images for the PSNR gate (tests/golden/converged_<alias>_<W>x<H>.npz).

Uses the C++ restatement of the reference (oracle/port, pinned bit-identical to the unmodified reference by
tests/test_oracle_port.py) in `procs` processes, each re-seeding the reference's MT19937 after the scene build (independent streams), and
averages the per-pixel float64 means — the reference itself would need hours single-threaded for the same sample count."""
import multiprocessing as mp
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))

JOBS = [  # alias, W, H, total spp. The north-star scenes (final scene, Cornell box, Cornell smoke) get 64 Ki spp so that the
          # fixture's own noise limit sits above 46 dB and the GPU test can assert PSNR >= 40 dB outright.
    (8, 128, 128, 65536), (6, 128, 128, 65536), (7, 96, 96, 65536), (1, 160, 90, 4096), (11, 160, 90, 2048), (5, 160, 90, 4096),
]
CHUNK = 1024     # spp per work item (one independent MT19937 stream each)


def work(args):
    alias, w, h, spp, seed = args
    import oracle_lib as O
    s = O.OracleScene(alias)
    _, lin, ms, rays = s.render(w, h, spp, reseed=seed)      # independent, non-overlapping streams
    return lin, rays


def main():
    """usage: make_converged.py [procs] [alias ...]   (default: 8 processes, every job)"""
    procs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    only = {int(a) for a in sys.argv[2:]}
    out = ROOT / "tests" / "golden"
    for alias, w, h, spp in JOBS:
        if only and alias not in only:
            continue
        t = time.time()
        n_items = max(2, spp // CHUNK)
        each = spp // n_items
        with mp.Pool(procs) as pool:
            res = pool.map(work, [(alias, w, h, each, 90001 + 7919 * i) for i in range(n_items)], chunksize=1)
        lin = np.mean([r[0] for r in res], axis=0)
        half_a = np.mean([r[0] for r in res[: n_items // 2]], axis=0)
        half_b = np.mean([r[0] for r in res[n_items // 2:]], axis=0)
        rays = sum(r[1] for r in res)
        np.savez_compressed(out / f"converged_{alias}_{w}x{h}.npz", lin=lin.astype(np.float32), half_a=half_a.astype(np.float32),
                            half_b=half_b.astype(np.float32), spp=each * n_items, alias=alias, rays_per_sample=rays / (w * h * each * n_items))
        print(alias, w, h, each * n_items, f"{time.time() - t:.0f}s rays/sample {rays / (w * h * each * n_items):.3f}", flush=True)


if __name__ == "__main__":
    main()
