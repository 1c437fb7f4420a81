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

JOBS = [  # alias, W, H, total spp
    (8, 128, 128, 8192), (6, 128, 128, 8192), (1, 160, 90, 4096), (7, 96, 96, 4096), (11, 160, 90, 2048), (5, 160, 90, 4096),
]


def work(args):
    alias, w, h, spp, seed = args
    import oracle_lib as O
    s = O.OracleScene(alias)
    _, lin, ms, rays = s.render(w, h, spp, reseed=seed)      # independent, non-overlapping streams
    return lin, rays


def main():
    procs = 8
    out = ROOT / "tests" / "golden"
    for alias, w, h, spp in JOBS:
        t = time.time()
        each = spp // procs
        with mp.Pool(procs) as pool:
            res = pool.map(work, [(alias, w, h, each, 90001 + 7919 * i) for i in range(procs)])
        lin = np.mean([r[0] for r in res], axis=0)
        half_a = np.mean([r[0] for r in res[: procs // 2]], axis=0)
        half_b = np.mean([r[0] for r in res[procs // 2:]], axis=0)
        rays = sum(r[1] for r in res)
        np.savez_compressed(out / f"converged_{alias}_{w}x{h}.npz", lin=lin.astype(np.float32), half_a=half_a.astype(np.float32),
                            half_b=half_b.astype(np.float32), spp=each * procs, alias=alias, rays_per_sample=rays / (w * h * each * procs))
        print(alias, w, h, each * procs, f"{time.time() - t:.0f}s rays/sample {rays / (w * h * each * procs):.3f}", flush=True)


if __name__ == "__main__":
    main()
