#!/usr/bin/env python3
"""bench.py — the headline benchmark: samples/s (and Mrays/s) of the path-tracing hot path on the reference's own
scenes, on N B200s, next to the reference CPU renderer.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W     # the reference's CPU renderer on the host cores

Workload (config.workload): BASELINE.json configs[2], the Next Week FINAL SCENE at 800x800, the scene the north_star
target is quoted on. One "step" = one pass of the hot path over one batch = `--spp-per-step` (default 256) samples for
every pixel; 10 000 spp is 39 such steps and samples/s does not depend on how many are run. Steps accumulate into one
image (distinct per-pixel sample indices), so K steps give a K*256-spp render.

  value     whole-job samples/s with the scene resident in HBM (CUDA-event time of artb_render_device, max over ranks)
  e2e       the same metric through artb_upload_scene + artb_render with HOST buffers: scene H2D, image D2H inside the
            timed region
  roofline  extend kernel: algorithmic bytes/ray x rays / summed per-launch CUDA-event time, against the
            measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the unmodified reference (oracle/_ref) on the box's host cores, bounded sample (N=1, rank 0 only)

Multi-GPU: scene replicated, sample split with one Philox stream per rank, one SUM-reduce of the float frame per step
(another_raytracer_b200/distributed.py); "weak" scaling: every rank renders spp-per-step samples per pixel per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("ARTB_ASSETS", str(ROOT / "assets"))

# algorithmic bytes per ray on the REFERENCE's trees: 32 B per node visit + primitive records + 128 B state
B_RAY = {1: 2354, 6: 768, 7: 768, 8: 1322, 9: 1275, 10: 1346, 11: 539}
PROFILE_EVERY = 8          # the kernels of every 8th wave are timed with CUDA events (roofline.avg_launch_ms and the per-kernel times)

WORKLOADS = {
    "final": dict(alias=8, width=800, height=800, name="next_week_final_scene_800x800", full_spp=10000),
    "cornell": dict(alias=6, width=600, height=600, name="cornell_box_600x600", full_spp=1000),
    "random": dict(alias=1, width=400, height=225, name="rtiow_random_spheres_400x225", full_spp=10),
    "cow": dict(alias=10, width=1920, height=1080, name="cow_obj_1920x1080", full_spp=256),
    "dino": dict(alias=11, width=1920, height=1080, name="dino_obj_1920x1080", full_spp=256),
    # BASELINE.json configs[4]: 1 M-triangle height field at 4K; the tree is built on the GPU (bvh_build.cuh)
    "heightfield": dict(alias=12, width=3840, height=2160, name="heightfield_1M_triangles_3840x2160", full_spp=64,
                        scene_kwargs=dict(mesh_as_soup=True, device_bvh=True)),
}


def ncu_traffic(workload: str):
    """DRAM bytes per launch of the closest-hit kernels from the committed ncu --set full capture OF THIS WORKLOAD
    (profiles/traffic.json, one entry per workload, written by tools/ncu_summary.py traffic <report> <workload>); None when no
    capture of this workload is committed — a capture of another scene is not evidence for this one."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        t = json.loads(p.read_text())[workload]
        return {"bytes_per_launch": int(sum(v for k, v in t["kernels"].items() if k in ("k_traverse", "k_pairs", "k_finish", "k_walk", "k_bin"))),
                "per_kernel": t["kernels"], "source": t.get("source")}
    except Exception:
        return None


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in Path(self.path).read_text().splitlines():
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------------------------
# the reference's CPU renderer on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(wl: dict, procs: int, spp: int = 1) -> dict:
    """`procs` concurrent single-mode processes of the UNMODIFIED reference engine (oracle/_ref/ref_engine_WxHxS), each
    rendering the whole image at `spp` with a different RNG offset: the all-core upper bound of SURVEY 8d row (iii)
    (the shipped parallel mode is hard-wired to 4 racy threads and scales 1.0-1.6x). Falls back to the C++ restatement
    (oracle/_ref/liboracle_port.so) when the reference binary for this size was not built."""
    w, h, alias = wl["width"], wl["height"], wl["alias"]
    exe = ROOT / "oracle" / "_ref" / f"ref_engine_{w}x{h}x{spp}"
    samples_each = w * h * spp
    if exe.exists():
        with tempfile.TemporaryDirectory() as tmp:
            env = dict(os.environ, ARTB_ASSETS=str(ROOT / "assets"))
            t0 = time.perf_counter()
            ps = [subprocess.Popen([str(exe), str(alias), "single", os.path.join(tmp, f"o{i}.raw"), str(1000 + 7919 * i)],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env) for i in range(procs)]
            ms = []
            for p in ps:
                out, _ = p.communicate()
                line = [l for l in out.splitlines() if l.startswith("{")]
                ms.append(float(json.loads(line[-1])["elapsed_ms"]))
            wall = time.perf_counter() - t0
        sps = procs * samples_each / (max(ms) / 1e3)
        return dict(value=sps / 1e6, unit="Msamples/s", cores=procs, kind="reference", wall_s=wall,
                    sample=f"{procs} concurrent engine_mode::single processes of the unmodified reference, each {w}x{h}x{spp}spp "
                           f"(engine's own elapsed ms, slowest process)")
    # restatement fallback
    sys.path.insert(0, str(ROOT / "tests"))
    import multiprocessing as mp
    rows = max(1, h // 16)
    with mp.Pool(procs) as pool:
        t0 = time.perf_counter()
        res = pool.starmap(_port_rows, [(alias, w, h, spp, i * (h // procs), min(h, i * (h // procs) + rows), 1000 + 7919 * i)
                                         for i in range(procs)])
        wall = time.perf_counter() - t0
    sps = sum(r[0] for r in res) / (max(r[1] for r in res) / 1e3)
    return dict(value=sps / 1e6, unit="Msamples/s", cores=procs, kind="port", wall_s=wall,
                sample=f"{procs} processes of the C++ restatement (oracle/port), {rows} rows of {w}x{h}x{spp}spp each")


def cpu_reference_modes(wl: dict) -> list:
    """the reference's own render modes next to the all-core bound, each ONE process of the
    unmodified engine (oracle/_ref/ref_engine_*): engine_mode::single (1 thread), parallel_stripes and parallel_images as
    shipped (hard-wired to 4 worker threads, engine.h:335-445; parallel_images renders spp/4 per worker, so it needs the
    4-spp binary). Rows whose binary is missing are skipped."""
    w, h, alias = wl["width"], wl["height"], wl["alias"]
    rows = []
    for mode, arg, spp, cores in (("single", "single", 1, 1), ("parallel_stripes", "stripes", 1, 4), ("parallel_images", "images", 4, 4)):
        exe = ROOT / "oracle" / "_ref" / f"ref_engine_{w}x{h}x{spp}"
        if not exe.exists():
            continue
        with tempfile.TemporaryDirectory() as tmp:
            env = dict(os.environ, ARTB_ASSETS=str(ROOT / "assets"))
            try:
                out = subprocess.run([str(exe), str(alias), arg, os.path.join(tmp, "o.raw")], capture_output=True, text=True, env=env,
                                     timeout=600).stdout
                ms = float(json.loads([l for l in out.splitlines() if l.startswith("{")][-1])["elapsed_ms"])
            except Exception:
                continue
        if ms > 0:
            rows.append({"mode": f"engine_mode::{mode}", "threads": cores, "value": w * h * spp / (ms / 1e3) / 1e6, "unit": "Msamples/s",
                         "sample": f"one process, {w}x{h}x{spp}spp, engine's own elapsed ms"})
    return rows


def _port_rows(alias, w, h, spp, r0, r1, burn):
    import oracle_lib as O
    s = O.OracleScene(alias)
    _, _, ms, _ = s.render(w, h, spp, rows=(r0, r1), burn=burn)
    return (r1 - r0) * w * spp, ms


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    procs = os.cpu_count() or 1
    vals = []
    for _ in range(args.warmup):
        cpu_reference_run(wl, procs)
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = cpu_reference_run(wl, procs)
        vals.append(last["value"])
    total = time.perf_counter() - t0
    value = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": "samples_per_sec", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "scene_alias": wl["alias"], "width": wl["width"], "height": wl["height"],
                   "max_depth": 50, "spp_per_step": 1, "note": "samples/s is spp-independent; CPU step = 1 spp per process"},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": last["cores"], "kind": last["kind"], "sample": last["sample"]},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------------------
# this repo's CUDA path
# ------------------------------------------------------------------------------------------------------------------
def run_cuda_arm(args, wl):
    import numpy as np
    import torch
    import torch.distributed as dist
    import another_raytracer_b200 as A
    from another_raytracer_b200 import distributed as D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this backend has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    W, H, alias, S = wl["width"], wl["height"], wl["alias"], args.spp_per_step
    scene = A.Scene.build(alias, **wl.get("scene_kwargs", {}))
    ctx = A.Context(local, args.pool)
    ctx.upload(scene)
    bvh_build = ctx.last_bvh_build
    cam = scene.camera(W, H)
    tiles = args.partition == "tiles"
    step_count = torch.zeros((H, W), dtype=torch.int32, device=dev) if args.variance_mask else None
    step_accum = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    total_accum = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    rgb = torch.zeros((H, W, 3), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # Everything a step does is queued on ONE stream — the library's (artb_stream): the L2 flush, the render's kernels, the
    # NCCL reduce of the frames and the accumulation on rank 0 — so a pair of CUDA events on that stream times the whole
    # step, the exchange included (the reduce used to sit outside the timed value).
    lib_stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.warmup + args.steps)]

    def one_step(k: int) -> dict:
        with torch.cuda.stream(lib_stream):
            flush.zero_()                                                # L2 flush between timed iterations
            e0, e1 = ev[k]
            e0.record(lib_stream)
            # partition A (default): every rank renders S samples of every pixel on its own Philox stream and index range;
            # partition B (--partition tiles, BASELINE configs[4]): every rank renders S samples of its interleaved tiles
            opt = A.RenderOptions(W, H, S, seed=args.seed, stream=0 if tiles else rank,
                                  sample_begin=(k if tiles else k * world + rank) * S, profile=PROFILE_EVERY,
                                  tile_parts=world if tiles else 1, tile_part=rank if tiles else 0,
                                  mode=1 if args.variance_mask else 0)
            st = ctx.render_device(cam, opt, step_accum.data_ptr(), step_count.data_ptr() if step_count is not None else None)
            D.reduce_frames(step_accum)                                  # the one exchange of the path (no-op at N=1)
            if rank == 0:
                total_accum.add_(step_accum)
            e1.record(lib_stream)
        return st

    for k in range(args.warmup):
        one_step(k)
    torch.cuda.synchronize()
    total_accum.zero_()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()                                                      # every rank samples ITS GPU's clocks
    t0 = time.perf_counter()
    stats = [one_step(args.warmup + k) for k in range(args.steps)]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()

    render_ms = sum(s["elapsed_ms"] for s in stats)                      # the library's own events: render kernels only
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev[args.warmup:])     # render + reduce + accumulate, per step
    t = torch.tensor([dev_ms, wall * 1e3, render_ms, -(clocks["sm_mhz"] or 0.0)], dtype=torch.float64, device=dev)
    tmin = torch.tensor([render_ms], dtype=torch.float64, device=dev)
    agg = torch.tensor([sum(s["samples"] for s in stats), sum(s["rays"] for s in stats), sum(s["launches"] for s in stats)],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max, render_ms_max, min_rank_sm_mhz = float(t[0]), float(t[1]), float(t[2]), -float(t[3])
    render_ms_min = float(tmin[0])
    samples, rays, launches = (float(x) for x in agg)

    # resolve the accumulated image on rank 0 (write_color) and keep a checksum so the work cannot be skipped
    checksum = None
    if rank == 0:
        ctx.resolve_device(total_accum.data_ptr(), None, W * H, S * (1 if tiles else world) * args.steps, rgb.data_ptr())
        torch.cuda.synchronize()
        checksum = int(rgb.to(torch.int64).sum().item())

    # e2e: the public API with HOST buffers, every step: scene arrays H2D (artb_upload_scene), render, image D2H.
    #   N = 1: artb_render (RGB8 + float sums copied back by the library).
    #   N > 1: every rank uploads and renders its share (artb_render_device), ONE reduce, rank 0 resolves (artb_resolve_device)
    #          and copies RGB8 + float sums to pinned host memory — the exchange is inside the timed region.
    e2e = None
    if args.e2e_steps > 0:
        e_samples, h2d, d2h = 0, 0, 0
        # one untimed call first: artb_render allocates its device scratch on first use (tens of milliseconds of cudaMalloc)
        ctx.upload(scene)
        ctx.render(cam, A.RenderOptions(W, H, min(S, 8), seed=args.seed, stream=rank, sample_begin=999 * S), want_accum=True)
        host_rgb = torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True)
        host_acc = torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True)
        barrier()
        te = time.perf_counter()
        for k in range(args.e2e_steps):
            ctx.upload(scene)                                            # scene arrays: pageable host -> HBM
            opt = A.RenderOptions(W, H, S, seed=args.seed, stream=rank, sample_begin=(1000 + k * world + rank) * S)
            if world == 1:
                img, st, acc, _ = ctx.render(cam, opt, want_accum=True)    # rgb8 + float sums copied back to the host
                d2h += st["d2h_bytes"]
            else:
                with torch.cuda.stream(lib_stream):
                    st = ctx.render_device(cam, opt, step_accum.data_ptr(), None)
                    D.reduce_frames(step_accum)
                    if rank == 0:
                        ctx.resolve_device(step_accum.data_ptr(), None, W * H, S * world, rgb.data_ptr())
                lib_stream.synchronize()
                if rank == 0:
                    # (copies on torch's own stream: a pinned block remembers the stream it was last used on, and the
                    # library's stream is gone by the time torch releases its host cache at exit)
                    host_rgb.copy_(rgb)
                    host_acc.copy_(step_accum)
                    d2h += host_rgb.numel() + host_acc.numel() * 4
            e_samples += st["samples"]
            if os.environ.get("ARTB_BENCH_DEBUG"):
                print(f"e2e step {k}: wall so far {time.perf_counter() - te:.3f} s, device {st['elapsed_ms']:.1f} ms", file=sys.stderr)
            h2d += scene.info["flat_bytes"] + st["h2d_bytes"]
        barrier()
        e_wall = time.perf_counter() - te
        tt = torch.tensor([e_wall], dtype=torch.float64, device=dev)
        ss = torch.tensor([float(e_samples), float(h2d), float(d2h)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(ss, op=dist.ReduceOp.SUM)
        e2e = {"value": float(ss[0]) / float(tt[0]) / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(float(ss[1]) / args.e2e_steps), "d2h_bytes_per_step": int(float(ss[2]) / args.e2e_steps),
               "steps": args.e2e_steps, "reduce_in_timed_region": world > 1}

    # BASELINE.json configs[4] next to the headline (so that the driver's scaling run carries it): the 1 M-triangle height
    # field at 4K, tree built on the GPU, TILE-sharded across the ranks, per-pixel variance mask with a 64-spp cap.
    config5 = run_config5(args, A, D, torch, dist, rank, world, local, dev) if args.config5_steps > 0 else None

    if rank == 0:
        ext_ms = sum(s["extend_ms"] for s in stats)
        ext_launches = sum(s["extend_launches"] for s in stats)
        rank0_rays = sum(s["rays"] for s in stats)
        peak, peak_src = measured_peak_gbs()
        b_ray = B_RAY.get(alias)
        roofline = None
        waves = sum(s["waves"] for s in stats)
        if b_ray and ext_ms > 0:
            # the kernels of every PROFILE_EVERY-th wave are timed (events on every wave cost 1.7 % of the step): mean rays per
            # wave over ALL waves / mean time of the timed waves
            achieved = b_ray * (rank0_rays / max(waves, 1)) / (ext_ms / max(ext_launches, 1) * 1e-3) / 1e9
            tr = ncu_traffic(args.workload)
            roofline = {"bound": "hbm", "kernel": "closest-hit stage = k_traverse + k_pairs + k_finish (+ k_extend overflow pass)",
                        "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": tr["bytes_per_launch"] if tr else None,
                        "traffic_per_kernel": tr["per_kernel"] if tr else None, "traffic_source": tr["source"] if tr else None,
                        "peak_source": peak_src,
                        "algorithmic_bytes_per_ray": b_ray, "rays_per_launch": rank0_rays / max(waves, 1),
                        "avg_launch_ms": ext_ms / max(ext_launches, 1), "launches": int(waves), "launches_timed": int(ext_launches),
                        "timed_every": PROFILE_EVERY,
                        "extend_share_of_step": (ext_ms / max(ext_launches, 1)) * waves / max(sum(s["elapsed_ms"] for s in stats), 1e-9),
                        "traverse_ms_per_launch": sum(s["traverse_ms"] for s in stats) / max(ext_launches, 1),
                        "pairs_ms_per_launch": sum(s["pairs_ms"] for s in stats) / max(ext_launches, 1),
                        "finish_ms_per_launch": sum(s["finish_ms"] for s in stats) / max(ext_launches, 1),
                        "overflow_rays": int(sum(s["overflow_rays"] for s in stats))}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            c = cpu_reference_run(wl, os.cpu_count() or 1)
            cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
            # the reference's own modes beside the all-core bound: single, parallel_stripes and
            # parallel_images with the 4 threads they ship with
            cpu["rows"] = cpu_reference_modes(wl) + [{"mode": "all-core: one engine_mode::single process per core", "threads": c["cores"],
                                                      "value": c["value"], "unit": c["unit"], "sample": c["sample"]}]
        value = samples / (dev_ms_max * 1e-3) / 1e6
        clocks["min_rank_sm_mhz"] = min_rank_sm_mhz
        line = {
            "metric": "samples_per_sec", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "scene_alias": alias, "width": W, "height": H, "max_depth": 50,
                       "spp_per_step": S, "full_config_spp": wl["full_spp"], "partition": f"{args.partition} x{world}", "variance_mask": bool(args.variance_mask), "pool_paths": args.pool,
                       "seed": args.seed, "l2": "256 MiB flush buffer written between timed steps; path pool > L2",
                       "device_bvh_build_ms": bvh_build["build_ms"] if bvh_build["n_nodes"] else None},
            "mrays_per_sec": rays / (dev_ms_max * 1e-3) / 1e6, "rays_per_sample": rays / max(samples, 1),
            "wall_ms_per_step": wall_ms_max / max(args.steps, 1),
            # the same steps without the exchange: the library's own event time of the render kernels, slowest / fastest rank
            "render_ms_per_step": {"max_rank": render_ms_max / max(args.steps, 1), "min_rank": render_ms_min / max(args.steps, 1)},
            "timed_region": "CUDA events on the library stream around render kernels + NCCL reduce + accumulate (L2 flush before, outside)",
            "config5": config5,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "image_checksum": checksum,
        }
        print(json.dumps(line), flush=True)
    # torch objects that were used on the library's stream are released before the context that owns that stream
    torch.cuda.synchronize()
    del ev, lib_stream
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


def run_config5(args, A, D, torch, dist, rank, world, local, dev):
    """One bounded pass of BASELINE.json configs[4]: 3840x2160, 999 698 triangles (alias 12 as one triangle_soup, tree built
    by the device builder), interleaved 32x32 tiles over the ranks, variance mask capped at 64 spp. Reports whole-job
    Msamples/s (samples all ranks traced / slowest rank's device time incl. the reduce) and the load balance."""
    W, H, cap = 3840, 2160, 64
    scene = A.Scene.build(12, mesh_as_soup=True, device_bvh=True)
    ctx = A.Context(local, args.pool)
    ctx.upload(scene)
    build = ctx.last_bvh_build
    cam = scene.camera(W, H)
    accum = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    count = torch.zeros((H, W), dtype=torch.int32, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    out = []
    for k in range(1 + args.config5_steps):                               # one untimed warm-up
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            opt = A.RenderOptions(W, H, cap, seed=args.seed + k, mode=1, tile_size=32, tile_parts=world, tile_part=rank,
                                  adaptive_min_spp=16, adaptive_batch_spp=16, adaptive_threshold=0.02)
            st = ctx.render_device(cam, opt, accum.data_ptr(), count.data_ptr())
            D.reduce_frames(accum, count)
            e1.record(stream)
        stream.synchronize()
        if k:
            out.append((e0.elapsed_time(e1), st))
    ms = sum(o[0] for o in out)
    mine = torch.tensor([ms, float(sum(o[1]["samples"] for o in out)), float(sum(o[1]["rays"] for o in out)),
                         float(out[-1][1]["pixels_active"] + out[-1][1]["pixels_converged"])], dtype=torch.float64, device=dev)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    ctx.close()
    scene.close()
    per = [[float(x) for x in r] for r in allr]
    ms_max = max(p[0] for p in per)
    samples = sum(p[1] for p in per)
    return {"workload": "heightfield_1M_triangles_3840x2160 (BASELINE configs[4])", "partition": f"tiles 32x32 x{world}", "variance_mask": True,
            "spp_cap": cap, "steps": args.config5_steps, "value": samples / (ms_max * 1e-3) / 1e6, "unit": "Msamples/s",
            "mrays_per_sec": sum(p[2] for p in per) / (ms_max * 1e-3) / 1e6, "ms_per_step": ms_max / args.config5_steps,
            "mean_spp": samples / args.config5_steps / (W * H), "device_bvh_build_ms": build["build_ms"], "bvh_nodes": build["n_nodes"],
            "per_rank": {"ms": [p[0] / args.config5_steps for p in per], "samples": [p[1] / args.config5_steps for p in per],
                         "pixels": [p[3] for p in per]},
            "imbalance": {"samples_max_over_mean": max(p[1] for p in per) / (samples / world),
                          "ms_max_over_mean": ms_max / (sum(p[0] for p in per) / world)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="final", choices=sorted(WORKLOADS))
    ap.add_argument("--spp-per-step", type=int, default=256)
    ap.add_argument("--pool", type=int, default=1 << 21)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config5-steps", type=int, default=2, help="timed steps of the BASELINE configs[4] sub-record (0 = skip)")
    ap.add_argument("--partition", default="samples", choices=["samples", "tiles"])
    ap.add_argument("--variance-mask", action="store_true", help="per-pixel variance mask (adaptive subsampling, north_star 3)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference_arm(args, wl)
    return run_cuda_arm(args, wl)


if __name__ == "__main__":
    sys.exit(main())
