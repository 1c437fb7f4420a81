"""bench.py's driver contract, the parts that run without a GPU: the reference arm (`--impl reference`: the UNMODIFIED reference's
CPU renderer, oracle/_ref/ref_engine_*, or the oracle port where that binary is missing) prints ONE JSON line with the keys the
driver reads, and the CUDA arm refuses to run without a device instead of falling back to anything."""
import json
import os
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent


def run_bench(*args, timeout=300):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", ""))
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


def test_reference_arm_prints_the_contract_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    if "unavailable" in d:          # never on a checkout where oracle/ was built; the contract allows it
        return
    assert d["metric"] == "samples_per_sec" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["dtype"] == "f64"
    assert d["config"]["workload"] == "next_week_final_scene_800x800" and d["config"]["width"] == 800 and d["config"]["max_depth"] == 50
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_cuda_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    r = run_bench("--steps", "1", "--warmup", "1", "--no-cpu-baseline", timeout=120)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]      # no bench line from a box without a GPU
