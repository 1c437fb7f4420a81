#!/bin/bash
# GPU session r2l: k_walk (shared-memory stack, primitive-step vote thresholds) on the height field; walk parity tests
mkdir -p gpurun_out
summ() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(f\"{d['value']:8.1f} Msamples/s  {d['mrays_per_sec']:7.0f} Mrays/s step {d['ms_per_step']:7.2f} ms checksum {d['image_checksum']}\")
"; }
B="--steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --config5-steps 0"
( for lib in libartb libartb_vote14 libartb_vote13 libartb_vote23; do printf "heightfield %-14s " $lib; ARTB_LIB=$PWD/another_raytracer_b200/$lib.so python bench.py --workload heightfield --spp-per-step 32 $B 2>&1 | summ; done ) > gpurun_out/r2l_walk.txt 2>&1
cat gpurun_out/r2l_walk.txt
python -m pytest tests/test_gpu_walk_parity.py tests/test_gpu_bvh_build.py tests/test_gpu_full_size_properties.py -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -4
