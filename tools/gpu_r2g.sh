#!/bin/bash
# GPU session r2g: media tested as (ray, medium) pairs in k_pairs vs ray by ray in k_finish
mkdir -p gpurun_out
summ() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline'] or {}
print(f\"{d['value']:8.1f} Msamples/s  {d['mrays_per_sec']:7.0f} Mrays/s step {d['ms_per_step']:7.2f} ms  trav {r.get('traverse_ms_per_launch',0):.4f} pairs {r.get('pairs_ms_per_launch',0):.4f} fin {r.get('finish_ms_per_launch',0):.4f} checksum {d['image_checksum']}\")
"; }
B="--steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --config5-steps 0"
( for wl in final cow dino; do
    printf "%-6s media in k_pairs   " $wl; ARTB_MEDIA_PAIRS=1 python bench.py --workload $wl $B 2>&1 | summ
    printf "%-6s media in k_finish  " $wl; ARTB_MEDIA_PAIRS=0 python bench.py --workload $wl $B 2>&1 | summ
    printf "%-6s no medium code     " $wl; ARTB_LIB=$PWD/another_raytracer_b200/libartb_nomediacode.so python bench.py --workload $wl $B 2>&1 | summ
  done ) > gpurun_out/r2g_media.txt 2>&1
cat gpurun_out/r2g_media.txt
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2g_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_tests.log
tail -12 gpurun_out/r2g_tests.log
