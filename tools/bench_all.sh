#!/bin/bash
# usage: tools/bench_all.sh [extra bench.py args]   — every scene workload on one GPU, each next to the reference's CPU
# renderer timed on the box's host cores in the same run; one JSON line per workload on stdout
for w in random cornell final cow dino; do
  python bench.py --workload $w --steps 3 --warmup 3 --spp-per-step ${SPP:-128} "$@" 2>/dev/null | tail -1
done
