#!/bin/bash
# k_walk variants: slot-swapped refill, lane-count rule, against the previous build
echo "== heightfield"; C5=1 timeout 1200 bash tools/bench_variants.sh --workload heightfield --spp-per-step 32
