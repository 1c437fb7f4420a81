#!/bin/bash
# k_walk prefetch variants: A/B on the height field and the cow forced through k_walk
echo "== heightfield"; C5=1 timeout 900 bash tools/bench_variants.sh --workload heightfield --spp-per-step 32
echo "== cow through k_walk"; ARTB_TWO_PHASE=0 timeout 600 bash tools/bench_variants.sh --workload cow --spp-per-step 64
