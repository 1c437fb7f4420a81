#!/bin/bash
# scene table of the final build (2 steps of 256 spp, heightfield 32 spp)
for w in final cornell cow dino random; do printf "%-12s" $w; timeout 600 bash tools/bench_variants.sh --workload $w | cut -c46-; done
printf "%-12s" heightfield; C5=2 timeout 600 bash tools/bench_variants.sh --workload heightfield --spp-per-step 32 | cut -c46-
