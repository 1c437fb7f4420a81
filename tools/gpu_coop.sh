#!/bin/bash
# k_shade register budget: 3 / 4 / 5 / 6 blocks per SM
for w in final cornell; do echo "== $w"; timeout 600 bash tools/bench_variants.sh --workload $w; done
