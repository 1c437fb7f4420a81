#!/bin/bash
# streaming cache hints on the pair buffer: A/B on final, cow, cornell
for w in final cow cornell; do echo "== $w"; timeout 600 bash tools/bench_variants.sh --workload $w; done
