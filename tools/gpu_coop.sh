#!/bin/bash
# L2 access-policy window over the path pool (a fraction of its lines persists), final scene and Cornell box
for mb in 0 80 48 24; do
  if [ $mb = 0 ]; then export ARTB_L2_PERSIST=0; else export ARTB_L2_PERSIST=3 ARTB_L2_PERSIST_MB=$mb; fi
  echo "== final, persisting MB=$mb"; timeout 600 bash tools/bench_variants.sh --workload final
done
for mb in 0 80; do
  if [ $mb = 0 ]; then export ARTB_L2_PERSIST=0; else export ARTB_L2_PERSIST=3 ARTB_L2_PERSIST_MB=$mb; fi
  echo "== cornell, persisting MB=$mb"; timeout 600 bash tools/bench_variants.sh --workload cornell
done
