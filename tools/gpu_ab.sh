for wl in ${WLS:-cow dino final}; do echo "== $wl"; bash tools/bench_variants.sh --workload $wl; done
