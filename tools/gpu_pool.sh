# bench.py over pool sizes (path slots) — per-kernel ms per wave and rays per wave beside the value
B="--steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 0 --config5-steps 0"
for p in ${POOLS:-1572864 1703936 1835008 1966080 2097152 2228224}; do
  printf "pool=%s " $p
  python bench.py $B --workload ${WL:-final} --pool $p 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(f\"{d['value']:8.1f} Msamples/s step {d['ms_per_step']:.2f} rays/wave {r['rays_per_launch']:.0f} trav {r['traverse_ms_per_launch']:.4f} pairs {r['pairs_ms_per_launch']:.4f} fin {r['finish_ms_per_launch']:.4f}\")"
done
