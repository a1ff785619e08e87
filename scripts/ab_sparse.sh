#!/bin/bash
# A/B of the sparse cost-volume backward on ONE box (closure time, lookup-backward / mark / build-backward kernel times)
cat > /tmp/ab_sparse_parse.py <<'PY'
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = {r["name"]: r["avg_us"] for r in d["kernels"]}
print(sys.argv[1], d["ms_per_step"], d["gpu_launches_per_step"], {n: v for n, v in k.items() if "corr" in n and ("backward" in n or "occupancy" in n)})
PY
for rep in 1 2 3; do
  for v in 0 1; do
    env PCFA_BWD_SPARSE=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline --universal-pairs 0 2>/dev/null | python /tmp/ab_sparse_parse.py PCFA_BWD_SPARSE=$v
  done
done
