#!/bin/bash
# A/B of the closure time on ONE box: alternate settings of an environment switch, several processes each
# (cuDNN autotuning varies per process).  usage: ab_closure.sh VAR [value0 value1]
VAR=${1:-PCFA_CONV_ACT}; V0=${2:-0}; V1=${3:-1}
for rep in 1 2 3; do
  for v in $V0 $V1; do
    env $VAR=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline --universal-pairs 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k={r['name']:r['avg_us'] for r in d['kernels']}
print('$VAR=$v', d['ms_per_step'], d['gpu_launches_per_step'], d['attack']['outer_step_ms'], 'lookup fwd/bwd us', k.get('pcfa_corr_lookup_forward_cl'), k.get('pcfa_corr_lookup_backward_cl'))"
  done
done
