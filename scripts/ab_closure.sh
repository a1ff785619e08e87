#!/bin/bash
# A/B of the closure time on ONE box: alternate settings of an environment switch, several processes each
# (cuDNN autotuning varies per process).  usage: ab_closure.sh VAR
VAR=${1:-PCFA_CONV_ACT}
for rep in 1 2 3; do
  for v in 0 1; do
    env $VAR=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline --universal-pairs 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v', d['ms_per_step'], d['gpu_launches_per_step'], d['attack']['outer_step_ms'])"
  done
done
