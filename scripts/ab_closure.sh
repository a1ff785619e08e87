#!/bin/bash
# A/B of the closure time on ONE box: alternate settings, several processes each (cuDNN autotuning varies per process)
for rep in 1 2 3; do
  for v in 0 1; do
    PCFA_CONV_ACT=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline --universal-pairs 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('conv_act=$v', d['ms_per_step'], d['gpu_launches_per_step'])"
  done
done
