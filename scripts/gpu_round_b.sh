#!/bin/bash
python -m pytest tests/test_gpu_net.py tests/test_gpu_parity_r2.py tests/test_gpu_attack.py -q -m gpu -x 2>&1 | tail -6
python -m pytest tests/test_gpu_ops.py -q -m gpu -k "softmax" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --universal-pairs 0 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -3 gpurun_out/bench_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'], d.get('attack'))
PY
