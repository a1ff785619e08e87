#!/bin/bash
# lookup kernels: correctness + ncu device time / instruction counts for both generations, and the microbench
python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_ops.py -q -m gpu -k "lookup or corrblock" 2>&1 | tail -3
python -m pytest tests/test_gpu_net.py -q -m gpu -k "channels_last" 2>&1 | tail -2
for i in 1 2; do
  PCFA_LOOKUP_IMPL=$i ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,sm__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:lookup -s 20 -c 4 --csv --log-file gpurun_out/lk_m$i.csv python scripts/bench_lookup.py > /dev/null 2>&1
done
B=8 PCFA_LOOKUP_IMPL=2 python scripts/bench_lookup.py
PCFA_LOOKUP_IMPL=2 python scripts/bench_lookup.py
