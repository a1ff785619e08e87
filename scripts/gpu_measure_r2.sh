#!/bin/bash
# Round-2 evidence run (one GPU): bench line, ncu launch list of the same command, closure breakdowns, other configs.
set -x
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2600 --csv --log-file gpurun_out/launches_bench_r2.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --universal-pairs 0 > gpurun_out/bench_under_ncu.log 2>&1
python scripts/summarise_launches.py gpurun_out/launches_bench_r2.csv > gpurun_out/launches_bench_r2_summary.txt 2>&1
python scripts/profile_closure.py cl 70 RAFT 1 > gpurun_out/closure_kernels_raft_r2.txt 2>&1
python scripts/profile_closure.py nchw 45 GMA 1 > gpurun_out/closure_kernels_gma_r2.txt 2>&1
python scripts/bench_configs.py > gpurun_out/bench_configs_r2.log 2>&1
python scripts/bench_vs_reference_gpu.py > gpurun_out/bvr_r2.log 2>&1
python scripts/attack_breakdown.py > gpurun_out/attack_breakdown_r2.txt 2>&1
for b in 1 8; do for i in 1 2; do B=$b PCFA_LOOKUP_IMPL=$i python scripts/bench_lookup.py; done; done > gpurun_out/bench_lookup_r2.txt 2>&1
tail -3 gpurun_out/bench_configs_r2.log
# one --set full capture of the correlation kernels (a single closure inside a profiler range)
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'corr_lookup|corr_pyramid|prep_targets|bw_prep|bw_unpool|occ_mark' -o gpurun_out/corr_r2 -f python scripts/run_closure.py > gpurun_out/ncu_corr_r2.log 2>&1
python scripts/g_sparsity.py 1.0 > gpurun_out/g_sparsity_r2.txt 2>&1; python scripts/g_sparsity.py 0.5 >> gpurun_out/g_sparsity_r2.txt 2>&1
