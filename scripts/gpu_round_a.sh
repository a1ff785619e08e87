#!/bin/bash
bash scripts/gpu_lookup_check.sh
for m in 0 1; do for f in randn corr; do PCFA_FWD_MODE=$m FEAT=$f python scripts/fwd_mode.py; done; done
PCFA_FWD_MODE=1 python -m pytest tests/test_gpu_ops.py tests/test_gpu_net.py tests/test_gpu_parity_r2.py -q -m gpu -x 2>&1 | tail -15
