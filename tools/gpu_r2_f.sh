#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_deferred_headline.py tests/test_gpu_deferred.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -4
for cfg in "3 0" "2 0" "2 1" "3 1" "4 0"; do set -- $cfg; echo "MINB=$1 PF=$2"; CTR_SWEEP_MINB=$1 CTR_SWEEP_PF=$2 timeout -k 5 300 python tools/time_sweep.py fresh 2>&1 | tail -1; done | tee gpurun_out/r02_time_sweep_pf.txt
bash tools/gpu_ncu_all.sh 2>&1 | tail -14
