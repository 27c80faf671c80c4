#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout -k 5 300 python tools/time_sweep.py fresh parked 2>&1 | tail -2 | tee gpurun_out/r02_time_sweep_h.txt
bash tools/gpu_ncu_all.sh 2>&1 | tail -12
du -sh gpurun_out
