#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_din.py tests/test_gpu_din_cli.py tests/test_gpu_kernels.py -x -q 2>&1 | tail -4
timeout -k 5 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_din_step3.csv python tools/profile_step.py din > /dev/null 2>&1; echo "din list rc=$?"
python tools/summarize_launches.py gpurun_out/r02_launches_din_step3.csv > gpurun_out/r02_launches_din_step3.txt; head -14 gpurun_out/r02_launches_din_step3.txt; rm -f gpurun_out/r02_launches_din_step3.csv
timeout -k 5 200 python tools/bench_models.py din 2>&1 | tail -3
