#!/bin/bash
# round-2 call A: packed sweep validation + timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deferred_headline.py tests/test_gpu_deferred.py -x -q --durations=8 2>&1 | tail -25
./build/ubench_f32x2 > gpurun_out/r02_ubench_f32x2.txt 2>&1; cat gpurun_out/r02_ubench_f32x2.txt
timeout 900 python tools/time_sweep.py fresh early parked verylong --scalar > gpurun_out/r02_time_sweep.txt 2>&1; cat gpurun_out/r02_time_sweep.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
