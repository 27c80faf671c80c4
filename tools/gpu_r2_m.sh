#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 180 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_din.py tests/test_gpu_din_cli.py -x -q 2>&1 | tail -3
CTR_GEMM_WS=1 timeout -k 5 150 python tools/bench_gemm.py 2>&1 | tail -6 | tee gpurun_out/r02_bench_gemm_ws1_pf3.txt
