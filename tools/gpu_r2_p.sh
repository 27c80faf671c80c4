#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_din.py tests/test_gpu_deepfm.py tests/test_gpu_dcn.py tests/test_gpu_nfm_pnn_afm.py tests/test_gpu_batch_norm.py -x -q 2>&1 | tail -3
timeout -k 5 150 python tools/bench_gemm.py 2>&1 | tail -6 | tee gpurun_out/r02_bench_gemm_final.txt
timeout -k 5 200 python tools/bench_models.py din 2>&1 | tail -3 | tee gpurun_out/r02_bench_din.txt
