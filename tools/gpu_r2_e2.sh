#!/bin/bash
# 2 GPUs: row-sharded tables -- NCCL parity test (log kept under profiles/) and a bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout -k 5 500 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_deferred.py -v 2>&1 | tail -40 | tee gpurun_out/r02_sharded_2gpu_pytest.txt
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 32 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r02_bench_n2.err; cat gpurun_out/r02_bench_n2.json | cut -c1-1500
