#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 500 python -m pytest tests/test_gpu_sharded.py -q 2>&1 | tail -5 | tee gpurun_out/r02_sharded_2gpu_pytest_graphs.txt
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 32 --warmup 5 > gpurun_out/r02_bench_n2b.json 2> gpurun_out/r02_bench_n2b.err; echo "bench2 rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r02_bench_n2b.err | tail -3; cut -c1-400 gpurun_out/r02_bench_n2b.json
