#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 32 --warmup 5 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "n4 rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r02_bench_n4.err | tail -3; cut -c1-330 gpurun_out/r02_bench_n4.json
