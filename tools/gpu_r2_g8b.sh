#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout -k 5 600 $TR bench.py --gpus 8 --steps 32 --warmup 5 > gpurun_out/r02_bench_n8_200M_b.json 2> gpurun_out/r02_bench_n8_200M_b.err; echo "n8 rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r02_bench_n8_200M_b.err | tail -3
cut -c1-330 gpurun_out/r02_bench_n8_200M_b.json
