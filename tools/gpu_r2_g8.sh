#!/bin/bash
# 8 GPUs: BASELINE.json configs[4] (1e9-row table, row-sharded) and the 200M-row scaling point
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
NCCL_DEBUG=INFO timeout -k 5 700 $TR bench.py --gpus 8 --steps 32 --warmup 5 --vocab 1000000000 > gpurun_out/r02_bench_n8_c5_1B.json 2> gpurun_out/r02_bench_n8_c5_1B.err; echo "c5 rc=$?"
grep -E "NVLS|Connected all|via P2P|nChannels|comm 0x.*rank 0" gpurun_out/r02_bench_n8_c5_1B.err | head -12 > gpurun_out/r02_nccl_info_n8.txt; grep -v "NCCL INFO" gpurun_out/r02_bench_n8_c5_1B.err | tail -5
cut -c1-1200 gpurun_out/r02_bench_n8_c5_1B.json
timeout -k 5 600 $TR bench.py --gpus 8 --steps 32 --warmup 5 > gpurun_out/r02_bench_n8_200M.json 2> gpurun_out/r02_bench_n8_200M.err; echo "n8 rc=$?"; tail -3 gpurun_out/r02_bench_n8_200M.err
cut -c1-1200 gpurun_out/r02_bench_n8_200M.json
