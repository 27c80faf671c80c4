#!/bin/bash
# round-2 call B: where does the packed sweep spend its time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
./build/ubench_f32x2 > gpurun_out/r02_ubench_f32x2.txt 2>&1; cat gpurun_out/r02_ubench_f32x2.txt
for mb in 2 3 4; do echo "MINB=$mb"; CTR_SWEEP_MINB=$mb timeout 300 python tools/time_sweep.py fresh parked 2>&1 | tail -2; done | tee gpurun_out/r02_time_sweep_minb.txt
echo "K=1"; timeout 300 python tools/time_sweep.py fresh parked --k 1 --scalar 2>&1 | tail -4 | tee gpurun_out/r02_time_sweep_k1.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:epoch_sweep_adam_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_sweep_adam \
    python tools/time_sweep.py fresh > gpurun_out/r02_ncu_sweep.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02_ncu_sweep.log
timeout 900 python bench.py > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; echo "bench rc=$?"; tail -5 gpurun_out/r02_bench_b.err; cat gpurun_out/r02_bench_b.json
timeout 600 python -m pytest tests/test_gpu_batch_norm.py -x -q 2>&1 | tail -15
timeout 600 python tools/bench_kernels.py 2>&1 | tee gpurun_out/r02_bench_k1_tma.txt
