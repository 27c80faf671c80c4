#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deferred_headline.py tests/test_gpu_batch_norm.py tests/test_gpu_deferred.py -x -q 2>&1 | tail -12
timeout 300 python tools/time_sweep.py fresh parked 2>&1 | tail -2 | tee gpurun_out/r02_time_sweep_c.txt
timeout 900 python bench.py > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_c.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_c.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "sweep", d["roofline"]["avg_launch_ms"], "frac", d["roofline"]["frac"],
      "launches", d["gpu_launches"], "steady", d.get("steady_state"), "lazy", round(d["lazy"]["value"]), "infer", round(d["infer"]["value"]), "text", d.get("e2e_text", {}).get("value"), "cpu", d.get("cpu_baseline", {}).get("value"))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_c.csv \
    python bench.py --steps 16 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_ncu_list.log 2>&1; echo "ncu list rc=$?"
CTR_BENCH_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_c_eager.csv \
    python bench.py --steps 16 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_ncu_list_eager.log 2>&1; echo "ncu list eager rc=$?"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
