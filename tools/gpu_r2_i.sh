#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_deferred_headline.py tests/test_gpu_deferred.py tests/test_gpu_kernels.py -x -q 2>&1 | tail -3
echo "K=1 sweep"; timeout -k 5 200 python tools/time_sweep.py fresh parked --k 1 2>&1 | tail -2 | tee gpurun_out/r02_time_sweep_k1_pf.txt
timeout -k 5 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_din_step.csv python tools/profile_step.py din > /dev/null 2>&1; echo "din list rc=$?"
python tools/summarize_launches.py gpurun_out/r02_launches_din_step.csv | head -45
timeout -k 5 900 python bench.py > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_i.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_i.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "sweep", d["roofline"]["avg_launch_ms"], "frac", d["roofline"]["frac"],
      "launches", d["gpu_launches"], "steady", d.get("steady_state", {}).get("value"), "lazy", round(d["lazy"]["value"]), "infer", round(d["infer"]["value"]), "text", d.get("e2e_text", {}).get("value"),
      "dcn", d.get("configs[2]_dcn", {}).get("value"), "din", d.get("configs[3]_din", {}).get("value"), "cpu", d.get("cpu_baseline", {}).get("value"))
PY
