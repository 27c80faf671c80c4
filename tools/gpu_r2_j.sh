#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout -k 5 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_din_step2.csv python tools/profile_step.py din > /dev/null 2>&1; echo "din list rc=$?"
python tools/summarize_launches.py gpurun_out/r02_launches_din_step2.csv | head -16
timeout -k 5 300 ncu --profile-from-start off --set full --clock-control none -k regex:"bn_" -c 10 -f -o /tmp/bn python tools/profile_step.py deepfm_bn > /dev/null 2>&1; python tools/ncu_summary.py /tmp/bn.ncu-rep gpurun_out/r02_ncu_bn_kernels.txt | grep -E "^==|duration" 
timeout -k 5 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_j.json 2> gpurun_out/r02_bench_j.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_j.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_j.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "sweep", d["roofline"]["avg_launch_ms"], "frac", d["roofline"]["frac"],
      "launches", d["gpu_launches"], "steady", d.get("steady_state", {}).get("value"), "lazy", round(d["lazy"]["value"]), "infer", round(d["infer"]["value"]), "text", d.get("e2e_text", {}).get("value"),
      "dcn", d.get("configs[2]_dcn", {}).get("value"), "din", d.get("configs[3]_din", {}).get("value"))
PY
