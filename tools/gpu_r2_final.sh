#!/bin/bash
# final validation on one B200: GPU tests, smoke, the bench line, the launch list of the same command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 700 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout -k 5 120 python -c "from __graft_entry__ import smoke; smoke()" 2>&1 | tail -2
timeout -k 5 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_final.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_final.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "sweep", d["roofline"]["avg_launch_ms"], "frac", d["roofline"]["frac"],
      "launches", d["gpu_launches"], "steady", d.get("steady_state", {}).get("value"), "lazy", round(d["lazy"]["value"]), "infer", round(d["infer"]["value"]), "text", d.get("e2e_text", {}).get("value"),
      "dcn", d.get("configs[2]_dcn", {}).get("value"), "din", d.get("configs[3]_din", {}).get("value"), "cpu", d.get("cpu_baseline", {}).get("value"), d["clocks"])
PY
timeout -k 5 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_final.csv \
    python bench.py --steps 16 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_ncu_list_final.log 2>&1; echo "ncu list rc=$?"
python tools/summarize_launches.py gpurun_out/r02_launches_final.csv > gpurun_out/r02_launches_final.txt; head -14 gpurun_out/r02_launches_final.txt; rm -f gpurun_out/r02_launches_final.csv
