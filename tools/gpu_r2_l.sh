#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 180 python -m pytest tests/test_gpu_kernels.py -x -q -k "fc or gemm or mlp" 2>&1 | tail -3
CTR_GEMM_WS=1 timeout -k 5 150 python tools/bench_gemm.py 2>&1 | tail -6 | tee gpurun_out/r02_bench_gemm_ws1_fast.txt
timeout -k 5 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout -k 5 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_l.json 2> gpurun_out/r02_bench_l.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_l.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_l.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "sweep", d["roofline"]["avg_launch_ms"],
      "steady", d.get("steady_state", {}).get("value"), "lazy", round(d["lazy"]["value"]), "infer", round(d["infer"]["value"]), "text", d.get("e2e_text", {}).get("value"),
      "dcn", d.get("configs[2]_dcn", {}).get("value"), "din", d.get("configs[3]_din", {}).get("value"))
PY
timeout -k 5 120 python -c "from __graft_entry__ import smoke; smoke()" 2>&1 | tail -2
