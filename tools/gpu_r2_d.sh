#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== fc / gemm tests with the warp-specialised kernel"; timeout -k 5 180 python -m pytest tests/test_gpu_kernels.py -x -q -k "fc or gemm or mlp" 2>&1 | tail -5
echo "== gemm bench WS=0"; CTR_GEMM_WS=0 timeout -k 5 150 python tools/bench_gemm.py 2>&1 | tail -6 | tee gpurun_out/r02_bench_gemm_ws0.txt
echo "== gemm bench WS=1"; CTR_GEMM_WS=1 timeout -k 5 150 python tools/bench_gemm.py 2>&1 | tail -6 | tee gpurun_out/r02_bench_gemm_ws1.txt
echo "== model tests"; timeout -k 5 400 python -m pytest tests/test_gpu_deepfm.py tests/test_gpu_din.py tests/test_gpu_dcn.py tests/test_gpu_sharded.py tests/test_gpu_nfm_pnn_afm.py -x -q 2>&1 | tail -5
timeout -k 5 900 python bench.py > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_d.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_d.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "sweep", d["roofline"]["avg_launch_ms"], "frac", d["roofline"]["frac"],
      "launches", d["gpu_launches"], "steady", d.get("steady_state", {}).get("value"), "lazy", round(d["lazy"]["value"]), "infer", round(d["infer"]["value"]), "text", d.get("e2e_text", {}).get("value"),
      "dcn", d.get("configs[2]_dcn"), "din", d.get("configs[3]_din"), "cpu", d.get("cpu_baseline", {}).get("value"))
PY
timeout -k 5 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
