#!/bin/bash
cd "$(dirname "$0")/.."
timeout -k 5 300 python -m pytest tests/test_gpu_deferred_headline.py -x -q -k "graph" 2>&1 | tail -3
timeout -k 5 300 python bench.py --no-cpu-baseline --steps 32 > gpurun_out/r02_bench_q.json 2> gpurun_out/r02_bench_q.err; echo "rc=$?"; tail -2 gpurun_out/r02_bench_q.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_q.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "infer", round(d["infer"]["value"]), "lazy", round(d["lazy"]["value"]))
PY
