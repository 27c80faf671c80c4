#!/bin/bash
# One GPU-box visit: model/CLI/tokenizer parity tests, the bench line, the ncu launch list + full capture of the
# dominant kernel, the DCN/DIN side benches, smoke().  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_deepfm.py tests/test_gpu_dcn.py tests/test_gpu_nfm_pnn_afm.py tests/test_gpu_din.py \
    tests/test_gpu_cli.py tests/test_gpu_libsvm.py tests/test_gpu_wide_deep.py -q -x --durations=4 2>&1 | tail -14
timeout 300 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01_final.csv \
    python bench.py --steps 16 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench_final.log 2>&1; echo "ncu list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:epoch_sweep_kernel -c 1 -o gpurun_out/prof_epoch_r01_final -f \
    python tools/tune_epoch.py 4 > gpurun_out/ncu_full_final.log 2>&1; echo "ncu full rc=$?"
timeout 200 python tools/bench_models.py > gpurun_out/bench_models_final.json 2> gpurun_out/bench_models_final.err; echo "models rc=$?"
timeout 100 python -c "from __graft_entry__ import smoke; smoke(); print('smoke ok')" 2>&1 | tail -2
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_r01_final.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "sweep ms", round(d["roofline"]["avg_launch_ms"], 2),
      "lazy", round(d["lazy"]["value"]), "infer", round(d["infer"]["value"]), "exact", round(d["exact_every_step"]["value"]),
      "cpu", d.get("cpu_baseline", {}).get("value"), d["clocks"])
PY
