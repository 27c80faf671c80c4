#!/usr/bin/env python
"""K1 (ctr_fm_embed_fwd) with rows fetched by 128-bit loads (default) vs staged in shared memory by TMA bulk copies
(CTR_FM_EMBED_TMA=1), at the DeepFM config-2 shape and at wider rows.  Prints one JSON line per (K, variant) with the
CUDA-event time, the achieved HBM GB/s on the algorithmic bytes (ids + vals + rows + w + x + y) and a
checksum; the two variants must agree bit for bit (same arithmetic order).

  python tools/bench_kernels.py [--rows 50000000]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, torch
sys.path.insert(0, %(root)r)
from tf_repos_b200 import ops, synth
d = torch.device("cuda:0")
N, K, B, F = %(n)d, %(k)d, %(b)d, 39
V = torch.empty(N, K, device=d); ops.init_trunc_normal(V, 0.01, 1)
W = torch.empty(N, device=d); ops.init_trunc_normal(W, 0.01, 2)
batches = [synth.criteo_batch(B, N, F, seed=i, device=d) for i in range(8)]
x = torch.empty(B, F * K, device=d); yw = torch.empty(B, device=d); yv = torch.empty(B, device=d); S = torch.empty(B, K, device=d)
oob = torch.zeros(2, dtype=torch.int32, device=d)
def run(i):
    ids, vals, _ = batches[i %% 8]
    ops.fm_embed_fwd(ids, vals, V, W, ops.FM_DEEPFM, x=x, y_w=yw, y2=yv, S=S, oob=oob)
for i in range(5): run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 40
e0.record()
for i in range(reps): run(i)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
run(0); torch.cuda.synchronize()
alg = B * (F * (4 + 4 + 4 * K + 4) + 4 * F * K + 8 + 4 * K)
print(json.dumps({"kernel": "fm_embed_fwd", "variant": %(variant)r, "K": K, "B": B, "rows": N, "us": us,
                  "GBps": alg / us / 1e3, "bytes": alg,
                  "checksum": [float(x.double().sum()), float(yv.double().sum()), float(yw.double().sum())],
                  "x_hash": int(x.view(torch.int32).long().sum())}))
'''
rows = int(sys.argv[sys.argv.index("--rows") + 1]) if "--rows" in sys.argv else 50_000_000
out = []
for K, B in ((16, 8192), (32, 8192), (64, 4096), (128, 4096)):
    n = max(min(rows, int(12e9 // (K * 4))), 1000)
    res = {}
    for variant, env in (("ldg128", "0"), ("tma_bulk", "1")):
        r = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT, n=n, k=K, b=B, variant=variant)],
                           env=dict(os.environ, CTR_FM_EMBED_TMA=env), capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(json.dumps({"K": K, "variant": variant, "error": r.stderr[-400:]}))
            continue
        res[variant] = json.loads(line[-1])
        print(line[-1], flush=True)
    if len(res) == 2:
        same = res["ldg128"]["x_hash"] == res["tma_bulk"]["x_hash"] and res["ldg128"]["checksum"] == res["tma_bulk"]["checksum"]
        print(json.dumps({"K": K, "bit_identical": same, "tma_over_ldg_time": res["tma_bulk"]["us"] / res["ldg128"]["us"]}), flush=True)
