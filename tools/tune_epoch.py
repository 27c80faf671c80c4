#!/usr/bin/env python
"""Times ctr_epoch_sweep (Adam, config-2 fm_v: 2e8 x 16, 16 steps per pass) for each CTR_EPOCH_CFG."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from tf_repos_b200 import ops, engine
d = torch.device("cuda:0")
N, K, P = 200_000_000, 16, 16
ost = engine.OptimizerState("Adam", 5e-4, 1e-4, d)
var = torch.empty(N * K, device=d); ops.init_trunc_normal(var, 1e-4, 1)
m = torch.zeros(N * K, device=d); v = torch.zeros(N * K, device=d)
last = torch.zeros(N, dtype=torch.uint8, device=d)
part = torch.zeros(ops.epoch_max_steps() * ops.epoch_partials_count(), dtype=torch.float64, device=d)
for j in range(P): ost.tick_epoch(j)
# like a real epoch: ~1.6 %% of the rows were gathered by some batch and are already past some step
touched = torch.randint(0, N, (16 * 200_000,), device=d)
lvals = torch.randint(1, P + 1, (touched.numel(),), device=d, dtype=torch.uint8)
ts = []
for _ in range(4):
    last[touched] = lvals
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.epoch_sweep(ost.opt, var, m, v, last, N, K, ost.record(0), ost.lr_table, P, True, part); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort()
print("cfg", %r, "median %%.2f ms best %%.2f ms per %%d-step pass -> %%.3f ms per step" %% (ts[len(ts)//2], ts[0], P, ts[0] / P))
'''
for cfg in sys.argv[1:] or ["0", "1", "2", "3"]:
    env = dict(os.environ, CTR_EPOCH_CFG=cfg)
    subprocess.run([sys.executable, "-c", CHILD % (ROOT, cfg)], env=env, check=False)
