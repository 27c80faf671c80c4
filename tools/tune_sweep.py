#!/usr/bin/env python
"""Times ctr_opt_dense_sweep (Adam, 3.2e9 elements = config-2 fm_v) for each CTR_SWEEP_CFG."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from tf_repos_b200 import ops, engine
d = torch.device("cuda:0")
n = 3_200_000_000
ost = engine.OptimizerState("Adam", 5e-4, 1e-4, d)
var = torch.empty(n, device=d); ops.init_trunc_normal(var, 1e-4, 1)
m = torch.zeros(n, device=d); v = torch.zeros(n, device=d)
part = torch.zeros(ops.sweep_partials_count(), device=d)
ost.tick()
for _ in range(3): ops.opt_dense_sweep(ost.opt, var, m, v, ost.record(0), part)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.opt_dense_sweep(ost.opt, var, m, v, ost.record(0), part); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort()
print("cfg", %r, "median %%.3f ms  best %%.3f ms  -> %%.0f GB/s (median)" %% (ts[5], ts[0], n * 24 / ts[5] / 1e6))
'''
for cfg in sys.argv[1:] or ["0", "1", "2", "3"]:
    env = dict(os.environ, CTR_SWEEP_CFG=cfg)
    subprocess.run([sys.executable, "-c", CHILD % (ROOT, cfg)], env=env, check=False)
