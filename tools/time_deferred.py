#!/usr/bin/env python
"""Times DeepFM train steps at config 2 for update_mode exact vs exact_deferred (several epoch lengths)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tf_repos_b200 import synth  # noqa: E402
from tf_repos_b200.deepfm import DeepFM  # noqa: E402

N, B, F, K = int(os.environ.get("VOCAB", 200_000_000)), 8192, 39, 16
dev = torch.device("cuda:0")
batches = [synth.criteo_batch(B, N, F, seed=i, device=dev) for i in range(8)]
for mode, P in [("exact", 1)] + [("exact_deferred", p) for p in (int(x) for x in sys.argv[1:] or [4, 8, 16])]:
    m = DeepFM(F, N, K, B, update_mode=mode, epoch_steps=P, device=dev)
    m.updater.sweep_events = []
    steps = max(2 * P, 8)
    for i in range(P + 2 if mode != "exact" else 3):
        m.train_step(*batches[i % 8])
    m.flush() if mode != "exact" else None
    # align to an epoch boundary
    while m.epoch_pos != 0:
        m.train_step(*batches[0])
    m.updater.sweep_events = []
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        m.train_step(*batches[i % 8])
    e1.record()
    torch.cuda.synchronize()
    sw = [a.elapsed_time(b) for a, b in m.updater.sweep_events]
    ms = e0.elapsed_time(e1) / steps
    print(f"{mode:15s} P={P:2d}: {ms:8.3f} ms/step  {B / ms * 1e3 / 1e6:6.3f} M samples/s ; fm_v sweeps: "
          f"{len(sw)} x {sum(sw) / max(len(sw), 1):.2f} ms", flush=True)
    del m
    torch.cuda.empty_cache()
