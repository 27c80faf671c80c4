#!/usr/bin/env python
"""In-situ device time per C-ABI entry point during DeepFM train steps (CUDA events around every call;
warm caches, real step context -- unlike the ncu launch list, whose times are cold-cache and serialised).

  python tools/time_phases.py [mode] [steps]      mode: exact_deferred (default) | exact | lazy
"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tf_repos_b200 import ops, synth  # noqa: E402
from tf_repos_b200.deepfm import DeepFM  # noqa: E402


class TimedLib:
    def __init__(self, lib):
        self._lib, self.records, self.on = lib, [], False

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("ctr_") or name in ("ctr_last_error", "ctr_launch_count"):
            return fn

        def call(*a):
            if not self.on:
                return fn(*a)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a)
            e1.record()
            key = name
            if name == "ctr_epoch_rows":   # (opt, apply, ..., K at 10, ..., j at 13)
                key = f"{name}[apply={a[1]},K={a[10]},j={'lo' if a[13] < 8 else 'hi'}]"
            self.records.append((key, e0, e1))
            return r
        return call


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "exact_deferred"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    N, B, F, K = int(os.environ.get("VOCAB", 200_000_000)), 8192, 39, 16
    dev = torch.device("cuda:0")
    tl = TimedLib(ops._L)
    ops._L = tl
    batches = [synth.criteo_batch(B, N, F, seed=i, device=dev) for i in range(8)]
    m = DeepFM(F, N, K, B, update_mode=mode, epoch_steps=16, device=dev)
    for i in range(5):
        m.train_step(*batches[i % 8])
    while mode == "exact_deferred" and m.epoch_pos != 0:
        m.train_step(*batches[0])
    torch.cuda.synchronize()
    tl.on = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        m.train_step(*batches[i % 8])
    e1.record()
    torch.cuda.synchronize()
    tl.on = False
    total = e0.elapsed_time(e1)
    agg = collections.OrderedDict()
    for name, a, b in tl.records:
        t = agg.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += a.elapsed_time(b)
    print(f"# DeepFM c2 {mode}: {steps} steps, {total / steps:.3f} ms/step (with event overhead)")
    print(f"# {'entry point':32s} {'calls/step':>10s} {'us/call':>10s} {'us/step':>10s} {'share':>7s}")
    acc = 0.0
    for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        acc += ms
        print(f"  {name:32s} {cnt / steps:10.2f} {ms / cnt * 1e3:10.1f} {ms / steps * 1e3:10.1f} {ms / total * 100:6.1f}%")
    print(f"# inside entry points: {acc / steps:.3f} ms/step; outside (torch ops, gaps): {(total - acc) / steps:.3f} ms/step")


if __name__ == "__main__":
    main()
