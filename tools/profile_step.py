#!/usr/bin/env python
"""One training step of a model between cudaProfilerStart/Stop, for `ncu --profile-from-start off --set full ...`
(north_star: every kernel evidenced by an ncu capture).  Shapes follow BASELINE.json's configs with the vocabulary cut
to 2e7 rows (the table sweeps are profiled at full size by tools/time_sweep.py).

  ncu --profile-from-start off --set full --clock-control none -k regex:<names> -o out python tools/profile_step.py deepfm
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tf_repos_b200 import ops, synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "deepfm"
dev = torch.device("cuda:0")
N = int(os.environ.get("VOCAB", 20_000_000))
kw = dict(l2_reg=1e-4, learning_rate=5e-4, optimizer="Adam", update_mode="exact_deferred", epoch_steps=4, device=dev)


def run(step, n_warm=5):
    for i in range(n_warm):
        step(i)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step(n_warm)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if which in ("deepfm", "dcn", "pnn", "nfm", "afm", "deepfm_bn"):
    B, F = 8192, 39
    bt = [synth.criteo_batch(B, N, F, seed=i, device=dev) for i in range(8)]
    if which == "deepfm":
        from tf_repos_b200.deepfm import DeepFM
        m = DeepFM(F, N, 16, B, **kw)
    elif which == "deepfm_bn":
        from tf_repos_b200.deepfm import DeepFM
        m = DeepFM(F, N, 16, B, batch_norm=True, **kw)
    elif which == "dcn":
        from tf_repos_b200.dcn import DCN
        m = DCN(F, N, 16, B, cross_layers=6, **kw)
    elif which == "pnn":
        from tf_repos_b200.pnn import PNN
        m = PNN(F, N, 16, B, model_type="Inner", **kw)
    elif which == "nfm":
        from tf_repos_b200.nfm import NFM
        m = NFM(F, N, 64, B, **{**kw, "l2_reg": 1e-3})
    else:
        from tf_repos_b200.afm import AFM
        m = AFM(F, N, 64, 1024, **{**kw, "l2_reg": 1e-3})
        bt = [synth.criteo_batch(1024, N, F, seed=i, device=dev) for i in range(8)]
    run(lambda i: m.train_step(*bt[i % 8]))
elif which == "din":
    from tf_repos_b200.din import DIN
    B, Fp, K, P = 4096, 11, 32, 100
    bt = []
    for i in range(4):
        b, l = synth.din_batch(B, N, Fp, P, 8, seed=i)
        bt.append(({k: v.to(dev) for k, v in b.items()}, l.to(dev)))
    m = DIN(Fp, N, K, B, P, max_a_int=8, **kw)
    run(lambda i: m.train_step(*bt[i % 4]), n_warm=4)
elif which == "libsvm":
    import io
    ids, vals, labels = synth.criteo_batch(65536, N, 39, seed=1)
    buf = io.StringIO()
    for r in range(ids.shape[0]):
        buf.write("%d " % int(labels[r]))
        buf.write(" ".join("%d:%s" % (int(i), ("%.6f" % v) if v != 1.0 else "1") for i, v in zip(ids[r].tolist(), vals[r].tolist())))
        buf.write("\n")
    text = torch.frombuffer(bytearray(buf.getvalue().encode()), dtype=torch.uint8).to(dev)
    run(lambda i: ops.parse_libsvm_device(text, 39, 65536, final_chunk=True), n_warm=2)
elif which == "wide_deep":
    from tf_repos_b200.wide_deep import WideDeep
    B = 4096
    m = WideDeep(32, B, device=dev)
    g = torch.Generator().manual_seed(0)
    dense = torch.rand(B, 13, generator=g).to(dev)
    cat = torch.randint(0, 10_000, (B, 26), generator=g, dtype=torch.int32).to(dev)
    labels = (torch.rand(B, generator=g) < 0.25).float().to(dev)
    run(lambda i: m.train_step(dense, cat, labels), n_warm=3)
print("done", which)
