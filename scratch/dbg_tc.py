import sys, torch
sys.path.insert(0, '.')
from tf_repos_b200 import ops
d = torch.device("cuda:0")
torch.manual_seed(0)
for (M, Kd, Nd) in [(8192, 624, 256), (8192, 256, 128), (1000, 128, 64), (256, 312, 400)]:
    x = torch.randn(M, Kd, device=d); W = torch.randn(Kd, Nd, device=d) / Kd ** 0.5; b = torch.randn(Nd, device=d) * 0.1
    out = torch.empty(M, Nd, device=d)
    ops.fc_fwd(x, W, b, None, 1.0, 1, out)
    ref = torch.relu(x.double() @ W.double() + b.double())
    err = (out.double() - ref).abs()
    print(f"fwd {M}x{Kd}x{Nd}: max err {err.max().item():.3e} (scale {ref.abs().max().item():.2f}); bad>1e-4: {(err > 1e-4).sum().item()}")
    for trial in range(3):
        dOut = torch.randn(M, Nd, device=d)
        dZ_ref = dOut.double() * (out > 0)
        dIn = torch.empty(M, Kd, device=d); dW = torch.empty(Kd, Nd, device=d); db = torch.empty(Nd, device=d)
        ws = torch.empty(ops.fc_bwd_workspace_bytes(M, Kd, Nd), dtype=torch.uint8, device=d)
        dO = dOut.clone()
        ops.fc_bwd(x, W, out, None, 1.0, dO, 1, dIn, dW, db, ws)
        e1 = (dIn.double() - dZ_ref @ W.double().t()).abs()
        e2 = (dW.double() - x.double().t() @ dZ_ref).abs()
        bad = (e1 > 1e-3).nonzero()
        print(f"  bwd trial {trial}: dIn max {e1.max().item():.3e} bad {bad.shape[0]} rows {sorted(set(bad[:,0].tolist()))[:8]} cols[min,max] "
              f"{(bad[:,1].min().item(), bad[:,1].max().item()) if bad.numel() else None}; dW max {e2.max().item():.3e} bad {(e2 > 1e-2).sum().item()}")
