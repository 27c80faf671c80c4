import torch, numpy as np, sys
sys.path.insert(0, '.')
from oracle import tf_semantics as tfs
from tf_repos_b200 import engine, ops
d = torch.device('cuda:0')
N, n, lr, l2, K = 997, 300, 0.01, 1e-3, 1
g = torch.Generator().manual_seed(K)
ost = engine.OptimizerState('Adagrad', lr, l2, d)
var = (torch.randn(N, K, generator=g) * 0.1).float()
acc = (torch.rand(N, K, generator=g) * 0.01 + 1e-8).float()
for step in range(3):
    dv, da = var.to(d), acc.to(d)
    part = torch.zeros(ops.sweep_partials_count(), device=d)
    ops.opt_dense_sweep(ost.opt, dv, da, None, ost.record(0), part)
    G = torch.tensor(l2) * var
    rv, ra = var.clone(), acc.clone()
    tfs.adagrad_(rv, ra, G, torch.tensor(lr))
    bad = (dv.cpu() != rv).view(-1).nonzero().view(-1)
    print('step', step, 'mismatch var', bad.numel(), 'acc', (da.cpu() != ra).sum().item())
    for i in bad[:5].tolist():
        v, a = np.float32(var.view(-1)[i]), np.float32(acc.view(-1)[i])
        gg = np.float32(l2) * v
        a2 = a + gg * gg
        r = np.float32(1) / np.sqrt(a2)
        np_v = v - (np.float32(lr) * gg) * r
        print(i, 'gpu %.9e cpu %.9e numpy %.9e' % (dv.view(-1)[i].item(), rv.view(-1)[i].item(), np_v), 'acc gpu %.9e cpu %.9e np %.9e' % (da.view(-1)[i].item(), ra.view(-1)[i].item(), a2))
        x = torch.tensor([a2]); print('  torch 1/sqrt: %.9e  np %.9e  torch.rsqrt %.9e' % ((torch.ones(())/torch.sqrt(x)).item(), r, torch.rsqrt(x).item()))
    var, acc = rv, ra
