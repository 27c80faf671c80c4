import torch, numpy as np, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import tf_semantics as tfs
from tf_repos_b200 import engine, ops
from test_gpu_kernels import _oracle_rows
d = torch.device('cuda:0')
def run(opt_name, K):
    N, n, lr, l2 = 997, 300, 0.01, 1e-3
    g = torch.Generator().manual_seed(K)
    ost = engine.OptimizerState(opt_name, lr, l2, d)
    adam = tfs.AdamHyper(lr)
    var = (torch.randn(N, K, generator=g) * 0.1).float()
    ns = ost.n_slots
    slots = [(torch.rand(N, K, generator=g) * 0.01 + ost.slot_init(i)).float() for i in range(ns)]
    uniq = torch.sort(torch.randperm(N, generator=g)[:n])[0].to(torch.int32)
    g_uniq = (torch.randn(n, K, generator=g) * 0.05).float()
    touched = torch.zeros(N, dtype=torch.bool); touched[uniq.long()] = True
    for step in range(2):
        ost.tick()
        dv, ds = var.to(d), [s.to(d) for s in slots]
        s1 = ds[1] if ns > 1 else None
        n_uniq = torch.tensor([n], dtype=torch.int32, device=d)
        stage = torch.zeros(3 * n * K, device=d)
        ops.opt_sparse_rows(ost.opt, dv, ds[0], s1, uniq.to(d), n_uniq, g_uniq.to(d), n, K, ost.record(0), stage)
        part = torch.zeros(ops.sweep_partials_count(), device=d)
        ops.opt_dense_sweep(ost.opt, dv, ds[0], s1, ost.record(0), part)
        ops.opt_patch_rows(dv, ds[0], s1, uniq.to(d), n_uniq, stage, n, K, ns)
        G = torch.tensor(l2) * var
        G[uniq.long()] = g_uniq + G[uniq.long()]
        rv, rs = _oracle_rows(opt_name, var, slots, G, None, lr, True, adam)
        bad = (dv.cpu() != rv).nonzero()
        print(opt_name, K, 'step', step, 'var mismatches', bad.shape[0], 'of which touched rows', int(touched[bad[:,0]].sum()),
              'slot mismatches', [int((a.cpu()!=b).sum()) for a,b in zip(ds, rs)])
        for r, c in bad[:4].tolist():
            print('   row', r, 'touched', bool(touched[r]), 'gpu %.9e cpu %.9e  var_old %.9e G %.9e slot0_old %.9e' % (dv[r,c].item(), rv[r,c].item(), var[r,c].item(), G[r,c].item(), slots[0][r,c].item()))
            if opt_name == 'Adagrad':
                v, a, gg = np.float32(var[r,c]), np.float32(slots[0][r,c]), np.float32(G[r,c])
                a2 = np.float32(a + gg*gg); rr = np.float32(1)/np.sqrt(a2)
                print('      numpy: %.9e ; acc new gpu %.9e cpu %.9e np %.9e' % (v - (np.float32(lr)*gg)*rr, ds[0][r,c].item(), rs[0][r,c].item(), a2))
        var, slots = rv, rs
        adam.finish()
run('Adagrad', 1); run('Adam', 16); run('Momentum', 16)
