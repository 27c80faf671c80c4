"""DCN (Deep & Cross, DCN.py:105-230): cross-network kernel parity and N-step end-to-end parity vs the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(got, ref, rtol, what=""):
    got = got.detach().cpu().double().numpy(); ref = ref.detach().cpu().double().numpy()
    s = max(float(np.abs(ref).max()), 1e-30)
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol * s, err_msg=what)


@pytest.mark.parametrize("D,L", [(624, 6), (312, 3), (1248, 2), (8, 1), (2048, 4), (100, 5)])
def test_cross_fwd_bwd_vs_autograd(D, L):
    from tf_repos_b200 import ops
    d = torch.device("cuda:0")
    B = 257
    g = torch.Generator().manual_seed(D + L)
    x0 = torch.randn(B, D, generator=g) * 0.5
    w = torch.randn(L, D, generator=g) / D ** 0.5
    b = torch.randn(L, D, generator=g) * 0.1
    dxL = torch.randn(B, D, generator=g)
    dx_in = torch.randn(B, D, generator=g)
    x0d, wd, bd = x0.double().requires_grad_(), w.double().requires_grad_(), b.double().requires_grad_()
    xl = x0d
    s_ref = []
    for l in range(L):
        xlw = xl @ wd[l].reshape(-1, 1)
        s_ref.append(xlw.reshape(-1))
        xl = x0d * xlw + xl + bd[l]
    xl.backward(dxL.double())
    xL = torch.empty(B, D, device=d); s = torch.empty(B, L, device=d)
    ops.cross_fwd(x0.to(d), w.to(d), b.to(d), xL, s)
    _close(xL, xl, 1e-5, "x_L")
    _close(s, torch.stack(s_ref, 1), 1e-5, "s")
    dx0 = torch.empty(B, D, device=d); dw = torch.empty(L, D, device=d); db = torch.empty(L, D, device=d)
    ws = torch.empty(ops.cross_bwd_workspace_bytes(B, D, L), dtype=torch.uint8, device=d)
    ops.cross_bwd(x0.to(d), w.to(d), b.to(d), s, dxL.to(d), dx_in.to(d), dx0, dw, db, ws)
    _close(dx0, x0d.grad + dx_in.double(), 1e-5, "dx0")
    _close(dw, wd.grad, 1e-5, "dw")
    _close(db, bd.grad, 1e-5, "db")
    dw2 = torch.empty_like(dw); db2 = torch.empty_like(db)
    ops.cross_bwd(x0.to(d), w.to(d), b.to(d), s, dxL.to(d), None, dx0, dw2, db2, ws)
    assert torch.equal(dw, dw2) and torch.equal(db, db2), "deterministic"
    _close(dx0, x0d.grad, 1e-5, "dx0 without dx_in")


def test_cross_known_answer_zero_w():
    """w = 0 => x_L = x0 + sum_l b_l."""
    from tf_repos_b200 import ops
    d = torch.device("cuda:0")
    B, D, L = 33, 624, 6
    x0 = torch.randn(B, D); b = torch.randn(L, D)
    xL = torch.empty(B, D, device=d); s = torch.empty(B, L, device=d)
    ops.cross_fwd(x0.to(d), torch.zeros(L, D, device=d), b.to(d), xL, s)
    ref = x0.clone()
    for l in range(L):
        ref = ref + b[l]
    assert torch.equal(xL.cpu(), ref) and torch.all(s == 0)


@pytest.mark.parametrize("opt,mode", [("Adam", "exact"), ("Adam", "exact_deferred"), ("Adagrad", "exact"), ("Adam", "lazy")])
def test_dcn_train_steps_match_oracle(opt, mode):
    from oracle import models as om
    from tf_repos_b200 import synth
    from tf_repos_b200.dcn import DCN
    B, N, K, F, L = 256, 10_000, 8, 39, 3
    lr = 5e-4 if opt == "Adam" else 0.01
    ref = om.DCN(F, N, K, deep_layers="64,32", cross_layers=L, dropout="1.0,1.0", l2_reg=1e-4, learning_rate=lr,
                 optimizer=opt, update_mode=("lazy" if mode == "lazy" else "exact"), seed=5)
    g = torch.Generator().manual_seed(9)
    ref.params["emb"].copy_(torch.randn(N, K, generator=g) * 0.1)
    gpu = DCN(F, N, K, B, deep_layers="64,32", cross_layers=L, dropout="1.0,1.0", l2_reg=1e-4, learning_rate=lr,
              optimizer=opt, update_mode=mode, epoch_steps=3, device="cuda:0")
    gpu.load_variables(ref.params)
    for step in range(4):
        ids, vals, labels = synth.criteo_batch(B, N, F, seed=200 + step)
        batch = {"feat_ids": ids.long(), "feat_vals": vals}
        prob = gpu.predict(ids.cuda(), vals.cuda())
        out = ref.predict(batch)
        _close(gpu.y[:B], out["y"], 1e-5, f"logits step {step}")
        loss_ref = ref.train_step(batch, labels)
        parts = gpu.train_step(ids.cuda(), vals.cuda(), labels.cuda())
        gpu.check_ids()
        if mode == "exact":
            assert abs(gpu.loss_value(parts) - loss_ref) <= 1e-5 * abs(loss_ref)
        vs = gpu.variables()
        for name in ("emb", "cross_w", "cross_b", "Deep-Network/mlp0/weights", "DCN-out/out_layer/weights"):
            _close(vs[name], ref.params[name], 2e-5, f"{name} after step {step} ({opt},{mode})")
