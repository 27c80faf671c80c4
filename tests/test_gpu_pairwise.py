"""K7/K8 interaction ops (PNN inner/outer product, AFM pairwise attention pooling) vs fp64 autograd."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(got, ref, rtol=1e-5, what=""):
    got = got.detach().cpu().double().numpy(); ref = ref.detach().cpu().double().numpy()
    s = max(float(np.abs(ref).max()), 1e-30)
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol * s, err_msg=what)


def _pairs(F):
    row, col = [], []
    for i in range(F - 1):              # PNN.py:144-147
        for j in range(i + 1, F):
            row.append(i); col.append(j)
    return row, col


@pytest.mark.parametrize("F,K", [(39, 16), (39, 8), (5, 4), (2, 32), (13, 10)])
@pytest.mark.parametrize("outer", [False, True])
def test_pnn_product_fwd_bwd(F, K, outer):
    from tf_repos_b200 import ops
    if outer and F * K > 200:
        F = 7      # keep the outer tail (P*K*K floats/sample) small
    d = torch.device("cuda:0")
    B = 19
    P = F * (F - 1) // 2
    g = torch.Generator().manual_seed(F * K)
    x = torch.randn(B, F * K, generator=g)
    xd = x.double().requires_grad_()
    e = xd.reshape(B, F, K)
    row, col = _pairs(F)
    p, q = e[:, row], e[:, col]
    if outer:
        tail = torch.einsum("api,apj->apij", p, q).reshape(B, P * K * K)          # PNN.py:166
    else:
        tail = (p * q).sum(-1).reshape(B, P)                                       # PNN.py:152
    z_ref = torch.cat([xd, tail], 1)
    dz = torch.randn(z_ref.shape, generator=g)
    z_ref.backward(dz.double())
    z = torch.empty(z_ref.shape, device=d)
    ops.pnn_product_fwd(x.to(d), B, F, K, outer, z)
    _close(z, z_ref, what="z")
    assert torch.equal(z[:, :F * K].cpu(), x)
    dX = torch.empty(B, F * K, device=d)
    ops.pnn_product_bwd(x.to(d), dz.to(d), B, F, K, outer, dX)
    _close(dX, xd.grad, what="dX")


@pytest.mark.parametrize("F,K,drop", [(39, 16, False), (10, 256, True), (3, 8, True)])
def test_afm_pairs_and_pool(F, K, drop):
    from tf_repos_b200 import ops
    d = torch.device("cuda:0")
    B = 11
    P = F * (F - 1) // 2
    g = torch.Generator().manual_seed(F + K)
    x = torch.randn(B, F * K, generator=g) * 0.7
    logit = torch.randn(B, P, generator=g)
    keep = 0.8
    mask = (torch.rand(B, P, generator=g) < keep).float() if drop else None
    xd, ld = x.double().requires_grad_(), logit.double().requires_grad_()
    e = xd.reshape(B, F, K)
    row, col = _pairs(F)
    pw_ref = e[:, row] * e[:, col]                                                 # AFM.py:134-138
    att = torch.softmax(ld, dim=1)                                                 # :151
    w = att / keep * mask.double() if drop else att                                # :152-153
    y_ref = (w.unsqueeze(-1) * pw_ref).sum(1)                                      # :156
    dy = torch.randn(B, K, generator=g)
    y_ref.backward(dy.double())
    pw = torch.empty(B * P, K, device=d)
    ops.afm_pairs_fwd(x.to(d), B, F, K, pw)
    _close(pw, pw_ref.reshape(B * P, K), what="pw")
    att_o = torch.empty(B * P, device=d); y = torch.empty(B, K, device=d)
    mk = mask.to(d).reshape(-1) if drop else None
    ops.afm_pool_fwd(pw, logit.to(d).reshape(-1), mk, keep, B, P, K, att_o, y)
    _close(att_o, att.reshape(-1), what="softmax"); _close(y, y_ref, what="y_emb")
    dpw = torch.empty(B * P, K, device=d); dlogit = torch.empty(B * P, device=d)
    ops.afm_pool_bwd(pw, att_o, mk, keep, dy.to(d), B, P, K, dpw, dlogit)
    _close(dlogit, ld.grad.reshape(-1), what="dlogit")
    dX = torch.empty(B, F * K, device=d)
    ops.afm_pairs_bwd(x.to(d), dpw, B, F, K, dX)
    _close(dX, xd.grad, what="dX through pool + pairs")
    # equal logits => mean of the pair products
    ops.afm_pool_fwd(pw, torch.zeros(B * P, device=d), None, 1.0, B, P, K, att_o, y)
    _close(y, pw_ref.mean(1), what="equal logits -> mean")


def test_dropout_apply():
    from tf_repos_b200 import ops
    d = torch.device("cuda:0")
    x = torch.randn(1000, device=d); m = (torch.rand(1000, device=d) < 0.5).float(); o = torch.empty_like(x)
    ops.dropout_apply(x, m, 0.5, o)
    assert torch.equal(o, x / 0.5 * m)
