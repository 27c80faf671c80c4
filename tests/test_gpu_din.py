"""DIN (DIN.py:101-257): kernel parity of the pooling pieces and N-step end-to-end parity vs the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(got, ref, rtol, what=""):
    got = got.detach().cpu().double().numpy(); ref = ref.detach().cpu().double().numpy()
    s = max(float(np.abs(ref).max()), 1e-30)
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol * s, err_msg=what)


def _cuda(batch):
    return {k: v.cuda() for k, v in batch.items()}


@pytest.mark.parametrize("K", [8, 32])
def test_gather_bag_pool_kernels(K):
    from tf_repos_b200 import ops, synth
    d = torch.device("cuda:0")
    B, P, N = 50, 13, 1000
    batch, _ = synth.din_batch(B, N, Fp=5, P=P, seed=K)
    g = torch.Generator().manual_seed(0)
    V = torch.randn(N, K, generator=g)
    # gather into a strided slice of a wider buffer
    Dx = 5 * K + 8
    x = torch.zeros(B, Dx, device=d)
    ops.gather_scale_rows(batch["feat_ids"].reshape(-1).to(d), None, V.to(d), x, 5, Dx)
    assert torch.equal(x[:, :5 * K].cpu(), V[batch["feat_ids"].long()].reshape(B, 5 * K)) and torch.all(x[:, 5 * K:] == 0)
    # weighted padded gather
    E = torch.empty(B * P, K, device=d)
    ids, wgt = batch["u_ids"][0].reshape(-1), batch["u_wgt"][0].reshape(-1)
    ops.gather_scale_rows(ids.to(d), wgt.to(d), V.to(d), E, 1, K)
    E_ref = V[ids.long()] * wgt[:, None]
    assert torch.equal(E.cpu(), E_ref)
    # bag sum fwd / bwd
    out = torch.zeros(B, K + 4, device=d)
    ops.bag_sum_fwd(batch["a_int_ids"].to(d), None, batch["a_int_off"].to(d), V.to(d), out[:, 4:], K + 4)
    off = batch["a_int_off"].long()
    seg = torch.repeat_interleave(torch.arange(B), off[1:] - off[:-1])
    ref = torch.zeros(B, K).index_add(0, seg, V[batch["a_int_ids"].long()])
    _close(out[:, 4:], ref, 1e-6)
    d_out = torch.randn(B, K, generator=g)
    g_rows = torch.empty(batch["a_int_ids"].numel(), K, device=d)
    ops.bag_sum_bwd(d_out.to(d), K, None, batch["a_int_off"].to(d), K, g_rows)
    assert torch.equal(g_rows.cpu(), d_out[seg])
    # attention pooling fwd / bwd vs autograd
    z = torch.randn(B * P, generator=g)
    Ed, zd = E_ref.double().requires_grad_(), z.double().requires_grad_()
    mask = (batch["u_ids"][0] > 0).double().unsqueeze(-1)
    u_ref = ((Ed.reshape(B, P, K) * torch.sigmoid(zd).reshape(B, P, 1)) * mask).sum(1)
    du = torch.randn(B, K, generator=g)
    u_ref.backward(du.double())
    att = torch.empty(B * P, device=d); u = torch.empty(B, K, device=d)
    ops.din_pool_fwd(E, z.to(d), ids.to(d), B, P, K, att, u, K)
    _close(u, u_ref, 1e-5); _close(att, torch.sigmoid(zd), 1e-6)
    dE = torch.empty(B * P, K, device=d); dz = torch.empty(B * P, device=d)
    ops.din_pool_bwd(E, att, ids.to(d), du.to(d), K, B, P, K, dE, dz)
    _close(dE, Ed.grad, 1e-5); _close(dz, zd.grad, 1e-5)


@pytest.mark.parametrize("opt,mode,attn", [("Adam", "exact", True), ("Adam", "exact_deferred", True),
                                           ("Adagrad", "lazy", True), ("Adam", "exact", False)])
def test_din_train_steps_match_oracle(opt, mode, attn):
    from oracle import models as om
    from tf_repos_b200 import synth
    from tf_repos_b200.din import DIN
    B, N, K, Fp, P = 64, 5000, 8, 11, 9
    lr = 5e-4 if opt == "Adam" else 0.01
    kw = dict(deep_layers="16,8", dropout="1.0,1.0", attention_layers="256", attention_pooling=attn, l2_reg=1e-4,
              learning_rate=lr, optimizer=opt)
    ref = om.DIN(Fp, N, K, update_mode=("lazy" if mode == "lazy" else "exact"), seed=4, **kw)
    g = torch.Generator().manual_seed(1)
    ref.params["embeddings"].copy_(torch.randn(N, K, generator=g) * 0.1)
    gpu = DIN(Fp, N, K, B, P, max_a_int=8, update_mode=mode, epoch_steps=3, device="cuda:0", **kw)
    gpu.load_variables(ref.params)
    for step in range(4):
        batch, labels = synth.din_batch(B, N, Fp, P, 8, seed=50 + step)
        lb = {k: (v.long() if v.dtype == torch.int32 else v) for k, v in batch.items()}
        gpu.predict(_cuda(batch))
        out = ref.predict(lb)
        _close(gpu.y, out["y"], 1e-5, f"logits step {step}")
        loss_ref = ref.train_step(lb, labels)
        parts = gpu.train_step(_cuda(batch), labels.cuda())
        gpu.check_ids()
        if mode == "exact":
            assert abs(gpu.loss_value(parts) - loss_ref) <= 1e-5 * abs(loss_ref)
        vs = gpu.variables()
        names = ["embeddings", "MLP-layer/mlp0/weights", "DIN-out/din_out/weights"]
        if attn:
            names += ["Field-wise-Pooling-layer/att_fc0/weights", "Field-wise-Pooling-layer/att_fc0/biases",
                      "Field-wise-Pooling-layer/att_out/weights", "Field-wise-Pooling-layer/att_out/biases"]
        for name in names:
            _close(vs[name], ref.params[name], 2e-5, f"{name} after step {step} ({opt},{mode},attn={attn})")


def test_din_single_position_known_answer():
    """P = 1: u = mask * sigmoid(att) * e  (SURVEY.md 8c)."""
    from tf_repos_b200 import ops
    d = torch.device("cuda:0")
    B, K = 7, 8
    E = torch.randn(B, K, device=d); z = torch.randn(B, device=d)
    ids = torch.tensor([3, 0, 5, 1, 0, 9, 2], dtype=torch.int32, device=d)
    att = torch.empty(B, device=d); u = torch.empty(B, K, device=d)
    ops.din_pool_fwd(E, z, ids, B, 1, K, att, u, K)
    ref = E * torch.sigmoid(z)[:, None] * (ids > 0).float()[:, None]
    assert torch.allclose(u, ref, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("mode", ["exact", "exact_deferred"])
def test_din_partial_final_batch_equals_the_smaller_batch(mode):
    """repeat-before-batch leaves one partial batch at the end of training (DIN.py:93-94): the CUDA model takes it padded
    to its buffer size with n_valid; the step must equal the oracle's step on the n_valid-sample batch."""
    from oracle import models as om
    from tf_repos_b200 import synth
    from tf_repos_b200.din import DIN
    B, n, N, K, Fp, P = 64, 23, 5000, 8, 11, 9
    kw = dict(deep_layers="16,8", dropout="1.0,1.0", attention_layers="256", attention_pooling=True, l2_reg=1e-4,
              learning_rate=5e-4, optimizer="Adam")
    ref = om.DIN(Fp, N, K, update_mode="exact", seed=4, **kw)
    g = torch.Generator().manual_seed(1)
    ref.params["embeddings"].copy_(torch.randn(N, K, generator=g) * 0.1)
    gpu = DIN(Fp, N, K, B, P, max_a_int=8, update_mode=mode, epoch_steps=3, device="cuda:0", **kw)
    gpu.load_variables(ref.params)
    # one full step, then the partial one
    batch, labels = synth.din_batch(B, N, Fp, P, 8, seed=7)
    ref.train_step({k: (v.long() if v.dtype == torch.int32 else v) for k, v in batch.items()}, labels)
    gpu.train_step(_cuda(batch), labels.cuda())
    small, lab_s = synth.din_batch(n, N, Fp, P, 8, seed=8)
    loss_ref = ref.train_step({k: (v.long() if v.dtype == torch.int32 else v) for k, v in small.items()}, lab_s)
    # pad to B with copies of sample 0 (what din_main.make_batch does)
    pad = B - n
    off = small["a_int_off"]
    first_bag = small["a_int_ids"][off[0]:off[1]]
    padded = {
        "feat_ids": torch.cat([small["feat_ids"], small["feat_ids"][:1].repeat(pad, 1)]),
        "a_ids": torch.cat([small["a_ids"], small["a_ids"][:, :1].repeat(1, pad)], dim=1),
        "a_int_ids": torch.cat([small["a_int_ids"], first_bag.repeat(pad)]),
        "a_int_off": torch.cat([off, off[-1] + (torch.arange(1, pad + 1, dtype=off.dtype) * first_bag.numel())]),
        "u_ids": torch.cat([small["u_ids"], small["u_ids"][:, :1].repeat(1, pad, 1)], dim=1),
        "u_wgt": torch.cat([small["u_wgt"], small["u_wgt"][:, :1].repeat(1, pad, 1)], dim=1),
    }
    labels_p = torch.cat([lab_s, lab_s[:1].repeat(pad)])
    parts = gpu.train_step(_cuda(padded), labels_p.cuda(), n_valid=n)
    if mode == "exact":
        assert abs(gpu.loss_value(parts) - loss_ref) <= 1e-5 * abs(loss_ref), (gpu.loss_value(parts), loss_ref)
    vs = gpu.variables()
    for name in ("embeddings", "MLP-layer/mlp0/weights", "DIN-out/din_out/weights", "Field-wise-Pooling-layer/att_fc0/weights",
                 "Field-wise-Pooling-layer/att_out/biases"):
        _close(vs[name], ref.params[name], 2e-5, f"{name} after the partial batch ({mode})")
