"""--batch_norm on the CUDA path (csrc/batch_norm.cu): batch_norm_layer of DeepFM.py:159-160,231-235 -- after the relu,
before the dropout; TRAIN uses batch moments and updates the moving statistics in place, EVAL/PREDICT uses the moving
statistics -- against oracle.tf_semantics.batch_norm (kernels vs an fp64 autograd twin, models vs the oracle models)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(got, ref, rtol, what):
    got = got.detach().cpu().double().numpy(); ref = ref.detach().cpu().double().numpy()
    scale = max(float(np.abs(ref).max()), 1e-30)
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol * scale, err_msg=what)


@pytest.mark.parametrize("n,H,keep", [(8192, 256, 0.5), (300, 64, 1.0), (77, 40, 0.8), (1, 16, 1.0)])
def test_bn_kernels_match_fp64_autograd(n, H, keep):
    from oracle import tf_semantics as tfs
    from tf_repos_b200 import ops
    g = torch.Generator().manual_seed(n + H)
    x = torch.relu(torch.randn(n, H, generator=g)) * 2.0
    gamma = 1.0 + 0.3 * torch.randn(H, generator=g); beta = 0.2 * torch.randn(H, generator=g)
    mm0 = torch.randn(H, generator=g) * 0.1; mv0 = 1.0 + 0.1 * torch.rand(H, generator=g)
    mask = (torch.rand(n, H, generator=g) < keep).float() if keep < 1.0 else None
    d_out = torch.randn(n, H, generator=g)
    # fp64 twin of the oracle's formula
    x64 = x.double().requires_grad_(True); g64 = gamma.double().requires_grad_(True); b64 = beta.double().requires_grad_(True)
    mm64, mv64 = mm0.double().clone(), mv0.double().clone()
    y64 = tfs.batch_norm(x64, g64, b64, mm64, mv64, True, 0.9)
    if mask is not None:
        y64 = y64 / keep * mask.double()
    y64.backward(d_out.double())
    dev = "cuda"
    xd, gd, bd = x.to(dev), gamma.to(dev), beta.to(dev)
    mm, mv = mm0.to(dev), mv0.to(dev)
    out = torch.empty(n, H, device=dev); sm = torch.empty(H, device=dev); sv = torch.empty(H, device=dev)
    md = mask.to(dev) if mask is not None else None
    ops.bn_fwd(xd, gd, bd, mm, mv, True, 0.9, md, keep, out, sm, sv)
    # n == 1: x == mean, and TF's x*inv + (beta - mean*inv) cancels to within an ulp of x*inv rather than exactly
    _close(out, y64.detach(), 2e-6 if n > 1 else 1e-5, "bn forward")
    _close(mm, mm64, 2e-6, "moving_mean"); _close(mv, mv64, 2e-6, "moving_variance")
    dx = torch.empty(n, H, device=dev); dg = torch.empty(H, device=dev); db = torch.empty(H, device=dev)
    ops.bn_bwd(d_out.to(dev), xd, sm, sv, gd, md, keep, dx, dg, db)
    tol = 2e-5 if n > 1 else 1e-3       # n == 1: var = 0, everything cancels against eps
    _close(dx, x64.grad, tol, "d_x"); _close(dg, g64.grad, tol, "d_gamma"); _close(db, b64.grad, 2e-6, "d_beta")
    # inference branch: moving statistics, no dropout, nothing else written
    mm_before = mm.clone()
    ops.bn_fwd(xd, gd, bd, mm, mv, False, 0.9, None, 1.0, out)
    y_inf = tfs.batch_norm(x.double(), gamma.double(), beta.double(), mm64, mv64, False, 0.9)
    _close(out, y_inf, 2e-6, "bn inference"); assert torch.equal(mm, mm_before)


def _deepfm_pair(B=256, N=10_000, K=8, dropout="1.0,1.0,1.0"):
    from oracle import models as om
    from tf_repos_b200.deepfm import DeepFM
    ref = om.DeepFM(39, N, K, deep_layers="64,32,16", dropout=dropout, l2_reg=1e-4, learning_rate=5e-4, optimizer="Adam",
                    update_mode="exact", seed=3, batch_norm=True, batch_norm_decay=0.9)
    g = torch.Generator().manual_seed(5)
    ref.params["fm_v"].copy_(torch.randn(N, K, generator=g) * 0.1)
    ref.params["fm_w"].copy_(torch.randn(N, generator=g) * 0.1)
    for i in range(3):   # non-trivial gamma / beta
        ref.params[f"Deep-part/bn_{i}/gamma"].copy_(1.0 + 0.2 * torch.randn_like(ref.params[f"Deep-part/bn_{i}/gamma"]))
        ref.params[f"Deep-part/bn_{i}/beta"].copy_(0.1 * torch.randn_like(ref.params[f"Deep-part/bn_{i}/beta"]))
    gpu = DeepFM(39, N, K, B, deep_layers="64,32,16", dropout=dropout, l2_reg=1e-4, learning_rate=5e-4, optimizer="Adam",
                 update_mode="exact", device="cuda:0", batch_norm=True, batch_norm_decay=0.9)
    gpu.load_variables(ref.params)
    return ref, gpu


def test_deepfm_with_batch_norm_matches_oracle():
    from tf_repos_b200 import synth
    B, N = 256, 10_000
    ref, gpu = _deepfm_pair(B, N)
    for step in range(3):
        ids, vals, labels = synth.criteo_batch(B, N, 39, seed=40 + step)
        batch = {"feat_ids": ids.long(), "feat_vals": vals}
        loss_ref = ref.train_step(batch, labels)
        parts = gpu.train_step(ids.cuda(), vals.cuda(), labels.cuda())
        assert abs(gpu.loss_value(parts) - loss_ref) <= 2e-5 * abs(loss_ref), (step, gpu.loss_value(parts), loss_ref)
        vs = gpu.variables()
        for name in ("fm_v", "fm_bias", "Deep-part/mlp0/weights", "Deep-part/bn_0/gamma", "Deep-part/bn_2/beta",
                     "Deep-part/deep_out/weights"):
            _close(vs[name], ref.params[name], 5e-5, f"{name} after step {step}")
        for name, v in ref.bn_state.items():
            _close(vs[name], v, 1e-5, f"{name} after step {step}")
    # PREDICT: moving statistics
    ids, vals, _ = synth.criteo_batch(B, N, 39, seed=99)
    prob = gpu.predict(ids.cuda(), vals.cuda())
    out = ref.predict({"feat_ids": ids.long(), "feat_vals": vals})
    _close(gpu.y[:B], out["y"], 2e-5, "eval logits"); _close(prob, out["prob"], 2e-5, "eval prob")


def test_batch_norm_checkpoint_round_trip_and_flag(tmp_path):
    """moving statistics travel with model.variables() (checkpoint / export / TF-name mapping)"""
    from tf_repos_b200 import synth, tf_names
    ref, gpu = _deepfm_pair(128, 5000)
    ids, vals, labels = synth.criteo_batch(128, 5000, 39, seed=1, device="cuda")
    gpu.train_step(ids, vals, labels)
    st = tf_names.state_dict_tf(gpu)
    assert "Deep-part/bn_0/moving_mean" in st and "Deep-part/bn_1/gamma/Adam" in st
    assert "Deep-part/bn_0/moving_mean/Adam" not in st        # non-trainable: no slots
    _, gpu2 = _deepfm_pair(128, 5000)
    tf_names.load_state_dict_tf(gpu2, st)
    assert torch.equal(gpu.predict(ids, vals), gpu2.predict(ids, vals))


@pytest.mark.parametrize("model", ["DCN", "NFM", "PNN"])
def test_other_models_accept_batch_norm(model):
    from oracle import models as om
    from tf_repos_b200 import synth
    from tf_repos_b200.dcn import DCN
    from tf_repos_b200.nfm import NFM
    from tf_repos_b200.pnn import PNN
    B, N, K = 128, 5000, 8
    kw = dict(deep_layers="32,16", dropout="1.0,1.0", l2_reg=1e-4, learning_rate=1e-3, optimizer="Adam")
    if model == "DCN":
        ref = om.DCN(39, N, K, cross_layers=2, update_mode="exact", seed=2, batch_norm=True, **kw)
        gpu = DCN(39, N, K, B, cross_layers=2, update_mode="exact", device="cuda:0", batch_norm=True, **kw)
    elif model == "NFM":
        kw["dropout"] = "1.0,1.0,1.0"
        ref = om.NFM(39, N, K, update_mode="exact", seed=2, batch_norm=True, **kw)
        gpu = NFM(39, N, K, B, update_mode="exact", device="cuda:0", batch_norm=True, **kw)
    else:
        ref = om.PNN(39, N, K, model_type="Inner", update_mode="exact", seed=2, batch_norm=True, **kw)
        gpu = PNN(39, N, K, B, model_type="Inner", update_mode="exact", device="cuda:0", batch_norm=True, **kw)
    g = torch.Generator().manual_seed(9)
    ref.params[ref.tables[-1]].copy_(torch.randn(N, K, generator=g) * 0.1)
    gpu.load_variables(ref.params)
    for step in range(2):
        ids, vals, labels = synth.criteo_batch(B, N, 39, seed=60 + step)
        loss_ref = ref.train_step({"feat_ids": ids.long(), "feat_vals": vals}, labels)
        parts = gpu.train_step(ids.cuda(), vals.cuda(), labels.cuda())
        assert abs(gpu.loss_value(parts) - loss_ref) <= 3e-5 * abs(loss_ref), (model, step, gpu.loss_value(parts), loss_ref)
    vs = gpu.variables()
    for name, v in ref.bn_state.items():
        _close(vs[name], v, 2e-5, name)
