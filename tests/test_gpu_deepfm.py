"""End-to-end DeepFM parity: N training steps of the CUDA engine vs the CPU oracle on the same seeded
Criteo-layout batches (config 1 of BASELINE.json: 39 fields, 10k vocab, k=8), logits and updated
embedding rows within 1e-5 relative; exact (TensorFlow) and lazy update modes; all four optimizers."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pair(opt="Adam", mode="exact", B=256, N=10_000, K=8, l2=1e-4, lr=5e-4, layers="256,128,64"):
    from oracle import models as om
    from tf_repos_b200.deepfm import DeepFM
    F = 39
    ref = om.DeepFM(F, N, K, deep_layers=layers, dropout="1.0,1.0,1.0", l2_reg=l2, learning_rate=lr,
                    optimizer=opt, update_mode=mode, seed=11)
    # make the tables O(0.1) so that gradients and L2 terms are well above fp32 noise
    g = torch.Generator().manual_seed(5)
    ref.params["fm_v"].copy_(torch.randn(N, K, generator=g) * 0.1)
    ref.params["fm_w"].copy_(torch.randn(N, generator=g) * 0.1)
    gpu = DeepFM(F, N, K, B, deep_layers=layers, dropout="1.0,1.0,1.0", l2_reg=l2, learning_rate=lr,
                 optimizer=opt, update_mode=mode, device="cuda:0")
    gpu.load_variables(ref.params)
    return ref, gpu


def _assert_close(got, ref, rtol, what):
    got = got.detach().cpu().double().numpy(); ref = ref.detach().cpu().double().numpy()
    scale = max(float(np.abs(ref).max()), 1e-30)
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol * scale, err_msg=what)


@pytest.mark.parametrize("opt,mode", [("Adam", "exact"), ("Adam", "lazy"), ("Adagrad", "exact"),
                                      ("Momentum", "exact"), ("ftrl", "exact"), ("Adagrad", "lazy")])
def test_deepfm_train_steps_match_oracle(opt, mode):
    from tf_repos_b200 import synth
    B, N = 256, 10_000
    lr = 5e-4 if opt == "Adam" else 0.01
    ref, gpu = _pair(opt, mode, B=B, N=N, lr=lr)
    for step in range(4):
        ids, vals, labels = synth.criteo_batch(B, N, 39, seed=100 + step)
        batch = {"feat_ids": ids.long(), "feat_vals": vals}
        # forward parity before the update
        prob = gpu.predict(ids.cuda(), vals.cuda())
        out = ref.predict(batch)
        _assert_close(gpu.y[:B], out["y"], 1e-5, f"logits step {step}")
        _assert_close(prob, out["prob"], 1e-5, f"prob step {step}")
        loss_ref = ref.train_step(batch, labels)
        parts = gpu.train_step(ids.cuda(), vals.cuda(), labels.cuda())
        gpu.check_ids()
        if mode == "exact":
            assert abs(gpu.loss_value(parts) - loss_ref) <= 1e-5 * abs(loss_ref), (gpu.loss_value(parts), loss_ref)
        vs = gpu.variables()
        for name in ("fm_v", "fm_w", "fm_bias", "Deep-part/mlp0/weights", "Deep-part/mlp2/biases",
                     "Deep-part/deep_out/weights"):
            # a few fp32 ulps of the gradient are amplified by Adam's normalisation on near-zero
            # gradients; 1e-5 relative to the variable's scale is the north-star tolerance
            _assert_close(vs[name], ref.params[name], 2e-5 if step else 1e-5, f"{name} after step {step} ({opt},{mode})")


def test_deepfm_exact_moves_every_row_lazy_does_not():
    from tf_repos_b200 import synth
    B, N = 128, 5000
    for mode in ("exact", "lazy"):
        ref, gpu = _pair("Adam", mode, B=B, N=N)
        before = gpu.fm_v.var.clone()
        ids, vals, labels = synth.criteo_batch(B, N, 39, seed=1)
        gpu.train_step(ids.cuda(), vals.cuda(), labels.cuda())
        moved = (gpu.fm_v.var != before).any(dim=1)
        touched = torch.zeros(N, dtype=torch.bool, device="cuda"); touched[ids.long().cuda().view(-1)] = True
        if mode == "exact":
            assert moved.all(), "TensorFlow's Adam + dense L2 gradient updates every row"
        else:
            assert torch.equal(moved, touched & moved) and moved.sum() > 0


def test_deepfm_step_is_deterministic():
    from tf_repos_b200 import synth
    B, N = 256, 10_000
    outs = []
    for _ in range(2):
        _, gpu = _pair("Adam", "exact", B=B, N=N)
        for step in range(2):
            ids, vals, labels = synth.criteo_batch(B, N, 39, seed=step)
            gpu.train_step(ids.cuda(), vals.cuda(), labels.cuda())
        outs.append((gpu.fm_v.var.clone(), gpu.fm_w.var.clone(), gpu.dense.flat.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_deepfm_full_batch_properties_large_vocab():
    """Config-2 shaped (B=8192, F=39, K=16) on a 20M-row table: size-independent properties."""
    from tf_repos_b200 import ops, synth
    from tf_repos_b200.deepfm import DeepFM
    B, N, K = 8192, 20_000_000, 16
    m = DeepFM(39, N, K, B, l2_reg=1e-4, update_mode="exact", device="cuda:0", dropout="1.0,1.0,1.0")
    ids, vals, labels = synth.criteo_batch(B, N, 39, seed=7, device="cuda")
    v0 = m.fm_v.var.clone()
    m.train_step(ids, vals, labels)
    m.check_ids()
    torch.cuda.synchronize()
    # (1) K3 invariants at full size: perm is a permutation, sorted ids ascending, inverse round-trips
    uw = m.updater.uw
    n = B * 39
    U = uw.n_uniq.item()
    flat = ids.view(-1)
    perm = uw.perm[:n].long()
    assert torch.equal(torch.sort(perm)[0], torch.arange(n, device="cuda"))
    s = flat[perm]
    assert torch.all(s[1:] >= s[:-1])
    assert torch.equal(uw.uniq[:U][uw.inverse[:n].long()], flat)
    assert torch.all(uw.uniq[1:U] > uw.uniq[: U - 1])
    assert uw.seg_offsets[U].item() == n
    # (2) linearity of the scatter-add: sum over unique rows == sum over occurrences
    tot_u = m.updater.g_uniq[: U * K].view(U, K).double().sum(0)
    tot_o = m.g_rows.double().sum(0)
    assert torch.allclose(tot_u, tot_o, rtol=1e-6, atol=1e-9)
    # (3) exact mode: every untouched row took the closed-form first Adam step with g = l2*var
    untouched = torch.ones(N, dtype=torch.bool, device="cuda"); untouched[flat.long()] = False
    w0 = v0[untouched]
    f = lambda x: torch.tensor(x, dtype=torch.float32, device="cuda")
    g = f(1e-4) * w0
    m1 = g * (f(1.0) - f(0.9)); v1 = (g * g) * (f(1.0) - f(0.999))
    lr_t = f(5e-4) * torch.sqrt(f(1.0) - f(0.999)) / (f(1.0) - f(0.9))
    expect = w0 - (lr_t * m1) / (torch.sqrt(v1) + f(1e-8))
    assert torch.equal(m.fm_v.var[untouched], expect), "dense sweep must be bit-exact with the fp32 formula"
    assert torch.equal(m.fm_v.slots[0][untouched], m1) and torch.equal(m.fm_v.slots[1][untouched], v1)
