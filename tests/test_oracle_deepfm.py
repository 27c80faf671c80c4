"""Pins the CPU oracle itself (the reference ships no tests or golden vectors -- parity is
unpinned against TensorFlow, see oracle/tf_semantics.py): closed-form known answers, fp64
autograd cross-checks of the hand-derived gradients the CUDA kernels implement, TF optimizer
arithmetic on hand-computed cases, and the committed golden fixtures."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import models as om
from oracle import tf_semantics as tfs

HERE = os.path.dirname(os.path.abspath(__file__))

# the one full 39-field sample the reference contains (Serving_pipeline/deep_fm_serving_client.cpp:42-45
# gives ids {1..13, 15, 555, 1078, 17797, ..., 111823} for feature_size=117581); values: 13 scaled
# continuous features then 26 ones.
SERVING_IDS = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 555, 1078, 17797, 26190, 26341, 28570, 35361,
               35613, 35984, 48424, 51364, 64053, 65964, 66206, 71628, 84088, 84119, 86889, 88280, 88283,
               100288, 100300, 102447, 109932, 111823]
SERVING_VALS = [0.05, 0.006633, 0.05, 0, 0.021594, 0.008, 0.15, 0.04, 0.362, 0.1, 0.2, 0, 0.04] + [1.0] * 26


def _small(opt="Adam", mode="exact", dtype=torch.float32, F=5, N=50, K=4, **kw):
    m = om.DeepFM(F, N, K, deep_layers="8,4", dropout="1.0,1.0", optimizer=opt, update_mode=mode,
                  dtype=dtype, seed=1, **kw)
    g = torch.Generator().manual_seed(2)
    m.params["fm_v"].copy_(torch.randn(N, K, generator=g, dtype=torch.float64) * 0.3)
    m.params["fm_w"].copy_(torch.randn(N, generator=g, dtype=torch.float64) * 0.3)
    m.params["fm_bias"].fill_(0.25)
    for n_, p in m.params.items():
        if n_.endswith("biases"):
            p.copy_(torch.randn(p.shape, generator=g, dtype=torch.float64) * 0.1)
    return m


def _batch(B, F, N, seed=0):
    g = torch.Generator().manual_seed(seed)
    return ({"feat_ids": torch.randint(0, N, (B, F), generator=g), "feat_vals": torch.rand(B, F, generator=g)},
            (torch.rand(B, generator=g) < 0.4).float())


# ---- closed-form known answers ---------------------------------------------------
def test_all_zero_values_gives_bias_plus_mlp_of_zero():
    m = _small()
    batch, _ = _batch(6, 5, 50)
    batch["feat_vals"] = torch.zeros(6, 5)
    out = m.predict(batch)
    h = torch.zeros(1, 5 * 4)
    for i in range(2):
        h = torch.relu(h @ m.params[f"Deep-part/mlp{i}/weights"] + m.params[f"Deep-part/mlp{i}/biases"])
    y_d = (h @ m.params["Deep-part/deep_out/weights"] + m.params["Deep-part/deep_out/biases"]).item()
    assert torch.all(out["y_w"] == 0) and torch.all(out["y_v"] == 0)
    np.testing.assert_allclose(out["y"].numpy(), np.full(6, 0.25 + y_d, dtype=np.float32), rtol=1e-6)


def test_single_active_field_has_zero_second_order():
    m = _small()
    batch, _ = _batch(6, 5, 50)
    v = torch.zeros(6, 5); v[:, 2] = batch["feat_vals"][:, 2]
    batch["feat_vals"] = v
    assert torch.all(m.predict(batch)["y_v"] == 0)


def test_two_fields_second_order_is_inner_product():
    m = om.DeepFM(2, 30, 8, deep_layers="4", dropout="1.0", seed=0)
    m.params["fm_v"].copy_(torch.randn(30, 8) * 0.5)
    batch, _ = _batch(9, 2, 30)
    out = m.predict(batch)
    e = m.params["fm_v"][batch["feat_ids"]] * batch["feat_vals"][..., None]
    np.testing.assert_allclose(out["y_v"].numpy(), (e[:, 0] * e[:, 1]).sum(1).numpy(), rtol=2e-5, atol=1e-6)


def test_loss_formula_and_ignored_fc_regulariser():
    """loss = mean CE + l2*l2_loss(fm_w) + l2*l2_loss(fm_v) and nothing for the MLP weights (quirk Q4)."""
    m = _small(l2_reg=0.01)
    batch, labels = _batch(7, 5, 50)
    loss, out, _, dg = m.gradients(batch, labels)
    y = out["y"].double()
    ce = (torch.clamp(y, min=0) - y * labels.double() + torch.log1p(torch.exp(-y.abs()))).mean()
    reg = 0.01 * 0.5 * (m.params["fm_w"].double() ** 2).sum() + 0.01 * 0.5 * (m.params["fm_v"].double() ** 2).sum()
    assert abs(float(loss) - float(ce + reg)) < 1e-6
    # dense (MLP) gradients carry no L2 term: zero weights' grad does not depend on l2_reg
    m2 = _small(l2_reg=0.5)
    _, _, _, dg2 = m2.gradients(batch, labels)
    assert torch.equal(dg["Deep-part/mlp0/weights"], dg2["Deep-part/mlp0/weights"])


# ---- hand-derived gradients (what K2 implements) vs fp64 autograd ----------------------------------
def test_fm_backward_formula_matches_autograd_fp64():
    m = _small(dtype=torch.float64)
    B, F, K = 11, 5, 4
    batch, labels = _batch(B, F, 50, seed=3)
    _, out, _, _ = m.gradients(batch, labels)
    per = out["per_occurrence"]
    y = out["y"]
    dy = (tfs.sigmoid(y) - labels.double()) / B
    vals = batch["feat_vals"].double()
    e = m.params["fm_v"][batch["feat_ids"]] * vals[..., None]
    S = e.sum(1)
    # dX: backprop through the MLP by hand
    x = e.reshape(B, F * K)
    W0, b0 = m.params["Deep-part/mlp0/weights"], m.params["Deep-part/mlp0/biases"]
    W1, b1 = m.params["Deep-part/mlp1/weights"], m.params["Deep-part/mlp1/biases"]
    Wo = m.params["Deep-part/deep_out/weights"]
    h0 = torch.relu(x @ W0 + b0); h1 = torch.relu(h0 @ W1 + b1)
    d1 = (dy[:, None] * Wo.view(1, -1)) * (h1 > 0)
    d0 = (d1 @ W1.t()) * (h0 > 0)
    dX = (d0 @ W0.t()).reshape(B, F, K)
    g_e = dy[:, None, None] * (S[:, None, :] - e) + dX          # K2's formula
    np.testing.assert_allclose(per["v"].numpy(), (g_e * vals[..., None]).numpy(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(per["w"].numpy(), (dy[:, None] * vals).numpy(), rtol=1e-10, atol=1e-14)


def test_fp32_oracle_tracks_fp64_oracle():
    a, b = _small(), _small(dtype=torch.float64)
    for step in range(3):
        batch, labels = _batch(16, 5, 50, seed=step)
        la = a.train_step(batch, labels); lb = b.train_step(batch, labels)
        assert abs(la - lb) < 1e-5 * abs(lb)
    for n_ in a.params:
        np.testing.assert_allclose(a.params[n_].numpy(), b.params[n_].numpy(), rtol=0, atol=2e-5 * float(b.params[n_].abs().max()))


# ---- TF optimizer semantics -------------------------------------------------------------------------
def test_adam_first_step_is_lr_times_sign_for_large_gradients():
    var = torch.tensor([1.0, -2.0, 3.0]); m = torch.zeros(3); v = torch.zeros(3)
    g = torch.tensor([0.5, -0.25, 4.0])
    a = tfs.AdamHyper(0.01)
    tfs.adam_dense_(var, m, v, g, a.lr_t(), a.b1, a.b2, a.eps)
    np.testing.assert_allclose(var.numpy(), np.array([1.0 - 0.01, -2.0 + 0.01, 3.0 - 0.01]), rtol=1e-6)
    var2 = torch.tensor([1.0, -2.0, 3.0]); m2 = torch.zeros(3); v2 = torch.zeros(3)
    tfs.adam_sparse_(var2, m2, v2, g, a.lr_t(), a.b1, a.b2, a.eps)
    np.testing.assert_allclose(var2.numpy(), var.numpy(), rtol=1e-6)
    np.testing.assert_allclose(m2.numpy(), 0.1 * g.numpy(), rtol=1e-6)
    # fp32(1) - fp32(0.999) = 0.00099998713 (TF computes 1-beta2 in fp32 too)
    np.testing.assert_allclose(v2.numpy(), 0.001 * g.numpy() ** 2, rtol=2e-5)


def test_adam_beta_powers_are_running_fp32_products():
    a = tfs.AdamHyper(5e-4)
    b1 = np.float32(0.9); b2 = np.float32(0.999); p1, p2 = b1, b2
    for t in range(1, 30):
        expect = np.float32(5e-4) * np.sqrt(np.float32(1) - p2) / (np.float32(1) - p1)
        assert a.lr_t().item() == expect
        a.finish(); p1 = p1 * b1; p2 = p2 * b2


def test_adagrad_momentum_ftrl_hand_values():
    var = torch.tensor([1.0]); acc = torch.tensor([1e-8]); g = torch.tensor([0.5])
    tfs.adagrad_(var, acc, g, torch.tensor(0.1))
    assert abs(acc.item() - (1e-8 + 0.25)) < 1e-7 and abs(var.item() - (1 - 0.1 * 0.5 / math.sqrt(0.25 + 1e-8))) < 1e-6
    var = torch.tensor([1.0]); acc = torch.tensor([2.0])
    tfs.momentum_(var, acc, g, torch.tensor(0.1), torch.tensor(0.95))
    assert abs(acc.item() - 2.4) < 1e-6 and abs(var.item() - (1 - 0.24)) < 1e-6
    var = torch.tensor([1.0]); accum = torch.tensor([0.1]); lin = torch.tensor([0.0])
    tfs.ftrl_(var, accum, lin, g, 0.1)
    na = 0.1 + 0.25
    l_ = 0.5 - (math.sqrt(na) - math.sqrt(0.1)) / 0.1 * 1.0
    assert abs(lin.item() - l_) < 1e-5 and abs(var.item() - (-l_ / (math.sqrt(na) / 0.1))) < 1e-5
    assert abs(accum.item() - na) < 1e-6


def test_exact_mode_updates_every_row_and_lazy_only_gathered():
    for mode in ("exact", "lazy"):
        m = _small(mode=mode)
        before = m.params["fm_v"].clone()
        batch, labels = _batch(4, 5, 50)
        m.train_step(batch, labels)
        moved = (m.params["fm_v"] != before).any(1)
        touched = torch.zeros(50, dtype=torch.bool); touched[batch["feat_ids"].reshape(-1)] = True
        if mode == "exact":
            assert moved.all()
        else:
            assert torch.equal(moved, touched)


def test_adam_without_l2_still_decays_untouched_rows():
    """Non-lazy sparse Adam: a row gathered at step 1 keeps moving at step 2."""
    m = _small(l2_reg=0.0)
    b1, l1 = _batch(4, 5, 50, seed=0)
    m.train_step(b1, l1)
    snap = m.params["fm_v"].clone()
    ids2 = torch.full((4, 5), 49); ids2[:, 0] = 48
    m.train_step({"feat_ids": ids2, "feat_vals": torch.ones(4, 5)}, l1)
    first = torch.zeros(50, dtype=torch.bool); first[b1["feat_ids"].reshape(-1)] = True
    first[48] = first[49] = False
    assert (m.params["fm_v"][first] != snap[first]).any(1).all()


def test_dedup_matches_dense_scatter_and_is_ascending():
    rng = np.random.default_rng(0)
    idx = rng.integers(0, 20, size=200)
    vals = rng.standard_normal((200, 3)).astype(np.float32)
    summed, uniq = tfs.deduplicate_indexed_slices(vals, idx)
    assert np.all(np.diff(uniq) > 0)
    dense = np.zeros((20, 3)); np.add.at(dense, idx, vals.astype(np.float64))
    np.testing.assert_allclose(summed, dense[uniq], rtol=1e-5, atol=1e-6)
    perm, u2, inv, seg = tfs.unique_segment_reference(idx.astype(np.int32))
    assert np.array_equal(u2, uniq) and np.array_equal(u2[inv], idx) and seg[-1] == 200
    assert np.array_equal(np.sort(perm), np.arange(200)) and np.all(np.diff(idx[perm]) >= 0)


def test_auc_matches_exact_auc_closely():
    rng = np.random.default_rng(1)
    y = rng.random(5000) < 0.3
    p = np.clip(rng.normal(0.3 + 0.2 * y, 0.15), 0, 1)
    from sklearn.metrics import roc_auc_score
    assert abs(tfs.auc(y, p) - roc_auc_score(y, p)) < 2e-3


def test_out_of_range_id_raises_like_tf():
    m = _small()
    with pytest.raises(IndexError):
        m.predict({"feat_ids": torch.full((1, 5), 50), "feat_vals": torch.ones(1, 5)})


def test_unknown_optimizer_is_a_name_error_like_the_reference():
    with pytest.raises(NameError):
        _small(opt="GD")


# ---- golden fixtures ---------------------------------------------------------------------------------
def test_golden_serving_sample():
    """tests/golden/deepfm_serving_sample.json was generated by tests/golden/make_golden.py from this
    oracle (NOT from TensorFlow): it guards the restatement against drift."""
    path = os.path.join(HERE, "golden", "deepfm_serving_sample.json")
    gold = json.load(open(path))
    from tests.golden.make_golden import build_serving_case
    m, batch, labels = build_serving_case()
    out = m.predict(batch)
    np.testing.assert_allclose(out["y"].numpy(), np.array(gold["logits"], dtype=np.float32), rtol=1e-6, atol=1e-7)
    losses = [m.train_step(batch, labels) for _ in range(3)]
    np.testing.assert_allclose(losses, gold["losses"], rtol=1e-6)
    rows = m.params["fm_v"][torch.tensor(gold["row_ids"])].numpy()
    np.testing.assert_allclose(rows, np.array(gold["fm_v_rows_after_3_steps"], dtype=np.float32), rtol=1e-5, atol=1e-8)
    assert gold["serving_ids"] == SERVING_IDS
