"""The oracle against an independent implementation of the same published algorithms: a textbook DeepFM written on
torch.nn.functional (embedding, linear, binary_cross_entropy_with_logits) with the L2 terms INSIDE the autograd loss, trained
by torch.optim's Adagrad / SGD-momentum / Adam.  Nothing under oracle/ is used on that side.

What this pins (and what it does not): the reference's loss is `mean(sigmoid_ce) + l2*l2_loss(fm_w) + l2*l2_loss(fm_v)`
(DeepFM.py:188-190) minimised by a stock optimizer (DeepFM.py:204-211); differentiating the L2 terms over the whole table
is what makes EVERY row move every step.  The oracle restates that through TF's IndexedSlices mechanics (dense gradient
concatenated with the de-duplicated gather gradients, non-lazy sparse Adam); here the same trajectory must come out of
plain dense autograd + a third-party optimizer.  It is NOT TensorFlow: rounding order, beta-power bookkeeping and Adam's
epsilon placement (TF: sqrt(v)+eps under lr_t; torch: sqrt(v_hat)+eps) are outside this check, so it runs in fp64 with
tolerances 1e-12 (Adagrad, Momentum), 1e-9 (Adam with epsilon 1e-14 on both sides, where the placement cannot matter) and
1e-3 (Adam with the reference's epsilon 1e-8: measured 2.8e-4 on the MLP weights after 6 steps, exactly 10^6 times the
epsilon-1e-14 difference, i.e. all of it is the placement -- at step 1 TF's epsilon acts 1/sqrt(1-beta2) = 31.6 times
larger than torch's) of each variable's scale."""
import pytest
import torch
import torch.nn.functional as Fnn

from oracle import models as om

F64 = torch.float64
B, F, N, K = 64, 6, 40, 4
LAYERS = [16, 8]
L2, LR = 1e-2, 5e-3


def _batches(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        ids = torch.randint(0, N - 7, (B, F), generator=g)          # rows N-7.. are never gathered: they must still move
        vals = torch.rand(B, F, generator=g, dtype=F64) + 0.1
        labels = (torch.rand(B, generator=g) < 0.3).to(F64)
        out.append((ids, vals, labels))
    return out


def _textbook_loss(p, ids, vals, labels):
    w = Fnn.embedding(ids, p["fm_w"].unsqueeze(1)).squeeze(-1)                     # [B, F]
    e = Fnn.embedding(ids, p["fm_v"]) * vals.unsqueeze(-1)                          # [B, F, K]
    first = (w * vals).sum(1)
    second = 0.5 * (e.sum(1).pow(2) - e.pow(2).sum(1)).sum(1)
    h = e.reshape(B, F * K)
    for i in range(len(LAYERS)):
        h = Fnn.relu(Fnn.linear(h, p[f"Deep-part/mlp{i}/weights"].t(), p[f"Deep-part/mlp{i}/biases"]))
    deep = Fnn.linear(h, p["Deep-part/deep_out/weights"].t(), p["Deep-part/deep_out/biases"]).squeeze(1)
    logit = p["fm_bias"] + first + second + deep
    return (Fnn.binary_cross_entropy_with_logits(logit, labels)
            + L2 * 0.5 * p["fm_w"].pow(2).sum() + L2 * 0.5 * p["fm_v"].pow(2).sum())


def _torch_optimizer(name, params, eps=1e-8):
    if name == "Adagrad":       # DeepFM.py:207: AdagradOptimizer(lr, initial_accumulator_value=1e-8); TF's Adagrad has no epsilon
        return torch.optim.Adagrad(params, lr=LR, initial_accumulator_value=1e-8, eps=0.0)
    if name == "Momentum":      # DeepFM.py:209: MomentumOptimizer(lr, momentum=0.95)
        return torch.optim.SGD(params, lr=LR, momentum=0.95)
    return torch.optim.Adam(params, lr=LR, betas=(0.9, 0.999), eps=eps)            # DeepFM.py:205


def _run_pair(opt, eps):
    ref = om.DeepFM(F, N, K, deep_layers=LAYERS, dropout="1.0,1.0", l2_reg=L2, learning_rate=LR, optimizer=opt,
                    dtype=F64, seed=5)
    if eps is not None:
        ref.adam.eps = torch.tensor(eps, dtype=F64)
    g = torch.Generator().manual_seed(1)
    ref.params["fm_w"].copy_(torch.randn(N, generator=g, dtype=F64) * 0.3)
    ref.params["fm_v"].copy_(torch.randn(N, K, generator=g, dtype=F64) * 0.3)
    book = {n: p.clone().requires_grad_() for n, p in ref.params.items()}
    start = {n: p.clone() for n, p in ref.params.items()}
    optim = _torch_optimizer(opt, list(book.values()), eps)
    losses = []
    for ids, vals, labels in _batches(6):
        loss_ref = ref.train_step({"feat_ids": ids, "feat_vals": vals}, labels)
        optim.zero_grad()
        loss = _textbook_loss(book, ids, vals, labels)
        loss.backward()
        optim.step()
        losses.append((loss.item(), loss_ref))                                     # the loss BEFORE this step's update
    diff = {n: (book[n].detach() - p).abs().max().item() / (p.abs().max().item() or 1.0) for n, p in ref.params.items()}
    return ref, start, losses, diff


@pytest.mark.parametrize("opt,tol,eps", [("Adagrad", 1e-12, None), ("Momentum", 1e-12, None), ("Adam", 1e-3, 1e-8),
                                         ("Adam", 1e-9, 1e-14)])
def test_oracle_trajectory_equals_textbook_autograd_plus_torch_optim(opt, tol, eps):
    ref, start, losses, diff = _run_pair(opt, eps)
    for loss, loss_ref in losses:
        assert abs(loss - loss_ref) <= max(tol, 1e-9) * abs(loss_ref)
    assert max(diff.values()) <= tol, diff
    # every table row moved, gathered or not (the never-gathered tail included)
    assert bool(((ref.params["fm_v"] - start["fm_v"]).abs() > 0).all())
    assert bool(((ref.params["fm_w"] - start["fm_w"]).abs() > 0).all())


def test_adam_difference_is_all_epsilon_placement():
    """The residual against torch.optim.Adam is linear in epsilon (so it vanishes with it): nothing else differs."""
    d8, d14 = _run_pair("Adam", 1e-8)[3], _run_pair("Adam", 1e-14)[3]
    for n in d8:
        assert d14[n] > 0 and abs(d8[n] / d14[n] / 1e6 - 1.0) <= 0.05, (n, d8[n], d14[n])


def test_lazy_mode_is_not_the_reference_semantics():
    """The same textbook run must DISAGREE with the oracle's `lazy` mode on never-gathered rows (guards the test above
    against passing vacuously)."""
    ref = om.DeepFM(F, N, K, deep_layers=LAYERS, dropout="1.0,1.0", l2_reg=L2, learning_rate=LR, optimizer="Adam",
                    dtype=F64, seed=5, update_mode="lazy")
    before = ref.params["fm_v"].clone()
    for ids, vals, labels in _batches(2):
        ref.train_step({"feat_ids": ids, "feat_vals": vals}, labels)
    assert torch.equal(ref.params["fm_v"][N - 7:], before[N - 7:])
