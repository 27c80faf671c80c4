"""exact_deferred mode (csrc/epoch.cu) must reproduce the exact mode (sweep every row every step) BIT FOR
BIT: the untouched-row update is replayed lazily, in the same fp32 operation order."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _models(opt, l2, P, B=128, N=3000, K=8, lr=None):
    from tf_repos_b200.deepfm import DeepFM
    lr = lr if lr is not None else (5e-4 if opt == "Adam" else 0.01)
    kw = dict(deep_layers="32,16", dropout="1.0,1.0", l2_reg=l2, learning_rate=lr, optimizer=opt, device="cuda:0")
    a = DeepFM(39, N, K, B, update_mode="exact", **kw)
    b = DeepFM(39, N, K, B, update_mode="exact_deferred", epoch_steps=P, **kw)
    g = torch.Generator().manual_seed(3)
    vals = {"fm_v": torch.randn(N, K, generator=g) * 0.1, "fm_w": torch.randn(N, generator=g) * 0.1}
    a.load_variables(vals); b.load_variables(vals)
    b.dense.flat.copy_(a.dense.flat)
    return a, b


def _same(a, b, what):
    b.flush()
    for ta, tb in ((a.fm_v, b.fm_v), (a.fm_w, b.fm_w)):
        assert torch.equal(ta.var, tb.var), f"{what}: {ta.name} var"
        for sa, sb in zip(ta.slots, tb.slots):
            assert torch.equal(sa, sb), f"{what}: {ta.name} slot"
    assert torch.equal(a.dense.flat, b.dense.flat), f"{what}: dense"


@pytest.mark.parametrize("opt,l2", [("Adam", 1e-4), ("Adam", 0.0), ("Adagrad", 1e-3), ("Momentum", 1e-3), ("ftrl", 1e-3)])
@pytest.mark.parametrize("P", [1, 3, 8])
def test_deferred_state_is_bit_identical(opt, l2, P):
    from tf_repos_b200 import synth
    a, b = _models(opt, l2, P)
    B, N = a.B, a.N
    for step in range(2 * P + 3):
        ids, vals, labels = synth.criteo_batch(B, N, 39, seed=step, device="cuda")
        if step % 3 == 0:  # some batches re-gather rows of the previous batch (catch-up by 1 step)
            ids[:, 20:] = ids_prev[:, 20:] if step else ids[:, 20:]
        ids_prev = ids
        la = a.train_step(ids, vals, labels)
        lb = b.train_step(ids, vals, labels)
        assert torch.equal(la[0], lb[0]), f"CE differs at step {step}"
        if step in (0, P, 2 * P + 2):
            _same(a, b, f"{opt} l2={l2} P={P} after step {step}")
    _same(a, b, "final")


def test_deferred_loss_terms_and_predict_mid_epoch():
    from tf_repos_b200 import synth
    P = 4
    a, b = _models("Adam", 1e-3, P)
    B, N = a.B, a.N
    regs = []
    for step in range(P):
        ids, vals, labels = synth.criteo_batch(B, N, 39, seed=10 + step, device="cuda")
        if step == 2:  # predict in the middle of an epoch must see up-to-date rows
            pa = a.predict(ids, vals).clone(); pb = b.predict(ids, vals).clone()
            assert torch.equal(pa, pb)
        la = a.train_step(ids, vals, labels)
        b.train_step(ids, vals, labels)
        regs.append(la[1:].clone())  # exact mode: [l2*l2_loss(fm_w), l2*l2_loss(fm_v)] of this step
    got = b.epoch_reg_terms()        # [2, P] for the epoch that just closed
    want = torch.stack(regs, dim=1)
    assert torch.allclose(got, want, rtol=1e-5, atol=0), (got, want)
    _same(a, b, "after one epoch")


def test_deferred_large_rows_k16_and_k256():
    """vector row kernels (K=16: 4 lanes/row; K=256: 32 lanes x 2)"""
    from tf_repos_b200 import synth
    for K in (16, 256):
        a, b = _models("Adam", 1e-4, 4, B=64, N=2000, K=K)
        for step in range(6):
            ids, vals, labels = synth.criteo_batch(64, 2000, 39, seed=step, device="cuda")
            a.train_step(ids, vals, labels); b.train_step(ids, vals, labels)
        _same(a, b, f"K={K}")


def test_deferred_bit_identical_on_zero_denormal_and_tiny_state():
    """Rows nothing gathers are pulled to 0 by l2 + Adam and their m underflows: the sweep's grouped fast path has a
    zero-group shortcut, a guarded range down to 2^-100 and the compiler's slow path below it.  All three must leave
    the same bits as the every-step sweep."""
    from tf_repos_b200 import synth
    P = 3
    a, b = _models("Adam", 1e-4, P, N=4096)
    N, K = a.N, a.K
    g = torch.Generator().manual_seed(11)
    scales = torch.tensor([0.0, 1e-42, 1e-39, 1e-36, 1e-33, 1e-30, 1e-26, 1e-20, 1e-12, 1e-6, 1e-2, 1.0])
    v = torch.randn(N, K, generator=g) * scales[torch.randint(0, len(scales), (N, 1), generator=g)]
    v[::7] = torch.randn(N, K, generator=g)[::7] * scales[torch.randint(0, len(scales), (N, K), generator=g)][::7]  # mixed groups
    w = torch.randn(N, generator=g) * scales[torch.randint(0, len(scales), (N,), generator=g)]
    for m in (a, b):
        m.load_variables({"fm_v": v, "fm_w": w})
        # slots in every regime too: zero, denormal and tiny first moments; second moments from 0 upwards
        m.fm_v.slots[0].copy_((torch.randn(N, K, generator=torch.Generator().manual_seed(5)) * 1e-38).cuda() * (torch.arange(N).cuda() % 3 == 0).float().view(N, 1))
        m.fm_v.slots[1].copy_((torch.rand(N, K, generator=torch.Generator().manual_seed(6)) * 1e-30).cuda() * (torch.arange(N).cuda() % 2 == 0).float().view(N, 1))
    for step in range(2 * P + 2):
        ids, vals, labels = synth.criteo_batch(a.B, N, 39, seed=70 + step, device="cuda")
        a.train_step(ids, vals, labels); b.train_step(ids, vals, labels)
        _same(a, b, f"extreme state after step {step}")
