"""The exact-deferred update at the configuration bench.py quotes its headline on -- Adam, l2 1e-4, K=16, epochs of 16
steps, batch 8192, deep_layers 256,128,64, dropout 0.5 -- must leave the same bits as sweeping every row every step
(`exact`), at a vocabulary large enough (2e7 rows) that the sweep's grid-stride loop runs many iterations and both
register slots of every thread are live.  Also: the state rows nothing gathers park in after a long run (var ~FLT_MIN,
denormal m: the packed sweep's scaled loops), rows that span warps / CTAs (K=12, K=256), and the packed Adam loops
against the scalar step (ctr_selftest_adam_packed).  Reference semantics: DeepFM.py:188-213 as restated in
oracle/tf_semantics.py; `exact` itself is checked against the oracle in test_gpu_deepfm.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pair(N, K, P, B, layers="256,128,64", dropout="0.5,0.5,0.5", l2=1e-4, lr=5e-4):
    from tf_repos_b200.deepfm import DeepFM
    kw = dict(deep_layers=layers, dropout=dropout, l2_reg=l2, learning_rate=lr, optimizer="Adam", device="cuda:0")
    a = DeepFM(39, N, K, B, update_mode="exact", **kw)
    b = DeepFM(39, N, K, B, update_mode="exact_deferred", epoch_steps=P, **kw)
    b.fm_v.var.copy_(a.fm_v.var); b.fm_w.var.copy_(a.fm_w.var)
    b.dense.flat.copy_(a.dense.flat)
    return a, b


def _same(a, b, what):
    b.flush()
    for ta, tb in ((a.fm_v, b.fm_v), (a.fm_w, b.fm_w)):
        assert torch.equal(ta.var, tb.var), f"{what}: {ta.name} var"
        for i, (sa, sb) in enumerate(zip(ta.slots, tb.slots)):
            assert torch.equal(sa, sb), f"{what}: {ta.name} slot {i}"
    assert torch.equal(a.dense.flat, b.dense.flat), f"{what}: dense"


def _run(a, b, n_steps, check_at, seed0=0, flush_at=()):
    from tf_repos_b200 import synth
    for step in range(n_steps):
        ids, vals, labels = synth.criteo_batch(a.B, a.N, 39, seed=seed0 + step, device="cuda")
        la = a.train_step(ids, vals, labels)
        lb = b.train_step(ids, vals, labels)
        assert torch.equal(la[0], lb[0]), f"CE differs at step {step}"
        if step in flush_at:
            b.flush()                      # mid-epoch flush; training continues inside the same epoch
        if step in check_at:
            _same(a, b, f"after step {step}")


def test_headline_config_bit_identical_2e7_rows():
    """bench.py's configuration with the vocabulary cut to 2e7 rows (3.2e8 elements = 1400 sweep iterations per
    thread slot): two full epochs, a mid-epoch flush, and a final partial epoch."""
    a, b = _pair(20_000_000, 16, 16, 8192)
    _run(a, b, 16 + 16 + 7, check_at=(15, 20, 31, 38), flush_at=(20, 34))


@pytest.mark.parametrize("mode,P", [("exact_deferred", 4), ("lazy", 1), ("exact", 1)])
def test_graph_replayed_steps_equal_eager_steps(mode, P):
    """train_step_graphed (one CUDA graph per epoch position) must leave exactly the bits of train_step."""
    from tf_repos_b200 import synth
    from tf_repos_b200.deepfm import DeepFM
    kw = dict(deep_layers="64,32", dropout="0.5,0.5", l2_reg=1e-4, learning_rate=5e-4, optimizer="Adam", device="cuda:0")
    a = DeepFM(39, 50_000, 16, 512, update_mode=mode, epoch_steps=P, **kw)
    b = DeepFM(39, 50_000, 16, 512, update_mode=mode, epoch_steps=P, **kw)
    b.fm_v.var.copy_(a.fm_v.var); b.fm_w.var.copy_(a.fm_w.var); b.dense.flat.copy_(a.dense.flat)
    for step in range(3 * max(P, 2) + 1):
        ids, vals, labels = synth.criteo_batch(512, 50_000, 39, seed=step, device="cuda")
        la = a.train_step(ids, vals, labels).clone()
        lb = b.train_step_graphed(ids, vals, labels).clone()
        assert torch.equal(la[0], lb[0]), f"CE differs at step {step}"
    assert len(b._graphs) >= 1
    a.flush(); b.flush()
    for ta, tb in ((a.fm_v, b.fm_v), (a.fm_w, b.fm_w)):
        assert torch.equal(ta.var, tb.var)
        for sa, sb in zip(ta.slots, tb.slots):
            assert torch.equal(sa, sb)
    assert torch.equal(a.dense.flat, b.dense.flat) and a.global_step == b.global_step
    ids, vals, _ = synth.criteo_batch(512, 50_000, 39, seed=99, device="cuda")
    want = a.predict(ids, vals).clone()
    for _ in range(3):                      # eager visit, capture, replay
        assert torch.equal(b.predict_graphed(ids, vals), want)


def _park(model, kind, seed):
    """Overwrite the tables' state with what rows nothing gathers look like after a long run."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    for t in (model.fm_v, model.fm_w):
        shape = t.var.shape
        u = lambda: torch.rand(shape, device="cuda", generator=g)
        sign = lambda: torch.where(u() < 0.5, -1.0, 1.0)
        flt_min = 2.0 ** -126
        t.var.copy_(sign() * (0.25 + 4.0 * u()) * flt_min)          # around FLT_MIN, some denormal
        m = sign() * (u() * 4e-42)                                  # a few thousand denormal ulps
        m = torch.where(u() < 0.2, torch.zeros_like(m), m)          # and exact zeros of both signs
        t.slots[0].copy_(m * sign())
        if kind == "parked":                                        # second moment still in the normal range
            t.slots[1].copy_((0.5 + u()) * 1e-24)
        else:                                                       # very long run: denormal and zero second moments
            v = u() * 1e-40
            t.slots[1].copy_(torch.where(u() < 0.2, torch.zeros_like(v), v))


@pytest.mark.parametrize("kind", ["parked", "very_long_run"])
def test_headline_config_bit_identical_from_parked_state(kind):
    a, b = _pair(20_000_000, 16, 16, 8192)
    _park(a, kind, 5)
    for ta, tb in ((a.fm_v, b.fm_v), (a.fm_w, b.fm_w)):
        tb.var.copy_(ta.var)
        for sa, sb in zip(ta.slots, tb.slots):
            sb.copy_(sa)
    # late in training: lr_t is lr to within a few ulps
    for m in (a, b):
        m.opt.state[0] = 0.9 ** 2000; m.opt.state[1] = 0.999 ** 2000
    _run(a, b, 16 + 5, check_at=(15, 20), seed0=100, flush_at=(18,))


@pytest.mark.parametrize("K,N", [(12, 400_000), (256, 24_000), (20, 300_000)])
def test_rows_spanning_warps_bit_identical(K, N):
    """K/4 not a power of two (K=12, 20) or > 32 float4 per row (K=256): a row's float4s sit in different warps /
    CTAs / grid-stride iterations of the sweep.  Several iterations per thread plus a mid-epoch flush."""
    a, b = _pair(N, K, 4, 256, layers="32,16", dropout="1.0,1.0")
    _run(a, b, 11, check_at=(3, 6, 10), flush_at=(5,))


@pytest.mark.parametrize("opt,K,N", [("Adagrad", 12, 200_000), ("Momentum", 256, 12_000), ("ftrl", 16, 300_000)])
def test_non_adam_sweeps_rows_spanning_warps(opt, K, N):
    from tf_repos_b200 import synth
    from tf_repos_b200.deepfm import DeepFM
    kw = dict(deep_layers="32,16", dropout="1.0,1.0", l2_reg=1e-3, learning_rate=0.01, optimizer=opt, device="cuda:0")
    a = DeepFM(39, N, K, 256, update_mode="exact", **kw)
    b = DeepFM(39, N, K, 256, update_mode="exact_deferred", epoch_steps=4, **kw)
    b.fm_v.var.copy_(a.fm_v.var); b.fm_w.var.copy_(a.fm_w.var); b.dense.flat.copy_(a.dense.flat)
    _run(a, b, 10, check_at=(3, 9), flush_at=(5,))


@pytest.mark.parametrize("regime", [0, 1, 2])
def test_packed_adam_loops_match_scalar_step(regime):
    """FMUL2/FADD2/FFMA2 loops (csrc/adam_packed.cuh) vs step_sparse<ADAM>, bit for bit, on 2^26 random 8-element
    states per (regime, steps); the validity check may reject a trajectory, never accept a wrong one."""
    from tf_repos_b200 import ops
    for steps, seed in ((1, 1), (3, 77), (4, 12345)):
        for lr, l2 in ((5e-4, 1e-4), (1.0e-2, 0.0), (3e-3, 1e-2)):
            bad, rejected, total = ops.selftest_adam_packed(regime, seed, 1 << 24, steps, lr, l2, torch.device("cuda"))
            assert bad == 0, (regime, steps, lr, l2, bad, rejected, total)
            assert total == 1 << 24 and rejected < 0.5 * total, (regime, steps, lr, l2, rejected, total)
