"""GPU parity tests, kernel by kernel, through the C ABI (tf_repos_b200.ops -> libctr_b200.so)
against the CPU oracle (oracle/).  Integer outputs must be bit-exact; fp32 outputs within
rtol 1e-5 (+ an atol scaled to the magnitude of the terms that were summed); the optimizer apply
is bit-exact given the same gradient."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _dev():
    return torch.device("cuda:0")


def _close(got, ref, scale=None, rtol=RTOL, what=""):
    got = got.detach().cpu().double().numpy() if torch.is_tensor(got) else np.asarray(got, dtype=np.float64)
    ref = ref.detach().cpu().double().numpy() if torch.is_tensor(ref) else np.asarray(ref, dtype=np.float64)
    s = float(np.abs(ref).max()) if scale is None else float(scale)
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol * max(s, 1e-30), err_msg=what)


def _rand_batch(B, F, N, K, seed, dup_heavy=False):
    g = torch.Generator().manual_seed(seed)
    if dup_heavy:
        ids = torch.randint(0, min(N, 7), (B, F), generator=g)
    else:
        ids = torch.randint(0, N, (B, F), generator=g)
    if B * F >= 2:
        ids.view(-1)[0] = 0
        ids.view(-1)[-1] = N - 1
    vals = torch.rand(B, F, generator=g) * 2 - 0.5
    V = torch.randn(N, K, generator=g) * 0.3
    W = torch.randn(N, generator=g) * 0.3
    return ids.to(torch.int32), vals.float(), V.float(), W.float()


# -------------------------------------------------------------------------------------------------
# K1 forward
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K", [4, 8, 16, 32, 64, 128, 256, 10])
@pytest.mark.parametrize("F", [1, 2, 39, 70])
def test_fm_embed_fwd_deepfm(K, F):
    from tf_repos_b200 import ops
    B, N = 37, 1000
    ids, vals, V, W = _rand_batch(B, F, N, K, seed=K * 100 + F)
    e = V[ids.long()] * vals[..., None]
    S = e.sum(1)
    y_v = 0.5 * (S ** 2 - (e ** 2).sum(1)).sum(1)
    y_w = (W[ids.long()] * vals).sum(1)
    d = _dev()
    x = torch.empty(B, F * K, device=d); yw = torch.empty(B, device=d)
    y2 = torch.empty(B, device=d); Sg = torch.empty(B, K, device=d)
    oob = torch.zeros(2, dtype=torch.int32, device=d)
    ops.fm_embed_fwd(ids.to(d), vals.to(d), V.to(d), W.to(d), ops.FM_DEEPFM, x=x, y_w=yw, y2=y2, S=Sg, oob=oob)
    assert torch.equal(x.cpu(), e.reshape(B, F * K)), "x = V[ids]*vals must be bit-exact (one multiply)"
    _close(Sg, S, what="S")
    _close(yw, y_w, scale=(W[ids.long()] * vals).abs().sum(1).max(), what="y_w")
    _close(y2, y_v, scale=(S ** 2).sum(1).max(), what="y_v")
    assert oob.tolist() == [0, 0]


@pytest.mark.parametrize("K", [8, 16, 64])
def test_fm_embed_fwd_int64_ids_and_modes(K):
    from tf_repos_b200 import ops
    B, F, N = 19, 39, 500
    ids, vals, V, W = _rand_batch(B, F, N, K, seed=K)
    e = V[ids.long()] * vals[..., None]
    S = e.sum(1)
    bi = 0.5 * (S ** 2 - (e ** 2).sum(1))
    d = _dev()
    # NFM mode with int64 ids (serving signature dtype, DeepFM.py:362), no x
    y2 = torch.empty(B, K, device=d); Sg = torch.empty(B, K, device=d); yw = torch.empty(B, device=d)
    ops.fm_embed_fwd(ids.long().to(d), vals.to(d), V.to(d), W.to(d), ops.FM_NFM, x=None, y_w=yw, y2=y2, S=Sg)
    _close(y2, bi, scale=(S ** 2).max(), what="bi")
    # PLAIN mode, no W
    x = torch.empty(B, F * K, device=d)
    ops.fm_embed_fwd(ids.to(d), vals.to(d), V.to(d), None, ops.FM_PLAIN, x=x)
    assert torch.equal(x.cpu(), e.reshape(B, F * K))


def test_fm_embed_fwd_known_answers_and_oob():
    from tf_repos_b200 import ops
    d = _dev()
    B, F, N, K = 4, 39, 100, 16
    ids, vals, V, W = _rand_batch(B, F, N, K, seed=3)
    # single active field => y_v == 0 exactly (S == e)
    v1 = torch.zeros_like(vals); v1[:, 5] = vals[:, 5]
    x = torch.empty(B, F * K, device=d); yw = torch.empty(B, device=d)
    y2 = torch.empty(B, device=d); Sg = torch.empty(B, K, device=d)
    ops.fm_embed_fwd(ids.to(d), v1.to(d), V.to(d), W.to(d), ops.FM_DEEPFM, x=x, y_w=yw, y2=y2, S=Sg)
    assert torch.all(y2 == 0)
    # all-zero values => everything 0
    ops.fm_embed_fwd(ids.to(d), torch.zeros_like(vals).to(d), V.to(d), W.to(d), ops.FM_DEEPFM, x=x, y_w=yw, y2=y2, S=Sg)
    assert torch.all(x == 0) and torch.all(yw == 0) and torch.all(y2 == 0)
    # out-of-range ids are counted (TF raises InvalidArgumentError) and contribute zero
    bad = ids.clone(); bad[1, 3] = N + 5; bad[2, 0] = -1
    oob = torch.zeros(2, dtype=torch.int32, device=d)
    ops.fm_embed_fwd(bad.to(d), vals.to(d), V.to(d), W.to(d), ops.FM_DEEPFM, x=x, y_w=yw, y2=y2, S=Sg, oob=oob)
    assert oob[0].item() == 2 and oob[1].item() in (N + 5, -1)
    assert torch.all(x.view(B, F, K)[1, 3] == 0) and torch.all(x.view(B, F, K)[2, 0] == 0)


def test_fm_embed_fwd_empty_batch():
    from tf_repos_b200 import ops
    d = _dev()
    ids = torch.zeros(0, 39, dtype=torch.int32, device=d); vals = torch.zeros(0, 39, device=d)
    V = torch.zeros(10, 16, device=d)
    ops.fm_embed_fwd(ids, vals, V, None, ops.FM_PLAIN, x=torch.empty(0, 39 * 16, device=d))


# -------------------------------------------------------------------------------------------------
# K2 backward vs autograd (fp64 truth and fp32)
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K", [4, 8, 16, 32, 256, 10])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_fm_embed_bwd(K, mode):
    from tf_repos_b200 import ops
    B, F, N = 23, 39, 300
    ids, vals, V, W = _rand_batch(B, F, N, K, seed=K + mode)
    g = torch.Generator().manual_seed(1)
    rows = V[ids.long()].double().requires_grad_()
    wr = W[ids.long()].double().requires_grad_()
    e = rows * vals.double()[..., None]
    S = e.sum(1)
    dX = torch.randn(B, F * K, generator=g).double() * 0.1
    dyv = torch.randn(B, generator=g).double()
    dbi = torch.randn(B, K, generator=g).double()
    dyw = torch.randn(B, generator=g).double()
    obj = (e.reshape(B, -1) * dX).sum() + ((wr * vals.double()).sum(1) * dyw).sum()
    if mode == 0:
        obj = obj + ((0.5 * (S ** 2 - (e ** 2).sum(1)).sum(1)) * dyv).sum()
    elif mode == 1:
        obj = obj + ((0.5 * (S ** 2 - (e ** 2).sum(1))) * dbi).sum()
    obj.backward()
    d = _dev()
    x32 = (V[ids.long()] * vals[..., None]).reshape(B, F * K)
    S32 = (V[ids.long()] * vals[..., None]).sum(1)
    g_rows = torch.empty(B * F, K, device=d); g_w = torch.empty(B * F, device=d)
    dy2 = None if mode == 2 else (dyv if mode == 0 else dbi).float().to(d).contiguous()
    ops.fm_embed_bwd(vals.to(d), x32.to(d), S32.to(d), dX.float().to(d), dy2, dyw.float().to(d), K, mode, g_rows, g_w)
    _close(g_rows, rows.grad.reshape(B * F, K), what="g_rows")
    _close(g_w, wr.grad.reshape(-1), what="g_w")


# -------------------------------------------------------------------------------------------------
# K3: unique / inverse / segment offsets / perm are bit-exact
# -------------------------------------------------------------------------------------------------
def _check_unique(ids_np, N):
    from oracle import tf_semantics as tfs
    from tf_repos_b200 import ops
    d = _dev()
    n = ids_np.shape[0]
    uw = ops.UniqueWorkspace(n, N, d)
    ops.unique_segment(torch.from_numpy(ids_np).to(d), uw)
    torch.cuda.synchronize()
    perm, uniq, inverse, seg = tfs.unique_segment_reference(ids_np)
    U = uniq.shape[0]
    assert uw.n_uniq.item() == U
    if n:
        np.testing.assert_array_equal(uw.perm.cpu().numpy()[:n], perm)
        np.testing.assert_array_equal(uw.uniq.cpu().numpy()[:U], uniq)
        np.testing.assert_array_equal(uw.inverse.cpu().numpy()[:n], inverse)
    np.testing.assert_array_equal(uw.seg_offsets.cpu().numpy()[: U + 1], seg)
    ll = uw.long_list.cpu().numpy()
    lens = seg[1:] - seg[:-1]
    expect_long = set(np.nonzero(lens > ops.LONG_SEG)[0].tolist())
    assert ll[0] == len(expect_long) and set(ll[1: 1 + ll[0]].tolist()) == expect_long
    return uw


@pytest.mark.parametrize("n,N", [(1, 10), (2, 2), (5, 1), (31, 100), (2048, 10_000), (2049, 117_581),
                                 (50_000, 1 << 20), (319_488, 200_000_000), (100_000, 1_000_000_000),
                                 (70_000, (1 << 31) - 1)])
def test_unique_segment_random(n, N):
    rng = np.random.default_rng(n)
    ids = rng.integers(0, N, size=n, dtype=np.int64).astype(np.int32)
    ids[0] = N - 1
    _check_unique(ids, N)


def test_unique_segment_edge_cases():
    _check_unique(np.zeros(0, dtype=np.int32), 100)                       # empty
    _check_unique(np.full(5000, 7, dtype=np.int32), 100)                  # one long run
    _check_unique(np.arange(4096, dtype=np.int32)[::-1].copy(), 4096)     # all distinct, reversed
    rng = np.random.default_rng(0)
    ids = rng.integers(0, 5, size=10_000).astype(np.int32)                # 5 very long runs
    _check_unique(ids, 1000)
    from tf_repos_b200 import synth
    ids, _, _ = synth.criteo_batch(8192, 200_000_000, 39, seed=1)         # Criteo layout: 13 runs of 8192
    _check_unique(ids.reshape(-1).numpy(), 200_000_000)


@pytest.mark.parametrize("K", [4, 8, 16, 32, 64, 128, 256, 10])
def test_segment_sum_rows(K):
    from tf_repos_b200 import ops
    d = _dev()
    rng = np.random.default_rng(K)
    n = 6000
    ids = np.concatenate([rng.integers(0, 3, size=3000), rng.integers(3, 2000, size=3000)]).astype(np.int32)
    rng.shuffle(ids)
    uw = _check_unique(ids, 2000)
    g = rng.standard_normal((n, K)).astype(np.float32)
    gw = rng.standard_normal(n).astype(np.float32)
    U = uw.n_uniq.item()
    g_uniq = torch.zeros(n, K, device=d); gw_uniq = torch.zeros(n, device=d)
    ops.segment_sum_rows(torch.from_numpy(g).to(d), torch.from_numpy(gw).to(d), uw, K, g_uniq, gw_uniq)
    uniq, inv = np.unique(ids, return_inverse=True)
    ref = np.zeros((U, K)); np.add.at(ref, inv, g.astype(np.float64))
    refw = np.zeros(U); np.add.at(refw, inv, gw.astype(np.float64))
    mag = np.zeros((U, K)); np.add.at(mag, inv, np.abs(g).astype(np.float64))
    err = np.abs(g_uniq.cpu().numpy()[:U] - ref)
    assert np.all(err <= 2e-6 * mag + 1e-30), f"max rel err {np.max(err / (mag + 1e-30))}"
    _close(gw_uniq[:U], refw, scale=np.abs(gw).sum() / 3)
    # run-to-run determinism (fixed trees, no float atomics)
    g2 = torch.zeros(n, K, device=d)
    ops.segment_sum_rows(torch.from_numpy(g).to(d), None, uw, K, g2, None)
    assert torch.equal(g2[:U], g_uniq[:U])
    # short runs are summed in occurrence order == TF's CPU order: bit-exact vs sequential fp32
    seq = np.zeros((U, K), dtype=np.float32); np.add.at(seq, inv, g)
    lens = np.bincount(inv)
    short = lens <= ops.LONG_SEG
    np.testing.assert_array_equal(g_uniq.cpu().numpy()[:U][short], seq[short])


# -------------------------------------------------------------------------------------------------
# K4: optimizers are bit-exact against the oracle given the same gradient
# -------------------------------------------------------------------------------------------------
def _oracle_rows(opt_name, var, slots, g, l2, lr, step_fn_sparse=True, adam=None):
    from oracle import tf_semantics as tfs
    var = var.clone(); slots = [s.clone() for s in slots]
    g = g + torch.tensor(l2) * var if l2 is not None else g
    if opt_name == "Adam":
        fn = tfs.adam_sparse_ if step_fn_sparse else tfs.adam_dense_
        fn(var, slots[0], slots[1], g, adam.lr_t(), adam.b1, adam.b2, adam.eps)
    elif opt_name == "Adagrad":
        tfs.adagrad_(var, slots[0], g, torch.tensor(lr))
    elif opt_name == "Momentum":
        tfs.momentum_(var, slots[0], g, torch.tensor(lr), torch.tensor(0.95))
    else:
        tfs.ftrl_(var, slots[0], slots[1], g, lr)
    return var, slots


@pytest.mark.parametrize("opt_name", ["Adam", "Adagrad", "Momentum", "ftrl"])
@pytest.mark.parametrize("K", [1, 16, 10, 256])
def test_optimizer_sparse_sweep_patch_bit_exact(opt_name, K):
    from oracle import tf_semantics as tfs
    from tf_repos_b200 import engine, ops
    d = _dev()
    N, n, lr, l2 = 997, 300, 0.01, 1e-3
    g = torch.Generator().manual_seed(K)
    ost = engine.OptimizerState(opt_name, lr, l2, d)
    adam = tfs.AdamHyper(lr)
    var = (torch.randn(N, K, generator=g) * 0.1).float()
    ns = ost.n_slots
    slots = [(torch.rand(N, K, generator=g) * 0.01 + ost.slot_init(i)).float() for i in range(ns)]
    uniq = torch.sort(torch.randperm(N, generator=g)[:n])[0].to(torch.int32)
    g_uniq = (torch.randn(n, K, generator=g) * 0.05).float()
    for step in range(3):
        ost.tick()
        dv, ds = var.to(d), [s.to(d) for s in slots]
        s1 = ds[1] if ns > 1 else None
        n_uniq = torch.tensor([n], dtype=torch.int32, device=d)
        # (a) exact composition: stage -> sweep -> patch
        stage = torch.zeros(3 * n * K, device=d)
        ops.opt_sparse_rows(ost.opt, dv, ds[0], s1, uniq.to(d), n_uniq, g_uniq.to(d), n, K, ost.record(0), stage)
        part = torch.zeros(ops.sweep_partials_count(), device=d)
        ops.opt_dense_sweep(ost.opt, dv, ds[0], s1, ost.record(0), part)
        ops.opt_patch_rows(dv, ds[0], s1, uniq.to(d), n_uniq, stage, n, K, ns)
        # oracle: every row is an index of the sparse apply with G = l2*var (+ segment sum)
        G = torch.tensor(l2) * var
        G[uniq.long()] = g_uniq + G[uniq.long()]
        rv, rs = _oracle_rows(opt_name, var, slots, G, None, lr, True, adam)
        assert torch.equal(dv.cpu(), rv), f"{opt_name} var step {step}"
        for a, b in zip(ds, rs):
            assert torch.equal(a.cpu(), b)
        _close(part.sum().cpu(), (var.double() ** 2).sum(), rtol=1e-5)
        # (b) lazy: rows only, in place
        dv2, ds2 = var.to(d), [s.to(d) for s in slots]
        ops.opt_sparse_rows(ost.opt, dv2, ds2[0], ds2[1] if ns > 1 else None, uniq.to(d), n_uniq, g_uniq.to(d), n, K,
                            ost.record(0), None)
        lv = var.clone(); ls = [s.clone() for s in slots]
        sub_v, sub_s = _oracle_rows(opt_name, var[uniq.long()], [s[uniq.long()] for s in slots],
                                    g_uniq, l2, lr, True, adam)
        lv[uniq.long()] = sub_v
        for s, r in zip(ls, sub_s):
            s[uniq.long()] = r
        assert torch.equal(dv2.cpu(), lv)
        for a, b in zip(ds2, ls):
            assert torch.equal(a.cpu(), b)
        var, slots = rv, rs
        adam.finish()


@pytest.mark.parametrize("opt_name", ["Adam", "Adagrad", "Momentum", "ftrl"])
def test_optimizer_dense_grad_bit_exact(opt_name):
    from oracle import tf_semantics as tfs
    from tf_repos_b200 import engine, ops
    d = _dev()
    n, lr = 10_001, 0.003
    g = torch.Generator().manual_seed(5)
    for rec, l2 in ((1, None), (2, 2e-3)):
        ost = engine.OptimizerState(opt_name, lr, 2e-3, d)
        adam = tfs.AdamHyper(lr)
        var = (torch.randn(n, generator=g) * 0.1).float()
        slots = [(torch.rand(n, generator=g) * 0.01 + ost.slot_init(i)).float() for i in range(ost.n_slots)]
        for step in range(3):
            grad = (torch.randn(n, generator=g) * 0.02).float()
            ost.tick()
            dv, ds = var.to(d), [s.to(d) for s in slots]
            ops.opt_dense_grad(ost.opt, dv, ds[0], ds[1] if ost.n_slots > 1 else None, grad.to(d), ost.record(rec))
            var, slots = _oracle_rows(opt_name, var, slots, grad, l2, lr, False, adam)
            assert torch.equal(dv.cpu(), var)
            for a, b in zip(ds, slots):
                assert torch.equal(a.cpu(), b)
            adam.finish()


def test_adam_tick_matches_tf_beta_powers():
    from oracle import tf_semantics as tfs
    from tf_repos_b200 import engine
    ost = engine.OptimizerState("Adam", 5e-4, 1e-4, _dev())
    adam = tfs.AdamHyper(5e-4)
    for _ in range(50):
        ost.tick()
        assert ost.hyper[0, 0].item() == adam.lr_t().item()
        assert ost.hyper[1, 0].item() == adam.lr_t().item()
        adam.finish()
    assert ost.state[3].item() == 50.0


def test_logit_loss_and_reductions():
    from oracle import tf_semantics as tfs
    from tf_repos_b200 import ops
    d = _dev()
    g = torch.Generator().manual_seed(0)
    for B in (1, 37, 8192, 10_000):
        ya, yb, yc = [torch.randn(B, generator=g) * 3 for _ in range(3)]
        bias = torch.tensor([0.3]); lab = (torch.rand(B, generator=g) < 0.3).float()
        y = bias + ya + yb + yc
        y_o, p_o, dy_o = [torch.empty(B, device=d) for _ in range(3)]
        l_o = torch.zeros(1, device=d); db_o = torch.zeros(1, device=d)
        ops.logit_loss(bias.to(d), ya.to(d), yb.to(d), yc.to(d), lab.to(d), B, y=y_o, pred=p_o, loss_ce=l_o, dy=dy_o, dbias=db_o)
        assert torch.equal(y_o.cpu(), y)
        _close(p_o, tfs.sigmoid(y.double()), rtol=2e-6)
        ce = tfs.sigmoid_cross_entropy_with_logits(y.double(), lab.double()).mean()
        _close(l_o, ce, rtol=1e-5)
        dy = (tfs.sigmoid(y.double()) - lab.double()) / B
        _close(dy_o, dy, rtol=1e-5)
        _close(db_o, dy.sum(), scale=dy.abs().sum(), rtol=1e-5)
    t = torch.randn(1_000_003, generator=g)
    out = torch.zeros(1, device=d); ws = torch.empty(1024, device=d)
    ops.l2_loss(t.to(d), out, ws)
    _close(out, (t.double() ** 2).sum() / 2, rtol=1e-5)


def test_init_trunc_normal_distribution():
    from tf_repos_b200 import ops
    t = torch.empty(2_000_000, device=_dev())
    ops.init_trunc_normal(t, 0.01, 123)
    assert t.abs().max().item() <= 0.02 + 1e-9
    assert abs(t.mean().item()) < 1e-4
    # std of a 2-sigma truncated normal is 0.8796*sigma
    assert abs(t.std().item() / 0.01 - 0.8796) < 0.01
    t2 = torch.empty_like(t); ops.init_trunc_normal(t2, 0.01, 123)
    assert torch.equal(t, t2)


# -------------------------------------------------------------------------------------------------
# dense layers (csrc/fc.cu) vs fp64 torch
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,Kd,Nd", [(8192, 624, 256), (8192, 256, 128), (1000, 128, 64), (37, 50, 19), (256, 312, 400)])
@pytest.mark.parametrize("drop", [False, True])
def test_fc_fwd_bwd(M, Kd, Nd, drop):
    from tf_repos_b200 import ops
    d = _dev()
    g = torch.Generator().manual_seed(M + Nd)
    x = torch.randn(M, Kd, generator=g); W = torch.randn(Kd, Nd, generator=g) / Kd ** 0.5
    b = torch.randn(Nd, generator=g) * 0.1
    keep = 0.8
    mask = (torch.rand(M, Nd, generator=g) < keep).float() if drop else None
    dOut = torch.randn(M, Nd, generator=g)
    xd, Wd, bd = x.double().requires_grad_(), W.double().requires_grad_(), b.double().requires_grad_()
    out = torch.empty(M, Nd, device=d)
    ops.fc_fwd(x.to(d), W.to(d), b.to(d), mask.to(d) if drop else None, keep, 1, out)
    pre = xd @ Wd + bd
    # a pre-activation within rounding distance of 0 may land on either side of the relu: the reference
    # takes the GPU's side for those (otherwise one flipped unit perturbs a whole row of dIn)
    act = (out.cpu() > 0).double() if not drop else ((out.cpu() > 0) | (mask == 0) & (pre.detach() > 0)).double()
    near = pre.detach().abs() < 1e-4
    gate = torch.where(near, act, (pre.detach() > 0).double())
    z = pre * gate
    out_ref = z / keep * mask.double() if drop else z
    out_ref.backward(dOut.double())
    _close(out, out_ref, rtol=2e-6 * Kd ** 0.5, what="fc out")
    dO = dOut.to(d).clone()
    dIn = torch.empty(M, Kd, device=d); dW = torch.empty(Kd, Nd, device=d); db = torch.empty(Nd, device=d)
    ws = torch.empty(ops.fc_bwd_workspace_bytes(M, Kd, Nd), dtype=torch.uint8, device=d)
    ops.fc_bwd(x.to(d), W.to(d), out, mask.to(d) if drop else None, keep, dO, 1, dIn, dW, db, ws)
    _close(dIn, xd.grad, rtol=1e-5, what="dIn")
    _close(dW, Wd.grad, rtol=1e-5, what="dW")
    _close(db, bd.grad, rtol=1e-5, what="db")
    # determinism
    dW2 = torch.empty_like(dW); dO2 = dOut.to(d).clone()
    ops.fc_bwd(x.to(d), W.to(d), out, mask.to(d) if drop else None, keep, dO2, 1, None, dW2, db, ws)
    assert torch.equal(dW, dW2)


@pytest.mark.parametrize("Ka,Kb", [(64, 0), (624, 64), (5, 3)])
def test_fc1_fwd_bwd(Ka, Kb):
    from tf_repos_b200 import ops
    d = _dev()
    M = 777
    g = torch.Generator().manual_seed(Ka)
    a = torch.randn(M, Ka, generator=g); bb = torch.randn(M, Kb, generator=g) if Kb else None
    w = torch.randn(Ka + Kb, generator=g); bias = torch.tensor([0.3]); dy = torch.randn(M, generator=g)
    cat = torch.cat([a, bb], 1) if Kb else a
    catd, wd, bd = cat.double().requires_grad_(), w.double().requires_grad_(), bias.double().requires_grad_()
    y_ref = catd @ wd + bd
    y_ref.backward(dy.double())
    y = torch.empty(M, device=d)
    ops.fc1_fwd(a.to(d), bb.to(d) if Kb else None, w.to(d), bias.to(d), y)
    _close(y, y_ref, rtol=1e-5)
    d_a = torch.empty(M, Ka, device=d); d_b = torch.empty(M, Kb, device=d) if Kb else None
    dw = torch.empty(Ka + Kb, device=d); db = torch.empty(1, device=d)
    ws = torch.empty(ops.fc1_bwd_workspace_bytes(M, Ka, Kb), dtype=torch.uint8, device=d)
    ops.fc1_bwd(a.to(d), bb.to(d) if Kb else None, w.to(d), dy.to(d), d_a, d_b, dw, db, ws)
    _close(d_a, catd.grad[:, :Ka], rtol=1e-6)
    if Kb:
        _close(d_b, catd.grad[:, Ka:], rtol=1e-6)
    _close(dw, wd.grad, rtol=1e-5)
    _close(db, bd.grad, scale=dy.abs().sum(), rtol=1e-5)


def test_dropout_mask_rate_and_step_dependence():
    from tf_repos_b200 import ops
    d = _dev()
    m = torch.empty(1_000_000, device=d); m2 = torch.empty_like(m)
    step = torch.tensor([3.0], device=d)
    ops.dropout_mask(m, 0.8, 7, step)
    assert set(m.unique().tolist()) <= {0.0, 1.0} and abs(m.mean().item() - 0.8) < 2e-3
    ops.dropout_mask(m2, 0.8, 7, step)
    assert torch.equal(m, m2)
    step.fill_(4.0)
    ops.dropout_mask(m2, 0.8, 7, step)
    assert not torch.equal(m, m2)


@pytest.mark.gpu
def test_inrange_sqrt_div_fast_paths_are_correctly_rounded():
    """The epoch sweeps' grouped IEEE sqrt/div fast paths (optim_steps.cuh) vs the compiler's sqrt.rn/div.rn:
    bit-identical on 2^31 seeded operands (random + hard mantissa patterns over the whole guarded range)."""
    from tf_repos_b200 import ops
    for seed in (1, 12345):
        bad_sqrt, bad_div = ops.selftest_divsqrt(seed, 1 << 30, torch.device("cuda"))
        assert (bad_sqrt, bad_div) == (0, 0)
