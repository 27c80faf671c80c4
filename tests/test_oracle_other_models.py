"""Known answers for the oracle's DCN / NFM / PNN / AFM / DIN forward passes and losses: each model_fn of the reference is
restated here a second time as per-sample, per-element Python loops (no tensor ops, no code shared with oracle/), read
line by line from the script, and the oracle (fp64) must reproduce the logits to 1e-12.  This is what catches a
transposed index, a wrong pair order, a mask on the wrong axis or a misplaced concat in the vectorised restatement the
GPU parity tests trust.  (The DeepFM oracle has its own known-answer file, test_oracle_deepfm.py.)

Reference lines: DCN.py:134-145,179-183,198-199; NFM.py:118-128,152-155; PNN.py:131-167,190-193; AFM.py:123-167;
DIN.py:143-183,199-226."""
import math

import pytest
import torch

from oracle import models as om

F64 = torch.float64


def _fill(model, seed, scale=0.4):
    """Random values everywhere (the initialisers leave biases at 0, which would hide a missing bias add)."""
    g = torch.Generator().manual_seed(seed)
    for n, p in model.params.items():
        p.copy_(torch.randn(p.shape, generator=g, dtype=F64) * scale)
    return {n: p.tolist() for n, p in model.params.items()}


def _fc(x, W, b, act=None):
    out = []
    for j in range(len(b)):
        s = b[j]
        for i in range(len(x)):
            s += x[i] * W[i][j]
        if act == "relu":
            s = max(s, 0.0)
        elif act == "sigmoid":
            s = 1.0 / (1.0 + math.exp(-s))
        out.append(s)
    return out


def _mlp(x, p, scope, n_layers):
    for i in range(n_layers):
        x = _fc(x, p[f"{scope}/mlp{i}/weights"], p[f"{scope}/mlp{i}/biases"], "relu")
    return x


def _ce(y, z):            # tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log(1 + exp(-|x|))
    return max(y, 0.0) - y * z + math.log1p(math.exp(-abs(y)))


def _l2(t):               # tf.nn.l2_loss = sum(t^2)/2
    flat = torch.tensor(t, dtype=F64).reshape(-1).tolist()
    return sum(v * v for v in flat) / 2


def _libsvm_batch(B, F, N, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, N, (B, F), generator=g)
    vals = torch.rand(B, F, generator=g, dtype=F64) + 0.2
    labels = (torch.rand(B, generator=g) < 0.4).to(F64)
    return ids, vals, labels


def _check(model, batch, labels, y_loops, reg_loops):
    y = model.predict(batch)["y"].tolist()
    assert len(y) == len(y_loops)
    for a, b in zip(y, y_loops):
        assert abs(a - b) <= 1e-12 * max(1.0, abs(b)), (a, b)
    loss = sum(_ce(v, z) for v, z in zip(y_loops, labels.tolist())) / len(y_loops) + reg_loops
    got = model.evaluate(batch, labels)["loss"]
    assert abs(got - loss) <= 1e-12 * max(1.0, abs(loss)), (got, loss)


def test_dcn_cross_network_and_stacked_head():
    B, F, N, K, L = 3, 4, 17, 3, 3
    m = om.DCN(F, N, K, deep_layers="5,4", cross_layers=L, dropout="1.0,1.0", l2_reg=0.01, dtype=F64, seed=1)
    p = _fill(m, 11)
    ids, vals, labels = _libsvm_batch(B, F, N, 3)
    ys = []
    for b in range(B):
        x0 = [p["emb"][ids[b, f]][k] * float(vals[b, f]) for f in range(F) for k in range(K)]       # [F*K], field-major
        xl = list(x0)
        for l in range(L):
            xlw = sum(xl[d] * p["cross_w"][l][d] for d in range(F * K))                             # xl . w_l (a scalar)
            xl = [x0[d] * xlw + xl[d] + p["cross_b"][l][d] for d in range(F * K)]
        h = _mlp(x0, p, "Deep-Network", 2)                                                          # the deep net reads x0
        ys.append(_fc(xl + h, p["DCN-out/out_layer/weights"], p["DCN-out/out_layer/biases"])[0])    # concat [xL, deep]
    reg = 0.01 * (_l2(p["cross_b"]) + _l2(p["cross_w"]) + _l2(p["emb"]))
    _check(m, {"feat_ids": ids, "feat_vals": vals}, labels, ys, reg)


def test_nfm_bi_interaction_pooling():
    B, F, N, K = 3, 5, 19, 4
    m = om.NFM(F, N, K, deep_layers="6,3", dropout="1.0,1.0,1.0", l2_reg=0.02, dtype=F64, seed=2)
    p = _fill(m, 12)
    ids, vals, labels = _libsvm_batch(B, F, N, 4)
    ys = []
    for b in range(B):
        lin = sum(p["linear"][ids[b, f]] * float(vals[b, f]) for f in range(F))
        e = [[p["emb"][ids[b, f]][k] * float(vals[b, f]) for k in range(K)] for f in range(F)]
        bi = [0.5 * (sum(e[f][k] for f in range(F)) ** 2 - sum(e[f][k] ** 2 for f in range(F))) for k in range(K)]
        h = _mlp(bi, p, "Deep-part", 2)
        ys.append(p["bias"][0] + lin + _fc(h, p["Deep-part/deep_out/weights"], p["Deep-part/deep_out/biases"])[0])
    _check(m, {"feat_ids": ids, "feat_vals": vals}, labels, ys, 0.02 * (_l2(p["linear"]) + _l2(p["emb"])))


@pytest.mark.parametrize("model_type", ["FNN", "Inner", "Outer"])
def test_pnn_product_layer_pair_order(model_type):
    B, F, N, K = 3, 4, 13, 3
    m = om.PNN(F, N, K, model_type=model_type, deep_layers="5,3", dropout="1.0,1.0", l2_reg=0.03, dtype=F64, seed=3)
    p = _fill(m, 13)
    ids, vals, labels = _libsvm_batch(B, F, N, 5)
    ys = []
    for b in range(B):
        lin = sum(p["linear"][ids[b, f]] * float(vals[b, f]) for f in range(F))
        e = [[p["emb"][ids[b, f]][k] * float(vals[b, f]) for k in range(K)] for f in range(F)]
        z = [e[f][k] for f in range(F) for k in range(K)]
        pairs = [(i, j) for i in range(F - 1) for j in range(i + 1, F)]                              # row/col lists
        if model_type == "Inner":
            z = z + [sum(e[i][k] * e[j][k] for k in range(K)) for i, j in pairs]
        elif model_type == "Outer":                                                                  # 'api,apj->apij'
            z = z + [e[i][a] * e[j][c] for i, j in pairs for a in range(K) for c in range(K)]
        h = _mlp(z, p, "Deep-part", 2)
        ys.append(p["bias"][0] + lin + _fc(h, p["Deep-part/deep_out/weights"], p["Deep-part/deep_out/biases"])[0])
    _check(m, {"feat_ids": ids, "feat_vals": vals}, labels, ys, 0.03 * (_l2(p["linear"]) + _l2(p["emb"])))


def test_afm_attention_softmax_is_over_the_pairs():
    B, F, N, K = 3, 4, 11, 3
    m = om.AFM(F, N, K, attention_layers="5", dropout="1.0,1.0", l2_reg=0.5, dtype=F64, seed=4)
    p = _fill(m, 14)
    ids, vals, labels = _libsvm_batch(B, F, N, 6)
    A, PO = "Attention-part", "Attention-based-Pooling"
    ys = []
    for b in range(B):
        lin = sum(p["linear"][ids[b, f]] * float(vals[b, f]) for f in range(F))
        e = [[p["emb"][ids[b, f]][k] * float(vals[b, f]) for k in range(K)] for f in range(F)]
        pw = [[e[i][k] * e[j][k] for k in range(K)] for i in range(F) for j in range(i + 1, F)]      # pair-major
        a = [_fc(_fc(v, p[f"{A}/mlp0/weights"], p[f"{A}/mlp0/biases"], "relu"),
                 p[f"{A}/attention_out/weights"], p[f"{A}/attention_out/biases"])[0] for v in pw]
        mx = max(a)
        ex = [math.exp(v - mx) for v in a]
        soft = [v / sum(ex) for v in ex]                                                             # softmax over pairs
        y_emb = [sum(soft[q] * pw[q][k] for q in range(len(pw))) for k in range(K)]
        ys.append(p["bias"][0] + lin + _fc(y_emb, p[f"{PO}/deep_out/weights"], p[f"{PO}/deep_out/biases"])[0])
    _check(m, {"feat_ids": ids, "feat_vals": vals}, labels, ys, 0.5 * (_l2(p["linear"]) + _l2(p["emb"])))


def _din_batch(B, Fp, N, P, seed):
    g = torch.Generator().manual_seed(seed)
    feat_ids = torch.randint(1, N, (B, Fp), generator=g)
    a_ids = torch.randint(1, N, (3, B), generator=g)
    n_int = torch.randint(1, 4, (B,), generator=g)                     # ad "int" ids: 1..3 per sample (multi-hot)
    a_int_off = torch.cat([torch.zeros(1, dtype=torch.long), n_int.cumsum(0)])
    a_int_ids = torch.randint(1, N, (int(a_int_off[-1]),), generator=g)
    u_ids = torch.randint(1, N, (4, B, P), generator=g)
    u_wgt = torch.rand(4, B, P, generator=g, dtype=F64) + 0.1
    for f in range(4):                                                 # ragged behaviours: 0-padded ids AND weights
        for b in range(B):
            n = int(torch.randint(0, P + 1, (1,), generator=g))
            u_ids[f, b, n:] = 0
            u_wgt[f, b, n:] = 0.0
    labels = (torch.rand(B, generator=g) < 0.4).to(F64)
    return {"feat_ids": feat_ids, "a_ids": a_ids, "a_int_ids": a_int_ids, "a_int_off": a_int_off.to(torch.int32),
            "u_ids": u_ids, "u_wgt": u_wgt}, labels


@pytest.mark.parametrize("attention_pooling", [True, False])
def test_din_attention_unit_mask_and_concat_order(attention_pooling):
    B, Fp, N, K, P = 3, 2, 23, 3, 4
    m = om.DIN(Fp, N, K, deep_layers="6,4", dropout="1.0,1.0", attention_layers="7", attention_pooling=attention_pooling,
               l2_reg=0.05, dtype=F64, seed=5)
    p = _fill(m, 15)
    batch, labels = _din_batch(B, Fp, N, P, 7)
    # an id-0 slot with a NON-zero weight: embedding row 0 still enters the attention MLP, the mask removes it from the sum
    batch["u_ids"][1, 0, P - 1] = 0
    batch["u_wgt"][1, 0, P - 1] = 0.7
    E, ATT = p["embeddings"], "Field-wise-Pooling-layer"
    off = batch["a_int_off"].tolist()
    ys = []
    for b in range(B):
        common = [E[batch["feat_ids"][b, f]][k] for f in range(Fp) for k in range(K)]
        a = [[E[batch["a_ids"][t, b]][k] for k in range(K)] for t in range(3)]
        a.append([sum(E[batch["a_int_ids"][i]][k] for i in range(off[b], off[b + 1])) for k in range(K)])   # sum combiner
        u = []
        for f in range(4):                                             # cat, shop, brand, int -- each against its own ad field
            acc = [0.0] * K
            for q in range(P):
                i, w = int(batch["u_ids"][f, b, q]), float(batch["u_wgt"][f, b, q])
                ub = [E[i][k] * w for k in range(K)]
                if attention_pooling:
                    x = ub + [ub[k] - a[f][k] for k in range(K)] + a[f]
                    x = _fc(x, p[f"{ATT}/att_fc0/weights"], p[f"{ATT}/att_fc0/biases"], "relu")      # width = deep_layers[0]
                    att = _fc(x, p[f"{ATT}/att_out/weights"], p[f"{ATT}/att_out/biases"], "sigmoid")[0]
                    keep = 1.0 if i > 0 else 0.0                                                    # dense_ids > 0
                    acc = [acc[k] + ub[k] * att * keep for k in range(K)]
                else:                                                  # embedding_lookup_sparse(..., combiner="sum")
                    acc = [acc[k] + ub[k] for k in range(K)]
            u.append(acc)
        x = common + u[0] + u[1] + u[2] + u[3] + a[0] + a[1] + a[2] + a[3]
        h = _mlp(x, p, "MLP-layer", 2)
        ys.append(_fc(h, p["DIN-out/din_out/weights"], p["DIN-out/din_out/biases"])[0])
    if not attention_pooling:
        # sparse "sum" pooling only sees the entries the SparseTensor holds; the dense 0-padded layout agrees because the
        # padded weights are 0 -- undo the deliberate (id 0, weight 0.7) slot, which a SparseTensor could not contain
        batch["u_wgt"][1, 0, P - 1] = 0.0
        ys[0] = None
    y = m.predict(batch)["y"].tolist()
    for bb, (got, want) in enumerate(zip(y, ys)):
        if want is not None:
            assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (bb, got, want)
    if attention_pooling:
        loss = sum(_ce(v, z) for v, z in zip(ys, labels.tolist())) / B + 0.05 * _l2(E)
        assert abs(m.evaluate(batch, labels)["loss"] - loss) <= 1e-12 * max(1.0, abs(loss))
        assert m.layers[0] == 6 and p[f"{ATT}/att_fc0/biases"].__len__() == 6      # quirk Q5: not attention_layers' 7
