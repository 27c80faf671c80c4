"""oracle/wide_deep.py (restated TF canned estimators of wide_n_deep.py:92-151): closed-form known answers, an fp64 twin,
the CSV reader and the flag surface.  CPU only."""
import math
import os
import subprocess
import sys

import numpy as np
import torch

from oracle import tf_semantics as tfs
from oracle import wide_deep as owd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _batch(B, seed=0, oob=False):
    g = torch.Generator().manual_seed(seed)
    dense = torch.rand(B, 13, generator=g)
    cat = torch.randint(0, 10000, (B, 26), generator=g, dtype=torch.int64).to(torch.int32)
    if oob:
        cat[0, 0] = 12345
        cat[1, 3] = -7
    labels = (torch.rand(B, generator=g) < 0.25).float()
    return dense, cat, labels


def test_column_order_and_learning_rates():
    assert owd.NUM_SORTED == [0, 9, 10, 11, 12, 1, 2, 3, 4, 5, 6, 7, 8]      # I1, I10..I13, I2..I9
    assert owd.WideDeep(model_type="deep").dnn_lr == 0.05
    m = owd.WideDeep(model_type="wide_n_deep")
    assert m.dnn_lr == 0.001 and m.linear_lr == 0.005
    assert abs(owd.WideDeep(model_type="wide").linear_lr - 1 / math.sqrt(39)) < 1e-12


def test_wide_first_step_closed_form():
    """All linear weights start at 0 => logit 0, p = 0.5; one Ftrl step from accum 0.1:
    w = -g*lr/sqrt(0.1 + g^2) with g = sum_b (0.5 - y_b) * x_b (SUM loss)."""
    m = owd.WideDeep(model_type="wide")
    dense, cat, labels = _batch(5, seed=3)
    assert torch.equal(m.predict(dense, cat)["prob"], torch.full((5,), 0.5))
    loss = m.train_step(dense, cat, labels)
    assert abs(loss - 5 * math.log(2.0)) < 1e-5
    dy = 0.5 - labels
    g_bias = float(dy.sum())
    lr = m.linear_lr
    want = -g_bias * lr / math.sqrt(0.1 + g_bias * g_bias)
    assert abs(float(m.params["linear/linear_model/bias_weights"]) - want) < 1e-6
    j = 4
    g = float((dy * dense[:, j]).sum())
    assert abs(float(m.params["linear/linear_model/I5/weights"]) - (-g * lr / math.sqrt(0.1 + g * g))) < 1e-6
    # a categorical weight that one example hit: g = dy_b
    idx = int(cat[2, 7])
    w = float(m.params["linear/linear_model/C21/weights"][idx])
    hits = [b for b in range(5) if int(cat[b, 7]) == idx]
    gg = float(dy[hits].sum())
    assert abs(w - (-gg * lr / math.sqrt(0.1 + gg * gg))) < 1e-6
    # untouched rows stay zero
    assert float(m.params["linear/linear_model/C21/weights"].abs().sum()) - abs(w) < 1e-6 or len(set(int(c) for c in cat[:, 7])) > 1


def test_out_of_range_ids_map_to_bucket_zero():
    m = owd.WideDeep(model_type="wide_n_deep", embedding_size=8, deep_layers="16,8", seed=1)
    dense, cat, _ = _batch(4, seed=5, oob=True)
    fixed = cat.clone()
    fixed[0, 0] = 0
    fixed[1, 3] = 0
    assert torch.equal(m.predict(dense, cat)["y"], m.predict(dense, fixed)["y"])


def test_fp64_twin_stays_close_over_steps():
    ms = [owd.WideDeep(model_type="wide_n_deep", embedding_size=8, deep_layers="32,16", seed=2, dtype=dt)
          for dt in (torch.float32, torch.float64)]
    for n in ms[0].params:
        ms[1].params[n] = ms[0].params[n].double().clone()
    for s in range(4):
        dense, cat, labels = _batch(64, seed=10 + s)
        l32 = ms[0].train_step(dense, cat, labels)
        l64 = ms[1].train_step(dense.double(), cat, labels.double())
        assert abs(l32 - l64) < 1e-4 * max(1.0, abs(l64))
    for n in ms[0].params:
        a, b = ms[0].params[n].double(), ms[1].params[n]
        assert float((a - b).abs().max()) < 2e-5 * max(1.0, float(b.abs().max())), n


def test_deep_only_adagrad_first_step_on_logits_bias():
    m = owd.WideDeep(model_type="deep", embedding_size=4, deep_layers="8", seed=4)
    dense, cat, labels = _batch(16, seed=7)
    p0 = m.predict(dense, cat)["prob"]
    g = float((p0 - labels).sum())
    m.train_step(dense, cat, labels)
    want = -0.05 * g / math.sqrt(0.1 + g * g)
    assert abs(float(m.params["dnn/logits/bias"]) - want) < 1e-6


def test_csv_reader_defaults_and_errors(tmp_path):
    from tf_repos_b200 import wide_deep_main as wm
    line1 = "1," + ",".join("%.2f" % (0.1 * i) for i in range(13)) + "," + ",".join(str(100 + i) for i in range(26))
    line2 = "0," + ",".join("" for _ in range(13)) + "," + ",".join("" for _ in range(26))     # all defaults
    p = os.path.join(tmp_path, "tr.csv")
    open(p, "w").write(line1 + "\n" + line2 + "\n")
    labels, dense, cat = wm.decode_csv_file(p)
    assert labels.tolist() == [1.0, 0.0] and dense.shape == (2, 13) and cat.shape == (2, 26)
    assert np.allclose(dense[0], [0.1 * i for i in range(13)], atol=1e-6) and not dense[1].any() and not cat[1].any()
    assert cat[0].tolist() == [100 + i for i in range(26)]
    lab, d, c = owd.parse_csv_line(line1)
    assert lab == 1.0 and c == cat[0].tolist()
    open(p, "w").write("1,2,3\n")
    try:
        wm.decode_csv_file(p)
        assert False
    except ValueError as e:
        assert "Expect 40 fields" in str(e)
    # batching: repeat before batch, last partial batch kept
    open(p, "w").write("".join(line1 + "\n" for _ in range(5)))
    sizes = [b[2].shape[0] for b in wm.input_fn([p], num_epochs=2, batch_size=4)]
    assert sizes == [4, 4, 2]


def test_cli_rejects_unknown_flag_and_lists_reference_flags():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "Model_pipeline", "wide_n_deep.py"), "--feature_size=3"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "Unknown command line flag 'feature_size'" in (r.stderr + r.stdout)
