"""CPU tests of the drop-in surface: libsvm parser / input_fn vs the oracle restatement, flag parsing,
AUC, quirks.  (No GPU: the native parser is host code in libctr_b200.so.)"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _toy(tmp_path, rows=103, F=39, N=10_000, seed=0, name="tr.libsvm"):
    from tf_repos_b200 import synth
    ids, vals, labels = synth.criteo_batch(rows, N, F, seed=seed)
    p = str(tmp_path / name)
    synth.write_libsvm(p, ids, vals, labels)
    return p, ids, vals, labels


def test_native_parser_matches_oracle_parser(tmp_path):
    from oracle import libsvm as ol
    from tf_repos_b200 import input_fn as inp
    p, ids, vals, labels = _toy(tmp_path)
    i2, v2, l2 = inp.decode_libsvm_file(p)
    ref = [ol.decode_libsvm(line) for line in open(p)]
    np.testing.assert_array_equal(i2, np.stack([r[0] for r in ref]))
    np.testing.assert_array_equal(v2, np.stack([r[1] for r in ref]))      # strtof == np.float32(str): bit-exact
    np.testing.assert_array_equal(l2, np.array([r[2] for r in ref], dtype=np.float32))
    np.testing.assert_array_equal(i2, ids.numpy())


def test_reference_comment_line_and_ragged_rows(tmp_path):
    """The sample line in the reference (DeepFM.py:62) has 27 pairs; runs of spaces are skipped like
    tf.string_split does; a row with a different pair count is an error."""
    from tf_repos_b200 import input_fn as inp
    line = ("1 1:0.5 2:0.03519 3:1 4:0.02567 7:0.03708 8:0.01705 9:0.06296 10:0.18185 11:0.02497 12:1 14:0.02565 "
            "15:0.03267 17:0.0247 18:0.03158 20:1 22:1 23:0.13169 24:0.02933 27:0.18159 31:0.0177 34:0.02888 38:1 "
            "51:1 63:1 132:1 164:1 236:1")
    p = tmp_path / "a.libsvm"
    p.write_text(line + "\n" + line.replace(" 2:", "   2:") + "\n\n")
    ids, vals, labels = inp.decode_libsvm_file(str(p))
    assert ids.shape == (2, 27) and ids[0, -1] == 236 and vals[0, 1] == np.float32(0.03519) and labels.tolist() == [1, 1]
    assert np.array_equal(ids[0], ids[1])
    (tmp_path / "b.libsvm").write_text(line + "\n" + "0 1:1 2:1\n")
    with pytest.raises(ValueError, match="field_size"):
        inp.decode_libsvm_file(str(tmp_path / "b.libsvm"))
    (tmp_path / "c.libsvm").write_text("1 3-0.5 4:1\n")
    with pytest.raises(ValueError):
        inp.decode_libsvm_file(str(tmp_path / "c.libsvm"), field_size=2)
    (tmp_path / "empty.libsvm").write_text("")
    e = inp.decode_libsvm_file(str(tmp_path / "empty.libsvm"), field_size=39)
    assert e[0].shape == (0, 39)


def test_input_fn_batches_like_the_reference(tmp_path):
    """repeat before batch: batches straddle epochs; last partial batch kept; shapes [B,F,1]."""
    from oracle import libsvm as ol
    from tf_repos_b200 import input_fn as inp
    p, ids, vals, labels = _toy(tmp_path, rows=103)
    p2, *_ = _toy(tmp_path, rows=50, seed=1, name="tr2.libsvm")
    got = list(inp.input_fn([p, p2], batch_size=32, num_epochs=2))
    ref = list(ol.input_fn([p, p2], batch_size=32, num_epochs=2))
    assert len(got) == len(ref) == 10 and got[-1][1].shape[0] == (2 * 153) % 32
    for (gf, gl), (rf, rl) in zip(got, ref):
        assert gf["feat_ids"].dtype == torch.int32 and gf["feat_vals"].dtype == torch.float32
        np.testing.assert_array_equal(gf["feat_ids"].numpy(), rf["feat_ids"])
        np.testing.assert_array_equal(gf["feat_vals"].numpy(), rf["feat_vals"])
        np.testing.assert_array_equal(gl.numpy(), rl)
    sh = list(inp.input_fn([p], batch_size=1000, num_epochs=1, perform_shuffle=True))
    assert sorted(sh[0][1].tolist()) == sorted(labels.tolist()) and sh[0][0]["feat_ids"].shape == (103, 39, 1)


def test_large_file_is_parsed_in_parallel_chunks(tmp_path, monkeypatch):
    from tf_repos_b200 import input_fn as inp
    p, ids, vals, labels = _toy(tmp_path, rows=5000)
    monkeypatch.setattr(inp, "CHUNK", 10_000)       # force many chunks
    i2, v2, l2 = inp.decode_libsvm_file(p)
    np.testing.assert_array_equal(i2, ids.numpy()); np.testing.assert_array_equal(l2, labels.numpy())


def test_flags_match_reference_defaults_and_syntax():
    import importlib
    from tf_repos_b200 import flags
    importlib.reload(flags)
    flags.define_common()
    F = flags.FLAGS
    # defaults of DeepFM.py:35-60
    assert (F.embedding_size, F.batch_size, F.num_epochs, F.learning_rate, F.l2_reg, F.optimizer, F.deep_layers,
            F.dropout, F.batch_norm, F.task_type, F.log_steps, F.num_threads) == \
           (32, 64, 10, 0.0005, 0.0001, "Adam", "256,128,64", "0.5,0.5,0.5", False, "train", 1000, 16)
    # the canonical invocation of deep_ctr/run.sh:13
    rest = F._parse("--task_type=train --learning_rate=0.0005 --optimizer=Adam --num_epochs=1 --batch_size=256 "
                    "--field_size=39 --feature_size=117581 --deep_layers=400,400,400 --dropout=0.5,0.5,0.5 "
                    "--log_steps=1000 --num_threads=8 --model_dir=./model_ckpt/criteo/DeepFM/ --data_dir=./data/criteo/".split())
    assert rest == [] and F.feature_size == 117581 and F.deep_layers == "400,400,400" and F.batch_size == 256
    F._parse(["--clear_existing_model", "--batch_norm=True", "--field_size", "7"])
    assert F.clear_existing_model and F.batch_norm and F.field_size == 7
    with pytest.raises(SystemExit):
        F._parse(["--num_cross_layers=6"])          # the real flag is --cross_layers (DCN.py:52)
    with pytest.raises(AttributeError):
        F.chief_hosts                                # quirk Q1: never defined (DeepFM.py:240)


def test_auc_200_matches_oracle_and_sklearn():
    from oracle import tf_semantics as tfs
    from tf_repos_b200.estimator import auc_200
    rng = np.random.default_rng(0)
    y = rng.random(20_000) < 0.25
    p = np.clip(rng.normal(0.3 + 0.15 * y, 0.2), 0, 1).astype(np.float32)
    a, b = auc_200(y, p), tfs.auc(y, p)
    assert abs(a - b) < 1e-6
    from sklearn.metrics import roc_auc_score
    assert abs(a - roc_auc_score(y, p)) < 2e-3
