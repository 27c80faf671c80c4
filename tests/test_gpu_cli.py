"""Config 1 of BASELINE.json end to end: a 1k-row libsvm toy (39 fields, 10k vocab, k=8) through the
drop-in surface (input_fn -> model -> pred.txt), against the oracle fed by its own parser."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write(tmp, name, rows, seed, N=10_000):
    from tf_repos_b200 import synth
    ids, vals, labels = synth.criteo_batch(rows, N, 39, seed=seed)
    synth.write_libsvm(os.path.join(tmp, name), ids, vals, labels)


def test_train_on_libsvm_with_partial_last_batch_matches_oracle(tmp_path):
    from oracle import libsvm as ol
    from oracle import models as om
    from tf_repos_b200.deepfm import DeepFM
    from tf_repos_b200.input_fn import input_fn
    tmp = str(tmp_path)
    _write(tmp, "tr.libsvm", 1000, 1); _write(tmp, "te.libsvm", 100, 2)
    F, N, K, B = 39, 10_000, 8, 256
    ref = om.DeepFM(F, N, K, deep_layers="32,16", dropout="1.0,1.0", seed=2)
    g = torch.Generator().manual_seed(0)
    ref.params["fm_v"].copy_(torch.randn(N, K, generator=g) * 0.1)
    ref.params["fm_w"].copy_(torch.randn(N, generator=g) * 0.1)
    gpu = DeepFM(F, N, K, B, deep_layers="32,16", dropout="1.0,1.0", update_mode="exact_deferred", epoch_steps=3,
                 device="cuda:0")
    gpu.load_variables(ref.params)
    n_batches = 0
    for (gf, gl), (rf, rl) in zip(input_fn([tmp + "/tr.libsvm"], batch_size=B), ol.input_fn([tmp + "/tr.libsvm"], batch_size=B)):
        gpu.train_step(gf["feat_ids"].reshape(-1, F).cuda(), gf["feat_vals"].reshape(-1, F).cuda(), gl.cuda())
        ref.train_step({"feat_ids": torch.from_numpy(rf["feat_ids"]).long().reshape(-1, F),
                        "feat_vals": torch.from_numpy(rf["feat_vals"]).reshape(-1, F)}, torch.from_numpy(rl))
        n_batches += 1
    assert n_batches == 4 and gl.shape[0] == 1000 - 3 * 256
    gpu.check_ids()
    for name in ("fm_v", "fm_w", "Deep-part/mlp0/weights"):
        a, b = gpu.variables()[name].cpu(), ref.params[name]
        assert (a - b).abs().max().item() <= 2e-5 * b.abs().max().item(), name
    (tf, tl) = next(ol.input_fn([tmp + "/te.libsvm"], batch_size=100))
    batch = {"feat_ids": torch.from_numpy(tf["feat_ids"]).long().reshape(-1, F), "feat_vals": torch.from_numpy(tf["feat_vals"]).reshape(-1, F)}
    p_ref = ref.predict(batch)["prob"].numpy()
    p = gpu.predict(batch["feat_ids"].int().cuda(), batch["feat_vals"].cuda()).cpu().numpy()
    np.testing.assert_allclose(p, p_ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("script,extra", [("DeepFM.py", []), ("DCN.py", ["--cross_layers=2"])])
def test_cli_train_eval_infer_export(tmp_path, script, extra):
    tmp = str(tmp_path)
    os.makedirs(tmp + "/data"); os.makedirs(tmp + "/ckpt")
    for name, rows, seed in (("tr0.libsvm", 600, 1), ("tr1.libsvm", 400, 2), ("va.libsvm", 200, 3), ("te.libsvm", 150, 4)):
        _write(tmp + "/data", name, rows, seed)
    common = [sys.executable, os.path.join(ROOT, "Model_pipeline", script), "--field_size=39", "--feature_size=10000",
              "--embedding_size=8", "--batch_size=128", "--deep_layers=32,16", "--dropout=0.8,0.8", "--log_steps=5",
              "--num_epochs=1", "--data_dir=" + tmp + "/data", "--model_dir=" + tmp + "/ckpt/m_", "--dt_dir=20260922"] + extra

    def run(*args):
        r = subprocess.run(common + list(args), capture_output=True, text=True, timeout=280)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        return r.stdout
    out = run("--task_type=train")
    assert "global_step/sec" in out and "Loss for final step" in out
    assert os.path.exists(tmp + "/ckpt/m_20260922/ctr_b200.ckpt"), "model_dir + dt_dir (quirk Q2)"
    ev = json.loads(run("--task_type=eval").strip().splitlines()[-1])
    assert 0.0 <= ev["auc"] <= 1.0 and ev["global_step"] == 8
    out2 = run("--task_type=train")                 # resumes from the checkpoint like tf.estimator
    assert "restored checkpoint" in out2
    assert json.loads(run("--task_type=eval").strip().splitlines()[-1])["global_step"] == 16
    run("--task_type=infer")
    lines = open(tmp + "/data/pred.txt").read().split("\n")
    assert len(lines) == 151 and lines[-1] == "" and all(len(l.split(".")[1]) == 6 and 0 <= float(l) <= 1 for l in lines[:-1])
    run("--task_type=export", "--servable_model_dir=" + tmp + "/export")
    sub = os.listdir(tmp + "/export")
    sig = json.load(open(os.path.join(tmp, "export", sub[0], "signature.json")))
    assert sig["inputs"]["feat_ids"] == {"dtype": "int64", "shape": [None, 39]} and "prob" in sig["outputs"]
    # the exported model answers the serving clients' request shape (int64 ids) with pred.txt's numbers (8f-4)
    from tf_repos_b200.input_fn import decode_libsvm_file
    from tf_repos_b200.serving import Servable
    ids, vals, _ = decode_libsvm_file(tmp + "/data/te.libsvm", 39)
    s = Servable.load(os.path.join(tmp, "export", sub[0]), max_batch=64)
    prob = s.predict(ids.astype(np.int64), vals)                      # 150 rows through a 64-row servable
    want = np.array([float(l) for l in lines[:-1]], dtype=np.float32)
    np.testing.assert_allclose(prob.numpy(), want, atol=1e-6)
    one = s.predict(torch.from_numpy(ids[:1].astype(np.int64)).reshape(1, 39, 1), vals[:1])   # feat_ids[1,39] request
    assert abs(float(one[0]) - want[0]) <= 1e-6


@pytest.mark.parametrize("optimizer", ["Adam", "Adagrad", "ftrl"])
def test_tf_named_training_state_roundtrip(tmp_path, optimizer):
    """tf_names: variables + optimizer slots + beta powers + global_step under their tf.train.Saver names; a model
    restored from the archive continues bit-identically."""
    from tf_repos_b200 import synth, tf_names
    from tf_repos_b200.deepfm import DeepFM
    F, N, K, B = 39, 5000, 8, 64
    mk = lambda: DeepFM(F, N, K, B, deep_layers="16,8", dropout="1.0,1.0", optimizer=optimizer, update_mode="exact_deferred",
                        epoch_steps=4, device="cuda:0", seed=5)
    a = mk()
    batches = [synth.criteo_batch(B, N, F, seed=40 + i, device="cuda:0") for i in range(6)]
    for b in batches[:3]:
        a.train_step(*b)
    sd = tf_names.state_dict_tf(a)
    slot = {"Adam": "Adam", "Adagrad": "Adagrad", "ftrl": "Ftrl"}[optimizer]
    assert sd["fm_v"].shape == (N, K) and sd["fm_v/" + slot].shape == (N, K) and sd["Deep-part/mlp0/weights/" + slot].shape == (F * K, 16)
    assert int(sd["global_step"]) == 3 and (("beta1_power" in sd) == (optimizer == "Adam"))
    path = os.path.join(str(tmp_path), "state.npz")
    tf_names.export_npz(a, path)
    b2 = mk()
    tf_names.import_npz(b2, path)
    for b in batches[3:]:
        a.train_step(*b); b2.train_step(*b)
    va, vb = a.variables(), b2.variables()
    for name in va:
        assert torch.equal(va[name], vb[name]), name
