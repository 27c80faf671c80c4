"""DIN drop-in script end to end on TFRecord input (Model_pipeline/DIN.py, tf_repos_b200/din_main.py): train / eval /
infer / export, and the inference numbers against the oracle DIN fed the same checkpoint."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_din(path, n, seed, F=11, N=5000, maxlen=6):
    from tf_repos_b200 import tfrecord as tfr
    rng = np.random.RandomState(seed)
    recs = []
    for _ in range(n):
        ex = {"y": np.float32(rng.rand() < 0.3), "z": np.float32(0.0), "feat_ids": rng.randint(1, N, F).astype(np.int64),
              "a_catids": np.int64(rng.randint(1, N)), "a_shopids": np.int64(rng.randint(1, N)),
              "a_brandids": np.int64(rng.randint(1, N)), "a_intids": rng.randint(1, N, rng.randint(0, 4)).astype(np.int64)}
        for f in ("cat", "shop", "brand", "int"):
            ln = rng.randint(0, maxlen + 1)
            ex["u_%sids" % f] = rng.randint(1, N, ln).astype(np.int64)
            ex["u_%svals" % f] = (rng.rand(ln) * 3).astype(np.float32)
        recs.append(tfr.encode_example(ex))
    tfr.write_records(path, recs)


def test_din_cli_train_eval_infer_export_and_oracle_inference(tmp_path):
    tmp = str(tmp_path)
    os.makedirs(tmp + "/data/tr"); os.makedirs(tmp + "/data/te"); os.makedirs(tmp + "/ckpt")
    _write_din(tmp + "/data/tr/part0.tfrecord", 120, 1); _write_din(tmp + "/data/tr/part1.tfrecord", 80, 2)
    _write_din(tmp + "/data/te/part0.tfrecord", 70, 3)
    common = [sys.executable, os.path.join(ROOT, "Model_pipeline", "DIN.py"), "--field_size=11", "--feature_size=5000",
              "--embedding_size=8", "--batch_size=64", "--deep_layers=16,8", "--dropout=0.9,0.9", "--log_steps=1",
              "--num_epochs=1", "--data_dir=" + tmp + "/data", "--model_dir=" + tmp + "/ckpt/m_", "--dt_dir=20260922"]

    def run(*args):
        r = subprocess.run(common + list(args), capture_output=True, text=True, timeout=280)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        return r.stdout
    out = run("--task_type=train")
    assert "skipping the final partial batch" not in out and "Loss for final step" in out
    ev = json.loads(run("--task_type=eval").strip().splitlines()[-1])
    assert ev["global_step"] == 4 and 0.0 <= ev["auc"] <= 1.0     # 200 samples / 64: three full batches + the partial one (kept, DIN.py:93-94)
    run("--task_type=infer")
    lines = open(tmp + "/data/pred.txt").read().split("\n")
    assert len(lines) == 71 and lines[-1] == ""
    run("--task_type=export", "--servable_model_dir=" + tmp + "/export")
    # the oracle DIN with the checkpoint's variables scores the test set to the same numbers
    from oracle import models as om
    from tf_repos_b200 import din_main as dm
    st = torch.load(tmp + "/ckpt/m_20260922/ctr_b200.ckpt", map_location="cpu")
    ref = om.DIN(11, 5000, 8, deep_layers="16,8", dropout="0.9,0.9", attention_layers="256", seed=0)
    for k, v in st["variables"].items():
        ref.params[k] = v.float().reshape(ref.params[k].shape).clone()
    d = dm.decode_tfrecord_files([tmp + "/data/te/part0.tfrecord"], 11)
    meta = json.load(open(tmp + "/ckpt/m_20260922/din_shapes.json"))
    want = []
    for idx in dm.index_stream(70, 1, 64):
        batch, _, n = dm.make_batch(d, idx, 64, meta["P"], "cpu")
        lb = {k: (v.long() if v.dtype == torch.int32 else v) for k, v in batch.items()}
        want.append(ref.predict(lb)["prob"][:n].numpy())
    want = np.concatenate(want)
    got = np.array([float(l) for l in lines[:-1]], dtype=np.float32)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6)
