"""Generates tests/golden/*.json from the CPU oracle (oracle/), NOT from TensorFlow -- the reference
cannot run here (Python-2 / TF-1.4, no TensorFlow installed) and ships no golden vectors, so these
fixtures only guard the restatement (and the CUDA path) against drift.  Re-generate with
    python -m tests.golden.make_golden
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import models as om  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
# Serving_pipeline/deep_fm_serving_client.cpp:42-45 (feature_size 117581, field_size 39: run.sh:13)
SERVING_IDS = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 555, 1078, 17797, 26190, 26341, 28570, 35361,
               35613, 35984, 48424, 51364, 64053, 65964, 66206, 71628, 84088, 84119, 86889, 88280, 88283,
               100288, 100300, 102447, 109932, 111823]
SERVING_VALS = [0.05, 0.006633, 0.05, 0, 0.021594, 0.008, 0.15, 0.04, 0.362, 0.1, 0.2, 0, 0.04] + [1.0] * 26
ROW_IDS = [0, 1, 13, 15, 555, 111823, 117580]


def build_serving_case():
    """DeepFM with the reference's Criteo shape (run.sh:13: feature_size=117581, field_size=39,
    embedding_size=32, deep_layers=400,400,400 shortened to 32,16), 4-row batch whose first row is the
    serving-client sample."""
    F, N, K = 39, 117581, 32
    m = om.DeepFM(F, N, K, deep_layers="32,16", dropout="1.0,1.0", l2_reg=1e-4, learning_rate=5e-4,
                  optimizer="Adam", seed=20260922)
    g = torch.Generator().manual_seed(7)
    m.params["fm_v"].copy_(torch.randn(N, K, generator=g) * 0.05)
    m.params["fm_w"].copy_(torch.randn(N, generator=g) * 0.05)
    ids = torch.tensor([SERVING_IDS] * 4)
    ids[1, 13:] = torch.randint(14, N, (26,), generator=g)
    ids[2, 13:] = torch.randint(14, N, (26,), generator=g)
    ids[3, 13:] = ids[1, 13:]
    vals = torch.tensor([SERVING_VALS] * 4, dtype=torch.float32)
    vals[1:, :13] = torch.rand(3, 13, generator=g)
    labels = torch.tensor([1.0, 0.0, 0.0, 1.0])
    return m, {"feat_ids": ids, "feat_vals": vals}, labels


def main():
    m, batch, labels = build_serving_case()
    out = m.predict(batch)
    gold = {"serving_ids": SERVING_IDS, "serving_vals": SERVING_VALS,
            "logits": [float(v) for v in out["y"]], "prob": [float(v) for v in out["prob"]]}
    gold["losses"] = [m.train_step(batch, labels) for _ in range(3)]
    gold["row_ids"] = ROW_IDS
    gold["fm_v_rows_after_3_steps"] = m.params["fm_v"][torch.tensor(ROW_IDS)].tolist()
    gold["fm_w_after_3_steps"] = m.params["fm_w"][torch.tensor(ROW_IDS)].tolist()
    with open(os.path.join(HERE, "deepfm_serving_sample.json"), "w") as fo:
        json.dump(gold, fo, indent=1)
    print("wrote deepfm_serving_sample.json:", gold["logits"], gold["losses"])


if __name__ == "__main__":
    main()
