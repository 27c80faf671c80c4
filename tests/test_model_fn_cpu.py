"""tf_repos_b200/model_fn.py (the DeepFM.py:100-221 model_fn contract) against a CPU stand-in model: reshaping of
[B,F,1] features, mode dispatch, what each mode returns."""
import numpy as np
import pytest
import torch

from tf_repos_b200.model_fn import EstimatorSpec, ModeKeys, model_fn


class _Stub:
    device = torch.device("cpu")

    def __init__(self):
        self.calls = []
        self.y = torch.zeros(8)

    def predict(self, ids, vals):
        self.calls.append(("predict", tuple(ids.shape), ids.dtype, tuple(vals.shape)))
        self.y[: ids.shape[0]] = vals.sum(1) - 1.0
        return torch.sigmoid(self.y[: ids.shape[0]])

    def train_step(self, ids, vals, labels):
        self.calls.append(("train", tuple(ids.shape), tuple(labels.shape)))
        return torch.tensor([0.5, 0.01, 0.02])


def test_modes():
    m = _Stub()
    params = {"field_size": 3, "feature_size": 10, "embedding_size": 4}
    ids = torch.tensor([[[1], [2], [3]], [[4], [5], [6]]], dtype=torch.int64)       # [B,F,1] like the input_fn
    vals = torch.tensor([[[1.0], [0.5], [0.0]], [[0.0], [0.0], [0.25]]])
    feats = {"feat_ids": ids, "feat_vals": vals}
    spec = model_fn(feats, None, ModeKeys.PREDICT, params, model=m)
    assert isinstance(spec, EstimatorSpec) and set(spec.predictions) == {"prob"} and spec.predictions["prob"].shape == (2,)
    assert spec.export_outputs["serving_default"]["prob"] is spec.predictions["prob"]
    assert m.calls[-1] == ("predict", (2, 3), torch.int64, (2, 3))
    labels = torch.tensor([1.0, 0.0])
    spec = model_fn(feats, labels, ModeKeys.EVAL, params, model=m)
    y = np.array([0.5, -0.75])
    want = float((np.maximum(y, 0) - y * labels.numpy() + np.log1p(np.exp(-np.abs(y)))).mean())
    assert abs(spec.loss - want) < 1e-6 and 0.0 <= spec.eval_metric_ops["auc"] <= 1.0
    spec = model_fn(feats, labels, ModeKeys.TRAIN, params, model=m)
    assert spec.train_op is not None and m.calls[-1][0] == "predict"            # nothing ran yet
    out = spec.train_op()
    assert m.calls[-1] == ("train", (2, 3), (2,)) and out.tolist() == pytest.approx([0.5, 0.01, 0.02])
    with pytest.raises(ValueError):
        model_fn(feats, None, ModeKeys.TRAIN, params, model=m)
    with pytest.raises(ValueError):
        model_fn(feats, labels, "export", params, model=m)
