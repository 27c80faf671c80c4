"""CPU restatement of deep_ctr/Model_pipeline/wide_n_deep.py:92-151 -- the TF canned estimators
LinearClassifier / DNNClassifier / DNNLinearCombinedClassifier over the Criteo CSV columns.

TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see oracle/tf_semantics.py).  The reference script only wires
feature columns into canned estimators, so ALL arithmetic lives in TensorFlow 1.4 (un-vendored).  What is
restated here, each marked [TF-sem] and -- as SURVEY.md A.8 says -- with lower confidence than the hand-written
model_fns because no call site in the reference shows these defaults:

  feature columns (wide_n_deep.py:92-107)
    I1..I13  numeric_column (float, CSV default 0.0)
    C14..C39 categorical_column_with_identity(num_buckets=10000, default_value=0): an id outside [0, 10000)
             becomes 0
    embedding_column(dimension=K): own table [10000, K] per column, combiner 'mean' (one id per example: the row
             itself), initializer truncated_normal(stddev = 1/sqrt(K))
  input_layer / linear_model order columns by NAME: C14(_embedding) .. C39(_embedding), I1, I10, I11, I12, I13,
             I2 .. I9
  DNN      hidden Dense(relu, glorot_uniform kernel, zero bias) x len(hidden_units) -> Dense(1); no dropout
  linear   logit = sum_cols contribution + bias; all weights start at zero
  combined logits = dnn_logits + linear_logits
  head     binary logistic head: loss = SUM over the batch of sigmoid cross-entropy (TF 1.4 canned heads reduce
           with SUM, not the mean), prediction probabilities[:, 1] = sigmoid(logits)
  optimizers
    DNNClassifier                 Adagrad(0.05,  initial_accumulator_value 0.1)
    LinearClassifier              Ftrl(min(0.2,   1/sqrt(#columns)))
    DNNLinearCombinedClassifier   Adagrad(0.001, 0.1) on dnn/*  and  Ftrl(min(0.005, 1/sqrt(#linear columns))) on linear/*
    sparse gradients are de-duplicated (summed per id) before the apply, like every tf.train.Optimizer.

Variable names follow the TF checkpoint layout of the canned estimators.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import tf_semantics as tfs

F32 = torch.float32
N_NUM, N_CAT, NUM_BUCKETS = 13, 26, 10000
NUM_NAMES = ["I%d" % i for i in range(1, 14)]
CAT_NAMES = ["C%d" % i for i in range(14, 40)]
# position of numeric column j in the name-sorted order, and the sorted list itself
NUM_SORTED = sorted(range(N_NUM), key=lambda j: NUM_NAMES[j])          # [0, 9, 10, 11, 12, 1, ..., 8]


def parse_csv_line(line: str) -> Tuple[float, List[float], List[int]]:
    """tf.decode_csv(line, record_defaults=[[0.0]] + 13*[[0.0]] + 26*[[0]]) (wide_n_deep.py:55-71) [TF-sem]:
    exactly 40 comma-separated fields; an empty field takes the default."""
    cols = line.rstrip("\r\n").split(",")
    if len(cols) != 1 + N_NUM + N_CAT:
        raise ValueError("Expect %d fields but have %d in record" % (1 + N_NUM + N_CAT, len(cols)))
    f = lambda s: float(np.float32(s)) if s.strip() != "" else 0.0
    i = lambda s: int(s) if s.strip() != "" else 0
    return f(cols[0]), [f(c) for c in cols[1:1 + N_NUM]], [i(c) for c in cols[1 + N_NUM:]]


class WideDeep:
    def __init__(self, embedding_size=32, deep_layers="256,128,64", model_type="wide_n_deep", seed=0, dtype=F32):
        assert model_type in ("wide", "deep", "wide_n_deep")
        self.K, self.model_type, self.dtype = embedding_size, model_type, dtype
        self.layers = [int(t) for t in deep_layers.split(",")] if isinstance(deep_layers, str) else list(deep_layers)
        self.has_dnn = model_type != "wide"
        self.has_linear = model_type != "deep"
        self.dnn_lr = 0.05 if model_type == "deep" else 0.001
        n_lin = N_NUM + N_CAT
        self.linear_lr = min(0.2 if model_type == "wide" else 0.005, 1.0 / math.sqrt(n_lin))
        gen = torch.Generator().manual_seed(seed)
        self.params: Dict[str, torch.Tensor] = {}
        if self.has_dnn:
            for c in CAT_NAMES:
                self.params[self.emb_name(c)] = tfs.truncated_normal((NUM_BUCKETS, self.K), 1.0 / math.sqrt(self.K), gen, dtype)
            d = N_CAT * self.K + N_NUM
            for i, w in enumerate(self.layers):
                self.params[f"dnn/hiddenlayer_{i}/kernel"] = tfs.xavier_uniform((d, w), gen, dtype)
                self.params[f"dnn/hiddenlayer_{i}/bias"] = torch.zeros(w, dtype=dtype)
                d = w
            self.params["dnn/logits/kernel"] = tfs.xavier_uniform((d, 1), gen, dtype)
            self.params["dnn/logits/bias"] = torch.zeros(1, dtype=dtype)
        if self.has_linear:
            for c in CAT_NAMES:
                self.params[f"linear/linear_model/{c}/weights"] = torch.zeros(NUM_BUCKETS, 1, dtype=dtype)
            for c in NUM_NAMES:
                self.params[f"linear/linear_model/{c}/weights"] = torch.zeros(1, 1, dtype=dtype)
            self.params["linear/linear_model/bias_weights"] = torch.zeros(1, dtype=dtype)
        self.slots: Dict[str, List[torch.Tensor]] = {}
        for n, p in self.params.items():
            if n.startswith("dnn/"):
                self.slots[n] = [torch.full_like(p, 0.1)]                       # Adagrad accumulator
            else:
                self.slots[n] = [torch.full_like(p, 0.1), torch.zeros_like(p)]  # Ftrl accum, linear
        self.global_step = 0

    @staticmethod
    def emb_name(col: str) -> str:
        return f"dnn/input_from_feature_columns/input_layer/{col}_embedding/embedding_weights"

    @staticmethod
    def clamp_ids(cat: torch.Tensor) -> torch.Tensor:
        """categorical_column_with_identity(default_value=0) [TF-sem]"""
        cat = cat.long()
        return torch.where((cat < 0) | (cat >= NUM_BUCKETS), torch.zeros_like(cat), cat)

    # ---- forward -----------------------------------------------------------------------------------------
    def _forward(self, P: Dict[str, torch.Tensor], rows: Dict[str, torch.Tensor], lin_rows: Dict[str, torch.Tensor],
                 dense: torch.Tensor) -> torch.Tensor:
        B = dense.shape[0]
        logits = torch.zeros(B, dtype=self.dtype)
        if self.has_dnn:
            x = torch.cat([rows[c] for c in CAT_NAMES] + [dense[:, j:j + 1] for j in NUM_SORTED], dim=1)
            for i in range(len(self.layers)):
                x = torch.relu(x @ P[f"dnn/hiddenlayer_{i}/kernel"] + P[f"dnn/hiddenlayer_{i}/bias"])
            logits = logits + (x @ P["dnn/logits/kernel"] + P["dnn/logits/bias"]).reshape(B)
        if self.has_linear:
            lin = torch.zeros(B, dtype=self.dtype)
            for c in CAT_NAMES:
                lin = lin + lin_rows[c].reshape(B)
            for j in NUM_SORTED:
                lin = lin + (dense[:, j:j + 1] @ P[f"linear/linear_model/{NUM_NAMES[j]}/weights"]).reshape(B)
            logits = logits + (lin + P["linear/linear_model/bias_weights"])
        return logits

    def _gather(self, cat: torch.Tensor, grad: bool):
        ids = self.clamp_ids(cat)
        rows, lin_rows = {}, {}
        for f, c in enumerate(CAT_NAMES):
            if self.has_dnn:
                r = self.params[self.emb_name(c)][ids[:, f]]
                rows[c] = r.detach().requires_grad_() if grad else r
            if self.has_linear:
                r = self.params[f"linear/linear_model/{c}/weights"][ids[:, f]]
                lin_rows[c] = r.detach().requires_grad_() if grad else r
        return ids, rows, lin_rows

    def predict(self, dense: torch.Tensor, cat: torch.Tensor) -> Dict[str, torch.Tensor]:
        with torch.no_grad():
            _, rows, lin_rows = self._gather(cat, False)
            y = self._forward(self.params, rows, lin_rows, dense.to(self.dtype))
            return {"y": y, "prob": tfs.sigmoid(y)}

    # ---- one train step -------------------------------------------------------------------------------------
    def train_step(self, dense: torch.Tensor, cat: torch.Tensor, labels: torch.Tensor) -> float:
        dense = dense.to(self.dtype)
        ids, rows, lin_rows = self._gather(cat, True)
        table_names = {self.emb_name(c) for c in CAT_NAMES} | {f"linear/linear_model/{c}/weights" for c in CAT_NAMES}
        P = {n: (p if n in table_names else p.detach().requires_grad_()) for n, p in self.params.items()}
        y = self._forward(P, rows, lin_rows, dense)
        loss = tfs.sigmoid_cross_entropy_with_logits(y, labels.to(self.dtype)).sum()     # SUM over the batch
        loss.backward()
        dt = self.dtype
        for f, c in enumerate(CAT_NAMES):
            for name, r, is_dnn in ((self.emb_name(c), rows.get(c), True),
                                    (f"linear/linear_model/{c}/weights", lin_rows.get(c), False)):
                if r is None:
                    continue
                summed, uniq = tfs.deduplicate_indexed_slices(r.grad.numpy(), ids[:, f].numpy())
                uniq_t = torch.from_numpy(uniq.astype(np.int64))
                g = torch.from_numpy(summed)
                var, slots = self.params[name], self.slots[name]
                rv, rs = var[uniq_t], [s[uniq_t] for s in slots]
                if is_dnn:
                    tfs.adagrad_(rv, rs[0], g, torch.tensor(self.dnn_lr, dtype=dt))
                else:
                    tfs.ftrl_(rv, rs[0], rs[1], g, self.linear_lr)
                var[uniq_t] = rv
                for s, r_ in zip(slots, rs):
                    s[uniq_t] = r_
        for n, p in P.items():
            if n in table_names:
                continue
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            if n.startswith("dnn/"):
                tfs.adagrad_(self.params[n], self.slots[n][0], g, torch.tensor(self.dnn_lr, dtype=dt))
            else:
                tfs.ftrl_(self.params[n], self.slots[n][0], self.slots[n][1], g, self.linear_lr)
        self.global_step += 1
        return float(loss.detach())
