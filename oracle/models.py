"""CPU restatement of the reference model_fns (deep_ctr/Model_pipeline/*.py), op for op.

TEST INFRASTRUCTURE ONLY (see oracle/tf_semantics.py header; PARITY UNPINNED -- no TensorFlow, no
reference golden vectors exist).  The backward pass is torch autograd on the restated forward
(the reference's backward is TF autodiff of the same graph); the gradient aggregation and the
optimizer apply are restated explicitly because their TF semantics are what make "updated
embedding rows" non-trivial (dense L2 gradient, non-lazy sparse Adam).

Variable names are the TF checkpoint names of the reference graph.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import tf_semantics as tfs

F32 = torch.float32


def _ints(s) -> List[int]:
    return [int(t) for t in s.split(",")] if isinstance(s, str) else list(s)


def _floats(s) -> List[float]:
    return [float(t) for t in s.split(",")] if isinstance(s, str) else list(s)


class OracleModel:
    """Shared machinery: variables, slots, TF-exact (or lazy) train step.

    Sub-classes define
      tables      : names of variables read through embedding_lookup (sparse gradients)
      l2_vars     : names of variables inside `l2_reg * tf.nn.l2_loss(.)` terms, in loss order
      sites(batch): {site: (table_name, ids LongTensor)} -- one entry per embedding_lookup call
      forward(rows, dense, batch, train, masks) -> {"y": logits, ...}
    """

    tables: Tuple[str, ...] = ()
    l2_vars: Tuple[str, ...] = ()

    def __init__(self, l2_reg=1e-4, learning_rate=5e-4, optimizer="Adam", dtype=F32, update_mode="exact"):
        self.dtype = dtype
        self.l2_reg = float(l2_reg)
        self.learning_rate = float(learning_rate)
        self.optimizer = optimizer
        self.update_mode = update_mode  # "exact" = TensorFlow; "lazy" = gathered rows only
        self.params: Dict[str, torch.Tensor] = {}
        self.slots: Dict[str, List[torch.Tensor]] = {}
        self.global_step = 0
        self.adam = tfs.AdamHyper(learning_rate, dtype=dtype) if optimizer == "Adam" else None

    # -- variables ---------------------------------------------------------------------------
    def add_param(self, name: str, value: torch.Tensor):
        self.params[name] = value.to(self.dtype).contiguous()

    def init_slots(self):
        for name, p in self.params.items():
            if self.optimizer == "Adam":
                self.slots[name] = [torch.zeros_like(p), torch.zeros_like(p)]
            elif self.optimizer == "Adagrad":
                self.slots[name] = [torch.full_like(p, 1e-8)]  # initial_accumulator_value (DeepFM.py:207)
            elif self.optimizer == "Momentum":
                self.slots[name] = [torch.zeros_like(p)]
            elif self.optimizer == "ftrl":
                self.slots[name] = [torch.full_like(p, 0.1), torch.zeros_like(p)]  # accum, linear
            else:
                raise NameError(f"optimizer {self.optimizer!r}: the reference has no branch for it "
                                "(DeepFM.py:204-211 -> NameError)")

    # -- to be provided ------------------------------------------------------------------------
    def sites(self, batch) -> Dict[str, Tuple[str, torch.Tensor]]:
        raise NotImplementedError

    def forward(self, rows, dense, batch, train: bool, masks=None) -> Dict[str, torch.Tensor]:
        raise NotImplementedError

    # -- inference / eval ------------------------------------------------------------------------
    def _gather(self, batch, grad: bool):
        rows = {}
        for site, (tname, ids) in self.sites(batch).items():
            r = tfs.embedding_lookup(self.params[tname], ids)
            rows[site] = r.detach().requires_grad_() if grad else r
        return rows

    def predict(self, batch) -> Dict[str, torch.Tensor]:
        with torch.no_grad():
            dense = {n: p for n, p in self.params.items() if n not in self.tables}
            out = self.forward(self._gather(batch, False), dense, batch, train=False)
            out["prob"] = tfs.sigmoid(out["y"])
        return out

    def reg_loss(self) -> torch.Tensor:
        t = torch.zeros((), dtype=self.dtype)
        for n in self.l2_vars:
            t = t + torch.tensor(self.l2_reg, dtype=self.dtype) * tfs.l2_loss(self.params[n])
        return t

    def evaluate(self, batch, labels) -> Dict[str, float]:
        out = self.predict(batch)
        ce = tfs.sigmoid_cross_entropy_with_logits(out["y"], labels.to(self.dtype)).mean()
        loss = ce
        for n in self.l2_vars:
            loss = loss + torch.tensor(self.l2_reg, dtype=self.dtype) * tfs.l2_loss(self.params[n])
        return {"loss": float(loss), "auc": tfs.auc(labels.numpy(), out["prob"].float().numpy())}

    # -- one optimizer.minimize(loss) ----------------------------------------------------------------
    def gradients(self, batch, labels, masks=None):
        """Returns (loss, out, table_grads {name: (summed [U,..], uniq ids)}, dense_grads {name: g})."""
        rows = self._gather(batch, True)
        dense = {n: p.detach().requires_grad_() for n, p in self.params.items() if n not in self.tables}
        out = self.forward(rows, dense, batch, train=True, masks=masks)
        ce = tfs.sigmoid_cross_entropy_with_logits(out["y"], labels.to(self.dtype)).mean()
        ce.backward()
        loss = ce.detach()
        for n in self.l2_vars:  # loss = ce + l2*l2_loss(a) + l2*l2_loss(b) ... (DeepFM.py:188-190)
            loss = loss + torch.tensor(self.l2_reg, dtype=self.dtype) * tfs.l2_loss(self.params[n])
        sites = self.sites(batch)
        table_grads = {}
        for tname in self.tables:
            vals, idx = [], []
            p = self.params[tname]
            for site, (tn, ids) in sites.items():
                if tn != tname:
                    continue
                g = rows[site].grad
                if g is None:
                    g = torch.zeros_like(rows[site])
                vals.append(g.reshape((-1,) + tuple(p.shape[1:])).numpy())
                idx.append(ids.reshape(-1).numpy())
            summed, uniq = tfs.deduplicate_indexed_slices(np.concatenate(vals), np.concatenate(idx))
            table_grads[tname] = (torch.from_numpy(summed), torch.from_numpy(uniq.astype(np.int64)))
        dense_grads = {}
        for n, p in dense.items():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            if n in self.l2_vars and self.l2_reg != 0.0:
                g = g + torch.tensor(self.l2_reg, dtype=self.dtype) * self.params[n]  # AddN of the two gradients
            dense_grads[n] = g
        out = {k: v.detach() for k, v in out.items()}
        out["per_occurrence"] = {s: (rows[s].grad.detach() if rows[s].grad is not None else None) for s in rows}
        return loss, out, table_grads, dense_grads

    def train_step(self, batch, labels, masks=None) -> float:
        loss, _, table_grads, dense_grads = self.gradients(batch, labels, masks)
        self.apply_gradients(table_grads, dense_grads)
        return float(loss)

    def apply_gradients(self, table_grads, dense_grads):
        dt = self.dtype
        l2 = torch.tensor(self.l2_reg, dtype=dt)
        if self.optimizer == "Adam":
            lr_t, b1, b2, eps = self.adam.lr_t(), self.adam.b1, self.adam.b2, self.adam.eps
        lr = torch.tensor(self.learning_rate, dtype=dt)
        for tname, (summed, uniq) in table_grads.items():
            var = self.params[tname]
            slots = self.slots[tname]
            regularised = tname in self.l2_vars and self.l2_reg != 0.0
            if self.update_mode == "lazy":
                # gathered rows only; g = segment_sum + l2*var[row]
                g = summed + (l2 * var[uniq] if regularised else 0)
                rv = var[uniq]
                rs = [s[uniq] for s in slots]
                self._apply_rows(rv, rs, g, sparse=True)
                var[uniq] = rv
                for s, r in zip(slots, rs):
                    s[uniq] = r
                continue
            # TensorFlow: the dense l2 gradient is converted to IndexedSlices over ALL rows and
            # concatenated with the gather gradients => every row is an index of the sparse apply.
            if regularised:
                if var.numel() >= (1 << 24):
                    G = tfs._scratch(var, 2)
                    torch.mul(var, l2, out=G)
                else:
                    G = l2 * var
                G[uniq] = summed + G[uniq]
                self._apply_rows(var, slots, G, sparse=True)
            elif self.optimizer == "Adam":
                # non-lazy sparse Adam: m,v decay and var moves for every row; only g is sparse
                one = torch.ones((), dtype=dt)
                m, v = slots
                m.mul_(b1)
                m[uniq] = m[uniq] + summed * (one - b1)
                v.mul_(b2)
                v[uniq] = v[uniq] + (summed * summed) * (one - b2)
                var.sub_((lr_t * m) / (tfs.ieee_sqrt(v) + eps))
            else:
                rv = var[uniq]
                rs = [s[uniq] for s in slots]
                self._apply_rows(rv, rs, summed, sparse=True)
                var[uniq] = rv
                for s, r in zip(slots, rs):
                    s[uniq] = r
        for n, g in dense_grads.items():
            self._apply_rows(self.params[n], self.slots[n], g, sparse=False)
        if self.optimizer == "Adam":
            self.adam.finish()
        self.global_step += 1

    def _apply_rows(self, var, slots, g, sparse: bool):
        dt = self.dtype
        lr = torch.tensor(self.learning_rate, dtype=dt)
        if self.optimizer == "Adam":
            a = self.adam
            (tfs.adam_sparse_ if sparse else tfs.adam_dense_)(var, slots[0], slots[1], g, a.lr_t(), a.b1, a.b2, a.eps)
        elif self.optimizer == "Adagrad":
            tfs.adagrad_(var, slots[0], g, lr)
        elif self.optimizer == "Momentum":
            tfs.momentum_(var, slots[0], g, lr, torch.tensor(0.95, dtype=dt))
        elif self.optimizer == "ftrl":
            tfs.ftrl_(var, slots[0], slots[1], g, lr)


class MLPMixin:
    """The 'Deep-part' MLP shared by every libsvm model (DeepFM.py:137-167)."""

    def build_mlp(self, in_dim: int, layers: List[int], gen, scope="Deep-part", out_scope="deep_out",
                  out_dim_in: Optional[int] = None, batch_norm=False):
        d = in_dim
        for i, h in enumerate(layers):
            self.add_param(f"{scope}/mlp{i}/weights", tfs.xavier_uniform((d, h), gen, self.dtype))
            self.add_param(f"{scope}/mlp{i}/biases", torch.zeros(h, dtype=self.dtype))
            if batch_norm:
                self.add_param(f"{scope}/bn_{i}/gamma", torch.ones(h, dtype=self.dtype))
                self.add_param(f"{scope}/bn_{i}/beta", torch.zeros(h, dtype=self.dtype))
                self.bn_state[f"{scope}/bn_{i}/moving_mean"] = torch.zeros(h, dtype=self.dtype)
                self.bn_state[f"{scope}/bn_{i}/moving_variance"] = torch.ones(h, dtype=self.dtype)
            d = h
        if out_scope:
            od = out_dim_in if out_dim_in is not None else d
            name = out_scope if "/" in out_scope else f"{scope}/{out_scope}"
            self.add_param(f"{name}/weights", tfs.xavier_uniform((od, 1), gen, self.dtype))
            self.add_param(f"{name}/biases", torch.zeros(1, dtype=self.dtype))

    def run_mlp(self, x, dense, layers, keep, train, masks, scope="Deep-part", batch_norm=False, bn_decay=0.9):
        for i in range(len(layers)):
            x = tfs.fully_connected(x, dense[f"{scope}/mlp{i}/weights"], dense[f"{scope}/mlp{i}/biases"], "relu")
            if batch_norm:  # after relu (DeepFM.py:159-160)
                x = tfs.batch_norm(x, dense[f"{scope}/bn_{i}/gamma"], dense[f"{scope}/bn_{i}/beta"],
                                   self.bn_state[f"{scope}/bn_{i}/moving_mean"],
                                   self.bn_state[f"{scope}/bn_{i}/moving_variance"], train, bn_decay)
            if train:  # DeepFM.py:161-162
                x = tfs.dropout(x, keep[i], None if masks is None else masks[i])
        return x


class DeepFM(OracleModel, MLPMixin):
    """DeepFM.py:100-221."""

    tables = ("fm_w", "fm_v")
    l2_vars = ("fm_w", "fm_v")  # DeepFM.py:188-190

    def __init__(self, field_size, feature_size, embedding_size, deep_layers="256,128,64",
                 dropout="0.5,0.5,0.5", batch_norm=False, batch_norm_decay=0.9, seed=0, **kw):
        super().__init__(**kw)
        self.F, self.N, self.K = field_size, feature_size, embedding_size
        self.layers, self.keep = _ints(deep_layers), _floats(dropout)
        self.batch_norm, self.bn_decay = batch_norm, batch_norm_decay
        self.bn_state = {}
        gen = torch.Generator().manual_seed(seed)
        self.add_param("fm_bias", torch.zeros(1))                                  # DeepFM.py:114
        self.add_param("fm_w", tfs.glorot_normal((self.N,), gen, self.dtype))      # :115
        self.add_param("fm_v", tfs.glorot_normal((self.N, self.K), gen, self.dtype))  # :116
        self.build_mlp(self.F * self.K, self.layers, gen, batch_norm=batch_norm)
        self.init_slots()

    def sites(self, batch):
        ids = batch["feat_ids"].reshape(-1, self.F)
        return {"w": ("fm_w", ids), "v": ("fm_v", ids)}

    def forward(self, rows, dense, batch, train, masks=None):
        B = rows["v"].shape[0]
        vals = batch["feat_vals"].reshape(-1, self.F).to(self.dtype)
        y_w = (rows["w"] * vals).sum(1)                                            # :125-127
        emb = rows["v"] * vals.reshape(-1, self.F, 1)                              # :130-132
        sum_square = emb.sum(1) ** 2
        square_sum = (emb ** 2).sum(1)
        y_v = 0.5 * (sum_square - square_sum).sum(1)                               # :133-135
        x = emb.reshape(B, self.F * self.K)                                        # :151
        h = self.run_mlp(x, dense, self.layers, self.keep, train, masks, batch_norm=self.batch_norm,
                         bn_decay=self.bn_decay)
        y_d = tfs.fully_connected(h, dense["Deep-part/deep_out/weights"], dense["Deep-part/deep_out/biases"],
                                  None).reshape(-1)                                # :165-167
        y = dense["fm_bias"] * torch.ones_like(y_d) + y_w + y_v + y_d              # :172-175
        return {"y": y, "y_w": y_w, "y_v": y_v, "y_d": y_d, "x": x, "S": emb.sum(1)}


class DCN(OracleModel, MLPMixin):
    """DCN.py:105-230 (Deep & Cross).  No first-order term, no bias variable."""

    tables = ("emb",)
    l2_vars = ("cross_b", "cross_w", "emb")  # DCN.py:198-199, in loss order

    def __init__(self, field_size, feature_size, embedding_size, deep_layers="256,128,64", cross_layers=3,
                 dropout="0.5,0.5,0.5", batch_norm=False, batch_norm_decay=0.9, seed=0, **kw):
        super().__init__(**kw)
        self.F, self.N, self.K, self.L = field_size, feature_size, embedding_size, int(cross_layers)
        self.layers, self.keep = _ints(deep_layers), _floats(dropout)
        self.batch_norm, self.bn_decay = batch_norm, batch_norm_decay
        self.bn_state = {}
        D = self.F * self.K
        gen = torch.Generator().manual_seed(seed)
        self.add_param("cross_b", tfs.glorot_normal((self.L, D), gen, self.dtype))     # DCN.py:118-119
        self.add_param("cross_w", tfs.glorot_normal((self.L, D), gen, self.dtype))     # :120-121
        self.add_param("emb", tfs.glorot_normal((self.N, self.K), gen, self.dtype))    # :122-123
        self.build_mlp(D, self.layers, gen, scope="Deep-Network", out_scope="DCN-out/out_layer",
                       out_dim_in=D + (self.layers[-1] if self.layers else D), batch_norm=batch_norm)
        self.init_slots()

    def sites(self, batch):
        return {"emb": ("emb", batch["feat_ids"].reshape(-1, self.F))}

    def forward(self, rows, dense, batch, train, masks=None):
        B = rows["emb"].shape[0]
        vals = batch["feat_vals"].reshape(-1, self.F, 1).to(self.dtype)
        x0 = (rows["emb"] * vals).reshape(B, self.F * self.K)                          # :134-138
        xl = x0
        for l in range(self.L):                                                        # :140-145
            wl = dense["cross_w"][l].reshape(-1, 1)
            xlw = xl @ wl
            xl = x0 * xlw + xl + dense["cross_b"][l]
        h = self.run_mlp(x0, dense, self.layers, self.keep, train, masks, scope="Deep-Network",
                         batch_norm=self.batch_norm, bn_decay=self.bn_decay)           # :147-176
        x_stack = torch.cat([xl, h], 1)                                                # :179
        y = tfs.fully_connected(x_stack, dense["DCN-out/out_layer/weights"],
                                dense["DCN-out/out_layer/biases"], None).reshape(-1)        # :180-183
        return {"y": y, "x0": x0, "xL": xl}


class _LinearEmb(OracleModel, MLPMixin):
    """Shared by NFM / PNN / AFM: variables `bias`, `linear`, `emb`; L2 on linear and emb."""

    tables = ("linear", "emb")
    l2_vars = ("linear", "emb")  # NFM.py:169, PNN.py:207, AFM.py:181

    def _base(self, field_size, feature_size, embedding_size, gen):
        self.F, self.N, self.K = field_size, feature_size, embedding_size
        self.bn_state = {}
        self.add_param("bias", torch.zeros(1))
        self.add_param("linear", tfs.glorot_normal((self.N,), gen, self.dtype))
        self.add_param("emb", tfs.glorot_normal((self.N, self.K), gen, self.dtype))

    def sites(self, batch):
        ids = batch["feat_ids"].reshape(-1, self.F)
        return {"w": ("linear", ids), "v": ("emb", ids)}

    def _lin_emb(self, rows, batch):
        vals = batch["feat_vals"].reshape(-1, self.F).to(self.dtype)
        y_linear = (rows["w"] * vals).sum(1)
        emb = rows["v"] * vals.reshape(-1, self.F, 1)
        return y_linear, emb


class NFM(_LinearEmb):
    """NFM.py:94-200."""

    def __init__(self, field_size, feature_size, embedding_size, deep_layers="128,64", dropout="0.5,0.8,0.8", seed=0,
                 batch_norm=False, batch_norm_decay=0.9, **kw):
        kw.setdefault("l2_reg", 0.001); kw.setdefault("learning_rate", 0.05)
        super().__init__(**kw)
        self.layers, self.keep = _ints(deep_layers), _floats(dropout)
        self.batch_norm, self.bn_decay = batch_norm, batch_norm_decay           # NFM.py:142-143
        gen = torch.Generator().manual_seed(seed)
        self._base(field_size, feature_size, embedding_size, gen)
        self.build_mlp(self.K, self.layers, gen, batch_norm=batch_norm)
        self.init_slots()

    def forward(self, rows, dense, batch, train, masks=None):
        y_linear, emb = self._lin_emb(rows, batch)
        x = 0.5 * (emb.sum(1) ** 2 - (emb ** 2).sum(1))                                   # NFM.py:126-128
        if train:
            x = tfs.dropout(x, self.keep[0], None if masks is None else masks.get("bi"))  # :136-137
        h = self.run_mlp(x, dense, self.layers, self.keep, train, None if masks is None else masks.get("mlp"),
                         batch_norm=self.batch_norm, bn_decay=self.bn_decay)
        y_d = tfs.fully_connected(h, dense["Deep-part/deep_out/weights"], dense["Deep-part/deep_out/biases"], None).reshape(-1)
        y = dense["bias"] * torch.ones_like(y_d) + y_linear + y_d                          # :152-155
        return {"y": y}


class PNN(_LinearEmb):
    """PNN.py:102-238, model_type in {FNN, Inner, Outer}."""

    def __init__(self, field_size, feature_size, embedding_size, model_type="Inner", deep_layers="256,128,64",
                 dropout="0.5,0.5,0.5", seed=0, batch_norm=False, batch_norm_decay=0.9, **kw):
        super().__init__(**kw)
        self.model_type = model_type
        self.batch_norm, self.bn_decay = batch_norm, batch_norm_decay           # PNN.py:180-181
        self.layers, self.keep = _ints(deep_layers), _floats(dropout)
        gen = torch.Generator().manual_seed(seed)
        self._base(field_size, feature_size, embedding_size, gen)
        P = field_size * (field_size - 1) // 2
        dz = field_size * embedding_size + {"FNN": 0, "Inner": P, "Outer": P * embedding_size ** 2}[model_type]
        self.build_mlp(dz, self.layers, gen, batch_norm=batch_norm)
        self.init_slots()

    def forward(self, rows, dense, batch, train, masks=None):
        y_linear, emb = self._lin_emb(rows, batch)
        B, F, K = emb.shape
        x = emb.reshape(B, F * K)
        if self.model_type == "FNN":
            z = x
        else:
            row, col = [], []
            for i in range(F - 1):                                                         # PNN.py:144-147
                for j in range(i + 1, F):
                    row.append(i); col.append(j)
            p, q = emb[:, row], emb[:, col]
            if self.model_type == "Inner":
                z = torch.cat([x, (p * q).sum(-1)], 1)                                     # :152-153
            else:
                z = torch.cat([x, torch.einsum("api,apj->apij", p, q).reshape(B, -1)], 1)  # :166-167
        h = self.run_mlp(z, dense, self.layers, self.keep, train, None if masks is None else masks.get("mlp"),
                         batch_norm=self.batch_norm, bn_decay=self.bn_decay)
        y_d = tfs.fully_connected(h, dense["Deep-part/deep_out/weights"], dense["Deep-part/deep_out/biases"], None).reshape(-1)
        return {"y": dense["bias"] * torch.ones_like(y_d) + y_linear + y_d}                # :190-193


class AFM(_LinearEmb):
    """AFM.py:99-212."""

    ATT, POOL = "Attention-part", "Attention-based-Pooling"

    def __init__(self, field_size, feature_size, embedding_size, attention_layers="256", dropout="1.0,0.5", seed=0, **kw):
        kw.setdefault("l2_reg", 1.0); kw.setdefault("learning_rate", 0.1)
        super().__init__(**kw)
        self.att_layers, self.keep = _ints(attention_layers), _floats(dropout)
        gen = torch.Generator().manual_seed(seed)
        self._base(field_size, feature_size, embedding_size, gen)
        d = self.K
        for i, a in enumerate(self.att_layers):
            self.add_param(f"{self.ATT}/mlp{i}/weights", tfs.xavier_uniform((d, a), gen, self.dtype))
            self.add_param(f"{self.ATT}/mlp{i}/biases", torch.zeros(a, dtype=self.dtype))
            d = a
        self.add_param(f"{self.ATT}/attention_out/weights", tfs.xavier_uniform((d, 1), gen, self.dtype))
        self.add_param(f"{self.ATT}/attention_out/biases", torch.zeros(1, dtype=self.dtype))
        self.add_param(f"{self.POOL}/deep_out/weights", tfs.xavier_uniform((self.K, 1), gen, self.dtype))
        self.add_param(f"{self.POOL}/deep_out/biases", torch.zeros(1, dtype=self.dtype))
        self.init_slots()

    def forward(self, rows, dense, batch, train, masks=None):
        y_linear, emb = self._lin_emb(rows, batch)
        B, F, K = emb.shape
        prods = [emb[:, i, :] * emb[:, j, :] for i in range(F) for j in range(i + 1, F)]   # AFM.py:134-136
        pw = torch.stack(prods).permute(1, 0, 2)                                           # :137-138
        P = pw.shape[1]
        h = pw.reshape(-1, K)
        for i in range(len(self.att_layers)):
            h = tfs.fully_connected(h, dense[f"{self.ATT}/mlp{i}/weights"], dense[f"{self.ATT}/mlp{i}/biases"], "relu")
        aij = tfs.fully_connected(h, dense[f"{self.ATT}/attention_out/weights"], dense[f"{self.ATT}/attention_out/biases"], None)
        soft = torch.softmax(aij.reshape(B, P, 1), dim=1)                                  # :151
        if train:
            soft = tfs.dropout(soft, self.keep[0], None if masks is None or masks.get("att") is None
                               else masks["att"].reshape(B, P, 1))                         # :152-153
        y_emb = (soft * pw).sum(1)                                                         # :156
        if train:
            y_emb = tfs.dropout(y_emb, self.keep[1], None if masks is None else masks.get("pool"))  # :157-158
        y_deep = tfs.fully_connected(y_emb, dense[f"{self.POOL}/deep_out/weights"], dense[f"{self.POOL}/deep_out/biases"],
                                     None).reshape(-1)
        return {"y": dense["bias"] * torch.ones_like(y_deep) + y_linear + y_deep}          # :164-167


class DIN(OracleModel, MLPMixin):
    """DIN.py:101-257.  batch keys: feat_ids [B,F'], a_ids [3,B], a_int_ids [nnz] + a_int_off [B+1],
    u_ids [4,B,P] / u_wgt [4,B,P] (0-padded, as sparse_tensor_to_dense produces, DIN.py:153-154)."""

    tables = ("embeddings",)
    l2_vars = ("embeddings",)  # DIN.py:226
    ATT = "Field-wise-Pooling-layer"

    def __init__(self, field_size, feature_size, embedding_size, deep_layers="256,128,64", dropout="0.5,0.5,0.5",
                 attention_layers="256", attention_pooling=True, seed=0, **kw):
        super().__init__(**kw)
        self.Fp, self.N, self.K = field_size, feature_size, embedding_size
        self.layers, self.keep = _ints(deep_layers), _floats(dropout)
        self.att_layers = _ints(attention_layers)
        self.attention_pooling = attention_pooling
        self.bn_state = {}
        K = self.K
        gen = torch.Generator().manual_seed(seed)
        self.add_param("embeddings", tfs.glorot_normal((self.N, K), gen, self.dtype))     # DIN.py:115
        if attention_pooling:
            d = 3 * K
            for i in range(len(self.att_layers)):                                        # widths = layers[i] (Q5)
                self.add_param(f"{self.ATT}/att_fc{i}/weights", tfs.xavier_uniform((d, self.layers[i]), gen, self.dtype))
                self.add_param(f"{self.ATT}/att_fc{i}/biases", torch.zeros(self.layers[i], dtype=self.dtype))
                d = self.layers[i]
            self.add_param(f"{self.ATT}/att_out/weights", tfs.xavier_uniform((d, 1), gen, self.dtype))
            self.add_param(f"{self.ATT}/att_out/biases", torch.zeros(1, dtype=self.dtype))
        self.build_mlp(self.Fp * K + 8 * K, self.layers, gen, scope="MLP-layer", out_scope="DIN-out/din_out")
        self.init_slots()

    def sites(self, batch):
        s = {"common": ("embeddings", batch["feat_ids"]),
             "a_cat": ("embeddings", batch["a_ids"][0]), "a_shop": ("embeddings", batch["a_ids"][1]),
             "a_brand": ("embeddings", batch["a_ids"][2]), "a_int": ("embeddings", batch["a_int_ids"])}
        for f, nm in enumerate(("cat", "shop", "brand", "int")):
            s[f"u_{nm}"] = ("embeddings", batch["u_ids"][f])
        return s

    def _attention_unit(self, dense_emb_raw, ids, wgt, a_emb, dense, train, masks, f):
        K = self.K
        B, P = ids.shape
        dense_emb = dense_emb_raw * wgt.unsqueeze(-1).to(self.dtype)                      # :155-156
        mask = (ids > 0).to(self.dtype).unsqueeze(-1)                                     # :157
        ub = dense_emb.reshape(-1, K)
        ax = a_emb.repeat(1, P).reshape(-1, K)                                            # :160-161
        x = torch.cat([ub, ub - ax, ax], 1)                                               # :162
        for i in range(len(self.att_layers)):
            x = tfs.fully_connected(x, dense[f"{self.ATT}/att_fc{i}/weights"], dense[f"{self.ATT}/att_fc{i}/biases"], "relu")
            if train:
                m = None if masks is None or masks.get("att") is None else masks["att"][f]
                x = tfs.dropout(x, self.keep[i], m)                                       # :167-168
        att = tfs.fully_connected(x, dense[f"{self.ATT}/att_out/weights"], dense[f"{self.ATT}/att_out/biases"], "sigmoid")
        att = att.reshape(B, P, 1)                                                        # :169-170
        return ((dense_emb * att) * mask).sum(1)                                          # :171-172

    def forward(self, rows, dense, batch, train, masks=None):
        K = self.K
        B = batch["feat_ids"].shape[0]
        common = rows["common"].reshape(B, self.Fp * K)
        a = [rows["a_cat"], rows["a_shop"], rows["a_brand"]]
        off = batch["a_int_off"].long()
        seg = torch.repeat_interleave(torch.arange(B), off[1:] - off[:-1])
        a_int = torch.zeros(B, K, dtype=self.dtype).index_add(0, seg, rows["a_int"])       # :148 (sum combiner)
        a.append(a_int)
        u = []
        for f, nm in enumerate(("cat", "shop", "brand", "int")):
            ids, wgt = batch["u_ids"][f], batch["u_wgt"][f]
            if self.attention_pooling:
                u.append(self._attention_unit(rows[f"u_{nm}"], ids, wgt, a[f], dense, train, masks, f))
            else:
                u.append((rows[f"u_{nm}"] * wgt.unsqueeze(-1).to(self.dtype)).sum(1))      # :180-183
        x = torch.cat([common] + u + a, 1)                                                 # :199
        mm = None if masks is None else masks.get("mlp")
        h = self.run_mlp(x, dense, self.layers, self.keep, train, mm, scope="MLP-layer")
        y = tfs.fully_connected(h, dense["DIN-out/din_out/weights"], dense["DIN-out/din_out/biases"], None).reshape(-1)
        return {"y": y, "x": x}
