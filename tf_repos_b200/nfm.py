"""Neural Factorization Machine: mirror of `model_fn` in deep_ctr/Model_pipeline/NFM.py:94-200.
Variables: `bias [1]`, `linear [N]`, `emb [N,K]` (NFM.py:106-108), `Deep-part/mlp{i}`, `Deep-part/deep_out`.
y = bias + sum_f linear[id]*val + MLP(dropout(0.5*((sum e)^2 - sum e^2)))   (NFM.py:118-155).
dropout[0] is applied to the bi-interaction vector AND again after layer 0 (NFM.py:136-137,144-145)."""
from __future__ import annotations

import torch

from . import ops
from .base import CTRModel, floats, ints
from .engine import DenseVars
from .mlp import MLP


class NFM(CTRModel):
    table_name, linear_name, bias_name = "emb", "linear", "bias"

    def __init__(self, field_size, feature_size, embedding_size, batch_size, deep_layers="128,64",
                 dropout="0.5,0.8,0.8", l2_reg=0.001, learning_rate=0.05, optimizer="Adam", update_mode="exact",
                 device="cuda", seed=0, world=1, epoch_steps=8, batch_norm=False, batch_norm_decay=0.9):
        self.layers, self.keep = ints(deep_layers), floats(dropout)
        self.batch_norm, self.bn_decay = bool(batch_norm), float(batch_norm_decay)
        super().__init__(field_size, feature_size, embedding_size, batch_size, l2_reg, learning_rate, optimizer,
                         update_mode, device, seed, world, epoch_steps)

    def _build(self):
        B, F, K, dev = self.B, self.F, self.K, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.mlp = MLP(K, self.layers, self.keep, B, dev, seed=self.seed, batch_norm=self.batch_norm, bn_decay=self.bn_decay)
        self.dense = DenseVars([("bias", (1,))] + self.mlp.specs(), self.opt, dev)
        self.mlp.init(self.dense, torch.Generator().manual_seed(self.seed))
        self.x = torch.empty(B, F * K, **f32)
        self.S = torch.empty(B, K, **f32)
        self.y_w = torch.empty(B, **f32)
        self.bi = torch.empty(B, K, **f32)
        self.bi_d = torch.empty(B, K, **f32)
        self.bi_mask = torch.empty(B, K, **f32)
        self.d_bi = torch.empty(B, K, **f32)
        self.d_last = torch.empty(B, self.mlp.out_in, **f32)
        self._bi_active = None

    def _forward(self, ids, vals, train, masks=None):
        B = ids.shape[0]
        ops.fm_embed_fwd(ids, vals, self.V.var, self.W.var, ops.FM_NFM, x=self.x[:B], y_w=self.y_w[:B], y2=self.bi[:B],
                         S=self.S[:B], oob=self.oob)                                         # NFM.py:118-128
        inp = self.bi[:B]
        self._bi_active = None
        if train:                                                                            # NFM.py:136-137
            m = None
            if masks is not None and masks.get("bi") is not None:
                m = masks["bi"]
            elif self.keep[0] < 1.0:
                m = self.bi_mask[:B]
                ops.dropout_mask(m, self.keep[0], self.seed * 131 + 77, self.opt.state[3:4])
            if m is not None:
                ops.dropout_apply(self.bi[:B], m, self.keep[0], self.bi_d[:B])
                inp, self._bi_active = self.bi_d[:B], m
        self._inp = inp
        mm = masks.get("mlp") if masks else None
        self._a = self.mlp.forward_hidden(inp, self.dense, train, mm, step_dev=self.opt.state[3:4])
        y_d = self.mlp.forward_out(self._a, self.dense)
        return self.dense["bias"], self.y_w[:B], y_d, None                                   # NFM.py:152-155

    def _backward(self, ids, vals):
        B = ids.shape[0]
        dy = self.dy[:B]
        self.mlp.backward_out(self._a, dy, self.dense, self.d_last[:B])
        d_in = self.mlp.backward_hidden(self._inp, self.d_last[:B], self.dense)
        if self._bi_active is not None:
            ops.dropout_apply(d_in, self._bi_active, self.keep[0], self.d_bi[:B])
            d_in = self.d_bi[:B]
        ops.fm_embed_bwd(vals, self.x[:B], self.S[:B], None, d_in, dy, self.K, ops.FM_NFM,
                         self.g_rows[: B * self.F], self.g_w[: B * self.F])
