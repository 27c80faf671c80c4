"""DeepFM on the B200 engine: host-side mirror of `model_fn` in deep_ctr/Model_pipeline/DeepFM.py:100-221.

Same parameters (`field_size, feature_size, embedding_size, l2_reg, learning_rate, deep_layers,
dropout`, DeepFM.py:329-338), same variable names (`fm_bias, fm_w, fm_v, Deep-part/mlp{i}/...`),
same modes (TRAIN / EVAL / PREDICT).  The sparse path (gather, FM interaction, gradient
scatter-add, optimizer) runs in hand-written sm_100a kernels through the C ABI.

update_mode
  "exact": TensorFlow semantics -- every table row moves every step (dense L2 gradient +
           non-lazy sparse Adam).  This is what `python DeepFM.py` computes.
  "exact_deferred": bit-identical state to "exact", but the update of rows nothing gathered is
           replayed lazily (csrc/epoch.cu): one pass over HBM per `epoch_steps` steps instead of one
           per step.  The l2*l2_loss terms of `loss` become available at the end of each epoch.
  "lazy" : only gathered rows are updated (what LazyAdam would do); NOT the reference's result.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops
from .engine import DenseVars, OptimizerState, SparseUpdater, Table
from .mlp import MLP


def _ints(s):
    return [int(t) for t in s.split(",")] if isinstance(s, str) else list(s)


def _floats(s):
    return [float(t) for t in s.split(",")] if isinstance(s, str) else list(s)


class DeepFM:
    def __init__(self, field_size: int, feature_size: int, embedding_size: int, batch_size: int,
                 deep_layers="256,128,64", dropout="0.5,0.5,0.5", l2_reg: float = 1e-4,
                 learning_rate: float = 5e-4, optimizer: str = "Adam", update_mode: str = "exact",
                 device="cuda", seed: int = 0, world: int = 1, epoch_steps: int = 8):
        assert update_mode in ("exact", "exact_deferred", "lazy")
        self.F, self.N, self.K, self.B = field_size, feature_size, embedding_size, batch_size
        self.layers, self.keep = _ints(deep_layers), _floats(dropout)
        self.l2_reg, self.update_mode = float(l2_reg), update_mode
        self.device = torch.device(device)
        dev = self.device
        # data parallel (world > 1): tables are replicated; every rank all-gathers the per-occurrence
        # sparse gradients and applies the identical de-duplicated update (synchronous DP replaces the
        # reference's asynchronous parameter server, DeepFM.py:237-282 -- documented deviation)
        self.world = world
        self.opt = OptimizerState(optimizer, learning_rate, l2_reg, dev)
        # ---- variables (DeepFM.py:114-116) ------------------------------------------------------
        self.fm_v = Table("fm_v", self.N, self.K, self.opt, dev, seed=seed * 2 + 1)
        self.fm_w = Table("fm_w", self.N, 1, self.opt, dev, seed=seed * 2 + 2)
        self.mlp = MLP(self.F * self.K, self.layers, self.keep, self.B, dev)
        self.dense = DenseVars([("fm_bias", (1,))] + self.mlp.specs(), self.opt, dev)
        self.mlp.init(self.dense, torch.Generator().manual_seed(seed))
        # ---- step buffers -------------------------------------------------------------------------
        B, F, K = self.B, self.F, self.K
        f32 = dict(dtype=torch.float32, device=dev)
        self.x = torch.empty(B, F * K, **f32)
        self.S = torch.empty(B, K, **f32)
        self.y_w = torch.empty(B, **f32)
        self.y_v = torch.empty(B, **f32)
        self.y = torch.empty(B, **f32)
        self.pred = torch.empty(B, **f32)
        self.dy = torch.empty(B, **f32)
        self.loss_ce = self.dense.tail[0:1]
        self.g_rows = torch.empty(B * F, K, **f32)
        self.g_w = torch.empty(B * F, **f32)
        self.oob = torch.zeros(2, dtype=torch.int32, device=dev)
        self.d_last = torch.empty(B, self.mlp.out_in, **f32)
        G = world
        self.updater = SparseUpdater(G * B * F, self.N, K, self.opt, dev, with_scalar_table=True)
        if G > 1:
            self.ids_all = torch.empty(G * B * F, dtype=torch.int32, device=dev)
            self.g_rows_all = torch.empty(G * B * F, K, **f32)
            self.g_w_all = torch.empty(G * B * F, **f32)
        self.global_step = 0
        self.epoch_steps, self.epoch_pos = epoch_steps, 0
        if update_mode == "exact_deferred":
            # Adagrad/Momentum/Ftrl with l2_reg == 0 are truly sparse in TF: nothing to defer
            if self.l2_reg == 0.0 and optimizer != "Adam":
                self.update_mode = "exact"
            else:
                self.updater.enable_epochs(epoch_steps, [self.fm_v, self.fm_w])

    # ---- variable access by TF name ------------------------------------------------------------------
    def flush(self):
        """exact_deferred: bring every row to the current step (no-op otherwise)."""
        if self.update_mode == "exact_deferred" and self.epoch_pos > 0:
            self.updater.epoch_sweep([self.fm_v, self.fm_w], self.epoch_pos, reset=False, l2_reg=self.l2_reg)

    def set_update_mode(self, mode: str):
        """Switch between exact / exact_deferred / lazy on a live model (state stays consistent)."""
        assert mode in ("exact", "exact_deferred", "lazy")
        if self.update_mode == "exact_deferred" and self.epoch_pos > 0:
            self.updater.epoch_sweep([self.fm_v, self.fm_w], self.epoch_pos, reset=True, l2_reg=self.l2_reg)
            self.epoch_pos = 0
        if mode == "exact_deferred" and not hasattr(self.updater, "ep"):
            self.updater.enable_epochs(self.epoch_steps, [self.fm_v, self.fm_w])
        self.update_mode = mode

    def variables(self) -> Dict[str, torch.Tensor]:
        self.flush()
        out = {"fm_v": self.fm_v.var, "fm_w": self.fm_w.var}
        out.update(self.dense.views)
        return out

    def load_variables(self, values: Dict[str, torch.Tensor]):
        for name, v in values.items():
            self.variables()[name].copy_(v.to(self.device, torch.float32).reshape(self.variables()[name].shape))

    # ---- f(x) ------------------------------------------------------------------------------------------
    def _forward(self, ids, vals, train: bool, masks=None):
        B = ids.shape[0]
        ops.fm_embed_fwd(ids, vals, self.fm_v.var, self.fm_w.var, ops.FM_DEEPFM, x=self.x[:B], y_w=self.y_w[:B],
                         y2=self.y_v[:B], S=self.S[:B], oob=self.oob)
        a = self.mlp.forward_hidden(self.x[:B], self.dense, train, masks, step_dev=self.opt.state[3:4])
        y_d = self.mlp.forward_out(a, self.dense)
        return a, y_d

    def predict(self, ids: torch.Tensor, vals: torch.Tensor) -> torch.Tensor:
        """mode == PREDICT (DeepFM.py:178-185): returns prob [B]."""
        B = ids.shape[0]
        self.flush()
        _, y_d = self._forward(ids, vals, train=False)
        ops.logit_loss(self.dense["fm_bias"], self.y_w[:B], self.y_v[:B], y_d, None, B, y=self.y[:B],
                       pred=self.pred[:B])
        return self.pred[:B]

    def check_ids(self):
        """TF raises InvalidArgumentError for ids outside [0, feature_size); we count them on device."""
        cnt, first = self.oob.tolist()
        if cnt:
            self.oob.zero_()
            raise IndexError(f"{cnt} feature ids outside [0, {self.N}) (first: {first}); "
                             "TensorFlow would raise InvalidArgumentError")

    def train_step(self, ids: torch.Tensor, vals: torch.Tensor, labels: torch.Tensor, masks=None) -> torch.Tensor:
        """mode == TRAIN: one optimizer.minimize(loss) (DeepFM.py:188-213).  Returns a device tensor
        [3] = {mean CE, l2*l2_loss(fm_w), l2*l2_loss(fm_v)} whose left-to-right sum is `loss`
        (the L2 terms are produced by the dense sweep in exact mode; zeros in lazy mode)."""
        B, F, K = ids.shape[0], self.F, self.K
        assert B == self.B, "train_step is specialised for the configured batch size"
        deferred = self.update_mode == "exact_deferred"
        if deferred:
            j = self.epoch_pos
            if j == 0:
                self.updater.epoch_begin()
            self.opt.tick_epoch(j)
            ids_u = ids.reshape(-1)
            if self.world > 1:
                import torch.distributed as dist
                dist.all_gather_into_tensor(self.ids_all, ids_u)
                ids_u = self.ids_all
            # gathered rows (of every rank) must hold the state at the start of this step
            self.updater.unique(ids_u)
            self.updater.epoch_rows([(self.fm_v, None), (self.fm_w, None)], j, apply=False)
        else:
            self.opt.tick()
        a, y_d = self._forward(ids, vals, train=True, masks=masks)
        ops.logit_loss(self.dense["fm_bias"], self.y_w, self.y_v, y_d, labels, B, y=self.y, pred=self.pred,
                       loss_ce=self.loss_ce, dy=self.dy, dbias=self.dense.grads["fm_bias"], B_total=B * self.world)
        self.mlp.backward_out(a, self.dy, self.dense, self.d_last)
        dX = self.mlp.backward_hidden(self.x, self.d_last, self.dense)
        ops.fm_embed_bwd(vals, self.x, self.S, dX, self.dy, self.dy, K, ops.FM_DEEPFM, self.g_rows, self.g_w)
        g_rows, g_w = self.g_rows, self.g_w
        if self.world > 1:
            import torch.distributed as dist
            if not deferred:
                dist.all_gather_into_tensor(self.ids_all, ids.reshape(-1))
            dist.all_gather_into_tensor(self.g_rows_all, self.g_rows)
            dist.all_gather_into_tensor(self.g_w_all, self.g_w)
            dist.all_reduce(self.dense.grad)  # dense gradients + the loss tail, summed over ranks
            g_rows, g_w = self.g_rows_all, self.g_w_all
        if deferred:
            self.updater.segment_sum(g_rows, g_w)
            self.updater.epoch_rows([(self.fm_v, self.updater.g_uniq), (self.fm_w, self.updater.gw_uniq)],
                                    self.epoch_pos, apply=True)
            self.epoch_pos += 1
            if self.epoch_pos == self.epoch_steps:
                self.updater.epoch_sweep([self.fm_v, self.fm_w], self.epoch_steps, reset=True, l2_reg=self.l2_reg)
                self.epoch_pos = 0
        else:
            self.updater.dedup(self.ids_all if self.world > 1 else ids.reshape(-1), g_rows, g_w)
            self.updater.apply(self.fm_v, self.fm_w, exact=(self.update_mode == "exact"), l2_reg=self.l2_reg)
        self.dense.apply()
        self.global_step += 1
        return torch.cat([self.loss_ce, self.updater.reg[1:2], self.updater.reg[0:1]])

    def epoch_reg_terms(self) -> torch.Tensor:
        """exact_deferred: [2, epoch_steps] = l2*l2_loss(fm_w), l2*l2_loss(fm_v) for every step of the
        epoch that just ended (valid right after the step that closed the epoch)."""
        ep = self.updater.ep
        return torch.stack([ep["fm_w"]["reg"][: self.epoch_steps], ep["fm_v"]["reg"][: self.epoch_steps]])

    def loss_value(self, parts: torch.Tensor) -> float:
        p = parts.tolist()
        return (p[0] + p[1]) + p[2]
