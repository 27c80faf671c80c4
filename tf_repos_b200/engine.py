"""Device-side building blocks shared by every model: embedding tables with TF-semantics
optimizer state, the de-duplicated sparse update, the flat buffer of dense variables and the
hyper-parameter records the kernels read.

Everything numerical happens in libctr_b200.so (tf_repos_b200/csrc); torch provides memory.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops

# hyper record layout (see include/ctr_b200.h): {lr_t, beta1, beta2, eps, l2_reg, aux0, aux1, aux2}
HYPER_TABLE, HYPER_DENSE, HYPER_DENSE_L2 = 0, 1, 2


class OptimizerState:
    """Host mirror of tf.train.*Optimizer construction (DeepFM.py:204-211) + device hyper records.

    record 0: embedding tables (l2_reg as given)      -- sparse rows / dense sweep
    record 1: dense variables without L2 (MLP; the fully_connected regularizer is a no-op, A.3)
    record 2: dense variables with L2 (DCN cross_w / cross_b)
    """

    def __init__(self, optimizer: str, learning_rate: float, l2_reg: float, device, adagrad_init: float = 1e-8):
        self.adagrad_init = adagrad_init   # initial_accumulator_value: 1e-8 in DeepFM.py:207, 0.1 in the canned estimators
        if optimizer not in ops.OPT_BY_NAME:
            # the reference has no branch for e.g. 'GD' although the flag help lists it (DeepFM.py:50,204-211)
            raise NameError(f"optimizer {optimizer!r} is not one of {sorted(ops.OPT_BY_NAME)}")
        self.name = optimizer
        self.opt = ops.OPT_BY_NAME[optimizer]
        self.n_slots = ops.OPT_SLOTS[self.opt]
        self.device = device
        aux = (0.0, 0.0, 0.0)
        beta1, beta2, eps = 0.9, 0.999, 1e-8
        if optimizer == "Momentum":
            aux = (0.95, 0.0, 0.0)
        elif optimizer == "ftrl":
            aux = (-0.5, 0.0, 0.0)  # learning_rate_power, l1, l2 (FtrlOptimizer defaults)
        rec = lambda l2: [learning_rate, beta1, beta2, eps, l2, *aux]
        self.hyper = torch.tensor([rec(l2_reg), rec(0.0), rec(l2_reg)], dtype=torch.float32, device=device)
        # {beta1_power, beta2_power, lr, global_step}; Adam's powers start at beta (TF _create_slots)
        self.state = torch.tensor([beta1, beta2, learning_rate, 0.0], dtype=torch.float32, device=device)

    def slot_init(self, slot: int) -> float:
        if self.name == "Adagrad":
            return self.adagrad_init
        if self.name == "ftrl" and slot == 0:
            return 0.1           # FtrlOptimizer initial_accumulator_value default
        return 0.0

    def tick(self):
        """Start of a step.  For Adam: lr_t from the current beta powers, then advance them."""
        if self.name == "Adam":
            ops.adam_tick(self.state, self.hyper)
        else:  # only the global-step counter advances (it seeds the dropout masks)
            self.tick_epoch(0)

    def tick_epoch(self, j: int):
        """exact-deferred mode: like tick(), and records this step's lr_t in lr_table[j]."""
        if self.lr_table is None:
            self.lr_table = torch.zeros(ops.epoch_max_steps(), dtype=torch.float32, device=self.device)
        ops.epoch_tick(self.state, self.hyper, self.lr_table, j, self.name == "Adam")

    lr_table = None

    def record(self, which: int) -> torch.Tensor:
        return self.hyper[which]


class Table:
    """One embedding variable [N, K] (K == 1 for the first-order weights `fm_w` [N]) + its slots."""

    def __init__(self, name: str, N: int, K: int, opt: OptimizerState, device, init_std: Optional[float] = None,
                 seed: int = 0, value: Optional[torch.Tensor] = None):
        self.name, self.N, self.K = name, N, K
        shape = (N,) if K == 1 else (N, K)
        self.var = torch.empty(shape, dtype=torch.float32, device=device)
        if value is not None:
            self.var.copy_(value.reshape(shape))
        else:
            if init_std is None:  # glorot_normal_initializer (DeepFM.py:115-116): sqrt(2/(fan_in+fan_out))
                init_std = math.sqrt(2.0 / (N + N)) if K == 1 else math.sqrt(2.0 / (N + K))
            ops.init_trunc_normal(self.var, init_std, seed)
        self.slots: List[torch.Tensor] = []
        for s in range(opt.n_slots):
            t = torch.empty(shape, dtype=torch.float32, device=device)
            ops.fill(t, opt.slot_init(s))
            self.slots.append(t)

    def slot(self, i):
        return self.slots[i] if i < len(self.slots) else None


class SparseUpdater:
    """K3 + K4 for a group of tables that are all gathered with the SAME ids (DeepFM: fm_v and fm_w).

    exact mode (TensorFlow semantics): gathered rows are computed first into a stage
    buffer from the pre-step state, the dense sweep then advances every row with g = l2*var, and
    the staged rows are patched back.  lazy mode updates the gathered rows only.
    """

    def __init__(self, n_ids: int, N: int, K: int, opt: OptimizerState, device, with_scalar_table: bool):
        self.n, self.N, self.K, self.opt = n_ids, N, K, opt
        self.uw = ops.UniqueWorkspace(n_ids, N, device)
        f32 = dict(dtype=torch.float32, device=device)
        self.g_uniq = torch.empty(max(n_ids, 1) * K, **f32)
        self.gw_uniq = torch.empty(max(n_ids, 1), **f32) if with_scalar_table else None
        self.stage_v = torch.empty(3 * max(n_ids, 1) * K, **f32)
        self.stage_w = torch.empty(3 * max(n_ids, 1), **f32) if with_scalar_table else None
        self.n_part = ops.sweep_partials_count()
        self.partials_v = torch.zeros(self.n_part, **f32)
        self.partials_w = torch.zeros(self.n_part, **f32)
        self.red_ws = torch.empty(1024, **f32)
        # [l2*l2_loss(V), l2*l2_loss(W)] of the PRE-step tables (what `loss` of this step contains)
        self.reg = torch.zeros(2, **f32)
        self.sweep_events = None  # set to [] to collect (start, end) CUDA events around the V sweep
        self.sweep_steps = []     # parallel to sweep_events: steps replayed by each pass (deferred mode)

    def dedup(self, ids_flat: torch.Tensor, g_rows: torch.Tensor, g_w: Optional[torch.Tensor]):
        ops.unique_segment(ids_flat, self.uw)
        ops.segment_sum_rows(g_rows, g_w, self.uw, self.K, self.g_uniq, self.gw_uniq if g_w is not None else None)

    # ---- exact-deferred ("epoch") mode: csrc/epoch.cu ------------------------------------------------
    def enable_epochs(self, P: int, tables: Sequence[Table]):
        """Allocates the per-row `last` bytes and the per-step sum(var^2) accumulators."""
        pmax = ops.epoch_max_steps()
        assert 1 <= P <= pmax
        dev = tables[0].var.device
        self.P = P
        self.flush_pos = 0   # steps of the current epoch every row's stored state already contains (mid-epoch flush)
        self.n_epart = ops.epoch_partials_count()
        self.ep = {}
        for t in tables:
            # rows gathered since the last sweep, collected by the packed Adam sweep for its second pass
            cap = max(min(self.n * pmax, t.N), 1)
            self.ep[t.name] = dict(
                list=torch.empty(cap, dtype=torch.int32, device=dev),
                list_count=torch.zeros(1, dtype=torch.int32, device=dev),
                last=torch.zeros(t.N, dtype=torch.uint8, device=dev),
                ss=torch.zeros(pmax, dtype=torch.float64, device=dev),
                partials=torch.zeros(pmax * self.n_epart, dtype=torch.float64, device=dev),
                reg=torch.zeros(pmax, dtype=torch.float32, device=dev))

    def unique(self, ids_flat: torch.Tensor):
        ops.unique_segment(ids_flat, self.uw)

    def segment_sum(self, g_rows, g_w):
        ops.segment_sum_rows(g_rows, g_w, self.uw, self.K, self.g_uniq, self.gw_uniq if g_w is not None else None)

    def epoch_rows(self, tables_g, j: int, apply: bool):
        """tables_g: [(Table, g_uniq or None)].  apply=False: catch the gathered rows up to the start of
        step j; apply=True: take step j with the de-duplicated gradient."""
        o, uw = self.opt, self.uw
        if (len(tables_g) == 2 and tables_g[0][0].K in ops.EPOCH_ROWS2_K and tables_g[1][0].K == 1
                and tables_g[0][0].N == tables_g[1][0].N):
            (V, gv), (W, gw) = tables_g            # fm_v + fm_w: one launch for both tables
            ev, ew = self.ep[V.name], self.ep[W.name]
            ops.epoch_rows2(o.opt, apply, V, W, ev["last"], ew["last"], uw.uniq, uw.n_uniq, gv if apply else None,
                            gw if apply else None, self.n, o.record(HYPER_TABLE), o.lr_table, j, ev["ss"], ew["ss"])
            return
        for t, g in tables_g:
            e = self.ep[t.name]
            ops.epoch_rows(o.opt, apply, t.var, t.slot(0), t.slot(1), e["last"], uw.uniq, uw.n_uniq,
                           g if apply else None, self.n, t.K, o.record(HYPER_TABLE), o.lr_table, j, e["ss"])

    def epoch_sweep(self, tables: Sequence[Table], upto: int, reset: bool, l2_reg: float):
        """All rows -> state after `upto` steps of this epoch; per-step l2*l2_loss terms -> ep[.]['reg']."""
        o = self.opt
        for t in tables:
            e = self.ep[t.name]
            ev = None
            if self.sweep_events is not None and t.K > 1:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            ops.epoch_sweep(o.opt, t.var, t.slot(0), t.slot(1), e["last"], t.N, t.K, o.record(HYPER_TABLE),
                            o.lr_table, self.flush_pos, upto, reset, e["partials"], e["list"], e["list_count"],
                            e["ss"])
            if ev is not None:
                ev[1].record()
                self.sweep_events.append(ev)
                self.sweep_steps.append(upto - self.flush_pos)   # optimizer steps this pass replayed per element
            # accumulate: a mid-epoch flush and the epoch-end sweep each contribute their share
            ops.epoch_reg_loss(e["ss"], e["partials"], self.n_epart, upto, 0.5 * l2_reg, e["reg"], accumulate=True)
        self.flush_pos = 0 if reset else upto

    def epoch_begin(self):
        for e in self.ep.values():
            ops.fill(e["reg"], 0.0)

    def apply(self, V: Table, W: Optional[Table], exact: bool, l2_reg: float):
        o, uw, hyper = self.opt, self.uw, self.opt.record(HYPER_TABLE)
        n = self.n
        # TF's sparse Adagrad/Momentum/Ftrl touch only gathered rows unless the dense L2 gradient
        # makes every row an index; sparse Adam decays every row regardless.
        sweep = exact and (l2_reg != 0.0 or o.name == "Adam")
        tabs = [(V, self.g_uniq, self.stage_v, self.partials_v, 0)]
        if W is not None:
            tabs.append((W, self.gw_uniq, self.stage_w, self.partials_w, 1))
        for t, g, stage, partials, ri in tabs:
            ops.opt_sparse_rows(o.opt, t.var, t.slot(0), t.slot(1), uw.uniq, uw.n_uniq, g, n, t.K, hyper,
                                stage if sweep else None)
            if sweep:
                ev = None
                if self.sweep_events is not None and ri == 0:  # bench.py: time the dominant kernel live
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record()
                ops.opt_dense_sweep(o.opt, t.var, t.slot(0), t.slot(1), hyper, partials)
                if ev is not None:
                    ev[1].record()
                    self.sweep_events.append(ev)
                ops.opt_patch_rows(t.var, t.slot(0), t.slot(1), uw.uniq, uw.n_uniq, stage, n, t.K, o.n_slots)
                ops.reduce_sum(partials, 0.5 * l2_reg, self.reg[ri:ri + 1], self.red_ws)


class DenseVars:
    """All dense variables of a model in ONE flat fp32 buffer (+ flat grads and slots) so that the
    optimizer apply is a single launch.  Views keep the TF variable names."""

    def __init__(self, specs: Sequence[Tuple[str, Tuple[int, ...]]], opt: OptimizerState, device,
                 l2_names: Sequence[str] = (), tail: int = 4):
        # variables with L2 first, so each group is one contiguous range
        specs = [s for s in specs if s[0] in l2_names] + [s for s in specs if s[0] not in l2_names]
        self.opt = opt
        sizes = [int(math.prod(shape)) for _, shape in specs]
        pad = lambda x: (x + 3) // 4 * 4
        offs, o = [], 0
        for sz in sizes:
            offs.append(o)
            o += pad(sz)
        self.total = o
        self.n_l2 = sum(pad(sz) for (nm, _), sz in zip(specs, sizes) if nm in l2_names)
        f32 = dict(dtype=torch.float32, device=device)
        self.flat = torch.zeros(max(self.total, 4), **f32)
        # `tail` extra floats ride along in the gradient buffer (per-rank loss terms) so that ONE
        # all-reduce covers dense gradients + loss under data parallelism
        self.grad = torch.zeros(max(self.total, 4) + tail, **f32)
        self.tail = self.grad[max(self.total, 4):]
        self.slots = []
        for s in range(opt.n_slots):
            t = torch.empty(max(self.total, 4), **f32)
            ops.fill(t, opt.slot_init(s))
            self.slots.append(t)
        self.views: Dict[str, torch.Tensor] = {}
        self.grads: Dict[str, torch.Tensor] = {}
        for (nm, shape), off, sz in zip(specs, offs, sizes):
            self.views[nm] = self.flat[off:off + sz].view(shape)
            self.grads[nm] = self.grad[off:off + sz].view(shape)

    def __getitem__(self, name):
        return self.views[name]

    def apply(self):
        o = self.opt
        s1 = lambda a, b: (self.slots[1][a:b] if o.n_slots > 1 else None)
        if self.n_l2:
            ops.opt_dense_grad(o.opt, self.flat[:self.n_l2], self.slots[0][:self.n_l2], s1(0, self.n_l2),
                               self.grad[:self.n_l2], o.record(HYPER_DENSE_L2))
        if self.total > self.n_l2:
            a, b = self.n_l2, self.total
            ops.opt_dense_grad(o.opt, self.flat[a:b], self.slots[0][a:b], s1(a, b), self.grad[a:b],
                               o.record(HYPER_DENSE))
