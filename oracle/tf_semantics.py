"""CPU restatement of the TensorFlow-1.x op semantics the reference's model_fns rely on.

TEST INFRASTRUCTURE ONLY.  Nothing under tf_repos_b200/ may import this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.

PARITY UNPINNED: TensorFlow is an un-vendored, unpinned dependency of the reference ("version:
1.4", deep_ctr/README.md:36); it is not installed here, the reference scripts are Python-2 only and
the reference ships no tests, golden vectors or checkpoints.  Every function below therefore
restates the *published* TF 1.x behaviour of the op named in its docstring (marked [TF-sem]) and
cites the reference call site (file:line under deep_ctr/Model_pipeline/) that uses it.  The
restatement itself is pinned by closed-form known answers and by fp64 autograd cross-checks in
tests/test_oracle_*.py.

All arithmetic is done with torch CPU tensors, one torch op per TF op, so every intermediate is
rounded to fp32 exactly where TF's op-by-op executor rounds it (no FMA contraction across ops).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

F32 = torch.float32


# ---------------------------------------------------------------------------------------------
# initialisers
# ---------------------------------------------------------------------------------------------
def glorot_normal(shape, gen: torch.Generator, dtype=F32) -> torch.Tensor:
    """tf.glorot_normal_initializer (DeepFM.py:115-116) [TF-sem]: variance_scaling(scale=1,
    mode='fan_avg', distribution='normal') -> truncated normal, stddev = sqrt(2/(fan_in+fan_out));
    TF 1.4 applies no 0.8796 truncation correction.  Rank-1 shape [n]: fan_in = fan_out = n."""
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    else:
        fan_in, fan_out = shape[-2], shape[-1]
    std = math.sqrt(2.0 / (fan_in + fan_out))
    return truncated_normal(shape, std, gen, dtype)


def truncated_normal(shape, std: float, gen: torch.Generator, dtype=F32) -> torch.Tensor:
    """tf.truncated_normal [TF-sem]: values more than 2 stddev from the mean are re-drawn."""
    out = torch.randn(shape, generator=gen, dtype=torch.float64)
    bad = out.abs() > 2.0
    while bool(bad.any()):
        out[bad] = torch.randn(int(bad.sum()), generator=gen, dtype=torch.float64)
        bad = out.abs() > 2.0
    return (out * std).to(dtype)


def xavier_uniform(shape, gen: torch.Generator, dtype=F32) -> torch.Tensor:
    """tf.contrib.layers.fully_connected default weights_initializer = xavier_initializer()
    (DeepFM.py:156) [TF-sem]: uniform(+-sqrt(6/(fan_in+fan_out)))."""
    fan_in, fan_out = shape
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * lim).to(dtype)


# ---------------------------------------------------------------------------------------------
# forward ops
# ---------------------------------------------------------------------------------------------
def embedding_lookup(params: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """tf.nn.embedding_lookup (DeepFM.py:126,130) [TF-sem]: a gather on axis 0; on CPU an id
    outside [0, N) raises InvalidArgumentError."""
    n = params.shape[0]
    if ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= n):
        raise IndexError(f"indices out of range [0, {n})  (TF: InvalidArgumentError)")
    return params[ids.long()]


def l2_loss(t: torch.Tensor) -> torch.Tensor:
    """tf.nn.l2_loss (DeepFM.py:189-190) [TF-sem]: sum(t**2)/2."""
    return (t * t).sum() / 2


def sigmoid(x: torch.Tensor) -> torch.Tensor:
    """tf.sigmoid (DeepFM.py:176) [TF-sem]: 1/(1+exp(-x))."""
    return 1.0 / (1.0 + torch.exp(-x))


def sigmoid_cross_entropy_with_logits(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """tf.nn.sigmoid_cross_entropy_with_logits (DeepFM.py:188) [TF-sem]:
    max(x,0) - x*z + log(1+exp(-|x|)), written with tf.where(x >= 0, ...) selects exactly like TF's
    nn_impl.py so that autodiff at x == 0 gives sigmoid(0) - z (clamp/abs would give the subgradient
    -z there; logits are exactly 0 whenever every relu of the last hidden layer is dead)."""
    zeros = torch.zeros_like(logits)
    cond = logits >= zeros
    relu_logits = torch.where(cond, logits, zeros)
    neg_abs_logits = torch.where(cond, -logits, logits)
    return relu_logits - logits * labels + torch.log1p(torch.exp(neg_abs_logits))


def fully_connected(x, W, b, activation: Optional[str] = "relu"):
    """tf.contrib.layers.fully_connected (DeepFM.py:156,165) [TF-sem]: activation_fn defaults to
    relu; weights [in,out]; biases zero-initialised.  Its weights_regularizer only populates the
    REGULARIZATION_LOSSES collection, which the reference never adds to `loss` -> no effect."""
    y = x @ W + b
    if activation == "relu":
        y = torch.relu(y)
    elif activation == "sigmoid":
        y = sigmoid(y)
    elif activation is not None:
        raise ValueError(activation)
    return y


def dropout(x, keep_prob: float, mask: Optional[torch.Tensor]):
    """tf.nn.dropout(x, keep_prob) (DeepFM.py:162) [TF-sem]: x / keep_prob * binary_mask, TRAIN
    only.  TF's Philox stream is not reproducible here: parity runs inject `mask` (0/1) or use
    keep_prob = 1 (mask None => identity)."""
    if mask is None:
        return x
    return x / keep_prob * mask


def batch_norm(x, gamma, beta, moving_mean, moving_var, train: bool, decay: float, eps: float = 1e-3):
    """tf.contrib.layers.batch_norm(decay, center, scale, updates_collections=None)
    (DeepFM.py:231-235) [TF-sem]: epsilon 0.001; train => batch moments (biased variance) and
    in-place moving-average update  mv = mv*decay + batch*(1-decay); infer => moving stats."""
    if train:
        mean = x.mean(0)
        var = x.var(0, unbiased=False)
        with torch.no_grad():
            moving_mean.mul_(decay).add_(mean.detach() * (1 - decay))
            moving_var.mul_(decay).add_(var.detach() * (1 - decay))
    else:
        mean, var = moving_mean, moving_var
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


def auc(labels: np.ndarray, preds: np.ndarray, num_thresholds: int = 200) -> float:
    """tf.metrics.auc (DeepFM.py:194) [TF-sem]: ROC curve at `num_thresholds` thresholds
    ((i+1)/(n-1) for the inner ones, endpoints -eps and 1+eps, eps = 1e-7), trapezoidal sum of
    tp_rate over fp_rate with eps-guarded divisions."""
    kepsilon = 1e-7
    thr = [(i + 1) * 1.0 / (num_thresholds - 1) for i in range(num_thresholds - 2)]
    thr = np.array([0.0 - kepsilon] + thr + [1.0 + kepsilon], dtype=np.float32)
    labels = labels.astype(bool)
    preds = preds.astype(np.float32)
    pred_pos = preds[None, :] > thr[:, None]
    tp = (pred_pos & labels[None, :]).sum(1).astype(np.float32)
    fp = (pred_pos & ~labels[None, :]).sum(1).astype(np.float32)
    fn = (~pred_pos & labels[None, :]).sum(1).astype(np.float32)
    tn = (~pred_pos & ~labels[None, :]).sum(1).astype(np.float32)
    eps = np.float32(1.0e-6)
    rec = (tp + eps) / (tp + fn + eps)
    fp_rate = fp / (fp + tn + eps)
    x, y = fp_rate, rec
    return float(np.sum((x[: num_thresholds - 1] - x[1:]) * (y[: num_thresholds - 1] + y[1:]) / 2.0))


# ---------------------------------------------------------------------------------------------
# gradient aggregation
# ---------------------------------------------------------------------------------------------
def deduplicate_indexed_slices(values: np.ndarray, indices: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """optimizer._deduplicate_indexed_slices [TF-sem]: unique_indices, new_positions =
    tf.unique(indices); summed = tf.unsorted_segment_sum(values, new_positions, n_unique).
    The CPU kernel accumulates occurrences sequentially in input order (fp32).  tf.unique returns
    first-occurrence order, which is unobservable after the scatter; this restatement returns the
    ids ASCENDING (the order the CUDA path defines), with identical per-id sums."""
    uniq, inverse = np.unique(indices, return_inverse=True)
    summed = np.zeros((uniq.shape[0],) + values.shape[1:], dtype=values.dtype)
    np.add.at(summed, inverse, values)  # unbuffered, sequential in occurrence order
    return summed, uniq


def unique_segment_reference(ids: np.ndarray):
    """Integer outputs of the K3 contract (bit-exact targets): perm, uniq, inverse, seg_offsets."""
    ids = np.asarray(ids)
    perm = np.argsort(ids, kind="stable").astype(np.int32)
    uniq, inverse, counts = np.unique(ids, return_inverse=True, return_counts=True)
    seg = np.zeros(uniq.shape[0] + 1, dtype=np.int32)
    np.cumsum(counts, out=seg[1:])
    return perm, uniq.astype(np.int32), inverse.astype(np.int32), seg


# ---------------------------------------------------------------------------------------------
# optimizers (fp32, one rounding per elementary op, exactly the expression order of TF 1.x)
# ---------------------------------------------------------------------------------------------
def f32(x) -> torch.Tensor:
    return torch.tensor(x, dtype=F32)


# torch's CPU sqrt goes through MKL VML and is NOT correctly rounded (0.7 % of random fp32 inputs
# differ from IEEE by 1 ulp, measured); numpy's is (hardware sqrtps).  Parity runs need the IEEE
# result (the CUDA kernels use __fsqrt_rn); the timed cpu_baseline may flip FAST_SQRT to use all
# host threads.
FAST_SQRT = False


def ieee_sqrt(t: torch.Tensor) -> torch.Tensor:
    if FAST_SQRT or t.dtype != F32:
        return torch.sqrt(t)
    return torch.from_numpy(np.atleast_1d(np.sqrt(t.detach().numpy()))).reshape(t.shape)


class AdamHyper:
    """tf.train.AdamOptimizer(learning_rate, 0.9, 0.999, 1e-8) (DeepFM.py:205) [TF-sem].
    beta{1,2}_power are fp32 variables initialised to beta{1,2} and multiplied by beta{1,2} in
    _finish() after every apply."""

    def __init__(self, lr, beta1=0.9, beta2=0.999, eps=1e-8, dtype=F32):
        self.dtype = dtype
        self.lr = torch.tensor(lr, dtype=dtype)
        self.b1 = torch.tensor(beta1, dtype=dtype)
        self.b2 = torch.tensor(beta2, dtype=dtype)
        self.eps = torch.tensor(eps, dtype=dtype)
        self.b1p = self.b1.clone()
        self.b2p = self.b2.clone()

    def lr_t(self) -> torch.Tensor:
        """lr * sqrt(1 - beta2_power) / (1 - beta1_power), evaluated left to right in fp32."""
        one = torch.tensor(1.0, dtype=self.dtype)
        return (self.lr * ieee_sqrt(one - self.b2p)) / (one - self.b1p)

    def finish(self):
        self.b1p = self.b1p * self.b1
        self.b2p = self.b2p * self.b2


_SCRATCH = {}


def _scratch(like: torch.Tensor, slot: int) -> torch.Tensor:
    key = (tuple(like.shape), like.dtype, slot)
    if key not in _SCRATCH:
        _SCRATCH.clear() if len(_SCRATCH) > 8 else None
        _SCRATCH[key] = torch.empty_like(like)
    return _SCRATCH[key]


def adam_sparse_(var, m, v, g, lr_t, b1, b2, eps):
    """AdamOptimizer._apply_sparse_shared [TF-sem] on the rows given (in place):
        m <- m*b1 ; m += g*(1-b1) ; v <- v*b2 ; v += (g*g)*(1-b2)
        var <- var - (lr_t*m)/(sqrt(v)+eps)
    (TF decays m and v of EVERY row and updates every row of var; callers pass whole tables.)
    Large tables take the same op-by-op arithmetic through two reused scratch buffers (the timed CPU
    baseline would otherwise spend most of its time page-faulting fresh 12.8 GB temporaries, which TF's
    BFC allocator does not do)."""
    one = torch.ones((), dtype=var.dtype)
    if var.numel() >= (1 << 24) and g.shape == var.shape:
        t = _scratch(var, 0)
        torch.mul(g, one - b1, out=t); m.mul_(b1); m.add_(t)
        torch.mul(g, g, out=t); t.mul_(one - b2); v.mul_(b2); v.add_(t)
        torch.mul(m, lr_t, out=t)
        d = _scratch(var, 1)
        if FAST_SQRT or var.dtype != F32:
            torch.sqrt(v, out=d)
        else:
            d.copy_(ieee_sqrt(v))
        d.add_(eps); t.div_(d); var.sub_(t)
        return
    m.mul_(b1).add_(g * (one - b1))
    v.mul_(b2).add_((g * g) * (one - b2))
    var.sub_((lr_t * m) / (ieee_sqrt(v) + eps))


def adam_dense_(var, m, v, g, lr_t, b1, b2, eps):
    """training_ops.apply_adam CPU functor [TF-sem] (dense variables, no nesterov):
        m += (g - m)*(1-b1) ; v += (g*g - v)*(1-b2) ; var -= (m*alpha)/(sqrt(v)+eps)"""
    one = torch.ones((), dtype=var.dtype)
    m.add_((g - m) * (one - b1))
    v.add_((g * g - v) * (one - b2))
    var.sub_((m * lr_t) / (ieee_sqrt(v) + eps))


def adagrad_(var, acc, g, lr):
    """ApplyAdagrad / SparseApplyAdagrad [TF-sem] (DeepFM.py:207; initial_accumulator_value=1e-8):
        acc += g*g ; var -= lr*g*rsqrt(acc)   (rsqrt restated as 1/sqrt, both IEEE)"""
    acc.add_(g * g)
    var.sub_((lr * g) * (torch.ones((), dtype=var.dtype) / ieee_sqrt(acc)))


def momentum_(var, acc, g, lr, momentum):
    """ApplyMomentum / SparseApplyMomentum [TF-sem] (DeepFM.py:209, momentum=0.95, no nesterov):
        acc <- acc*momentum + g ; var -= acc*lr"""
    acc.mul_(momentum).add_(g)
    var.sub_(acc * lr)


def ftrl_(var, accum, linear, g, lr, lr_power=-0.5, l1=0.0, l2=0.0):
    """ApplyFtrl [TF-sem] (DeepFM.py:211; defaults learning_rate_power=-0.5,
    initial_accumulator_value=0.1, l1=l2=0)."""
    dt = var.dtype
    lr = torch.as_tensor(lr, dtype=dt)
    l1t = torch.tensor(l1, dtype=dt)
    l2t = torch.tensor(l2, dtype=dt)
    new_accum = accum + g * g
    if lr_power == -0.5:
        pn, po = ieee_sqrt(new_accum), ieee_sqrt(accum)
    else:
        pn, po = torch.pow(new_accum, -lr_power), torch.pow(accum, -lr_power)
    linear.add_(g - ((pn - po) / lr) * var)
    x = l1t * torch.sign(linear) - linear
    y = pn / lr + torch.tensor(2.0, dtype=dt) * l2t
    var.copy_(torch.where(linear.abs() > l1t, x / y, torch.zeros((), dtype=dt)))
    accum.copy_(new_accum)
