"""The dense 'Deep-part' (DeepFM.py:137-167): fully_connected(relu) -> [BN] -> dropout stacks.

fp32 throughout (logit parity target is 1e-5 relative, which rules out TF32/BF16 tensor-core
inputs without split-precision emulation).  Forward/backward buffers are allocated once so the
whole step can be captured in a CUDA graph.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from .engine import DenseVars


class MLP:
    """layers: hidden widths; final_out: append a Linear(->1, identity) named `{scope}/{out_scope}`."""

    def __init__(self, in_dim: int, layers: Sequence[int], keep_prob: Sequence[float], B: int, device,
                 scope: str = "Deep-part", out_scope: Optional[str] = "deep_out", out_extra_in: int = 0):
        self.in_dim, self.layers, self.keep = in_dim, list(layers), list(keep_prob)
        self.scope, self.out_scope, self.B, self.device = scope, out_scope, B, device
        self.out_in = (self.layers[-1] if self.layers else in_dim) + out_extra_in
        f32 = dict(dtype=torch.float32, device=device)
        self.h = [torch.empty(B, w, **f32) for w in self.layers]          # post-activation (post-dropout)
        self.mask = [None] * len(self.layers)                              # keep-scaled dropout masks
        self.dh = [torch.empty(B, w, **f32) for w in self.layers]
        self.y = torch.empty(B, **f32)
        self.dx = torch.empty(B, in_dim, **f32)

    def specs(self):
        out, d = [], self.in_dim
        for i, w in enumerate(self.layers):
            out += [(f"{self.scope}/mlp{i}/weights", (d, w)), (f"{self.scope}/mlp{i}/biases", (w,))]
            d = w
        if self.out_scope:
            out += [(f"{self.scope}/{self.out_scope}/weights", (self.out_in, 1)),
                    (f"{self.scope}/{self.out_scope}/biases", (1,))]
        return out

    def init(self, dv: DenseVars, gen: torch.Generator):
        """xavier_uniform weights, zero biases (tf.contrib.layers.fully_connected defaults)."""
        for name, shape in self.specs():
            if name.endswith("weights"):
                lim = (6.0 / (shape[0] + shape[1])) ** 0.5
                w = (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * lim
                dv[name].copy_(w.to(torch.float32))

    # ---- forward -------------------------------------------------------------------------------
    def forward_hidden(self, x: torch.Tensor, dv: DenseVars, train: bool, masks=None) -> torch.Tensor:
        a = x
        for i in range(len(self.layers)):
            W, b = dv[f"{self.scope}/mlp{i}/weights"], dv[f"{self.scope}/mlp{i}/biases"]
            h = self.h[i][: a.shape[0]]
            torch.addmm(b, a, W, out=h)
            h.relu_()
            self.mask[i] = None
            if train and (masks is not None or self.keep[i] < 1.0):
                if masks is not None:
                    m = masks[i] / self.keep[i]
                else:
                    m = torch.empty_like(h).bernoulli_(self.keep[i]).div_(self.keep[i])
                self.mask[i] = m
                h.mul_(m)
            a = h
        return a

    def forward_out(self, a: torch.Tensor, dv: DenseVars) -> torch.Tensor:
        W, b = dv[f"{self.scope}/{self.out_scope}/weights"], dv[f"{self.scope}/{self.out_scope}/biases"]
        y = self.y[: a.shape[0]]
        torch.addmv(b.expand(a.shape[0]), a, W.view(-1), out=y)
        return y

    # ---- backward ------------------------------------------------------------------------------
    def backward_out(self, a: torch.Tensor, dy: torch.Tensor, dv: DenseVars, da: torch.Tensor):
        """y = a @ W + b ; dy [B] -> dW, db, da."""
        W = dv[f"{self.scope}/{self.out_scope}/weights"]
        torch.mv(a.t(), dy, out=dv.grads[f"{self.scope}/{self.out_scope}/weights"].view(-1))
        torch.sum(dy, dim=0, keepdim=True, out=dv.grads[f"{self.scope}/{self.out_scope}/biases"])
        torch.mul(dy.unsqueeze(1), W.view(1, -1), out=da)

    def backward_hidden(self, x: torch.Tensor, d_last: torch.Tensor, dv: DenseVars, need_dx: bool = True):
        """d_last: gradient w.r.t. the last hidden activation (post-dropout)."""
        d = d_last
        for i in reversed(range(len(self.layers))):
            W = dv[f"{self.scope}/mlp{i}/weights"]
            h = self.h[i][: d.shape[0]]
            if self.mask[i] is not None:
                d.mul_(self.mask[i])
            d.mul_(h > 0)  # relu' (h is post-dropout: h>0 iff relu output >0 and kept)
            a = self.h[i - 1][: d.shape[0]] if i > 0 else x
            torch.mm(a.t(), d, out=dv.grads[f"{self.scope}/mlp{i}/weights"])
            torch.sum(d, dim=0, out=dv.grads[f"{self.scope}/mlp{i}/biases"])
            if i > 0:
                nd = self.dh[i - 1][: d.shape[0]]
                torch.mm(d, W.t(), out=nd)
                d = nd
            elif need_dx:
                torch.mm(d, W.t(), out=self.dx[: d.shape[0]])
        return self.dx[: d.shape[0]] if need_dx else None
