"""The epoch sweep's grouped IEEE fast path (csrc/optim_steps.cuh) only pays off while the operands of rows nothing
gathers stay inside its guarded range.  This CPU simulation of that recurrence (TF's non-lazy Adam with the dense L2
gradient, reference defaults: lr 5e-4, l2 1e-4, glorot init of a 2e8 x 16 table) reads the guard constants from the
CUDA header and checks (a) the range covers the first several hundred training steps, (b) what the long-run state looks
like: var parks just above FLT_MIN, m at a few denormal ulps, and lr_t*m is either exactly 0 (most elements) or a
denormal -- the regime the planned power-of-two-scaled fast path has to cover (known gaps).  (The bit-exactness of the fast path itself is a GPU test:
ctr_selftest_divsqrt and tests/test_gpu_deferred.py.)"""
import math
import os
import re

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _guards():
    src = open(os.path.join(ROOT, "tf_repos_b200", "csrc", "optim_steps.cuh")).read()
    val = lambda name: float(re.search(name + r"\s*=\s*([0-9.e+-]+)f", src).group(1))
    return val("SQRT_LO"), val("SQRT_HI"), val("DIV_LO"), val("DIV_HI")


def test_guard_constants_are_the_documented_powers_of_two():
    s_lo, s_hi, d_lo, d_hi = _guards()
    assert math.isclose(s_lo, 2.0 ** -101, rel_tol=1e-6) and math.isclose(s_hi, 2.0 ** 40, rel_tol=1e-6)
    assert math.isclose(d_lo, 2.0 ** -100, rel_tol=1e-6) and math.isclose(d_hi, 2.0 ** 60, rel_tol=1e-6)


def test_untouched_row_recurrence_stays_in_range_then_parks_at_denormals():
    s_lo, s_hi, d_lo, d_hi = _guards()
    f = lambda t: torch.tensor(t, dtype=torch.float32)
    n = 80_000
    g = torch.Generator().manual_seed(0)
    std = math.sqrt(2.0 / (200_000_000 + 16))
    x = (torch.randn(n, generator=g) * std).clamp(-2 * std, 2 * std).float()
    m, v = torch.zeros(n), torch.zeros(n)
    lr, b1, b2, eps, l2 = f(5e-4), f(0.9), f(0.999), f(1e-8), f(1e-4)
    b1p, b2p = b1.clone(), b2.clone()
    out_of_range_groups = {}
    all_zero_groups = {}
    for t in range(1, 1801):
        lr_t = (lr * torch.sqrt(1 - b2p)) / (1 - b1p)
        gr = l2 * x
        m = m * b1 + gr * (1 - b1)
        v = v * b2 + (gr * gr) * (1 - b2)
        a = lr_t * m
        zero = a == 0
        bad = (~zero) & ((v < s_lo) | (v > s_hi) | (a.abs() < d_lo) | (a.abs() > d_hi))
        allz = zero.view(-1, 8).all(1)
        slow = (~allz) & (bad | zero).view(-1, 8).any(1)          # groups that would take the per-element slow path
        x = x - a / (torch.sqrt(v) + eps)
        b1p, b2p = b1p * b1, b2p * b2
        if t in (16, 64, 160, 320, 480, 1800):
            out_of_range_groups[t] = slow.float().mean().item()
            all_zero_groups[t] = allz.float().mean().item()
    for t in (16, 64, 160, 320, 480):
        assert out_of_range_groups[t] < 1e-3, (t, out_of_range_groups)
    # long run: every lr_t*m is 0 or denormal, most are 0, but few 8-element groups are ALL zero
    flt_min = 2.0 ** -126
    assert bool((a.abs() < flt_min).all()) and zero.float().mean().item() > 0.5
    assert all_zero_groups[1800] < 0.5 and out_of_range_groups[1800] > 0.5
    assert 1e-39 < x.abs().median().item() < 1e-36 and m.abs().median().item() < 1e-41
