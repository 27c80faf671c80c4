"""oracle.tf_semantics.batch_norm (tf.contrib.layers.batch_norm as batch_norm_layer uses it, DeepFM.py:231-235) against
torch.nn.functional.batch_norm, a third-party implementation of the same published layer: outputs and gradients in TRAIN
mode, outputs in INFER mode, moving mean.  The moving VARIANCE is where the two differ by design -- TF's non-fused path
feeds the biased batch variance into the moving average, torch the unbiased one -- so it is compared after the n/(n-1)
correction, which also documents that choice."""
import torch
import torch.nn.functional as Fnn

from oracle import tf_semantics as tfs

F64 = torch.float64


def _data(n=37, h=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, h, generator=g, dtype=F64) * 2 + 0.5
    gamma = torch.rand(h, generator=g, dtype=F64) + 0.5
    beta = torch.randn(h, generator=g, dtype=F64)
    return x, gamma, beta


def test_train_mode_output_gradients_and_moving_statistics():
    x, gamma, beta = _data()
    n, decay, eps = x.shape[0], 0.9, 1e-3
    mm, mv = torch.zeros(5, dtype=F64), torch.ones(5, dtype=F64)
    a = [t.clone().requires_grad_() for t in (x, gamma, beta)]
    y = tfs.batch_norm(a[0], a[1], a[2], mm, mv, True, decay, eps)
    rm, rv = torch.zeros(5, dtype=F64), torch.ones(5, dtype=F64)
    b = [t.clone().requires_grad_() for t in (x, gamma, beta)]
    z = Fnn.batch_norm(b[0], rm, rv, b[1], b[2], training=True, momentum=1 - decay, eps=eps)
    assert torch.allclose(y, z, rtol=1e-12, atol=1e-12)
    w = torch.linspace(-1, 1, y.numel(), dtype=F64).reshape(y.shape)
    (y * w).sum().backward()
    (z * w).sum().backward()
    for p, q in zip(a, b):
        assert torch.allclose(p.grad, q.grad, rtol=1e-10, atol=1e-12)
    assert torch.allclose(mm, rm, rtol=1e-12, atol=1e-14)
    batch_var = x.var(0, unbiased=False)
    assert torch.allclose(mv, 0.9 * torch.ones(5, dtype=F64) + 0.1 * batch_var, rtol=1e-12)
    assert torch.allclose(rv, 0.9 * torch.ones(5, dtype=F64) + 0.1 * batch_var * n / (n - 1), rtol=1e-12)


def test_infer_mode_uses_the_moving_statistics():
    x, gamma, beta = _data(seed=1)
    mm = torch.randn(5, dtype=F64, generator=torch.Generator().manual_seed(2))
    mv = torch.rand(5, dtype=F64, generator=torch.Generator().manual_seed(3)) + 0.2
    mm0, mv0 = mm.clone(), mv.clone()
    y = tfs.batch_norm(x, gamma, beta, mm, mv, False, 0.9, 1e-3)
    z = Fnn.batch_norm(x, mm0.clone(), mv0.clone(), gamma, beta, training=False, eps=1e-3)
    assert torch.allclose(y, z, rtol=1e-12, atol=1e-12)
    assert torch.equal(mm, mm0) and torch.equal(mv, mv0)          # untouched outside TRAIN
