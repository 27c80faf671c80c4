"""wide_n_deep on the CUDA path (tf_repos_b200/wide_deep.py, csrc/wide_deep.cu) against oracle/wide_deep.py:
logits, loss and every updated variable for the three model types, incl. out-of-range ids and a partial batch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batch(B, seed=0, nb=10000):
    g = torch.Generator().manual_seed(seed)
    dense = torch.rand(B, 13, generator=g)
    cat = torch.randint(0, nb, (B, 26), generator=g, dtype=torch.int64).to(torch.int32)
    cat[0, 0] = 12345            # out of range -> bucket 0
    cat[1, 5] = -3
    cat[:, 2] = 7                # one id hit by every sample (a long run for the segment sum)
    labels = (torch.rand(B, generator=g) < 0.3).float()
    return dense, cat, labels


@pytest.mark.parametrize("model_type", ["wide", "deep", "wide_n_deep"])
def test_train_steps_match_oracle(model_type):
    from oracle import wide_deep as owd
    from tf_repos_b200.wide_deep import WideDeep
    K, layers, B = 8, "32,16", 96
    o = owd.WideDeep(embedding_size=K, deep_layers=layers, model_type=model_type, seed=3)
    m = WideDeep(embedding_size=K, batch_size=B, deep_layers=layers, model_type=model_type, seed=3)
    if model_type == "wide":     # all-zero start is a degenerate parity case: give the linear part some weights
        g = torch.Generator().manual_seed(1)
        for n in o.params:
            o.params[n] = (torch.randn(o.params[n].shape, generator=g) * 0.05).float()
    m.load_variables(o.params)
    dev = m.device
    for step in range(4):
        Bs = B if step != 2 else 50                       # a partial batch in the middle
        dense, cat, labels = _batch(Bs, seed=20 + step)
        want_y = o.predict(dense, cat)["y"]
        got_p = m.predict(dense.to(dev), cat.to(dev))
        assert torch.allclose(m.y[:Bs].cpu(), want_y, rtol=1e-5, atol=1e-6)
        assert torch.allclose(got_p.cpu(), torch.sigmoid(want_y), rtol=1e-5, atol=1e-6)
        want_loss = o.train_step(dense, cat, labels)
        got_loss = float(m.train_step(dense.to(dev), cat.to(dev), labels.to(dev)))
        assert abs(got_loss - want_loss) <= 1e-5 * max(1.0, abs(want_loss))
        for name, v in m.variables().items():
            w = o.params[name]
            assert torch.allclose(v.cpu().reshape(w.shape), w, rtol=2e-5, atol=2e-6), (step, name)
