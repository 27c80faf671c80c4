"""tf_repos_b200/tf_names.py on a CPU stand-in for a model: the TF checkpoint names of variables, optimizer slots,
beta powers and global_step, and the .npz round trip.  (The same on a real CUDA model: tests/test_gpu_cli.py.)"""
import os
import types

import numpy as np
import pytest
import torch

from tf_repos_b200 import tf_names


class _Table:
    def __init__(self, name, shape, n_slots, seed):
        g = torch.Generator().manual_seed(seed)
        self.name = name
        self.var = torch.randn(shape, generator=g)
        self.slots = [torch.rand(shape, generator=g) for _ in range(n_slots)]


class _Dense:
    def __init__(self, specs, n_slots, seed):
        g = torch.Generator().manual_seed(seed)
        total = sum(int(np.prod(s)) for _, s in specs)
        self.flat = torch.randn(total, generator=g)
        self.slots = [torch.rand(total, generator=g) for _ in range(n_slots)]
        self.views, off = {}, 0
        for name, shape in specs:
            n = int(np.prod(shape))
            self.views[name] = self.flat[off:off + n].view(shape)
            off += n


def _model(opt, seed):
    n_slots = {"Adam": 2, "Adagrad": 1, "Momentum": 1, "ftrl": 2}[opt]
    m = types.SimpleNamespace()
    m.tables = [_Table("fm_v", (50, 4), n_slots, seed), _Table("fm_w", (50,), n_slots, seed + 1)]
    m.dense = _Dense([("fm_bias", (1,)), ("Deep-part/mlp0/weights", (8, 3)), ("Deep-part/mlp0/biases", (3,))], n_slots, seed + 2)
    m.opt = types.SimpleNamespace(name=opt, state=torch.tensor([0.9 ** 3, 0.999 ** 3, 5e-4, 3.0]))
    m.global_step, m.device = 3, torch.device("cpu")
    m.flush = lambda: None
    m.variables = lambda: {**{t.name: t.var for t in m.tables}, **m.dense.views}
    return m


@pytest.mark.parametrize("opt,slot_names", [("Adam", ["Adam", "Adam_1"]), ("Adagrad", ["Adagrad"]), ("Momentum", ["Momentum"]),
                                             ("ftrl", ["Ftrl", "Ftrl_1"])])
def test_names_and_roundtrip(tmp_path, opt, slot_names):
    a = _model(opt, 1)
    sd = tf_names.state_dict_tf(a)
    want = {"fm_v", "fm_w", "fm_bias", "Deep-part/mlp0/weights", "Deep-part/mlp0/biases", "global_step"}
    for v in ("fm_v", "fm_w", "fm_bias", "Deep-part/mlp0/weights", "Deep-part/mlp0/biases"):
        want |= {f"{v}/{s}" for s in slot_names}
    if opt == "Adam":
        want |= {"beta1_power", "beta2_power"}
    assert set(sd) == want
    assert sd["Deep-part/mlp0/weights/" + slot_names[0]].shape == (8, 3) and sd["fm_v/" + slot_names[-1]].shape == (50, 4)
    assert np.array_equal(sd["fm_w/" + slot_names[0]], a.tables[1].slots[0].numpy())
    # dense slots are cut out of the flat slot buffer at the variable's own offset
    off = 1
    assert np.array_equal(sd["Deep-part/mlp0/weights/" + slot_names[0]].reshape(-1), a.dense.slots[0][off:off + 24].numpy())
    path = os.path.join(tmp_path, "s.npz")
    tf_names.export_npz(a, path)
    b = _model(opt, 99)                       # different contents everywhere
    b.global_step = 0
    tf_names.import_npz(b, path)
    sb = tf_names.state_dict_tf(b)
    for k in sd:
        assert np.array_equal(np.asarray(sd[k]), np.asarray(sb[k])), k
    assert b.global_step == 3 and float(b.opt.state[3]) == 3.0
    # strict import reports what is missing
    sd.pop("fm_v/" + slot_names[0])
    with pytest.raises(KeyError):
        tf_names.load_state_dict_tf(_model(opt, 5), sd, strict=True)
    tf_names.load_state_dict_tf(_model(opt, 5), sd, strict=False)
