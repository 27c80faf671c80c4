"""TF-checkpoint naming.  `model.variables()` already uses the reference graph's variable names
(`fm_bias`, `fm_w`, `fm_v`, `Deep-part/mlp0/weights`, ...: DeepFM.py:114-116,156,165); this module adds the names
tf.train.Saver gives the OPTIMIZER state, so a whole training state can travel as one `{tf name: array}` mapping:

    Adam      <var>/Adam (m), <var>/Adam_1 (v), beta1_power, beta2_power      [TF-sem: slot names "m"/"v" are saved as
    Adagrad   <var>/Adagrad                                                     Adam / Adam_1 by Optimizer._slot_dict order]
    Momentum  <var>/Momentum
    Ftrl      <var>/Ftrl (accum), <var>/Ftrl_1 (linear)
    global_step

`export_npz` / `import_npz` write/read that mapping as a NumPy archive ("/" kept in the keys).  On the TensorFlow side
`tf.train.load_checkpoint(path).get_tensor(name)` produces, and `tf.assign` / `init_from_checkpoint` consumes, exactly
these names and shapes.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

SLOT_SUFFIX = {"Adam": ["Adam", "Adam_1"], "Adagrad": ["Adagrad"], "Momentum": ["Momentum"], "ftrl": ["Ftrl", "Ftrl_1"]}


def state_dict_tf(model) -> Dict[str, np.ndarray]:
    """{TF checkpoint name: array} for variables, optimizer slots, Adam beta powers and global_step."""
    out: Dict[str, np.ndarray] = {}
    for name, v in model.variables().items():
        out[name] = v.detach().cpu().numpy().copy()
    suf = SLOT_SUFFIX[model.opt.name]
    for t in model.tables:
        for s, sfx in zip(t.slots, suf):
            out[f"{t.name}/{sfx}"] = s.detach().cpu().numpy().copy()
    for k, sfx in enumerate(suf):
        flat = model.dense.slots[k]
        for name, view in model.dense.views.items():
            off = (view.data_ptr() - model.dense.flat.data_ptr()) // 4
            out[f"{name}/{sfx}"] = flat[off:off + view.numel()].view(view.shape).detach().cpu().numpy().copy()
    st = model.opt.state.detach().cpu().numpy()
    if model.opt.name == "Adam":
        out["beta1_power"], out["beta2_power"] = np.float32(st[0]), np.float32(st[1])
    out["global_step"] = np.int64(model.global_step)
    return out


def load_state_dict_tf(model, values: Dict[str, np.ndarray], strict: bool = True):
    model.flush()
    suf = SLOT_SUFFIX[model.opt.name]
    want = state_dict_tf(model).keys() if strict else ()
    missing = [k for k in want if k not in values]
    if missing:
        raise KeyError(f"missing TF variables: {missing[:5]}{' ...' if len(missing) > 5 else ''}")
    dev = model.device
    tens = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).to(dev)
    for name, dst in model.variables().items():
        if name in values:
            dst.copy_(tens(values[name]).reshape(dst.shape))
    for t in model.tables:
        for s, sfx in zip(t.slots, suf):
            if f"{t.name}/{sfx}" in values:
                s.copy_(tens(values[f"{t.name}/{sfx}"]).reshape(s.shape))
    for k, sfx in enumerate(suf):
        flat = model.dense.slots[k]
        for name, view in model.dense.views.items():
            key = f"{name}/{sfx}"
            if key in values:
                off = (view.data_ptr() - model.dense.flat.data_ptr()) // 4
                flat[off:off + view.numel()].copy_(tens(values[key]).reshape(-1))
    if model.opt.name == "Adam" and "beta1_power" in values:
        model.opt.state[0] = float(values["beta1_power"]); model.opt.state[1] = float(values["beta2_power"])
    if "global_step" in values:
        model.global_step = int(values["global_step"])
        model.opt.state[3] = float(model.global_step)


def export_npz(model, path: str):
    np.savez(path, **{k.replace("/", "|"): v for k, v in state_dict_tf(model).items()})


def import_npz(model, path: str, strict: bool = True):
    with np.load(path) as z:
        load_state_dict_tf(model, {k.replace("|", "/"): z[k] for k in z.files}, strict)
