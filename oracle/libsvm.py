"""Pure-Python restatement of input_fn / decode_libsvm (deep_ctr/Model_pipeline/DeepFM.py:63-98).
TEST INFRASTRUCTURE ONLY (small files).  [TF-sem]: tf.string_split skips empty tokens;
string_to_number(float32) behaves like strtof, int32 like strtol."""
from __future__ import annotations

from typing import Iterator, List, Sequence, Tuple

import numpy as np


def decode_libsvm(line: str) -> Tuple[np.ndarray, np.ndarray, np.float32]:
    columns = [t for t in line.rstrip("\r\n").split(" ") if t != ""]          # tf.string_split([line], ' ')
    label = np.float32(columns[0])                                             # :70
    splits = [[p for p in tok.split(":") if p != ""] for tok in columns[1:]]   # :71
    if any(len(s) != 2 for s in splits):
        raise ValueError("reshape(splits.values, splits.dense_shape) fails: a token is not <id>:<val>")
    feat_ids = np.array([int(s[0]) for s in splits], dtype=np.int32)           # :74
    feat_vals = np.array([np.float32(s[1]) for s in splits], dtype=np.float32)  # :75
    return feat_ids, feat_vals, label


def input_fn(filenames: Sequence[str], batch_size: int = 32, num_epochs: int = 1) -> Iterator:
    """TextLineDataset -> map(decode) -> repeat(num_epochs) -> batch(batch_size); yields
    ({"feat_ids": int32 [B,F,1], "feat_vals": f32 [B,F,1]}, labels f32 [B])."""
    rows: List = []
    for _ in range(num_epochs):
        for fn in filenames:
            with open(fn) as fh:
                for line in fh:
                    if line.strip() == "":
                        continue
                    rows.append(decode_libsvm(line))
    for lo in range(0, len(rows), batch_size):
        chunk = rows[lo: lo + batch_size]
        yield ({"feat_ids": np.stack([c[0] for c in chunk])[..., None],
                "feat_vals": np.stack([c[1] for c in chunk])[..., None]},
               np.array([c[2] for c in chunk], dtype=np.float32))
