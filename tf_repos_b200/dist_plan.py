"""Host-side (numpy) statement of the row-sharding plan used by tf_repos_b200/sharded.py; the CUDA routing
kernels (csrc/shard.cu) are tested against it and the world_size-2 gloo tests exercise it on CPU."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def owner_of(ids: np.ndarray, G: int) -> np.ndarray:
    return ids % G


def local_row(ids: np.ndarray, G: int) -> np.ndarray:
    return ids // G


def local_rows(N: int, G: int, rank: int) -> int:
    return (N - rank + G - 1) // G


def route_plan(uniq: np.ndarray, G: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """uniq: ascending unique ids of a batch -> (counts [G], order [U] (stable, bucket-major), local_ids [U])."""
    own = owner_of(uniq, G)
    order = np.argsort(own, kind="stable")
    counts = np.bincount(own, minlength=G)
    return counts.astype(np.int64), order.astype(np.int64), local_row(uniq[order], G).astype(np.int64)


def split_sizes(counts: np.ndarray) -> List[int]:
    return [int(c) for c in counts]


def composite_keys(ids: np.ndarray, N: int, G: int) -> np.ndarray:
    """key(id) = owner * ceil(N/G) + local row: sorting the keys orders ids by (owner, id) -- what csrc/shard.cu's
    ctr_shard_keys emits and tf_repos_b200/sharded.py sorts (one radix sort per step)."""
    npad = (N + G - 1) // G
    return owner_of(ids, G) * npad + local_row(ids, G)


def key_plan(ids: np.ndarray, N: int, G: int):
    """(counts [G], local_ids [U] bucket-major, cache_pos [n]) from the composite keys of a batch's ids."""
    npad = (N + G - 1) // G
    keys = composite_keys(ids.astype(np.int64), N, G)
    uniq, inverse = np.unique(keys, return_inverse=True)
    own = uniq // npad
    return np.bincount(own, minlength=G).astype(np.int64), (uniq - own * npad).astype(np.int64), inverse.astype(np.int64)
