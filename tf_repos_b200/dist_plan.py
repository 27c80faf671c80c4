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
