"""libsvm `input_fn`, same contract as deep_ctr/Model_pipeline/DeepFM.py:63-98:

    input_fn(filenames, batch_size=32, num_epochs=1, perform_shuffle=False)
      -> iterator of ({"feat_ids": int32 [B,F,1], "feat_vals": float32 [B,F,1]}, labels float32 [B])

TextLineDataset -> decode_libsvm -> [shuffle(256)] -> repeat(num_epochs) -> batch(batch_size): `repeat`
comes BEFORE `batch`, so batches straddle file and epoch boundaries and only the very last batch may be
partial (it is kept).  Tokenising is done by the native parser in libctr_b200.so (ctr_parse_libsvm,
csrc/libsvm_host.cu) on a small thread pool (the reference uses num_parallel_calls=10).
Tensors are returned in pinned host memory when CUDA is available (ready for an async H2D copy).
"""
from __future__ import annotations

import ctypes
import random
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Iterator, List, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib

_L = _lib.raw()
CHUNK = 32 << 20  # bytes per parse task


def _split_chunks(data: bytes) -> List[Tuple[int, int]]:
    out, pos, n = [], 0, len(data)
    while pos < n:
        end = min(n, pos + CHUNK)
        if end < n:
            nl = data.find(b"\n", end)
            end = n if nl < 0 else nl + 1
        out.append((pos, end))
        pos = end
    return out


def _parse(data: bytes, lo: int, hi: int, F: int):
    view = memoryview(data)[lo:hi]
    buf = ctypes.c_char_p(bytes(view)) if lo or hi != len(data) else ctypes.c_char_p(data)
    n_bytes = hi - lo
    max_rows = max(1, n_bytes // max(2 * F + 2, 1))  # every row has >= 2F+2 characters
    ids = np.empty((max_rows, F), dtype=np.int32)
    vals = np.empty((max_rows, F), dtype=np.float32)
    labels = np.empty(max_rows, dtype=np.float32)
    consumed = ctypes.c_size_t(0)
    rows = _L.ctr_parse_libsvm(buf, n_bytes, F, max_rows, 1, ids.ctypes.data, vals.ctypes.data,
                               labels.ctypes.data, ctypes.byref(consumed))
    if rows < 0:
        raise ValueError(_lib.last_error())
    return ids[:rows], vals[:rows], labels[:rows]


def decode_libsvm_file(path: str, field_size: int = 0, threads: int = 10):
    """Whole file -> (ids int32 [n,F], vals f32 [n,F], labels f32 [n]).  field_size 0 = infer from line 1."""
    with open(path, "rb") as fh:
        data = fh.read()
    F = field_size or _L.ctr_libsvm_count_fields(data, len(data))
    if F <= 0 or len(data) == 0:
        return (np.empty((0, max(field_size, 0)), np.int32), np.empty((0, max(field_size, 0)), np.float32),
                np.empty(0, np.float32))
    chunks = _split_chunks(data)
    if len(chunks) == 1:
        parts = [_parse(data, chunks[0][0], chunks[0][1], F)]
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            parts = list(ex.map(lambda c: _parse(data, c[0], c[1], F), chunks))
    return (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]),
            np.concatenate([p[2] for p in parts]))


def decode_libsvm_file_device(path: str, field_size: int = 0, device="cuda", chunk_bytes: int = 256 << 20):
    """Whole file -> (ids, vals, labels) as CUDA tensors, tokenised on the GPU (ctr_parse_libsvm_device): the file
    is read in chunks cut at line ends, copied to the device and parsed there; a chunk the device parser
    declines (blank/malformed line, a number only strtof may decide) is re-parsed by the host parser, so results and
    error messages are exactly those of decode_libsvm_file."""
    from . import ops
    with open(path, "rb") as fh:
        data = fh.read()
    F = field_size or _L.ctr_libsvm_count_fields(data, len(data))
    dev = torch.device(device)
    if F <= 0 or len(data) == 0:
        return (torch.empty(0, max(field_size, 0), dtype=torch.int32, device=dev),
                torch.empty(0, max(field_size, 0), dtype=torch.float32, device=dev),
                torch.empty(0, dtype=torch.float32, device=dev))
    parts = []
    pos, n = 0, len(data)
    while pos < n:
        end = min(n, pos + chunk_bytes)
        if end < n:
            nl = data.find(b"\n", end)
            end = n if nl < 0 else nl + 1
        raw = np.frombuffer(data, dtype=np.uint8, count=end - pos, offset=pos)
        text = torch.from_numpy(raw.copy()).to(dev, non_blocking=False)
        max_rows = max(1, (end - pos) // max(2 * F + 2, 1))
        ids, vals, labels, consumed, needs_host = ops.parse_libsvm_device(text, F, max_rows, final_chunk=True)
        if needs_host or consumed != end - pos:
            h = _parse(data, pos, end, F)
            ids, vals, labels = (torch.from_numpy(a.copy()).to(dev) for a in h)
        parts.append((ids, vals, labels))
        pos = end
    if len(parts) == 1:
        return parts[0]
    return tuple(torch.cat([p[k] for p in parts]) for k in range(3))


def _pin(t: torch.Tensor) -> torch.Tensor:
    return t.pin_memory() if torch.cuda.is_available() else t


def _input_fn_device(files, batch_size, num_epochs, field_size, device):
    """Same batching (repeat before batch, last partial batch kept), tensors tokenised on and left on the GPU."""
    carry = None
    for _ in range(num_epochs):
        for path in files:
            ids, vals, labels = decode_libsvm_file_device(path, field_size, device)
            if carry is not None:
                ids, vals, labels = (torch.cat([c, t]) for c, t in zip(carry, (ids, vals, labels)))
                carry = None
            n_full = (labels.shape[0] // batch_size) * batch_size
            for lo in range(0, n_full, batch_size):
                hi = lo + batch_size
                yield ({"feat_ids": ids[lo:hi].unsqueeze(-1), "feat_vals": vals[lo:hi].unsqueeze(-1)}, labels[lo:hi])
            if n_full < labels.shape[0]:
                carry = (ids[n_full:], vals[n_full:], labels[n_full:])
    if carry is not None and carry[2].shape[0]:
        yield ({"feat_ids": carry[0].unsqueeze(-1), "feat_vals": carry[1].unsqueeze(-1)}, carry[2])


def input_fn(filenames: Union[str, Sequence[str]], batch_size: int = 32, num_epochs: int = 1,
             perform_shuffle: bool = False, field_size: int = 0,
             device=None) -> Iterator[Tuple[Dict[str, torch.Tensor], torch.Tensor]]:
    """device=None: host parser, pinned host tensors.  device="cuda[:i]": the text is copied to the GPU and
    tokenised there (ctr_parse_libsvm_device); batches are CUDA tensors.  Identical values either way."""
    print("Parsing", filenames)  # DeepFM.py:64
    files = [filenames] if isinstance(filenames, str) else list(filenames)
    if device is not None and not perform_shuffle:
        yield from _input_fn_device(files, batch_size, num_epochs, field_size, device)
        return

    def rows():
        for _ in range(num_epochs):
            for path in files:
                ids, vals, labels = decode_libsvm_file(path, field_size)
                if perform_shuffle:  # tf.data shuffle(buffer_size=256) window semantics
                    buf: List[int] = []
                    order = []
                    for i in range(len(labels)):
                        if len(buf) < 256:
                            buf.append(i)
                            continue
                        j = random.randrange(256)
                        order.append(buf[j]); buf[j] = i
                    random.shuffle(buf)
                    order.extend(buf)
                    ids, vals, labels = ids[order], vals[order], labels[order]
                yield ids, vals, labels

    carry = None
    for ids, vals, labels in rows():
        if carry is not None:
            ids = np.concatenate([carry[0], ids]); vals = np.concatenate([carry[1], vals])
            labels = np.concatenate([carry[2], labels])
            carry = None
        n_full = (len(labels) // batch_size) * batch_size
        for lo in range(0, n_full, batch_size):
            hi = lo + batch_size
            yield ({"feat_ids": _pin(torch.from_numpy(ids[lo:hi].copy()).unsqueeze(-1)),
                    "feat_vals": _pin(torch.from_numpy(vals[lo:hi].copy()).unsqueeze(-1))},
                   _pin(torch.from_numpy(labels[lo:hi].copy())))
        if n_full < len(labels):
            carry = (ids[n_full:], vals[n_full:], labels[n_full:])
    if carry is not None and len(carry[2]):
        yield ({"feat_ids": _pin(torch.from_numpy(carry[0].copy()).unsqueeze(-1)),
                "feat_vals": _pin(torch.from_numpy(carry[1].copy()).unsqueeze(-1))},
               _pin(torch.from_numpy(carry[2].copy())))
