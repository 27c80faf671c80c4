"""TFRecord / tf.Example reader+writer (tf_repos_b200/tfrecord.py) and the DIN input pipeline (din_main.py): CPU only."""
import os
import struct

import numpy as np
import pytest

from tf_repos_b200 import tfrecord as tfr


def test_crc32c_known_answers():
    assert tfr.crc32c(b"123456789") == 0xE3069283            # the CRC-32C check value
    assert tfr.crc32c(b"") == 0
    assert tfr.crc32c(bytes(32)) == 0x8A9136AA               # RFC 3720 B.4: 32 bytes of zeros
    assert tfr.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43      # RFC 3720 B.4: 32 bytes of ones


def test_example_roundtrip_and_unpacked_encoding(tmp_path):
    ex = {"y": np.float32(1.0), "z": np.float32(0.0), "feat_ids": np.arange(11, dtype=np.int64) * 1000 + 7,
          "u_catids": np.array([5, 6, 7], dtype=np.int64), "u_catvals": np.array([0.5, 1.25, 3.0], dtype=np.float32),
          "a_catids": np.int64(9), "a_intids": np.array([], dtype=np.int64), "neg": np.array([-1, -(2 ** 40)], dtype=np.int64)}
    data = tfr.encode_example(ex)
    got = tfr.parse_example(data)
    for k, v in ex.items():
        assert np.array_equal(np.atleast_1d(v), got[k]), k
        assert got[k].dtype == (np.float32 if np.asarray(v).dtype.kind == "f" else np.int64)
    # hand-built UNPACKED Int64List / FloatList (field 1 repeated as varint / fixed32): parsers must accept both
    i64 = b"".join(b"\x08" + tfr._enc_varint(v) for v in (3, 300))
    f32 = b"".join(b"\x0d" + struct.pack("<f", v) for v in (0.5, 2.0))
    entries = tfr._ld(1, tfr._ld(1, b"a") + tfr._ld(2, tfr._ld(3, i64))) + tfr._ld(1, tfr._ld(1, b"b") + tfr._ld(2, tfr._ld(2, f32)))
    got = tfr.parse_example(tfr._ld(1, entries))
    assert got["a"].tolist() == [3, 300] and got["b"].tolist() == [0.5, 2.0]
    # framing
    p = os.path.join(tmp_path, "x.tfrecord")
    tfr.write_records(p, [data, b"", data[:10]])
    recs = list(tfr.read_records(p, verify_crc=True))
    assert recs == [data, b"", data[:10]]
    raw = bytearray(open(p, "rb").read())
    raw[20] ^= 1
    open(p, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        list(tfr.read_records(p, verify_crc=True))
    open(p, "wb").write(bytes(raw[:-3]))
    with pytest.raises(ValueError):
        list(tfr.read_records(p))


def _write_din(path, n, seed, F=11, N=5000, maxlen=6):
    rng = np.random.RandomState(seed)
    recs = []
    for _ in range(n):
        ex = {"y": np.float32(rng.rand() < 0.3), "z": np.float32(0.0), "feat_ids": rng.randint(1, N, F).astype(np.int64),
              "a_catids": np.int64(rng.randint(1, N)), "a_shopids": np.int64(rng.randint(1, N)),
              "a_brandids": np.int64(rng.randint(1, N)), "a_intids": rng.randint(1, N, rng.randint(0, 4)).astype(np.int64)}
        for f in ("cat", "shop", "brand", "int"):
            ln = rng.randint(0, maxlen + 1)
            ex["u_%sids" % f] = rng.randint(1, N, ln).astype(np.int64)
            ex["u_%svals" % f] = (rng.rand(ln) * 3).astype(np.float32)
        recs.append(tfr.encode_example(ex))
    tfr.write_records(path, recs)


def test_din_input_pipeline(tmp_path):
    import torch
    from tf_repos_b200 import din_main as dm
    p = os.path.join(tmp_path, "a.tfrecord")
    _write_din(p, 37, seed=1)
    d = dm.decode_tfrecord_files([p], 11)
    assert len(d["y"]) == 37 and d["feat_ids"][0].shape == (11,)
    P, A = dm.max_lengths(d)
    assert 1 <= P <= 6 and 1 <= A <= 3
    assert [len(b) for b in dm.index_stream(37, 2, 16)] == [16, 16, 16, 16, 10]     # repeat before batch
    stream = list(dm.index_stream(37, 2, 16))
    assert stream[2][:6] == [32, 33, 34, 35, 36, 0]                                   # straddles the epoch boundary
    batch, labels, n = dm.make_batch(d, list(range(5)), 8, P, "cpu")
    assert n == 5 and batch["feat_ids"].shape == (8, 11) and batch["u_ids"].shape == (4, 8, P) and labels.shape == (8,)
    assert batch["feat_ids"].dtype == torch.int32 and batch["a_ids"].shape == (3, 8)
    for b in range(5):
        ids = d["u_shopids"][b]
        assert batch["u_ids"][1, b, :len(ids)].tolist() == ids.tolist() and not batch["u_ids"][1, b, len(ids):].any()
        assert np.allclose(batch["u_wgt"][1, b, :len(ids)].numpy(), d["u_shopvals"][b])
        lo, hi = int(batch["a_int_off"][b]), int(batch["a_int_off"][b + 1])
        assert batch["a_int_ids"][lo:hi].tolist() == d["a_int"][b].tolist()
    assert torch.equal(batch["feat_ids"][5], batch["feat_ids"][0])                    # padding = copies of sample 0
    with pytest.raises(ValueError):
        dm.decode_tfrecord_files([p], 12)                                             # field_size mismatch
