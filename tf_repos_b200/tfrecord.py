"""TFRecord files of tf.train.Example protos, read and written without TensorFlow (DIN's input format:
deep_ctr/Model_pipeline/DIN.py:57-99, `tf.data.TFRecordDataset(...).map(tf.parse_single_example)`).

Record framing [TF-sem, tensorflow/core/lib/io/record_writer.cc]:
    uint64 length (LE) | uint32 masked_crc32c(length bytes) | data[length] | uint32 masked_crc32c(data)
    masked_crc = ((crc >> 15) | (crc << 17)) + 0xa282ead8   (mod 2^32), crc = CRC-32C (Castagnoli)
Example wire format [TF-sem, tensorflow/core/example/{example,feature}.proto]:
    Example  { Features features = 1; }
    Features { map<string, Feature> feature = 1; }        // map entry: key = 1 (string), value = 2 (Feature)
    Feature  { oneof kind { BytesList bytes_list = 1; FloatList float_list = 2; Int64List int64_list = 3; } }
    FloatList { repeated float value = 1 [packed = true]; }   Int64List { repeated int64 value = 1 [packed = true]; }
Both packed and unpacked encodings of the repeated fields are accepted on read (protobuf parsers must).
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Tuple, Union

import numpy as np

# ---- CRC-32C -------------------------------------------------------------------------------------------------
_POLY = 0x82F63B78
_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ _POLY if _c & 1 else _c >> 1
    _TABLE.append(_c)


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- record framing ------------------------------------------------------------------------------------------
def read_records(path: str, verify_crc: bool = False) -> Iterator[bytes]:
    with open(path, "rb") as fh:
        buf = fh.read()
    pos, n = 0, len(buf)
    while pos < n:
        if pos + 12 > n:
            raise ValueError(f"{path}: truncated record header at byte {pos}")
        (length,) = struct.unpack_from("<Q", buf, pos)
        if verify_crc and struct.unpack_from("<I", buf, pos + 8)[0] != masked_crc(buf[pos:pos + 8]):
            raise ValueError(f"{path}: corrupted record length at byte {pos}")
        start, end = pos + 12, pos + 12 + length
        if end + 4 > n:
            raise ValueError(f"{path}: truncated record at byte {pos}")
        data = buf[start:end]
        if verify_crc and struct.unpack_from("<I", buf, end)[0] != masked_crc(data):
            raise ValueError(f"{path}: corrupted record data at byte {pos}")
        yield data
        pos = end + 4


def write_records(path: str, records) -> None:
    with open(path, "wb") as fh:
        for data in records:
            hdr = struct.pack("<Q", len(data))
            fh.write(hdr); fh.write(struct.pack("<I", masked_crc(hdr)))
            fh.write(data); fh.write(struct.pack("<I", masked_crc(data)))


# ---- protobuf wire format --------------------------------------------------------------------------------------
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    r, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        r |= (b & 0x7F) << shift
        if not b & 0x80:
            return r, pos
        shift += 7


def _fields(buf: bytes) -> Iterator[Tuple[int, int, Union[int, bytes]]]:
    """(field number, wire type, value) of one message level"""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, v


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_feature(buf: bytes):
    for num, wt, v in _fields(buf):
        if num == 2:      # FloatList
            vals: List[float] = []
            for n2, w2, x in _fields(v):
                if n2 == 1 and w2 == 2:
                    vals.extend(np.frombuffer(x, dtype="<f4").tolist())
                elif n2 == 1 and w2 == 5:
                    vals.append(struct.unpack("<f", x)[0])
            return np.asarray(vals, dtype=np.float32)
        if num == 3:      # Int64List
            ints: List[int] = []
            for n2, w2, x in _fields(v):
                if n2 == 1 and w2 == 2:
                    p = 0
                    while p < len(x):
                        t, p = _varint(x, p)
                        ints.append(_signed64(t))
                elif n2 == 1 and w2 == 0:
                    ints.append(_signed64(x))
            return np.asarray(ints, dtype=np.int64)
        if num == 1:      # BytesList
            return [x for n2, w2, x in _fields(v) if n2 == 1]
    return np.asarray([], dtype=np.float32)   # a Feature with no kind set


def parse_example(data: bytes) -> Dict[str, Union[np.ndarray, list]]:
    out: Dict[str, Union[np.ndarray, list]] = {}
    for num, wt, v in _fields(data):
        if num != 1:
            continue
        for n2, w2, entry in _fields(v):          # Features.feature map entries
            if n2 != 1:
                continue
            key, feat = None, b""
            for n3, w3, x in _fields(entry):
                if n3 == 1:
                    key = x.decode("utf-8")
                elif n3 == 2:
                    feat = x
            if key is not None:
                out[key] = _parse_feature(feat)
    return out


def _enc_varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(num: int, payload: bytes) -> bytes:
    return _enc_varint((num << 3) | 2) + _enc_varint(len(payload)) + payload


def encode_example(features: Dict[str, Union[np.ndarray, list, float, int]]) -> bytes:
    """float32 arrays/scalars -> FloatList, integer arrays/scalars -> Int64List (packed), like tf.train.Example"""
    entries = b""
    for key in sorted(features):
        v = np.atleast_1d(np.asarray(features[key]))
        if v.dtype.kind == "f":
            feat = _ld(2, _ld(1, v.astype("<f4").tobytes()) if v.size else b"")
        else:
            feat = _ld(3, _ld(1, b"".join(_enc_varint(int(x)) for x in v)) if v.size else b"")
        entries += _ld(1, _ld(1, key.encode()) + _ld(2, feat))
    return _ld(1, entries)
