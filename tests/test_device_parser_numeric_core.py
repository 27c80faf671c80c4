"""The numeric core of the GPU libsvm tokenizer (csrc/libsvm_device.cu::parse_float), restated in Python and checked on
the CPU against EXACT rational arithmetic: whenever the fast path accepts a decimal, fp32(RN_double(m / 10^k)) must be the
correctly rounded fp32 of the decimal value (what strtof returns); decimals chosen next to fp32 rounding boundaries, where
the double rounding could bite, must be declined (or still correct).  The kernel itself is compared with strtof on the
GPU in tests/test_gpu_libsvm.py; this test pins the reasoning behind its guard band."""
import random
import struct
from fractions import Fraction

import numpy as np

POW10 = [10.0 ** k for k in range(23)]          # exact doubles up to 1e22
FLT_MIN, FLT_MAX = 1.1754943508222875e-38, 3.4028234663852886e38


def device_parse(sig_digits: int, m: int, exp10: int):
    """Returns ('ok', float32) | ('host', None): the decision and value of parse_float for mantissa m (sig_digits
    significant digits) and decimal exponent exp10."""
    if m == 0:
        return "ok", np.float32(0.0)
    if sig_digits > 15 or exp10 < -22 or exp10 > 22:
        return "host", None
    d = float(m) / POW10[-exp10] if exp10 < 0 else float(m) * POW10[exp10]        # ONE correctly rounded double op
    if not (FLT_MIN <= d <= FLT_MAX):
        return "host", None
    low = struct.unpack("<Q", struct.pack("<d", d))[0] & 0x1FFFFFFF
    if 0x0FFFFFFF <= low <= 0x10000001:
        return "host", None
    return "ok", np.float32(d)


def exact_f32(fr: Fraction) -> np.float32:
    """correctly rounded (nearest-even) float32 of a positive rational inside the normal range"""
    x = np.float32(float(fr))
    best = None
    for c in (x, np.nextafter(x, np.float32(np.inf)), np.nextafter(x, np.float32(0))):
        err = abs(Fraction(float(c)) - fr)
        even = (struct.unpack("<I", struct.pack("<f", float(c)))[0] & 1) == 0
        key = (err, 0 if even else 1)
        if best is None or key < best[0]:
            best = (key, c)
    return np.float32(best[1])


def test_accepted_decimals_are_correctly_rounded_random():
    rng = random.Random(3)
    accepted = 0
    for _ in range(40000):
        nd = rng.randint(1, 15)
        m = rng.randint(10 ** (nd - 1), 10 ** nd - 1)
        exp10 = rng.randint(-22, 22 - 0)
        st, val = device_parse(nd, m, exp10)
        if st != "ok":
            continue
        accepted += 1
        want = exact_f32(Fraction(m) * Fraction(10) ** exp10)
        assert val.tobytes() == want.tobytes(), (m, exp10, val, want)
    assert accepted > 20000


def test_decimals_next_to_fp32_rounding_boundaries():
    """15-digit decimals within 1e-15 (relative) of the midpoint between two adjacent fp32 values: the double nearest to
    them is usually the midpoint itself, so fp32(double) would round half-to-even blindly.  They must be declined, or be right."""
    rng = random.Random(5)
    declined = 0
    for _ in range(4000):
        bits = rng.randint(0x20000000, 0x5F000000)                    # ~1e-19 .. 1e19
        lo = struct.unpack("<f", struct.pack("<I", bits))[0]
        hi = struct.unpack("<f", struct.pack("<I", bits + 1))[0]
        mid = (Fraction(lo) + Fraction(hi)) / 2
        # a 15-significant-digit decimal just below / above the midpoint
        e = len(str(int(mid))) if mid >= 1 else -len(str(int(1 / mid))) + 1
        scale = Fraction(10) ** (15 - e)
        base = int(mid * scale)
        for m in (base, base + 1):
            nd = len(str(m))
            exp10 = -(15 - e)
            if nd > 15 or not (-22 <= exp10 <= 22):
                continue
            st, val = device_parse(nd, m, exp10)
            if st == "host":
                declined += 1
                continue
            want = exact_f32(Fraction(m) * Fraction(10) ** exp10)
            assert val.tobytes() == want.tobytes(), (m, exp10, val, want)
    assert declined > 100          # the guard band is what protects these cases
