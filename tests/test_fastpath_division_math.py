"""The arithmetic behind the epoch sweep's division fast path (csrc/optim_steps.cuh::div_rn_inrange), emulated on the
CPU with exact rationals: r = refine(seed 1/b); q = RN(a*r); t = RN(a - b*q) [one FMA]; result = RN(q + r*t) [one FMA].
With a correctly rounded seed the sequence returns the correctly rounded quotient over the WHOLE guarded range
(|a| in [2^-100, 2^60), b in [2^-51, 2^21)) -- in particular the FMA residual stays exactly representable down to
|a| = 2^-100 -- and the power-of-two scaling of the packed sweep's S loops (csrc/adam_packed.cuh) extends it to denormal numerators.  (How the hardware's MUFU.RCP seed
behaves is what ctr_selftest_divsqrt checks on the GPU; this test pins the range reasoning.)"""
import random
import struct
from fractions import Fraction

import numpy as np

f32 = np.float32


def rn32(fr: Fraction) -> np.float32:
    if fr == 0:
        return f32(0.0)
    x = f32(float(fr))
    best = None
    for c in (x, np.nextafter(x, f32(np.inf)), np.nextafter(x, f32(-np.inf))):
        if not np.isfinite(c):
            continue
        err = abs(Fraction(float(c)) - fr)
        even = (struct.unpack("<I", struct.pack("<f", float(c)))[0] & 1) == 0
        key = (err, 0 if even else 1)
        if best is None or key < best[0]:
            best = (key, c)
    return f32(best[1])


def fma(a, b, c):
    return rn32(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def div_fast(a, b):
    r = rn32(Fraction(1) / Fraction(float(b)))            # seed: correctly rounded reciprocal
    e = fma(-b, r, f32(1.0))
    r = fma(r, e, r)
    q = fma(a, r, f32(0.0))
    t = fma(-b, q, a)
    return fma(r, t, q)


def rand_f32(rng, e_lo, e_hi):
    e = rng.randint(e_lo, e_hi)
    man = rng.getrandbits(23)
    k = rng.randrange(8)
    if k == 0:
        man = 0x7FFFFF
    elif k == 1:
        man = 0
    elif k == 2:
        man = 1
    return f32(struct.unpack("<f", struct.pack("<I", ((e + 127) << 23) | man))[0])


def test_fast_division_is_correctly_rounded_over_the_guarded_range():
    rng = random.Random(1)
    for i in range(5000):
        a = rand_f32(rng, -100, -56) if i % 2 else rand_f32(rng, -100, 59)
        if rng.random() < 0.5:
            a = -a
        b = rand_f32(rng, -51, 20)
        want = rn32(Fraction(float(a)) / Fraction(float(b)))
        got = div_fast(a, b)
        assert want.tobytes() == got.tobytes(), (a, b, want, got)
    for z in (f32(0.0), f32(-0.0)):                        # a zero numerator gives a zero
        assert div_fast(z, f32(3e-7)) == 0.0


def test_scaled_division_handles_denormal_numerators():
    """(a * 2^64) / b * 2^-64 for denormal / tiny a and b in [2^-50, 2^-25): exact scaling, normal quotient."""
    rng = random.Random(2)
    two64, twom64 = f32(2.0 ** 64), f32(2.0 ** -64)
    for _ in range(3000):
        if rng.random() < 0.75:
            a = f32(struct.unpack("<f", struct.pack("<I", rng.randint(1, 0x7FFFFF)))[0])          # denormal
        else:
            a = rand_f32(rng, -126, -101)
        if rng.random() < 0.5:
            a = -a
        b = rand_f32(rng, -50, -26)
        a_s = rn32(Fraction(float(a)) * Fraction(2) ** 64)
        assert Fraction(float(a_s)) == Fraction(float(a)) * Fraction(2) ** 64                   # the up-scaling is exact
        q = rn32(Fraction(float(div_fast(a_s, b))) * Fraction(1, 2 ** 64))
        want = rn32(Fraction(float(a)) / Fraction(float(b)))
        assert abs(float(want)) >= 2.0 ** -126 and want.tobytes() == q.tobytes(), (a, b, want, q)


def test_scaled_sqrt_handles_denormal_and_zero_second_moments():
    """Loop S2 of the packed sweep (csrc/adam_packed.cuh): sqrt(v) for a denormal v is taken as sqrt(v*2^48)*2^-24 --
    the up-scaling is exact, the correctly rounded root of the scaled operand is a normal number and the down-scaling
    is exact -- and v == 0 is clamped to 2^-101 before the root: sqrt(0)+eps must still be eps for eps >= 2^-40."""
    rng = random.Random(3)
    for _ in range(4000):
        bits = rng.randint(1, 0x7FFFFF) >> rng.randint(0, 22)              # denormals of every magnitude
        v = np.array([max(bits, 1)], dtype=np.uint32).view(np.float32)[0]
        want = np.sqrt(v)                                                   # numpy's float32 sqrt is IEEE (correctly rounded)
        vs = np.float32(float(v) * 2.0 ** 48)
        assert float(vs) == float(v) * 2.0 ** 48 and float(vs) >= 2.0 ** -101
        got = np.float32(float(np.sqrt(vs)) * 2.0 ** -24)
        assert got.tobytes() == want.tobytes(), (v, got, want)
    for eps in (np.float32(1e-8), np.float32(2.0 ** -40), np.float32(2.0 ** -26)):
        clamp_root = np.float32(float(np.sqrt(np.float32(2.0 ** -101))) * 2.0 ** -24)
        assert np.float32(clamp_root + eps).tobytes() == eps.tobytes()       # == sqrt(0) + eps
