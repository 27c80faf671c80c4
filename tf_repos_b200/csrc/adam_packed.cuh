// adam_packed.cuh -- the untouched-row Adam step (g = l2*var) on PAIRS of elements with the
// sm_100 packed fp32 instructions (FMUL2 / FADD2 / FFMA2: two IEEE-rounded lanes per issue slot), for the
// epoch sweeps (epoch_adam.cu).  Every lane performs exactly the operations of step_sparse<ADAM>
// (optim_steps.cuh) in the same order with the same roundings, so the result is bit-identical; what changes
// is the number of issue slots: 13.5 per element-step instead of 29.5, which moves the sweep's limiter from
// instruction issue to the MUFU pipe (2 MUFU per element-step: 16 lanes/clk/SM).
//
// Three things make that possible:
//  * mul.rn.f32x2 followed by add.rn.f32x2 IS contracted by ptxas 12.9 into FFMA2 (unlike the scalar .rn
//    forms; -fmad=false does not stop it).  A product that feeds an add is therefore written
//    fma(a, b, nz) with nz a RUN-TIME -0.0f (kernel argument): RN(a*b + (-0)) == RN(a*b) for every a*b
//    including both zeros, and the following add has no multiply left to absorb.
//  * no range check, branch or select inside the step loop.  The IEEE sqrt/div fast paths
//    (sqrt_rn_inrange / div_rn_inrange) are exact on a range; instead of testing it per element and step,
//    the loop keeps two running min/max trackers (one FMNMX3 per pair each) and the caller validates the
//    whole trajectory AFTERWARDS.  If validation fails, nothing has been stored: the caller reloads the
//    state and replays it with the checked scalar path (adam_untouched, optim_steps.cuh).
//  * the numerator is carried negated (na = -lr_t*m; RN arithmetic is odd-symmetric), so the final
//    var - q is a plain FADD2 / FFMA2 without a negation of a register pair.
//
// MODE 0 ("A"): every |lr_t*m| in [2^-100, ..): the plain fast paths.
// MODE 1 ("S1"): tiny or zero numerators (|lr_t*m| <= 2^-37) over a small denominator (v <= 2^-52,
//   2^-40 <= eps <= 2^-26): (a*2^64)/b is inside the divider's exact range and scaling its correctly
//   rounded quotient back by 2^-64 is exact because |a/b| >= 2^-149/2^-25 is a normal number.  The
//   sign of a zero numerator is restored with one LOP3 (copysign) -- it decides -0 - (-0) = +0.
//   This is the state rows nothing gathers park in (l2 + Adam pull them to ~FLT_MIN, m underflows).
// MODE 2 ("S2"): as S1, and the second moment may be denormal or zero (v < 2^-101, reached after ~26 k
//   steps): sqrt(v*2^48)*2^-24, clamped below at 2^-101 (sqrt(0)+eps == eps for eps >= 2^-40).
#pragma once
#include "optim_steps.cuh"

namespace ctr {

struct AdamPk {
  float l2, b1, b2, omb1, omb2, eps, nz;
};

__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float copysign_bits(float mag, float sgn) {
  return __uint_as_float((__float_as_uint(mag) & 0x7fffffffu) | (__float_as_uint(sgn) & 0x80000000u));
}

constexpr float PK_SC64 = 18446744073709551616.f /* 2^64 */, PK_ISC64 = 5.421010862427522e-20f /* 2^-64 */;
constexpr float PK_SC48 = 281474976710656.f /* 2^48 */, PK_ISC24 = 5.9604644775390625e-8f /* 2^-24 */;
constexpr float PK_S_A_HI = 7.2759576e-12f /* 2^-37 */, PK_S_V_HI = 2.2204460e-16f /* 2^-52 */;
constexpr float PK_S_EPS_LO = 9.0949470e-13f /* 2^-40 */, PK_S_EPS_HI = 1.4901161e-8f /* 2^-26 */;

// One step for one pair.  nlr = -lr_t.  trk_a: running min (MODE 0) / max (MODE 1,2) of |lr_t*m|.
// The second moment needs no per-step tracker: v' = b2*v + (1-b2)*g^2 >= RN(b2*v), so every v_s is bounded below
// by the FIRST updated value times b2^(s-1) and above by the LAST one divided by b2^(n-s) (pk_valid).
template <int MODE>
__device__ __forceinline__ void adam_pk_step(float2& x, float2& m, float2& v, float nlr, const AdamPk& c,
                                             float& trk_a) {
  const float2 NZ = bc2(c.nz);
  const float2 g = __fmul2_rn(bc2(c.l2), x);
  m = __fadd2_rn(__ffma2_rn(m, bc2(c.b1), NZ), __ffma2_rn(g, bc2(c.omb1), NZ));
  v = __fadd2_rn(__ffma2_rn(v, bc2(c.b2), NZ), __ffma2_rn(__fmul2_rn(g, g), bc2(c.omb2), NZ));
  const float2 na = __fmul2_rn(bc2(nlr), m);
  float2 b;
  if (MODE == 2) {
    float2 vs = __fmul2_rn(v, bc2(PK_SC48));
    vs.x = fmaxf(vs.x, SQRT_LO); vs.y = fmaxf(vs.y, SQRT_LO);
    const float2 r = make_float2(mufu_rsq(vs.x), mufu_rsq(vs.y));
    const float2 y = __fmul2_rn(vs, r);
    const float2 hh = __fmul2_rn(r, bc2(0.5f));
    const float2 e = __ffma2_rn(neg2(y), y, vs);
    const float2 sq = __ffma2_rn(e, hh, y);
    b = __ffma2_rn(sq, bc2(PK_ISC24), bc2(c.eps));   // sq*2^-24 is exact: one rounding, that of sqrt(v)+eps
  } else {
    const float2 r = make_float2(mufu_rsq(v.x), mufu_rsq(v.y));
    const float2 y = __fmul2_rn(v, r);
    const float2 hh = __fmul2_rn(r, bc2(0.5f));
    const float2 e = __ffma2_rn(neg2(y), y, v);
    const float2 sq = __ffma2_rn(e, hh, y);
    b = __fadd2_rn(sq, bc2(c.eps));
  }
  float2 rc = make_float2(mufu_rcp(b.x), mufu_rcp(b.y));
  const float2 e2 = __ffma2_rn(neg2(b), rc, bc2(1.f));
  rc = __ffma2_rn(rc, e2, rc);
  if (MODE == 0) {
    const float2 q = __ffma2_rn(na, rc, bc2(0.f));
    const float2 t = __ffma2_rn(neg2(b), q, na);
    const float2 res = __ffma2_rn(rc, t, q);        // == -(lr_t*m / b), correctly rounded
    x = __fadd2_rn(x, res);
    trk_a = fminf(fminf(trk_a, fabsf(na.x)), fabsf(na.y));
  } else {
    const float2 nas = __fmul2_rn(na, bc2(PK_SC64));
    const float2 q = __ffma2_rn(nas, rc, bc2(0.f));
    const float2 t = __ffma2_rn(neg2(b), q, nas);
    float2 qs = __ffma2_rn(rc, t, q);                // == -(lr_t*m / b) * 2^64; a zero lost its sign
    qs.x = copysign_bits(qs.x, na.x); qs.y = copysign_bits(qs.y, na.y);
    x = __ffma2_rn(qs, bc2(PK_ISC64), x);            // qs*2^-64 is exact: one rounding, that of var - q
    trk_a = fmaxf(fmaxf(trk_a, fabsf(na.x)), fabsf(na.y));
  }
}

// NP pairs, steps [s0, s1) with nlr_s[s] = -lr_t of step s; q2[s-s0]... the caller accumulates sum(var^2)
// through `ssq(s, value)`.
struct PkTrackers {
  float a;    // running min (MODE 0) / max (MODE 1, 2) of |lr_t*m|
  float v1;   // min over the elements of the second moment after the FIRST step (set by the caller)
};
template <int MODE> __device__ __forceinline__ PkTrackers pk_trackers_init() {
  PkTrackers t;
  t.a = (MODE == 0) ? 3.0e38f : 0.f;
  t.v1 = 3.0e38f;
  return t;
}

// post-hoc validity of a whole trajectory (see the header comment).  vfin_max/xfin: max of the final v and
// "all final x finite"; b2n = b2^n_steps.
template <int MODE>
__device__ __forceinline__ bool pk_valid(const PkTrackers& t, float vfin_max, bool x_finite, float b2n) {
  // v_s <= v_final / b2^(n-s) (the second moment cannot fall faster than b2 per step): bound every v_s from the last one
  const float vhi = (MODE == 0 ? SQRT_HI : PK_S_V_HI) * b2n * 0.99f;
  bool ok = x_finite && vfin_max <= vhi;
  // v_s >= v_1 * b2^(s-1) * (1 - 2^-24)^(s-1): demand v_1 * b2^n >= 1.01 * 2^-101
  const bool v_lo_ok = t.v1 * b2n >= SQRT_LO * 1.01f;
  if (MODE == 0) ok = ok && t.a >= DIV_LO && v_lo_ok;
  if (MODE == 1) ok = ok && t.a <= PK_S_A_HI && v_lo_ok;
  if (MODE == 2) ok = ok && t.a <= PK_S_A_HI;
  return ok;
}

// static preconditions on the hyper-parameters for the packed loops
__device__ __forceinline__ bool pk_hyper_ok(const Hyper& h, bool scaled) {
  bool ok = h.b2 > 0.f && h.b2 < 1.f && h.b1 >= 0.f && h.b1 < 1.f && h.l2 >= 0.f;
  if (scaled) ok = ok && h.eps >= PK_S_EPS_LO && h.eps <= PK_S_EPS_HI;
  else ok = ok && h.eps >= 0.f && h.eps <= 524288.f;
  return ok;
}

}  // namespace ctr
