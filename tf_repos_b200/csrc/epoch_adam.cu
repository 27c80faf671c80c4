// epoch_adam.cu -- the Adam epoch sweep of the exact-deferred update (epoch.cu) on the packed fp32 pipe.
//
// Replaces (bit for bit) what epoch_sweep_kernel<ADAM> does for rows nothing gathered since `from`:
// replay the untouched-row Adam step (g = l2*var; DeepFM.py:189-190,205 [TF-sem]) for steps from..upto-1
// in registers, one pass over HBM.  Differences in structure:
//   * the step loop is adam_pk_step (adam_packed.cuh): FMUL2/FADD2/FFMA2, no range check / branch / select
//     per step; the trajectory is validated afterwards and replayed by the checked scalar path if it left
//     the exact range of the IEEE fast paths (nothing has been stored at that point);
//   * rows a batch gathered since `from` (last[row] > from) are NOT handled here: their ids are appended
//     to `list` and caught up by ctr_epoch_rows(apply=0, j=upto) right after this kernel (they are ~1.6 %
//     of the rows but sit in ~25 % of the warps; select-masking them costs more than a second small pass).
//     Their register slots are filled with a copy of a neighbouring untouched row so that the packed loop
//     and its trackers only ever see untouched-row state;
//   * this kernel never writes `last` (rows may span warps / CTAs when K/4 is not a power of two <= 32):
//     after an epoch-end sweep (from == 0, reset) untouched rows already hold 0; otherwise
//     epoch_last_kernel rewrites the bytes in a separate pass.
//
// Bound: 2 MUFU per element-step at 16 MUFU lanes/clk/SM => 22.8 ms per 16-step pass over 3.2e9 elements
// at 1.9 GHz; HBM traffic stays 24 B/element per pass (11.7 ms at the measured peak).
#include <stdlib.h>

#include "adam_packed.cuh"

namespace ctr {

constexpr int EPOCH_MAX_A = 32;
constexpr int SWEEP_THREADS = 256;

// dynamic shared memory: nlr[32] | ss_thr[nsteps][256] | ss_tmp[nsteps][256]
struct SweepSmem {
  float* nlr;
  float* ss_thr;
  float* ss_tmp;
};
__device__ __forceinline__ SweepSmem sweep_smem(float* base, int nsteps) {
  SweepSmem s;
  s.nlr = base;
  s.ss_thr = base + EPOCH_MAX_A;
  s.ss_tmp = s.ss_thr + nsteps * SWEEP_THREADS;
  return s;
}

// NP pairs through steps [from, upto) on the packed pipe.  Returns false if the trajectory left the exact range
// (state and ss_tmp are then garbage: the caller reloads).  wq: weight of this thread's sum(var^2) (dummy slots).
template <int MODE, int NP, bool WEIGHTED = false>
__device__ __forceinline__ bool pk_run(float2 (&x)[NP], float2 (&m)[NP], float2 (&v)[NP], const AdamPk& c,
                                       const SweepSmem& sm, int from, int upto, float b2n,
                                       const float2* w = nullptr) {
  PkTrackers t = pk_trackers_init<MODE>();
  // first step apart: its second moments give the lower bound of the whole trajectory's (pk_valid)
  {
    const float nlr = sm.nlr[from];
    float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      q2 = __ffma2_rn(WEIGHTED ? __fmul2_rn(x[p], w[p]) : x[p], x[p], q2);
      adam_pk_step<MODE>(x[p], m[p], v[p], nlr, c, t.a);
      t.v1 = fminf(fminf(t.v1, v[p].x), v[p].y);
    }
    sm.ss_tmp[threadIdx.x] = q2.x + q2.y;
  }
#pragma unroll 2
  for (int s = from + 1; s < upto; ++s) {
    const float nlr = sm.nlr[s];
    float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      q2 = __ffma2_rn(WEIGHTED ? __fmul2_rn(x[p], w[p]) : x[p], x[p], q2);
      adam_pk_step<MODE>(x[p], m[p], v[p], nlr, c, t.a);
    }
    sm.ss_tmp[(s - from) * SWEEP_THREADS + threadIdx.x] = q2.x + q2.y;
  }
  float vmax = 0.f, fin = 0.f;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    vmax = fmaxf(fmaxf(vmax, v[p].x), v[p].y);
    fin += (x[p].x - x[p].x) + (x[p].y - x[p].y);   // 0 iff both finite
  }
  return pk_valid<MODE>(t, vmax, fin == 0.f, b2n);
}

// regime guess from an estimate of the first step's numerator and second moment (heuristic only: validity
// is established after the fact by pk_valid)
template <int NP>
__device__ __forceinline__ int pk_guess(const float2 (&x)[NP], const float2 (&m)[NP], const float2 (&v)[NP],
                                        const AdamPk& c, float lr0, bool okA, bool okS) {
  float emin = 3.0e38f, emax = 0.f, vmin = 3.0e38f;
  const float gl = c.omb1 * c.l2, gv = c.omb2 * c.l2 * c.l2;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const float ex = fabsf(fmaf(c.b1, m[p].x, gl * x[p].x)) * lr0, ey = fabsf(fmaf(c.b1, m[p].y, gl * x[p].y)) * lr0;
    emin = fminf(fminf(emin, ex), ey); emax = fmaxf(fmaxf(emax, ex), ey);
    const float vx = fmaf(c.b2, v[p].x, gv * x[p].x * x[p].x), vy = fmaf(c.b2, v[p].y, gv * x[p].y * x[p].y);
    vmin = fminf(fminf(vmin, vx), vy);
  }
  if (okA && emin >= 8.0779357e-28f /* 2^-90 */ && vmin >= 2.5243549e-29f /* 2^-95 */) return 0;
  if (okS && emax <= 9.0949470e-13f /* 2^-40 */) return vmin >= 2.5243549e-29f ? 1 : 2;
  return 3;
}

// ---- K % 4 == 0: a row is K/4 consecutive float4 ------------------------------------------------------------
// PREFETCH: the next grid-stride iteration's 6 float4 + `last` bytes are requested before this iteration's step loop
// (26 more live registers: use with MINB = 2), so no warp waits on HBM between iterations.
template <int MINB, bool PREFETCH = false>
__global__ void __launch_bounds__(SWEEP_THREADS, MINB)
epoch_sweep_adam_kernel(float* __restrict__ var, float* __restrict__ slot0, float* __restrict__ slot1,
                        const uint8_t* __restrict__ last, int64_t n4, int f4_per_row, int sh,
                        const float* __restrict__ hyper, const float* __restrict__ lr_table, int from, int upto,
                        double* __restrict__ ss_partials, int n_partials, int32_t* __restrict__ list,
                        int32_t* __restrict__ list_count, int64_t list_cap, float nz) {
  constexpr int U = 2, NP = 2 * U;
  extern __shared__ float smem_dyn[];
  const int nsteps = upto - from;
  const SweepSmem sm = sweep_smem(smem_dyn, nsteps);
  if (threadIdx.x < EPOCH_MAX_A) sm.nlr[threadIdx.x] = (threadIdx.x < upto) ? -lr_table[threadIdx.x] : 0.f;
  for (int s = 0; s < nsteps; ++s) sm.ss_thr[s * SWEEP_THREADS + threadIdx.x] = 0.f;
  __syncthreads();
  const Hyper h0 = load_hyper(hyper);
  AdamPk c;
  c.l2 = h0.l2; c.b1 = h0.b1; c.b2 = h0.b2; c.omb1 = __fsub_rn(1.f, h0.b1); c.omb2 = __fsub_rn(1.f, h0.b2);
  c.eps = h0.eps; c.nz = nz;
  const bool okA = pk_hyper_ok(h0, false), okS = pk_hyper_ok(h0, true);
  float b2n = 1.f;
  for (int s = from; s < upto; ++s) b2n *= h0.b2;
  const float lr0 = fabsf(sm.nlr[from]);
  const AdamConsts ac = adam_consts(h0);
  float4* v4 = reinterpret_cast<float4*>(var);
  float4* a4 = reinterpret_cast<float4*>(slot0);
  float4* b4 = reinterpret_cast<float4*>(slot1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  struct Raw { float4 X[U], M[U], V[U]; int l0[U]; };
  auto load_raw = [&](int64_t j0, Raw& r) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = j0 + u * stride;
      r.l0[u] = -1;                                  // out of range
      if (i < n4) {
        r.X[u] = ld_stream4(v4 + i); r.M[u] = ld_stream4(a4 + i); r.V[u] = ld_stream4(b4 + i);
        r.l0[u] = last[sh >= 0 ? (i >> sh) : (i / f4_per_row)];
      }
    }
  };
  Raw cur, nxt;
  const int64_t i_first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (PREFETCH && i_first < n4) load_raw(i_first, cur);
  for (int64_t i0 = i_first; i0 < n4; i0 += U * stride) {
    if (!PREFETCH) load_raw(i0, cur);
    else if (i0 + U * stride < n4) load_raw(i0 + U * stride, nxt);
    float2 x[NP], m[NP], v[NP];
    bool act[U];
    int nact = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      act[u] = false;
      if (cur.l0[u] >= 0) {
        x[2 * u] = make_float2(cur.X[u].x, cur.X[u].y); x[2 * u + 1] = make_float2(cur.X[u].z, cur.X[u].w);
        m[2 * u] = make_float2(cur.M[u].x, cur.M[u].y); m[2 * u + 1] = make_float2(cur.M[u].z, cur.M[u].w);
        v[2 * u] = make_float2(cur.V[u].x, cur.V[u].y); v[2 * u + 1] = make_float2(cur.V[u].z, cur.V[u].w);
        const int64_t row = sh >= 0 ? (i >> sh) : (i / f4_per_row);
        const int l0 = cur.l0[u];
        act[u] = l0 == from;
        if (l0 > from && row * f4_per_row == i) {   // gathered since `from`: second pass (head lane appends)
          const int pos = atomicAdd(list_count, 1);
          if (pos < list_cap) list[pos] = (int32_t)row;
        }
      }
      nact += act[u] ? 1 : 0;
    }
    if (PREFETCH) cur = nxt;
    if (nact == 0) continue;
    if (!act[0]) { x[0] = x[2]; x[1] = x[3]; m[0] = m[2]; m[1] = m[3]; v[0] = v[2]; v[1] = v[3]; }
    if (!act[1]) { x[2] = x[0]; x[3] = x[1]; m[2] = m[0]; m[3] = m[1]; v[2] = v[0]; v[3] = v[1]; }
    const float wq = nact == U ? 1.f : 0.5f;   // a dummy slot is an exact copy of the other one
    const int mode = pk_guess<NP>(x, m, v, c, lr0, okA, okS);
    bool done = false;
    if (mode == 0) done = pk_run<0, NP>(x, m, v, c, sm, from, upto, b2n);
    else if (mode == 1) done = pk_run<1, NP>(x, m, v, c, sm, from, upto, b2n);
    else if (mode == 2) done = pk_run<2, NP>(x, m, v, c, sm, from, upto, b2n);
    if (done) {
      for (int s = 0; s < nsteps; ++s)
        sm.ss_thr[s * SWEEP_THREADS + threadIdx.x] += wq * sm.ss_tmp[s * SWEEP_THREADS + threadIdx.x];
    } else {
      // checked scalar path (per-group range check, compiler's sqrt.rn/div.rn outside it), from the stored state
      float4 X[U], M[U], V[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + (act[u] ? u : (1 - u)) * stride;
        X[u] = ld_stream4(v4 + i); M[u] = ld_stream4(a4 + i); V[u] = ld_stream4(b4 + i);
      }
      Hyper h = h0;
#pragma unroll 1
      for (int s = from; s < upto; ++s) {
        h.lr = -sm.nlr[s];
        float q = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) q += (X[u].x * X[u].x + X[u].y * X[u].y) + (X[u].z * X[u].z + X[u].w * X[u].w);
        adam_untouched<U>(X, M, V, h, ac);
        sm.ss_thr[(s - from) * SWEEP_THREADS + threadIdx.x] += wq * q;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        x[2 * u] = make_float2(X[u].x, X[u].y); x[2 * u + 1] = make_float2(X[u].z, X[u].w);
        m[2 * u] = make_float2(M[u].x, M[u].y); m[2 * u + 1] = make_float2(M[u].z, M[u].w);
        v[2 * u] = make_float2(V[u].x, V[u].y); v[2 * u + 1] = make_float2(V[u].z, V[u].w);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (act[u]) {
        const int64_t i = i0 + u * stride;
        st_stream4(v4 + i, make_float4(x[2 * u].x, x[2 * u].y, x[2 * u + 1].x, x[2 * u + 1].y));
        st_stream4(a4 + i, make_float4(m[2 * u].x, m[2 * u].y, m[2 * u + 1].x, m[2 * u + 1].y));
        st_stream4(b4 + i, make_float4(v[2 * u].x, v[2 * u].y, v[2 * u + 1].x, v[2 * u + 1].y));
      }
    }
  }
  __syncthreads();
  // warp w reduces the per-thread accumulators of steps w, w+8, ... (fixed order => deterministic)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int s = warp; s < upto; s += 8) {
    double q = 0.0;
    if (s >= from) {
#pragma unroll
      for (int k = 0; k < 8; ++k) q += (double)sm.ss_thr[(s - from) * SWEEP_THREADS + lane + 32 * k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(FULL_MASK, q, o);
    }
    if (lane == 0) ss_partials[(int64_t)s * n_partials + blockIdx.x] = q;
  }
}

// ---- K == 1 (first-order weights): a float4 holds 4 rows, each with its own `last` byte ------------------------
template <int MINB, bool PREFETCH = false>
__global__ void __launch_bounds__(SWEEP_THREADS, MINB)
epoch_sweep_adam_k1_kernel(float* __restrict__ var, float* __restrict__ slot0, float* __restrict__ slot1,
                           const uint8_t* __restrict__ last, int64_t n4, const float* __restrict__ hyper,
                           const float* __restrict__ lr_table, int from, int upto, double* __restrict__ ss_partials,
                           int n_partials, int32_t* __restrict__ list, int32_t* __restrict__ list_count,
                           int64_t list_cap, float nz) {
  constexpr int U = 2, NP = 2 * U, NE = 4 * U;
  extern __shared__ float smem_dyn[];
  const int nsteps = upto - from;
  const SweepSmem sm = sweep_smem(smem_dyn, nsteps);
  if (threadIdx.x < EPOCH_MAX_A) sm.nlr[threadIdx.x] = (threadIdx.x < upto) ? -lr_table[threadIdx.x] : 0.f;
  for (int s = 0; s < nsteps; ++s) sm.ss_thr[s * SWEEP_THREADS + threadIdx.x] = 0.f;
  __syncthreads();
  const Hyper h0 = load_hyper(hyper);
  AdamPk c;
  c.l2 = h0.l2; c.b1 = h0.b1; c.b2 = h0.b2; c.omb1 = __fsub_rn(1.f, h0.b1); c.omb2 = __fsub_rn(1.f, h0.b2);
  c.eps = h0.eps; c.nz = nz;
  const bool okA = pk_hyper_ok(h0, false), okS = pk_hyper_ok(h0, true);
  float b2n = 1.f;
  for (int s = from; s < upto; ++s) b2n *= h0.b2;
  const float lr0 = fabsf(sm.nlr[from]);
  float4* v4 = reinterpret_cast<float4*>(var);
  float4* a4 = reinterpret_cast<float4*>(slot0);
  float4* b4 = reinterpret_cast<float4*>(slot1);
  const uint32_t* l4 = reinterpret_cast<const uint32_t*>(last);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  struct Raw { float4 X[U], M[U], V[U]; uint32_t lw[U]; bool in[U]; };
  auto load_raw = [&](int64_t j0, Raw& r) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = j0 + u * stride;
      r.in[u] = i < n4;
      r.lw[u] = 0xffffffffu;
      if (r.in[u]) { r.X[u] = ld_stream4(v4 + i); r.M[u] = ld_stream4(a4 + i); r.V[u] = ld_stream4(b4 + i); r.lw[u] = l4[i]; }
      else { r.X[u] = r.M[u] = r.V[u] = f4_zero(); }
    }
  };
  Raw cur, nxt;
  const int64_t i_first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (PREFETCH && i_first < n4) load_raw(i_first, cur);
  for (int64_t i0 = i_first; i0 < n4; i0 += U * stride) {
    if (!PREFETCH) load_raw(i0, cur);
    else if (i0 + U * stride < n4) load_raw(i0 + U * stride, nxt);
    float xe[NE], me[NE], ve[NE];
    unsigned actm = 0;            // bit e: element e replays from..upto-1 here
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      xe[4 * u] = cur.X[u].x; xe[4 * u + 1] = cur.X[u].y; xe[4 * u + 2] = cur.X[u].z; xe[4 * u + 3] = cur.X[u].w;
      me[4 * u] = cur.M[u].x; me[4 * u + 1] = cur.M[u].y; me[4 * u + 2] = cur.M[u].z; me[4 * u + 3] = cur.M[u].w;
      ve[4 * u] = cur.V[u].x; ve[4 * u + 1] = cur.V[u].y; ve[4 * u + 2] = cur.V[u].z; ve[4 * u + 3] = cur.V[u].w;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int l0 = (int)((cur.lw[u] >> (8 * e)) & 255u);
        if (cur.in[u] && l0 == from) actm |= 1u << (4 * u + e);
        if (cur.in[u] && l0 > from) {      // gathered since `from`: second pass
          const int pos = atomicAdd(list_count, 1);
          if (pos < list_cap) list[pos] = (int32_t)(4 * i + e);
        }
      }
    }
    if (PREFETCH) cur = nxt;
    if (actm == 0) continue;
    // donor for the inactive slots: the first active element (chained selects, highest index first)
    float xd = 0.f, md = 0.f, vd = 0.f;
#pragma unroll
    for (int e = NE - 1; e >= 0; --e) if (actm & (1u << e)) { xd = xe[e]; md = me[e]; vd = ve[e]; }
    float2 x[NP], m[NP], v[NP], w[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const bool a0 = actm & (1u << (2 * p)), a1 = actm & (1u << (2 * p + 1));
      w[p] = make_float2(a0 ? 1.f : 0.f, a1 ? 1.f : 0.f);   // dummy slots do not count in sum(var^2)
      x[p] = make_float2(a0 ? xe[2 * p] : xd, a1 ? xe[2 * p + 1] : xd);
      m[p] = make_float2(a0 ? me[2 * p] : md, a1 ? me[2 * p + 1] : md);
      v[p] = make_float2(a0 ? ve[2 * p] : vd, a1 ? ve[2 * p + 1] : vd);
    }
    const int mode = pk_guess<NP>(x, m, v, c, lr0, okA, okS);
    bool done = false;
    if (mode == 0) done = pk_run<0, NP, true>(x, m, v, c, sm, from, upto, b2n, w);
    else if (mode == 1) done = pk_run<1, NP, true>(x, m, v, c, sm, from, upto, b2n, w);
    else if (mode == 2) done = pk_run<2, NP, true>(x, m, v, c, sm, from, upto, b2n, w);
    if (done) {
      for (int s = 0; s < nsteps; ++s)
        sm.ss_thr[s * SWEEP_THREADS + threadIdx.x] += sm.ss_tmp[s * SWEEP_THREADS + threadIdx.x];
    } else {
      // rare (a trajectory outside the exact range): scalar replay of the active elements from the STORED state,
      // IEEE sqrt/div from the compiler
      float ye[NE], ne[NE], ue[NE];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * stride;
        float4 X = f4_zero(), M = f4_zero(), V = f4_zero();
        if (i < n4) { X = ld_stream4(v4 + i); M = ld_stream4(a4 + i); V = ld_stream4(b4 + i); }
        ye[4 * u] = X.x; ye[4 * u + 1] = X.y; ye[4 * u + 2] = X.z; ye[4 * u + 3] = X.w;
        ne[4 * u] = M.x; ne[4 * u + 1] = M.y; ne[4 * u + 2] = M.z; ne[4 * u + 3] = M.w;
        ue[4 * u] = V.x; ue[4 * u + 1] = V.y; ue[4 * u + 2] = V.z; ue[4 * u + 3] = V.w;
      }
      Hyper h = h0;
#pragma unroll 1
      for (int s = from; s < upto; ++s) {
        h.lr = -sm.nlr[s];
        float q = 0.f;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          if (actm & (1u << e)) { q += ye[e] * ye[e]; step_sparse<CTR_OPT_ADAM>(ye[e], ne[e], ue[e], __fmul_rn(h.l2, ye[e]), h); }
        }
        sm.ss_thr[(s - from) * SWEEP_THREADS + threadIdx.x] += q;
      }
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        x[p] = make_float2(ye[2 * p], ye[2 * p + 1]); m[p] = make_float2(ne[2 * p], ne[2 * p + 1]);
        v[p] = make_float2(ue[2 * p], ue[2 * p + 1]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      const unsigned am = (actm >> (4 * u)) & 15u;
      if (am) {
        // inactive elements keep their stored value (the second pass owns them): element-wise stores
        const float ox[4] = {x[2 * u].x, x[2 * u].y, x[2 * u + 1].x, x[2 * u + 1].y};
        const float om[4] = {m[2 * u].x, m[2 * u].y, m[2 * u + 1].x, m[2 * u + 1].y};
        const float ov[4] = {v[2 * u].x, v[2 * u].y, v[2 * u + 1].x, v[2 * u + 1].y};
        if (am == 15u) {
          st_stream4(v4 + i, make_float4(ox[0], ox[1], ox[2], ox[3]));
          st_stream4(a4 + i, make_float4(om[0], om[1], om[2], om[3]));
          st_stream4(b4 + i, make_float4(ov[0], ov[1], ov[2], ov[3]));
        } else {
          float* px = var + 4 * i; float* pm = slot0 + 4 * i; float* pv = slot1 + 4 * i;
#pragma unroll
          for (int e = 0; e < 4; ++e) if (am & (1u << e)) { px[e] = ox[e]; pm[e] = om[e]; pv[e] = ov[e]; }
        }
      }
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int s = warp; s < upto; s += 8) {
    double q = 0.0;
    if (s >= from) {
#pragma unroll
      for (int k = 0; k < 8; ++k) q += (double)sm.ss_thr[(s - from) * SWEEP_THREADS + lane + 32 * k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(FULL_MASK, q, o);
    }
    if (lane == 0) ss_partials[(int64_t)s * n_partials + blockIdx.x] = q;
  }
}

// ---- self-test: packed loops vs the scalar step, bit for bit ------------------------------------------------
__device__ __forceinline__ uint64_t st_mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// float with a uniformly drawn exponent in [e_lo, e_hi] (e_lo < -126 => denormals, drawn as raw mantissas) and
// a random or stress-pattern mantissa; sign from bit 63
__device__ __forceinline__ float st_float(uint64_t r, int e_lo, int e_hi, bool allow_zero) {
  const int e = e_lo + (int)((r >> 40) % (uint64_t)(e_hi - e_lo + 1));
  uint32_t man = (uint32_t)r & 0x7FFFFFu;
  switch ((r >> 32) & 15u) {
    case 0: man = 0x7FFFFFu; break;
    case 1: man = 0u; break;
    case 2: man = 1u; break;
    case 3: man = 0x7FFFFEu; break;
    case 4: man = 1u << ((r >> 36) % 23); break;
    default: break;
  }
  uint32_t bits;
  if (e < -126) {                 // denormal: value = man * 2^-149 with the top bits shifted out by the "exponent"
    const int shift = min(-126 - e, 23);
    bits = man >> shift;
    if (!allow_zero && bits == 0) bits = 1u;
  } else {
    bits = ((uint32_t)(e + 127) << 23) | man;
  }
  if (r >> 63) bits |= 0x80000000u;
  return __uint_as_float(bits);
}

// regime 0: A (|lr*m| and v in the normal exact range); 1: S1 (tiny/denormal/zero m, v >= 2^-101);
// 2: S2 (also v denormal / zero).  Each thread draws NP pairs, runs `steps` steps with the packed loop and with
// step_sparse<ADAM>, and counts elements whose (var, m, v) bits differ, plus trajectories pk_valid rejected.
template <int MODE>
__global__ void __launch_bounds__(256) selftest_adam_packed_kernel(uint64_t seed, int64_t n, int steps, float lr,
                                                                  float l2, float nz,
                                                                  unsigned long long* __restrict__ out) {
  constexpr int NP = 4;
  __shared__ float smem[EPOCH_MAX_A + 2 * EPOCH_MAX_A * 256 / 8];   // nlr + a short ss_tmp (steps <= 4)
  SweepSmem sm;
  sm.nlr = smem; sm.ss_thr = smem + EPOCH_MAX_A; sm.ss_tmp = smem + EPOCH_MAX_A;
  Hyper h;
  h.b1 = 0.9f; h.b2 = 0.999f; h.eps = 1e-8f; h.l2 = l2; h.a0 = h.a1 = h.a2 = 0.f; h.lr = lr;
  if (threadIdx.x < EPOCH_MAX_A) sm.nlr[threadIdx.x] = -lr * (1.f + 0.03125f * threadIdx.x);
  __syncthreads();
  AdamPk c;
  c.l2 = h.l2; c.b1 = h.b1; c.b2 = h.b2; c.omb1 = __fsub_rn(1.f, h.b1); c.omb2 = __fsub_rn(1.f, h.b2);
  c.eps = h.eps; c.nz = nz;
  float b2n = 1.f;
  for (int s = 0; s < steps; ++s) b2n *= h.b2;
  unsigned long long bad = 0, rejected = 0, total = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float2 x[NP], m[NP], v[NP];
    float xr[2 * NP], mr[2 * NP], vr[2 * NP];
#pragma unroll
    for (int e = 0; e < 2 * NP; ++e) {
      const uint64_t r0 = st_mix64(seed + 24 * (uint64_t)i + 3 * e), r1 = st_mix64(seed + 24 * (uint64_t)i + 3 * e + 1),
                     r2 = st_mix64(seed + 24 * (uint64_t)i + 3 * e + 2);
      float xv, mv, vv;
      if (MODE == 0) {
        // lr*m in [2^-98, 2^20]: m exponent from -98-log2(lr) upwards; v in [2^-99, 2^36]; x anything moderate
        xv = st_float(r0, -60, 10, true);
        mv = st_float(r1, -85, 25, false);
        vv = fabsf(st_float(r2, -98, 36, false));
      } else {
        xv = (r0 & 7) ? st_float(r0, -135, -110, true) : st_float(r0, -160, -120, true);   // ~FLT_MIN, denormal, zero
        mv = (r1 & 3) ? st_float(r1, -172, -127, true) : st_float(r1, -126, -45, true);    // mostly denormal / zero
        vv = (MODE == 1) ? fabsf(st_float(r2, -98, -54, false)) : fabsf(st_float(r2, -175, -54, true));
        if ((r2 >> 20 & 15) == 0 && MODE == 2) vv = 0.f;
      }
      xr[e] = xv; mr[e] = mv; vr[e] = vv;
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      x[p] = make_float2(xr[2 * p], xr[2 * p + 1]); m[p] = make_float2(mr[2 * p], mr[2 * p + 1]);
      v[p] = make_float2(vr[2 * p], vr[2 * p + 1]);
    }
    const bool ok = pk_run<MODE, NP>(x, m, v, c, sm, 0, steps, b2n);
    for (int s = 0; s < steps; ++s) {
      h.lr = -sm.nlr[s];
#pragma unroll
      for (int e = 0; e < 2 * NP; ++e) step_sparse<CTR_OPT_ADAM>(xr[e], mr[e], vr[e], __fmul_rn(h.l2, xr[e]), h);
    }
    ++total;
    if (!ok) { ++rejected; continue; }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const float gx[2] = {x[p].x, x[p].y}, gm[2] = {m[p].x, m[p].y}, gv[2] = {v[p].x, v[p].y};
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int e = 2 * p + k;
        if (__float_as_uint(gx[k]) != __float_as_uint(xr[e]) || __float_as_uint(gm[k]) != __float_as_uint(mr[e]) ||
            __float_as_uint(gv[k]) != __float_as_uint(vr[e])) ++bad;
      }
    }
  }
  if (bad) atomicAdd(&out[0], bad);
  if (rejected) atomicAdd(&out[1], rejected);
  if (total) atomicAdd(&out[2], total);
}

// launcher used by ctr_epoch_sweep (epoch.cu).  Returns false if this path does not apply.
// CTR_SWEEP_MINB (tuning hook, tools/time_sweep.py): resident CTAs per SM the kernel is compiled for (2, 3, 4)
template <int MINB>
static void launch_sweep_minb(float* var, float* slot0, float* slot1, const uint8_t* last, int64_t n_rows, int K,
                              const float* hyper, const float* lr_table, int from, int upto, double* ss_partials,
                              int n_partials, int32_t* list, int32_t* list_count, int64_t list_cap, cudaStream_t st) {
  static bool attr = false;
  const int nsteps = upto - from;
  const size_t smem = (EPOCH_MAX_A + 2 * (size_t)nsteps * SWEEP_THREADS) * sizeof(float);
  if (!attr) {
    const int mx = (EPOCH_MAX_A + 2 * EPOCH_MAX_A * SWEEP_THREADS) * (int)sizeof(float);
    cudaFuncSetAttribute(epoch_sweep_adam_kernel<MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    cudaFuncSetAttribute(epoch_sweep_adam_k1_kernel<MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
    attr = true;
  }
  const float nz = -0.0f;
  const int grid = sm_count() * MINB;
  static int pf = -1;
  if (pf < 0) { const char* e = getenv("CTR_SWEEP_PF"); pf = e ? atoi(e) : 1; }
  if (K % 4 == 0) {
    const int f4 = K / 4;
    const int sh = (f4 & (f4 - 1)) == 0 ? (31 - __builtin_clz((unsigned)f4)) : -1;
    if (pf) {
      static bool attr_pf = false;
      if (!attr_pf) {
        cudaFuncSetAttribute(epoch_sweep_adam_kernel<MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (EPOCH_MAX_A + 2 * EPOCH_MAX_A * SWEEP_THREADS) * (int)sizeof(float));
        attr_pf = true;
      }
      epoch_sweep_adam_kernel<MINB, true><<<grid, SWEEP_THREADS, smem, st>>>(var, slot0, slot1, last, n_rows * f4, f4, sh,
                                                                             hyper, lr_table, from, upto, ss_partials,
                                                                             n_partials, list, list_count, list_cap, nz);
    } else {
      epoch_sweep_adam_kernel<MINB><<<grid, SWEEP_THREADS, smem, st>>>(var, slot0, slot1, last, n_rows * f4, f4, sh, hyper,
                                                                       lr_table, from, upto, ss_partials, n_partials, list,
                                                                       list_count, list_cap, nz);
    }
  } else {
    if (pf) {
      static bool attr_pf1 = false;
      if (!attr_pf1) {
        cudaFuncSetAttribute(epoch_sweep_adam_k1_kernel<MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (EPOCH_MAX_A + 2 * EPOCH_MAX_A * SWEEP_THREADS) * (int)sizeof(float));
        attr_pf1 = true;
      }
      epoch_sweep_adam_k1_kernel<MINB, true><<<grid, SWEEP_THREADS, smem, st>>>(var, slot0, slot1, last, n_rows / 4, hyper,
                                                                                lr_table, from, upto, ss_partials,
                                                                                n_partials, list, list_count, list_cap, nz);
    } else {
      epoch_sweep_adam_k1_kernel<MINB><<<grid, SWEEP_THREADS, smem, st>>>(var, slot0, slot1, last, n_rows / 4, hyper,
                                                                          lr_table, from, upto, ss_partials, n_partials,
                                                                          list, list_count, list_cap, nz);
    }
  }
}

bool launch_epoch_sweep_adam(float* var, float* slot0, float* slot1, const uint8_t* last, int64_t n_rows, int K,
                             const float* hyper, const float* lr_table, int from, int upto, double* ss_partials,
                             int n_partials, int32_t* list, int32_t* list_count, int64_t list_cap, int grid,
                             cudaStream_t st) {
  (void)grid;
  if (!(K % 4 == 0 || (K == 1 && n_rows % 4 == 0 && ((uintptr_t)last & 3) == 0))) return false;
  static int minb = 0;
  if (!minb) {
    const char* e = getenv("CTR_SWEEP_MINB");
    minb = e ? atoi(e) : 2;   // measured (profiles/r02_time_sweep_pf.txt): 2 CTAs/SM + register prefetch 44.4 ms, 3 CTAs/SM 48.0 ms
    if (minb < 2 || minb > 4) minb = 2;
  }
#define SW_ARGS var, slot0, slot1, last, n_rows, K, hyper, lr_table, from, upto, ss_partials, n_partials, list, list_count, list_cap, st
  if (minb == 2) launch_sweep_minb<2>(SW_ARGS);
  else if (minb == 4) launch_sweep_minb<4>(SW_ARGS);
  else launch_sweep_minb<3>(SW_ARGS);
#undef SW_ARGS
  return true;
}

}  // namespace ctr

using namespace ctr;

extern "C" int ctr_selftest_adam_packed(int regime, uint64_t seed, int64_t n, int steps, float lr, float l2,
                                        int64_t* out3, ctr_stream_t stream) {
  CTR_REQUIRE(out3 && n >= 0 && steps >= 1 && steps <= 4 && regime >= 0 && regime <= 2, CTR_ERR_INVALID_ARG,
              "ctr_selftest_adam_packed: bad args");
  cudaStream_t st = as_stream(stream);
  CTR_REQUIRE(cudaMemsetAsync(out3, 0, 3 * sizeof(int64_t), st) == cudaSuccess, CTR_ERR_CUDA,
              "ctr_selftest_adam_packed: memset failed");
  if (n == 0) return CTR_OK;
  unsigned long long* o = reinterpret_cast<unsigned long long*>(out3);
  const float nz = -0.0f;
  const int grid = sm_count() * 4;
  if (regime == 0) selftest_adam_packed_kernel<0><<<grid, 256, 0, st>>>(seed, n, steps, lr, l2, nz, o);
  else if (regime == 1) selftest_adam_packed_kernel<1><<<grid, 256, 0, st>>>(seed, n, steps, lr, l2, nz, o);
  else selftest_adam_packed_kernel<2><<<grid, 256, 0, st>>>(seed, n, steps, lr, l2, nz, o);
  CTR_LAUNCHED("ctr_selftest_adam_packed");
  return CTR_OK;
}
