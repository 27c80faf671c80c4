// tc_gemm.cu -- the dense GEMMs of the MLP / attention layers on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM), at fp32-class accuracy through a 3xTF32 split.
//
// Why 3xTF32: the parity target is 1e-5 relative on the logits; a plain TF32 (10-bit mantissa) or
// BF16 product loses 1e-3.  Each fp32 operand is split on the fly into hi = rn_tf32(a) and
// lo = a - hi (exact); D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi on the tensor cores recovers ~2^-21
// relative accuracy with fp32 accumulation in TMEM, at 1/3 of the TF32 rate (still ~5x the fp32
// SIMT pipe).
//
// Structure of one CTA (128 threads, one 128 x BN output tile, BN <= 128):
//   loop over 32-wide reduction slices, 2-stage ring:
//     all threads : global -> registers (any operand orientation; the gather transposes for free)
//                   -> hi/lo split -> st.shared in the canonical K-major SWIZZLE_128B UMMA layout
//                   fence.proxy.async ; __syncthreads
//     thread 0    : 4 k-steps x 3 tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8) ;
//                   tcgen05.commit -> mbarrier of the stage (frees it for the refill two slices later)
//   epilogue      : tcgen05.ld 32x32b (thread t owns accumulator row t) -> bias / group bias / relu /
//                   dropout / accumulate -> global
// Operands are staged by plain loads rather than TMA because the three products of a layer need
// three different operand orientations of fp32 data that must be split anyway.
#include "common.cuh"

namespace ctr {

constexpr int TC_BM = 128, TC_BK = 32, TC_STAGES = 2;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;            // 16 KB: one 128 x 32 fp32 operand tile
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;           // A_hi, A_lo, B_hi, B_lo
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   start_address[0,14) = addr>>4 ; LBO[16,30) = 1 (unused for swizzled K-major) ; SBO[32,46) = 1024 B >> 4
//   version[46,48) = 1 (sm_100) ; layout_type[61,64) = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, both K-major
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                      // c_format  = F32
  d |= 2u << 7;                      // a_format  = TF32
  d |= 2u << 10;                     // b_format  = TF32
  d |= (uint32_t)(N >> 3) << 17;     // n_dim
  d |= (uint32_t)(M >> 4) << 24;     // m_dim
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// byte offset of (row, 16-byte chunk c) inside a K-major SWIZZLE_128B tile of 32 fp32 per row
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

// C[i][j] = sum_r A(i,r) * B(r,j)
//   A_RC: A(i,r) = A[i*lda + r] else A[r*lda + i];  B_RC: B(r,j) = B[j*ldb + r] else B[r*ldb + j]
// EPI 0: store (split-R chunk z to C + z*M*ldc)  1: act(acc + bias + gbias[i/gP]) (/keep*mask)  2: C += acc
template <bool A_RC, bool B_RC, int EPI>
__global__ void __launch_bounds__(128, 1)
tc_gemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C,
               int ldc, int M, int N, int R, const float* __restrict__ bias, int act, const float* __restrict__ mask,
               float keep, const float* __restrict__ gbias, int gP) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_stage[TC_STAGES];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B wants 1024 B alignment
  const int i0 = blockIdx.y * TC_BM, j0 = blockIdx.x * TC_BM;
  const int n_here = min(N - j0, TC_BM);
  const int n_pad = (n_here + 15) & ~15;                       // UMMA_N: multiple of 16 for M = 128
  const int r_chunk = (R + gridDim.z - 1) / gridDim.z;
  const int r_begin = blockIdx.z * r_chunk, r_end = min(R, r_begin + r_chunk);
  const int KT = (r_end - r_begin + TC_BK - 1) / TC_BK;

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) mbar_init(&bar_stage[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {   // TMEM: 2 x 128 columns x 128 lanes of fp32 accumulators (main | small terms)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = make_idesc(TC_BM, n_pad);

  for (int kt = 0; kt < KT; ++kt) {
    const int s = kt % TC_STAGES;
    if (kt >= TC_STAGES) {   // the MMAs that read this stage two slices ago must be done
      mbar_wait(&bar_stage[s], (uint32_t)(((kt / TC_STAGES) - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    uint8_t* st = smem + s * TC_STAGE_BYTES;
    const int r0 = r_begin + kt * TC_BK;
    // ---- operand A: thread t stages row i0+t ----
    {
      float v[TC_BK];
      const int gi = i0 + tid;
      if (A_RC) {
        const float* p = A + (int64_t)gi * lda + r0;
        const bool full = gi < M && r0 + TC_BK <= r_end && ((((uintptr_t)p) & 15) == 0);
        if (full) {
#pragma unroll
          for (int c = 0; c < 8; ++c) { const float4 t = *reinterpret_cast<const float4*>(p + 4 * c); v[4*c] = t.x; v[4*c+1] = t.y; v[4*c+2] = t.z; v[4*c+3] = t.w; }
        } else {
#pragma unroll
          for (int k = 0; k < TC_BK; ++k) v[k] = (gi < M && r0 + k < r_end) ? p[k] : 0.f;
        }
      } else {
#pragma unroll
        for (int k = 0; k < TC_BK; ++k) v[k] = (gi < M && r0 + k < r_end) ? A[(int64_t)(r0 + k) * lda + gi] : 0.f;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 hi, lo;
        hi.x = tf32_hi(v[4*c]); hi.y = tf32_hi(v[4*c+1]); hi.z = tf32_hi(v[4*c+2]); hi.w = tf32_hi(v[4*c+3]);
        lo.x = v[4*c] - hi.x; lo.y = v[4*c+1] - hi.y; lo.z = v[4*c+2] - hi.z; lo.w = v[4*c+3] - hi.w;
        const uint32_t off = sw128_off(tid, c);
        *reinterpret_cast<float4*>(st + off) = hi;
        *reinterpret_cast<float4*>(st + TC_TILE_BYTES + off) = lo;
      }
    }
    // ---- operand B: thread t stages output column j0+t ----
    if (tid < n_pad) {
      float v[TC_BK];
      const int gj = j0 + tid;
      if (B_RC) {
        const float* p = B + (int64_t)gj * ldb + r0;
        const bool full = gj < N && r0 + TC_BK <= r_end && ((((uintptr_t)p) & 15) == 0);
        if (full) {
#pragma unroll
          for (int c = 0; c < 8; ++c) { const float4 t = *reinterpret_cast<const float4*>(p + 4 * c); v[4*c] = t.x; v[4*c+1] = t.y; v[4*c+2] = t.z; v[4*c+3] = t.w; }
        } else {
#pragma unroll
          for (int k = 0; k < TC_BK; ++k) v[k] = (gj < N && r0 + k < r_end) ? p[k] : 0.f;
        }
      } else {
#pragma unroll
        for (int k = 0; k < TC_BK; ++k) v[k] = (gj < N && r0 + k < r_end) ? B[(int64_t)(r0 + k) * ldb + gj] : 0.f;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 hi, lo;
        hi.x = tf32_hi(v[4*c]); hi.y = tf32_hi(v[4*c+1]); hi.z = tf32_hi(v[4*c+2]); hi.w = tf32_hi(v[4*c+3]);
        lo.x = v[4*c] - hi.x; lo.y = v[4*c+1] - hi.y; lo.z = v[4*c+2] - hi.z; lo.w = v[4*c+3] - hi.w;
        const uint32_t off = sw128_off(tid, c);
        *reinterpret_cast<float4*>(st + 2 * TC_TILE_BYTES + off) = hi;
        *reinterpret_cast<float4*>(st + 3 * TC_TILE_BYTES + off) = lo;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(st), a_lo = a_hi + TC_TILE_BYTES, b_hi = a_hi + 2 * TC_TILE_BYTES, b_lo = a_hi + 3 * TC_TILE_BYTES;
#pragma unroll
      for (int k = 0; k < TC_BK / 8; ++k) {                        // UMMA_K = 8 tf32 = 32 bytes inside the 128 B swizzle atom
        const uint32_t ko = k * 32;
        // The tensor core accumulates in fp32 with truncation, so every accumulation costs ~2^-24 of the
        // accumulator, with a bias.  The two cross terms (2^-11 of the result) go to their own accumulator:
        // the main one then sees one accumulation per k-step instead of three.
        const uint32_t first = (kt > 0 || k > 0) ? 1u : 0u;
        umma_tf32(tmem_d + 128, make_desc(a_lo + ko), make_desc(b_hi + ko), idesc, first);
        umma_tf32(tmem_d + 128, make_desc(a_hi + ko), make_desc(b_lo + ko), idesc, 1u);
        umma_tf32(tmem_d, make_desc(a_hi + ko), make_desc(b_hi + ko), idesc, first);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_stage[s])) : "memory");
    }
  }
  if (KT > 0) {   // the last commit tracks every MMA issued before it
    mbar_wait(&bar_stage[(KT - 1) % TC_STAGES], (uint32_t)(((KT - 1) / TC_STAGES) & 1));
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: thread t <-> accumulator row t (TMEM lane 32*warp + lane) ----
  const int gi = i0 + tid;
  float* Cz = C + (EPI == 0 ? (int64_t)blockIdx.z * M * ldc : 0);
  for (int c0 = 0; c0 < n_pad; c0 += 16) {
    uint32_t r[16], r2[16];
    if (KT > 0) {
      const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3]), "=r"(r2[4]), "=r"(r2[5]), "=r"(r2[6]), "=r"(r2[7]), "=r"(r2[8]),
            "=r"(r2[9]), "=r"(r2[10]), "=r"(r2[11]), "=r"(r2[12]), "=r"(r2[13]), "=r"(r2[14]), "=r"(r2[15])
          : "r"(taddr + 128));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 16; ++q) r[q] = __float_as_uint(__uint_as_float(r[q]) + __uint_as_float(r2[q]));
    } else {
#pragma unroll
      for (int q = 0; q < 16; ++q) r[q] = 0u;
    }
    if (gi < M) {
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int gj = j0 + c0 + q;
        if (gj < N) {
          float v = __uint_as_float(r[q]);
          if (EPI == 2) v += Cz[(int64_t)gi * ldc + gj];
          if (EPI == 1) {
            if (bias) v += bias[gj];
            if (gbias) v += gbias[(int64_t)(gi / gP) * N + gj];
            if (act == 1) v = fmaxf(v, 0.f);
            if (mask) v = __fdiv_rn(v, keep) * mask[(int64_t)gi * ldc + gj];
          }
          Cz[(int64_t)gi * ldc + gj] = v;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256));
}

template <bool A_RC, bool B_RC, int EPI>
static int launch_tc(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int R, int S,
                     const float* bias, int act, const float* mask, float keep, const float* gbias, int gP,
                     cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(tc_gemm_kernel<A_RC, B_RC, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    attr = true;
  }
  dim3 grid((N + TC_BM - 1) / TC_BM, (M + TC_BM - 1) / TC_BM, S);
  tc_gemm_kernel<A_RC, B_RC, EPI><<<grid, 128, TC_SMEM_BYTES, st>>>(A, lda, B, ldb, C, ldc, M, N, R, bias, act, mask, keep,
                                                                    gbias, gP);
  return 0;
}

// entry used by fc.cu: kind 0 = forward (A_RC, !B_RC, EPI 1), 1 = dIn (A_RC, B_RC, EPI 0 / 2), 2 = dW (!A_RC, !B_RC, EPI 0)
int tc_gemm_dispatch(int kind, int epi, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N,
                     int R, int S, const float* bias, int act, const float* mask, float keep, const float* gbias, int gP,
                     cudaStream_t st) {
  if (kind == 0) return launch_tc<true, false, 1>(A, lda, B, ldb, C, ldc, M, N, R, S, bias, act, mask, keep, gbias, gP, st);
  if (kind == 1 && epi == 2) return launch_tc<true, true, 2>(A, lda, B, ldb, C, ldc, M, N, R, S, bias, act, mask, keep, gbias, gP, st);
  if (kind == 1) return launch_tc<true, true, 0>(A, lda, B, ldb, C, ldc, M, N, R, S, bias, act, mask, keep, gbias, gP, st);
  return launch_tc<false, false, 0>(A, lda, B, ldb, C, ldc, M, N, R, S, bias, act, mask, keep, gbias, gP, st);
}

}  // namespace ctr
