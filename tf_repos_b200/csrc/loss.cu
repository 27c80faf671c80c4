// loss.cu -- logit assembly, sigmoid, sigmoid cross-entropy and its gradient.
//
// Replaces (DeepFM.py): :172-176  y = y_bias + y_w + y_v + y_d ; pred = sigmoid(y)
//                       :188      reduce_mean(sigmoid_cross_entropy_with_logits(y, labels))
// and the autodiff of both [TF-sem]: dy = (sigmoid(y) - t) * (1/B); d fm_bias = sum(dy).
// One CTA (B = 8192 => 8 samples per thread); the reductions use a fixed tree => deterministic.
#include "common.cuh"

namespace ctr {

constexpr int LOSS_THREADS = 1024;

__device__ __forceinline__ float block_sum_1024(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = sh[threadIdx.x];
    t = warp_sum(t);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

__global__ void __launch_bounds__(LOSS_THREADS)
logit_loss_kernel(const float* __restrict__ bias, const float* __restrict__ y_a,
                  const float* __restrict__ y_b, const float* __restrict__ y_c,
                  const float* __restrict__ labels, int B, int B_total, float* __restrict__ y_out,
                  float* __restrict__ pred, float* __restrict__ loss_ce, float* __restrict__ dy,
                  float* __restrict__ dbias) {
  __shared__ float sh[32];
  const float b0 = bias ? bias[0] : 0.f;
  const float invB = __fdiv_rn(1.f, (float)B_total);  // mean over the GLOBAL batch (data parallel)
  float lsum = 0.f, dsum = 0.f;
  for (int i = threadIdx.x; i < B; i += LOSS_THREADS) {
    float y = b0;                       // left-to-right like the reference expression
    if (y_a) y = __fadd_rn(y, y_a[i]);
    if (y_b) y = __fadd_rn(y, y_b[i]);
    if (y_c) y = __fadd_rn(y, y_c[i]);
    const float p = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-y)));
    if (y_out) y_out[i] = y;
    if (pred) pred[i] = p;
    if (labels) {
      const float t = labels[i];
      // max(x,0) - x*z + log(1+exp(-|x|))
      lsum += fmaxf(y, 0.f) - y * t + log1pf(expf(-fabsf(y)));
      const float d = __fmul_rn(__fsub_rn(p, t), invB);
      if (dy) dy[i] = d;
      dsum += d;
    }
  }
  if (labels) {
    float l = block_sum_1024(lsum, sh);
    float d = block_sum_1024(dsum, sh);
    if (threadIdx.x == 0) {
      if (loss_ce) loss_ce[0] = l * invB;
      if (dbias) dbias[0] = d;
    }
  }
}

}  // namespace ctr

using namespace ctr;

extern "C" int ctr_logit_loss(const float* bias, const float* y_a, const float* y_b, const float* y_c,
                              const float* labels, int B, int B_total, float* y, float* pred,
                              float* loss_ce, float* dy, float* dbias, ctr_stream_t stream) {
  // B_total == 1 with B > 1: SUM reduction (the canned estimators' head, wide_n_deep.py)
  CTR_REQUIRE(B >= 0 && (B_total >= B || B_total == 1), CTR_ERR_INVALID_ARG,
              "ctr_logit_loss: need 0 <= B <= B_total (or B_total == 1 for a summed loss)");
  if (B == 0) return CTR_OK;
  CTR_REQUIRE(y_a || y_b || y_c, CTR_ERR_INVALID_ARG, "ctr_logit_loss: no logit term given");
  logit_loss_kernel<<<1, LOSS_THREADS, 0, as_stream(stream)>>>(bias, y_a, y_b, y_c, labels, B, B_total, y, pred,
                                                               loss_ce, dy, dbias);
  CTR_LAUNCHED("ctr_logit_loss");
  return CTR_OK;
}
