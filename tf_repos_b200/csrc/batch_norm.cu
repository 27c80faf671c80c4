// batch_norm.cu -- tf.contrib.layers.batch_norm(decay, center=True, scale=True, updates_collections=None) as the
// reference applies it AFTER the relu of each hidden layer (batch_norm_layer, DeepFM.py:159-160,231-235; same in
// DCN.py:171, PNN.py:181, NFM.py:143, DIN.py:206), followed by the layer's dropout (DeepFM.py:161-162).
//
// [TF-sem] non-fused path for rank-2 inputs: mean, variance = tf.nn.moments(x, [0]) (biased variance, computed as
// mean((x - mean)^2)); y = tf.nn.batch_normalization: inv = rsqrt(var + 0.001) * gamma; y = x*inv + (beta - mean*inv);
// TRAIN: moving_x -= (moving_x - batch_x) * (1 - decay), applied in place with the forward (updates_collections=None);
// EVAL / PREDICT: moving statistics.  Gradients flow through the batch moments (autodiff of the same graph).
//
// Layout: x [n, H] row-major, H <= a few hundred.  Column reductions: 32-column slabs x 32 row chunks (256 CTAs at H = 256),
// 8 row groups x 32 columns per CTA (128 B coalesced row segments), sequential accumulation per thread, a fixed 8-way tree,
// then a fixed-order merge of the chunks (Chan's formula for the variance): deterministic, no atomics.
#include "common.cuh"

namespace ctr {

constexpr int BN_ROWG = 8;
constexpr int BN_CHUNKS = 32;   // row chunks per column slab: 8 slabs x 32 chunks = 256 CTAs at H = 256

// chunk blockIdx.y of the rows: its own mean (pass 1) and M2 = sum (x - mean_c)^2 (pass 2), per column
__global__ void __launch_bounds__(256)
bn_chunk_stats_kernel(const float* __restrict__ x, int n, int H, float* __restrict__ part_mean,
                      float* __restrict__ part_m2) {
  __shared__ float red[BN_ROWG][33];
  __shared__ float mu_s[32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), g = threadIdx.x >> 5, ch = blockIdx.y;
  const int r0 = (int)((int64_t)n * ch / BN_CHUNKS), r1 = (int)((int64_t)n * (ch + 1) / BN_CHUNKS);
  const bool in = c < H;
  float s = 0.f;
  if (in) for (int r = r0 + g; r < r1; r += BN_ROWG) s += x[(int64_t)r * H + c];
  red[g][threadIdx.x & 31] = s;
  __syncthreads();
  if (g == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < BN_ROWG; ++k) t += red[k][threadIdx.x];
    mu_s[threadIdx.x] = r1 > r0 ? t / (float)(r1 - r0) : 0.f;
  }
  __syncthreads();
  const float mu = mu_s[threadIdx.x & 31];
  float q = 0.f;
  if (in) for (int r = r0 + g; r < r1; r += BN_ROWG) { const float d = x[(int64_t)r * H + c] - mu; q += d * d; }
  __syncthreads();
  red[g][threadIdx.x & 31] = q;
  __syncthreads();
  if (g == 0 && in) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < BN_ROWG; ++k) t += red[k][threadIdx.x];
    part_mean[(int64_t)ch * H + c] = mu;
    part_m2[(int64_t)ch * H + c] = t;
  }
}

// merge the chunks in chunk order (Chan et al.): mean = sum n_c mean_c / n ; M2 = sum M2_c + sum n_c (mean_c - mean)^2
// (= tf.nn.moments' mean((x - mean)^2) up to rounding); TRAIN also updates the moving statistics in place
__global__ void bn_merge_stats_kernel(const float* __restrict__ part_mean, const float* __restrict__ part_m2, int n,
                                      int H, float* __restrict__ mean, float* __restrict__ var,
                                      float* __restrict__ moving_mean, float* __restrict__ moving_var, float decay) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  float tot = 0.f;
  for (int ch = 0; ch < BN_CHUNKS; ++ch) {
    const int nc = (int)((int64_t)n * (ch + 1) / BN_CHUNKS) - (int)((int64_t)n * ch / BN_CHUNKS);
    tot += (float)nc * part_mean[(int64_t)ch * H + c];
  }
  const float mu = tot / (float)n;
  float m2 = 0.f;
  for (int ch = 0; ch < BN_CHUNKS; ++ch) {
    const int nc = (int)((int64_t)n * (ch + 1) / BN_CHUNKS) - (int)((int64_t)n * ch / BN_CHUNKS);
    const float d = part_mean[(int64_t)ch * H + c] - mu;
    m2 += part_m2[(int64_t)ch * H + c] + (float)nc * d * d;
  }
  const float v = m2 / (float)n;
  mean[c] = mu; var[c] = v;
  if (moving_mean) {   // assign_moving_average: variable -= (variable - value) * (1 - decay)
    const float omd = 1.f - decay;
    moving_mean[c] = moving_mean[c] - (moving_mean[c] - mu) * omd;
    moving_var[c] = moving_var[c] - (moving_var[c] - v) * omd;
  }
}

// out = dropout(x*inv + (beta - mean*inv)), inv = rsqrt(var + eps)*gamma
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, int64_t total, int H, const float* __restrict__ mean,
                const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                float eps, const float* __restrict__ mask, float keep, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % H);
    const float inv = (1.f / sqrtf(var[c] + eps)) * gamma[c];
    float y = x[i] * inv + (beta[c] - mean[c] * inv);
    if (mask) y = __fdiv_rn(y, keep) * mask[i];
    out[i] = y;
  }
}

// partial dgamma / dbeta of row chunk blockIdx.y: dY = d_out (/keep*mask), xhat = (x - mean)*rsqrt(var+eps)
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ d_out, const float* __restrict__ x, int n, int H,
                     const float* __restrict__ mean, const float* __restrict__ var, float eps,
                     const float* __restrict__ mask, float keep, float* __restrict__ part_g,
                     float* __restrict__ part_b) {
  __shared__ float red_g[BN_ROWG][33], red_b[BN_ROWG][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), g = threadIdx.x >> 5, ch = blockIdx.y;
  const int r0 = (int)((int64_t)n * ch / BN_CHUNKS), r1 = (int)((int64_t)n * (ch + 1) / BN_CHUNKS);
  const bool in = c < H;
  float sg = 0.f, sb = 0.f;
  if (in) {
    const float mu = mean[c], rstd = 1.f / sqrtf(var[c] + eps);
    for (int r = r0 + g; r < r1; r += BN_ROWG) {
      const int64_t i = (int64_t)r * H + c;
      float dy = d_out[i];
      if (mask) dy = __fdiv_rn(dy, keep) * mask[i];
      sb += dy;
      sg += dy * ((x[i] - mu) * rstd);
    }
  }
  red_g[g][threadIdx.x & 31] = sg; red_b[g][threadIdx.x & 31] = sb;
  __syncthreads();
  if (g == 0 && in) {
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int k = 0; k < BN_ROWG; ++k) { tg += red_g[k][threadIdx.x]; tb += red_b[k][threadIdx.x]; }
    part_g[(int64_t)ch * H + c] = tg; part_b[(int64_t)ch * H + c] = tb;
  }
}

__global__ void bn_bwd_merge_kernel(const float* __restrict__ part_g, const float* __restrict__ part_b, int H,
                                    float* __restrict__ d_gamma, float* __restrict__ d_beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  float tg = 0.f, tb = 0.f;
  for (int ch = 0; ch < BN_CHUNKS; ++ch) { tg += part_g[(int64_t)ch * H + c]; tb += part_b[(int64_t)ch * H + c]; }
  d_gamma[c] = tg; d_beta[c] = tb;
}

// d_x = gamma*rstd * (dY - d_beta/n - xhat*d_gamma/n)
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ d_out, const float* __restrict__ x, int64_t total, int n, int H,
                    const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ gamma,
                    float eps, const float* __restrict__ mask, float keep, const float* __restrict__ d_gamma,
                    const float* __restrict__ d_beta, float* __restrict__ d_x) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float inv_n = 1.f / (float)n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % H);
    const float rstd = 1.f / sqrtf(var[c] + eps);
    float dy = d_out[i];
    if (mask) dy = __fdiv_rn(dy, keep) * mask[i];
    const float xhat = (x[i] - mean[c]) * rstd;
    d_x[i] = gamma[c] * rstd * (dy - d_beta[c] * inv_n - xhat * (d_gamma[c] * inv_n));
  }
}

}  // namespace ctr

using namespace ctr;

extern "C" {

size_t ctr_bn_workspace_bytes(int H) { return H > 0 ? (size_t)2 * BN_CHUNKS * (size_t)H * sizeof(float) : 0; }

int ctr_bn_fwd(const float* x, int n, int H, const float* gamma, const float* beta, float* moving_mean,
               float* moving_var, int train, float decay, float eps, const float* mask, float keep, float* out,
               float* save_mean, float* save_var, void* ws, size_t ws_bytes, ctr_stream_t stream) {
  CTR_REQUIRE(n >= 0 && H > 0, CTR_ERR_INVALID_ARG, "ctr_bn_fwd: bad n/H");
  if (n == 0) return CTR_OK;
  CTR_REQUIRE(x && gamma && beta && moving_mean && moving_var && out, CTR_ERR_INVALID_ARG, "ctr_bn_fwd: null buffer");
  CTR_REQUIRE(!train || (save_mean && save_var), CTR_ERR_INVALID_ARG, "ctr_bn_fwd: save_mean/save_var required in TRAIN mode");
  CTR_REQUIRE(!train || (ws && ws_bytes >= ctr_bn_workspace_bytes(H)), CTR_ERR_WORKSPACE, "ctr_bn_fwd: workspace too small");
  CTR_REQUIRE(!mask || keep > 0.f, CTR_ERR_INVALID_ARG, "ctr_bn_fwd: keep must be > 0 with a mask");
  cudaStream_t st = as_stream(stream);
  const int64_t total = (int64_t)n * H;
  const int64_t gmax = (int64_t)sm_count() * 8, gwant = ceil_div64(total, 256);
  const int grid = (int)(gwant < gmax ? gwant : gmax);
  if (train) {
    float* pm = reinterpret_cast<float*>(ws);
    float* pq = pm + (size_t)BN_CHUNKS * H;
    bn_chunk_stats_kernel<<<dim3((H + 31) / 32, BN_CHUNKS), 256, 0, st>>>(x, n, H, pm, pq);
    CTR_LAUNCHED("ctr_bn_fwd(chunk stats)");
    bn_merge_stats_kernel<<<(H + 127) / 128, 128, 0, st>>>(pm, pq, n, H, save_mean, save_var, moving_mean, moving_var, decay);
    CTR_LAUNCHED("ctr_bn_fwd(merge)");
    bn_apply_kernel<<<grid, 256, 0, st>>>(x, total, H, save_mean, save_var, gamma, beta, eps, mask, keep, out);
  } else {
    bn_apply_kernel<<<grid, 256, 0, st>>>(x, total, H, moving_mean, moving_var, gamma, beta, eps, nullptr, 1.f, out);
  }
  CTR_LAUNCHED("ctr_bn_fwd");
  return CTR_OK;
}

int ctr_bn_bwd(const float* d_out, const float* x, int n, int H, const float* save_mean, const float* save_var,
               const float* gamma, float eps, const float* mask, float keep, float* d_x, float* d_gamma, float* d_beta,
               void* ws, size_t ws_bytes, ctr_stream_t stream) {
  CTR_REQUIRE(n >= 0 && H > 0, CTR_ERR_INVALID_ARG, "ctr_bn_bwd: bad n/H");
  if (n == 0) return CTR_OK;
  CTR_REQUIRE(d_out && x && save_mean && save_var && gamma && d_x && d_gamma && d_beta, CTR_ERR_INVALID_ARG,
              "ctr_bn_bwd: null buffer");
  CTR_REQUIRE(!mask || keep > 0.f, CTR_ERR_INVALID_ARG, "ctr_bn_bwd: keep must be > 0 with a mask");
  CTR_REQUIRE(ws && ws_bytes >= ctr_bn_workspace_bytes(H), CTR_ERR_WORKSPACE, "ctr_bn_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int64_t total = (int64_t)n * H;
  float* pg = reinterpret_cast<float*>(ws);
  float* pb = pg + (size_t)BN_CHUNKS * H;
  bn_bwd_reduce_kernel<<<dim3((H + 31) / 32, BN_CHUNKS), 256, 0, st>>>(d_out, x, n, H, save_mean, save_var, eps, mask, keep, pg, pb);
  CTR_LAUNCHED("ctr_bn_bwd(reduce)");
  bn_bwd_merge_kernel<<<(H + 127) / 128, 128, 0, st>>>(pg, pb, H, d_gamma, d_beta);
  CTR_LAUNCHED("ctr_bn_bwd(merge)");
  const int64_t gmax = (int64_t)sm_count() * 8, gwant = ceil_div64(total, 256);
  const int grid = (int)(gwant < gmax ? gwant : gmax);
  bn_bwd_apply_kernel<<<grid, 256, 0, st>>>(d_out, x, total, n, H, save_mean, save_var, gamma, eps, mask, keep, d_gamma,
                                            d_beta, d_x);
  CTR_LAUNCHED("ctr_bn_bwd");
  return CTR_OK;
}

}  // extern "C"
