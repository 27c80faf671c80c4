// wide_deep.cu -- the feature-column front end of wide_n_deep.py:92-107.
//
//   categorical_column_with_identity(num_buckets, default_value=0)  -> id outside [0, NB) becomes 0
//   embedding_column(dimension=K), one table per column              -> row gather into the DNN input
//   numeric_column x 13                                              -> appended in the name-sorted column order
//   linear_model over the same columns                               -> per-sample sum of the selected weights
//
// The Fc per-column tables are stored stacked: column f owns rows [f*NB, (f+1)*NB) of `emb` ([Fc*NB, K]) and of
// `wide_cat` ([Fc*NB]); flat_ids = f*NB + clamped id is what the de-duplication / optimizer kernels see.
// x row layout (input_layer sorts columns by name): [C14_embedding .. C39_embedding | numerics in num_perm order].
#include "common.cuh"

namespace ctr {

// one warp per sample
__global__ void __launch_bounds__(256)
wd_input_fwd_kernel(const int32_t* __restrict__ ids, const float* __restrict__ dense, const float* __restrict__ emb,
                    const float* __restrict__ wide_cat, const float* __restrict__ wide_num,
                    const float* __restrict__ wide_bias, const int32_t* __restrict__ num_perm, int B, int Fc, int Fd,
                    int NB, int K, int32_t* __restrict__ flat_ids, float* __restrict__ x, float* __restrict__ lin) {
  const int b = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int D = Fc * K + Fd;
  float acc = 0.f;
  for (int f = lane; f < Fc; f += 32) {
    int id = ids[(int64_t)b * Fc + f];
    if (id < 0 || id >= NB) id = 0;
    const int fid = f * NB + id;
    flat_ids[(int64_t)b * Fc + f] = fid;
    if (wide_cat) acc += wide_cat[fid];
  }
  if (wide_num) {
    for (int j = lane; j < Fd; j += 32) acc += dense[(int64_t)b * Fd + j] * wide_num[j];
  }
  if (lin) {
    acc = warp_sum(acc);
    if (lane == 0) lin[b] = acc + (wide_bias ? wide_bias[0] : 0.f);
  }
  if (emb) {
    __syncwarp();   // flat_ids of this sample were written by this warp
    float* xr = x + (int64_t)b * D;
    const int n_el = Fc * K;
    for (int e = lane; e < n_el; e += 32) {
      const int f = e / K, k = e - f * K;
      int id = ids[(int64_t)b * Fc + f];
      if (id < 0 || id >= NB) id = 0;
      xr[e] = emb[((int64_t)f * NB + id) * K + k];
    }
    for (int j = lane; j < Fd; j += 32) xr[n_el + j] = dense[(int64_t)b * Fd + num_perm[j]];
  }
}

// per-occurrence gradients: g_rows[b*Fc+f, :] = dX[b, f*K : (f+1)*K], g_cat[b*Fc+f] = dy[b]
__global__ void __launch_bounds__(256)
wd_input_bwd_kernel(const float* __restrict__ dX, const float* __restrict__ dy, int B, int Fc, int Fd, int K,
                    float* __restrict__ g_rows, float* __restrict__ g_cat) {
  const int64_t n = (int64_t)B * Fc * K;
  const int D = Fc * K + Fd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / (Fc * K);
    const int e = (int)(i - b * Fc * K);
    if (g_rows) g_rows[i] = dX[b * D + e];
    if (g_cat && (e % K) == 0) g_cat[b * Fc + e / K] = dy[b];
  }
}

// g_num[j] = sum_b dy[b] * dense[b, j] (j < Fd) ; g_bias = sum_b dy[b].  One CTA per output, fixed tree.
__global__ void __launch_bounds__(256)
wd_linear_dense_grad_kernel(const float* __restrict__ dy, const float* __restrict__ dense, int B, int Fd,
                            float* __restrict__ g_num, float* __restrict__ g_bias) {
  __shared__ float red[256];
  const int j = blockIdx.x;   // j == Fd -> bias
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) s += (j < Fd) ? dy[b] * dense[(int64_t)b * Fd + j] : dy[b];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (j < Fd) g_num[j] = red[0];
    else g_bias[0] = red[0];
  }
}

}  // namespace ctr

using namespace ctr;

extern "C" {

int ctr_wd_input_fwd(const int32_t* ids, const float* dense, const float* emb, const float* wide_cat,
                     const float* wide_num, const float* wide_bias, const int32_t* num_perm, int B, int Fc, int Fd,
                     int NB, int K, int32_t* flat_ids, float* x, float* lin, ctr_stream_t stream) {
  CTR_REQUIRE(B >= 0 && Fc > 0 && Fd >= 0 && NB > 0 && K > 0, CTR_ERR_INVALID_ARG, "ctr_wd_input_fwd: bad sizes");
  if (B == 0) return CTR_OK;
  CTR_REQUIRE(ids && flat_ids && (Fd == 0 || dense), CTR_ERR_INVALID_ARG, "ctr_wd_input_fwd: null ids/dense");
  CTR_REQUIRE(!emb || (x && (Fd == 0 || num_perm)), CTR_ERR_INVALID_ARG, "ctr_wd_input_fwd: emb needs x and num_perm");
  CTR_REQUIRE(!(wide_cat || wide_num) || lin, CTR_ERR_INVALID_ARG, "ctr_wd_input_fwd: wide part needs lin");
  const int64_t threads = (int64_t)B * 32;
  wd_input_fwd_kernel<<<(unsigned)ceil_div64(threads, 256), 256, 0, as_stream(stream)>>>(
      ids, dense, emb, wide_cat, wide_num, wide_bias, num_perm, B, Fc, Fd, NB, K, flat_ids, x, lin);
  CTR_LAUNCHED("ctr_wd_input_fwd");
  return CTR_OK;
}

int ctr_wd_input_bwd(const float* dX, const float* dy, const float* dense, int B, int Fc, int Fd, int K, float* g_rows,
                     float* g_cat, float* g_num, float* g_bias, ctr_stream_t stream) {
  CTR_REQUIRE(B >= 0 && Fc > 0 && Fd >= 0 && K > 0, CTR_ERR_INVALID_ARG, "ctr_wd_input_bwd: bad sizes");
  if (B == 0) return CTR_OK;
  CTR_REQUIRE(!g_rows || dX, CTR_ERR_INVALID_ARG, "ctr_wd_input_bwd: g_rows needs dX");
  CTR_REQUIRE(!(g_cat || g_num || g_bias) || dy, CTR_ERR_INVALID_ARG, "ctr_wd_input_bwd: wide gradients need dy");
  CTR_REQUIRE(!g_num || (dense && g_bias), CTR_ERR_INVALID_ARG, "ctr_wd_input_bwd: g_num needs dense and g_bias");
  cudaStream_t st = as_stream(stream);
  if (g_rows || g_cat) {
    const int64_t n = (int64_t)B * Fc * K;
    const int grid = (int)std::min<int64_t>(ceil_div64(n, 256), (int64_t)sm_count() * 16);
    wd_input_bwd_kernel<<<grid, 256, 0, st>>>(dX, dy, B, Fc, Fd, K, g_rows, g_cat);
    CTR_LAUNCHED("ctr_wd_input_bwd(rows)");
  }
  if (g_num || g_bias) {
    wd_linear_dense_grad_kernel<<<(g_num ? Fd : 0) + 1, 256, 0, st>>>(dy, dense, B, g_num ? Fd : 0, g_num, g_bias);
    CTR_LAUNCHED("ctr_wd_input_bwd(dense)");
  }
  return CTR_OK;
}

}  // extern "C"
