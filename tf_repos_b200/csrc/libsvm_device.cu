// libsvm_device.cu -- libsvm tokenizer on the GPU (replaces decode_libsvm, DeepFM.py:65-81,
// when the text is already in device memory).
//
// Contract: whatever this path accepts, it converts to EXACTLY the bits the host parser (libsvm_host.cu:
// strtof / strtol) produces; everything it is not sure about is counted in `info` and the caller re-parses
// that chunk on the host (tf_repos_b200/input_fn.py does).  "Not sure" =
//   * a blank line, a malformed line, a pair count != F            (host parser owns the error messages)
//   * a number outside the fast decimal path: > 15 significant digits, |decimal exponent| > 22, inf/nan/hex,
//     a result outside the normal fp32 range, or a double that sits within one ulp of an fp32 rounding
//     boundary (fp32(RN_double(m / 10^k)) could then differ from the correctly rounded strtof by double rounding;
//     about 6e-9 of all values).
// Inside the fast path the conversion is exact: m < 2^53 and 10^k (k <= 22) are exact doubles, one IEEE
// double division/multiplication gives the correctly rounded double, and away from an fp32 boundary
// rounding that double to fp32 equals rounding the exact decimal value.
//
// Kernels: (1) count '\n' per 4 KB block; (2) scan the block counts; (3) emit line starts; (4) one thread
// per line walks its bytes (adjacent threads read adjacent lines, so sectors are shared through L1).
#include "common.cuh"

namespace ctr {

constexpr int LS_THREADS = 256, LS_BYTES_PER_THREAD = 16, LS_BLOCK_BYTES = LS_THREADS * LS_BYTES_PER_THREAD;

__device__ __forceinline__ int count_nl16(const unsigned char* __restrict__ t, int64_t pos, int64_t len, uint32_t& mask) {
  mask = 0;
  if (pos + 16 <= len && ((reinterpret_cast<uintptr_t>(t + pos) & 15) == 0)) {
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(t + pos));
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (((ws[k] >> (8 * b)) & 0xFFu) == '\n') mask |= 1u << (4 * k + b);
    }
  } else {
    for (int b = 0; b < 16; ++b)
      if (pos + b < len && t[pos + b] == '\n') mask |= 1u << b;
  }
  return __popc(mask);
}

__global__ void __launch_bounds__(LS_THREADS) ls_count_kernel(const unsigned char* __restrict__ text, int64_t len,
                                                             int32_t* __restrict__ block_counts) {
  __shared__ int warp_tot[LS_THREADS / 32];
  const int64_t pos = ((int64_t)blockIdx.x * LS_THREADS + threadIdx.x) * LS_BYTES_PER_THREAD;
  uint32_t mask;
  int c = pos < len ? count_nl16(text, pos, len, mask) : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL_MASK, c, o);
  if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < LS_THREADS / 32; ++w) s += warp_tot[w];
    block_counts[blockIdx.x] = s;
  }
}

// exclusive scan of block_counts (one CTA, sequential over tiles of 1024); total -> info_lines[0]
__global__ void __launch_bounds__(1024) ls_scan_kernel(int32_t* __restrict__ block_counts, int n_blocks,
                                                       int64_t* __restrict__ n_newlines) {
  __shared__ int64_t warp_sum_s[32];
  __shared__ int64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int64_t v = i < n_blocks ? block_counts[i] : 0;
    int64_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(FULL_MASK, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sum_s[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int64_t w = warp_sum_s[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t y = __shfl_up_sync(FULL_MASK, w, o);
        if (threadIdx.x >= o) w += y;
      }
      warp_sum_s[threadIdx.x] = w;   // inclusive over warps
    }
    __syncthreads();
    const int64_t before = carry_s + (threadIdx.x >= 32 ? warp_sum_s[(threadIdx.x >> 5) - 1] : 0) + (x - v);
    // block counts are < 2^31 in total for any buffer this API accepts (len < 2^31 * 1 byte per newline)
    if (i < n_blocks) block_counts[i] = (int32_t)before;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) n_newlines[0] = carry_s;
}

// line_start[k+1] = position just after the k-th '\n' (k < max_rows); line_start[0] = 0
__global__ void __launch_bounds__(LS_THREADS) ls_emit_kernel(const unsigned char* __restrict__ text, int64_t len,
                                                            const int32_t* __restrict__ block_offsets, int64_t max_rows,
                                                            int64_t* __restrict__ line_start) {
  __shared__ int warp_tot[LS_THREADS / 32];
  const int64_t pos = ((int64_t)blockIdx.x * LS_THREADS + threadIdx.x) * LS_BYTES_PER_THREAD;
  uint32_t mask = 0;
  const int c = pos < len ? count_nl16(text, pos, len, mask) : 0;
  int x = c;
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(FULL_MASK, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
  __syncthreads();
  int before = x - c;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += warp_tot[w];
  int64_t k = (int64_t)block_offsets[blockIdx.x] + before;
  if (blockIdx.x == 0 && threadIdx.x == 0) line_start[0] = 0;
  while (mask) {
    const int b = __ffs(mask) - 1;
    mask &= mask - 1;
    if (k < max_rows) line_start[k + 1] = pos + b + 1;
    ++k;
  }
}

__constant__ double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                  1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

enum { LS_OK = 0, LS_BAD = 1, LS_HOST = 2 };

// decimal float at p (no leading spaces): [+-]digits[.digits][(e|E)[+-]digits].  Returns LS_OK and advances p,
// LS_BAD if no number starts here, LS_HOST if the host's strtof has to decide.
__device__ __forceinline__ int parse_float(const unsigned char* __restrict__ t, int64_t& p, int64_t e, float& out) {
  int64_t q = p;
  bool neg = false;
  if (q < e && (t[q] == '+' || t[q] == '-')) { neg = t[q] == '-'; ++q; }
  uint64_t m = 0;
  int sig = 0, exp10 = 0, any = 0;
  bool dropped = false;
  while (q < e && t[q] >= '0' && t[q] <= '9') {
    any = 1;
    const int d = t[q] - '0';
    if (sig < 18) { if (sig || d) { m = m * 10 + d; ++sig; } }
    else { dropped = true; ++exp10; }
    ++q;
  }
  if (q < e && t[q] == '.') {
    ++q;
    while (q < e && t[q] >= '0' && t[q] <= '9') {
      any = 1;
      const int d = t[q] - '0';
      if (sig < 18) { if (sig || d) { m = m * 10 + d; ++sig; } --exp10; }
      else dropped = true;
      ++q;
    }
  }
  if (!any) {
    // "inf", "nan", "0x..." and friends are the host's business; anything else is not a number
    const unsigned char c = q < e ? t[q] : 0;
    return (c == 'i' || c == 'I' || c == 'n' || c == 'N') ? LS_HOST : LS_BAD;
  }
  if (q < e && (t[q] == 'e' || t[q] == 'E')) {
    int64_t r = q + 1;
    bool eneg = false;
    if (r < e && (t[r] == '+' || t[r] == '-')) { eneg = t[r] == '-'; ++r; }
    if (r < e && t[r] >= '0' && t[r] <= '9') {
      int ex = 0;
      while (r < e && t[r] >= '0' && t[r] <= '9') { if (ex < 10000) ex = ex * 10 + (t[r] - '0'); ++r; }
      exp10 += eneg ? -ex : ex;
      q = r;
    }  // else: "1e" / "1e+" -> strtof stops before the 'e'
  }
  if (q < e && (t[q] == 'x' || t[q] == 'X')) return LS_HOST;   // "0x1p3": hex float
  p = q;
  if (m == 0) { out = neg ? -0.0f : 0.0f; return LS_OK; }
  if (dropped || sig > 15 || exp10 < -22 || exp10 > 22) return LS_HOST;
  const double d = exp10 < 0 ? __ddiv_rn((double)m, kPow10[-exp10]) : __dmul_rn((double)m, kPow10[exp10]);
  if (!(d >= 1.1754943508222875e-38 && d <= 3.4028234663852886e38)) return LS_HOST;   // fp32 subnormal / overflow
  const uint64_t low = (uint64_t)__double_as_longlong(d) & 0x1FFFFFFFull;              // bits below the fp32 mantissa
  if (low >= 0x0FFFFFFFull && low <= 0x10000001ull) return LS_HOST;                     // next to a rounding boundary
  const float f = __double2float_rn(d);
  out = neg ? -f : f;
  return LS_OK;
}

// one thread per line.  status[0] = blank lines, [1] = malformed lines, [2] = lines with a number for the host
__global__ void __launch_bounds__(128) ls_parse_kernel(const unsigned char* __restrict__ text, int64_t len,
                                                      const int64_t* __restrict__ line_start,
                                                      const int64_t* __restrict__ n_newlines, int64_t max_rows, int F,
                                                      int final_chunk, int32_t* __restrict__ ids, float* __restrict__ vals,
                                                      float* __restrict__ labels, int64_t* __restrict__ info) {
  const int64_t nn = n_newlines[0];
  const bool tail = final_chunk && len > 0 && text[len - 1] != '\n';
  int64_t n_lines = nn + (tail ? 1 : 0);
  if (n_lines > max_rows) n_lines = max_rows;
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row == 0) {
    info[0] = n_lines;
    info[1] = n_lines == 0 ? 0 : ((n_lines <= nn) ? line_start[n_lines] : len);   // bytes consumed
  }
  if (row >= n_lines) return;
  int64_t p = line_start[row];
  int64_t e = (row < nn) ? line_start[row + 1] - 1 : len;     // exclusive end, '\n' dropped
  if (e > p && text[e - 1] == '\r') --e;
  while (p < e && text[p] == ' ') ++p;
  if (p == e) { atomicAdd(reinterpret_cast<unsigned long long*>(&info[2]), 1ull); return; }
  int st = LS_OK;
  float lab = 0.f;
  st |= parse_float(text, p, e, lab);
  int f = 0;
  int32_t* id_row = ids + row * F;
  float* val_row = vals + row * F;
  while (st == LS_OK) {
    if (p < e && text[p] != ' ') { st = LS_BAD; break; }      // a number must be followed by a space or the end
    while (p < e && text[p] == ' ') ++p;
    if (p >= e) break;
    if (f >= F) { st = LS_BAD; break; }
    // id: [+-]digits ':'   (strtol; more than 9 digits could overflow int32 -> host)
    bool neg = false;
    if (text[p] == '+' || text[p] == '-') { neg = text[p] == '-'; ++p; }
    int64_t v = 0;
    int nd = 0;
    while (p < e && text[p] >= '0' && text[p] <= '9') { if (nd < 12) v = v * 10 + (text[p] - '0'); ++nd; ++p; }
    if (nd == 0 || p >= e || text[p] != ':') { st = LS_BAD; break; }
    if (nd > 9) { st = LS_HOST; break; }
    ++p;
    if (p < e && (text[p] == ' ' || text[p] == '\t')) { st = LS_HOST; break; }   // strtof would skip the blank
    float val = 0.f;
    st |= parse_float(text, p, e, val);
    if (st != LS_OK) break;
    id_row[f] = (int32_t)(neg ? -v : v);
    val_row[f] = val;
    ++f;
  }
  if (st == LS_OK && f != F) st = LS_BAD;
  if (st & LS_BAD) atomicAdd(reinterpret_cast<unsigned long long*>(&info[3]), 1ull);
  else if (st & LS_HOST) atomicAdd(reinterpret_cast<unsigned long long*>(&info[4]), 1ull);
  else labels[row] = lab;
}

}  // namespace ctr

using namespace ctr;

extern "C" {

size_t ctr_parse_libsvm_device_workspace_bytes(size_t len, int64_t max_rows) {
  const size_t n_blocks = (len + LS_BLOCK_BYTES - 1) / LS_BLOCK_BYTES;
  return n_blocks * sizeof(int32_t) + 16 + (size_t)(max_rows + 1) * sizeof(int64_t) + 16;
}

int ctr_parse_libsvm_device(const char* text, size_t len, int F, int64_t max_rows, int final_chunk, int32_t* ids,
                            float* vals, float* labels, int64_t* info, void* ws, size_t ws_bytes, ctr_stream_t stream) {
  CTR_REQUIRE(F > 0 && max_rows >= 0 && info && (len == 0 || text), CTR_ERR_INVALID_ARG,
              "ctr_parse_libsvm_device: bad arguments");
  CTR_REQUIRE(len < ((size_t)1 << 32), CTR_ERR_INVALID_ARG, "ctr_parse_libsvm_device: buffer too large (len < 2^32)");
  CTR_REQUIRE(max_rows == 0 || (ids && vals && labels), CTR_ERR_INVALID_ARG, "ctr_parse_libsvm_device: null output");
  CTR_REQUIRE(ws && ws_bytes >= ctr_parse_libsvm_device_workspace_bytes(len, max_rows), CTR_ERR_WORKSPACE,
              "ctr_parse_libsvm_device: workspace too small");
  cudaStream_t st = as_stream(stream);
  CTR_REQUIRE(cudaMemsetAsync(info, 0, 5 * sizeof(int64_t), st) == cudaSuccess, CTR_ERR_CUDA,
              "ctr_parse_libsvm_device: memset failed");
  if (len == 0 || max_rows == 0) return CTR_OK;
  const int n_blocks = (int)((len + LS_BLOCK_BYTES - 1) / LS_BLOCK_BYTES);
  int32_t* block_counts = reinterpret_cast<int32_t*>(ws);
  int64_t* n_newlines = reinterpret_cast<int64_t*>(reinterpret_cast<uint8_t*>(ws) + (((size_t)n_blocks * 4 + 15) & ~(size_t)15));
  int64_t* line_start = n_newlines + 2;
  const unsigned char* t = reinterpret_cast<const unsigned char*>(text);
  ls_count_kernel<<<n_blocks, LS_THREADS, 0, st>>>(t, (int64_t)len, block_counts);
  CTR_LAUNCHED("ctr_parse_libsvm_device(count)");
  ls_scan_kernel<<<1, 1024, 0, st>>>(block_counts, n_blocks, n_newlines);
  CTR_LAUNCHED("ctr_parse_libsvm_device(scan)");
  ls_emit_kernel<<<n_blocks, LS_THREADS, 0, st>>>(t, (int64_t)len, block_counts, max_rows, line_start);
  CTR_LAUNCHED("ctr_parse_libsvm_device(emit)");
  ls_parse_kernel<<<(unsigned)ceil_div64(max_rows, 128), 128, 0, st>>>(t, (int64_t)len, line_start, n_newlines, max_rows, F,
                                                                       final_chunk, ids, vals, labels, info);
  CTR_LAUNCHED("ctr_parse_libsvm_device(parse)");
  return CTR_OK;
}

}  // extern "C"
