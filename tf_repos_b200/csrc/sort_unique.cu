// sort_unique.cu -- K3: de-duplication of the embedding-gradient IndexedSlices.
//
// Replaces optimizer._deduplicate_indexed_slices (tf.unique + tf.unsorted_segment_sum) which
// optimizer.minimize (DeepFM.py:213) applies to the FM_V / FM_W gradients [TF-sem].
//
// Algorithm: stable LSD radix sort of (id, position) -- ceil(bits/10) passes of <=10-bit digits
// (3 passes cover N <= 2^30, i.e. the 1e9-row table of config 5) -- then run-length encoding.
// Every pass is three small kernels (tile histogram + digit totals -> per-digit column scan over
// the tiles -> stable scatter), linear in n, no inter-CTA spinning, so the whole thing is
// graph-capturable and cannot hang.
// n = B*F = 319 488 at config 2: all buffers (2.5 MB) stay L2-resident; the cost is launch
// latency, not bandwidth.
#include "common.cuh"

namespace ctr {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 2048 keys per CTA
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int MAX_DIGIT_BITS = 10;
constexpr int MAX_BINS = 1 << MAX_DIGIT_BITS;

__device__ __forceinline__ uint32_t sanitize_key(int32_t id, int64_t N) {
  // ids outside [0,N) are an error the forward already flagged; keep the sort memory-safe
  return (id < 0 || (int64_t)id >= N) ? 0u : (uint32_t)id;
}

__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const int32_t* __restrict__ keys, int64_t n, int64_t N, int shift, int nbins,
                  int32_t* __restrict__ hist, int32_t* __restrict__ totals) {
  __shared__ int32_t sh[MAX_BINS];
  for (int d = threadIdx.x; d < nbins; d += SORT_THREADS) sh[d] = 0;
  __syncthreads();
  const int64_t start = (int64_t)blockIdx.x * SORT_TILE;
  const uint32_t mask = (uint32_t)nbins - 1u;
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    int64_t i = start + r * SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&sh[(sanitize_key(keys[i], N) >> shift) & mask], 1);
  }
  __syncthreads();
  // tile-major [tile][digit] (coalesced here, in the column scan and in the scatter) + digit totals
  for (int d = threadIdx.x; d < nbins; d += SORT_THREADS) {
    const int32_t c = sh[d];
    hist[(int64_t)blockIdx.x * nbins + d] = c;
    if (c) atomicAdd(&totals[d], c);
  }
}

// hist[tile][digit] -> exclusive prefix over tiles, per digit column, in place.
// One CTA = 32 columns x 32 row segments: thread (seg, col) sums its segment (warp = 32 consecutive
// columns of one segment => 128 B coalesced rows), a 32-step scan over the segments in shared
// memory, then the segment is rewritten with its running prefix.  Linear in n_tiles.
__global__ void __launch_bounds__(1024)
radix_colscan_kernel(int32_t* __restrict__ hist, int nbins, int n_tiles) {
  __shared__ int32_t seg_sum[32][33];
  const int col_l = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + col_l;
  const int rows_per = (n_tiles + 31) / 32;
  const int r0 = min(n_tiles, seg * rows_per), r1 = min(n_tiles, r0 + rows_per);
  int32_t sum = 0;
  if (col < nbins)
    for (int r = r0; r < r1; ++r) sum += hist[(int64_t)r * nbins + col];
  seg_sum[seg][col_l] = sum;
  __syncthreads();
  if (seg == 0) {
    int32_t run = 0;
#pragma unroll
    for (int sgm = 0; sgm < 32; ++sgm) {
      const int32_t v = seg_sum[sgm][col_l];
      seg_sum[sgm][col_l] = run;
      run += v;
    }
  }
  __syncthreads();
  if (col < nbins) {
    int32_t run = seg_sum[seg][col_l];
    for (int r = r0; r < r1; ++r) {
      const int64_t i = (int64_t)r * nbins + col;
      const int32_t v = hist[i];
      hist[i] = run;
      run += v;
    }
  }
}

// stable scatter of one tile.  Order inside the tile is (warp, round, lane) == index order.
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const int32_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in,
                     int32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out, int64_t n,
                     int64_t N, int shift, int nbins, const int32_t* __restrict__ tile_prefix,
                     const int32_t* __restrict__ totals) {
  extern __shared__ int32_t warp_hist[];  // [SORT_WARPS][nbins]
  __shared__ int32_t digit_base[MAX_BINS];  // exclusive scan of the digit totals
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int d = tid; d < SORT_WARPS * nbins; d += SORT_THREADS) warp_hist[d] = 0;
  for (int d = tid; d < nbins; d += SORT_THREADS) digit_base[d] = totals[d];
  __syncthreads();
  if (w == 0) {  // one warp: lane-contiguous chunks, then a shuffle scan over the 32 chunk sums
    const int per = (nbins + 31) / 32;
    const int lo = min(nbins, lane * per), hi = min(nbins, lo + per);
    int32_t sum = 0;
    for (int d = lo; d < hi; ++d) sum += digit_base[d];
    int32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(FULL_MASK, incl, o);
      if (lane >= o) incl += t;
    }
    int32_t run = incl - sum;
    for (int d = lo; d < hi; ++d) {
      const int32_t v = digit_base[d];
      digit_base[d] = run;
      run += v;
    }
  }
  __syncthreads();
  const uint32_t mask = (uint32_t)nbins - 1u;
  const int64_t wstart = (int64_t)blockIdx.x * SORT_TILE + (int64_t)w * (SORT_ITEMS * 32);
  int32_t* my_hist = warp_hist + w * nbins;
  uint32_t key[SORT_ITEMS];
  int32_t val[SORT_ITEMS], rank[SORT_ITEMS], dig[SORT_ITEMS];
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    const int64_t i = wstart + r * 32 + lane;
    const bool ok = i < n;
    key[r] = ok ? sanitize_key(keys_in[i], N) : 0u;
    val[r] = ok ? (vals_in ? vals_in[i] : (int32_t)i) : 0;
    dig[r] = ok ? (int32_t)((key[r] >> shift) & mask) : -1 - lane;  // invalid lanes match nobody
    const uint32_t peers = __match_any_sync(FULL_MASK, dig[r]);
    const int leader = __ffs(peers) - 1;
    int32_t base = 0;
    if (ok && lane == leader) {
      base = my_hist[dig[r]];
      my_hist[dig[r]] = base + __popc(peers);
    }
    base = __shfl_sync(FULL_MASK, base, leader);
    rank[r] = base + __popc(peers & lt_mask);
    __syncwarp();
  }
  __syncthreads();
  // per digit: exclusive prefix over the warps of this tile + the tile's global offset
  for (int d = tid; d < nbins; d += SORT_THREADS) {
    int32_t run = digit_base[d] + tile_prefix[(int64_t)blockIdx.x * nbins + d];
#pragma unroll
    for (int ww = 0; ww < SORT_WARPS; ++ww) {
      int32_t cnt = warp_hist[ww * nbins + d];
      warp_hist[ww * nbins + d] = run;
      run += cnt;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    if (dig[r] >= 0) {
      const int32_t dst = my_hist[dig[r]] + rank[r];
      keys_out[dst] = (int32_t)key[r];
      vals_out[dst] = val[r];
    }
  }
}

// ---- run-length encoding of the sorted keys ---------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS)
heads_count_kernel(const int32_t* __restrict__ keys, int64_t n, int32_t* __restrict__ tile_counts,
                   int n_tiles, int32_t* __restrict__ long_list) {
  __shared__ int32_t wsum[SORT_WARPS];
  const int64_t start = (int64_t)blockIdx.x * SORT_TILE;
  int32_t c = 0;
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    int64_t i = start + r * SORT_THREADS + threadIdx.x;
    if (i < n) c += (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL_MASK, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t t = 0;
#pragma unroll
    for (int ww = 0; ww < SORT_WARPS; ++ww) t += wsum[ww];
    tile_counts[blockIdx.x] = t;
    if (blockIdx.x == 0) long_list[0] = 0;
  }
}

__global__ void __launch_bounds__(SORT_THREADS)
heads_emit_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ perm, int64_t n,
                  const int32_t* __restrict__ tile_counts, int32_t* __restrict__ uniq,
                  int32_t* __restrict__ seg_offsets, int32_t* __restrict__ inverse,
                  int32_t* __restrict__ n_uniq) {
  __shared__ int32_t warp_tot[SORT_WARPS];
  __shared__ int32_t tile_off_s[SORT_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  {  // runs that start in earlier tiles
    int32_t c = 0;
    for (int t = tid; t < (int)blockIdx.x; t += SORT_THREADS) c += tile_counts[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL_MASK, c, o);
    if (lane == 0) tile_off_s[w] = c;
  }
  // blocked arrangement: thread t owns items [t*ITEMS, (t+1)*ITEMS) of the tile
  const int64_t start = (int64_t)blockIdx.x * SORT_TILE + (int64_t)tid * SORT_ITEMS;
  int32_t k[SORT_ITEMS];
  bool head[SORT_ITEMS];
  int32_t prev = (start > 0 && start - 1 < n) ? keys[start - 1] : -1;
  int32_t cnt = 0;
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    const int64_t i = start + r;
    k[r] = (i < n) ? keys[i] : 0;
    head[r] = (i < n) && (i == 0 || k[r] != prev);
    prev = k[r];
    cnt += head[r] ? 1 : 0;
  }
  int32_t incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(FULL_MASK, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  int32_t wbase = 0;
#pragma unroll
  for (int ww = 0; ww < SORT_WARPS; ++ww) wbase += (ww < w) ? warp_tot[ww] : 0;
  int32_t tile_off = 0;
#pragma unroll
  for (int ww = 0; ww < SORT_WARPS; ++ww) tile_off += tile_off_s[ww];
  int32_t u = tile_off + wbase + incl - cnt - 1;  // index of the run before my first item
#pragma unroll
  for (int r = 0; r < SORT_ITEMS; ++r) {
    const int64_t i = start + r;
    if (i < n) {
      if (head[r]) {
        ++u;
        uniq[u] = k[r];
        seg_offsets[u] = (int32_t)i;
      }
      inverse[perm[i]] = u;
      if (i == n - 1) {
        seg_offsets[u + 1] = (int32_t)n;
        n_uniq[0] = u + 1;
      }
    }
  }
}

__global__ void long_list_kernel(const int32_t* __restrict__ seg_offsets,
                                 const int32_t* __restrict__ n_uniq, int32_t* __restrict__ long_list,
                                 int64_t n) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n || u >= n_uniq[0]) return;
  if (seg_offsets[u + 1] - seg_offsets[u] > CTR_LONG_SEG) {
    int32_t slot = atomicAdd(&long_list[0], 1);
    long_list[1 + slot] = (int32_t)u;
  }
}

__global__ void empty_unique_kernel(int32_t* n_uniq, int32_t* seg_offsets, int32_t* long_list) {
  n_uniq[0] = 0; seg_offsets[0] = 0; long_list[0] = 0;
}

struct SortPlan {
  int bits, passes, digit_bits, n_tiles;
  size_t off_keys_a, off_keys_b, off_vals_a, off_hist, off_tiles, off_totals, total;
};

static SortPlan make_plan(int64_t n, int64_t N) {
  SortPlan p;
  int bits = 1;
  while (bits < 31 && ((int64_t)1 << bits) < N) ++bits;
  p.bits = bits;
  p.passes = (bits + MAX_DIGIT_BITS - 1) / MAX_DIGIT_BITS;
  p.digit_bits = (bits + p.passes - 1) / p.passes;
  p.n_tiles = (int)ceil_div64(n > 0 ? n : 1, SORT_TILE);
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t o = 0;
  p.off_keys_a = o; o = align(o + (size_t)n * 4);
  p.off_keys_b = o; o = align(o + (size_t)n * 4);
  p.off_vals_a = o; o = align(o + (size_t)n * 4);
  p.off_hist = o;   o = align(o + ((size_t)MAX_BINS * p.n_tiles + 1) * 4);
  p.off_tiles = o;  o = align(o + ((size_t)p.n_tiles + 1) * 4);
  p.off_totals = o; o = align(o + (size_t)4 * MAX_BINS * 4);  // one totals row per pass (<= 4 passes)
  p.total = o;
  return p;
}

}  // namespace ctr

using namespace ctr;

extern "C" {

size_t ctr_unique_segment_workspace_bytes(int64_t n, int64_t N) {
  if (n < 0 || N <= 0) return 0;
  return make_plan(n, N).total;
}

int ctr_unique_segment(const int32_t* ids, int64_t n, int64_t N, int32_t* perm, int32_t* uniq,
                       int32_t* inverse, int32_t* seg_offsets, int32_t* n_uniq, int32_t* long_list,
                       void* ws, size_t ws_bytes, ctr_stream_t stream) {
  CTR_REQUIRE(n >= 0 && N > 0, CTR_ERR_INVALID_ARG, "ctr_unique_segment: bad n=%lld N=%lld",
              (long long)n, (long long)N);
  CTR_REQUIRE(n < ((int64_t)1 << 31) - SORT_TILE && N <= ((int64_t)1 << 31) - 1, CTR_ERR_UNSUPPORTED,
              "ctr_unique_segment: n and N must fit int32");
  CTR_REQUIRE(n_uniq && seg_offsets && long_list, CTR_ERR_INVALID_ARG,
              "ctr_unique_segment: null n_uniq/seg_offsets/long_list");
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    empty_unique_kernel<<<1, 1, 0, st>>>(n_uniq, seg_offsets, long_list);
    CTR_LAUNCHED("ctr_unique_segment(empty)");
    return CTR_OK;
  }
  CTR_REQUIRE(ids && perm && uniq && inverse, CTR_ERR_INVALID_ARG, "ctr_unique_segment: null buffer");
  SortPlan p = make_plan(n, N);
  CTR_REQUIRE(ws && ws_bytes >= p.total, CTR_ERR_WORKSPACE,
              "ctr_unique_segment: workspace %zu < required %zu", ws_bytes, p.total);
  char* base = reinterpret_cast<char*>(ws);
  int32_t* keys_ab[2] = {reinterpret_cast<int32_t*>(base + p.off_keys_a),
                         reinterpret_cast<int32_t*>(base + p.off_keys_b)};
  int32_t* vals_a = reinterpret_cast<int32_t*>(base + p.off_vals_a);
  int32_t* hist = reinterpret_cast<int32_t*>(base + p.off_hist);
  int32_t* tiles = reinterpret_cast<int32_t*>(base + p.off_tiles);
  int32_t* totals = reinterpret_cast<int32_t*>(base + p.off_totals);
  const int nbins = 1 << p.digit_bits;
  if (cudaMemsetAsync(totals, 0, (size_t)4 * MAX_BINS * 4, st) != cudaSuccess) {
    set_error("ctr_unique_segment: memset failed");
    return CTR_ERR_CUDA;
  }
  static_assert((SORT_WARPS + 1) * MAX_BINS * 4 <= 48 * 1024, "scatter smem must fit the default carve-out");
  const int32_t* kin = ids;
  const int32_t* vin = nullptr;  // implicit iota
  for (int pass = 0; pass < p.passes; ++pass) {
    const int shift = pass * p.digit_bits;
    int32_t* kout = keys_ab[pass & 1];
    // the last pass must land the positions in `perm`
    int32_t* vout = (((p.passes - 1 - pass) & 1) == 0) ? perm : vals_a;
    int32_t* tot = totals + pass * MAX_BINS;
    radix_hist_kernel<<<p.n_tiles, SORT_THREADS, 0, st>>>(kin, n, N, shift, nbins, hist, tot);
    CTR_LAUNCHED("radix_hist");
    radix_colscan_kernel<<<(nbins + 31) / 32, 1024, 0, st>>>(hist, nbins, p.n_tiles);
    CTR_LAUNCHED("radix_colscan");
    radix_scatter_kernel<<<p.n_tiles, SORT_THREADS, SORT_WARPS * nbins * 4, st>>>(
        kin, vin, kout, vout, n, N, shift, nbins, hist, tot);
    CTR_LAUNCHED("radix_scatter");
    kin = kout;
    vin = vout;
  }
  heads_count_kernel<<<p.n_tiles, SORT_THREADS, 0, st>>>(kin, n, tiles, p.n_tiles, long_list);
  CTR_LAUNCHED("heads_count");
  heads_emit_kernel<<<p.n_tiles, SORT_THREADS, 0, st>>>(kin, perm, n, tiles, uniq, seg_offsets,
                                                        inverse, n_uniq);
  CTR_LAUNCHED("heads_emit");
  long_list_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(seg_offsets, n_uniq, long_list, n);
  CTR_LAUNCHED("long_list");
  return CTR_OK;
}

}  // extern "C"
