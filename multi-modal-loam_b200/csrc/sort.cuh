// Device-wide primitives used by the voxel filter and the map builder:
//   * exclusive scan of an int array (single CTA, tiled)
//   * stable LSD radix sort of (u32 key, u32 value) pairs, 8-bit digits
// Element counts are read from device memory (`n_dev`) so callers can chain stages without
// a host round trip; grids are sized from a host-side upper bound.
#pragma once
#include "common.cuh"

namespace mml {

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;

// data[i] <- sum_{k<i} data[k]; *total_out (may be null) <- sum of all.  n = *n_dev (or n_host if n_dev null)
static __global__ void __launch_bounds__(kScanThreads) k_exclusive_scan(int* __restrict__ data, const int* __restrict__ n_dev,
                                                                 int n_host, int* __restrict__ total_out) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  const int n = n_dev ? *n_dev : n_host;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += kScanThreads * kScanItems) {
    int v[kScanItems];
    int tsum = 0;
    const int i0 = base + threadIdx.x * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      v[k] = (i0 + k < n) ? data[i0 + k] : 0;
      tsum += v[k];
    }
    int x = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sum[lane];
      int xs = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, xs, d);
        if (lane >= d) xs += y;
      }
      warp_sum[lane] = xs - w;  // exclusive over warps
    }
    __syncthreads();
    const int carry = carry_s;
    int run = carry + warp_sum[warp] + (x - tsum);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
      if (i0 + k < n) data[i0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == kScanThreads - 1) carry_s = run;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

// ---- single-pass scan of large arrays: decoupled look-back (one read and one write per element). Tiles are
// handed out by an atomic ticket so that a tile's predecessors are always resident or finished; a tile publishes its
// aggregate, looks back over its predecessors' status words (AGGREGATE / INCLUSIVE, value in the low 32 bits) until
// it meets an inclusive prefix, then publishes its own inclusive prefix.
constexpr int kLbThreads = 256;
constexpr int kLbItems = 16;
constexpr int kLbTile = kLbThreads * kLbItems;  // 4096 elements per tile
constexpr unsigned long long kLbAggregate = 1ull << 32, kLbInclusive = 2ull << 32;

static __global__ void __launch_bounds__(kLbThreads) k_scan_lookback(int* __restrict__ data, const int* __restrict__ n_dev, int n_host,
                                                              int* __restrict__ total_out, unsigned long long* __restrict__ status,
                                                              unsigned* __restrict__ ticket) {
  __shared__ int s_tile, s_prefix;
  __shared__ int warp_sum[kLbThreads / 32];
  const int n = n_dev ? *n_dev : n_host;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  const int base = tile * kLbTile;
  if (base >= n) return;
  // blocked arrangement: thread t owns items [t * kLbItems, (t + 1) * kLbItems) of the tile
  int v[kLbItems];
  int tsum = 0;
  const int i0 = base + threadIdx.x * kLbItems;
  if (i0 + kLbItems <= n) {
    const int4* p = reinterpret_cast<const int4*>(data + i0);
#pragma unroll
    for (int k = 0; k < kLbItems / 4; k++) {
      const int4 q = p[k];
      v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kLbItems; k++) v[k] = (i0 + k < n) ? data[i0 + k] : 0;
  }
#pragma unroll
  for (int k = 0; k < kLbItems; k++) tsum += v[k];
  int x = tsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) warp_sum[warp] = x;
  __syncthreads();
  int woff = 0, tile_sum = 0;
#pragma unroll
  for (int w = 0; w < kLbThreads / 32; w++) {
    if (w < warp) woff += warp_sum[w];
    tile_sum += warp_sum[w];
  }
  if (threadIdx.x == 0) {
    int prefix = 0;
    if (tile == 0) {
      atomicExch(&status[0], kLbInclusive | (unsigned)tile_sum);
    } else {
      atomicExch(&status[tile], kLbAggregate | (unsigned)tile_sum);
      for (int t = tile - 1;; t--) {
        unsigned long long st;
        do { st = atomicAdd(&status[t], 0ull); } while ((st >> 32) == 0);
        prefix += (int)(unsigned)st;
        if ((st >> 32) == 2) break;
      }
      atomicExch(&status[tile], kLbInclusive | (unsigned)(prefix + tile_sum));
    }
    s_prefix = prefix;
    if (total_out && base + kLbTile >= n) *total_out = prefix + tile_sum;
  }
  __syncthreads();
  int run = s_prefix + woff + (x - tsum);
  if (i0 + kLbItems <= n) {
    int4* p = reinterpret_cast<int4*>(data + i0);
#pragma unroll
    for (int k = 0; k < kLbItems / 4; k++) {
      int4 q;
      q.x = run; run += v[4 * k];
      q.y = run; run += v[4 * k + 1];
      q.z = run; run += v[4 * k + 2];
      q.w = run; run += v[4 * k + 3];
      p[k] = q;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kLbItems; k++) {
      if (i0 + k < n) data[i0 + k] = run;
      run += v[k];
    }
  }
}

// data[i] <- sum_{k<i} data[k] for i < n (n = *n_dev when given, at most n_max), *total_out (may be null) <- the sum.
// Small arrays: one CTA, one launch (latency). Large arrays: the look-back scan over div_up(n_max, 4096) tiles.
static inline int exclusive_scan_device(mml_ctx* ctx, int* data, const int* n_dev, int n_max, int* total_out) {
  if (n_max <= 4 * kScanThreads * kScanItems) {
    k_exclusive_scan<<<1, kScanThreads, 0, ctx->stream>>>(data, n_dev, n_dev ? 0 : n_max, total_out);
    MML_LAUNCHED(ctx);
    return MML_OK;
  }
  const int tiles = div_up(n_max, kLbTile);
  MML_CUDA(ctx, ctx->scan_state.reserve(sizeof(unsigned long long) * (size_t)(tiles + 2)));
  MML_CUDA(ctx, cudaMemsetAsync(ctx->scan_state.p, 0, sizeof(unsigned long long) * (size_t)(tiles + 2), ctx->stream));
  unsigned long long* status = ctx->scan_state.as<unsigned long long>() + 1;
  if (total_out && n_dev) MML_CUDA(ctx, cudaMemsetAsync(total_out, 0, sizeof(int), ctx->stream));  // n may be 0
  k_scan_lookback<<<tiles, kLbThreads, 0, ctx->stream>>>(data, n_dev, n_dev ? 0 : n_max, total_out, status,
                                                       reinterpret_cast<unsigned*>(ctx->scan_state.p));
  MML_LAUNCHED(ctx);
  return MML_OK;
}

constexpr int kRadixTile = 1024;  // elements per CTA (4 rounds of 256)

static __global__ void __launch_bounds__(256) k_radix_hist(const unsigned* __restrict__ keys, const int* __restrict__ n_dev,
                                                    int shift, int nblocks, int* __restrict__ hist) {
  __shared__ int h[256];
  const int n = *n_dev;
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * kRadixTile;
  for (int r = threadIdx.x; r < kRadixTile; r += 256) {
    const int i = base + r;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1);
  }
  __syncthreads();
  hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

static __global__ void __launch_bounds__(256) k_radix_scatter(const unsigned* __restrict__ keys_in, const unsigned* __restrict__ vals_in,
                                                       unsigned* __restrict__ keys_out, unsigned* __restrict__ vals_out,
                                                       const int* __restrict__ n_dev, int shift, int nblocks,
                                                       const int* __restrict__ hist_scanned) {
  __shared__ int wcnt[8][256];
  __shared__ int base[256];
  const int n = *n_dev;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  base[threadIdx.x] = hist_scanned[threadIdx.x * nblocks + blockIdx.x];
  const int tile0 = blockIdx.x * kRadixTile;
  if (tile0 >= n) return;
  for (int r0 = 0; r0 < kRadixTile; r0 += 256) {
    if (tile0 + r0 >= n) break;  // uniform
#pragma unroll
    for (int w = 0; w < 8; w++) wcnt[w][threadIdx.x] = 0;
    __syncthreads();
    const int i = tile0 + r0 + threadIdx.x;
    const bool act = i < n;
    unsigned key = 0, val = 0;
    int d = 256 + lane;  // inactive lanes never match an active digit
    if (act) {
      key = keys_in[i];
      val = vals_in[i];
      d = (key >> shift) & 255u;
    }
    const unsigned mask = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(mask & ((1u << lane) - 1u));
    if (act && rank == 0) wcnt[warp][d] = __popc(mask);
    __syncthreads();
    if (act) {
      int off = base[d] + rank;
      for (int w = 0; w < warp; w++) off += wcnt[w][d];
      keys_out[off] = key;
      vals_out[off] = val;
    }
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) t += wcnt[w][threadIdx.x];
    base[threadIdx.x] += t;
    __syncthreads();
  }
}

// Sort n (= *n_dev <= n_max) pairs by key, stable. Result ends in keys[0]/vals[0] after an
// even number of passes (4). hist must hold 256 * ceil(n_max / kRadixTile) ints.
static inline int radix_sort_pairs(mml_ctx* ctx, unsigned* keys[2], unsigned* vals[2], const int* n_dev, int n_max,
                                   int* hist) {
  const int nblocks = div_up(n_max > 0 ? n_max : 1, kRadixTile);
  for (int pass = 0; pass < 4; pass++) {
    const int shift = 8 * pass;
    const int a = pass & 1, b = a ^ 1;
    k_radix_hist<<<nblocks, 256, 0, ctx->stream>>>(keys[a], n_dev, shift, nblocks, hist);
    MML_LAUNCHED(ctx);
    MML_CHECK(exclusive_scan_device(ctx, hist, nullptr, 256 * nblocks, nullptr));
    k_radix_scatter<<<nblocks, 256, 0, ctx->stream>>>(keys[a], vals[a], keys[b], vals[b], n_dev, shift, nblocks, hist);
    MML_LAUNCHED(ctx);
  }
  MML_CUDA(ctx, cudaGetLastError());
  return MML_OK;
}

}  // namespace mml
