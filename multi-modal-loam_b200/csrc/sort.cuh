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
    k_exclusive_scan<<<1, kScanThreads, 0, ctx->stream>>>(hist, nullptr, 256 * nblocks, nullptr);
    MML_LAUNCHED(ctx);
    k_radix_scatter<<<nblocks, 256, 0, ctx->stream>>>(keys[a], vals[a], keys[b], vals[b], n_dev, shift, nblocks, hist);
    MML_LAUNCHED(ctx);
  }
  MML_CUDA(ctx, cudaGetLastError());
  return MML_OK;
}

}  // namespace mml
