// Spatial ordering of large query sets (map-sized sweeps, SURVEY.md §8 d S4/S5).
// Scans arrive in scan-line / voxel order and are already coherent; a caller-supplied set of 10^5-10^6
// queries may not be. Sorting them once per frame by a Morton code of their (sensor-frame) coordinates makes
// the lanes of a warp walk the same hash cells during association (coalesced point loads, L1 hits). A rigid
// pose change between iterations preserves that coherence, so one sort serves every association of the frame.
// Features are still written to the original slot (perm[i] = original index of sorted query i).
#include "common.cuh"
#include "sort.cuh"

namespace mml {

__device__ __forceinline__ unsigned spread10(unsigned v) {  // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void __launch_bounds__(256) k_morton_keys(const float4* __restrict__ q, int n, float inv_cell, float ox, float oy,
                                                     float oz, unsigned* __restrict__ keys, unsigned* __restrict__ vals) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float4 p = q[i];
  const int x = min(max((int)((p.x - ox) * inv_cell), 0), 1023);
  const int y = min(max((int)((p.y - oy) * inv_cell), 0), 1023);
  const int z = min(max((int)((p.z - oz) * inv_cell), 0), 1023);
  keys[i] = spread10((unsigned)x) | (spread10((unsigned)y) << 1) | (spread10((unsigned)z) << 2);
  vals[i] = (unsigned)i;
}

__global__ void __launch_bounds__(256) k_gather_queries(const float4* __restrict__ src, const unsigned* __restrict__ perm, int n,
                                                        float4* __restrict__ dst) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  dst[i] = src[perm[i]];
}

}  // namespace mml

using namespace mml;

int mml_bbox_device(mml_ctx* ctx, const float4* pts_d, int n, float* mn3, float* mx3);

// Sorts the n queries in q_d in place (through tmp) and leaves perm_d[i] = original index of sorted query i.
int mml_sort_queries_device(mml_ctx* ctx, float4* q_d, int n, mml::DevBuf& perm_buf) {
  if (n <= 0) return MML_OK;
  cudaStream_t st = ctx->stream;
  float mn[3], mx[3];
  MML_CHECK(mml_bbox_device(ctx, q_d, n, mn, mx));
  float ext = 1e-3f;
  for (int a = 0; a < 3; a++) ext = fmaxf(ext, mx[a] - mn[a]);
  const float inv_cell = 1023.0f / ext;
  const int nblocks = div_up(n, kRadixTile);
  for (int k = 0; k < 2; k++) {
    MML_CUDA(ctx, ctx->vox_keys[k].reserve(sizeof(unsigned) * (size_t)n));
    MML_CUDA(ctx, ctx->vox_vals[k].reserve(sizeof(unsigned) * (size_t)n));
  }
  MML_CUDA(ctx, ctx->vox_hist.reserve(sizeof(int) * (256 * (size_t)nblocks + 16)));
  MML_CUDA(ctx, ctx->tmp_e.reserve(sizeof(float4) * (size_t)n));
  MML_CUDA(ctx, perm_buf.reserve(sizeof(unsigned) * (size_t)n));
  MML_CUDA(ctx, ctx->counters.reserve(256));
  unsigned* keys[2] = {ctx->vox_keys[0].as<unsigned>(), ctx->vox_keys[1].as<unsigned>()};
  unsigned* vals[2] = {ctx->vox_vals[0].as<unsigned>(), ctx->vox_vals[1].as<unsigned>()};
  int* n_dev = ctx->counters.as<int>() + 32;
  MML_CUDA(ctx, cudaMemcpyAsync(n_dev, &n, sizeof(int), cudaMemcpyHostToDevice, st));
  k_morton_keys<<<div_up(n, 256), 256, 0, st>>>(q_d, n, inv_cell, mn[0], mn[1], mn[2], keys[0], vals[0]);
  MML_LAUNCHED(ctx);
  MML_CHECK(radix_sort_pairs(ctx, keys, vals, n_dev, n, ctx->vox_hist.as<int>()));
  MML_CUDA(ctx, cudaMemcpyAsync(perm_buf.p, vals[0], sizeof(unsigned) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  k_gather_queries<<<div_up(n, 256), 256, 0, st>>>(q_d, perm_buf.as<unsigned>(), n, ctx->tmp_e.as<float4>());
  MML_LAUNCHED(ctx);
  MML_CUDA(ctx, cudaMemcpyAsync(q_d, ctx->tmp_e.p, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  MML_CUDA(ctx, cudaStreamSynchronize(st));  // n lives on the host stack
  return MML_OK;
}
