// Per-point geometry stages around the extractor, device resident:
//   k_undistort        RemoveLidarDistortion          src/unionPoseEstimation.cpp:402-421
//   k_velo_*           ring + relative time           src/unionFeatureExtract.cpp:1136-1195
//   k_hori_filter      Horizon CustomPoint filter     src/unionFeatureExtract.cpp:985-998
//   label split        EstimateLidarPose label split  src/lio/Estimator.cpp:992-1011
//   voxel filter       pcl::VoxelGrid::filter         src/lio/Estimator.cpp:1015-1024
// Compiled with -fmad=false: float32 centroids and voxel indices are bit-identical to the
// CPU oracle; undistortion differs from it only through libm vs CUDA sin/acos (<= 1 ulp
// of float32 after rounding).
#include "common.cuh"
#include "sort.cuh"
#include "undistort.cuh"
#include <math.h>

namespace mml {

__global__ void __launch_bounds__(256) k_undistort(float4* __restrict__ pts, const float* __restrict__ s_arr, int n,
                                                   UndistortParams P) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  pts[i] = undistort_point(pts[i], (double)s_arr[i], P);
}

// ---- A2 -------------------------------------------------------------------------------
struct VeloOri { float startOri, endOri; };

// pass 1: ring id; first kept index at which the sweep passes half (FE.cpp:1175-1176)
__global__ void __launch_bounds__(256) k_velo_ring(const float4* __restrict__ pts, int n, VeloOri o,
                                                   int16_t* __restrict__ ring_out, int* __restrict__ half_idx) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  const float angle = (float)(atan((double)(p.z / sqrtf(p.x * p.x + p.y * p.y))) * 180 / M_PI);
  const int scanID = int((angle + 15) / 2 + 0.5);
  if (scanID > 15 || scanID < 0) {
    ring_out[i] = -1;
    return;
  }
  ring_out[i] = (int16_t)scanID;
  float ori = (float)(-atan2((double)p.y, (double)p.x));
  if (ori < o.startOri - M_PI / 2) ori += 2 * M_PI;
  else if (ori > o.startOri + M_PI * 3 / 2) ori -= 2 * M_PI;
  if (ori - o.startOri > M_PI) atomicMin(half_idx, i);
}

__global__ void __launch_bounds__(256) k_velo_time(const float4* __restrict__ pts, int n, VeloOri o,
                                                   const int16_t* __restrict__ ring, const int* __restrict__ half_idx,
                                                   float* __restrict__ reltime) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  if (ring[i] < 0) {
    reltime[i] = 0.f;
    return;
  }
  const float4 p = pts[i];
  float ori = (float)(-atan2((double)p.y, (double)p.x));
  if (i <= *half_idx) {
    if (ori < o.startOri - M_PI / 2) ori += 2 * M_PI;
    else if (ori > o.startOri + M_PI * 3 / 2) ori -= 2 * M_PI;
  } else {
    ori += 2 * M_PI;
    if (ori < o.endOri - M_PI * 3 / 2) ori += 2 * M_PI;
    else if (ori > o.endOri + M_PI / 2) ori -= 2 * M_PI;
  }
  reltime[i] = (ori - o.startOri) / (o.endOri - o.startOri);
}

// ---- A3 -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_hori_filter(const uint32_t* __restrict__ off, const float* __restrict__ xyz,
                                                     const uint8_t* __restrict__ line, int n, double time_span,
                                                     uint8_t* __restrict__ keep, float* __restrict__ reltime) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const bool k = !((int)line[i] > 5) && !((double)xyz[3 * i] < 0.01);
  keep[i] = k ? 1 : 0;
  float rt = 0.f;
  if (k) {
    const uint32_t t = off[i];
    const double sec = (double)(t / 1000000000u) + 1e-9 * (double)(t % 1000000000u);
    rt = (float)(sec / time_span);
  }
  reltime[i] = rt;
}

// ---- label split: stable compaction of label==1 and label==2 ---------------------------
__global__ void __launch_bounds__(256) k_label_flags(const uint8_t* __restrict__ label, int n, int* __restrict__ f1,
                                                     int* __restrict__ f2) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint8_t l = label[i];
  f1[i] = (l == 1);
  f2[i] = (l == 2);
}
__global__ void __launch_bounds__(256) k_label_scatter(const float4* __restrict__ pts, const uint8_t* __restrict__ label,
                                                       int n, const int* __restrict__ o1, const int* __restrict__ o2,
                                                       float4* __restrict__ corner, float4* __restrict__ surf) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint8_t l = label[i];
  if (l == 1) corner[o1[i]] = pts[i];
  else if (l == 2) surf[o2[i]] = pts[i];
}

// ---- voxel filter ----------------------------------------------------------------------
// bbox[0..2] = min, bbox[3..5] = max as order-preserving unsigned encodings of float
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_vox_bbox_init(unsigned* bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) bbox[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) k_vox_bbox(const float4* __restrict__ pts, const int* __restrict__ n_dev, int n_host,
                                                  unsigned* __restrict__ bbox) {
  const int n = n_dev ? *n_dev : n_host;
  unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const float4 p = pts[i];
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) continue;
    const unsigned e[3] = {f2ord(p.x), f2ord(p.y), f2ord(p.z)};
#pragma unroll
    for (int c = 0; c < 3; c++) {
      mn[c] = min(mn[c], e[c]);
      mx[c] = max(mx[c], e[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; c++) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mn[c] = min(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], d));
      mx[c] = max(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], d));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      atomicMin(&bbox[c], mn[c]);
      atomicMax(&bbox[3 + c], mx[c]);
    }
  }
}

// keys: PCL's linear voxel index; vals: point index. status[0] = 1 when the grid would
// overflow int32 (PCL then passes the cloud through unchanged).
__global__ void __launch_bounds__(256) k_vox_keys(const float4* __restrict__ pts, const int* __restrict__ n_dev,
                                                  const unsigned* __restrict__ bbox, float inv, unsigned* __restrict__ keys,
                                                  unsigned* __restrict__ vals, int* __restrict__ status) {
  const int n = *n_dev;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float mnx = ord2f(bbox[0]), mny = ord2f(bbox[1]), mnz = ord2f(bbox[2]);
  const float mxx = ord2f(bbox[3]), mxy = ord2f(bbox[4]), mxz = ord2f(bbox[5]);
  const long long dx = (long long)((mxx - mnx) * inv) + 1;
  const long long dy = (long long)((mxy - mny) * inv) + 1;
  const long long dz = (long long)((mxz - mnz) * inv) + 1;
  if (dx * dy * dz > 2147483647LL) {
    if (i == 0) status[0] = 1;
    keys[i] = (unsigned)i;
    vals[i] = (unsigned)i;
    return;
  }
  if (i == 0) status[0] = 0;
  const int minb0 = (int)floorf(mnx * inv), minb1 = (int)floorf(mny * inv), minb2 = (int)floorf(mnz * inv);
  const int maxb0 = (int)floorf(mxx * inv), maxb1 = (int)floorf(mxy * inv);
  const int div0 = maxb0 - minb0 + 1, div1 = maxb1 - minb1 + 1;
  const float4 p = pts[i];
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) {  // dropped like PCL drops them from a non-dense cloud: sorted last, never a head
    keys[i] = 0xFFFFFFFFu;
    vals[i] = (unsigned)i;
    return;
  }
  const int i0 = (int)(floorf(p.x * inv) - (float)minb0);
  const int i1 = (int)(floorf(p.y * inv) - (float)minb1);
  const int i2 = (int)(floorf(p.z * inv) - (float)minb2);
  keys[i] = (unsigned)(i0 * 1 + i1 * div0 + i2 * (div0 * div1));
  vals[i] = (unsigned)i;
}

__global__ void __launch_bounds__(256) k_vox_heads(const unsigned* __restrict__ keys, const int* __restrict__ n_dev,
                                                   int* __restrict__ head) {
  const int n = *n_dev;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  head[i] = (keys[i] != 0xFFFFFFFFu && (i == 0 || keys[i] != keys[i - 1])) ? 1 : 0;
}

// one thread per voxel head: sum members in sorted (= ascending input index) order
__global__ void __launch_bounds__(256) k_vox_centroid(const float4* __restrict__ pts, const unsigned* __restrict__ keys,
                                                      const unsigned* __restrict__ vals, const int* __restrict__ n_dev,
                                                      const int* __restrict__ out_pos, float4* __restrict__ out) {
  const int n = *n_dev;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const unsigned key = keys[i];
  if (key == 0xFFFFFFFFu || (i > 0 && keys[i - 1] == key)) return;
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  int j = i;
  for (; j < n && keys[j] == key; j++) {
    const float4 p = pts[vals[j]];
    sx += p.x; sy += p.y; sz += p.z; si += p.w;
  }
  const float cnt = (float)(j - i);
  out[out_pos[i]] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
}

}  // namespace mml

using namespace mml;

int mml_undistort_device(mml_ctx* ctx, float4* pts_d, const float* s_d, int n, const double* dR9, const double* dt3) {
  if (n <= 0) return MML_OK;
  const UndistortParams P = make_undistort_params(dR9, dt3);
  k_undistort<<<div_up(n, 256), 256, 0, ctx->stream>>>(pts_d, s_d, n, P);
  MML_LAUNCHED(ctx);
  MML_CUDA(ctx, cudaGetLastError());
  return MML_OK;
}

// Split labelled points into corner / surf clouds (stable). Counts land in cnt_d[0], cnt_d[1].
int mml_label_split_device(mml_ctx* ctx, const float4* pts_d, const uint8_t* label_d, int n, float4* corner_d,
                           float4* surf_d, int* cnt_d) {
  cudaStream_t st = ctx->stream;
  if (n <= 0) {
    MML_CUDA(ctx, cudaMemsetAsync(cnt_d, 0, 2 * sizeof(int), st));
    return MML_OK;
  }
  MML_CUDA(ctx, ctx->tmp_a.reserve(sizeof(int) * (size_t)n));
  MML_CUDA(ctx, ctx->tmp_b.reserve(sizeof(int) * (size_t)n));
  int* f1 = ctx->tmp_a.as<int>();
  int* f2 = ctx->tmp_b.as<int>();
  k_label_flags<<<div_up(n, 256), 256, 0, st>>>(label_d, n, f1, f2);
  MML_LAUNCHED(ctx);
  MML_CHECK(exclusive_scan_device(ctx, f1, nullptr, n, cnt_d));
  MML_CHECK(exclusive_scan_device(ctx, f2, nullptr, n, cnt_d + 1));
  k_label_scatter<<<div_up(n, 256), 256, 0, st>>>(pts_d, label_d, n, f1, f2, corner_d, surf_d);
  MML_LAUNCHED(ctx);
  MML_CUDA(ctx, cudaGetLastError());
  return MML_OK;
}

// Voxel-grid filter of n = *n_dev (<= n_max) points. out_d capacity n_max; *m_dev = count.
int mml_voxel_device(mml_ctx* ctx, const float4* pts_d, const int* n_dev, int n_max, float leaf, float4* out_d,
                     int* m_dev) {
  cudaStream_t st = ctx->stream;
  if (n_max <= 0) {
    MML_CUDA(ctx, cudaMemsetAsync(m_dev, 0, sizeof(int), st));
    return MML_OK;
  }
  const int nblocks = div_up(n_max, kRadixTile);
  for (int k = 0; k < 2; k++) {
    MML_CUDA(ctx, ctx->vox_keys[k].reserve(sizeof(unsigned) * (size_t)n_max));
    MML_CUDA(ctx, ctx->vox_vals[k].reserve(sizeof(unsigned) * (size_t)n_max));
  }
  MML_CUDA(ctx, ctx->vox_hist.reserve(sizeof(int) * (256 * (size_t)nblocks + (size_t)n_max + 16)));
  MML_CUDA(ctx, ctx->vox_bbox.reserve(64));
  unsigned* keys[2] = {ctx->vox_keys[0].as<unsigned>(), ctx->vox_keys[1].as<unsigned>()};
  unsigned* vals[2] = {ctx->vox_vals[0].as<unsigned>(), ctx->vox_vals[1].as<unsigned>()};
  int* hist = ctx->vox_hist.as<int>();
  int* head = hist + 256 * (size_t)nblocks;
  unsigned* bbox = ctx->vox_bbox.as<unsigned>();
  int* status = reinterpret_cast<int*>(bbox + 8);
  const float inv = 1.0f / leaf;
  const int g = div_up(n_max, 256);
  k_vox_bbox_init<<<1, 32, 0, st>>>(bbox);
  MML_LAUNCHED(ctx);
  k_vox_bbox<<<g < 4 * kNumSMs ? g : 4 * kNumSMs, 256, 0, st>>>(pts_d, n_dev, 0, bbox);
  MML_LAUNCHED(ctx);
  k_vox_keys<<<g, 256, 0, st>>>(pts_d, n_dev, bbox, inv, keys[0], vals[0], status);
  MML_LAUNCHED(ctx);
  MML_CHECK(radix_sort_pairs(ctx, keys, vals, n_dev, n_max, hist));
  k_vox_heads<<<g, 256, 0, st>>>(keys[0], n_dev, head);
  MML_LAUNCHED(ctx);
  MML_CHECK(exclusive_scan_device(ctx, head, n_dev, n_max, m_dev));
  MML_LAUNCHED(ctx);
  k_vox_centroid<<<g, 256, 0, st>>>(pts_d, keys[0], vals[0], n_dev, head, out_d);
  MML_LAUNCHED(ctx);
  MML_CUDA(ctx, cudaGetLastError());
  return MML_OK;
}

int mml_velo_ring_time_device(mml_ctx* ctx, const float4* pts_d, int n, const float* first_last_xy, int16_t* ring_d,
                              float* reltime_d) {
  if (n <= 0) return MML_OK;
  // FE.cpp:1136-1145 on the first / last point (host, libm like the reference)
  VeloOri o;
  o.startOri = (float)(-atan2((double)first_last_xy[1], (double)first_last_xy[0]));
  o.endOri = (float)(-atan2((double)first_last_xy[3], (double)first_last_xy[2]) + 2 * M_PI);
  if (o.endOri - o.startOri > 3 * M_PI) o.endOri -= 2 * M_PI;
  else if (o.endOri - o.startOri < M_PI) o.endOri += 2 * M_PI;
  MML_CUDA(ctx, ctx->tmp_c.reserve(64));
  int* half = ctx->tmp_c.as<int>();
  MML_CUDA(ctx, cudaMemsetAsync(half, 0x7f, sizeof(int), ctx->stream));
  k_velo_ring<<<div_up(n, 256), 256, 0, ctx->stream>>>(pts_d, n, o, ring_d, half);
  MML_LAUNCHED(ctx);
  k_velo_time<<<div_up(n, 256), 256, 0, ctx->stream>>>(pts_d, n, o, ring_d, half, reltime_d);
  MML_LAUNCHED(ctx);
  MML_CUDA(ctx, cudaGetLastError());
  return MML_OK;
}

int mml_hori_filter_device(mml_ctx* ctx, const uint32_t* off_d, const float* xyz_d, const uint8_t* line_d, int n,
                           uint32_t last_offset, uint8_t* keep_d, float* reltime_d) {
  if (n <= 0) return MML_OK;
  const double span = (double)(last_offset / 1000000000u) + 1e-9 * (double)(last_offset % 1000000000u);
  k_hori_filter<<<div_up(n, 256), 256, 0, ctx->stream>>>(off_d, xyz_d, line_d, n, span, keep_d, reltime_d);
  MML_LAUNCHED(ctx);
  MML_CUDA(ctx, cudaGetLastError());
  return MML_OK;
}

// bounding box of n device points, returned to the host (synchronises the stream)
int mml_bbox_device(mml_ctx* ctx, const float4* pts_d, int n, float* mn3, float* mx3) {
  cudaStream_t st = ctx->stream;
  MML_CUDA(ctx, ctx->vox_bbox.reserve(64));
  unsigned* bbox = ctx->vox_bbox.as<unsigned>();
  const int g = div_up(n, 256);
  k_vox_bbox_init<<<1, 32, 0, st>>>(bbox);
  MML_LAUNCHED(ctx);
  k_vox_bbox<<<g < 4 * kNumSMs ? g : 4 * kNumSMs, 256, 0, st>>>(pts_d, nullptr, n, bbox);
  MML_LAUNCHED(ctx);
  unsigned h[6];
  MML_CUDA(ctx, cudaMemcpyAsync(h, bbox, sizeof(h), cudaMemcpyDeviceToHost, st));
  MML_CUDA(ctx, cudaStreamSynchronize(st));
  for (int c = 0; c < 3; c++) {
    unsigned u = h[c], v = h[3 + c];
    unsigned a = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    unsigned b = (v & 0x80000000u) ? (v & 0x7fffffffu) : ~v;
    memcpy(&mn3[c], &a, 4);
    memcpy(&mx3[c], &b, 4);
  }
  return MML_OK;
}
