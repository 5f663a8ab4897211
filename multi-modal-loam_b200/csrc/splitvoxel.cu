// Fused label split + undistortion + voxel-grid filter: ONE launch, two CTAs (corner / surf).
//
// Replaces, for the per-scan path, EstimateLidarPose's label split and pcl::VoxelGrid::filter
// (src/lio/Estimator.cpp:992-1026) plus RemoveLidarDistortion (src/unionPoseEstimation.cpp:402-421)
// restricted to the labelled points (the only ones the Estimate loop reads). The general multi-kernel
// path (geometry.cu) stays for clouds with more than kSvCap labelled points of one kind.
//
// Per CTA (1024 threads, ~200 KB shared memory):
//   1. stable compaction of the points with this CTA's label (ballot scan), undistorted on the fly,
//      written to a compact HBM scratch; bounding box by block reduction
//   2. PCL's linear voxel index per point -> key64 = (voxel << 32) | compact position
//   3. bitonic sort of the keys in shared memory (ties impossible: positions are unique, so the order
//      inside a voxel is the input order, as the oracle defines it)
//   4. voxel heads -> block scan -> float32 centroid accumulated in input order -> output in voxel order
// Arithmetic is bit-identical to the multi-kernel path and to the CPU oracle (-fmad=false).
#include "common.cuh"
#include "undistort.cuh"
#include "eststate.cuh"

namespace mml {

constexpr int kSvCap = 16384;     // labelled points of one kind handled in shared memory
constexpr int kSvThreads = 1024;

struct SplitVoxelArgs {
  const float4* pts;      // scan (raw, input order)
  const float* s;         // sweep fraction (may be null when undistortion is disabled)
  const uint8_t* label;
  int n;
  UndistortParams U;
  float leaf[2];
  float4* scratch[2];     // compact undistorted labelled points, capacity kSvCap each
  float4* out[2];         // voxel centroids, capacity kSvCap each
  int* counts;            // [0..1] voxel output counts, [2..3] raw labelled counts, [4] overflow flag
  // chained odometry loop (odometry.cu): the motion used for undistortion and the start pose of the solve come
  // from the device-side pose history, and this launch also starts the scan's solve (k_est_init's job)
  const OdomDev* od;
  EstState* est;
  EstInit I;
  unsigned* assoc_stats_words;
  unsigned* acc_out_words;
  const int* fe_counters;  // the extraction's slot counters: n_sharp, n_flat, overflow
  int* counts_out;         // ChainOut.counts of this scan
};

__device__ __forceinline__ unsigned sv_f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sv_ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void __launch_bounds__(kSvThreads) k_split_voxel(SplitVoxelArgs A) {
  extern __shared__ __align__(16) unsigned char sv_smem[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(sv_smem);  // [kSvCap]
  __shared__ int s_warp[32];
  __shared__ int s_base, s_total;
  __shared__ unsigned s_bbox[6];
  const int kind = blockIdx.x;  // 0: label 1 (corner), 1: label 2 (surf)
  const uint8_t want = (uint8_t)(kind + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float4* scratch = A.scratch[kind];
  __shared__ UndistortParams sU;
  const UndistortParams* Up = &A.U;
  if (A.od) {
    Up = &sU;
    if (kind == 0) {  // fresh solver state and statistics (k_est_init)
      unsigned* w = reinterpret_cast<unsigned*>(A.est);
      for (int i = tid; i < (int)(sizeof(EstState) / 4); i += kSvThreads) w[i] = 0u;
      if (tid < 128) A.assoc_stats_words[tid] = 0u;
      if (tid < 64) A.acc_out_words[tid] = 0u;
    }
    __syncthreads();
    if (tid == 0 || (kind == 0 && tid == 32)) {
      // constant-velocity model: delta = T_before^-1 T_last, prediction = T_last delta (PE.cpp:847-852, 882-890)
      double Tinv[16], delta[16];
      rigid_inv(A.od->T_before, Tinv);
      mat4_mul(Tinv, A.od->T_last, delta);
      if (tid == 0) {
        const double dR[9] = {delta[0], delta[1], delta[2], delta[4], delta[5], delta[6], delta[8], delta[9], delta[10]};
        const double dt[3] = {delta[3], delta[7], delta[11]};
        sU = make_undistort_params(A.s ? dR : nullptr, A.s ? dt : nullptr);
      } else {
        double Tp[16];
        mat4_mul(A.od->T_last, delta, Tp);
        const double Rp[9] = {Tp[0], Tp[1], Tp[2], Tp[4], Tp[5], Tp[6], Tp[8], Tp[9], Tp[10]};
        const Quat qp = quat_from_R9(Rp);
        const double P[3] = {Tp[3], Tp[7], Tp[11]}, Q[4] = {qp.w, qp.x, qp.y, qp.z};
        est_fill(A.est, A.I, P, Q);
      }
    }
  }
  if (tid == 0) s_base = 0;
  if (tid < 3) s_bbox[tid] = 0xffffffffu;
  else if (tid < 6) s_bbox[tid] = 0u;
  __syncthreads();

  // ---- 1. stable compaction + undistortion + bounding box
  // every thread owns one contiguous, 16-byte aligned slice of the label array (one block scan in total)
  unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
  bool overflow = false;
  {
    const int slice = ((A.n + kSvThreads - 1) / kSvThreads + 15) & ~15;
    const int b0 = min(tid * slice, A.n), b1 = min(b0 + slice, A.n);
    int mine = 0;
    for (int i = b0; i < b1; i += 16) {
      if (i + 16 <= b1) {
        const uint4 v = *reinterpret_cast<const uint4*>(A.label + i);
        const unsigned wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
          for (int b = 0; b < 4; b++) mine += ((wds[q] >> (8 * b)) & 0xffu) == want;
      } else {
        for (int k = i; k < b1; k++) mine += A.label[k] == want;
      }
    }
    int x = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      const int v = s_warp[lane];
      int xs = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, xs, d);
        if (lane >= d) xs += y;
      }
      s_warp[lane] = xs - v;
      if (lane == 31) s_base = xs;
    }
    __syncthreads();
    // compact source indices first (they live in the key buffer until the keys are built) ...
    int pos = s_warp[warp] + (x - mine);
    if (mine) {
      for (int i = b0; i < b1; i++) {
        if (A.label[i] != want) continue;
        if (pos < kSvCap) keys[pos] = (unsigned long long)(unsigned)i;
        pos++;
      }
    }
    __syncthreads();
  }
  int cnt = s_base;
  if (cnt > kSvCap) { overflow = true; cnt = kSvCap; }
  // ... then undistort with the points dealt out evenly over the CTA (the labelled points cluster, so
  // doing this inside the slice loop would leave the whole CTA waiting for a few busy lanes)
  for (int k = tid; k < cnt; k += kSvThreads) {
    const int i = (int)keys[k];
    float4 p = A.pts[i];
    if (Up->enabled) p = undistort_point(p, (double)A.s[i], *Up);
    scratch[k] = p;
    const unsigned e[3] = {sv_f2ord(p.x), sv_f2ord(p.y), sv_f2ord(p.z)};
#pragma unroll
    for (int c = 0; c < 3; c++) { mn[c] = min(mn[c], e[c]); mx[c] = max(mx[c], e[c]); }
  }
#pragma unroll
  for (int c = 0; c < 3; c++) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mn[c] = min(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], d));
      mx[c] = max(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], d));
    }
    if (lane == 0) { atomicMin(&s_bbox[c], mn[c]); atomicMax(&s_bbox[3 + c], mx[c]); }
  }
  __threadfence_block();
  __syncthreads();
  if (tid == 0) {
    A.counts[2 + kind] = s_base;
    if (overflow) atomicExch(&A.counts[4], 1);
    if (A.counts_out) {
      if (overflow) A.counts_out[5] = 1;
      if (kind == 0) { A.counts_out[0] = A.fe_counters[0]; A.counts_out[1] = A.fe_counters[1]; A.counts_out[4] = A.fe_counters[2]; }
    }
  }
  if (cnt == 0) {
    if (tid == 0) {
      A.counts[kind] = 0;
      if (A.counts_out) A.counts_out[2 + kind] = 0;
    }
    return;
  }

  // ---- 2. keys (pcl::VoxelGrid::applyFilter index arithmetic)
  const float inv = 1.0f / A.leaf[kind];
  const float mnx = sv_ord2f(s_bbox[0]), mny = sv_ord2f(s_bbox[1]), mnz = sv_ord2f(s_bbox[2]);
  const float mxx = sv_ord2f(s_bbox[3]), mxy = sv_ord2f(s_bbox[4]), mxz = sv_ord2f(s_bbox[5]);
  const long long dx = (long long)((mxx - mnx) * inv) + 1;
  const long long dy = (long long)((mxy - mny) * inv) + 1;
  const long long dz = (long long)((mxz - mnz) * inv) + 1;
  const bool passthrough = dx * dy * dz > 2147483647LL;
  const int minb0 = (int)floorf(mnx * inv), minb1 = (int)floorf(mny * inv), minb2 = (int)floorf(mnz * inv);
  const int maxb0 = (int)floorf(mxx * inv), maxb1 = (int)floorf(mxy * inv);
  const int div0 = maxb0 - minb0 + 1, div1 = maxb1 - minb1 + 1;
  int N2 = 1;
  while (N2 < cnt) N2 <<= 1;
  for (int k = tid; k < N2; k += kSvThreads) {
    unsigned long long key = ~0ull;
    if (k < cnt) {
      unsigned vox = (unsigned)k;
      if (!passthrough) {
        const float4 p = scratch[k];
        const int i0 = (int)(floorf(p.x * inv) - (float)minb0);
        const int i1 = (int)(floorf(p.y * inv) - (float)minb1);
        const int i2 = (int)(floorf(p.z * inv) - (float)minb2);
        vox = (unsigned)(i0 * 1 + i1 * div0 + i2 * (div0 * div1));
      }
      key = ((unsigned long long)vox << 32) | (unsigned)k;
    }
    keys[k] = key;
  }
  __syncthreads();

  // ---- 3. bitonic sort, ascending
  for (int size = 2; size <= N2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (N2 >> 1); t += kSvThreads) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  }

  // ---- 4. heads -> exclusive scan -> centroids (members summed in input order)
  const int per = (cnt + kSvThreads - 1) / kSvThreads;
  const int k0 = tid * per, k1 = min(k0 + per, cnt);
  int heads = 0;
  for (int k = k0; k < k1; k++) heads += (k == 0 || (unsigned)(keys[k] >> 32) != (unsigned)(keys[k - 1] >> 32));
  int x = heads;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    const int v = s_warp[lane];
    int xs = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, xs, d);
      if (lane >= d) xs += y;
    }
    s_warp[lane] = xs - v;
    if (lane == 31) s_total = xs;
  }
  __syncthreads();
  int pos = s_warp[warp] + (x - heads);
  float4* out = A.out[kind];
  for (int k = k0; k < k1; k++) {
    const unsigned vox = (unsigned)(keys[k] >> 32);
    if (!(k == 0 || vox != (unsigned)(keys[k - 1] >> 32))) continue;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    int j = k;
    for (; j < cnt && (unsigned)(keys[j] >> 32) == vox; j++) {
      const float4 p = scratch[(unsigned)(keys[j] & 0xffffffffu)];
      sx += p.x; sy += p.y; sz += p.z; si += p.w;
    }
    const float c = (float)(j - k);
    out[pos++] = make_float4(sx / c, sy / c, sz / c, si / c);
  }
  if (tid == 0) {
    A.counts[kind] = s_total;
    if (A.counts_out) A.counts_out[2 + kind] = s_total;
  }
}

}  // namespace mml

using namespace mml;

int mml_split_voxel_capacity() { return kSvCap; }

// counts_d: int[5] = {n_corner_ds, n_surf_ds, n_corner_raw, n_surf_raw, overflow}
int mml_split_voxel_device(mml_ctx* ctx, const float4* pts_d, const float* s_d, const uint8_t* label_d, int n,
                           const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, float4* corner_out,
                           float4* surf_out, int* counts_d, const mml::SvChain* chain) {
  cudaStream_t st = ctx->stream;
  MML_CUDA(ctx, ctx->corner_raw.reserve(sizeof(float4) * (size_t)kSvCap));
  MML_CUDA(ctx, ctx->surf_raw.reserve(sizeof(float4) * (size_t)kSvCap));
  SplitVoxelArgs A;
  memset(&A, 0, sizeof(A));
  A.pts = pts_d;
  A.s = s_d;
  A.label = label_d;
  A.n = n;
  A.U = make_undistort_params(s_d ? dR9 : nullptr, s_d ? dt3 : nullptr);
  A.leaf[0] = leaf_corner;
  A.leaf[1] = leaf_surf;
  A.scratch[0] = ctx->corner_raw.as<float4>();
  A.scratch[1] = ctx->surf_raw.as<float4>();
  A.out[0] = corner_out;
  A.out[1] = surf_out;
  A.counts = counts_d;
  if (chain) {
    A.od = chain->od;
    A.est = chain->est;
    A.I = chain->I;
    A.assoc_stats_words = ctx->assoc_stats.as<unsigned>();
    A.acc_out_words = ctx->acc_out.as<unsigned>();
    A.fe_counters = chain->fe_counters;
    A.counts_out = chain->counts_out;
  } else {
    MML_CUDA(ctx, cudaMemsetAsync(counts_d + 4, 0, sizeof(int), st));
  }
  const size_t smem = sizeof(unsigned long long) * (size_t)kSvCap;
  static bool attr_set = false;
  if (!attr_set) {
    MML_CUDA(ctx, cudaFuncSetAttribute(k_split_voxel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  k_split_voxel<<<2, kSvThreads, smem, st>>>(A);
  MML_LAUNCHED(ctx);
  MML_CUDA(ctx, cudaGetLastError());
  return MML_OK;
}
