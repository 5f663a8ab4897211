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
constexpr int kSvCluster = 8;      // CTAs per label in the chained loop's launch (portable cluster size)

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
  // labelled indices already compacted by k_label_compact on the extraction stream (pipelined loop): step 1 is skipped
  const int* pre_idx[2];   // input-order indices of the points with label 1 / 2, capacity kSvCap each
  const int* pre_cnt;      // [2] raw labelled counts (may exceed kSvCap: overflow)
  unsigned* bbox_part;     // cluster launch: [2][CL][6] per-CTA bounding boxes
  unsigned long long* tl;  // MML_TIMELINE
};

__device__ __forceinline__ unsigned sv_f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sv_ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// Stable compaction of the indices i with label[i] == want, in input order. Every thread owns one contiguous,
// 16-byte aligned slice of the label array (one block scan in total). s_warp: 32 ints of shared scratch.
template <class Store>
__device__ __forceinline__ void compact_labels_impl(const uint8_t* __restrict__ label, int n, uint8_t want, int* s_warp,
                                                    int* s_total, Store store) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    const int slice = ((n + kSvThreads - 1) / kSvThreads + 15) & ~15;
    const int b0 = min(tid * slice, n), b1 = min(b0 + slice, n);
    int mine = 0;
    for (int i = b0; i < b1; i += 16) {
      if (i + 16 <= b1) {
        const uint4 v = *reinterpret_cast<const uint4*>(label + i);
        const unsigned wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
          for (int b = 0; b < 4; b++) mine += ((wds[q] >> (8 * b)) & 0xffu) == want;
      } else {
        for (int k = i; k < b1; k++) mine += label[k] == want;
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
      if (lane == 31) *s_total = xs;
    }
    __syncthreads();
    // compact source indices first (they live in the key buffer until the keys are built) ...
    int pos = s_warp[warp] + (x - mine);
    if (mine) {
      for (int i = b0; i < b1; i++) {
        if (label[i] != want) continue;
        if (pos < kSvCap) store(pos, i);
        pos++;
      }
    }
    __syncthreads();
  }
}
__device__ __forceinline__ void compact_labels(const uint8_t* label, int n, uint8_t want, unsigned long long* keys, int* s_warp,
                                               int* s_total) {
  compact_labels_impl(label, n, want, s_warp, s_total, [&](int pos, int i) { keys[pos] = (unsigned long long)(unsigned)i; });
}

// Pipelined loop: the same compaction as its own launch on the extraction stream (one CTA per label), so that the
// matcher's critical path starts at the undistortion.
__global__ void __launch_bounds__(kSvThreads) k_label_compact(const uint8_t* __restrict__ label, int n, int* __restrict__ idx0,
                                                              int* __restrict__ idx1, int* __restrict__ cnt) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int kind = blockIdx.x;
  int* idx = kind == 0 ? idx0 : idx1;
  compact_labels_impl(label, n, (uint8_t)(kind + 1), s_warp, &s_total, [&](int pos, int i) { idx[pos] = i; });
  if (threadIdx.x == 0) cnt[kind] = s_total;
}

// ---- thread-block cluster primitives and the register-resident sort of the fast path
__device__ __forceinline__ unsigned sv_cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void sv_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned long long sv_shfl_xor(unsigned long long v, int m) {
  unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
  lo = __shfl_xor_sync(0xffffffffu, lo, m);
  hi = __shfl_xor_sync(0xffffffffu, hi, m);
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long sv_min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned long long sv_max(unsigned long long a, unsigned long long b) { return a < b ? b : a; }

constexpr int kSvFast = 2 * kSvThreads;  // labelled points of one kind the register sort handles

// Ascending bitonic sort of 2048 keys held two per thread (a0 = element tid, a1 = element tid + 1024). Partners at
// distance < 32 are reached by warp shuffles, the others through shared memory (two buffers used alternately, so a
// step needs one block barrier); the distance-1024 step pairs a thread's own two keys. 20 barriers instead of 66.
__device__ __forceinline__ void sv_bitonic2048(unsigned long long& a0, unsigned long long& a1, unsigned long long* bufA,
                                               unsigned long long* bufB, int tid) {
  int flip = 0;
  for (int size = 2; size <= kSvFast; size <<= 1) {
    for (int j = size >> 1; j > 0; j >>= 1) {
      if (j == kSvThreads) {  // size == 2048: ascending
        const unsigned long long lo = sv_min(a0, a1), hi = sv_max(a0, a1);
        a0 = lo;
        a1 = hi;
        continue;
      }
      unsigned long long p0, p1;
      if (j >= 32) {
        unsigned long long* b = flip ? bufB : bufA;
        flip ^= 1;
        b[tid] = a0;
        b[tid + kSvThreads] = a1;
        __syncthreads();
        p0 = b[tid ^ j];
        p1 = b[(tid ^ j) + kSvThreads];
      } else {
        p0 = sv_shfl_xor(a0, j);
        p1 = sv_shfl_xor(a1, j);
      }
      const bool lower = (tid & j) == 0;
      const bool up0 = (tid & size) == 0, up1 = ((tid + kSvThreads) & size) == 0;
      a0 = (lower == up0) ? sv_min(a0, p0) : sv_max(a0, p0);
      a1 = (lower == up1) ? sv_min(a1, p1) : sv_max(a1, p1);
    }
  }
}

// The same network on 32-bit keys (voxel << 11 | position), one or two per thread, for N2 <= 2048 keys: a compare-
// exchange is one shuffle (or one shared-memory word) and an integer min / max. Threads past N2 carry padding.
// N2 is a template parameter and both loops are fully unrolled: with 32 warps on one SM the network is bound by
// instruction issue, so the loop bookkeeping and the direction predicates must not cost instructions per step.
template <int N2>
__device__ __forceinline__ void sv_bitonic32_n(unsigned& a0, unsigned& a1, unsigned* bufA, unsigned* bufB, int tid) {
  constexpr bool two = N2 > kSvThreads;
  int flip = 0;
#pragma unroll
  for (int size = 2; size <= N2; size <<= 1) {
#pragma unroll
    for (int j = size >> 1; j > 0; j >>= 1) {
      if (j == kSvThreads) {  // size == 2048: ascending
        const unsigned lo = min(a0, a1), hi = max(a0, a1);
        a0 = lo;
        a1 = hi;
        continue;
      }
      unsigned p0, p1 = 0xffffffffu;
      if (j >= 32) {
        unsigned* b = flip ? bufB : bufA;
        flip ^= 1;
        b[tid] = a0;
        if (two) b[tid + kSvThreads] = a1;
        __syncthreads();
        p0 = b[tid ^ j];
        if (two) p1 = b[(tid ^ j) + kSvThreads];
      } else {
        p0 = __shfl_xor_sync(0xffffffffu, a0, j);
        if (two) p1 = __shfl_xor_sync(0xffffffffu, a1, j);
      }
      const bool lower = (tid & j) == 0;
      const bool up0 = (tid & size) == 0, up1 = ((tid + kSvThreads) & size) == 0;
      a0 = (lower == up0) ? min(a0, p0) : max(a0, p0);
      if (two) a1 = (lower == up1) ? min(a1, p1) : max(a1, p1);
    }
  }
}
__device__ __forceinline__ void sv_bitonic32(unsigned& a0, unsigned& a1, int N2, unsigned* bufA, unsigned* bufB, int tid) {
  switch (N2) {  // block-uniform
    case 64: sv_bitonic32_n<64>(a0, a1, bufA, bufB, tid); break;
    case 128: sv_bitonic32_n<128>(a0, a1, bufA, bufB, tid); break;
    case 256: sv_bitonic32_n<256>(a0, a1, bufA, bufB, tid); break;
    case 512: sv_bitonic32_n<512>(a0, a1, bufA, bufB, tid); break;
    case 1024: sv_bitonic32_n<1024>(a0, a1, bufA, bufB, tid); break;
    default: sv_bitonic32_n<2048>(a0, a1, bufA, bufB, tid); break;
  }
}

// CL = 1: one CTA per label does everything (per-scan API, host-driven loop).
// CL > 1 (chained loop, labels pre-compacted): a cluster of CL CTAs per label shares the float64 undistortion - one SM's
// FP64 pipe would bound it - then the cluster's first CTA builds the voxel grid; CTA 1 of the corner cluster starts
// the scan's solve (k_est_init's job) off the voxel filter's critical path.
#ifndef MML_SV_MINB
#define MML_SV_MINB 1
#endif
template <int CL>
__global__ void __launch_bounds__(kSvThreads, MML_SV_MINB) k_split_voxel(SplitVoxelArgs A) {
  extern __shared__ __align__(16) unsigned char sv_smem[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(sv_smem);  // [kSvCap]; fast path: [2][2048] + points
  float4* spts = reinterpret_cast<float4*>(sv_smem + 2 * kSvFast * sizeof(unsigned long long));  // fast path: [2048]
  __shared__ int s_warp[32];
  __shared__ int s_base, s_total;
  __shared__ unsigned s_bbox[6];
  const int kind = blockIdx.x / CL;  // 0: label 1 (corner), 1: label 2 (surf)
  const unsigned rank = CL > 1 ? sv_cluster_rank() : 0u;
  const uint8_t want = (uint8_t)(kind + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float4* scratch = A.scratch[kind];
  if (blockIdx.x == CL && tid == 0) MML_TL(A.tl, 0);
#ifdef MML_SV_PROF
  long long tp[6] = {0, 0, 0, 0, 0, 0}, t0 = clock64(), t1;
#define SV_TICK(k) { t1 = clock64(); tp[k] += t1 - t0; t0 = t1; }
#else
#define SV_TICK(k)
#endif
  __shared__ UndistortParams sU;
  const UndistortParams* Up = &A.U;
  if (A.od) {
    Up = &sU;
    const bool starter = kind == 0 && rank == (CL > 1 ? 1u : 0u);
    if (starter) {  // fresh solver state and statistics (k_est_init)
      unsigned* w = reinterpret_cast<unsigned*>(A.est);
      for (int i = tid; i < (int)(sizeof(EstState) / 4); i += kSvThreads) w[i] = 0u;
      if (tid < 128) A.assoc_stats_words[tid] = 0u;
      if (tid < 64) A.acc_out_words[tid] = 0u;
      __syncthreads();
    }
    if (tid == 0 || (starter && tid == 32)) {
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
  unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
  bool overflow = false;
  const int* pre = A.pre_idx[kind];
  if (pre) {
    if (tid == 0) s_base = A.pre_cnt[kind];
    __syncthreads();
  } else {
    compact_labels(A.label, A.n, want, keys, s_warp, &s_base);  // CL == 1 only (the host never launches a cluster without pre)
  }
  int cnt = s_base;
  if (cnt > kSvCap) { overflow = true; cnt = kSvCap; }
  SV_TICK(0)
  // ... then undistort with the points dealt out evenly over the CTA(s) (the labelled points cluster, so
  // doing this inside the slice loop would leave the whole CTA waiting for a few busy lanes)
  for (int k = (int)rank * kSvThreads + tid; k < cnt; k += CL * kSvThreads) {
    const int i = pre ? pre[k] : (int)keys[k];
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
  if (CL > 1) {
    // per-CTA boxes through global memory; the cluster barrier also publishes every CTA's slice of `scratch`
    unsigned* part = A.bbox_part + (size_t)kind * CL * 6;
    if (tid < 6) part[rank * 6 + tid] = s_bbox[tid];
    sv_cluster_sync();
    if (rank != 0) return;
    if (tid < 6) {
      unsigned v = tid < 3 ? 0xffffffffu : 0u;
      for (int r = 0; r < CL; r++) {
        const unsigned o = __ldcg(part + r * 6 + tid);
        v = tid < 3 ? min(v, o) : max(v, o);
      }
      s_bbox[tid] = v;
    }
    __syncthreads();
  }
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

  SV_TICK(1)
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
  auto key_of = [&](int k, const float4& p) -> unsigned long long {
    unsigned vox = (unsigned)k;
    if (!passthrough) {
      const int i0 = (int)(floorf(p.x * inv) - (float)minb0);
      const int i1 = (int)(floorf(p.y * inv) - (float)minb1);
      const int i2 = (int)(floorf(p.z * inv) - (float)minb2);
      vox = (unsigned)(i0 * 1 + i1 * div0 + i2 * (div0 * div1));
    }
    return ((unsigned long long)vox << 32) | (unsigned)k;
  };
  const bool fast = cnt <= kSvFast;  // block-uniform
  if (fast && !passthrough && dx * dy * dz <= (1LL << 21)) {
    // ---- 3a. 32-bit keys (voxel << 11 | position) in registers, points staged in shared memory
    unsigned a0 = 0xffffffffu, a1 = 0xffffffffu;
    if (tid < cnt) { const float4 p = scratch[tid]; spts[tid] = p; a0 = ((unsigned)(key_of(tid, p) >> 32) << 11) | (unsigned)tid; }
    if (tid + kSvThreads < cnt) {
      const float4 p = scratch[tid + kSvThreads];
      spts[tid + kSvThreads] = p;
      a1 = ((unsigned)(key_of(tid + kSvThreads, p) >> 32) << 11) | (unsigned)(tid + kSvThreads);
    }
    int N2 = 64;
    while (N2 < cnt) N2 <<= 1;
    SV_TICK(2)
    unsigned* ex = reinterpret_cast<unsigned*>(sv_smem + 4 * kSvFast * sizeof(unsigned long long));  // past keys + points
    sv_bitonic32(a0, a1, N2, ex, ex + kSvFast, tid);
    keys[tid] = ((unsigned long long)(a0 >> 11) << 32) | (a0 & 2047u);
    keys[tid + kSvThreads] = ((unsigned long long)(a1 >> 11) << 32) | (a1 & 2047u);
    __syncthreads();
  } else if (fast) {
    // ---- 3a'. (voxel grids beyond 2^21 cells) 64-bit keys in registers
    unsigned long long a0 = ~0ull, a1 = ~0ull;
    if (tid < cnt) { const float4 p = scratch[tid]; spts[tid] = p; a0 = key_of(tid, p); }
    if (tid + kSvThreads < cnt) { const float4 p = scratch[tid + kSvThreads]; spts[tid + kSvThreads] = p; a1 = key_of(tid + kSvThreads, p); }
    SV_TICK(2)
    sv_bitonic2048(a0, a1, keys, keys + kSvFast, tid);
    __syncthreads();  // the last exchange buffer may still be read by a slower warp
    keys[tid] = a0;
    keys[tid + kSvThreads] = a1;
    __syncthreads();
  } else {
    int N2 = 1;
    while (N2 < cnt) N2 <<= 1;
    for (int k = tid; k < N2; k += kSvThreads) keys[k] = k < cnt ? key_of(k, scratch[k]) : ~0ull;
    __syncthreads();
    SV_TICK(2)
    // ---- 3b. bitonic sort in shared memory, ascending
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
  }

  SV_TICK(3)
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
  SV_TICK(4)
  int pos = s_warp[warp] + (x - heads);
  float4* out = A.out[kind];
  for (int k = k0; k < k1; k++) {
    const unsigned vox = (unsigned)(keys[k] >> 32);
    if (!(k == 0 || vox != (unsigned)(keys[k - 1] >> 32))) continue;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    int j = k;
    for (; j < cnt && (unsigned)(keys[j] >> 32) == vox; j++) {
      const unsigned m = (unsigned)(keys[j] & 0xffffffffu);
      const float4 p = fast ? spts[m] : scratch[m];
      sx += p.x; sy += p.y; sz += p.z; si += p.w;
    }
    const float c = (float)(j - k);
    out[pos++] = make_float4(sx / c, sy / c, sz / c, si / c);
  }
  if (tid == 0) {
    A.counts[kind] = s_total;
    if (A.counts_out) A.counts_out[2 + kind] = s_total;
    if (kind == 1) MML_TL(A.tl, 1);
  }
  SV_TICK(5)
#ifdef MML_SV_PROF
  if (tid == 0) printf("split_voxel kind %d cnt %d: init %lld undistort+bbox %lld keys %lld sort %lld heads %lld centroids %lld cycles\n",
                       kind, cnt, tp[0], tp[1], tp[2], tp[3], tp[4], tp[5]);
#endif
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
    A.pre_idx[0] = chain->pre_idx[0];
    A.pre_idx[1] = chain->pre_idx[1];
    A.pre_cnt = chain->pre_cnt;
    A.tl = ctx->timeline.as<unsigned long long>();
  } else {
    MML_CUDA(ctx, cudaMemsetAsync(counts_d + 4, 0, sizeof(int), st));
  }
  const size_t smem = sizeof(unsigned long long) * (size_t)kSvCap;
  static bool attr_set = false;
  if (!attr_set) {
    MML_CUDA(ctx, cudaFuncSetAttribute(k_split_voxel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MML_CUDA(ctx, cudaFuncSetAttribute(k_split_voxel<kSvCluster>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  if (chain && chain->pre_idx[0]) {
    MML_CUDA(ctx, ctx->sv_bbox.reserve(sizeof(unsigned) * 2 * kSvCluster * 6));
    A.bbox_part = ctx->sv_bbox.as<unsigned>();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * kSvCluster);
    cfg.blockDim = dim3(kSvThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kSvCluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    MML_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_split_voxel<kSvCluster>, A));
  } else {
    k_split_voxel<1><<<2, kSvThreads, smem, st>>>(A);
  }
  MML_LAUNCHED(ctx);
  MML_CUDA(ctx, cudaGetLastError());
  return MML_OK;
}

// label compaction on the caller's stream (the extraction stream of the pipelined loop): idx0/idx1 have capacity
// mml_split_voxel_capacity(), cnt_d = int[2] raw labelled counts
int mml_label_compact_device(mml_ctx* ctx, const uint8_t* label_d, int n, int* idx0, int* idx1, int* cnt_d) {
  k_label_compact<<<2, kSvThreads, 0, ctx->stream>>>(label_d, n, idx0, idx1, cnt_d);
  MML_LAUNCHED(ctx);
  MML_CUDA(ctx, cudaGetLastError());
  return MML_OK;
}
