// Feature extraction on the device: every line of every scan in one pass.
//
// Replaces feature_extraction::detectFeaturePoints (reference
// mm-loam/src/unionFeatureExtract.cpp:341-844) and the split / label glue around it
// (FE.cpp:1001-1023, 1209-1240). Results are bit-identical to the CPU oracle: all float32
// and float64 expressions keep the reference's operand order and the file is compiled with
// -fmad=false (the reference is built for baseline x86-64, no FMA contraction).
//
// Pipeline (all on ctx->stream, no host synchronisation inside):
//   E1  k_line_hist / k_line_scan / k_line_scatter   stable split of each scan by line id
//   E2  k_point_attr      per point: curvature, reflectivity difference and a 16-bit
//                         attribute word holding every per-point test of FE.cpp:407-451,
//                         543-806 (the parts of the algorithm with no sequential dependence)
//   E3  k_part_sort       50 parts per line, stable rank-by-counting in shared memory
//                         (replaces the O(m^2) insertion sorts of FE.cpp:458-479)
//   E4  k_select          one CTA per line; the order-dependent flat selection
//                         (FE.cpp:483-539) and the count_num stride walk (FE.cpp:543-650)
//                         run on two lanes over shared-memory state; labels are scattered
//                         back to input order.
// HBM traffic per input point: 16 B xyzi + 2 B line read, 1 B label written (the
// algorithmic 19 B of SURVEY.md §8 d) plus the line-sorted working copy.
#include "common.cuh"

namespace mml {

// attribute bits written by k_point_attr
enum : unsigned {
  A_CAND = 1u << 0,   // curvature < (0.02*depth)^2          FE.cpp:488
  A_FAR = 1u << 1,    // depth > 50                           FE.cpp:499
  A_ANGLE = 1u << 2,  // both incidence cosines > 0.966       FE.cpp:430
  A_GAP = 1u << 3,    // |p[i+1]-p[i]|^2 > 0.02               FE.cpp:499,512
  A_C300 = 1u << 4,   // reflectivity pick test               FE.cpp:534-535
  A_RF = 1u << 5,     // right side flat -> stride 4          FE.cpp:597-609
  A_C150 = 1u << 6,   // two-plane corner test passes         FE.cpp:612-647
  A_BRK100 = 1u << 7, // break point, flag 100                FE.cpp:677-753
  A_BRK101 = 1u << 8, // break point demoted to 101           FE.cpp:756-804
  A_NEAR = 1u << 9,   // range^2 < 1                          FE.cpp:824
  A_W3 = 1u << 10,    // thNumCurvSize == 3 at this point     FE.cpp:424-428
};

struct Chunk { int scan, start, count, pad; };
constexpr int kChunkPts = 1024;
constexpr int kMaxLines = 64;

struct V3d { double x, y, z; };
// Vector3d - Vector3d of widened points (FE.cpp:417-422): float64 subtraction.
__device__ __forceinline__ V3d vsub(const float4& a, const float4& b) {
  return {(double)a.x - (double)b.x, (double)a.y - (double)b.y, (double)a.z - (double)b.z};
}
// Eigen::Vector3d(a.x - b.x, ...) (FE.cpp:618-620, 625-627, 635-640, 680-682, 717-719, 772-774, 790-792):
// float32 subtraction, widened afterwards (-fmad=false keeps the subtraction a plain FSUB).
__device__ __forceinline__ V3d vsubf(const float4& a, const float4& b) {
  return {(double)(a.x - b.x), (double)(a.y - b.y), (double)(a.z - b.z)};
}
__device__ __forceinline__ double vdot(const V3d& a, const V3d& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ double vnorm(const V3d& a) { return sqrt(vdot(a, a)); }
__device__ __forceinline__ void vnormalize(V3d& a) {
  double z = vdot(a, a);
  if (z > 0) {
    double s = sqrt(z);
    a.x /= s; a.y /= s; a.z /= s;
  }
}
__device__ __forceinline__ float range3(const float4& p) { return sqrtf(p.x * p.x + p.y * p.y + p.z * p.z); }

// ---------------------------------------------------------------- E1: split by line
__global__ void __launch_bounds__(256) k_line_hist(const uint16_t* __restrict__ line_id, const Chunk* __restrict__ chunks,
                                                   int n_lines, int* __restrict__ hist) {
  __shared__ int h[kMaxLines];
  const Chunk c = chunks[blockIdx.x];
  if (threadIdx.x < kMaxLines) h[threadIdx.x] = 0;
  __syncthreads();
  for (int r = threadIdx.x; r < c.count; r += 256) {
    int l = line_id[c.start + r];
    if (l < n_lines) atomicAdd(&h[l], 1);
  }
  __syncthreads();
  if (threadIdx.x < n_lines) hist[(size_t)blockIdx.x * n_lines + threadIdx.x] = h[threadIdx.x];
}

// one CTA per scan: exclusive scan of the chunk histograms per line, then line offsets
__global__ void __launch_bounds__(256) k_line_scan(int* __restrict__ hist, const int* __restrict__ scan_chunk0,
                                                   const int* __restrict__ scan_off, int n_lines,
                                                   int* __restrict__ line_start, int* __restrict__ line_count) {
  __shared__ int cnt[kMaxLines];
  const int s = blockIdx.x;
  const int c0 = scan_chunk0[s], c1 = scan_chunk0[s + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int l = warp; l < n_lines; l += 8) {
    int carry = 0;
    for (int base = c0; base < c1; base += 32) {
      int c = base + lane;
      int v = (c < c1) ? hist[(size_t)c * n_lines + l] : 0;
      int x = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
      }
      if (c < c1) hist[(size_t)c * n_lines + l] = carry + x - v;
      carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) cnt[l] = carry;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = scan_off[s];
    for (int l = 0; l < n_lines; l++) {
      line_start[s * n_lines + l] = run;
      line_count[s * n_lines + l] = cnt[l];
      run += cnt[l];
    }
  }
}

__global__ void __launch_bounds__(256) k_line_scatter(const float4* __restrict__ xyzi, const uint16_t* __restrict__ line_id,
                                                      const Chunk* __restrict__ chunks, const int* __restrict__ hist,
                                                      const int* __restrict__ line_start, int n_lines,
                                                      float4* __restrict__ srt_xyzi, int* __restrict__ srt_src,
                                                      int* __restrict__ srt_line) {
  __shared__ int wcnt[8][kMaxLines + 1];
  __shared__ int base[kMaxLines + 1];
  const Chunk c = chunks[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x <= kMaxLines) base[threadIdx.x] = 0;
  for (int r0 = 0; r0 < c.count; r0 += 256) {
    for (int k = threadIdx.x; k < 8 * (kMaxLines + 1); k += 256) (&wcnt[0][0])[k] = 0;
    __syncthreads();
    const int r = r0 + threadIdx.x;
    const bool act = r < c.count;
    int l = kMaxLines;
    if (act) {
      int li = line_id[c.start + r];
      if (li < n_lines) l = li;
    }
    unsigned mask = __match_any_sync(0xffffffffu, l);
    int rank = __popc(mask & ((1u << lane) - 1u));
    if (rank == 0) wcnt[warp][l] = __popc(mask);
    __syncthreads();
    if (act && l < n_lines) {
      int off = base[l] + rank;
      for (int w = 0; w < warp; w++) off += wcnt[w][l];
      int pos = line_start[c.scan * n_lines + l] + hist[(size_t)blockIdx.x * n_lines + l] + off;
      srt_xyzi[pos] = xyzi[c.start + r];
      srt_src[pos] = c.start + r;
      srt_line[pos] = c.scan * n_lines + l;
    }
    __syncthreads();
    if (threadIdx.x < n_lines) {
      int t = 0;
      for (int w = 0; w < 8; w++) t += wcnt[w][threadIdx.x];
      base[threadIdx.x] += t;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- E2: per-point tests
__global__ void __launch_bounds__(256) k_point_attr(const float4* __restrict__ P, const int* __restrict__ srt_line,
                                                    const int* __restrict__ line_start, const int* __restrict__ line_count,
                                                    int n_total, float* __restrict__ curv_out, float* __restrict__ refl_out,
                                                    uint16_t* __restrict__ attr_out) {
  const int g = blockIdx.x * 256 + threadIdx.x;
  if (g >= n_total) return;
  const int gl = srt_line[g];
  if (gl < 0) { attr_out[g] = 0; curv_out[g] = 0.f; refl_out[g] = 0.f; return; }
  const int ls = line_start[gl], n = line_count[gl];
  const int i = g - ls;
  const float4* p = P + ls;  // p[i] is the line-local point i
  unsigned a = 0;
  float curv = 0.f, refl = 0.f;
  const float4 pi = p[i];
  if (i + 1 < n) {
    const float4 q = p[i + 1];
    float dX = q.x - pi.x, dY = q.y - pi.y, dZ = q.z - pi.z;
    if ((double)(dX * dX + dY * dY + dZ * dZ) > 0.02) a |= A_GAP;
  }
  if (i >= 5 && i < n - 5) {
    const float4 pm1 = p[i - 1], pp1 = p[i + 1], pm2 = p[i - 2], pp2 = p[i + 2];
    const float4 pm3 = p[i - 3], pp3 = p[i + 3], pm4 = p[i - 4], pp4 = p[i + 4];
    // ---- FE.cpp:407-451
    const float dis = range3(pi);
    const V3d cur = {(double)pi.x, (double)pi.y, (double)pi.z};
    // The reference evaluates the two grazing-angle cosines in float64 (Eigen Vector3d). A float32 evaluation is
    // within ~1e-6 of it, so it decides every case farther than 1e-4 from the 0.966 threshold; only the (rare)
    // borderline cases - and NaNs, for which every comparison below is false - pay for the float64 form. The
    // decision is the reference's bit for bit, the FP64 pipe is no longer what bounds this kernel.
    bool graze;
    {
      const float ax = pm1.x - pi.x, ay = pm1.y - pi.y, az = pm1.z - pi.z;
      const float bx = pp1.x - pi.x, by = pp1.y - pi.y, bz = pp1.z - pi.z;
      const float nc = sqrtf(pi.x * pi.x + pi.y * pi.y + pi.z * pi.z);
      const float al = fabsf((ax * pi.x + ay * pi.y + az * pi.z) / (sqrtf(ax * ax + ay * ay + az * az) * nc));
      const float an = fabsf((bx * pi.x + by * pi.y + bz * pi.z) / (sqrtf(bx * bx + by * by + bz * bz) * nc));
      const float kT = 0.966f, kM = 1e-4f;
      if (al < kT - kM || an < kT - kM) {
        graze = false;
      } else if (al > kT + kM && an > kT + kM) {
        graze = true;
      } else {
        const V3d dl = vsub(pm1, pi), dn = vsub(pp1, pi);
        const double ncur = vnorm(cur);
        const double angle_last = vdot(dl, cur) / (vnorm(dl) * ncur);
        const double angle_next = vdot(dn, cur) / (vnorm(dn) * ncur);
        graze = fabs(angle_last) > 0.966 && fabs(angle_next) > 0.966;
      }
    }
    const int w = (dis > 50.0f || graze) ? 2 : 3;
    if (graze) a |= A_ANGLE;
    if (w == 3) a |= A_W3;
    float diffX = 0.f, diffY = 0.f, diffZ = 0.f;
    float diffR = (float)(-2 * w) * pi.w;
    diffX += pm1.x + pp1.x; diffY += pm1.y + pp1.y; diffZ += pm1.z + pp1.z; diffR += pm1.w + pp1.w;
    diffX += pm2.x + pp2.x; diffY += pm2.y + pp2.y; diffZ += pm2.z + pp2.z; diffR += pm2.w + pp2.w;
    if (w == 3) {
      diffX += pm3.x + pp3.x; diffY += pm3.y + pp3.y; diffZ += pm3.z + pp3.z; diffR += pm3.w + pp3.w;
    }
    const float tw = (float)(2 * w);
    diffX -= tw * pi.x; diffY -= tw * pi.y; diffZ -= tw * pi.z;
    curv = diffX * diffX + diffY * diffY + diffZ * diffZ;
    refl = diffR;
    const float thF = 0.02f;
    if (curv < thF * dis * thF * dis) a |= A_CAND;
    if (dis > 50.0f) a |= A_FAR;
    if ((double)curv < 0.7 * (double)thF * (double)dis * (double)thF * (double)dis && (double)refl > 20.0) a |= A_C300;
    if (pi.x * pi.x + pi.y * pi.y + pi.z * pi.z < 1.0f) a |= A_NEAR;

    // ---- FE.cpp:543-650 (tests only; which i are visited is decided in k_select)
    {
      float lX = pm4.x + pm3.x - 4 * pm2.x + pm1.x + pi.x;
      float lY = pm4.y + pm3.y - 4 * pm2.y + pm1.y + pi.y;
      float lZ = pm4.z + pm3.z - 4 * pm2.z + pm1.z + pi.z;
      float left_curv = lX * lX + lY * lY + lZ * lZ;
      float rX = pp4.x + pp3.x - 4 * pp2.x + pp1.x + pi.x;
      float rY = pp4.y + pp3.y - 4 * pp2.y + pp1.y + pi.y;
      float rZ = pp4.z + pp3.z - 4 * pp2.z + pp1.z + pi.z;
      float right_curv = rX * rX + rY * rY + rZ * rZ;
      const bool lf = left_curv < thF * dis, rf = right_curv < thF * dis;
      if (rf) a |= A_RF;
      // same two-tier evaluation for the corner test (FE.cpp:571-650): float32 first, float64 only near a threshold
      int c150 = -1;  // -1: undecided
      if (lf && rf) {
        float nl[3] = {0.f, 0.f, 0.f}, nr[3] = {0.f, 0.f, 0.f};
        const float4 L[4] = {pm1, pm2, pm3, pm4};
        const float4 R[4] = {pp1, pp2, pp3, pp4};
#pragma unroll
        for (int k = 1; k < 5; k++) {
          const float wk = (float)k / 10.0f;
          {
            const float tx = L[k - 1].x - pi.x, ty = L[k - 1].y - pi.y, tz = L[k - 1].z - pi.z;
            const float inv = wk / sqrtf(tx * tx + ty * ty + tz * tz);
            nl[0] += inv * tx; nl[1] += inv * ty; nl[2] += inv * tz;
          }
          {
            const float tx = R[k - 1].x - pi.x, ty = R[k - 1].y - pi.y, tz = R[k - 1].z - pi.z;
            const float inv = wk / sqrtf(tx * tx + ty * ty + tz * tz);
            nr[0] += inv * tx; nr[1] += inv * ty; nr[2] += inv * tz;
          }
        }
        const float nnl = sqrtf(nl[0] * nl[0] + nl[1] * nl[1] + nl[2] * nl[2]);
        const float nnr = sqrtf(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
        const float cc = fabsf((nl[0] * nr[0] + nl[1] * nr[1] + nl[2] * nr[2]) / (nnl * nnr));
        const float l4x = pm4.x - pi.x, l4y = pm4.y - pi.y, l4z = pm4.z - pi.z;
        const float r4x = pp4.x - pi.x, r4y = pp4.y - pi.y, r4z = pp4.z - pi.z;
        const float ld = sqrtf(l4x * l4x + l4y * l4y + l4z * l4z), cd = sqrtf(r4x * r4x + r4y * r4y + r4z * r4z);
        const bool conditioned = nnl > 0.05f && nnr > 0.05f;  // cancelling direction sums: the ratio is ill-conditioned
        if ((conditioned && cc > 0.5f + 1e-3f) || ld < 0.05f - 1e-5f || cd < 0.05f - 1e-5f) c150 = 0;
        else if (conditioned && cc < 0.5f - 1e-3f && ld > 0.05f + 1e-5f && cd > 0.05f + 1e-5f) c150 = 1;
      }
      if (c150 == 1) a |= A_C150;
      if (lf && rf && c150 < 0) {
        V3d nl = {0, 0, 0}, nr = {0, 0, 0};
        const float4 L[4] = {pm1, pm2, pm3, pm4};
        const float4 R[4] = {pp1, pp2, pp3, pp4};
#pragma unroll
        for (int k = 1; k < 5; k++) {
          V3d t = vsubf(L[k - 1], pi);
          vnormalize(t);
          double wk = k / 10.0;
          nl.x += wk * t.x; nl.y += wk * t.y; nl.z += wk * t.z;
        }
#pragma unroll
        for (int k = 1; k < 5; k++) {
          V3d t = vsubf(R[k - 1], pi);
          vnormalize(t);
          double wk = k / 10.0;
          nr.x += wk * t.x; nr.y += wk * t.y; nr.z += wk * t.z;
        }
        double cc = fabs(vdot(nl, nr) / (vnorm(nl) * vnorm(nr)));
        double last_dis = vnorm(vsubf(pm4, pi));
        double current_dis = vnorm(vsubf(pp4, pi));
        if (cc < 0.5 && last_dis > 0.05 && current_dis > 0.05) a |= A_C150;
      }
    }
    // ---- FE.cpp:651-806
    {
      float dX1 = pp1.x - pi.x, dY1 = pp1.y - pi.y, dZ1 = pp1.z - pi.z;
      float diff_right = sqrtf(dX1 * dX1 + dY1 * dY1 + dZ1 * dZ1);
      float dX2 = pm1.x - pi.x, dY2 = pm1.y - pi.y, dZ2 = pm1.z - pi.z;
      float diff_left = sqrtf(dX2 * dX2 + dY2 * dY2 + dZ2 * dZ2);
      float depth_right = range3(pp1), depth_left = range3(pm1);
      bool f100 = false;
      if (fabsf(diff_right - diff_left) > 1.0f) {
        if (diff_right > diff_left) {
          V3d sv = vsubf(pm1, pi);
          double cc = fabs(vdot(sv, cur) / (vnorm(sv) * vnorm(cur)));
          if (cc < 0.95) {
            if (depth_right > depth_left) f100 = true;
            else if (depth_right == 0.f) f100 = true;
          }
        } else {
          V3d sv = vsubf(pp1, pi);
          double cc = fabs(vdot(sv, cur) / (vnorm(sv) * vnorm(cur)));
          if (cc < 0.95) {
            if (depth_right < depth_left) f100 = true;
            else if (depth_left == 0.f) f100 = true;
          }
        }
      }
      if (f100) {
        V3d nf = {0, 0, 0}, nb = {0, 0, 0};
        const float4 L[3] = {pm1, pm2, pm3};
        const float4 R[3] = {pp1, pp2, pp3};
#pragma unroll
        for (int k = 1; k < 4; k++) {
          if (range3(L[k - 1]) < 1.0f) continue;
          V3d t = vsubf(L[k - 1], pi);
          vnormalize(t);
          double wk = k / 6.0;
          nf.x += wk * t.x; nf.y += wk * t.y; nf.z += wk * t.z;
        }
#pragma unroll
        for (int k = 1; k < 4; k++) {
          if (range3(L[k - 1]) < 1.0f) continue;  // sic, FE.cpp:782 tests i-k for the back side too
          V3d t = vsubf(R[k - 1], pi);
          vnormalize(t);
          double wk = k / 6.0;
          nb.x += wk * t.x; nb.y += wk * t.y; nb.z += wk * t.z;
        }
        double cc = fabs(vdot(nf, nb) / (vnorm(nf) * vnorm(nb)));
        a |= (cc < 0.95) ? A_BRK100 : A_BRK101;
      }
    }
  }
  curv_out[g] = curv;
  refl_out[g] = refl;
  attr_out[g] = (uint16_t)a;
}

// ---------------------------------------------------------------- E3: per-part stable sort
__device__ __forceinline__ void part_bounds(int n, int j, int& sp, int& ep) {
  // FE.cpp:454-455 with scanStartInd = 5, scanEndInd = n - 6
  const int span = n - 11;
  sp = 5 + (int)(((long long)span * j) / kParts);
  ep = 5 + (int)(((long long)span * (j + 1)) / kParts) - 1;
}

__global__ void __launch_bounds__(128) k_part_sort(const float* __restrict__ curv, const float* __restrict__ refl,
                                                   const int* __restrict__ line_start, const int* __restrict__ line_count,
                                                   int* __restrict__ sort_ind, int* __restrict__ refl_ind, int max_m) {
  extern __shared__ float sm[];
  const int gl = blockIdx.x / kParts, j = blockIdx.x % kParts;
  const int n = line_count[gl], ls = line_start[gl];
  if (n < 11) return;
  int sp, ep;
  part_bounds(n, j, sp, ep);
  const int m = ep - sp + 1;
  if (m <= 0) return;
  float* sc = sm;
  float* sr = sm + max_m;
  for (int e = threadIdx.x; e < m; e += 128) {
    sc[e] = curv[ls + sp + e];
    sr[e] = refl[ls + sp + e];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < m; e += 128) {
    const float vc = sc[e], vr = sr[e];
    int rc = 0, rr = 0;
    for (int k = 0; k < m; k++) {
      const float c = sc[k], r = sr[k];
      rc += (c < vc) || (c == vc && k < e);
      rr += (r < vr) || (r == vr && k < e);
    }
    sort_ind[ls + sp + rc] = sp + e;
    refl_ind[ls + sp + rr] = sp + e;
  }
}

// ---------------------------------------------------------------- E4: sequential selection
// Dynamic shared memory: flags u8[n] | v150 bits u32[(n+31)/32] | part buffers int[2*max_m]
// | (ATTR_SMEM) attr u16[n].
template <bool ATTR_SMEM>
__global__ void __launch_bounds__(64) k_select(const uint16_t* __restrict__ attr_g, const int* __restrict__ sort_ind,
                                               const int* __restrict__ refl_ind, const int* __restrict__ srt_src,
                                               const int* __restrict__ line_start, const int* __restrict__ line_count,
                                               int n_lines, int max_n, int max_m, uint8_t* __restrict__ out_label,
                                               int* __restrict__ counters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int gl = blockIdx.x;
  const int n = line_count[gl], ls = line_start[gl];
  const int nflag = (max_n + 15) & ~15;
  const int nbits = ((max_n + 31) / 32 + 3) & ~3;
  uint8_t* flags = smem_raw;
  unsigned* v150 = reinterpret_cast<unsigned*>(smem_raw + nflag);
  int* pbuf = reinterpret_cast<int*>(smem_raw + nflag + 4 * nbits);
  uint16_t* attr_s = reinterpret_cast<uint16_t*>(smem_raw + nflag + 4 * nbits + 8 * (size_t)max_m);
  const uint16_t* attr_line = attr_g + ls;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < n; i += 64) {
    flags[i] = 0;
    if (ATTR_SMEM) attr_s[i] = attr_line[i];
  }
  for (int i = tid; i < (n + 31) / 32; i += 64) v150[i] = 0;
  __syncthreads();
  auto ATTR = [&](int i) -> unsigned { return ATTR_SMEM ? (unsigned)attr_s[i] : (unsigned)__ldg(attr_line + i); };

  if (n >= 11) {
    if (warp == 0) {
      // ---- FE.cpp:453-541: parts in order; lane 0 runs the order-dependent logic
      const int w = (ATTR(n - 6) & A_W3) ? 3 : 2;  // thNumCurvSize left by the last point (FE.cpp:424-428)
      int* s_sort = pbuf;
      int* s_refl = pbuf + max_m;
      for (int j = 0; j < kParts; j++) {
        int sp, ep;
        part_bounds(n, j, sp, ep);
        const int m = ep - sp + 1;
        if (m <= 0) continue;
        __syncwarp();
        for (int e = lane; e < m; e += 32) {
          s_sort[e] = sort_ind[ls + sp + e];
          s_refl[e] = refl_ind[ls + sp + e];
        }
        __syncwarp();
        if (lane == 0) {
          for (int k = 0; k < m; k++) {  // FE.cpp:483-519
            const int ind = s_sort[k];
            if (flags[ind] != 0) continue;
            const unsigned a = ATTR(ind);
            if (a & A_CAND) {
              flags[ind] = 3;
              if (!(a & A_FAR)) {
                for (int l = 1; l <= w; l++) {
                  if (ATTR(ind + l - 1) & A_GAP) break;
                  flags[ind + l] = 1;
                }
                for (int l = 1; l <= w; l++) {
                  if (ATTR(ind - l) & A_GAP) break;
                  flags[ind - l] = 1;
                }
              }
            }
          }
          int smallest = 1, sharpest = 1;
          for (int k = 0; k < m; k++) {  // FE.cpp:521-539
            const int ind = s_sort[k];
            const unsigned a = ATTR(ind);
            const int f = flags[ind];
            if ((f == 3 && smallest <= 1) || (f == 3 && (a & A_FAR)) || (a & A_ANGLE)) {
              smallest++;
              flags[ind] = 2;
            }
            const int idx = s_refl[k];
            if (sharpest <= 3 && (ATTR(idx) & A_C300)) {
              sharpest++;
              flags[idx] = 4;  // 300
            }
          }
        }
      }
    } else if (lane == 0) {
      // ---- FE.cpp:543-650: visit i = 5, then i += 4 if the right side was flat else 1
      int i = 5;
      while (i < n - 5) {
        const unsigned a = ATTR(i);
        if (a & A_C150) v150[i >> 5] |= 1u << (i & 31);
        i += (a & A_RF) ? 4 : 1;
      }
    }
  }
  __syncthreads();

  // ---- FE.cpp:818-842 + label write-back FE.cpp:1016-1023
  const int scan = gl / n_lines;
  int n_sharp = 0, n_flat = 0;
  for (int i0 = 0; i0 < n; i0 += 64) {
    const int i = i0 + tid;
    int label = 0;
    if (i < n && i >= 5 && i < n - 5) {
      const unsigned a = ATTR(i);
      if (!(a & A_NEAR)) {
        const bool v = (v150[i >> 5] >> (i & 31)) & 1u;
        if (a & A_BRK100) label = 1;
        else if (a & A_BRK101) label = 0;
        else if (v) label = 1;
        else if (flags[i] == 2) label = 2;
      }
    }
    if (i < n) out_label[srt_src[ls + i]] = (uint8_t)label;
    n_sharp += __popc(__ballot_sync(0xffffffffu, label == 1));
    n_flat += __popc(__ballot_sync(0xffffffffu, label == 2));
  }
  if (lane == 0) {
    if (n_sharp) atomicAdd(&counters[2 * scan], n_sharp);
    if (n_flat) atomicAdd(&counters[2 * scan + 1], n_flat);
  }
}

// ---------------------------------------------------------------- E4': part-parallel selection
// The flat selection of FE.cpp:453-541 is sequential across the 50 parts of a line only through
// neighbour suppression spilling over a part boundary: processing part j marks up to 3 points at
// the head of part j+1 (before j+1 runs) and up to 3 at the tail of part j-1 (after j-1 finished).
// Both spills are contiguous from the boundary, so part j's behaviour depends on its predecessor
// only through h_j in {0,1,2,3} = how many of its head points arrive pre-marked. Every part is
// therefore simulated for all four values of h_j in parallel (4 x 50 lanes), a 50-step chain
// picks the h_j that actually occurs, and the tail spills are applied last — the result is
// identical to the sequential order. The count_num stride walk (FE.cpp:543-650) is a 4-state
// automaton (distance to the next visited index), evaluated with a parallel scan of composed
// transition functions.
// Dynamic shared memory: flags u8[4][NF] | attr8 u8[NF] | v150 u32[nbits]
constexpr int kSelThreads = 256;
constexpr int kMinParallelN = 11 + 4 * kParts;  // every part holds >= 4 points: spills reach adjacent parts only

__device__ __forceinline__ unsigned compose4(unsigned f, unsigned g) {  // (g o f)[s] = g[f[s]], 2 bits per state
  unsigned r = 0;
#pragma unroll
  for (int s = 0; s < 4; s++) r |= ((g >> (2 * ((f >> (2 * s)) & 3u))) & 3u) << (2 * s);
  return r;
}

__global__ void __launch_bounds__(kSelThreads) k_select_par(const uint16_t* __restrict__ attr_g, const int* __restrict__ sort_ind,
                                                            const int* __restrict__ refl_ind, const int* __restrict__ srt_src,
                                                            const int* __restrict__ line_start, const int* __restrict__ line_count,
                                                            int n_lines, int max_n, uint8_t* __restrict__ out_label,
                                                            int* __restrict__ counters, int* __restrict__ overflow_flag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned char s_out_h[kParts][4], s_lsp[kParts][4], s_hsel[kParts + 1];
  __shared__ unsigned s_fn[kSelThreads];
  __shared__ unsigned s_warp_fn[kSelThreads / 32];
  const int gl = blockIdx.x;
  const int n = line_count[gl], ls = line_start[gl];
  const int NF = (max_n + 15) & ~15;
  if (n > NF) {  // does not fit: the host re-runs the sequential kernel
    if (threadIdx.x == 0) atomicExch(overflow_flag, 1);
    return;
  }
  uint8_t* flags4 = smem_raw;
  uint8_t* attr8 = smem_raw + 4 * (size_t)NF;
  unsigned* v150 = reinterpret_cast<unsigned*>(smem_raw + 5 * (size_t)NF);
  const uint16_t* attr_line = attr_g + ls;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < n; i += kSelThreads) {
    attr8[i] = (uint8_t)(attr_line[i] & 0xFFu);
    flags4[i] = 0; flags4[NF + i] = 0; flags4[2 * NF + i] = 0; flags4[3 * NF + i] = 0;
  }
  for (int i = tid; i < (n + 31) / 32; i += kSelThreads) v150[i] = 0;
  if (tid < kParts * 4) { (&s_out_h[0][0])[tid] = 0; (&s_lsp[0][0])[tid] = 0; }
  if (tid <= kParts) s_hsel[tid] = 0;
  __syncthreads();

  if (n >= 11) {
    const int w = (attr_line[n - 6] & A_W3) ? 3 : 2;  // thNumCurvSize left by the last point (FE.cpp:424-428)
    const bool par = n >= kMinParallelN;
    // ---- phase C: stride walk (FE.cpp:543-650) as a scan of transition functions over 256 contiguous chunks
    {
      const int span = n - 10;  // indices 5 .. n-6
      const int chunk = (span + kSelThreads - 1) / kSelThreads;
      const int a0 = 5 + tid * chunk, a1 = min(a0 + chunk, n - 5);
      unsigned fn = 0xE4u;  // identity: state s -> s
      {
        int st0 = 0, st1 = 1, st2 = 2, st3 = 3;
        for (int i = a0; i < a1; i++) {
          const int jump = (attr8[i] & A_RF) ? 3 : 0;
          st0 = st0 ? st0 - 1 : jump;
          st1 = st1 ? st1 - 1 : jump;
          st2 = st2 ? st2 - 1 : jump;
          st3 = st3 ? st3 - 1 : jump;
        }
        if (a0 < a1) fn = (unsigned)st0 | ((unsigned)st1 << 2) | ((unsigned)st2 << 4) | ((unsigned)st3 << 6);
      }
      // inclusive scan of composition within the warp, then across the 8 warps
      unsigned inc = fn;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned prev = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc = compose4(prev, inc);
      }
      if (lane == 31) s_warp_fn[warp] = inc;
      s_fn[tid] = inc;
      __syncthreads();
      unsigned pre = 0xE4u;  // composition of all chunks before this thread's
      for (int wv = 0; wv < warp; wv++) pre = compose4(pre, s_warp_fn[wv]);
      if (lane > 0) pre = compose4(pre, s_fn[tid - 1]);
      int sst = (int)(pre & 3u);  // state entering index 5 is 0 (index 5 is visited)
      for (int i = a0; i < a1; i++) {
        if (sst == 0) {
          const unsigned a = attr8[i];
          if (a & A_C150) atomicOr(&v150[i >> 5], 1u << (i & 31));
          sst = (a & A_RF) ? 3 : 0;
        } else {
          sst--;
        }
      }
    }
    __syncthreads();
    // ---- repack the attribute byte for the selection loops: bits 0-2 CAND/FAR/ANGLE, bit 3 C300, bits 4-5 /
    //      6-7 = how many right / left neighbours a pick at this index suppresses (the walks of FE.cpp:492-517
    //      with thNumCurvSize, the 0.02 gap test and the far test folded in). Copy 3 of the flags is the scratch.
    {
      uint8_t* tmp = flags4 + 3 * (size_t)NF;
      for (int i = tid; i < n; i += kSelThreads) {
        const unsigned a = attr8[i];
        int rr = 0, ll = 0;
        if (i >= 5 && i < n - 5 && !(a & A_FAR)) {
          for (int l = 1; l <= w; l++) {
            if (attr8[i + l - 1] & A_GAP) break;
            rr = l;
          }
          for (int l = 1; l <= w; l++) {
            if (attr8[i - l] & A_GAP) break;
            ll = l;
          }
        }
        tmp[i] = (uint8_t)((a & 7u) | (((a >> 4) & 1u) << 3) | ((unsigned)rr << 4) | ((unsigned)ll << 6));
      }
      __syncthreads();
      for (int i = tid; i < n; i += kSelThreads) {
        attr8[i] = tmp[i];
        tmp[i] = 0;
      }
      __syncthreads();
    }
    constexpr unsigned B_CAND = 1u, B_FAR = 2u, B_ANGLE = 4u, B_C300 = 8u;
    // ---- phase 1: flat selection. parallel: lane (j,h) simulates part j with h pre-marked head points;
    //      short lines: lane 0 runs all parts in order on copy 0 (spills written directly)
    const int n_sim = par ? kParts * 4 : 1;
    if (tid < n_sim) {
      const int j0 = par ? (tid >> 2) : 0, j1 = par ? j0 + 1 : kParts;
      const int h = par ? (tid & 3) : 0;
      uint8_t* F = flags4 + (size_t)h * NF;
      for (int j = j0; j < j1; j++) {
        int sp, ep;
        part_bounds(n, j, sp, ep);
        const int m = ep - sp + 1;
        if (m <= 0) continue;
        const int lo = par ? sp : 0, hi = par ? ep : n - 1;  // indices this lane may write
        if (par)
          for (int t = 0; t < h; t++) F[sp + t] = 1;
        int rs = 0, lsp = 0;
        const int* so = sort_ind + ls + sp;
        const int* ro = refl_ind + ls + sp;
        int ind_next = __ldg(so);
        for (int k = 0; k < m; k++) {  // FE.cpp:483-519
          const int ind = ind_next;
          if (k + 1 < m) ind_next = __ldg(so + k + 1);
          const unsigned a = attr8[ind];
          if (F[ind] == 0 && (a & B_CAND)) {
            F[ind] = 3;
            const int rr = (a >> 4) & 3, ll = (a >> 6) & 3;
#pragma unroll
            for (int l = 1; l <= 3; l++) {
              if (l <= rr) {
                const int idx = ind + l;
                if (idx <= hi) F[idx] = 1;
                else rs = max(rs, idx - hi);
              }
              if (l <= ll) {
                const int idx = ind - l;
                if (idx >= lo) F[idx] = 1;
                else lsp = max(lsp, lo - idx);
              }
            }
          }
        }
        int smallest = 1, sharpest = 1;
        for (int k = 0; k < m; k++) {  // FE.cpp:521-539
          const int ind = __ldg(so + k);
          const int idx = __ldg(ro + k);
          const unsigned a = attr8[ind];
          const int f = F[ind];
          if ((f == 3 && (smallest <= 1 || (a & B_FAR))) || (a & B_ANGLE)) {
            smallest++;
            F[ind] = 2;
          }
          if (sharpest <= 3 && (attr8[idx] & B_C300)) {
            sharpest++;
            F[idx] = 4;  // 300
          }
        }
        if (par) {
          s_out_h[j][h] = (unsigned char)min(rs, 3);
          s_lsp[j][h] = (unsigned char)min(lsp, 3);
        }
      }
    }
    __syncthreads();
    // ---- phase 2: which h_j actually occurs
    if (tid == 0 && par) {
      int h = 0;
      for (int j = 0; j < kParts; j++) {
        s_hsel[j] = (unsigned char)h;
        h = s_out_h[j][h];
      }
      s_hsel[kParts] = 0;
    }
  }
  __syncthreads();

  // ---- phase 3: FE.cpp:818-842 + label write-back FE.cpp:1016-1023
  const int scan = gl / n_lines;
  const bool par = n >= kMinParallelN;
  int n_sharp = 0, n_flat = 0;
  for (int i0 = 0; i0 < n; i0 += kSelThreads) {
    const int i = i0 + tid;
    int label = 0;
    if (i < n && i >= 5 && i < n - 5) {
      const unsigned a = attr_line[i];
      if (!(a & A_NEAR)) {
        const bool v = (v150[i >> 5] >> (i & 31)) & 1u;
        if (a & A_BRK100) label = 1;
        else if (a & A_BRK101) label = 0;
        else if (v) label = 1;
        else {
          int f;
          if (par) {
            int j = (int)(((long long)(i - 5) * kParts) / (n - 11));
            j = j > kParts - 1 ? kParts - 1 : j;
            int sp, ep;
            part_bounds(n, j, sp, ep);
            while (i < sp) { j--; part_bounds(n, j, sp, ep); }
            while (i > ep) { j++; part_bounds(n, j, sp, ep); }
            f = flags4[(size_t)s_hsel[j] * NF + i];
            if (j + 1 < kParts && i > ep - (int)s_lsp[j + 1][s_hsel[j + 1]]) f = 1;  // tail spill of part j+1
          } else {
            f = flags4[i];
          }
          if (f == 2) label = 2;
        }
      }
    }
    if (i < n) out_label[srt_src[ls + i]] = (uint8_t)label;
    n_sharp += __popc(__ballot_sync(0xffffffffu, label == 1));
    n_flat += __popc(__ballot_sync(0xffffffffu, label == 2));
  }
  if (lane == 0) {
    if (n_sharp) atomicAdd(&counters[2 * scan], n_sharp);
    if (n_flat) atomicAdd(&counters[2 * scan + 1], n_flat);
  }
}

}  // namespace mml

using namespace mml;

// Device-resident extraction. xyzi_d/line_d hold n_total points; labels go to label_d (u8),
// per-scan (n_sharp, n_flat) to ctx->counters[2*s..]. scan_off is a HOST array.
int mml_extract_device(mml_ctx* ctx, const float4* xyzi_d, const uint16_t* line_d, const int* scan_off, int n_scans,
                       int n_lines, uint8_t* label_d, bool force_sequential) {
  if (n_lines <= 0 || n_lines > kMaxLines) return mml_fail(ctx, MML_ERR_INVALID, "n_lines must be in [1,64]");
  const int n_total = scan_off[n_scans];
  const int TL = n_scans * n_lines;
  cudaStream_t st = ctx->stream;
  MML_CUDA(ctx, ctx->counters.reserve(sizeof(int) * (2 * (size_t)n_scans + 64)));
  // counters: per scan (n_sharp, n_flat), then the overflow flag. A pipelined caller redirects them to its slot.
  int* const counters = ctx->counters_alt ? ctx->counters_alt : ctx->counters.as<int>();
  MML_CUDA(ctx, cudaMemsetAsync(counters, 0, sizeof(int) * (2 * (size_t)n_scans + 1), st));
  if (n_total <= 0) return MML_OK;

  // host-side chunk table (metadata only)
  std::vector<Chunk> chunks;
  std::vector<int> scan_chunk0(n_scans + 1);
  int max_scan = 0;
  for (int s = 0; s < n_scans; s++) {
    scan_chunk0[s] = (int)chunks.size();
    const int a = scan_off[s], b = scan_off[s + 1];
    if (b < a) return mml_fail(ctx, MML_ERR_INVALID, "scan_offsets must be non-decreasing");
    max_scan = b - a > max_scan ? b - a : max_scan;
    for (int p = a; p < b; p += kChunkPts) chunks.push_back({s, p, (b - p < kChunkPts ? b - p : kChunkPts), 0});
  }
  scan_chunk0[n_scans] = (int)chunks.size();
  const int n_chunks = (int)chunks.size();

  const size_t meta_bytes = sizeof(Chunk) * n_chunks + sizeof(int) * (2 * (size_t)n_scans + 2);
  // the chunk table only depends on the scan offsets: re-upload only when they change
  const bool cached = ctx->chunk_tab.p && (int)ctx->last_scan_off.size() == n_scans + 1 &&
                      memcmp(ctx->last_scan_off.data(), scan_off, sizeof(int) * (n_scans + 1)) == 0;
  if (!cached) {
    MML_CUDA(ctx, ctx->pin_small.reserve(meta_bytes));
    MML_CUDA(ctx, ctx->chunk_tab.reserve(meta_bytes));
    // the pinned staging area may still be in flight from a previous call
    MML_CUDA(ctx, cudaStreamSynchronize(st));
    char* hp = ctx->pin_small.as<char>();
    memcpy(hp, chunks.data(), sizeof(Chunk) * n_chunks);
    memcpy(hp + sizeof(Chunk) * n_chunks, scan_chunk0.data(), sizeof(int) * (n_scans + 1));
    memcpy(hp + sizeof(Chunk) * n_chunks + sizeof(int) * (n_scans + 1), scan_off, sizeof(int) * (n_scans + 1));
    MML_CUDA(ctx, cudaMemcpyAsync(ctx->chunk_tab.p, hp, meta_bytes, cudaMemcpyHostToDevice, st));
    ctx->last_scan_off.assign(scan_off, scan_off + n_scans + 1);
  }
  const Chunk* chunks_d = ctx->chunk_tab.as<Chunk>();
  const int* scan_chunk0_d = reinterpret_cast<const int*>(ctx->chunk_tab.as<char>() + sizeof(Chunk) * n_chunks);
  const int* scan_off_d = scan_chunk0_d + (n_scans + 1);

  MML_CUDA(ctx, ctx->chunk_hist.reserve(sizeof(int) * (size_t)n_chunks * n_lines));
  MML_CUDA(ctx, ctx->line_start.reserve(sizeof(int) * ((size_t)TL + 1)));
  MML_CUDA(ctx, ctx->line_count.reserve(sizeof(int) * ((size_t)TL + 1)));
  MML_CUDA(ctx, ctx->srt_xyzi.reserve(sizeof(float4) * (size_t)n_total));
  MML_CUDA(ctx, ctx->srt_src.reserve(sizeof(int) * (size_t)n_total));
  MML_CUDA(ctx, ctx->srt_line.reserve(sizeof(int) * (size_t)n_total));
  MML_CUDA(ctx, ctx->curv.reserve(sizeof(float) * (size_t)n_total));
  MML_CUDA(ctx, ctx->refl.reserve(sizeof(float) * (size_t)n_total));
  MML_CUDA(ctx, ctx->attr.reserve(sizeof(uint16_t) * (size_t)n_total));
  MML_CUDA(ctx, ctx->sort_ind.reserve(sizeof(int) * (size_t)n_total));
  MML_CUDA(ctx, ctx->refl_ind.reserve(sizeof(int) * (size_t)n_total));

  int* hist = ctx->chunk_hist.as<int>();
  int* line_start = ctx->line_start.as<int>();
  int* line_count = ctx->line_count.as<int>();

  MML_CUDA(ctx, cudaMemsetAsync(ctx->srt_line.p, 0xFF, sizeof(int) * (size_t)n_total, st));
  MML_CUDA(ctx, cudaMemsetAsync(label_d, 0, (size_t)n_total, st));
  k_line_hist<<<n_chunks, 256, 0, st>>>(line_d, chunks_d, n_lines, hist);
  MML_LAUNCHED(ctx);
  k_line_scan<<<n_scans, 256, 0, st>>>(hist, scan_chunk0_d, scan_off_d, n_lines, line_start, line_count);
  MML_LAUNCHED(ctx);
  k_line_scatter<<<n_chunks, 256, 0, st>>>(xyzi_d, line_d, chunks_d, hist, line_start, n_lines, ctx->srt_xyzi.as<float4>(),
                                           ctx->srt_src.as<int>(), ctx->srt_line.as<int>());
  MML_LAUNCHED(ctx);
  k_point_attr<<<div_up(n_total, 256), 256, 0, st>>>(ctx->srt_xyzi.as<float4>(), ctx->srt_line.as<int>(), line_start,
                                                     line_count, n_total, ctx->curv.as<float>(), ctx->refl.as<float>(),
                                                     ctx->attr.as<uint16_t>());
  MML_LAUNCHED(ctx);

  // a line cannot be longer than its scan; a part holds at most ceil((n-11)/50)+1 points
  const int max_n = max_scan;
  const int max_m = max_n > 11 ? (max_n - 11) / kParts + 2 : 2;
  const size_t sort_smem = sizeof(float) * 2 * (size_t)max_m;
  if (sort_smem > 200 * 1024) return mml_fail(ctx, MML_ERR_CAPACITY, "scan line too long for k_part_sort shared memory");
  if (sort_smem > 48 * 1024)
    MML_CUDA(ctx, cudaFuncSetAttribute(k_part_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem));
  k_part_sort<<<TL * kParts, 128, sort_smem, st>>>(ctx->curv.as<float>(), ctx->refl.as<float>(), line_start, line_count,
                                                   ctx->sort_ind.as<int>(), ctx->refl_ind.as<int>(), max_m);
  MML_LAUNCHED(ctx);

  int* overflow_flag = counters + 2 * (size_t)n_scans;
  const size_t kMaxSmem = 227 * 1024;
  if (!force_sequential && ctx->sel_tier < 2) {
    // part-parallel selection. Shared memory is sized by a sticky per-context tier: tier 0 holds lines of up to
    // 8192 points (41 KB: 5 CTAs per SM, what VLP-16 / Horizon frames need), tier 1 up to 44032 points (one CTA
    // per SM, 240k-point Horizon clouds). A longer line raises the overflow flag; the caller bumps the tier and
    // re-runs (tier 2 = the sequential kernel).
    const int tier_cap = ctx->sel_tier == 0 ? 8192 : 44032;
    const int cap_n = max_n < tier_cap ? ((max_n + 15) & ~15) : tier_cap;
    const size_t par_smem = 5 * (size_t)cap_n + 4 * ((((size_t)cap_n + 31) / 32 + 3) & ~(size_t)3);
    if (par_smem > 48 * 1024)
      MML_CUDA(ctx, cudaFuncSetAttribute(k_select_par, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)par_smem));
    k_select_par<<<TL, kSelThreads, par_smem, st>>>(ctx->attr.as<uint16_t>(), ctx->sort_ind.as<int>(), ctx->refl_ind.as<int>(),
                                                    ctx->srt_src.as<int>(), line_start, line_count, n_lines, cap_n, label_d,
                                                    counters, overflow_flag);
  } else {
    // sequential fallback: size shared memory for the longest line actually present
    std::vector<int> lc(TL);
    MML_CUDA(ctx, cudaMemcpyAsync(lc.data(), line_count, sizeof(int) * TL, cudaMemcpyDeviceToHost, st));
    MML_CUDA(ctx, cudaStreamSynchronize(st));
    int longest = 0;
    for (int v : lc) longest = v > longest ? v : longest;
    const size_t nflag = ((size_t)longest + 15) & ~(size_t)15;
    const size_t nbits = (((size_t)longest + 31) / 32 + 3) & ~(size_t)3;
    const size_t base_smem = nflag + 4 * nbits + 8 * (size_t)max_m;
    if (base_smem > kMaxSmem) return mml_fail(ctx, MML_ERR_CAPACITY, "scan line too long for k_select shared memory");
    if (base_smem > 48 * 1024)
      MML_CUDA(ctx, cudaFuncSetAttribute(k_select<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)base_smem));
    k_select<false><<<TL, 64, base_smem, st>>>(ctx->attr.as<uint16_t>(), ctx->sort_ind.as<int>(), ctx->refl_ind.as<int>(),
                                               ctx->srt_src.as<int>(), line_start, line_count, n_lines, longest, max_m, label_d,
                                               counters);
  }
  MML_LAUNCHED(ctx);
  MML_CUDA(ctx, cudaGetLastError());
  return MML_OK;
}
