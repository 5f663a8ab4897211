// LiDAR factor evaluation shared by the Estimate kernels (accumulate.cu: one frame; windowsolve.cu: sliding window):
// pose linearisation, the two cost functors with analytic Jacobians and Huber, the 28-sum warp reduction and the
// thread-block-cluster primitives.
//   Cost_NavState_IMU_Line::operator()      include/utils/ceresfunc.h:412-440
//   Cost_NavState_IMU_Plan_Vec::operator()  include/utils/ceresfunc.h:533-555
//   HuberLoss + Corrector (rho'' <= 0)      mirrored by ResidualBlockInfo::Evaluate, CF.h:33-63
#pragma once
#include "smallmath.cuh"

namespace mml {

struct PoseLin {
  double R[9], t[3], Jr[9], Rbl[9], Pbl[3];
};

__device__ inline void right_jacobian(const double* phi, double* Jr) {
  const double th2 = (phi[0] * phi[0] + phi[1] * phi[1]) + phi[2] * phi[2];
  double a, b;
  if (th2 < 1e-12) {
    a = 0.5 - th2 / 24.0;
    b = 1.0 / 6.0 - th2 / 120.0;
  } else {
    const double th = sqrt(th2);
    a = (1.0 - cos(th)) / th2;
    b = (th - sin(th)) / (th2 * th);
  }
  const double K[9] = {0, -phi[2], phi[1], phi[2], 0, -phi[0], -phi[1], phi[0], 0};
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += K[3 * r + k] * K[3 * k + c];
      Jr[3 * r + c] = -a * K[3 * r + c] + b * s + (r == c ? 1.0 : 0.0);
    }
}

__device__ inline void make_pose(const double* x6, const double* Rbl, const double* Pbl, PoseLin& L) {
  const Quat q = so3_exp(x6 + 3);
  quat_to_R(q, L.R);
  L.t[0] = x6[0]; L.t[1] = x6[1]; L.t[2] = x6[2];
  right_jacobian(x6 + 3, L.Jr);
  for (int i = 0; i < 9; i++) L.Rbl[i] = Rbl[i];
  for (int i = 0; i < 3; i++) L.Pbl[i] = Pbl[i];
}

// accumulate one scalar residual r with dr/dP = gr (3) at body-frame point u
__device__ __forceinline__ void add_row(const PoseLin& L, const double* u, const double* gr, double r, double* acc) {
  // J = [gr^T | (Jr^T (u x R^T gr))^T]
  const double v0 = L.R[0] * gr[0] + L.R[3] * gr[1] + L.R[6] * gr[2];
  const double v1 = L.R[1] * gr[0] + L.R[4] * gr[1] + L.R[7] * gr[2];
  const double v2 = L.R[2] * gr[0] + L.R[5] * gr[1] + L.R[8] * gr[2];
  const double w0 = u[1] * v2 - u[2] * v1, w1 = u[2] * v0 - u[0] * v2, w2 = u[0] * v1 - u[1] * v0;
  double J[6];
  J[0] = gr[0]; J[1] = gr[1]; J[2] = gr[2];
  J[3] = L.Jr[0] * w0 + L.Jr[3] * w1 + L.Jr[6] * w2;
  J[4] = L.Jr[1] * w0 + L.Jr[4] * w1 + L.Jr[7] * w2;
  J[5] = L.Jr[2] * w0 + L.Jr[5] * w1 + L.Jr[8] * w2;
  int k = 7;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    acc[1 + i] += J[i] * r;
#pragma unroll
    for (int j = i; j < 6; j++) acc[k++] += J[i] * J[j];
  }
}

// compact 48 B records (csrc/associate.cu k_export / fit writers); false = empty slot
__device__ __forceinline__ bool load_line(const float4* __restrict__ f, int i, double* p, double* a, double* b) {
  const float4 f0 = __ldg(f + 3 * (size_t)i);
  if (!(f0.w == 1.f)) return false;
  const float4 f1 = __ldg(f + 3 * (size_t)i + 1), f2 = __ldg(f + 3 * (size_t)i + 2);
  p[0] = f0.x; p[1] = f0.y; p[2] = f0.z;
  a[0] = f1.x; a[1] = f1.y; a[2] = f1.z;
  b[0] = f1.w; b[1] = f2.x; b[2] = f2.y;
  return true;
}
__device__ __forceinline__ bool load_plane(const float4* __restrict__ f, int i, double* p, double* pp, double* n) {
  const float4 f0 = __ldg(f + 3 * (size_t)i);
  if (!(f0.w == 1.f)) return false;
  const float4 f1 = __ldg(f + 3 * (size_t)i + 1), f2 = __ldg(f + 3 * (size_t)i + 2);
  p[0] = f0.x; p[1] = f0.y; p[2] = f0.z;
  n[0] = f2.x; n[1] = f2.y; n[2] = f2.z;
  const double dist = (double)f1.w;
  pp[0] = (double)f1.x - dist * n[0]; pp[1] = (double)f1.y - dist * n[1]; pp[2] = (double)f1.z - dist * n[2];
  return true;
}

// 1/sqrt in float64: one MUFU seed + Newton steps on the device instead of a square root and a division
__device__ __forceinline__ double rsqrt64(double x) { return rsqrt(x); }

// Every thread's evaluation is one dependent float64 chain, so square roots followed by divisions are folded
// into reciprocal square roots and products (results move by an ulp or two against a literal transcription of
// the functors; the gate for this file is the 1e-4 m / 1e-4 rad pose tolerance, measured at ~1e-15).
//
// point-to-line residual + Jacobian + Huber, CF.h:412-440, accumulated into acc[28]
__device__ __forceinline__ void eval_line(const PoseLin& L, const double* p, const double* a, const double* b, double s_info,
                                          double ha, double* acc) {
  double u[3], P[3];
  for (int r = 0; r < 3; r++) u[r] = L.Rbl[3 * r] * p[0] + L.Rbl[3 * r + 1] * p[1] + L.Rbl[3 * r + 2] * p[2] + L.Pbl[r];
  for (int r = 0; r < 3; r++) P[r] = L.R[3 * r] * u[0] + L.R[3 * r + 1] * u[1] + L.R[3 * r + 2] * u[2] + L.t[r];
  const double ab[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
  const double inv_l12 = rsqrt64(ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2]);
  const double c0 = (P[0] - a[0]) * (P[1] - b[1]) - (P[0] - b[0]) * (P[1] - a[1]);
  const double c1 = (P[0] - a[0]) * (P[2] - b[2]) - (P[0] - b[0]) * (P[2] - a[2]);
  const double c2 = (P[1] - a[1]) * (P[2] - b[2]) - (P[1] - b[1]) * (P[2] - a[2]);
  const double cc = c0 * c0 + c1 * c1 + c2 * c2;
  const double inv_a012 = rsqrt64(cc);
  const double a012 = cc * inv_a012;
  const double ld2 = a012 * inv_l12;  // distance to the line, >= 0
  const double PP = P[0] * P[0] + P[1] * P[1] + P[2] * P[2];
  const double inv_sq = rsqrt64(sqrt(PP));  // 1 / |P|^(1/2)
  const double inv_PP = (inv_sq * inv_sq) * (inv_sq * inv_sq);
  const double w = 1.0 - 0.9 * ld2 * inv_sq;
  double r = s_info * w * ld2;
  const double ch[3] = {c2 * inv_a012, -c1 * inv_a012, c0 * inv_a012};
  const double gd[3] = {(ab[1] * ch[2] - ab[2] * ch[1]) * inv_l12, (ab[2] * ch[0] - ab[0] * ch[2]) * inv_l12,
                        (ab[0] * ch[1] - ab[1] * ch[0]) * inv_l12};
  const double k5 = 0.5 * ld2 * inv_sq * inv_PP;
  double gr[3];
  for (int k = 0; k < 3; k++) gr[k] = s_info * (w * gd[k] + ld2 * (-0.9 * (gd[k] * inv_sq - k5 * P[k])));
  // Huber, CF.h:33-63 with rho'' <= 0
  const double s = r * r;
  double k1 = 1.0, rho = s;
  if (ha > 0 && s > ha * ha) {
    const double rr = sqrt(s);
    k1 = sqrt(ha / rr);
    rho = 2 * ha * rr - ha * ha;
  }
  acc[0] += 0.5 * rho;
  r *= k1;
  gr[0] *= k1; gr[1] *= k1; gr[2] *= k1;
  add_row(L, u, gr, r, acc);
}

// point-to-plane (vector form) residual + Jacobian + Huber, CF.h:533-555
__device__ __forceinline__ void eval_plane(const PoseLin& L, const double* p, const double* pp, const double* n_f32, double s_info,
                                           double w_tan, double ha, double* acc) {
  // The factor's direction is the UNIT vector along the float32 normal: sqrt_info = info * (V U^T)^T from
  // JacobiSVD(e1 n^T) (EST.cpp:675-682) keeps the singular vector and drops the singular value |n| = 1 +- 6e-8.
  // (p_proj keeps the float32 normal as it is, EST.cpp:672-673.)
  const double inv_nn = rsqrt64(n_f32[0] * n_f32[0] + n_f32[1] * n_f32[1] + n_f32[2] * n_f32[2]);
  const double n[3] = {n_f32[0] * inv_nn, n_f32[1] * inv_nn, n_f32[2] * inv_nn};
  double u[3], P[3];
  for (int r = 0; r < 3; r++) u[r] = L.Rbl[3 * r] * p[0] + L.Rbl[3 * r + 1] * p[1] + L.Rbl[3 * r + 2] * p[2] + L.Pbl[r];
  for (int r = 0; r < 3; r++) P[r] = L.R[3 * r] * u[0] + L.R[3 * r + 1] * u[1] + L.R[3 * r + 2] * u[2] + L.t[r];
  const double e[3] = {P[0] - pp[0], P[1] - pp[1], P[2] - pp[2]};
  const double ee = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
  const double inv_en = rsqrt64(ee);
  const double en = ee * inv_en;
  const double PP = P[0] * P[0] + P[1] * P[1] + P[2] * P[2];
  const double inv_sq = rsqrt64(sqrt(PP));
  const double inv_PP = (inv_sq * inv_sq) * (inv_sq * inv_sq);
  const double w = 1.0 - 0.9 * en * inv_sq;
  const double k5 = 0.5 * en * inv_sq * inv_PP;
  double gw[3];
  for (int k = 0; k < 3; k++) gw[k] = -0.9 * ((e[k] * inv_en) * inv_sq - k5 * P[k]);
  // residual in the canonical basis [n t1 t2]: sqrt_info^T sqrt_info = (n n^T + w_t^2 (I - n n^T)) / lidar_m^2.
  // ||r||^2 needs no basis: (s w)^2 [ (n.e)^2 + w_t^2 (|e|^2 - (n.e)^2) ]
  const double ne = n[0] * e[0] + n[1] * e[1] + n[2] * e[2];
  const double sw = s_info * w;
  double s = sw * sw * ne * ne;
  if (w_tan != 0.0) s += sw * sw * w_tan * w_tan * fmax(ee - ne * ne, 0.0);
  double k1 = 1.0, rho = s;
  if (ha > 0 && s > ha * ha) {
    const double rr = sqrt(s);
    k1 = sqrt(ha / rr);
    rho = 2 * ha * rr - ha * ha;
  }
  acc[0] += 0.5 * rho;
  {
    const double sc = k1 * s_info;
    const double gr[3] = {sc * (w * n[0] + ne * gw[0]), sc * (w * n[1] + ne * gw[1]), sc * (w * n[2] + ne * gw[2])};
    add_row(L, u, gr, sc * w * ne, acc);
  }
  if (w_tan != 0.0) {  // window size 5 only (EST.cpp:1203): the two tangential rows
    int kk = 0;
    if (fabs(n[1]) < fabs(n[kk])) kk = 1;
    if (fabs(n[2]) < fabs(n[kk])) kk = 2;
    double ex[3] = {0, 0, 0};
    ex[kk] = 1.0;
    const double v[3] = {ex[1] * n[2] - ex[2] * n[1], ex[2] * n[0] - ex[0] * n[2], ex[0] * n[1] - ex[1] * n[0]};
    const double inv_nv = rsqrt64(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const double t1[3] = {v[0] * inv_nv, v[1] * inv_nv, v[2] * inv_nv};
    const double t2[3] = {n[1] * t1[2] - n[2] * t1[1], n[2] * t1[0] - n[0] * t1[2], n[0] * t1[1] - n[1] * t1[0]};
    const double sc = k1 * s_info * w_tan;
    const double b1 = t1[0] * e[0] + t1[1] * e[1] + t1[2] * e[2];
    const double b2 = t2[0] * e[0] + t2[1] * e[1] + t2[2] * e[2];
    const double g1[3] = {sc * (w * t1[0] + b1 * gw[0]), sc * (w * t1[1] + b1 * gw[1]), sc * (w * t1[2] + b1 * gw[2])};
    const double g2[3] = {sc * (w * t2[0] + b2 * gw[0]), sc * (w * t2[1] + b2 * gw[1]), sc * (w * t2[2] + b2 * gw[2])};
    add_row(L, u, g1, sc * w * b1, acc);
    add_row(L, u, g2, sc * w * b2, acc);
  }
}

// Sum acc[0..27] over the 32 lanes of a warp: each exchange step halves the number of values a lane carries
// (16 + 8 + 4 + 2 + 1 = 31 shuffles of a double instead of 28 x 5). On return lane c (< 28) holds the warp
// total of component c. Deterministic order.
__device__ __forceinline__ double warp_reduce28(const double* acc, int lane) {
  double v16[16], v8[8], v4[4], v2[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2, h1 = lane & 1;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const double hi = (i + 16 < 28) ? acc[i + 16] : 0.0;
    const double send = h16 ? acc[i] : hi, keep = h16 ? hi : acc[i];
    v16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const double send = h8 ? v16[i] : v16[i + 8], keep = h8 ? v16[i + 8] : v16[i];
    v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const double send = h4 ? v8[i] : v8[i + 4], keep = h4 ? v8[i + 4] : v8[i];
    v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const double send = h2 ? v4[i] : v4[i + 2], keep = h2 ? v4[i + 2] : v4[i];
    v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  const double send = h1 ? v2[0] : v2[1], keep = h1 ? v2[1] : v2[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

// R, t, J_r of the evaluation point: rotation on one thread, right Jacobian on another (different warps)
__device__ __forceinline__ void make_pose_split(const double* x6, const double* Rbl, const double* Pbl, PoseLin& L, int tid) {
  if (tid == 0) {
    const Quat q = so3_exp(x6 + 3);
    quat_to_R(q, L.R);
    L.t[0] = x6[0]; L.t[1] = x6[1]; L.t[2] = x6[2];
  } else if (tid == 32) {
    right_jacobian(x6 + 3, L.Jr);
  } else if (tid == 33) {
    for (int i = 0; i < 9; i++) L.Rbl[i] = Rbl[i];
    for (int i = 0; i < 3; i++) L.Pbl[i] = Pbl[i];
  }
}

// thread-block cluster primitives (PTX: barrier.cluster, mapa, ld.shared::cluster)
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem_f64(const double* local_smem_ptr, unsigned cta_rank) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(local_smem_ptr);
  unsigned ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(cta_rank));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}

__device__ __forceinline__ void st_dsmem_f64(double* local_smem_ptr, unsigned cta_rank, double v) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(local_smem_ptr);
  unsigned ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(cta_rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}

}  // namespace mml
