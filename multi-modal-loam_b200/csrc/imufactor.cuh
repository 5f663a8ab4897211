// IMU factor of the sliding-window solve, shared by the host solver (window.cu) and the device solver
// (windowsolve.cu). Reference: Cost_NavState_PRV_Bias::operator(), include/utils/ceresfunc.h:321-393, evaluated under
// forward-mode differentiation like ceres::AutoDiffCostFunction does (Jet<double, 30>): the functor text is
// instantiated for plain doubles, for a dual number over all 30 parameters (host: one thread forms the whole
// Jacobian) and for a dual number over ONE parameter (device: one thread per Jacobian column; the derivative
// arithmetic of a column is the same operation sequence in both forms, so both give the same values).
#pragma once
#include <cmath>
#include "../../include/mmloam_b200.h"

namespace mml {

// ---- scalar overloads used by the functor text below when it is instantiated for plain doubles
__host__ __device__ inline double dsqrt(double f) { return sqrt(f); }
__host__ __device__ inline double dsin(double f) { return sin(f); }
__host__ __device__ inline double dcos(double f) { return cos(f); }
__host__ __device__ inline double datan(double f) { return atan(f); }
__host__ __device__ inline double val(double x) { return x; }
// sine and cosine of the same argument (the exponential map needs both): one call where the argument reduction can be shared
__host__ __device__ inline void dsincos(double f, double& sn, double& cs) { sn = sin(f); cs = cos(f); }

template <class T> struct Q4 { T w, x, y, z; };
template <class T> __host__ __device__ inline Q4<T> qmul(const Q4<T>& a, const Q4<T>& b) {  // sophus/so3.hpp:326-340
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
template <class T> __host__ __device__ inline Q4<T> qconj(const Q4<T>& q) { return {q.w, -q.x, -q.y, -q.z}; }
template <class T> __host__ __device__ inline void qrot(const Q4<T>& q, const T* p, T* out) {  // so3.hpp:358-371
  T uv[3] = {q.y * p[2] - q.z * p[1], q.z * p[0] - q.x * p[2], q.x * p[1] - q.y * p[0]};
  for (int k = 0; k < 3; k++) uv[k] = uv[k] + uv[k];
  T c[3] = {q.y * uv[2] - q.z * uv[1], q.z * uv[0] - q.x * uv[2], q.x * uv[1] - q.y * uv[0]};
  for (int k = 0; k < 3; k++) out[k] = p[k] + q.w * uv[k] + c[k];
}
template <class T> __host__ __device__ inline Q4<T> qexp(const T* om) {  // so3.hpp:585-623, epsilon 1e-10
  T theta_sq = (om[0] * om[0] + om[1] * om[1]) + om[2] * om[2];
  T imag, real;
  if (val(theta_sq) < 1e-20) {
    T theta_po4 = theta_sq * theta_sq;
    imag = T(0.5) - T(1.0 / 48.0) * theta_sq + T(1.0 / 3840.0) * theta_po4;
    real = T(1.0) - T(1.0 / 8.0) * theta_sq + T(1.0 / 384.0) * theta_po4;
  } else {
    T theta = dsqrt(theta_sq);
    T half = T(0.5) * theta;
    T sn;
    dsincos(half, sn, real);
    imag = sn / theta;
  }
  return {real, imag * om[0], imag * om[1], imag * om[2]};
}
template <class T> __host__ __device__ inline void qlog(const Q4<T>& q, T* out) {  // so3.hpp:247-292
  T squared_n = (q.x * q.x + q.y * q.y) + q.z * q.z;
  T w = q.w;
  T f;
  if (val(squared_n) < 1e-20) {
    T squared_w = w * w;
    f = T(2.0) / w - T(2.0 / 3.0) * squared_n / (w * squared_w);
  } else {
    T n = dsqrt(squared_n);
    if (fabs(val(w)) < 1e-10) f = T(val(w) > 0 ? 3.14159265358979323846 : -3.14159265358979323846) / n;
    else f = T(2.0) * datan(n / w) / n;
  }
  out[0] = f * q.x; out[1] = f * q.y; out[2] = f * q.z;
}

// Cost_NavState_PRV_Bias::operator(), CF.h:331-377, before the multiplication by sqrt_information
template <class T>
__host__ __device__ void imu_residual(const mml_preint& m, const double* g, const T* pri, const T* vbi, const T* prj,
                                      const T* vbj, T* r) {
  const Q4<T> Ri = qexp(pri + 3), Rj = qexp(prj + 3);
  T dbg[3], dba[3];
  for (int k = 0; k < 3; k++) { dbg[k] = vbi[3 + k] - T(m.bg[k]); dba[k] = vbi[6 + k] - T(m.ba[k]); }
  const double dT = m.dt, dT2 = m.dt * m.dt;
  Q4<T> dRij;
  {  // Sophus::SO3<T>(quaternion) normalises, so3.hpp:487-494
    const double n = sqrt(((m.dq[1] * m.dq[1] + m.dq[2] * m.dq[2]) + m.dq[3] * m.dq[3]) + m.dq[0] * m.dq[0]);
    dRij = {T(m.dq[0] / n), T(m.dq[1] / n), T(m.dq[2] / n), T(m.dq[3] / n)};
  }
  const Q4<T> RiT = qconj(Ri);
#define MML_J(r0, c0, r, c) m.jac[((r0) + (r)) * 15 + (c0) + (c)]
  T a[3], ra[3];
  for (int k = 0; k < 3; k++) a[k] = prj[k] - pri[k] - vbi[k] * T(dT) - T(0.5 * g[k]) * T(dT2);
  qrot(RiT, a, ra);
  for (int k = 0; k < 3; k++) {
    T c = T(m.dp[k]) + ((T(MML_J(0, 9, k, 0)) * dbg[0] + T(MML_J(0, 9, k, 1)) * dbg[1]) + T(MML_J(0, 9, k, 2)) * dbg[2]) +
          ((T(MML_J(0, 12, k, 0)) * dba[0] + T(MML_J(0, 12, k, 1)) * dba[1]) + T(MML_J(0, 12, k, 2)) * dba[2]);
    r[k] = ra[k] - c;
  }
  T w[3];
  for (int k = 0; k < 3; k++) w[k] = (T(MML_J(3, 9, k, 0)) * dbg[0] + T(MML_J(3, 9, k, 1)) * dbg[1]) + T(MML_J(3, 9, k, 2)) * dbg[2];
  const Q4<T> dR_dbg = qexp(w);
  const Q4<T> rR = qmul(qmul(qconj(qmul(dRij, dR_dbg)), RiT), Rj);
  qlog(rR, r + 3);
  for (int k = 0; k < 3; k++) a[k] = vbj[k] - vbi[k] - T(g[k]) * T(dT);
  qrot(RiT, a, ra);
  for (int k = 0; k < 3; k++) {
    T c = T(m.dv[k]) + ((T(MML_J(6, 9, k, 0)) * dbg[0] + T(MML_J(6, 9, k, 1)) * dbg[1]) + T(MML_J(6, 9, k, 2)) * dbg[2]) +
          ((T(MML_J(6, 12, k, 0)) * dba[0] + T(MML_J(6, 12, k, 1)) * dba[1]) + T(MML_J(6, 12, k, 2)) * dba[2]);
    r[6 + k] = ra[k] - c;
  }
#undef MML_J
  for (int k = 0; k < 6; k++) r[9 + k] = vbj[3 + k] - vbi[3 + k];
}

// Forward-mode dual number over the 30 parameters of the IMU factor: what Ceres' Jet<double, 30> is. The functor is
// differentiated automatically, exactly as the reference does (same derivative values, no hand-derived Jacobian).
struct Dual30 {
  double a, v[30];
  __host__ __device__ Dual30() : a(0) { for (int i = 0; i < 30; i++) v[i] = 0; }
  __host__ __device__ Dual30(double s) : a(s) { for (int i = 0; i < 30; i++) v[i] = 0; }
};
__host__ __device__ inline Dual30 operator+(const Dual30& f, const Dual30& g) { Dual30 h; h.a = f.a + g.a; for (int i = 0; i < 30; i++) h.v[i] = f.v[i] + g.v[i]; return h; }
__host__ __device__ inline Dual30 operator-(const Dual30& f, const Dual30& g) { Dual30 h; h.a = f.a - g.a; for (int i = 0; i < 30; i++) h.v[i] = f.v[i] - g.v[i]; return h; }
__host__ __device__ inline Dual30 operator-(const Dual30& f) { Dual30 h; h.a = -f.a; for (int i = 0; i < 30; i++) h.v[i] = -f.v[i]; return h; }
__host__ __device__ inline Dual30 operator*(const Dual30& f, const Dual30& g) { Dual30 h; h.a = f.a * g.a; for (int i = 0; i < 30; i++) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
__host__ __device__ inline Dual30 operator/(const Dual30& f, const Dual30& g) { Dual30 h; const double gi = 1.0 / g.a, q = f.a * gi; h.a = q; for (int i = 0; i < 30; i++) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h; }
__host__ __device__ inline Dual30 chain30(double val_, double d, const Dual30& f) { Dual30 h; h.a = val_; for (int i = 0; i < 30; i++) h.v[i] = d * f.v[i]; return h; }
__host__ __device__ inline Dual30 dsqrt(const Dual30& f) { const double t = sqrt(f.a); return chain30(t, 1.0 / (2.0 * t), f); }
__host__ __device__ inline Dual30 dsin(const Dual30& f) { return chain30(sin(f.a), cos(f.a), f); }
__host__ __device__ inline Dual30 dcos(const Dual30& f) { return chain30(cos(f.a), -sin(f.a), f); }
__host__ __device__ inline Dual30 datan(const Dual30& f) { return chain30(atan(f.a), 1.0 / (1.0 + f.a * f.a), f); }
__host__ __device__ inline double val(const Dual30& x) { return x.a; }
__host__ __device__ inline void dsincos(const Dual30& f, Dual30& sn, Dual30& cs) { sn = dsin(f); cs = dcos(f); }

// one derivative direction: the device evaluates column c of the 15 x 30 Jacobian on its own thread
struct Dual1 {
  double a, v;
  __host__ __device__ Dual1() : a(0), v(0) {}
  __host__ __device__ Dual1(double s) : a(s), v(0) {}
  __host__ __device__ Dual1(double s, double d) : a(s), v(d) {}
};
__host__ __device__ inline Dual1 operator+(const Dual1& f, const Dual1& g) { return {f.a + g.a, f.v + g.v}; }
__host__ __device__ inline Dual1 operator-(const Dual1& f, const Dual1& g) { return {f.a - g.a, f.v - g.v}; }
__host__ __device__ inline Dual1 operator-(const Dual1& f) { return {-f.a, -f.v}; }
__host__ __device__ inline Dual1 operator*(const Dual1& f, const Dual1& g) { return {f.a * g.a, f.a * g.v + f.v * g.a}; }
__host__ __device__ inline Dual1 operator/(const Dual1& f, const Dual1& g) { const double gi = 1.0 / g.a, q = f.a * gi; return {q, (f.v - q * g.v) * gi}; }
__host__ __device__ inline Dual1 dsqrt(const Dual1& f) { const double t = sqrt(f.a); return {t, (1.0 / (2.0 * t)) * f.v}; }
__host__ __device__ inline Dual1 dsin(const Dual1& f) { return {sin(f.a), cos(f.a) * f.v}; }
__host__ __device__ inline Dual1 dcos(const Dual1& f) { return {cos(f.a), -sin(f.a) * f.v}; }
__host__ __device__ inline Dual1 datan(const Dual1& f) { return {atan(f.a), (1.0 / (1.0 + f.a * f.a)) * f.v}; }
__host__ __device__ inline double val(const Dual1& x) { return x.a; }
__host__ __device__ inline void dsincos(const Dual1& f, Dual1& sn, Dual1& cs) {
  double sv, cv;
#ifdef __CUDA_ARCH__
  sincos(f.a, &sv, &cv);  // one argument reduction for both
#else
  sv = sin(f.a); cv = cos(f.a);
#endif
  sn = {sv, cv * f.v};
  cs = {cv, -sv * f.v};
}

}  // namespace mml
