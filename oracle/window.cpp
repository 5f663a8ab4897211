// ORACLE (test infrastructure only — see oracle.h). Sliding-window rows (window sizes 2..4: IMU factors,
// no marginalisation; BASELINE config 3 is window 3):
//   orc_imu_preintegrate()  <- IMUIntegrator::PreIntegration          src/lio/IMUIntegrator.cpp:105-166
//   orc_imu_factor()        <- Cost_NavState_PRV_Bias::operator()     include/utils/ceresfunc.h:321-393
//                              with sqrt_information = LLT(cov^-1).matrixL().transpose()  EST.cpp:1240-1242
//   orc_estimate_window()   <- Estimator::Estimate for windowSize < 5  src/lio/Estimator.cpp:1143-1581
//   orc_imu_predict()       <- the pose prediction of process()       src/unionPoseEstimation.cpp:812-829
// Ceres differentiates the IMU functor automatically; the oracle does the same with forward-mode dual numbers
// over the 30 parameters (the functor text is restated once, templated on the scalar). The lidar rows keep the
// closed forms of residual.cpp. Checked against the reference text itself in tests/test_ref_pin.py.
#include "oracle.h"
#include "oracle_math.h"
#include "dogleg.h"
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {
void accumulate_pose(const double* lf, int nl, const double* pf, int np, const double* x6, const double* Tbl,
                     double lidar_m, double w_tan, double a, double* H36, double* g6, double* cost, int threads);

// ---- dual numbers over N directions ---------------------------------------------------------
template <int N> struct Dual {
  double a; double v[N];
  Dual() : a(0) { for (int i = 0; i < N; i++) v[i] = 0; }
  Dual(double s) : a(s) { for (int i = 0; i < N; i++) v[i] = 0; }
};
#define DUAL template <int N> inline Dual<N>
DUAL operator+(const Dual<N>& f, const Dual<N>& g) { Dual<N> h; h.a = f.a + g.a; for (int i = 0; i < N; i++) h.v[i] = f.v[i] + g.v[i]; return h; }
DUAL operator-(const Dual<N>& f, const Dual<N>& g) { Dual<N> h; h.a = f.a - g.a; for (int i = 0; i < N; i++) h.v[i] = f.v[i] - g.v[i]; return h; }
DUAL operator-(const Dual<N>& f) { Dual<N> h; h.a = -f.a; for (int i = 0; i < N; i++) h.v[i] = -f.v[i]; return h; }
DUAL operator*(const Dual<N>& f, const Dual<N>& g) { Dual<N> h; h.a = f.a * g.a; for (int i = 0; i < N; i++) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
DUAL operator/(const Dual<N>& f, const Dual<N>& g) { Dual<N> h; double gi = 1.0 / g.a, fg = f.a * gi; h.a = fg; for (int i = 0; i < N; i++) h.v[i] = (f.v[i] - fg * g.v[i]) * gi; return h; }
DUAL chain(double val, double d, const Dual<N>& f) { Dual<N> h; h.a = val; for (int i = 0; i < N; i++) h.v[i] = d * f.v[i]; return h; }
DUAL sqrt(const Dual<N>& f) { double t = std::sqrt(f.a); return chain(t, 1.0 / (2.0 * t), f); }
DUAL sin(const Dual<N>& f) { return chain(std::sin(f.a), std::cos(f.a), f); }
DUAL cos(const Dual<N>& f) { return chain(std::cos(f.a), -std::sin(f.a), f); }
DUAL atan(const Dual<N>& f) { return chain(std::atan(f.a), 1.0 / (1.0 + f.a * f.a), f); }
#undef DUAL
template <int N> inline bool operator<(const Dual<N>& f, double s) { return f.a < s; }
template <int N> inline bool operator>(const Dual<N>& f, double s) { return f.a > s; }
inline double val(double x) { return x; }
template <int N> inline double val(const Dual<N>& x) { return x.a; }
using std::sqrt; using std::sin; using std::cos; using std::atan;

// ---- quaternion algebra as Sophus / Eigen spell it, generic in the scalar ----------------------
template <class T> struct Q4 { T w, x, y, z; };
template <class T> Q4<T> qmul(const Q4<T>& a, const Q4<T>& b) {  // so3.hpp:326-340
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
template <class T> Q4<T> qconj(const Q4<T>& q) { return {q.w, -q.x, -q.y, -q.z}; }
template <class T> Q4<T> qnormalized(const Q4<T>& q) {  // SO3(quaternion) normalises, so3.hpp:302-308
  T n = sqrt(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}
template <class T> void qrot(const Q4<T>& q, const T* p, T* out) {  // so3.hpp:358-371
  T uv[3] = {q.y * p[2] - q.z * p[1], q.z * p[0] - q.x * p[2], q.x * p[1] - q.y * p[0]};
  for (int k = 0; k < 3; k++) uv[k] = uv[k] + uv[k];
  T c[3] = {q.y * uv[2] - q.z * uv[1], q.z * uv[0] - q.x * uv[2], q.x * uv[1] - q.y * uv[0]};
  for (int k = 0; k < 3; k++) out[k] = p[k] + q.w * uv[k] + c[k];
}
template <class T> Q4<T> qexp(const T* om) {  // so3.hpp:585-623
  T theta_sq = (om[0] * om[0] + om[1] * om[1]) + om[2] * om[2];
  T imag, real;
  if (theta_sq < 1e-10 * 1e-10) {
    T theta_po4 = theta_sq * theta_sq;
    imag = T(0.5) - T(1.0 / 48.0) * theta_sq + T(1.0 / 3840.0) * theta_po4;
    real = T(1.0) - T(1.0 / 8.0) * theta_sq + T(1.0 / 384.0) * theta_po4;
  } else {
    T theta = sqrt(theta_sq);
    T half = T(0.5) * theta;
    imag = sin(half) / theta;
    real = cos(half);
  }
  return {real, imag * om[0], imag * om[1], imag * om[2]};
}
template <class T> void qlog(const Q4<T>& q, T* out) {  // so3.hpp:247-292
  T squared_n = (q.x * q.x + q.y * q.y) + q.z * q.z;
  T w = q.w;
  T f;
  if (squared_n < 1e-10 * 1e-10) {
    T squared_w = w * w;
    f = T(2.0) / w - T(2.0 / 3.0) * squared_n / (w * squared_w);
  } else {
    T n = sqrt(squared_n);
    if (std::fabs(val(w)) < 1e-10) f = T(val(w) > 0 ? M_PI : -M_PI) / n;
    else f = T(2.0) * atan(n / w) / n;
  }
  out[0] = f * q.x; out[1] = f * q.y; out[2] = f * q.z;
}

struct Preint {
  double dq[4];       // w x y z
  double dp[3], dv[3], dt;
  double bg[3], ba[3];  // linearisation point
  double cov[225], jac[225];  // row-major 15x15, order P R V BG BA (IMU.h:86-93)
  double sqrt_info[225];      // LLT(cov^-1).matrixL().transpose(), row-major
};

// CF.h:331-377, unweighted residual (15); the caller applies sqrt_information.
template <class T>
void imu_residual(const Preint& m, const double* g, const T* pri, const T* vbi, const T* prj, const T* vbj, T* r) {
  const T* Pi = pri; const T* Pj = prj;
  Q4<T> Ri = qexp(pri + 3), Rj = qexp(prj + 3);
  const T* Vi = vbi; const T* Vj = vbj;
  T dbg[3], dba[3];
  for (int k = 0; k < 3; k++) { dbg[k] = vbi[3 + k] - T(m.bg[k]); dba[k] = vbi[6 + k] - T(m.ba[k]); }
  const double dT = m.dt, dT2 = m.dt * m.dt;
  Q4<T> dRij = qnormalized(Q4<T>{T(m.dq[0]), T(m.dq[1]), T(m.dq[2]), T(m.dq[3])});
  Q4<T> RiT = qconj(Ri);
  auto J = [&](int r0, int c0, int r, int c) { return m.jac[(r0 + r) * 15 + c0 + c]; };
  // rPij
  T a[3], ra[3];
  for (int k = 0; k < 3; k++) a[k] = Pj[k] - Pi[k] - Vi[k] * T(dT) - T(0.5 * g[k]) * T(dT2);
  qrot(RiT, a, ra);
  for (int k = 0; k < 3; k++) {
    T c = T(m.dp[k]) + ((T(J(0, 9, k, 0)) * dbg[0] + T(J(0, 9, k, 1)) * dbg[1]) + T(J(0, 9, k, 2)) * dbg[2]) +
          ((T(J(0, 12, k, 0)) * dba[0] + T(J(0, 12, k, 1)) * dba[1]) + T(J(0, 12, k, 2)) * dba[2]);
    r[k] = ra[k] - c;
  }
  // rPhiij
  T w[3];
  for (int k = 0; k < 3; k++) w[k] = (T(J(3, 9, k, 0)) * dbg[0] + T(J(3, 9, k, 1)) * dbg[1]) + T(J(3, 9, k, 2)) * dbg[2];
  Q4<T> dR_dbg = qexp(w);
  Q4<T> rR = qmul(qmul(qconj(qmul(dRij, dR_dbg)), RiT), Rj);
  qlog(rR, r + 3);
  // rVij
  for (int k = 0; k < 3; k++) a[k] = Vj[k] - Vi[k] - T(g[k]) * T(dT);
  qrot(RiT, a, ra);
  for (int k = 0; k < 3; k++) {
    T c = T(m.dv[k]) + ((T(J(6, 9, k, 0)) * dbg[0] + T(J(6, 9, k, 1)) * dbg[1]) + T(J(6, 9, k, 2)) * dbg[2]) +
          ((T(J(6, 12, k, 0)) * dba[0] + T(J(6, 12, k, 1)) * dba[1]) + T(J(6, 12, k, 2)) * dba[2]);
    r[6 + k] = ra[k] - c;
  }
  for (int k = 0; k < 6; k++) r[9 + k] = vbj[3 + k] - vbi[3 + k];
}

// r15 = sqrt_info * r, J (15 x 30, row-major, columns [pri 6 | vbi 9 | prj 6 | vbj 9])
void imu_factor(const Preint& m, const double* g, const double* pri, const double* vbi, const double* prj,
                const double* vbj, double* r15, double* J450) {
  using D = Dual<30>;
  D x[30];
  const double* src[4] = {pri, vbi, prj, vbj};
  const int sz[4] = {6, 9, 6, 9};
  int o = 0;
  for (int b = 0; b < 4; b++) for (int k = 0; k < sz[b]; k++, o++) { x[o] = D(src[b][k]); x[o].v[o] = 1.0; }
  D r[15];
  imu_residual<D>(m, g, x, x + 6, x + 15, x + 21, r);
  for (int i = 0; i < 15; i++) {
    double s = 0;
    for (int k = 0; k < 15; k++) s += m.sqrt_info[i * 15 + k] * r[k].a;
    r15[i] = s;
    if (J450) for (int c = 0; c < 30; c++) {
      double t = 0;
      for (int k = 0; k < 15; k++) t += m.sqrt_info[i * 15 + k] * r[k].v[c];
      J450[i * 30 + c] = t;
    }
  }
}

static void mat_mul(int n, int k, int m, const double* A, const double* B, double* C) {  // C[n x m] = A[n x k] B[k x m]
  for (int i = 0; i < n; i++) for (int j = 0; j < m; j++) {
    double s = 0;
    for (int t = 0; t < k; t++) s += A[i * k + t] * B[t * m + j];
    C[i * m + j] = s;
  }
}
static void hat(const double* v, double* K) { K[0] = 0; K[1] = -v[2]; K[2] = v[1]; K[3] = v[2]; K[4] = 0; K[5] = -v[0]; K[6] = -v[1]; K[7] = v[0]; K[8] = 0; }

// Gauss-Jordan inverse with partial pivoting (stands in for Eigen's PartialPivLU inverse)
static bool invert(int n, const double* A, double* inv) {
  std::vector<double> a((size_t)n * 2 * n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { a[(size_t)i * 2 * n + j] = A[i * n + j]; a[(size_t)i * 2 * n + n + j] = (i == j); }
  for (int c = 0; c < n; c++) {
    int p = c;
    for (int i = c + 1; i < n; i++) if (std::fabs(a[(size_t)i * 2 * n + c]) > std::fabs(a[(size_t)p * 2 * n + c])) p = i;
    if (a[(size_t)p * 2 * n + c] == 0.0) return false;
    if (p != c) for (int j = 0; j < 2 * n; j++) std::swap(a[(size_t)p * 2 * n + j], a[(size_t)c * 2 * n + j]);
    double d = a[(size_t)c * 2 * n + c];
    for (int j = 0; j < 2 * n; j++) a[(size_t)c * 2 * n + j] /= d;
    for (int i = 0; i < n; i++) if (i != c) {
      double f = a[(size_t)i * 2 * n + c];
      if (f == 0.0) continue;
      for (int j = 0; j < 2 * n; j++) a[(size_t)i * 2 * n + j] -= f * a[(size_t)c * 2 * n + j];
    }
  }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) inv[i * n + j] = a[(size_t)i * 2 * n + n + j];
  return true;
}

// IMU.cpp:105-166. t / gyr / acc: n samples (acc in units of g: scaled by gnorm = 9.805, IMU.h:84).
void preintegrate(const double* t, const double* gyr, const double* acc, int n, double last_time, const double* bg,
                  const double* ba, Preint& m) {
  const double acc_n = 0.08, gyr_n = 0.004, acc_w = 2.0e-4, gyr_w = 2.0e-5, gnorm = 9.805;  // IMU.h:79-84
  Quat dq = {1, 0, 0, 0};
  double dp[3] = {0, 0, 0}, dv[3] = {0, 0, 0}, dtime = 0;
  std::vector<double> cov(225, 0.0), jac(225, 0.0), noise(144, 0.0);
  for (int i = 0; i < 15; i++) jac[i * 15 + i] = 1.0;
  for (int i = 0; i < 3; i++) {
    noise[i * 12 + i] = gyr_n * gyr_n; noise[(3 + i) * 12 + 3 + i] = acc_n * acc_n;
    noise[(6 + i) * 12 + 6 + i] = gyr_w * gyr_w; noise[(9 + i) * 12 + 9 + i] = acc_w * acc_w;
  }
  double current_time = last_time;
  for (int s = 0; s < n; s++) {
    double g3[3] = {gyr[3 * s] - bg[0], gyr[3 * s + 1] - bg[1], gyr[3 * s + 2] - bg[2]};
    double a3[3] = {acc[3 * s] * gnorm - ba[0], acc[3 * s + 1] * gnorm - ba[1], acc[3 * s + 2] * gnorm - ba[2]};
    double dt = t[s] - current_time, dt2 = dt * dt;
    double gdt[3] = {g3[0] * dt, g3[1] * dt, g3[2] * dt};
    double dR[9];
    quat_to_R(so3_exp(gdt), dR);
    double Jr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double nrm = std::sqrt((gdt[0] * gdt[0] + gdt[1] * gdt[1]) + gdt[2] * gdt[2]);
    if (nrm > 0.00001) {
      double k[3] = {gdt[0] / nrm, gdt[1] / nrm, gdt[2] / nrm};
      double K[9], KK[9];
      hat(k, K);
      mat_mul(3, 3, 3, K, K, KK);
      double c1 = (1 - std::cos(nrm)) / nrm, c2 = 1 - std::sin(nrm) / nrm;
      for (int i = 0; i < 9; i++) Jr[i] = (i % 4 == 0 ? 1.0 : 0.0) - c1 * K[i] + c2 * KK[i];
    }
    double Rq[9], Ha[9], RH[9];
    quat_to_R(dq, Rq);
    hat(a3, Ha);
    mat_mul(3, 3, 3, Rq, Ha, RH);
    double A[225], B[180];
    std::memset(A, 0, sizeof(A)); std::memset(B, 0, sizeof(B));
    for (int i = 0; i < 15; i++) A[i * 15 + i] = 1.0;
    auto setA = [&](int r0, int c0, const double* M, double f) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) A[(r0 + r) * 15 + c0 + c] = f * M[3 * r + c]; };
    auto setB = [&](int r0, int c0, const double* M, double f) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) B[(r0 + r) * 12 + c0 + c] = f * M[3 * r + c]; };
    const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double dRT[9] = {dR[0], dR[3], dR[6], dR[1], dR[4], dR[7], dR[2], dR[5], dR[8]};
    setA(0, 3, RH, -0.5 * dt2); setA(0, 6, I3, dt); setA(0, 12, Rq, -0.5 * dt2);
    setA(3, 3, dRT, 1.0); setA(3, 9, Jr, -dt);
    setA(6, 3, RH, -dt); setA(6, 12, Rq, -dt);
    setB(0, 3, Rq, 0.5 * dt2); setB(3, 0, Jr, dt); setB(6, 3, Rq, dt); setB(9, 6, I3, dt); setB(12, 9, I3, dt);
    std::vector<double> tmp(225), tmp2(225), AT(225), BN(180), BT(180), BNB(225);
    mat_mul(15, 15, 15, A, jac.data(), tmp.data());
    jac = tmp;
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) AT[i * 15 + j] = A[j * 15 + i];
    mat_mul(15, 15, 15, A, cov.data(), tmp.data());
    mat_mul(15, 15, 15, tmp.data(), AT.data(), tmp2.data());
    mat_mul(15, 12, 12, B, noise.data(), BN.data());
    for (int i = 0; i < 12; i++) for (int j = 0; j < 15; j++) BT[i * 15 + j] = B[j * 12 + i];
    mat_mul(15, 12, 15, BN.data(), BT.data(), BNB.data());
    for (int i = 0; i < 225; i++) cov[i] = tmp2[i] + BNB[i];
    double Ra[3] = {Rq[0] * a3[0] + Rq[1] * a3[1] + Rq[2] * a3[2], Rq[3] * a3[0] + Rq[4] * a3[1] + Rq[5] * a3[2],
                    Rq[6] * a3[0] + Rq[7] * a3[1] + Rq[8] * a3[2]};
    for (int k = 0; k < 3; k++) dp[k] += dv[k] * dt + 0.5 * Ra[k] * dt2;
    for (int k = 0; k < 3; k++) dv[k] += Ra[k] * dt;
    double m3[9];
    mat_mul(3, 3, 3, Rq, dR, m3);
    Quat qt = quat_from_R(m3);
    if (qt.w < 0) { qt.w = -qt.w; qt.x = -qt.x; qt.y = -qt.y; qt.z = -qt.z; }
    dq = quat_normalized(qt);
    dtime += dt;
    current_time = t[s];
  }
  m.dq[0] = dq.w; m.dq[1] = dq.x; m.dq[2] = dq.y; m.dq[3] = dq.z;
  for (int k = 0; k < 3; k++) { m.dp[k] = dp[k]; m.dv[k] = dv[k]; m.bg[k] = bg[k]; m.ba[k] = ba[k]; }
  m.dt = dtime;
  std::memcpy(m.cov, cov.data(), sizeof(m.cov));
  std::memcpy(m.jac, jac.data(), sizeof(m.jac));
  // EST.cpp:1240-1242: LLT(cov^-1).matrixL().transpose()
  double inv[225], L[225];
  std::memset(L, 0, sizeof(L));
  std::memset(m.sqrt_info, 0, sizeof(m.sqrt_info));
  if (n > 0 && invert(15, m.cov, inv)) {
    for (int i = 0; i < 15; i++) for (int j = 0; j <= i; j++) {
      double s = inv[i * 15 + j];
      for (int k = 0; k < j; k++) s -= L[i * 15 + k] * L[j * 15 + k];
      L[i * 15 + j] = (i == j) ? std::sqrt(s) : s / L[j * 15 + j];
    }
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) m.sqrt_info[i * 15 + j] = L[j * 15 + i];
  }
}
}  // namespace orc

using namespace orc;

extern "C" int orc_associate_line_mt(const orc_map* m, const float* q_xyzi, int nq, const double* T, double thres_dist,
                                     double* feat, int* n_feat, int threads);
extern "C" int orc_associate_plane_mt(const orc_map* m, const float* q_xyzi, int nq, const double* T, double thres_dist,
                                      double* feat, int* n_feat, double* M9, int* n_normals, int threads);

extern "C" {

int orc_preint_size(void) { return (int)sizeof(Preint); }

int orc_imu_preintegrate(const double* t, const double* gyr, const double* acc, int n, double last_time,
                         const double* bg3, const double* ba3, void* preint_out) {
  preintegrate(t, gyr, acc, n, last_time, bg3, ba3, *(Preint*)preint_out);
  return 0;
}

int orc_imu_factor(const void* preint, const double* gravity3, const double* pri6, const double* vbi9,
                   const double* prj6, const double* vbj9, double* r15, double* J450) {
  imu_factor(*(const Preint*)preint, gravity3, pri6, vbi9, prj6, vbj9, r15, J450);
  return 0;
}

// PE.cpp:812-829: state of the new frame predicted from the previous one and the pre-integration.
// state rows: P(3) q_wxyz(4) V(3) bg(3) ba(3) = 16 doubles.
int orc_imu_predict(const double* prev16, const void* preint, double* next16) {
  const Preint& m = *(const Preint*)preint;
  Quat Qp = {prev16[3], prev16[4], prev16[5], prev16[6]};
  Quat dQ = {m.dq[0], m.dq[1], m.dq[2], m.dq[3]};
  Quat Q = quat_mul(Qp, dQ);
  double rp[3], rv[3];
  quat_rotate(Qp, m.dp, rp);
  quat_rotate(Qp, m.dv, rv);
  for (int k = 0; k < 3; k++) { next16[k] = prev16[k] + rp[k]; next16[7 + k] = prev16[7 + k] + rv[k]; }
  next16[3] = Q.w; next16[4] = Q.x; next16[5] = Q.y; next16[6] = Q.z;
  for (int k = 0; k < 6; k++) next16[10 + k] = prev16[10 + k];
  return 0;
}

// Estimator::Estimate (EST.cpp:1143-1581) for 1 <= W <= 4: every outer iteration re-associates all W frames,
// IMU factors between consecutive frames (no loss), lidar factors under Huber, convergence tested on the LAST
// frame (EST.cpp:1256-1257, 1441-1448). states: W x 16 doubles, in place. preints[f] (f >= 1) links f-1 -> f.
// stats (optional, 16): [outer, inner_total, n_line(last frame), n_plane(last frame), final_cost, min_sv, degenerate].
int orc_estimate_window(const orc_map* map, int W, const float* const* corner, const int* n_corner,
                        const float* const* surf, const int* n_surf, const double* exTlb, double* states,
                        const void* const* preints, const double* gravity3, const orc_est_params* prm, double* stats) {
  if (W < 1 || W > 4) return -1;
  double Rbl[9], Pbl[3];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Rbl[3 * r + c] = exTlb[4 * c + r];
  for (int r = 0; r < 3; r++) Pbl[r] = -1.0 * (Rbl[3 * r] * exTlb[3] + Rbl[3 * r + 1] * exTlb[7] + Rbl[3 * r + 2] * exTlb[11]);
  double Tbl[16] = {Rbl[0], Rbl[1], Rbl[2], Pbl[0], Rbl[3], Rbl[4], Rbl[5], Pbl[1], Rbl[6], Rbl[7], Rbl[8], Pbl[2], 0, 0, 0, 1};
  std::vector<std::vector<double>> lf(W), pf(W);
  for (int f = 0; f < W; f++) { lf[f].resize((size_t)std::max(n_corner[f], 1) * 12); pf[f].resize((size_t)std::max(n_surf[f], 1) * 12); }
  double thres = prm->thres0;
  const double huber_a = prm->use_huber ? 0.1 / prm->lidar_m : 0.0;
  int outer_done = 0, inner_total = 0, nl = 0, np = 0, is_degenerate = 0;
  double final_cost = 0, min_sv = -1;
  const int n = (W == 1) ? 6 : 15 * W;  // a lone frame's velocity / bias block has no residual: Ceres drops it

  for (int it = 0; it < prm->max_outer; ++it) {
    // vector2double, EST.cpp:937-950: x = [PR_0 .. PR_{W-1} | VB_0 .. VB_{W-1}]
    std::vector<double> x(15 * W);
    for (int f = 0; f < W; f++) {
      const double* s = states + 16 * f;
      for (int k = 0; k < 3; k++) x[6 * f + k] = s[k];
      so3_log(Quat{s[3], s[4], s[5], s[6]}, &x[6 * f + 3]);
      for (int k = 0; k < 9; k++) x[6 * W + 9 * f + k] = s[7 + k];
    }
    const double* sb = states + 16 * (W - 1);
    Quat q_before = {sb[3], sb[4], sb[5], sb[6]};
    double t_before[3] = {sb[0], sb[1], sb[2]};
    for (int f = 0; f < W; f++) {
      const double* s = states + 16 * f;
      double Rq[9], T[16] = {0};
      quat_to_R(Quat{s[3], s[4], s[5], s[6]}, Rq);
      for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) T[4 * r + c] = Rq[3 * r] * Rbl[c] + Rq[3 * r + 1] * Rbl[3 + c] + Rq[3 * r + 2] * Rbl[6 + c];
        T[4 * r + 3] = Rq[3 * r] * Pbl[0] + Rq[3 * r + 1] * Pbl[1] + Rq[3 * r + 2] * Pbl[2] + s[r];
      }
      T[15] = 1;
      double M9[9];
      int nn = 0;
      orc_associate_line_mt(map, corner[f], n_corner[f], T, thres, lf[f].data(), &nl, prm->threads);
      orc_associate_plane_mt(map, surf[f], n_surf[f], T, thres, pf[f].data(), &np, M9, &nn, prm->threads);
      min_sv = orc_localizability(M9, nn);
      if (min_sv < 3.0) is_degenerate = 1;  // EST.cpp:771-775, never cleared inside Estimate
    }
    thres = (it == 0) ? prm->thres1 : prm->thres2;

    EvalFn eval = [&](const double* xx, double* cost, double* H, double* g) {
      double c = 0;
      if (H) std::fill(H, H + (size_t)n * n, 0.0);
      if (g) std::fill(g, g + n, 0.0);
      for (int f = 0; f < W; f++) {
        double Hh[36], gg[6], cf;
        accumulate_pose(lf[f].data(), n_corner[f], pf[f].data(), n_surf[f], xx + 6 * f, Tbl, prm->lidar_m,
                        prm->plan_weight_tan, huber_a, Hh, gg, &cf, prm->threads);
        c += cf;
        if (H) for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) H[(size_t)(6 * f + i) * n + 6 * f + j] += Hh[6 * i + j];
        if (g) for (int i = 0; i < 6; i++) g[6 * f + i] += gg[i];
      }
      for (int f = 1; f < W; f++) {
        double r[15], J[450];
        const double* pri = xx + 6 * (f - 1); const double* prj = xx + 6 * f;
        const double* vbi = xx + 6 * W + 9 * (f - 1); const double* vbj = xx + 6 * W + 9 * f;
        imu_factor(*(const Preint*)preints[f], gravity3, pri, vbi, prj, vbj, r, (H || g) ? J : nullptr);
        for (int k = 0; k < 15; k++) c += 0.5 * r[k] * r[k];
        if (H || g) {
          const int off[4] = {6 * (f - 1), 6 * W + 9 * (f - 1), 6 * f, 6 * W + 9 * f};
          const int sz[4] = {6, 9, 6, 9}, col0[4] = {0, 6, 15, 21};
          for (int b = 0; b < 4; b++) for (int i = 0; i < sz[b]; i++) {
            const int gi = off[b] + i, ci = col0[b] + i;
            if (g) { double s = 0; for (int k = 0; k < 15; k++) s += J[k * 30 + ci] * r[k]; g[gi] += s; }
            if (H) for (int b2 = 0; b2 < 4; b2++) for (int j = 0; j < sz[b2]; j++) {
              double s = 0;
              for (int k = 0; k < 15; k++) s += J[k * 30 + ci] * J[k * 30 + col0[b2] + j];
              H[(size_t)gi * n + off[b2] + j] += s;
            }
          }
        }
      }
      *cost = c;
      return std::isfinite(c);
    };
    DoglegSummary s = dogleg_minimize(n, x.data(), eval, prm->max_inner);
    inner_total += s.iterations;
    final_cost = s.final_cost;
    // double2vector, EST.cpp:952-964
    for (int f = 0; f < W; f++) {
      double* st = states + 16 * f;
      for (int k = 0; k < 3; k++) st[k] = x[6 * f + k];
      Quat q = so3_exp(&x[6 * f + 3]);
      st[3] = q.w; st[4] = q.x; st[5] = q.y; st[6] = q.z;
      if (W > 1) for (int k = 0; k < 9; k++) st[7 + k] = x[6 * W + 9 * f + k];
    }
    outer_done = it + 1;
    Quat Q = {sb[3], sb[4], sb[5], sb[6]};
    Quat dq = quat_mul(q_before, Quat{Q.w, -Q.x, -Q.y, -Q.z});
    double deltaR = 2.0 * std::atan2(std::sqrt((dq.x * dq.x + dq.y * dq.y) + dq.z * dq.z), std::fabs(dq.w)) * 180.0 / M_PI;
    double dt[3] = {t_before[0] - sb[0], t_before[1] - sb[1], t_before[2] - sb[2]};
    double deltaT = std::sqrt((dt[0] * dt[0] + dt[1] * dt[1]) + dt[2] * dt[2]);
    if ((deltaR < 0.05 && deltaT < 0.05) || (it + 1) == prm->max_outer) break;
  }
  if (stats) {
    stats[0] = outer_done; stats[1] = inner_total; stats[2] = nl; stats[3] = np;
    stats[4] = final_cost; stats[5] = min_sv; stats[6] = is_degenerate;
  }
  return 0;
}

}  // extern "C"
