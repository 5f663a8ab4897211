// ORACLE (test infrastructure only — see oracle.h). Small dense maths restating the
// third-party routines the reference calls (Eigen 3.3, Sophus). Parity unpinned.
#ifndef MMLOAM_ORACLE_MATH_H
#define MMLOAM_ORACLE_MATH_H
#include <algorithm>
#include <cmath>
#include <vector>

namespace orc {

struct Quat { double w, x, y, z; };

inline Quat quat_normalized(Quat q) {
  double n = std::sqrt(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}

// Eigen 3.3 quaternionbase_assign_impl<Matrix3> (row-major R9)
inline Quat quat_from_R(const double* m) {
  auto M = [&](int r, int c) { return m[3 * r + c]; };
  Quat q;
  double t = (M(0, 0) + M(1, 1)) + M(2, 2);
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (M(2, 1) - M(1, 2)) * t;
    q.y = (M(0, 2) - M(2, 0)) * t;
    q.z = (M(1, 0) - M(0, 1)) * t;
  } else {
    int i = 0;
    if (M(1, 1) > M(0, 0)) i = 1;
    if (M(2, 2) > M(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (M(k, j) - M(j, k)) * t;
    v[j] = (M(j, i) + M(i, j)) * t;
    v[k] = (M(k, i) + M(i, k)) * t;
    q.x = v[0]; q.y = v[1]; q.z = v[2];
  }
  return q;
}

// Eigen 3.3 QuaternionBase::toRotationMatrix (row-major out)
inline void quat_to_R(const Quat& q, double* R) {
  double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// Eigen 3.3 QuaternionBase::slerp
inline Quat quat_slerp(const Quat& a, double t, const Quat& b) {
  const double one = 1.0 - 2.220446049250313e-16;
  double d = ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w;
  double absD = std::fabs(d);
  double s0, s1;
  if (absD >= one) {
    s0 = 1.0 - t;
    s1 = t;
  } else {
    double theta = std::acos(absD);
    double sinTheta = std::sin(theta);
    s0 = std::sin((1.0 - t) * theta) / sinTheta;
    s1 = std::sin(t * theta) / sinTheta;
  }
  if (d < 0) s1 = -s1;
  return {s0 * a.w + s1 * b.w, s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z};
}

// Eigen 3.3 QuaternionBase::_transformVector: uv = 2 * (q.vec x v); v + w*uv + q.vec x uv
inline void quat_rotate(const Quat& q, const double* v, double* out) {
  double uv[3] = {q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  double c[3] = {q.y * uv[2] - q.z * uv[1], q.z * uv[0] - q.x * uv[2], q.x * uv[1] - q.y * uv[0]};
  out[0] = v[0] + q.w * uv[0] + c[0];
  out[1] = v[1] + q.w * uv[1] + c[1];
  out[2] = v[2] + q.w * uv[2] + c[2];
}

inline Quat quat_mul(const Quat& a, const Quat& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
          a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}

// Sophus SO3::expAndTheta, so3.hpp:585-623 (epsilon = 1e-10, common.hpp:117)
inline Quat so3_exp(const double* om) {
  double theta_sq = (om[0] * om[0] + om[1] * om[1]) + om[2] * om[2];
  double imag, real;
  if (theta_sq < 1e-10 * 1e-10) {
    double theta_po4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    double theta = std::sqrt(theta_sq);
    double half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  return {real, imag * om[0], imag * om[1], imag * om[2]};
}

// Sophus SO3::logAndTheta, so3.hpp:247-292
inline void so3_log(const Quat& q, double* out) {
  double squared_n = (q.x * q.x + q.y * q.y) + q.z * q.z;
  double w = q.w;
  double f;
  if (squared_n < 1e-10 * 1e-10) {
    double squared_w = w * w;
    f = 2.0 / w - (2.0 / 3.0) * squared_n / (w * squared_w);
  } else {
    double n = std::sqrt(squared_n);
    if (std::fabs(w) < 1e-10) f = (w > 0 ? M_PI : -M_PI) / n;
    else f = 2.0 * std::atan(n / w) / n;
  }
  out[0] = f * q.x; out[1] = f * q.y; out[2] = f * q.z;
}

// Symmetric 3x3 eigen-decomposition, eigenvalues ascending, unit eigenvectors in the
// columns of V (row-major V[r*3+c]). Cyclic Jacobi; stands in for Eigen's
// SelfAdjointEigenSolver<Matrix3d> (tridiagonal QL) — same result to rounding, sign of
// eigenvectors arbitrary in both.
inline void eig3_sym(const double* Ain, double* evals, double* V) {
  double A[3][3] = {{Ain[0], Ain[1], Ain[2]}, {Ain[3], Ain[4], Ain[5]}, {Ain[6], Ain[7], Ain[8]}};
  double U[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 32; sweep++) {
    double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (A[p][q] == 0.0) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; k++) {  // A <- A * G
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {  // A <- G^T * A
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          double ukp = U[k][p], ukq = U[k][q];
          U[k][p] = c * ukp - s * ukq;
          U[k][q] = s * ukp + c * ukq;
        }
      }
  }
  int order[3] = {0, 1, 2};
  double d[3] = {A[0][0], A[1][1], A[2][2]};
  std::sort(order, order + 3, [&](int a, int b) { return d[a] < d[b]; });
  for (int c = 0; c < 3; c++) {
    evals[c] = d[order[c]];
    for (int r = 0; r < 3; r++) V[3 * r + c] = U[r][order[c]];
  }
}

// Least-squares solve of the 5x3 system A x = b by Householder QR with column pivoting
// (Eigen 3.3 ColPivHouseholderQR::solve: pivot = largest remaining column norm, rank by
// |R_kk| > eps * 3 * max|R_kk|, zero for the deficient components).
inline void qr5x3_solve(const double A_in[5][3], const double b_in[5], double x[3]) {
  double A[5][3], b[5];
  for (int i = 0; i < 5; i++) {
    b[i] = b_in[i];
    for (int j = 0; j < 3; j++) A[i][j] = A_in[i][j];
  }
  int perm[3] = {0, 1, 2};
  double maxpivot = 0;
  int rank = 0;
  double Rdiag[3];
  for (int k = 0; k < 3; k++) {
    // pivot: largest remaining column norm (recomputed exactly)
    int best = k;
    double bestn = -1;
    for (int j = k; j < 3; j++) {
      double s = 0;
      for (int i = k; i < 5; i++) s += A[i][j] * A[i][j];
      if (s > bestn) { bestn = s; best = j; }
    }
    if (best != k) {
      for (int i = 0; i < 5; i++) std::swap(A[i][k], A[i][best]);
      std::swap(perm[k], perm[best]);
    }
    // Householder vector for column k, rows k..4
    double tail = 0;
    for (int i = k + 1; i < 5; i++) tail += A[i][k] * A[i][k];
    double c0 = A[k][k];
    double beta, tau;
    double v[5] = {0, 0, 0, 0, 0};
    if (tail <= 1e-300) {
      tau = 0;
      beta = c0;
    } else {
      beta = std::sqrt(c0 * c0 + tail);
      if (c0 >= 0) beta = -beta;
      for (int i = k + 1; i < 5; i++) v[i] = A[i][k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    v[k] = 1.0;
    if (tau != 0) {
      for (int j = k + 1; j < 3; j++) {
        double s = 0;
        for (int i = k; i < 5; i++) s += v[i] * A[i][j];
        s *= tau;
        for (int i = k; i < 5; i++) A[i][j] -= s * v[i];
      }
      double s = 0;
      for (int i = k; i < 5; i++) s += v[i] * b[i];
      s *= tau;
      for (int i = k; i < 5; i++) b[i] -= s * v[i];
    }
    A[k][k] = beta;
    for (int i = k + 1; i < 5; i++) A[i][k] = 0;
    Rdiag[k] = beta;
    maxpivot = std::max(maxpivot, std::fabs(beta));
  }
  double thr = 2.220446049250313e-16 * 3.0 * maxpivot;
  for (int k = 0; k < 3; k++)
    if (std::fabs(Rdiag[k]) > thr) rank++;
  double y[3] = {0, 0, 0};
  for (int k = rank - 1; k >= 0; k--) {
    double s = b[k];
    for (int j = k + 1; j < rank; j++) s -= A[k][j] * y[j];
    y[k] = s / A[k][k];
  }
  for (int k = 0; k < 3; k++) x[perm[k]] = (k < rank) ? y[k] : 0.0;
}

// Dense symmetric positive-definite solve (Cholesky). Returns false when the
// factorisation meets a non-positive pivot.
inline bool chol_solve(int n, const double* A, const double* b, double* x) {
  std::vector<double> Lv((size_t)n * n), yv(n);
  double* L = Lv.data();
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      double s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
      if (i == j) {
        if (!(s > 0)) return false;
        L[i * n + i] = std::sqrt(s);
      } else
        L[i * n + j] = s / L[j * n + j];
    }
  double* y = yv.data();
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * y[k];
    y[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = y[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
  return true;
}

}  // namespace orc
#endif
