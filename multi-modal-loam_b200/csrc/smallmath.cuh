// Small dense double-precision routines used inside kernels (one thread each):
//   eig3_sym      symmetric 3x3 eigen-decomposition   (Eigen SelfAdjointEigenSolver<Matrix3d>, EST.cpp:251)
//   qr5x3_solve   5x3 least squares, pivoted QR       (Eigen colPivHouseholderQr().solve, EST.cpp:640)
//   so3_exp/log   rotation vector <-> unit quaternion  (Sophus so3.hpp:585-623, 247-292)
//   chol_solve    SPD solve for the 6W x 6W dogleg system
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace mml {

// eigenvalues ascending in ev[], unit eigenvectors in the columns of V (row-major V[3*r+c]).
// Cyclic Jacobi sweeps: converges to double rounding for 3x3 in <= 6 sweeps.
__host__ __device__ inline void eig3_sym(const double* Ain, double* ev, double* V) {
  double a00 = Ain[0], a01 = Ain[1], a02 = Ain[2], a11 = Ain[4], a12 = Ain[5], a22 = Ain[8];
  double u[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int sweep = 0; sweep < 32; sweep++) {
    const double off = a01 * a01 + a02 * a02 + a12 * a12;
    const double dg = a00 * a00 + a11 * a11 + a22 * a22;
    if (off == 0.0 || off <= 1e-32 * dg) break;
    // rotation (0,1)
    if (a01 != 0.0) {
      const double th = (a11 - a00) / (2.0 * a01);
      const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      const double b00 = a00 - t * a01, b11 = a11 + t * a01;
      const double b02 = c * a02 - s * a12, b12 = s * a02 + c * a12;
      a00 = b00; a11 = b11; a01 = 0.0; a02 = b02; a12 = b12;
      for (int k = 0; k < 3; k++) {
        const double p = u[3 * k], q = u[3 * k + 1];
        u[3 * k] = c * p - s * q;
        u[3 * k + 1] = s * p + c * q;
      }
    }
    // rotation (0,2)
    if (a02 != 0.0) {
      const double th = (a22 - a00) / (2.0 * a02);
      const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      const double b00 = a00 - t * a02, b22 = a22 + t * a02;
      const double b01 = c * a01 - s * a12, b12 = s * a01 + c * a12;
      a00 = b00; a22 = b22; a02 = 0.0; a01 = b01; a12 = b12;
      for (int k = 0; k < 3; k++) {
        const double p = u[3 * k], q = u[3 * k + 2];
        u[3 * k] = c * p - s * q;
        u[3 * k + 2] = s * p + c * q;
      }
    }
    // rotation (1,2)
    if (a12 != 0.0) {
      const double th = (a22 - a11) / (2.0 * a12);
      const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      const double b11 = a11 - t * a12, b22 = a22 + t * a12;
      const double b01 = c * a01 - s * a02, b02 = s * a01 + c * a02;
      a11 = b11; a22 = b22; a12 = 0.0; a01 = b01; a02 = b02;
      for (int k = 0; k < 3; k++) {
        const double p = u[3 * k + 1], q = u[3 * k + 2];
        u[3 * k + 1] = c * p - s * q;
        u[3 * k + 2] = s * p + c * q;
      }
    }
  }
  double d[3] = {a00, a11, a22};
  int o0 = 0, o1 = 1, o2 = 2;
  if (d[o1] < d[o0]) { int t = o0; o0 = o1; o1 = t; }
  if (d[o2] < d[o1]) { int t = o1; o1 = o2; o2 = t; }
  if (d[o1] < d[o0]) { int t = o0; o0 = o1; o1 = t; }
  const int ord[3] = {o0, o1, o2};
  for (int c = 0; c < 3; c++) {
    ev[c] = d[ord[c]];
    for (int r = 0; r < 3; r++) V[3 * r + c] = u[3 * r + ord[c]];
  }
}

// Smallest eigenvalue of a symmetric 3x3 matrix in closed form (trigonometric solution of the characteristic
// polynomial). Absolute error ~1e-15 x the largest eigenvalue: used where the value only meets a coarse threshold
// (checkLocalizability, EST.cpp:536-565: sqrt(lambda_min) against 2.0 / 3.0), never where the oracle's digits matter.
__host__ __device__ inline double eig3_sym_min(const double* A) {
  const double a00 = A[0], a01 = A[1], a02 = A[2], a11 = A[4], a12 = A[5], a22 = A[8];
  const double p1 = a01 * a01 + a02 * a02 + a12 * a12;
  const double q = (a00 + a11 + a22) / 3.0;
  const double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
  const double p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * p1;
  if (!(p2 > 0.0)) return q;  // multiple of the identity
  const double p = sqrt(p2 / 6.0);
  const double ip = 1.0 / p;
  const double c00 = b00 * ip, c11 = b11 * ip, c22 = b22 * ip, c01 = a01 * ip, c02 = a02 * ip, c12 = a12 * ip;
  double r = 0.5 * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) + c02 * (c01 * c12 - c11 * c02));
  r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
  const double phi = acos(r) / 3.0;
  return q + 2.0 * p * cos(phi + 2.0943951023931953);  // + 2 pi / 3: the smallest root
}

// least-squares solution of A x = b (5x3), Householder QR with column pivoting.
// Every loop is unrolled over compile-time indices and the column pivoting is done with conditional swaps, so the
// 5x3 matrix lives in registers (a run-time column index would put it in local memory); the arithmetic and its order
// are those of the plain triple loop.
__host__ __device__ inline void qr5x3_solve(double A[5][3], double b[5], double x[3]) {
  int perm0 = 0, perm1 = 1, perm2 = 2;
  double rdiag[3];
  double maxpivot = 0.0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    int best = k;
    double bestn = -1.0;
#pragma unroll
    for (int j = k; j < 3; j++) {
      double s = 0.0;
#pragma unroll
      for (int i = k; i < 5; i++) s += A[i][j] * A[i][j];
      if (s > bestn) { bestn = s; best = j; }
    }
#pragma unroll
    for (int j = k + 1; j < 3; j++) {
      if (best == j) {
#pragma unroll
        for (int i = 0; i < 5; i++) { const double t = A[i][k]; A[i][k] = A[i][j]; A[i][j] = t; }
        // perm[k] <-> perm[j]
        int& pk = k == 0 ? perm0 : (k == 1 ? perm1 : perm2);
        int& pj = j == 1 ? perm1 : perm2;
        const int t = pk; pk = pj; pj = t;
      }
    }
    double tail = 0.0;
#pragma unroll
    for (int i = k + 1; i < 5; i++) tail += A[i][k] * A[i][k];
    const double c0 = A[k][k];
    double beta, tau;
    double v[5] = {0, 0, 0, 0, 0};
    if (tail <= 1e-300) {
      tau = 0.0;
      beta = c0;
    } else {
      beta = sqrt(c0 * c0 + tail);
      if (c0 >= 0) beta = -beta;
#pragma unroll
      for (int i = k + 1; i < 5; i++) v[i] = A[i][k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    v[k] = 1.0;
    if (tau != 0.0) {
#pragma unroll
      for (int j = k + 1; j < 3; j++) {
        double s = 0.0;
#pragma unroll
        for (int i = k; i < 5; i++) s += v[i] * A[i][j];
        s *= tau;
#pragma unroll
        for (int i = k; i < 5; i++) A[i][j] -= s * v[i];
      }
      double s = 0.0;
#pragma unroll
      for (int i = k; i < 5; i++) s += v[i] * b[i];
      s *= tau;
#pragma unroll
      for (int i = k; i < 5; i++) b[i] -= s * v[i];
    }
    A[k][k] = beta;
    rdiag[k] = beta;
    maxpivot = fmax(maxpivot, fabs(beta));
  }
  const double thr = 2.220446049250313e-16 * 3.0 * maxpivot;
  int rank = 0;
#pragma unroll
  for (int k = 0; k < 3; k++)
    if (fabs(rdiag[k]) > thr) rank++;
  double y[3] = {0, 0, 0};
#pragma unroll
  for (int k = 2; k >= 0; k--) {
    if (k < rank) {
      double s = b[k];
#pragma unroll
      for (int j = k + 1; j < 3; j++)
        if (j < rank) s -= A[k][j] * y[j];
      y[k] = s / A[k][k];
    }
  }
  const double y0 = rank > 0 ? y[0] : 0.0, y1 = rank > 1 ? y[1] : 0.0, y2 = rank > 2 ? y[2] : 0.0;
  x[0] = perm0 == 0 ? y0 : (perm1 == 0 ? y1 : y2);
  x[1] = perm0 == 1 ? y0 : (perm1 == 1 ? y1 : y2);
  x[2] = perm0 == 2 ? y0 : (perm1 == 2 ? y1 : y2);
}

struct Quat { double w, x, y, z; };

__host__ __device__ inline Quat so3_exp(const double* om) {
  const double th2 = (om[0] * om[0] + om[1] * om[1]) + om[2] * om[2];
  double imag, real;
  if (th2 < 1e-10 * 1e-10) {
    const double th4 = th2 * th2;
    imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
    real = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
  } else {
    const double th = sqrt(th2);
    const double half = 0.5 * th;
    imag = sin(half) / th;
    real = cos(half);
  }
  return {real, imag * om[0], imag * om[1], imag * om[2]};
}

__host__ __device__ inline void so3_log(const Quat& q, double* out) {
  const double n2 = (q.x * q.x + q.y * q.y) + q.z * q.z;
  const double w = q.w;
  double f;
  if (n2 < 1e-10 * 1e-10) {
    f = 2.0 / w - (2.0 / 3.0) * n2 / (w * (w * w));
  } else {
    const double n = sqrt(n2);
    if (fabs(w) < 1e-10) f = (w > 0 ? 3.14159265358979323846 : -3.14159265358979323846) / n;
    else f = 2.0 * atan(n / w) / n;
  }
  out[0] = f * q.x; out[1] = f * q.y; out[2] = f * q.z;
}

__host__ __device__ inline void quat_to_R(const Quat& q, double* R) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

__host__ __device__ inline Quat quat_mul(const Quat& a, const Quat& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}

__host__ __device__ inline Quat quat_from_R9(const double* m) {
  Quat q;
  double t = (m[0] + m[4]) + m[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m[7] - m[5]) * t;
    q.y = (m[2] - m[6]) * t;
    q.z = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m[3 * k + j] - m[3 * j + k]) * t;
    v[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    v[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    q.x = v[0]; q.y = v[1]; q.z = v[2];
  }
  const double n = sqrt(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
  q.w /= n; q.x /= n; q.y /= n; q.z /= n;
  return q;
}

// SPD solve A x = b, n <= 6, by L D L^T (the pivots d_j are the squares of Cholesky's diagonal, so the
// positive-definiteness test is the same). One reciprocal per column and no square roots: the dependent chain
// is what a single solver thread waits on. false if a pivot is not positive.
template <int N>
__host__ __device__ inline bool chol_solve(const double* A, const double* b, double* x) {
  double L[N * N], inv[N], piv[N];
#pragma unroll
  for (int j = 0; j < N; j++) {
    double w[N];  // w_k = L_jk d_k
    double d = A[j * N + j];
#pragma unroll
    for (int k = 0; k < j; k++) {
      w[k] = L[j * N + k] * piv[k];
      d -= L[j * N + k] * w[k];
    }
    if (!(d > 0)) return false;
    piv[j] = d;
    inv[j] = 1.0 / d;
#pragma unroll
    for (int i = j + 1; i < N; i++) {
      double s = A[i * N + j];
#pragma unroll
      for (int k = 0; k < j; k++) s -= L[i * N + k] * w[k];
      L[i * N + j] = s * inv[j];
    }
  }
  double z[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) s -= L[i * N + k] * z[k];
    z[i] = s;
  }
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    double s = z[i] * inv[i];
#pragma unroll
    for (int k = i + 1; k < N; k++) s -= L[k * N + i] * x[k];
    x[i] = s;
  }
  return true;
}

}  // namespace mml
