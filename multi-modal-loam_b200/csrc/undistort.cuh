// RemoveLidarDistortion for one point (src/unionPoseEstimation.cpp:402-421), shared by the
// whole-cloud kernel (geometry.cu) and the fused split + voxel kernel (splitvoxel.cu).
// Restates Eigen 3.3 Quaternion(Matrix3).normalized(), Identity().slerp(s, q).normalized() and
// quaternion * vector in float64, operand order as in Eigen; -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace mml {

struct UndistortParams {
  double qw, qx, qy, qz;    // normalised quaternion of dRlc
  double theta, sin_theta;  // slerp angle between identity and q
  int lerp;                 // |d| >= 1 - eps: linear weights
  int neg;                  // d < 0
  int enabled;              // 0: pass points through unchanged
  int pad;
  double R[9];              // dRlc row-major
  double t[3];
};

__device__ __forceinline__ float4 undistort_point(float4 p, double s, const UndistortParams& P) {
  double s0, s1;
  if (P.lerp) {
    s0 = 1.0 - s;
    s1 = s;
  } else {
    s0 = sin((1.0 - s) * P.theta) / P.sin_theta;
    s1 = sin(s * P.theta) / P.sin_theta;
  }
  if (P.neg) s1 = -s1;
  // slerp(identity, q): identity = (w=1, 0,0,0)
  double w = s0 * 1.0 + s1 * P.qw, x = s0 * 0.0 + s1 * P.qx, y = s0 * 0.0 + s1 * P.qy, z = s0 * 0.0 + s1 * P.qz;
  const double nn = sqrt(((x * x + y * y) + z * z) + w * w);
  w /= nn; x /= nn; y /= nn; z /= nn;
  const double v0 = (double)p.x, v1 = (double)p.y, v2 = (double)p.z;
  double u0 = y * v2 - z * v1, u1 = z * v0 - x * v2, u2 = x * v1 - y * v0;
  u0 += u0; u1 += u1; u2 += u2;
  const double c0 = y * u2 - z * u1, c1 = z * u0 - x * u2, c2 = x * u1 - y * u0;
  const double r0 = v0 + w * u0 + c0, r1 = v1 + w * u1 + c1, r2 = v2 + w * u2 + c2;
  const double a0 = r0 + s * P.t[0] - P.t[0], a1 = r1 + s * P.t[1] - P.t[1], a2 = r2 + s * P.t[2] - P.t[2];
  p.x = (float)((P.R[0] * a0 + P.R[3] * a1) + P.R[6] * a2);
  p.y = (float)((P.R[1] * a0 + P.R[4] * a1) + P.R[7] * a2);
  p.z = (float)((P.R[2] * a0 + P.R[5] * a1) + P.R[8] * a2);
  return p;
}

// quaternion of dRlc and the slerp constants (host: libm, like the reference's host code; device: chained loop)
__host__ __device__ inline UndistortParams make_undistort_params(const double* dR9, const double* dt3) {
  UndistortParams P;
  memset(&P, 0, sizeof(P));
  if (!dR9 || !dt3) return P;
  P.enabled = 1;
  auto M = [&](int r, int c) { return dR9[3 * r + c]; };
  double q[4];
  double t = (M(0, 0) + M(1, 1)) + M(2, 2);
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (M(2, 1) - M(1, 2)) * t;
    q[2] = (M(0, 2) - M(2, 0)) * t;
    q[3] = (M(1, 0) - M(0, 1)) * t;
  } else {
    int i = 0;
    if (M(1, 1) > M(0, 0)) i = 1;
    if (M(2, 2) > M(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (M(k, j) - M(j, k)) * t;
    v[j] = (M(j, i) + M(i, j)) * t;
    v[k] = (M(k, i) + M(i, k)) * t;
    q[1] = v[0]; q[2] = v[1]; q[3] = v[2];
  }
  const double nn = sqrt(((q[1] * q[1] + q[2] * q[2]) + q[3] * q[3]) + q[0] * q[0]);
  P.qw = q[0] / nn; P.qx = q[1] / nn; P.qy = q[2] / nn; P.qz = q[3] / nn;
  const double d = P.qw;  // <identity, q>
  const double absD = fabs(d);
  const double one = 1.0 - 2.220446049250313e-16;
  P.lerp = absD >= one;
  P.neg = d < 0;
  P.theta = P.lerp ? 0.0 : acos(absD);
  P.sin_theta = P.lerp ? 1.0 : sin(P.theta);
  for (int i = 0; i < 9; i++) P.R[i] = dR9[i];
  for (int i = 0; i < 3; i++) P.t[i] = dt3[i];
  return P;
}

}  // namespace mml
