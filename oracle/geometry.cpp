// ORACLE (test infrastructure only — see oracle.h).
//   orc_undistort()        <- RemoveLidarDistortion   src/unionPoseEstimation.cpp:402-421
//   orc_point_to_map()     <- pointAssociateToMap     src/lio/Map_Manager.cpp:75-89
//   orc_cube_index()       <- FindUsedCornerMap/Surf  src/lio/Map_Manager.cpp:583-629
//   orc_voxel_downsample() <- pcl::VoxelGrid::filter as called at
//                             src/lio/Estimator.cpp:1015-1024, 1631-1637
//   orc_so3_exp/log        <- include/sophus/so3.hpp:585-623, 247-292
// Third-party arithmetic restated (parity unpinned): Eigen 3.3 Quaternion(Matrix3),
// Quaternion::slerp, Quaternion*Vector3; PCL 1.8 VoxelGrid::applyFilter.
#include "oracle.h"
#include "oracle_math.h"
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

using namespace orc;

extern "C" {

void orc_so3_exp(const double* phi3, double* q_wxyz4, double* R9) {
  Quat q = so3_exp(phi3);
  if (q_wxyz4) { q_wxyz4[0] = q.w; q_wxyz4[1] = q.x; q_wxyz4[2] = q.y; q_wxyz4[3] = q.z; }
  if (R9) quat_to_R(q, R9);
}
void orc_so3_log(const double* q_wxyz4, double* phi3) {
  Quat q = {q_wxyz4[0], q_wxyz4[1], q_wxyz4[2], q_wxyz4[3]};
  so3_log(q, phi3);
}

// PE.cpp:402-421
int orc_undistort(float* xyzi, const float* s_arr, int n, const double* dR, const double* dt) {
  Quat qlc = quat_normalized(quat_from_R(dR));  // PE.cpp:410
  const Quat ident = {1, 0, 0, 0};
  for (int i = 0; i < n; i++) {
    float s = s_arr[i];
    Quat dq = quat_normalized(quat_slerp(ident, (double)s, qlc));  // PE.cpp:411
    double dP[3] = {s * dt[0], s * dt[1], s * dt[2]};              // PE.cpp:412
    double v[3] = {(double)xyzi[4 * i], (double)xyzi[4 * i + 1], (double)xyzi[4 * i + 2]};
    double r[3];
    quat_rotate(dq, v, r);
    double sp[3] = {r[0] + dP[0] - dt[0], r[1] + dP[1] - dt[1], r[2] + dP[2] - dt[2]};  // startP - dtlc
    // dRlc.transpose() * (startP - dtlc)
    double po[3];
    for (int c = 0; c < 3; c++) po[c] = (dR[0 * 3 + c] * sp[0] + dR[1 * 3 + c] * sp[1]) + dR[2 * 3 + c] * sp[2];
    xyzi[4 * i] = (float)po[0];
    xyzi[4 * i + 1] = (float)po[1];
    xyzi[4 * i + 2] = (float)po[2];
  }
  return 0;
}

// MM.cpp:75-89: double transform, stored back to float32
void orc_point_to_map(const float* p3, const double* T, float* out3) {
  double pin[3] = {(double)p3[0], (double)p3[1], (double)p3[2]};
  for (int r = 0; r < 3; r++) {
    double v = ((T[4 * r] * pin[0] + T[4 * r + 1] * pin[1]) + T[4 * r + 2] * pin[2]) + T[4 * r + 3];
    out3[r] = (float)v;
  }
}

// MM.cpp:583-605 (a=cen_w, b=cen_h, c=cen_d); 21 x 11 x 21 cubes (MM.h:117-119)
int orc_cube_index(const float* p, int a, int b, int c) {
  const int W = 21, Hh = 11, D = 21;
  int cubeI = int((p[0] + 25.0) / 50.0) + c;
  int cubeJ = int((p[1] + 25.0) / 50.0) + a;
  int cubeK = int((p[2] + 25.0) / 50.0) + b;
  if (p[0] + 25.0 < 0) cubeI--;
  if (p[1] + 25.0 < 0) cubeJ--;
  if (p[2] + 25.0 < 0) cubeK--;
  if (cubeI >= 0 && cubeI < D && cubeJ >= 0 && cubeJ < W && cubeK >= 0 && cubeK < Hh)
    return cubeI + D * cubeJ + D * W * cubeK;  // MM.cpp:64-66 ToIndex
  return 5000;
}

// PCL 1.8 VoxelGrid<PointT>::applyFilter, downsample_all_data = true,
// min_points_per_voxel = 0, no filter field. Within a voxel PCL's std::sort leaves the
// point order unspecified; the oracle defines it as ascending input index.
int orc_voxel_downsample(const float* xyzi, int n, float leaf, float* out, int* m_out) {
  *m_out = 0;
  if (n <= 0) return 0;
  float inv = 1.0f / leaf;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = 0; i < n; i++) {
    const float* p = xyzi + 4 * i;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    for (int c = 0; c < 3; c++) {
      mn[c] = std::min(mn[c], p[c]);
      mx[c] = std::max(mx[c], p[c]);
    }
  }
  int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1;
  int64_t dy = (int64_t)((mx[1] - mn[1]) * inv) + 1;
  int64_t dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)INT_MAX) {  // PCL: warn and pass the input through
    std::memcpy(out, xyzi, sizeof(float) * 4 * n);
    *m_out = n;
    return 1;
  }
  int min_b[3], max_b[3], div_b[3];
  for (int c = 0; c < 3; c++) {
    min_b[c] = (int)std::floor(mn[c] * inv);
    max_b[c] = (int)std::floor(mx[c] * inv);
    div_b[c] = max_b[c] - min_b[c] + 1;
  }
  int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  std::vector<std::pair<unsigned, int>> iv;
  iv.reserve(n);
  for (int i = 0; i < n; i++) {
    const float* p = xyzi + 4 * i;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    int i0 = (int)(std::floor(p[0] * inv) - (float)min_b[0]);
    int i1 = (int)(std::floor(p[1] * inv) - (float)min_b[1]);
    int i2 = (int)(std::floor(p[2] * inv) - (float)min_b[2]);
    int idx = i0 * mul[0] + i1 * mul[1] + i2 * mul[2];
    iv.emplace_back((unsigned)idx, i);
  }
  std::stable_sort(iv.begin(), iv.end(),
                   [](const std::pair<unsigned, int>& a, const std::pair<unsigned, int>& b) { return a.first < b.first; });
  int m = 0;
  size_t k = 0;
  while (k < iv.size()) {
    size_t e = k + 1;
    while (e < iv.size() && iv[e].first == iv[k].first) e++;
    float sx = 0, sy = 0, sz = 0, si = 0;
    for (size_t j = k; j < e; j++) {
      const float* p = xyzi + 4 * iv[j].second;
      sx += p[0]; sy += p[1]; sz += p[2]; si += p[3];
    }
    float cnt = (float)(e - k);
    out[4 * m] = sx / cnt;
    out[4 * m + 1] = sy / cnt;
    out[4 * m + 2] = sz / cnt;
    out[4 * m + 3] = si / cnt;
    m++;
    k = e;
  }
  *m_out = m;
  return 0;
}

}  // extern "C"
