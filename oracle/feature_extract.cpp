// ORACLE (test infrastructure only — see oracle.h). CPU restatement of the
// reference's per-line feature detector and its line-splitting glue.
//
// Follows /root/reference/mm-loam/src/unionFeatureExtract.cpp:
//   detect_line()        <- detectFeaturePoints            FE.cpp:341-844
//   orc_velo_ring_time() <- getVeloFeature ring/time       FE.cpp:1136-1195
//   orc_hori_filter()    <- getHoriFeatureExtract filter   FE.cpp:985-998
//   orc_extract_scan()   <- split / label glue             FE.cpp:1001-1023, 1209-1240
//
// Defined readings of the reference's undefined behaviour:
//   * cloudAngle[] is zero-initialised per call (the reference reads stale stack).
//   * work arrays are sized to n (the reference overflows at 20000).
//   * inputs must be finite; non-finite points are compacted away and indices
//     are reported in the compacted numbering exactly as the reference does.
// Arithmetic: float32 where the reference uses float (including the float32 differences it
// feeds into Eigen::Vector3d constructors), float64 where it operates on Eigen::Vector3d; no FMA contraction (build with -ffp-contract=off, matching
// the reference's baseline x86-64 Release build, mm-loam/CMakeLists.txt:4-5).
// Eigen 3.3 reduces a fixed 3-vector dot/squaredNorm as (e0+e1)+e2.
#include "oracle.h"
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct P4 { float x, y, z, i; };

struct V3 {
  double x, y, z;
};
// Vector3d - Vector3d of points widened first (FE.cpp:417-422): float64 subtraction.
inline V3 sub(const P4& a, const P4& b) {
  return {(double)a.x - (double)b.x, (double)a.y - (double)b.y, (double)a.z - (double)b.z};
}
// Eigen::Vector3d(a.x - b.x, a.y - b.y, a.z - b.z) (FE.cpp:618-620, 625-627, 635-640, 680-682,
// 717-719, 772-774, 790-792): the subtraction is float32, the result is widened afterwards.
inline V3 subf(const P4& a, const P4& b) {
  return {(double)(a.x - b.x), (double)(a.y - b.y), (double)(a.z - b.z)};
}
inline double dot(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline double norm(const V3& a) { return std::sqrt(dot(a, a)); }
// Eigen 3.3 MatrixBase::normalize(): z = squaredNorm(); if (z > 0) *this /= sqrt(z)
inline void normalize(V3& a) {
  double z = dot(a, a);
  if (z > 0) {
    double s = std::sqrt(z);
    a.x /= s; a.y /= s; a.z /= s;
  }
}
inline float range3(const P4& p) { return std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z); }

// FE.cpp:341-844. `flags_out` (optional) receives the final CloudFeatureFlag[].
void detect_line(const P4* in, int n_in, std::vector<int>& sharp, std::vector<int>& flat,
                 std::vector<int>* flags_out) {
  // FE.cpp:353-359 constants
  int thNumCurvSize = 2;
  const float thDistanceFaraway = 50.0f;
  const int thNumFlat = 1;
  const int thPartNum = 50;
  const float thFlatThreshold = 0.02f;
  const float thLidarNearestDis = 1.0f;
  const float thBreakCornerDis = 1.0f;

  // FE.cpp:369-390 compaction of non-finite points
  std::vector<P4> pts;
  pts.reserve(n_in);
  for (int i = 0; i < n_in; i++) {
    if (!std::isfinite(in[i].x) || !std::isfinite(in[i].y) || !std::isfinite(in[i].z)) continue;
    pts.push_back(in[i]);
  }
  const int n = (int)pts.size();
  const P4* p = pts.data();

  std::vector<int> flag(n > 0 ? n : 1, 0);
  std::vector<float> curv(n > 0 ? n : 1, 0.f), depth(n > 0 ? n : 1, 0.f), refl(n > 0 ? n : 1, 0.f);
  std::vector<int> sortInd(n > 0 ? n : 1, 0), reflInd(n > 0 ? n : 1, 0), angle(n > 0 ? n : 1, 0);

  int count_num = 1;
  bool left_surf_flag = false, right_surf_flag = false;
  const int scanStartInd = 5;
  const int scanEndInd = n - 6;

  // FE.cpp:407-451 curvature, depth, incidence angles, reflectivity difference
  for (int i = 5; i < n - 5; i++) {
    float diffX = 0, diffY = 0, diffZ = 0;
    float dis = range3(p[i]);
    V3 cur = {(double)p[i].x, (double)p[i].y, (double)p[i].z};
    V3 dl = sub(p[i - 1], p[i]);
    V3 dn = sub(p[i + 1], p[i]);
    double angle_last = dot(dl, cur) / (norm(dl) * norm(cur));
    double angle_next = dot(dn, cur) / (norm(dn) * norm(cur));
    if (dis > thDistanceFaraway || (std::fabs(angle_last) > 0.966 && std::fabs(angle_next) > 0.966))
      thNumCurvSize = 2;
    else
      thNumCurvSize = 3;
    if (std::fabs(angle_last) > 0.966 && std::fabs(angle_next) > 0.966) angle[i] = 1;

    float diffR = -2 * thNumCurvSize * p[i].i;
    for (int j = 1; j <= thNumCurvSize; ++j) {
      diffX += p[i - j].x + p[i + j].x;
      diffY += p[i - j].y + p[i + j].y;
      diffZ += p[i - j].z + p[i + j].z;
      diffR += p[i - j].i + p[i + j].i;
    }
    diffX -= 2 * thNumCurvSize * p[i].x;
    diffY -= 2 * thNumCurvSize * p[i].y;
    diffZ -= 2 * thNumCurvSize * p[i].z;

    depth[i] = dis;
    curv[i] = diffX * diffX + diffY * diffY + diffZ * diffZ;
    sortInd[i] = i;
    refl[i] = diffR;
    reflInd[i] = i;
  }

  // FE.cpp:453-541 per-part sorts and flat selection
  for (int j = 0; j < thPartNum; j++) {
    int sp = scanStartInd + (scanEndInd - scanStartInd) * j / thPartNum;
    int ep = scanStartInd + (scanEndInd - scanStartInd) * (j + 1) / thPartNum - 1;

    // FE.cpp:458-479: insertion sorts, ascending, stable
    for (int k = sp + 1; k <= ep; k++)
      for (int l = k; l >= sp + 1; l--)
        if (curv[sortInd[l]] < curv[sortInd[l - 1]]) std::swap(sortInd[l - 1], sortInd[l]);
    for (int k = sp + 1; k <= ep; k++)
      for (int l = k; l >= sp + 1; l--)
        if (refl[reflInd[l]] < refl[reflInd[l - 1]]) std::swap(reflInd[l - 1], reflInd[l]);

    int smallestPickedNum = 1;
    int sharpestPickedNum = 1;
    // FE.cpp:483-519
    for (int k = sp; k <= ep; k++) {
      int ind = sortInd[k];
      if (flag[ind] != 0) continue;
      if (curv[ind] < thFlatThreshold * depth[ind] * thFlatThreshold * depth[ind]) {
        flag[ind] = 3;
        for (int l = 1; l <= thNumCurvSize; l++) {
          float dX = p[ind + l].x - p[ind + l - 1].x;
          float dY = p[ind + l].y - p[ind + l - 1].y;
          float dZ = p[ind + l].z - p[ind + l - 1].z;
          if (dX * dX + dY * dY + dZ * dZ > 0.02 || depth[ind] > thDistanceFaraway) break;
          flag[ind + l] = 1;
        }
        for (int l = -1; l >= -thNumCurvSize; l--) {
          float dX = p[ind + l].x - p[ind + l + 1].x;
          float dY = p[ind + l].y - p[ind + l + 1].y;
          float dZ = p[ind + l].z - p[ind + l + 1].z;
          if (dX * dX + dY * dY + dZ * dZ > 0.02 || depth[ind] > thDistanceFaraway) break;
          flag[ind + l] = 1;
        }
      }
    }
    // FE.cpp:521-539
    for (int k = sp; k <= ep; k++) {
      int ind = sortInd[k];
      if (((flag[ind] == 3) && (smallestPickedNum <= thNumFlat)) ||
          ((flag[ind] == 3) && (depth[ind] > thDistanceFaraway)) || angle[ind] == 1) {
        smallestPickedNum++;
        flag[ind] = 2;
      }
      int idx = reflInd[k];
      if (curv[idx] < 0.7 * thFlatThreshold * depth[idx] * thFlatThreshold * depth[idx] &&
          sharpestPickedNum <= 3 && refl[idx] > 20.0) {
        sharpestPickedNum++;
        flag[idx] = 300;
      }
    }
  }

  // FE.cpp:543-650 two-plane corner (flag 150), stride count_num
  for (int i = 5; i < n - 5; i += count_num) {
    float dep = range3(p[i]);
    float lX = p[i - 4].x + p[i - 3].x - 4 * p[i - 2].x + p[i - 1].x + p[i].x;
    float lY = p[i - 4].y + p[i - 3].y - 4 * p[i - 2].y + p[i - 1].y + p[i].y;
    float lZ = p[i - 4].z + p[i - 3].z - 4 * p[i - 2].z + p[i - 1].z + p[i].z;
    float left_curvature = lX * lX + lY * lY + lZ * lZ;
    left_surf_flag = left_curvature < thFlatThreshold * dep;

    float rX = p[i + 4].x + p[i + 3].x - 4 * p[i + 2].x + p[i + 1].x + p[i].x;
    float rY = p[i + 4].y + p[i + 3].y - 4 * p[i + 2].y + p[i + 1].y + p[i].y;
    float rZ = p[i + 4].z + p[i + 3].z - 4 * p[i + 2].z + p[i + 1].z + p[i].z;
    float right_curvature = rX * rX + rY * rY + rZ * rZ;
    if (right_curvature < thFlatThreshold * dep) {
      count_num = 4;
      right_surf_flag = true;
    } else {
      count_num = 1;
      right_surf_flag = false;
    }

    if (left_surf_flag && right_surf_flag) {
      V3 nl = {0, 0, 0}, nr = {0, 0, 0};
      for (int k = 1; k < 5; k++) {
        V3 t = subf(p[i - k], p[i]);
        normalize(t);
        double w = k / 10.0;
        nl.x += w * t.x; nl.y += w * t.y; nl.z += w * t.z;
      }
      for (int k = 1; k < 5; k++) {
        V3 t = subf(p[i + k], p[i]);
        normalize(t);
        double w = k / 10.0;
        nr.x += w * t.x; nr.y += w * t.y; nr.z += w * t.z;
      }
      double cc = std::fabs(dot(nl, nr) / (norm(nl) * norm(nr)));
      double last_dis = norm(subf(p[i - 4], p[i]));
      double current_dis = norm(subf(p[i + 4], p[i]));
      if (cc < 0.5 && last_dis > 0.05 && current_dis > 0.05) flag[i] = 150;
    }
  }

  // FE.cpp:651-806 break points (flag 100 / 101)
  for (int i = 5; i < n - 5; i++) {
    float diff_left[2], diff_right[2];
    for (int c = 1; c < 3; c++) {
      float dX1 = p[i + c].x - p[i].x, dY1 = p[i + c].y - p[i].y, dZ1 = p[i + c].z - p[i].z;
      diff_right[c - 1] = std::sqrt(dX1 * dX1 + dY1 * dY1 + dZ1 * dZ1);
      float dX2 = p[i - c].x - p[i].x, dY2 = p[i - c].y - p[i].y, dZ2 = p[i - c].z - p[i].z;
      diff_left[c - 1] = std::sqrt(dX2 * dX2 + dY2 * dY2 + dZ2 * dZ2);
    }
    float depth_right = range3(p[i + 1]);
    float depth_left = range3(p[i - 1]);

    if (std::fabs(diff_right[0] - diff_left[0]) > thBreakCornerDis) {
      V3 lidar_vector = {(double)p[i].x, (double)p[i].y, (double)p[i].z};
      if (diff_right[0] > diff_left[0]) {
        V3 surf_vector = subf(p[i - 1], p[i]);
        double cc = std::fabs(dot(surf_vector, lidar_vector) / (norm(surf_vector) * norm(lidar_vector)));
        if (cc < 0.95) {
          if (depth_right > depth_left) flag[i] = 100;
          else if (depth_right == 0) flag[i] = 100;
        }
      } else {
        V3 surf_vector = subf(p[i + 1], p[i]);
        double cc = std::fabs(dot(surf_vector, lidar_vector) / (norm(surf_vector) * norm(lidar_vector)));
        if (cc < 0.95) {
          if (depth_right < depth_left) flag[i] = 100;
          else if (depth_left == 0) flag[i] = 100;
        }
      }
    }

    // FE.cpp:756-804 direction-consistency test
    if (flag[i] == 100) {
      V3 nf = {0, 0, 0}, nb = {0, 0, 0};
      for (int k = 1; k < 4; k++) {
        float temp_depth = range3(p[i - k]);
        if (temp_depth < 1) continue;
        V3 t = subf(p[i - k], p[i]);
        normalize(t);
        double w = k / 6.0;
        nf.x += w * t.x; nf.y += w * t.y; nf.z += w * t.z;
      }
      for (int k = 1; k < 4; k++) {
        float temp_depth = range3(p[i - k]);  // sic: the reference tests i-k here too (FE.cpp:782)
        if (temp_depth < 1) continue;
        V3 t = subf(p[i + k], p[i]);
        normalize(t);
        double w = k / 6.0;
        nb.x += w * t.x; nb.y += w * t.y; nb.z += w * t.z;
      }
      double cc = std::fabs(dot(nf, nb) / (norm(nf) * norm(nb)));
      if (!(cc < 0.95)) flag[i] = 101;
    }
  }

  // FE.cpp:818-842 collection
  for (int i = 5; i < n - 5; i++) {
    float dis = p[i].x * p[i].x + p[i].y * p[i].y + p[i].z * p[i].z;
    if (dis < thLidarNearestDis * thLidarNearestDis) continue;
    if (flag[i] == 2) {
      flat.push_back(i);
      continue;
    }
    if (flag[i] == 100 || flag[i] == 150) sharp.push_back(i);
  }
  if (flags_out) {
    flags_out->assign(flag.begin(), flag.begin() + (n > 0 ? n : 0));
  }
}

}  // namespace

extern "C" {

int orc_detect_feature_points(const float* xyzi, int n, int* sharp, int* n_sharp, int* flat,
                              int* n_flat) {
  std::vector<int> s, f;
  detect_line(reinterpret_cast<const P4*>(xyzi), n, s, f, nullptr);
  std::memcpy(sharp, s.data(), s.size() * sizeof(int));
  std::memcpy(flat, f.data(), f.size() * sizeof(int));
  *n_sharp = (int)s.size();
  *n_flat = (int)f.size();
  return 0;
}

int orc_detect_feature_flags(const float* xyzi, int n, int* flags) {
  std::vector<int> s, f, fl;
  detect_line(reinterpret_cast<const P4*>(xyzi), n, s, f, &fl);
  std::memcpy(flags, fl.data(), fl.size() * sizeof(int));
  return (int)fl.size();
}

// FE.cpp:1136-1195. `atan`/`atan2` are taken as the double libm functions on the
// float arguments, results stored to float as the reference does.
int orc_velo_ring_time(const float* xyzi, int n, int16_t* line_out, float* reltime_out) {
  const P4* p = reinterpret_cast<const P4*>(xyzi);
  if (n <= 0) return 0;
  float startOri = -std::atan2((double)p[0].y, (double)p[0].x);
  float endOri = -std::atan2((double)p[n - 1].y, (double)p[n - 1].x) + 2 * M_PI;
  if (endOri - startOri > 3 * M_PI) endOri -= 2 * M_PI;
  else if (endOri - startOri < M_PI) endOri += 2 * M_PI;

  bool halfPassed = false;
  int kept = 0;
  for (int i = 0; i < n; i++) {
    float px = p[i].x, py = p[i].y, pz = p[i].z;
    float angle = std::atan((double)(pz / std::sqrt(px * px + py * py))) * 180 / M_PI;
    int scanID = int((angle + 15) / 2 + 0.5);
    if (scanID > 15 || scanID < 0) {
      line_out[i] = -1;
      reltime_out[i] = 0.f;
      continue;
    }
    float ori = -std::atan2((double)py, (double)px);
    if (!halfPassed) {
      if (ori < startOri - M_PI / 2) ori += 2 * M_PI;
      else if (ori > startOri + M_PI * 3 / 2) ori -= 2 * M_PI;
      if (ori - startOri > M_PI) halfPassed = true;
    } else {
      ori += 2 * M_PI;
      if (ori < endOri - M_PI * 3 / 2) ori += 2 * M_PI;
      else if (ori > endOri + M_PI / 2) ori -= 2 * M_PI;
    }
    float relTime = (ori - startOri) / (endOri - startOri);
    line_out[i] = (int16_t)scanID;
    reltime_out[i] = relTime;
    kept++;
  }
  return kept;
}

// FE.cpp:985-998. ros::Time().fromNSec(t).toSec() == sec + 1e-9 * nsec in double.
int orc_hori_filter(const uint32_t* offset_time, const float* xyz3, const uint8_t* line, int n,
                    uint8_t* keep, float* reltime_out) {
  if (n <= 0) return 0;
  auto to_sec = [](uint32_t t) {
    uint32_t sec = t / 1000000000u, nsec = t % 1000000000u;
    return (double)sec + 1e-9 * (double)nsec;
  };
  double timeSpan = to_sec(offset_time[n - 1]);
  int kept = 0;
  for (int i = 0; i < n; i++) {
    int line_num = (int)line[i];
    keep[i] = 0;
    reltime_out[i] = 0.f;
    if (line_num > 5) continue;
    if (xyz3[3 * i] < 0.01) continue;
    keep[i] = 1;
    reltime_out[i] = (float)(to_sec(offset_time[i]) / timeSpan);
    kept++;
  }
  return kept;
}

// FE.cpp:1001-1023 (Horizon, 6 std::threads) and FE.cpp:1209-1240 (Velodyne, serial).
int orc_extract_scan(const float* xyzi, const uint16_t* line_id, int n, int n_lines,
                     uint8_t* label_out, int threads) {
  const P4* p = reinterpret_cast<const P4*>(xyzi);
  std::vector<std::vector<P4>> vlines(n_lines);
  std::vector<std::vector<int>> vsrc(n_lines);
  for (int i = 0; i < n; i++) {
    int l = line_id[i];
    if (l < 0 || l >= n_lines) continue;
    vlines[l].push_back(p[i]);
    vsrc[l].push_back(i);
  }
  std::vector<std::vector<int>> vcorner(n_lines), vsurf(n_lines);
  auto work = [&](int l) { detect_line(vlines[l].data(), (int)vlines[l].size(), vcorner[l], vsurf[l], nullptr); };
  if (threads <= 1) {
    for (int l = 0; l < n_lines; l++) work(l);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
      th.emplace_back([&, t]() {
        for (int l = t; l < n_lines; l += threads) work(l);
      });
    for (auto& t : th) t.join();
  }
  std::memset(label_out, 0, n);
  for (int l = 0; l < n_lines; l++) {
    for (int j : vcorner[l]) label_out[vsrc[l][j]] = 1;
    for (int j : vsurf[l]) label_out[vsrc[l][j]] = 2;
  }
  return 0;
}

}  // extern "C"
