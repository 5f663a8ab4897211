// Header-only C++ adapters: the reference's call surface on top of the C-ABI (include/mmloam_b200.h).
//
// The reference (TIERS/multi-modal-loam, ROS 1 + PCL + Eigen + Ceres) calls the hot path through
// ordinary member functions. These adapters keep those names and argument meanings so that
// unionFeatureExtract.cpp / Estimator.cpp need a one-line change each:
//
//   feature_extraction::detectFeaturePoints(cloud, pointsLessSharp, pointsLessFlat)   FE.cpp:341-343
//   LidarFeatureExtractor::detectFeaturePoint(...)      (LIO-Livox spelling of the same method)
//   Estimator::processPointToLine(...)                                              EST.h:159-165
//   Estimator::processPointToPlanVec(..., bool& is_degenerate)                      EST.h:179-186
//   Estimator::EstimateLidarPose(...) for a one-frame list                          EST.h:211-214
//
// Point clouds are any contiguous container of a point type with float members
// x, y, z, intensity (pcl::PointXYZINormal qualifies; mmloam::PointXYZINormal below is a
// layout-compatible stand-in for builds without PCL). Conversion AoS (48 B) <-> float4 happens here.
// No CPU fallback: without a CUDA device the constructors throw.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/mmloam_b200.h"

namespace mmloam {

// 48-byte stand-in with the field layout of pcl::PointXYZINormal
struct alignas(16) PointXYZINormal {
  float x, y, z, _pad0;
  float normal_x, normal_y, normal_z, _pad1;
  float intensity, curvature, _pad2, _pad3;
};
static_assert(sizeof(PointXYZINormal) == 48, "layout must match pcl::PointXYZINormal");

class Context {
 public:
  explicit Context(int device = 0) {
    int rc = mml_ctx_create(device, 1, &h_);
    if (rc != MML_OK) throw std::runtime_error("mmloam_b200: no usable CUDA device (no CPU fallback), rc=" + std::to_string(rc));
  }
  ~Context() { mml_ctx_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  mml_ctx* get() const { return h_; }
  void check(int rc) const {
    if (rc != MML_OK) throw std::runtime_error(std::string("mmloam_b200: ") + mml_last_error(h_));
  }

 private:
  mml_ctx* h_ = nullptr;
};

template <class Cloud>
inline std::vector<float> to_xyzi(const Cloud& pts) {
  std::vector<float> out(4 * pts.size());
  for (size_t i = 0; i < pts.size(); i++) {
    out[4 * i] = pts[i].x; out[4 * i + 1] = pts[i].y; out[4 * i + 2] = pts[i].z; out[4 * i + 3] = pts[i].intensity;
  }
  return out;
}

// ------------------------------------------------------------------------------------------
class feature_extraction {
 public:
  explicit feature_extraction(Context& ctx) : ctx_(ctx) {}

  // FE.cpp:341-343: one scan line in; line-local indices appended to the two output vectors.
  template <class Cloud>
  void detectFeaturePoints(const Cloud& cloud, std::vector<int>& pointsLessSharp, std::vector<int>& pointsLessFlat) {
    const int n = (int)cloud.size();
    if (n == 0) return;
    std::vector<float> xyzi = to_xyzi(cloud);
    std::vector<uint16_t> line(n, 0);
    std::vector<uint8_t> label(n, 0);
    int ns = 0, nf = 0;
    ctx_.check(mml_extract_features(ctx_.get(), xyzi.data(), line.data(), n, 1, label.data(), &ns, &nf));
    for (int i = 0; i < n; i++) {
      if (label[i] == 1) pointsLessSharp.push_back(i);
      else if (label[i] == 2) pointsLessFlat.push_back(i);
    }
  }

  // All lines of a scan in one call (replaces the 6 std::threads of FE.cpp:1008-1015 and the serial
  // loop of FE.cpp:1228-1230): writes the reference's normal_z labels (0 / 1.0 / 2.0) in place.
  // The line id is read from normal_y as the reference stores it (FE.cpp:996, 1188).
  template <class Cloud>
  void labelScan(Cloud& cloud, int n_lines) {
    const int n = (int)cloud.size();
    if (n == 0) return;
    std::vector<float> xyzi = to_xyzi(cloud);
    std::vector<uint16_t> line(n);
    for (int i = 0; i < n; i++) line[i] = (uint16_t)(int)cloud[i].normal_y;
    std::vector<uint8_t> label(n, 0);
    ctx_.check(mml_extract_features(ctx_.get(), xyzi.data(), line.data(), n, n_lines, label.data(), nullptr, nullptr));
    for (int i = 0; i < n; i++) cloud[i].normal_z = (float)label[i];
  }

 protected:
  Context& ctx_;
};

// LIO-Livox spelling (SURVEY.md §0 naming caveat)
class LidarFeatureExtractor : public feature_extraction {
 public:
  using feature_extraction::feature_extraction;
  template <class Cloud>
  void detectFeaturePoint(const Cloud& cloud, std::vector<int>& pointsLessSharp, std::vector<int>& pointsLessFlat) {
    detectFeaturePoints(cloud, pointsLessSharp, pointsLessFlat);
  }
};

// ------------------------------------------------------------------------------------------
class Estimator {
 public:
  static const int SLIDEWINDOWSIZE = 5;  // EST.h:30

  struct FeatureLine {       // EST.h:59-84
    std::array<double, 3> pointOri, lineP1, lineP2;
    double error;
    bool valid;
  };
  struct FeaturePlanVec {    // EST.h:107-122; sqrt_info = diag(1,w_t,w_t)/lidar_m * [n t1 t2]^T
    std::array<double, 3> pointOri, pointProj, normal;
    std::array<double, 9> sqrt_info;
    double error;
    bool valid;
  };
  struct Pose {              // the P / Q members of Estimator::LidarFrame (EST.h:33-56)
    std::array<double, 3> P{0, 0, 0};
    std::array<double, 4> Q{1, 0, 0, 0};  // w x y z
  };

  Estimator(Context& ctx, float filter_corner, float filter_surf)
      : ctx_(ctx), filter_corner_(filter_corner), filter_surf_(filter_surf) {}

  // replaces kdtreeCornerFromLocal->setInputCloud etc. (EST.cpp:1159-1167) and the 14553
  // per-cube kd-tree copies (EST.cpp:1171-1179)
  template <class Cloud>
  void setMap(int kind, const Cloud& cloud, const int* cube_centre3 = nullptr) {
    std::vector<float> xyzi = to_xyzi(cloud);
    ctx_.check(mml_map_set(ctx_.get(), kind, xyzi.data(), (int)cloud.size(), cube_centre3));
  }

  // Estimator::MapIncrementLocal (EST.cpp:1585-1643; the non-feature cloud is not on the window-1 path): ring of
  // the last 50 frames, concatenation with the previous filtered map, voxel filter and rebuild of the local maps'
  // search structure on the device. transformTobeMapped = T_wl row-major.
  template <class Cloud>
  void MapIncrementLocal(const Cloud& laserCloudCornerStack, const Cloud& laserCloudSurfStack, const double* transformTobeMapped) {
    std::vector<float> c = to_xyzi(laserCloudCornerStack), s = to_xyzi(laserCloudSurfStack);
    ctx_.check(mml_local_map_push(ctx_.get(), c.data(), (int)laserCloudCornerStack.size(), s.data(),
                                  (int)laserCloudSurfStack.size(), transformTobeMapped, filter_corner_, filter_surf_, nullptr,
                                  nullptr));
  }

  // EST.h:159-165. m4d = T_wl row-major.
  template <class Cloud>
  void processPointToLine(std::vector<FeatureLine>& vLineFeatures, const Cloud& laserCloudCorner, const double* m4d) {
    const int nq = (int)laserCloudCorner.size();
    std::vector<float> q = to_xyzi(laserCloudCorner);
    std::vector<double> feat(12 * (size_t)(nq > 0 ? nq : 1));
    int nf = 0;
    ctx_.check(mml_associate(ctx_.get(), 0, q.data(), nq, m4d, thres_dist, feat.data(), &nf, nullptr, nullptr));
    for (int i = 0; i < nq; i++) {
      const double* f = &feat[12 * (size_t)i];
      if (f[10] < 0) continue;
      vLineFeatures.push_back({{f[0], f[1], f[2]}, {f[3], f[4], f[5]}, {f[6], f[7], f[8]}, f[9], f[10] > 0.5});
    }
  }

  // EST.h:179-186
  template <class Cloud>
  void processPointToPlanVec(std::vector<FeaturePlanVec>& vPlanFeatures, const Cloud& laserCloudSurf, const double* m4d,
                             bool& is_degenerate) {
    const int nq = (int)laserCloudSurf.size();
    std::vector<float> q = to_xyzi(laserCloudSurf);
    std::vector<double> feat(12 * (size_t)(nq > 0 ? nq : 1));
    int nf = 0, nn = 0;
    double M[9];
    ctx_.check(mml_associate(ctx_.get(), 1, q.data(), nq, m4d, thres_dist, feat.data(), &nf, M, &nn));
    const double s = 1.0 / lidar_m;
    for (int i = 0; i < nq; i++) {
      const double* f = &feat[12 * (size_t)i];
      if (f[10] < 0) continue;
      FeaturePlanVec p{{f[0], f[1], f[2]}, {f[3], f[4], f[5]}, {f[6], f[7], f[8]}, {}, f[9], f[10] > 0.5};
      double t1[3], t2[3];
      basis(&f[6], t1, t2);
      const double* B[3] = {&f[6], t1, t2};
      const double sc[3] = {s, s * plan_weight_tan, s * plan_weight_tan};
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) p.sqrt_info[3 * r + c] = sc[r] * B[r][c];
      vPlanFeatures.push_back(p);
    }
    // checkLocalizability, EST.cpp:536-565 + 771-775
    double sv = -1.0;
    if (nn > 10) sv = std::sqrt(std::max(min_eig3(M), 0.0));
    if (sv < 2.0) fail_detected_ = true;
    if (sv < 3.0) is_degenerate = true;
  }

  // EST.h:211-214 for a one-frame list (window size 1, the branch the shipped launch file runs):
  // label split + voxel filter (EST.cpp:992-1026) + Estimate (EST.cpp:1143-1581), on the device.
  template <class Cloud>
  void EstimateLidarPose(Pose& frame, const Cloud& laserCloud, const double* exTlb16) {
    std::vector<float> corner, surf;
    for (const auto& p : laserCloud) {
      if (std::fabs(p.normal_z - 1.0) < 1e-5) corner.insert(corner.end(), {p.x, p.y, p.z, p.intensity});
      else if (std::fabs(p.normal_z - 2.0) < 1e-5) surf.insert(surf.end(), {p.x, p.y, p.z, p.intensity});
    }
    std::vector<float> cds(corner.size() + 4), sds(surf.size() + 4);
    int nc = 0, ns = 0;
    ctx_.check(mml_voxel_downsample(ctx_.get(), corner.data(), (int)corner.size() / 4, filter_corner_, cds.data(), &nc));
    ctx_.check(mml_voxel_downsample(ctx_.get(), surf.data(), (int)surf.size() / 4, filter_surf_, sds.data(), &ns));
    mml_est_params prm;
    mml_est_params_default(&prm);
    double stats[16] = {0};
    ctx_.check(mml_estimate(ctx_.get(), cds.data(), nc, sds.data(), ns, exTlb16, frame.P.data(), frame.Q.data(), &prm, stats));
    fail_detected_ = stats[6] != 0.0;
  }

  bool failureDetected() const { return fail_detected_; }  // EST.h:278

  double thres_dist = 1.0;         // EST.h:336
  double plan_weight_tan = 0.0;    // EST.h:335
  double lidar_m = 1.5e-3;         // IMUIntegrator.h:83

 private:
  static void basis(const double* n, double* t1, double* t2) {
    int k = 0;
    if (std::fabs(n[1]) < std::fabs(n[k])) k = 1;
    if (std::fabs(n[2]) < std::fabs(n[k])) k = 2;
    double e[3] = {0, 0, 0};
    e[k] = 1.0;
    double v[3] = {e[1] * n[2] - e[2] * n[1], e[2] * n[0] - e[0] * n[2], e[0] * n[1] - e[1] * n[0]};
    const double nv = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int i = 0; i < 3; i++) t1[i] = v[i] / nv;
    t2[0] = n[1] * t1[2] - n[2] * t1[1];
    t2[1] = n[2] * t1[0] - n[0] * t1[2];
    t2[2] = n[0] * t1[1] - n[1] * t1[0];
  }
  // smallest eigenvalue of a symmetric 3x3 (closed form, only used for the 2.0 / 3.0 thresholds)
  static double min_eig3(const double* M) {
    const double p1 = M[1] * M[1] + M[2] * M[2] + M[5] * M[5];
    const double q = (M[0] + M[4] + M[8]) / 3.0;
    if (p1 == 0.0) return std::min(M[0], std::min(M[4], M[8]));
    const double p2 = (M[0] - q) * (M[0] - q) + (M[4] - q) * (M[4] - q) + (M[8] - q) * (M[8] - q) + 2 * p1;
    const double p = std::sqrt(p2 / 6.0);
    double B[9];
    for (int i = 0; i < 9; i++) B[i] = (M[i] - (i % 4 == 0 ? q : 0.0)) / p;
    const double detB = B[0] * (B[4] * B[8] - B[5] * B[7]) - B[1] * (B[3] * B[8] - B[5] * B[6]) + B[2] * (B[3] * B[7] - B[4] * B[6]);
    double r = detB / 2.0;
    r = r < -1 ? -1 : (r > 1 ? 1 : r);
    const double phi = std::acos(r) / 3.0;
    return q + 2 * p * std::cos(phi + 2.0 * M_PI / 3.0);
  }

  Context& ctx_;
  float filter_corner_, filter_surf_;
  bool fail_detected_ = false;
};

}  // namespace mmloam
