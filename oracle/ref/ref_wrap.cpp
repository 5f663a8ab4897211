// ORACLE / reference pin (test infrastructure only; never linked or loaded by the product).
//
// Compiles the reference's own hot-path text — extracted verbatim from /root/reference by
// oracle/ref/extract.sh into oracle/_ref/gen/*.inc — against the stand-in headers in
// oracle/ref/shim (no Eigen / PCL / ROS / Ceres / Sophus exist in this image), and exports a small
// C API so tests can run the REFERENCE TEXT on the same inputs as the oracle restatement and the
// CUDA path. What is the reference's and what is a stand-in:
//   reference text (verbatim): detectFeaturePoints, getHoriFeatureExtract, the ring/time/label body
//     of getVeloFeature, RemoveLidarDistortion, all of Map_Manager.{h,cpp}, IMUIntegrator.{h,cpp},
//     ceresfunc.{h,cpp} (cost functors, marginalisation), Estimator.{h,cpp} (association, Estimate,
//     EstimateLidarPose, MapIncrementLocal, ...);
//   stand-ins (restated, see each header): Eigen 3.3 subset, PCL 1.8 kd-tree / voxel grid, Ceres
//     Jet autodiff + Problem/Solve (the oracle's dogleg), Sophus SO3, ROS macros.
// Built by `make -C oracle ref` into oracle/_ref/libmmloam_ref.so (git-ignored, travels to the GPU box).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <iterator>
#include <list>
#include <memory>
#include <mutex>
#include <numeric>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>
#include <pthread.h>

#include "shim/ref_eigen.h"
#include "shim/ref_sophus.h"
#include "shim/ref_pcl.h"
#include "shim/ref_ros.h"
#include "shim/ref_ceres.h"

// ---- the map thread of Estimator (EST.cpp:92-145) runs `while(true) { ...; r.sleep(); }`. The
// stand-in for ros::Rate::sleep() parks the thread until the test asks for one more pass, which
// makes the asynchronous map update deterministic, and ends the thread on shutdown.
namespace mmlref {
struct Gate {
  std::mutex m; std::condition_variable cv;
  int allowed = 0, done = 0; bool shutdown = false;
};
static std::mutex g_gate_mu;
static Gate* g_gate_pending = nullptr;  // gate of the Estimator under construction
}  // namespace mmlref
namespace ros {
struct Rate {
  mmlref::Gate* g;
  explicit Rate(double) { std::lock_guard<std::mutex> l(mmlref::g_gate_mu); g = mmlref::g_gate_pending; }
  void sleep() {
    if (!g) { std::this_thread::sleep_for(std::chrono::milliseconds(1)); return; }
    std::unique_lock<std::mutex> l(g->m);
    g->done++;
    g->cv.notify_all();
    g->cv.wait(l, [&] { return g->allowed > 0 || g->shutdown; });
    if (g->shutdown) { l.unlock(); pthread_exit(nullptr); }
    g->allowed--;
  }
};
inline bool ok() { return true; }
}  // namespace ros

static int mml_ref_cap = 20000;  // capacity of detectFeaturePoints' work arrays (extract.sh, edit 1)

// The reference text prints solver summaries and debug values with std::cout; they go to a sink.
namespace std {
struct mml_null_stream_t {
  template <class T> mml_null_stream_t& operator<<(const T&) { return *this; }
  mml_null_stream_t& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
static mml_null_stream_t mml_null_stream;
}  // namespace std
#define cout mml_null_stream

#define private public
#define protected public
// ---- reference text -------------------------------------------------------------------------
#include "../_ref/gen/mm_h.inc"
#include "../_ref/gen/mm_cpp.inc"
#include "../_ref/gen/imu_h.inc"
#include "../_ref/gen/imu_cpp.inc"
#include "../_ref/gen/cf_h.inc"
#include "../_ref/gen/cf_cpp.inc"
#include "../_ref/gen/est_h.inc"
#include "../_ref/gen/est_cpp.inc"

typedef pcl::PointXYZINormal PointType;
#include "../_ref/gen/pe_undistort.inc"

class feature_extraction {
 public:
  int VELO_N_SCANS = 16;  // FE.cpp:192
#include "../_ref/gen/fe_detect.inc"
#include "../_ref/gen/fe_hori.inc"
  // FE.cpp:1113-1135 reads a PointCloud2 into lidar_cloud_in and drops NaN points; the body below
  // (FE.cpp:1135-1240) is the reference's: ring, relative time, line split, detector, labels.
  void veloBody(pcl::PointCloud<pcl::PointXYZI>& lidar_cloud_in, pcl::PointCloud<pcl::PointXYZINormal>::Ptr& laserCloudOut) {
#include "../_ref/gen/fe_velo_body.inc"
    laserCloudOut = laserCloud;
  }
};
#undef private
#undef protected
#undef cout

// ---- C API ------------------------------------------------------------------------------------
namespace {
using Cloud = pcl::PointCloud<PointType>;
Cloud::Ptr make_cloud(const float* xyzi, int n) {
  Cloud::Ptr c(new Cloud);
  c->reserve(n);
  for (int i = 0; i < n; i++) { PointType p; p.x = xyzi[4 * i]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.intensity = xyzi[4 * i + 3]; c->push_back(p); }
  return c;
}
void dump7(const Cloud& c, float* out) {
  for (size_t i = 0; i < c.points.size(); i++) {
    const PointType& p = c.points[i];
    float* o = out + 7 * i;
    o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = p.intensity; o[4] = p.normal_x; o[5] = p.normal_y; o[6] = p.normal_z;
  }
}
Eigen::Matrix4d mat4(const double* T16) {
  Eigen::Matrix4d T;
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T(r, c) = T16[4 * r + c];
  return T;
}
struct RefEst {
  mmlref::Gate gate;
  Estimator* est = nullptr;
};
}  // namespace

extern "C" {

const char* ref_describe() {
  return "reference text compiled verbatim (TIERS/multi-modal-loam @1daa518) against stand-in Eigen/PCL/Ceres/Sophus/ROS headers";
}

// A1: feature_extraction::detectFeaturePoints, FE.cpp:341-844
int ref_detect_feature_points(const float* xyzi, int n, int* sharp, int* n_sharp, int* flat, int* n_flat) {
  feature_extraction fe;
  mml_ref_cap = std::max(n + 16, 64);
  Cloud::Ptr c = make_cloud(xyzi, n);
  std::vector<int> s, f;
  fe.detectFeaturePoints(c, s, f);
  std::copy(s.begin(), s.end(), sharp);
  std::copy(f.begin(), f.end(), flat);
  *n_sharp = (int)s.size();
  *n_flat = (int)f.size();
  return 0;
}

// A3 + split + A1 + labels: feature_extraction::getHoriFeatureExtract, FE.cpp:952-1035.
// cloud7_out rows: x y z intensity normal_x(rel. time) normal_y(line) normal_z(label), capacity n.
int ref_hori_extract(const uint32_t* offset_time, const float* xyz3, const uint8_t* refl, const uint8_t* line, int n,
                     int used_line, float* cloud7_out, int* n_out, int* n_corner, int* n_surf) {
  auto msg = std::make_shared<livox_ros_driver::CustomMsg>();
  msg->points.resize(n);
  msg->point_num = n;
  for (int i = 0; i < n; i++) {
    auto& p = msg->points[i];
    p.offset_time = offset_time[i]; p.x = xyz3[3 * i]; p.y = xyz3[3 * i + 1]; p.z = xyz3[3 * i + 2];
    p.reflectivity = refl[i]; p.tag = 0; p.line = line[i];
  }
  feature_extraction fe;
  mml_ref_cap = std::max(n + 16, 64);
  Cloud::Ptr cloud(new Cloud), corner(new Cloud), surf(new Cloud);
  livox_ros_driver::CustomMsgConstPtr cmsg = msg;
  fe.getHoriFeatureExtract(cmsg, cloud, corner, surf, used_line);
  dump7(*cloud, cloud7_out);
  *n_out = (int)cloud->size(); *n_corner = (int)corner->size(); *n_surf = (int)surf->size();
  return 0;
}

// A2 + split + A1 + labels: body of feature_extraction::getVeloFeature, FE.cpp:1135-1240.
int ref_velo_extract(const float* xyzi, int n, float* cloud7_out, int* n_out) {
  pcl::PointCloud<pcl::PointXYZI> in;
  for (int i = 0; i < n; i++) { pcl::PointXYZI p; p.x = xyzi[4 * i]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.intensity = xyzi[4 * i + 3]; in.push_back(p); }
  feature_extraction fe;
  mml_ref_cap = std::max(n + 16, 64);
  Cloud::Ptr out;
  fe.veloBody(in, out);
  dump7(*out, cloud7_out);
  *n_out = (int)out->size();
  return 0;
}

// A4: RemoveLidarDistortion, PE.cpp:402-421 (in place; s = per-point sweep fraction)
int ref_undistort(float* xyzi, const float* s, int n, const double* dR9, const double* dt3) {
  Cloud::Ptr c = make_cloud(xyzi, n);
  for (int i = 0; i < n; i++) c->points[i].normal_x = s[i];
  Eigen::Matrix3d dR;
  for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) dR(r, k) = dR9[3 * r + k];
  Eigen::Vector3d dt(dt3[0], dt3[1], dt3[2]);
  RemoveLidarDistortion(c, dR, dt);
  for (int i = 0; i < n; i++) { xyzi[4 * i] = c->points[i].x; xyzi[4 * i + 1] = c->points[i].y; xyzi[4 * i + 2] = c->points[i].z; }
  return 0;
}

// A5: MAP_MANAGER::pointAssociateToMap MM.cpp:75-89, FindUsedCornerMap / FindUsedSurfMap MM.cpp:583-629
void ref_point_to_map(const float* p3, const double* T16, float* out3) {
  PointType pi, po;
  pi.x = p3[0]; pi.y = p3[1]; pi.z = p3[2];
  MAP_MANAGER::pointAssociateToMap(&pi, &po, mat4(T16));
  out3[0] = po.x; out3[1] = po.y; out3[2] = po.z;
}

// ---- Estimator object -------------------------------------------------------------------------
void* ref_est_create(float filter_corner, float filter_surf) {
  RefEst* h = new RefEst;
  {
    std::lock_guard<std::mutex> l(mmlref::g_gate_mu);
    mmlref::g_gate_pending = &h->gate;
  }
  h->est = new Estimator(filter_corner, filter_surf);  // starts threadMapIncrement (EST.cpp:58)
  {  // wait for the map thread's first pass, so that it holds this gate
    std::unique_lock<std::mutex> l(h->gate.m);
    h->gate.cv.wait(l, [&] { return h->gate.done > 0; });
  }
  {
    std::lock_guard<std::mutex> l(mmlref::g_gate_mu);
    mmlref::g_gate_pending = nullptr;
  }
  return h;
}
void ref_est_destroy(void* hv) {
  RefEst* h = (RefEst*)hv;
  {
    std::lock_guard<std::mutex> l(h->gate.m);
    h->gate.shutdown = true;
  }
  h->gate.cv.notify_all();
  h->est->threadMap.join();
  delete h->est;
  delete h;
}
// one pass of threadMapIncrement's loop (EST.cpp:101-143)
void ref_est_map_thread_step(void* hv) {
  RefEst* h = (RefEst*)hv;
  std::unique_lock<std::mutex> l(h->gate.m);
  int target = h->gate.done + 1;
  h->gate.allowed++;
  h->gate.cv.notify_all();
  h->gate.cv.wait(l, [&] { return h->gate.done >= target; });
}
int ref_cube_index(void* hv, const float* p3, int kind) {
  Estimator* e = ((RefEst*)hv)->est;
  PointType p; p.x = p3[0]; p.y = p3[1]; p.z = p3[2];
  MAP_MANAGER* mm = e->map_manager;
  return (int)(kind == 0 ? mm->FindUsedCornerMap(&p, mm->laserCloudCenWidth, mm->laserCloudCenHeight, mm->laserCloudCenDepth)
                         : mm->FindUsedSurfMap(&p, mm->laserCloudCenWidth, mm->laserCloudCenHeight, mm->laserCloudCenDepth));
}
// MAP_MANAGER::MapIncrement, MM.cpp:125-281 (points already in the map frame)
int ref_est_map_increment(void* hv, const float* corner, int nc, const float* surf, int ns, const double* T16) {
  Estimator* e = ((RefEst*)hv)->est;
  Cloud::Ptr c = make_cloud(corner, nc), s = make_cloud(surf, ns), nf(new Cloud);
  e->map_manager->MapIncrement(c, s, nf, mat4(T16));
  return 0;
}
// kind 0/1: cubes' clouds as Estimate will see them after its copy (laserCloud*_for_match), 2/3: the
// cubes' current clouds (laserCloud*Array). Concatenated in cube order; cen3 = (CenWidth, CenHeight, CenDepth).
int ref_est_get_global_map(void* hv, int kind, float* out_xyzi, int cap, int* m_out, int* cen3) {
  Estimator* e = ((RefEst*)hv)->est;
  MAP_MANAGER* mm = e->map_manager;
  int m = 0;
  for (int i = 0; i < MAP_MANAGER::laserCloudNum; i++) {
    const Cloud& c = kind == 0 ? mm->laserCloudCorner_for_match[i] : kind == 1 ? mm->laserCloudSurf_for_match[i]
                     : kind == 2 ? *mm->laserCloudCornerArray[i] : *mm->laserCloudSurfArray[i];
    for (const PointType& p : c.points) {
      if (m < cap) { out_xyzi[4 * m] = p.x; out_xyzi[4 * m + 1] = p.y; out_xyzi[4 * m + 2] = p.z; out_xyzi[4 * m + 3] = p.intensity; }
      m++;
    }
  }
  *m_out = m;
  if (kind <= 1) { cen3[0] = mm->laserCloudCenWidth_last; cen3[1] = mm->laserCloudCenHeight_last; cen3[2] = mm->laserCloudCenDepth_last; }
  else { cen3[0] = mm->laserCloudCenWidth; cen3[1] = mm->laserCloudCenHeight; cen3[2] = mm->laserCloudCenDepth; }
  return 0;
}
// Estimator::MapIncrementLocal, EST.cpp:1585-1643 (points in the LiDAR frame + T_wl)
int ref_est_map_increment_local(void* hv, const float* corner, int nc, const float* surf, int ns, const double* T16) {
  Estimator* e = ((RefEst*)hv)->est;
  Cloud::Ptr c = make_cloud(corner, nc), s = make_cloud(surf, ns), nf(new Cloud);
  e->MapIncrementLocal(c, s, nf, mat4(T16));
  return 0;
}
int ref_est_set_local_map(void* hv, int kind, const float* xyzi, int m) {
  Estimator* e = ((RefEst*)hv)->est;
  if (kind == 0) *e->laserCloudCornerFromLocal = *make_cloud(xyzi, m);
  else *e->laserCloudSurfFromLocal = *make_cloud(xyzi, m);
  return 0;
}
int ref_est_get_local_map(void* hv, int kind, float* out_xyzi, int cap, int* m_out) {
  Estimator* e = ((RefEst*)hv)->est;
  const Cloud& c = kind == 0 ? *e->laserCloudCornerFromLocal : *e->laserCloudSurfFromLocal;
  int m = 0;
  for (const PointType& p : c.points) {
    if (m < cap) { out_xyzi[4 * m] = p.x; out_xyzi[4 * m + 1] = p.y; out_xyzi[4 * m + 2] = p.z; out_xyzi[4 * m + 3] = p.intensity; }
    m++;
  }
  *m_out = m;
  return 0;
}
// what Estimate does before the iterations, EST.cpp:1159-1184: local kd-trees, copies of the cubes
static void est_prepare(Estimator* e) {
  if (e->laserCloudCornerFromLocal->points.size()) e->kdtreeCornerFromLocal->setInputCloud(e->laserCloudCornerFromLocal);
  if (e->laserCloudSurfFromLocal->points.size()) e->kdtreeSurfFromLocal->setInputCloud(e->laserCloudSurfFromLocal);
  MAP_MANAGER* mm = e->map_manager;
  for (int i = 0; i < 4851; i++) {
    e->CornerKdMap[i] = mm->getCornerKdMap(i);
    e->SurfKdMap[i] = mm->getSurfKdMap(i);
    e->GlobalSurfMap[i] = mm->laserCloudSurf_for_match[i];
    e->GlobalCornerMap[i] = mm->laserCloudCorner_for_match[i];
  }
  e->laserCenWidth_last = mm->get_laserCloudCenWidth_last();
  e->laserCenHeight_last = mm->get_laserCloudCenHeight_last();
  e->laserCenDepth_last = mm->get_laserCloudCenDepth_last();
}
// A7: Estimator::processPointToLine, EST.cpp:148-365. feat rows (12 doubles):
// pointOri(3) lineP1(3) lineP2(3) error valid(|error|>1e-5) source-row(-1: the reference does not record it)
int ref_est_associate_line(void* hv, const float* q_xyzi, int nq, const double* T_wl16, const double* exTlb16,
                           double thres_dist, double* feat, int* n_feat) {
  Estimator* e = ((RefEst*)hv)->est;
  est_prepare(e);
  e->thres_dist = thres_dist;
  std::vector<ceres::CostFunction*> edges;
  std::vector<Estimator::FeatureLine> v;
  Cloud::Ptr q = make_cloud(q_xyzi, nq);
  e->processPointToLine(edges, v, q, e->laserCloudCornerFromLocal, e->kdtreeCornerFromLocal, mat4(exTlb16), mat4(T_wl16));
  for (size_t i = 0; i < v.size(); i++) {
    double* o = feat + 12 * i;
    for (int k = 0; k < 3; k++) { o[k] = v[i].pointOri[k]; o[3 + k] = v[i].lineP1[k]; o[6 + k] = v[i].lineP2[k]; }
    o[9] = v[i].error; o[10] = std::fabs(v[i].error) > 1e-5 ? 1 : 0; o[11] = -1;
  }
  for (auto* c : edges) delete c;
  *n_feat = (int)v.size();
  return 0;
}
// A8: Estimator::processPointToPlanVec, EST.cpp:573-777 + checkLocalizability 536-565. feat rows (18 doubles):
// pointOri(3) pointProj(3) sqrt_info(9, row-major) error valid -1
int ref_est_associate_plane(void* hv, const float* q_xyzi, int nq, const double* T_wl16, const double* exTlb16,
                            double thres_dist, double plan_weight_tan, double* feat, int* n_feat, int* is_degenerate,
                            int* fail_detected) {
  Estimator* e = ((RefEst*)hv)->est;
  est_prepare(e);
  e->thres_dist = thres_dist;
  e->plan_weight_tan = plan_weight_tan;
  std::vector<ceres::CostFunction*> edges;
  std::vector<Estimator::FeaturePlanVec> v;
  Cloud::Ptr q = make_cloud(q_xyzi, nq);
  bool deg = false;
  e->_fail_detected = false;
  e->processPointToPlanVec(edges, v, q, e->laserCloudSurfFromLocal, e->kdtreeSurfFromLocal, mat4(exTlb16), mat4(T_wl16), deg);
  for (size_t i = 0; i < v.size(); i++) {
    double* o = feat + 18 * i;
    for (int k = 0; k < 3; k++) { o[k] = v[i].pointOri[k]; o[3 + k] = v[i].pointProj[k]; }
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) o[6 + 3 * r + c] = v[i].sqrt_info(r, c);
    o[15] = v[i].error; o[16] = std::fabs(v[i].error) > 1e-5 ? 1 : 0; o[17] = -1;
  }
  for (auto* c : edges) delete c;
  *n_feat = (int)v.size();
  *is_degenerate = deg ? 1 : 0;
  *fail_detected = e->_fail_detected ? 1 : 0;
  return 0;
}
double ref_localizability(void* hv, const double* normals, int n) {
  Estimator* e = ((RefEst*)hv)->est;
  std::vector<Eigen::Vector3d> v;
  for (int i = 0; i < n; i++) v.emplace_back(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
  return e->checkLocalizability(v);
}

// A9 / A10: the reference's cost functors through dual-number autodiff (CF.h:397-458, 517-570).
// kind 0: line, feat = pointOri(3) lineP1(3) lineP2(3); kind 1: plane-vec, feat = pointOri(3) pointProj(3) sqrt_info(9).
// Tbl16 = exTlb^-1 (EST.cpp:157-159). r3 / J18: residuals and row-major Jacobian (1x6 or 3x6).
int ref_residual(int kind, const double* feat, const double* x6, const double* Tbl16, double lidar_m, double* r3, double* J18) {
  Eigen::Matrix4d Tbl = mat4(Tbl16);
  const double* params[1] = {x6};
  double* jac[1] = {J18};
  ceres::CostFunction* c;
  if (kind == 0)
    c = Cost_NavState_IMU_Line::Create(Eigen::Vector3d(feat[0], feat[1], feat[2]), Eigen::Vector3d(feat[3], feat[4], feat[5]),
                                       Eigen::Vector3d(feat[6], feat[7], feat[8]), Tbl, Eigen::Matrix<double, 1, 1>(1 / lidar_m));
  else {
    Eigen::Matrix3d si;
    for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) si(r, k) = feat[6 + 3 * r + k];
    c = Cost_NavState_IMU_Plan_Vec::Create(Eigen::Vector3d(feat[0], feat[1], feat[2]), Eigen::Vector3d(feat[3], feat[4], feat[5]), Tbl, si);
  }
  bool ok = c->Evaluate(params, r3, jac);
  delete c;
  return ok ? 0 : -1;
}

// A12 (+A6, map hand-over): Estimator::EstimateLidarPose, EST.cpp:967-1141, window size 1.
// cloud7 rows as produced by ref_*_extract (label in column 6). P3 / q_wxyz4: body pose, in place.
int ref_est_estimate_lidar_pose(void* hv, const float* cloud7, int n, double* P3, double* q_wxyz4, const double* exTlb16,
                                int lidarMode, int* fail_detected) {
  Estimator* e = ((RefEst*)hv)->est;
  std::list<Estimator::LidarFrame> frames;
  frames.emplace_back();
  Estimator::LidarFrame& f = frames.back();
  f.laserCloud.reset(new Cloud);
  for (int i = 0; i < n; i++) {
    PointType p; const float* o = cloud7 + 7 * i;
    p.x = o[0]; p.y = o[1]; p.z = o[2]; p.intensity = o[3]; p.normal_x = o[4]; p.normal_y = o[5]; p.normal_z = o[6];
    f.laserCloud->push_back(p);
  }
  f.P = Eigen::Vector3d(P3[0], P3[1], P3[2]);
  f.Q = Eigen::Quaterniond(q_wxyz4[0], q_wxyz4[1], q_wxyz4[2], q_wxyz4[3]);
  Eigen::Vector3d g(0, 0, -9.805);
  e->EstimateLidarPose(frames, mat4(exTlb16), g, lidarMode);
  const Estimator::LidarFrame& r = frames.front();
  P3[0] = r.P.x(); P3[1] = r.P.y(); P3[2] = r.P.z();
  q_wxyz4[0] = r.Q.w(); q_wxyz4[1] = r.Q.x(); q_wxyz4[2] = r.Q.y(); q_wxyz4[3] = r.Q.z();
  *fail_detected = e->failureDetected() ? 1 : 0;
  return 0;
}

// ---- sliding window: IMUIntegrator (IMU.cpp verbatim), Cost_NavState_PRV_Bias (CF.h:321-393 verbatim, dual-number
// autodiff), Estimator::EstimateLidarPose / Estimate with W frames (EST.cpp verbatim) ---------------------------
static std::vector<sensor_msgs::ImuConstPtr> make_imu(const double* t, const double* gyr, const double* acc, int n) {
  std::vector<sensor_msgs::ImuConstPtr> v;
  for (int i = 0; i < n; i++) {
    auto m = std::make_shared<sensor_msgs::Imu>();
    m->header.stamp.fromSec(t[i]);
    m->angular_velocity.x = gyr[3 * i]; m->angular_velocity.y = gyr[3 * i + 1]; m->angular_velocity.z = gyr[3 * i + 2];
    m->linear_acceleration.x = acc[3 * i]; m->linear_acceleration.y = acc[3 * i + 1]; m->linear_acceleration.z = acc[3 * i + 2];
    v.push_back(m);
  }
  return v;
}
// out: dq(wxyz) dp dv dt | cov 15x15 row-major | jac 15x15 row-major  (11 + 225 + 225 doubles)
int ref_imu_preintegrate(const double* t, const double* gyr, const double* acc, int n, double last_time, const double* bg,
                         const double* ba, double* out) {
  IMUIntegrator imu(make_imu(t, gyr, acc, n));
  imu.PreIntegration(last_time, Eigen::Vector3d(bg[0], bg[1], bg[2]), Eigen::Vector3d(ba[0], ba[1], ba[2]));
  const Eigen::Quaterniond& q = imu.GetDeltaQ();
  out[0] = q.w(); out[1] = q.x(); out[2] = q.y(); out[3] = q.z();
  for (int k = 0; k < 3; k++) { out[4 + k] = imu.GetDeltaP()[k]; out[7 + k] = imu.GetDeltaV()[k]; }
  out[10] = imu.GetDeltaTime();
  for (int r = 0; r < 15; r++) for (int c = 0; c < 15; c++) { out[11 + 15 * r + c] = imu.GetCovariance()(r, c); out[236 + 15 * r + c] = imu.GetJacobian()(r, c); }
  return 0;
}
int ref_imu_factor(const double* t, const double* gyr, const double* acc, int n, double last_time, const double* bg,
                   const double* ba, const double* gravity, const double* pri, const double* vbi, const double* prj,
                   const double* vbj, double* r15, double* J450) {
  IMUIntegrator imu(make_imu(t, gyr, acc, n));
  imu.PreIntegration(last_time, Eigen::Vector3d(bg[0], bg[1], bg[2]), Eigen::Vector3d(ba[0], ba[1], ba[2]));
  Eigen::Vector3d g(gravity[0], gravity[1], gravity[2]);
  // EST.cpp:1238-1242
  ceres::CostFunction* c = Cost_NavState_PRV_Bias::Create(imu, g,
      Eigen::LLT<Eigen::Matrix<double, 15, 15>>(imu.GetCovariance().inverse()).matrixL().transpose());
  const double* params[4] = {pri, vbi, prj, vbj};
  double J0[90], J1[135], J2[90], J3[135];
  double* jac[4] = {J0, J1, J2, J3};
  bool ok = c->Evaluate(params, r15, jac);
  for (int r = 0; r < 15; r++) {
    for (int k = 0; k < 6; k++) { J450[30 * r + k] = J0[6 * r + k]; J450[30 * r + 15 + k] = J2[6 * r + k]; }
    for (int k = 0; k < 9; k++) { J450[30 * r + 6 + k] = J1[9 * r + k]; J450[30 * r + 21 + k] = J3[9 * r + k]; }
  }
  delete c;
  return ok ? 0 : -1;
}
// EstimateLidarPose on a list of W frames, as process() hands it over in the IMU-initialised state
// (PE.cpp:796-835): frame f >= 1 carries the IMU messages of (t_{f-1}, t_f] and was pre-integrated with the
// previous frame's biases. clouds: rows of 7 floats (label in column 6), concatenated; states W x 16 in place
// (P, q_wxyz, V, bg, ba); imu_*: concatenated samples, imu_n[f] per frame; stamps[f] = frame time.
int ref_est_estimate_window(void* hv, int W, const float* clouds7, const int* n_pts, double* states, const double* stamps,
                            const double* imu_t, const double* imu_gyr, const double* imu_acc, const int* imu_n,
                            const double* exTlb16, const double* gravity, int lidarMode, int* fail_detected) {
  Estimator* e = ((RefEst*)hv)->est;
  std::list<Estimator::LidarFrame> frames;
  size_t po = 0, io = 0;
  for (int f = 0; f < W; f++) {
    frames.emplace_back();
    Estimator::LidarFrame& fr = frames.back();
    fr.laserCloud.reset(new Cloud);
    for (int i = 0; i < n_pts[f]; i++) {
      PointType p; const float* o = clouds7 + 7 * (po + i);
      p.x = o[0]; p.y = o[1]; p.z = o[2]; p.intensity = o[3]; p.normal_x = o[4]; p.normal_y = o[5]; p.normal_z = o[6];
      fr.laserCloud->push_back(p);
    }
    po += n_pts[f];
    const double* s = states + 16 * f;
    fr.P = Eigen::Vector3d(s[0], s[1], s[2]);
    fr.Q = Eigen::Quaterniond(s[3], s[4], s[5], s[6]);
    fr.V = Eigen::Vector3d(s[7], s[8], s[9]);
    fr.bg = Eigen::Vector3d(s[10], s[11], s[12]);
    fr.ba = Eigen::Vector3d(s[13], s[14], s[15]);
    fr.timeStamp = stamps[f];
    if (f >= 1) {
      fr.imuIntegrator.PushIMUMsg(make_imu(imu_t + io, imu_gyr + 3 * io, imu_acc + 3 * io, imu_n[f]));
      const double* sp = states + 16 * (f - 1);
      fr.imuIntegrator.PreIntegration(stamps[f - 1], Eigen::Vector3d(sp[10], sp[11], sp[12]), Eigen::Vector3d(sp[13], sp[14], sp[15]));
    }
    io += imu_n[f];
  }
  Eigen::Vector3d g(gravity[0], gravity[1], gravity[2]);
  e->EstimateLidarPose(frames, mat4(exTlb16), g, lidarMode);
  int f = 0;
  for (const auto& fr : frames) {
    double* s = states + 16 * f++;
    for (int k = 0; k < 3; k++) { s[k] = fr.P[k]; s[7 + k] = fr.V[k]; s[10 + k] = fr.bg[k]; s[13 + k] = fr.ba[k]; }
    s[3] = fr.Q.w(); s[4] = fr.Q.x(); s[5] = fr.Q.y(); s[6] = fr.Q.z();
  }
  *fail_detected = e->failureDetected() ? 1 : 0;
  return 0;
}

}  // extern "C"
