/*
 * mm-loam hot-path ORACLE — C API.
 *
 * TEST INFRASTRUCTURE ONLY. This is a dependency-free CPU restatement of the
 * reference's scan-matching hot path (TIERS/multi-modal-loam @1daa518). Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it. The product (multi-modal-loam_b200/) never links or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors
 *  and cannot be compiled here (needs ROS + PCL + Eigen + Ceres,
 * none present, no network). Third-party arithmetic is restated from the
 * published algorithms:
 *   PCL 1.8  pcl::VoxelGrid, pcl::KdTreeFLANN (FLANN 1.9 L2_Simple<float>)
 *   Eigen 3.3 SelfAdjointEigenSolver<Matrix3d>, colPivHouseholderQr, slerp
 *   Ceres 2.1.0 TrustRegionMinimizer + DoglegStrategy(TRADITIONAL) + HuberLoss
 * Each function cites the reference file:line it follows
 * (paths relative to /root/reference/mm-loam).
 */
#ifndef MMLOAM_ORACLE_H
#define MMLOAM_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- A1: src/unionFeatureExtract.cpp:341-844 (one scan line) ------------- */
/* xyzi: n x 4 floats. Outputs line-local indices, capacity n each. */
int orc_detect_feature_points(const float* xyzi, int n,
                              int* sharp, int* n_sharp, int* flat, int* n_flat);
/* Debug view of the final CloudFeatureFlag[] (0,1,2,3,100,101,150,300). */
int orc_detect_feature_flags(const float* xyzi, int n, int* flags);

/* ---- A2: src/unionFeatureExtract.cpp:1136-1195 (VLP-16 ring + rel. time) -- */
/* line_out[i] = ring 0..15 or -1 (rejected). reltime_out[i] valid when kept. */
int orc_velo_ring_time(const float* xyzi, int n, int16_t* line_out, float* reltime_out);

/* ---- A3: src/unionFeatureExtract.cpp:985-998 (Horizon CustomPoint filter) - */
/* keep[i]=1 if line<=5 && x>=0.01 ; reltime = offset_time/offset_time_last   */
int orc_hori_filter(const uint32_t* offset_time, const float* xyz3, const uint8_t* line,
                    int n, uint8_t* keep, float* reltime_out);

/* ---- glue: FE.cpp:1001-1023 / 1209-1240: split by line, A1 per line, labels */
/* line_id[i] in [0,n_lines). label_out: 0 none, 1 corner, 2 surf.            */
int orc_extract_scan(const float* xyzi, const uint16_t* line_id, int n, int n_lines,
                     uint8_t* label_out, int threads);

/* ---- A4: src/unionPoseEstimation.cpp:402-421 ------------------------------ */
int orc_undistort(float* xyzi, const float* s, int n, const double* dR9, const double* dt3);

/* ---- A6: pcl::VoxelGrid as used at src/lio/Estimator.cpp:1015-1024 -------- */
/* out capacity n x 4; returns m via *m_out. Output ordered by voxel index.   */
int orc_voxel_downsample(const float* xyzi, int n, float leaf, float* out, int* m_out);

/* ---- A5: src/lio/Map_Manager.cpp:75-89, 583-605 --------------------------- */
void orc_point_to_map(const float* p3, const double* T16, float* out3);
int  orc_cube_index(const float* p3, int cen_w, int cen_h, int cen_d); /* 5000 = outside */

/* ---- map store used by A7/A8 (per-cube kd-trees + local kd-tree) ---------- */
typedef struct orc_map orc_map;
orc_map* orc_map_create(void);
void     orc_map_destroy(orc_map*);
/* kind: 0 corner-global, 1 surf-global, 2 corner-local, 3 surf-local.
 * Global kinds are binned into the 21x11x21 cubes of 50 m with the given centre
 * (cen_w, cen_h, cen_d) = (laserCloudCenWidth, CenHeight, CenDepth).          */
int orc_map_set(orc_map*, int kind, const float* xyzi, int m, const int* cen3);

/* exact 5-NN (ties broken by lower index) in a flat cloud; for tests. */
int orc_knn5_brute(const float* cloud_xyzi, int m, const float* q3, int* idx5, float* d2_5);
int orc_knn5_kdtree(const float* cloud_xyzi, int m, const float* q_xyzi, int nq,
                    int* idx5, float* d2_5);

/* ---- A7: src/lio/Estimator.cpp:148-365 ------------------------------------ */
/* feat: nq x 12 doubles [pointOri(3) lineP1(3) lineP2(3) error valid src];
 * slots with no feature have valid = -1.  valid = 1 if |error| > 1e-5 else 0. */
int orc_associate_line(const orc_map*, const float* q_xyzi, int nq, const double* T_wl16,
                       double thres_dist, double* feat, int* n_feat);
/* ---- A8: src/lio/Estimator.cpp:573-777 ------------------------------------ */
/* feat: nq x 12 doubles [pointOri(3) pointProj(3) n(3) error valid src].
 * normal_moment9: sum n n^T over accepted planes, *n_normals their count.    */
int orc_associate_plane(const orc_map*, const float* q_xyzi, int nq, const double* T_wl16,
                        double thres_dist, double* feat, int* n_feat,
                        double* normal_moment9, int* n_normals);
/* checkLocalizability, EST.cpp:536-565: returns min singular value or -1.    */
double orc_localizability(const double* normal_moment9, int n_normals);

/* ---- A9-A11: include/utils/ceresfunc.h:412-440, 533-555, 33-63 ------------ */
/* Accumulate robustified normal equations for ONE pose x6=[t, phi].
 * H36 row-major 6x6, g6 = J^T r, cost = 1/2 sum rho.  huber_a <= 0: no loss.  */
int orc_accumulate(const double* line_feat, int n_line, const double* plane_feat, int n_plane,
                   const double* x6, const double* T_bl16, double plan_weight_tan,
                   double huber_a, double* H36, double* g6, double* cost, int threads);
/* single-feature residual + analytic Jacobian (for FD tests).
 * kind 0 line (1 residual), 1 plane-vec (3 residuals, canonical basis).      */
int orc_residual(int kind, const double* feat12, const double* x6, const double* T_bl16,
                 double plan_weight_tan, double* r3, double* J18);

/* ---- SO3 helpers: include/sophus/so3.hpp:247-292, 585-623 ------------------ */
void orc_so3_exp(const double* phi3, double* q_wxyz4, double* R9);
void orc_so3_log(const double* q_wxyz4, double* phi3);

/* ---- A12: src/lio/Estimator.cpp:1143-1581, window size 1 (see orc_estimate_window for 2..4) */
typedef struct {
  int    max_outer;        /* 5   EST.cpp:1210 */
  int    max_inner;        /* 10  EST.cpp:1428 */
  double lidar_m;          /* 1.5e-3 IMUIntegrator.h:83 */
  double plan_weight_tan;  /* 0.0 (W != 5) EST.cpp:1206 */
  double thres0, thres1, thres2; /* 25, 10, 1  EST.cpp:1207,1377-1381 */
  int    use_huber;        /* 1 (W != 5) EST.cpp:1221 */
  int    threads;
} orc_est_params;
void orc_est_params_default(orc_est_params*);
/* One frame (W=1). P3/q_wxyz4 are the BODY pose, updated in place.
 * stats (optional, 16 doubles): [outer_iters, inner_iters_total, n_line_last,
 *  n_plane_last, final_cost, min_sv, is_degenerate, ...].                     */
int orc_estimate(const orc_map*, const float* corner_xyzi, int n_corner,
                 const float* surf_xyzi, int n_surf, const double* exTlb16,
                 double* P3, double* q_wxyz4, const orc_est_params*, double* stats);

/* ---- sliding window (2 <= W <= 4: IMU factors, no marginalisation) ------------------------------ */
/* IMUIntegrator::PreIntegration, src/lio/IMUIntegrator.cpp:105-166. t/gyr/acc: n samples (acc in g).
 * preint_out: opaque block of orc_preint_size() bytes.                                                */
int orc_preint_size(void);
int orc_imu_preintegrate(const double* t, const double* gyr, const double* acc, int n, double last_time,
                         const double* bg3, const double* ba3, void* preint_out);
/* Cost_NavState_PRV_Bias (include/utils/ceresfunc.h:321-393) weighted by LLT(cov^-1).matrixL()^T
 * (EST.cpp:1240-1242): r15 and J (15 x 30 row-major, columns [PR_i 6 | VBias_i 9 | PR_j 6 | VBias_j 9]). */
int orc_imu_factor(const void* preint, const double* gravity3, const double* pri6, const double* vbi9,
                   const double* prj6, const double* vbj9, double* r15, double* J450);
/* pose prediction of process(), src/unionPoseEstimation.cpp:812-829. state = P3 q_wxyz4 V3 bg3 ba3.   */
int orc_imu_predict(const double* prev16, const void* preint, double* next16);
/* Estimator::Estimate for window sizes 1..4, src/lio/Estimator.cpp:1143-1581. states: W x 16, in place. */
int orc_estimate_window(const orc_map*, int W, const float* const* corner, const int* n_corner,
                        const float* const* surf, const int* n_surf, const double* exTlb16, double* states,
                        const void* const* preints, const double* gravity3, const orc_est_params*, double* stats);

#ifdef __cplusplus
}
#endif
#endif
