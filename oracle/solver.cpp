// ORACLE (test infrastructure only — see oracle.h).
//   orc_estimate()  <- Estimator::Estimate, window sizes without IMU factors
//                      src/lio/Estimator.cpp:1143-1581 (control flow), 937-964 (state packing)
//   DoglegSolver    <- ceres::Solve with the options of EST.cpp:1425-1432:
//                      TrustRegionMinimizer + DoglegStrategy(TRADITIONAL_DOGLEG) + DENSE_SCHUR,
//                      max_num_iterations 10, Jacobi scaling, Ceres 2.1.0 defaults otherwise
//                      (initial radius 1e4, function_tolerance 1e-6, gradient_tolerance 1e-10,
//                       parameter_tolerance 1e-8, min_relative_decrease 1e-3, dogleg mu in
//                       [1e-8, 1] x10, radius *0.5 below 0.25 / max(r, 3|step|) above 0.75).
// Ceres is not in the tree or the container: the minimiser is restated from its published
// algorithm on the normal equations (H = J^T J, g = J^T r of the loss-corrected problem),
// which carry everything Ceres' step computation uses. Parity unpinned; converged poses are
// what the parity tests compare, not step sequences.
#include "oracle.h"
#include "oracle_math.h"
#include <cmath>
#include <cstring>
#include <functional>
#include <limits>
#include <vector>

#include "dogleg.h"
namespace orc {
void accumulate_pose(const double* lf, int nl, const double* pf, int np, const double* x6, const double* Tbl,
                     double lidar_m, double w_tan, double a, double* H36, double* g6, double* cost, int threads);

}  // namespace orc

using namespace orc;

extern "C" int orc_associate_line_mt(const orc_map* m, const float* q_xyzi, int nq, const double* T, double thres_dist,
                                     double* feat, int* n_feat, int threads);
extern "C" int orc_associate_plane_mt(const orc_map* m, const float* q_xyzi, int nq, const double* T, double thres_dist,
                                      double* feat, int* n_feat, double* M9, int* n_normals, int threads);

extern "C" {

void orc_est_params_default(orc_est_params* p) {
  p->max_outer = 5;
  p->max_inner = 10;
  p->lidar_m = 1.5e-3;
  p->plan_weight_tan = 0.0;
  p->thres0 = 25.0;
  p->thres1 = 10.0;
  p->thres2 = 1.0;
  p->use_huber = 1;
  p->threads = 1;
}

// EST.cpp:1143-1581 for windowSize == 1 (the branch the shipped launch file runs,
// SURVEY.md §3.3): re-associate every outer iteration, Huber(0.1/lidar_m), stop when
// dR < 0.05 deg and dT < 0.05 m.
int orc_estimate(const orc_map* map, const float* corner, int n_corner, const float* surf, int n_surf,
                 const double* exTlb, double* P3, double* q4, const orc_est_params* prm, double* stats) {
  // exRbl = R^T, exPbl = -R^T t (EST.cpp:1155-1156); Tbl = exTlb^-1 (EST.cpp:157-159)
  double Rbl[9], Pbl[3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) Rbl[3 * r + c] = exTlb[4 * c + r];
  for (int r = 0; r < 3; r++) Pbl[r] = -1.0 * (Rbl[3 * r] * exTlb[3] + Rbl[3 * r + 1] * exTlb[7] + Rbl[3 * r + 2] * exTlb[11]);
  double Tbl[16] = {Rbl[0], Rbl[1], Rbl[2], Pbl[0], Rbl[3], Rbl[4], Rbl[5], Pbl[1], Rbl[6], Rbl[7], Rbl[8], Pbl[2], 0, 0, 0, 1};

  std::vector<double> lf((size_t)std::max(n_corner, 1) * 12), pf((size_t)std::max(n_surf, 1) * 12);
  double thres = prm->thres0;
  double huber_a = prm->use_huber ? 0.1 / prm->lidar_m : 0.0;
  int outer_done = 0, inner_total = 0, nl = 0, np = 0, n_normals = 0;
  double final_cost = 0, min_sv = -1;
  int is_degenerate = 0;
  Quat Q = {q4[0], q4[1], q4[2], q4[3]};
  double P[3] = {P3[0], P3[1], P3[2]};

  for (int it = 0; it < prm->max_outer; ++it) {
    // vector2double, EST.cpp:937-950
    double x[6] = {P[0], P[1], P[2], 0, 0, 0};
    so3_log(Q, x + 3);
    Quat q_before = Q;
    double t_before[3] = {P[0], P[1], P[2]};
    // T_wl = [Q exRbl, Q exPbl + P]  (EST.cpp:1268-1270)
    double Rq[9];
    quat_to_R(Q, Rq);
    double T[16] = {0};
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) T[4 * r + c] = Rq[3 * r] * Rbl[c] + Rq[3 * r + 1] * Rbl[3 + c] + Rq[3 * r + 2] * Rbl[6 + c];
      T[4 * r + 3] = Rq[3 * r] * Pbl[0] + Rq[3 * r + 1] * Pbl[1] + Rq[3 * r + 2] * Pbl[2] + P[r];
    }
    T[15] = 1;
    double M9[9];
    orc_associate_line_mt(map, corner, n_corner, T, thres, lf.data(), &nl, prm->threads);
    orc_associate_plane_mt(map, surf, n_surf, T, thres, pf.data(), &np, M9, &n_normals, prm->threads);
    min_sv = orc_localizability(M9, n_normals);  // EST.cpp:771-775
    if (min_sv < 3.0) is_degenerate = 1;
    thres = (it == 0) ? prm->thres1 : prm->thres2;  // EST.cpp:1377-1381

    EvalFn eval = [&](const double* xx, double* cost, double* H, double* g) {
      double Hh[36], gg[6];
      accumulate_pose(lf.data(), n_corner, pf.data(), n_surf, xx, Tbl, prm->lidar_m, prm->plan_weight_tan, huber_a, Hh, gg,
                      cost, prm->threads);
      if (H) std::memcpy(H, Hh, sizeof(Hh));
      if (g) std::memcpy(g, gg, sizeof(gg));
      return std::isfinite(*cost);
    };
    DoglegSummary s = dogleg_minimize(6, x, eval, prm->max_inner);
    inner_total += s.iterations;
    final_cost = s.final_cost;
    // double2vector, EST.cpp:952-964
    P[0] = x[0]; P[1] = x[1]; P[2] = x[2];
    Q = so3_exp(x + 3);
    outer_done = it + 1;
    // EST.cpp:1441-1450: Eigen 3.3 angularDistance = 2*atan2(|vec(q1 q2*)|, |w|), in degrees
    Quat dq = quat_mul(q_before, Quat{Q.w, -Q.x, -Q.y, -Q.z});
    double deltaR = 2.0 * std::atan2(std::sqrt((dq.x * dq.x + dq.y * dq.y) + dq.z * dq.z), std::fabs(dq.w)) * 180.0 / M_PI;
    double dt[3] = {t_before[0] - P[0], t_before[1] - P[1], t_before[2] - P[2]};
    double deltaT = std::sqrt((dt[0] * dt[0] + dt[1] * dt[1]) + dt[2] * dt[2]);
    if ((deltaR < 0.05 && deltaT < 0.05) || (it + 1) == prm->max_outer) break;
  }
  P3[0] = P[0]; P3[1] = P[1]; P3[2] = P[2];
  q4[0] = Q.w; q4[1] = Q.x; q4[2] = Q.y; q4[3] = Q.z;
  if (stats) {
    stats[0] = outer_done; stats[1] = inner_total; stats[2] = nl; stats[3] = np;
    stats[4] = final_cost; stats[5] = min_sv; stats[6] = is_degenerate;
  }
  return 0;
}

}  // extern "C"
