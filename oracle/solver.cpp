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

namespace orc {
void accumulate_pose(const double* lf, int nl, const double* pf, int np, const double* x6, const double* Tbl,
                     double lidar_m, double w_tan, double a, double* H36, double* g6, double* cost, int threads);

// eval(x, &cost, H (n*n row-major) or null, g or null) -> false on numerical failure
using EvalFn = std::function<bool(const double*, double*, double*, double*)>;

struct DoglegSummary { int iterations = 0; int successful = 0; double initial_cost = 0, final_cost = 0; int termination = 0; };

DoglegSummary dogleg_minimize(int n, double* x_io, const EvalFn& eval, int max_num_iterations) {
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  const double min_relative_decrease = 1e-3, min_radius = 1e-32;
  const double min_diagonal = 1e-6, max_diagonal = 1e32;
  const double min_mu = 1e-8, max_mu = 1.0, mu_increase = 10.0;
  double radius = 1e4, mu = min_mu;
  bool reuse = false;
  DoglegSummary sum;

  std::vector<double> x(x_io, x_io + n), cand(n), H(n * n), g(n), Hc(n * n), gc(n);
  std::vector<double> scale(n, 1.0), Hs(n * n), gs(n), diag(n), grad(n), gn(n), step(n), delta(n);
  double alpha = 0, dogleg_step_norm = 0;
  double x_cost = 0;
  if (!eval(x.data(), &x_cost, H.data(), g.data())) { sum.termination = -1; return sum; }
  sum.initial_cost = x_cost;
  double minimum_cost = x_cost;
  // Jacobi scaling, computed once: 1 / (1 + ||J_col||)
  for (int i = 0; i < n; i++) scale[i] = 1.0 / (1.0 + std::sqrt(H[i * n + i]));
  auto apply_scale = [&]() {
    for (int i = 0; i < n; i++) {
      gs[i] = g[i] * scale[i];
      for (int j = 0; j < n; j++) Hs[i * n + j] = H[i * n + j] * scale[i] * scale[j];
    }
  };
  apply_scale();
  auto grad_max = [&]() {
    double m = 0;
    for (int i = 0; i < n; i++) m = std::max(m, std::fabs(g[i]));
    return m;
  };
  if (grad_max() <= gradient_tolerance) { sum.final_cost = x_cost; sum.termination = 1; return sum; }
  double x_norm = 0;
  for (int i = 0; i < n; i++) x_norm += x[i] * x[i];
  x_norm = std::sqrt(x_norm);
  int num_invalid = 0;

  for (int iter = 1; iter <= max_num_iterations; iter++) {
    sum.iterations = iter;
    // ---- DoglegStrategy::ComputeStep
    bool solve_ok = true;
    if (!reuse) {
      reuse = true;
      for (int i = 0; i < n; i++) diag[i] = std::sqrt(std::min(std::max(Hs[i * n + i], min_diagonal), max_diagonal));
      for (int i = 0; i < n; i++) grad[i] = gs[i] / diag[i];
      // Cauchy point: alpha = |grad|^2 / |J D^-1 grad|^2
      {
        std::vector<double> v(n);
        double gg = 0, vHv = 0;
        for (int i = 0; i < n; i++) { v[i] = grad[i] / diag[i]; gg += grad[i] * grad[i]; }
        for (int i = 0; i < n; i++) {
          double s = 0;
          for (int j = 0; j < n; j++) s += Hs[i * n + j] * v[j];
          vHv += v[i] * s;
        }
        alpha = gg / vHv;
      }
      // Gauss-Newton step with growing regularisation mu
      solve_ok = false;
      while (mu < max_mu) {
        std::vector<double> A(Hs);
        for (int i = 0; i < n; i++) A[i * n + i] += mu * diag[i] * diag[i];
        std::vector<double> y(n);
        bool ok = chol_solve(n, A.data(), gs.data(), y.data());
        if (ok)
          for (int i = 0; i < n; i++)
            if (!std::isfinite(y[i])) ok = false;
        if (!ok) { mu *= mu_increase; continue; }
        for (int i = 0; i < n; i++) gn[i] = -diag[i] * y[i];
        solve_ok = true;
        break;
      }
    }
    double model_cost_change = 0;
    bool step_valid = false;
    if (solve_ok) {
      // ---- ComputeTraditionalDoglegStep
      double gnorm = 0, gnn = 0;
      for (int i = 0; i < n; i++) { gnorm += grad[i] * grad[i]; gnn += gn[i] * gn[i]; }
      gnorm = std::sqrt(gnorm);
      gnn = std::sqrt(gnn);
      if (gnn <= radius) {
        for (int i = 0; i < n; i++) step[i] = gn[i];
        dogleg_step_norm = gnn;
      } else if (gnorm * alpha >= radius) {
        for (int i = 0; i < n; i++) step[i] = -(radius / gnorm) * grad[i];
        dogleg_step_norm = radius;
      } else {
        double b_dot_a = 0;
        for (int i = 0; i < n; i++) b_dot_a += grad[i] * gn[i];
        b_dot_a *= -alpha;
        double a_sq = std::pow(alpha * gnorm, 2.0);
        double bma = a_sq - 2 * b_dot_a + std::pow(gnn, 2.0);
        double c = b_dot_a - a_sq;
        double d = std::sqrt(c * c + bma * (std::pow(radius, 2.0) - a_sq));
        double beta = (c <= 0) ? (d - c) / bma : (radius * radius - a_sq) / (d + c);
        double sn = 0;
        for (int i = 0; i < n; i++) {
          step[i] = (-alpha * (1.0 - beta)) * grad[i] + beta * gn[i];
          sn += step[i] * step[i];
        }
        dogleg_step_norm = std::sqrt(sn);
      }
      for (int i = 0; i < n; i++) step[i] /= diag[i];
      // ---- model cost change = -(J s)^T (r + J s / 2) = -s^T g - s^T H s / 2
      double sg = 0, sHs = 0;
      for (int i = 0; i < n; i++) {
        double t = 0;
        for (int j = 0; j < n; j++) t += Hs[i * n + j] * step[j];
        sHs += step[i] * t;
        sg += step[i] * gs[i];
      }
      model_cost_change = -sg - 0.5 * sHs;
      step_valid = model_cost_change > 0.0;
    }
    if (!step_valid) {
      // HandleInvalidStep
      if (++num_invalid >= 5) { sum.termination = -2; break; }
      mu *= mu_increase;  // StepIsInvalid
      reuse = false;
      continue;
    }
    num_invalid = 0;
    double step_norm = 0;
    for (int i = 0; i < n; i++) {
      delta[i] = step[i] * scale[i];
      cand[i] = x[i] + delta[i];
      step_norm += delta[i] * delta[i];
    }
    step_norm = std::sqrt(step_norm);
    double cand_cost;
    if (!eval(cand.data(), &cand_cost, Hc.data(), gc.data()) || !std::isfinite(cand_cost))
      cand_cost = std::numeric_limits<double>::max();
    // ParameterToleranceReached
    if (step_norm <= parameter_tolerance * (x_norm + parameter_tolerance)) { sum.termination = 2; break; }
    // FunctionToleranceReached
    double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= function_tolerance * x_cost) { sum.termination = 3; break; }
    double relative_decrease = cost_change / model_cost_change;
    if (relative_decrease > min_relative_decrease) {
      // HandleSuccessfulStep
      x = cand;
      x_cost = cand_cost;
      H = Hc;
      g = gc;
      apply_scale();
      x_norm = 0;
      for (int i = 0; i < n; i++) x_norm += x[i] * x[i];
      x_norm = std::sqrt(x_norm);
      sum.successful++;
      if (relative_decrease < 0.25) radius *= 0.5;
      if (relative_decrease > 0.75) radius = std::max(radius, 3.0 * dogleg_step_norm);
      mu = std::max(min_mu, 2.0 * mu / mu_increase);
      reuse = false;
      if (x_cost < minimum_cost) {
        minimum_cost = x_cost;
        std::memcpy(x_io, x.data(), sizeof(double) * n);
      }
      if (grad_max() <= gradient_tolerance) { sum.termination = 1; break; }
    } else {
      radius *= 0.5;  // StepRejected
      reuse = true;
    }
    if (radius < min_radius) { sum.termination = 4; break; }
  }
  sum.final_cost = minimum_cost;
  return sum;
}
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
