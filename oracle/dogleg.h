// ORACLE (test infrastructure only — see oracle.h).
//   dogleg_minimize  <- ceres::Solve with the options of src/lio/Estimator.cpp:1425-1432:
//                      TrustRegionMinimizer + DoglegStrategy(TRADITIONAL_DOGLEG) + DENSE_SCHUR,
//                      max_num_iterations 10, Jacobi scaling, Ceres 2.1.0 defaults otherwise
//                      (initial radius 1e4, function_tolerance 1e-6, gradient_tolerance 1e-10,
//                       parameter_tolerance 1e-8, min_relative_decrease 1e-3, dogleg mu in
//                       [1e-8, 1] x10, radius *0.5 below 0.25 / max(r, 3|step|) above 0.75).
// Ceres is not in the tree or the container: the minimiser is restated from its published
// algorithm on the normal equations (H = J^T J, g = J^T r of the loss-corrected problem),
// which carry everything Ceres' step computation uses (DENSE_SCHUR solves the same linear
// system exactly, whatever the elimination order). One copy, shared by the oracle's solvers
// and by the ceres::Solve stand-in of the reference pin (oracle/ref/shim/ref_ceres.h).
#ifndef MMLOAM_ORACLE_DOGLEG_H
#define MMLOAM_ORACLE_DOGLEG_H
#include "oracle_math.h"
#include <cmath>
#include <cstring>
#include <functional>
#include <limits>
#include <vector>

namespace orc {
// eval(x, &cost, H (n*n row-major) or null, g or null) -> false on numerical failure
using EvalFn = std::function<bool(const double*, double*, double*, double*)>;

struct DoglegSummary { int iterations = 0; int successful = 0; double initial_cost = 0, final_cost = 0; int termination = 0; };

inline DoglegSummary dogleg_minimize(int n, double* x_io, const EvalFn& eval, int max_num_iterations) {
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
#endif
