// ORACLE (test infrastructure only — see oracle.h).
//   line residual       <- Cost_NavState_IMU_Line::operator()      include/utils/ceresfunc.h:412-440
//   plane-vec residual  <- Cost_NavState_IMU_Plan_Vec::operator()  include/utils/ceresfunc.h:533-555
//   Huber correction    <- ceres::HuberLoss + Corrector, mirrored by
//                          ResidualBlockInfo::Evaluate             include/utils/ceresfunc.h:33-63
// Ceres evaluates the Jacobians by automatic differentiation of those functors; the
// oracle uses the closed-form derivatives of the same expressions (checked against
// central finite differences in tests/test_oracle.py and against the reference's functors under
// dual-number autodiff in tests/test_ref_pin.py).
//
// Parameterisation (EST.cpp:937-950, 1227-1229): x = [t_wb (3), phi (3)], R_wb = Exp(phi),
// plain 6-vector update (no manifold). P_map = R_wb (R_bl p + P_bl) + t_wb (CF.h:418-423).
//
// The plane functor's sqrt_info = diag(1,w_t,w_t)/lidar_m * (V U^T)^T from
// JacobiSVD(e1 n^T) (EST.cpp:675-682). Cost, gradient and J^T J depend on it only through
// sqrt_info^T sqrt_info = (n n^T + w_t^2 (I - n n^T)) / lidar_m^2, so the oracle uses the
// canonical factor diag(1,w_t,w_t)/lidar_m * [n t1 t2]^T.
#include "oracle.h"
#include "oracle_math.h"
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

using namespace orc;

namespace orc {

struct PoseLin {
  double R[9];      // R_wb
  double t[3];
  double Jr[9];     // right Jacobian of SO(3) at phi
  double Rbl[9], Pbl[3];
};

static void right_jacobian(const double* phi, double* Jr) {
  double th2 = (phi[0] * phi[0] + phi[1] * phi[1]) + phi[2] * phi[2];
  double a, b;  // Jr = I - a [phi]x + b [phi]x^2
  if (th2 < 1e-12) {
    a = 0.5 - th2 / 24.0;
    b = 1.0 / 6.0 - th2 / 120.0;
  } else {
    double th = std::sqrt(th2);
    a = (1.0 - std::cos(th)) / th2;
    b = (th - std::sin(th)) / (th2 * th);
  }
  double K[9] = {0, -phi[2], phi[1], phi[2], 0, -phi[0], -phi[1], phi[0], 0};
  double K2[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += K[3 * r + k] * K[3 * k + c];
      K2[3 * r + c] = s;
    }
  for (int i = 0; i < 9; i++) Jr[i] = -a * K[i] + b * K2[i];
  Jr[0] += 1; Jr[4] += 1; Jr[8] += 1;
}

void make_pose(const double* x6, const double* Tbl16, PoseLin& L) {
  Quat q = so3_exp(x6 + 3);
  quat_to_R(q, L.R);
  L.t[0] = x6[0]; L.t[1] = x6[1]; L.t[2] = x6[2];
  right_jacobian(x6 + 3, L.Jr);
  // CF.h:405-408: qbl = Quaterniond(Tbl.topLeftCorner(3,3)).normalized()
  double Rm[9] = {Tbl16[0], Tbl16[1], Tbl16[2], Tbl16[4], Tbl16[5], Tbl16[6], Tbl16[8], Tbl16[9], Tbl16[10]};
  quat_to_R(quat_normalized(quat_from_R(Rm)), L.Rbl);
  L.Pbl[0] = Tbl16[3]; L.Pbl[1] = Tbl16[7]; L.Pbl[2] = Tbl16[11];
}

// P = R u + t and dP/dx (3x6) = [I | -R [u]x Jr]
static void map_point(const PoseLin& L, const double* p, double* P, double* JP) {
  double u[3];
  for (int r = 0; r < 3; r++) u[r] = L.Rbl[3 * r] * p[0] + L.Rbl[3 * r + 1] * p[1] + L.Rbl[3 * r + 2] * p[2] + L.Pbl[r];
  for (int r = 0; r < 3; r++) P[r] = L.R[3 * r] * u[0] + L.R[3 * r + 1] * u[1] + L.R[3 * r + 2] * u[2] + L.t[r];
  if (!JP) return;
  double ux[9] = {0, -u[2], u[1], u[2], 0, -u[0], -u[1], u[0], 0};
  double Rux[9], M[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += L.R[3 * r + k] * ux[3 * k + c];
      Rux[3 * r + c] = s;
    }
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += Rux[3 * r + k] * L.Jr[3 * k + c];
      M[3 * r + c] = -s;
    }
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      JP[6 * r + c] = (r == c) ? 1.0 : 0.0;
      JP[6 * r + 3 + c] = M[3 * r + c];
    }
}

// CF.h:425-437. feat = [p(3) a(3) b(3) ...]. Returns 1 residual, J 1x6.
void line_residual(const PoseLin& L, const double* f, double s_info, double* r, double* J) {
  const double *p = f, *a = f + 3, *b = f + 6;
  double P[3], JP[18];
  map_point(L, p, P, J ? JP : nullptr);
  double l12 = std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
  double c0 = (P[0] - a[0]) * (P[1] - b[1]) - (P[0] - b[0]) * (P[1] - a[1]);
  double c1 = (P[0] - a[0]) * (P[2] - b[2]) - (P[0] - b[0]) * (P[2] - a[2]);
  double c2 = (P[1] - a[1]) * (P[2] - b[2]) - (P[1] - b[1]) * (P[2] - a[2]);
  double a012 = std::sqrt(c0 * c0 + c1 * c1 + c2 * c2);
  double ld2 = a012 / l12;
  double PP = P[0] * P[0] + P[1] * P[1] + P[2] * P[2];
  double sq = std::sqrt(std::sqrt(PP));  // sqrt(||P||)
  double w = 1.0 - 0.9 * std::fabs(ld2) / sq;
  r[0] = s_info * w * ld2;
  if (!J) return;
  // true cross product c = (P-a)x(P-b) = (c2, -c1, c0); grad d = ((a-b) x c^) / l12
  double cx = c2, cy = -c1, cz = c0;
  double ch[3] = {cx / a012, cy / a012, cz / a012};
  double ab[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
  double gd[3] = {(ab[1] * ch[2] - ab[2] * ch[1]) / l12, (ab[2] * ch[0] - ab[0] * ch[2]) / l12,
                  (ab[0] * ch[1] - ab[1] * ch[0]) / l12};
  double k5 = 0.5 * ld2 / (PP * sq);  // 1/2 d ||P||^(-5/2)
  double gr[3];
  for (int k = 0; k < 3; k++) {
    double gw = -0.9 * (gd[k] / sq - k5 * P[k]);
    gr[k] = s_info * (w * gd[k] + ld2 * gw);
  }
  for (int c = 0; c < 6; c++) J[c] = gr[0] * JP[c] + gr[1] * JP[6 + c] + gr[2] * JP[12 + c];
}

void plane_basis(const double* n, double* t1, double* t2) {
  int k = 0;
  if (std::fabs(n[1]) < std::fabs(n[k])) k = 1;
  if (std::fabs(n[2]) < std::fabs(n[k])) k = 2;
  double e[3] = {0, 0, 0};
  e[k] = 1.0;
  double v[3] = {e[1] * n[2] - e[2] * n[1], e[2] * n[0] - e[0] * n[2], e[0] * n[1] - e[1] * n[0]};
  double nv = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  for (int i = 0; i < 3; i++) t1[i] = v[i] / nv;
  t2[0] = n[1] * t1[2] - n[2] * t1[1];
  t2[1] = n[2] * t1[0] - n[0] * t1[2];
  t2[2] = n[0] * t1[1] - n[1] * t1[0];
}

// CF.h:545-552. feat = [p(3) p_proj(3) n(3) ...]. 3 residuals, J 3x6 (row-major).
void plane_residual(const PoseLin& L, const double* f, double s_info, double w_tan, double* r, double* J) {
  const double *p = f, *pp = f + 3, *n_f32 = f + 6;
  // sqrt_info = info * (V U^T)^T with JacobiSVD(e1 n^T) (EST.cpp:675-682): the factor's direction is the unit
  // singular vector n / |n| in float64; the singular value |n| (1 +- 6e-8 for the float32 normal) is dropped.
  const double nn = std::sqrt((n_f32[0] * n_f32[0] + n_f32[1] * n_f32[1]) + n_f32[2] * n_f32[2]);
  const double n[3] = {n_f32[0] / nn, n_f32[1] / nn, n_f32[2] / nn};
  double P[3], JP[18];
  map_point(L, p, P, J ? JP : nullptr);
  double e[3] = {P[0] - pp[0], P[1] - pp[1], P[2] - pp[2]};
  double en = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
  double PP = P[0] * P[0] + P[1] * P[1] + P[2] * P[2];
  double sq = std::sqrt(std::sqrt(PP));
  double w = 1.0 - 0.9 * en / sq;
  double t1[3], t2[3];
  plane_basis(n, t1, t2);
  const double* B[3] = {n, t1, t2};
  double sc[3] = {s_info, s_info * w_tan, s_info * w_tan};
  for (int k = 0; k < 3; k++) r[k] = sc[k] * w * (B[k][0] * e[0] + B[k][1] * e[1] + B[k][2] * e[2]);
  if (!J) return;
  double k5 = 0.5 * en / (PP * sq);
  double gw[3];
  for (int k = 0; k < 3; k++) gw[k] = -0.9 * ((e[k] / en) / sq - k5 * P[k]);
  // A = w I + e gw^T ; rows_k = sc_k * B_k^T A
  for (int k = 0; k < 3; k++) {
    double be = B[k][0] * e[0] + B[k][1] * e[1] + B[k][2] * e[2];
    double row[3];
    for (int c = 0; c < 3; c++) row[c] = sc[k] * (w * B[k][c] + be * gw[c]);
    for (int c = 0; c < 6; c++) J[6 * k + c] = row[0] * JP[c] + row[1] * JP[6 + c] + row[2] * JP[12 + c];
  }
}

// ceres::HuberLoss(a) + Corrector with rho'' <= 0: scale residual and Jacobian by sqrt(rho').
// Returns rho(s).
inline double huber(double s, double a, double* sqrt_rho1) {
  if (a > 0 && s > a * a) {
    double rr = std::sqrt(s);
    *sqrt_rho1 = std::sqrt(a / rr);
    return 2 * a * rr - a * a;
  }
  *sqrt_rho1 = 1.0;
  return s;
}

struct Acc {
  double H[36], g[6], cost;
  Acc() { std::memset(this, 0, sizeof(*this)); }
  void add_rows(const double* r, const double* J, int nres, double a) {
    double s = 0;
    for (int k = 0; k < nres; k++) s += r[k] * r[k];
    double k1;
    double rho = huber(s, a, &k1);
    cost += 0.5 * rho;
    for (int k = 0; k < nres; k++) {
      double rk = k1 * r[k];
      double Jk[6];
      for (int c = 0; c < 6; c++) Jk[c] = k1 * J[6 * k + c];
      for (int i = 0; i < 6; i++) {
        g[i] += Jk[i] * rk;
        for (int j = 0; j < 6; j++) H[6 * i + j] += Jk[i] * Jk[j];
      }
    }
  }
  void merge(const Acc& o) {
    for (int i = 0; i < 36; i++) H[i] += o.H[i];
    for (int i = 0; i < 6; i++) g[i] += o.g[i];
    cost += o.cost;
  }
};

void accumulate_range(const PoseLin& L, const double* lf, int l0, int l1, const double* pf, int p0, int p1,
                      double s_info, double w_tan, double a, Acc& acc) {
  double r[3], J[18];
  for (int i = l0; i < l1; i++) {
    const double* f = lf + 12 * i;
    if (f[10] != 1.0) continue;  // only features with |error| > 1e-5 enter the problem
    line_residual(L, f, s_info, r, J);
    acc.add_rows(r, J, 1, a);
  }
  for (int i = p0; i < p1; i++) {
    const double* f = pf + 12 * i;
    if (f[10] != 1.0) continue;
    plane_residual(L, f, s_info, w_tan, r, J);
    acc.add_rows(r, J, 3, a);
  }
}

void accumulate_pose(const double* lf, int nl, const double* pf, int np, const double* x6, const double* Tbl,
                     double lidar_m, double w_tan, double a, double* H36, double* g6, double* cost, int threads) {
  PoseLin L;
  make_pose(x6, Tbl, L);
  double s_info = 1.0 / lidar_m;
  Acc total;
  if (threads <= 1) {
    accumulate_range(L, lf, 0, nl, pf, 0, np, s_info, w_tan, a, total);
  } else {
    std::vector<Acc> part(threads);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
      th.emplace_back([&, t]() {
        accumulate_range(L, lf, (int)((long)nl * t / threads), (int)((long)nl * (t + 1) / threads), pf,
                         (int)((long)np * t / threads), (int)((long)np * (t + 1) / threads), s_info, w_tan, a, part[t]);
      });
    for (auto& t : th) t.join();
    for (auto& p : part) total.merge(p);
  }
  std::memcpy(H36, total.H, sizeof(total.H));
  std::memcpy(g6, total.g, sizeof(total.g));
  *cost = total.cost;
}

}  // namespace orc

extern "C" {

int orc_accumulate(const double* lf, int nl, const double* pf, int np, const double* x6, const double* Tbl,
                   double w_tan, double huber_a, double* H36, double* g6, double* cost, int threads) {
  orc::accumulate_pose(lf, nl, pf, np, x6, Tbl, 1.5e-3, w_tan, huber_a, H36, g6, cost, threads);
  return 0;
}

int orc_residual(int kind, const double* feat12, const double* x6, const double* Tbl, double w_tan,
                 double* r3, double* J18) {
  orc::PoseLin L;
  orc::make_pose(x6, Tbl, L);
  double s_info = 1.0 / 1.5e-3;
  r3[0] = r3[1] = r3[2] = 0;
  if (J18) std::memset(J18, 0, sizeof(double) * 18);
  if (kind == 0) orc::line_residual(L, feat12, s_info, r3, J18);
  else orc::plane_residual(L, feat12, s_info, w_tan, r3, J18);
  return 0;
}

}  // extern "C"
