// mmloam_b200: the sliding-window solve on the device (SURVEY.md §8 f, F3: the small-system solver with IMU factors).
// Reference: Estimator::Estimate for windowSize < SLIDEWINDOWSIZE, src/lio/Estimator.cpp:1143-1581 — vector2double
// 937-950, IMU blocks 1235-1254 (Cost_NavState_PRV_Bias, include/utils/ceresfunc.h:321-393, under forward-mode
// differentiation like ceres::AutoDiffCostFunction), lidar blocks 1377-1418, ceres::Solve 1425-1432 (TrustRegionMinimizer
// with the traditional dogleg on the dense (15 W)-dimensional normal equations, Jacobi scaling, Ceres 2.1.0 defaults
// otherwise), double2vector 952-964, convergence test 1441-1450.
//
// One launch of k_solve_window = one outer iteration's whole trust-region loop, on ONE thread-block cluster of 16
// CTAs (16 SMs). Per evaluation:
//   * warps 0-7 of every CTA evaluate a slice of the lidar features of ONE frame (CTA rank mod W picks the frame):
//     residual, analytic Jacobian, Huber, 28 sums in registers, warp shuffles, then an all-gather of the per-CTA sums
//     through distributed shared memory (one cluster barrier);
//   * warps 8-10 evaluate the W-1 IMU factors meanwhile: one thread per Jacobian column runs the functor text on a
//     dual number that carries that column's derivative (30 columns per factor), then the whole CTA weights the
//     rows by sqrt_info and forms the factors' J^T J / J^T r straight into the dense system;
//   * every CTA then takes the SAME dogleg step on its own copy of the solver state (identical inputs, identical
//     code, fixed summation orders: identical steps, so nothing is broadcast): bulk vector / matrix work on all
//     warps, the Cholesky factorisation (left-looking, the right-hand side carried as an extra row) on warp 0.
// The tail (CTA 0) writes the states back, runs the convergence test, prepares the next association's transforms
// and, when the solve is over, publishes the result to mapped host memory and clears the WHILE condition.
#include "windowstate.cuh"
#include "smallmath.cuh"
#include "eststate.cuh"
#include "lidarfactor.cuh"
#include "imufactor.cuh"
#include <float.h>
#include <math.h>
#include <vector>

int mml_associate_launch(mml_ctx* ctx, int kind, const double* T16, float thres, const double* T_dev,
                         const float* thres_dev, const int* gate, const int* nq_dev, int cap);

namespace mml {

constexpr int kWLD = kWinN + 1;           // leading dimension (odd: conflict-free column walks in shared memory)
constexpr int kWLidarWarps = 8;
constexpr int kWLidarThreads = 32 * kWLidarWarps;
constexpr int kWThreads = kWLidarThreads + 96;   // + one thread per IMU Jacobian column (30 (W - 1) <= 90)
constexpr int kWWarps = kWThreads / 32;
constexpr int kWCluster = 16;

struct WinSolveArgs {
  WinDev* wd;
  const float4* f_line[kMaxWindow];       // by physical slot
  const float4* f_plane[kMaxWindow];
  const int* cnt[kMaxWindow];             // [n_corner, n_surf]
  const double* assoc_stats[kMaxWindow];
  double* host_out;                       // mapped: [16 W states][16 statistics at 64]
  volatile unsigned* host_seq;
  cudaGraphConditionalHandle cond;
  int use_cond;
};

struct WinShared {
  double Hn[kWinN * kWLD];                // normal equations of the newest evaluation
  double Hs[kWinN * kWLD];                // Jacobi-scaled normal equations of the current linearisation
  double A[(kWinN + 1) * kWLD];           // Hs + mu diag^2 (rows 0..n-1) and the right-hand side (row n); factor in place
  double x[kWinN], x_cand[kWinN], x_best[kWinN], gs[kWinN], gnew[kWinN], scale[kWinN], diag[kWinN], grad[kWinN], gn[kWinN],
      step[kWinN], tv[kWinN], tt[kWinN], ysol[kWinN];
  double raw[kMaxWindow - 1][31][15];     // unweighted IMU residual (column 30) and Jacobian columns
  double Jw[kMaxWindow - 1][31][15];      // weighted by sqrt_info
  double gather[2][kWCluster][28];
  double sred[kWLidarWarps][28];
  double tot[kMaxWindow][28];
  PoseLin L;
  mml_preint pre[kMaxWindow - 1];
  double gravity[3], Rbl_q[9], Pbl[3];
  double q_before[4], t_before[3];
  double cost, min_cost, cost_new, cost_imu, radius, mu, alpha, dogleg_norm, model_change, step_norm, x_norm;
  int gidx[kMaxWindow - 1][30];
  int lidx[kMaxWindow - 1][kWinN];
  int nnz;
  unsigned short nz[kWinN * (kWinN + 1) / 2];  // upper-triangle entries some factor touches: (row << 8) | column
  unsigned char tri[27 * 28 / 2][2];        // (row, column) of the idx-th entry of a lower triangle, row-major
  int pos[kWinN], unpos[kWinN];           // unknown g <-> frame-major position (the order the factorisation works in)
  int reuse, first, done, it, num_invalid, total_inner, evals, action, done_after, chol_ok;
#ifdef MML_WIN_DEVPROF
  long long prof[16], prof_t0;
#endif
};

#ifdef MML_WIN_DEVPROF
#define FTICK(k) if (tid == 0) { const long long t_ = clock64(); s.prof[k] += t_ - s.prof_t0; s.prof_t0 = t_; }
#else
#define FTICK(k)
#endif

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
// sum over i < n of a[i] * b[i], formed redundantly by every warp (same order in each: the result is uniform over the CTA)
__device__ __forceinline__ double wdot(const double* a, const double* b, int n, int lane) {
  double v = 0;
  for (int i = lane; i < n; i += 32) v += a[i] * b[i];
  return wsum(v);
}
// out = M v (n x n, leading dimension kWLD): warps stride over the rows
__device__ __forceinline__ void cta_matvec(const double* M, const double* v, double* out, int n, int warp, int lane) {
  for (int i = warp; i < n; i += kWWarps) {
    double t = 0;
    for (int j = lane; j < n; j += 32) t += M[i * kWLD + j] * v[j];
    t = wsum(t);
    if (lane == 0) out[i] = t;
  }
}

// Cholesky factorisation and solve of A y = b by the whole CTA. A holds the matrix in frame-major order
// ([P, log Q, V, bg, ba] of frame 0, then frame 1, ...) in rows 0..n-1 and the right-hand side as row n.
// In that order the normal equations are block tridiagonal with blocks of B = 15 (an IMU factor couples consecutive
// frames, a lidar factor one frame), so the factor has no entries outside the band: a column in frame block b only
// reaches the rows of blocks b and b + 1 and the right-hand side row.
// Right-looking, three columns per step: every thread factors the 3 x 3 diagonal block in registers (same inputs,
// same result, so the positive-definiteness test is uniform and nothing is broadcast), one thread per row solves
// the panel below it, one barrier, the trailing entries of the band take their rank-3 update spread over all
// threads, one barrier. What bounds a step is the dependent chain (three reciprocal square roots) and the two
// barriers, not the flops. The back substitution runs the same way from the last block up.
// On success ysol (caller's order) = A^-1 b and the function returns true (uniform over the CTA).
// Replaces chol_solve_n of the reference restatement (oracle/window.cpp), which factors the same matrix densely.
__device__ bool cta_chol_solve(WinShared& s, int n, int B, int tid) {
  double* A = s.A;
  for (int c0 = 0; c0 < n; c0 += 3) {
    const int bj = c0 / B;
    const int rend = min(n, B * (bj + 2));
    const int m = rend - (c0 + 3);  // panel rows below the diagonal block (band only); the right-hand side row is extra
    const double* D0 = A + c0 * kWLD + c0;
    __syncthreads();  // the previous step's update is complete
    const double d00 = D0[0], d10 = D0[kWLD], d11 = D0[kWLD + 1], d20 = D0[2 * kWLD], d21 = D0[2 * kWLD + 1], d22 = D0[2 * kWLD + 2];
    if (!(d00 > 0.0)) return false;
    const double i00 = rsqrt(d00);
    const double l10 = d10 * i00, l20 = d20 * i00;
    const double t11 = d11 - l10 * l10;
    if (!(t11 > 0.0)) return false;
    const double i11 = rsqrt(t11);
    const double l21 = (d21 - l20 * l10) * i11;
    const double t22 = d22 - l20 * l20 - l21 * l21;
    if (!(t22 > 0.0)) return false;
    const double i22 = rsqrt(t22);
    if (tid <= m) {
      const int r = tid < m ? c0 + 3 + tid : n;
      double* Ar = A + r * kWLD + c0;
      const double x0 = Ar[0] * i00;
      const double x1 = (Ar[1] - x0 * l10) * i11;
      const double x2 = (Ar[2] - x0 * l20 - x1 * l21) * i22;
      Ar[0] = x0; Ar[1] = x1; Ar[2] = x2;
    }
    __syncthreads();  // panel complete; every thread has read the diagonal block, which may now be overwritten
    if (tid == kWThreads - 1) {
      double* Dw = A + c0 * kWLD + c0;
      Dw[0] = d00 * i00;
      Dw[kWLD] = l10; Dw[kWLD + 1] = t11 * i11;
      Dw[2 * kWLD] = l20; Dw[2 * kWLD + 1] = l21; Dw[2 * kWLD + 2] = t22 * i22;
      s.tv[c0] = i00; s.tv[c0 + 1] = i11; s.tv[c0 + 2] = i22;  // 1 / L[j][j] for the back substitution
    }
    const int ntri = m * (m + 1) / 2;
    for (int idx = tid; idx < ntri + m; idx += kWThreads) {
      int r, c;
      if (idx < ntri) { r = c0 + 3 + s.tri[idx][0]; c = c0 + 3 + s.tri[idx][1]; }
      else { r = n; c = c0 + 3 + idx - ntri; }
      const double* Xr = A + r * kWLD + c0;
      const double* Xc = A + c * kWLD + c0;
      A[r * kWLD + c] -= (Xr[0] * Xc[0] + Xr[1] * Xc[1]) + Xr[2] * Xc[2];
    }
  }
  __syncthreads();
  // row n now holds L^-1 b; back substitution L^T y = (row n), three unknowns per step, last block first
  double* y = A + n * kWLD;
  for (int c0 = n - 3; c0 >= 0; c0 -= 3) {
    const double* D0 = A + c0 * kWLD + c0;
    const double x2 = y[c0 + 2] * s.tv[c0 + 2];
    const double x1 = (y[c0 + 1] - D0[2 * kWLD + 1] * x2) * s.tv[c0 + 1];
    const double x0 = (y[c0] - D0[kWLD] * x1 - D0[2 * kWLD] * x2) * s.tv[c0];
    const int bj = c0 / B, lo = bj > 0 ? B * (bj - 1) : 0;  // rows c0..c0+2 of the factor start at the previous block
    __syncthreads();
    const int i = lo + tid;
    if (i < c0) y[i] -= (D0[i - c0] * x0 + D0[kWLD + i - c0] * x1) + D0[2 * kWLD + i - c0] * x2;
    else if (tid == kWThreads - 1) { y[c0] = x0; y[c0 + 1] = x1; y[c0 + 2] = x2; }
    __syncthreads();
  }
  bool ok = true;
  for (int i = tid; i < n; i += kWThreads) {
    const double v = y[i];
    s.ysol[s.unpos[i]] = v;
    if (!isfinite(v)) ok = false;
  }
  return __syncthreads_and(ok) != 0;
}

// DoglegStrategy::ComputeStep + the model evaluation (DoglegN::compute_step of the host solver). All threads of the
// CTA call it; scalars that steer the control flow are formed redundantly per warp, so they are uniform.
__device__ bool win_compute_step(WinShared& s, int n, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  FTICK(2)
  if (!s.reuse) {
    for (int i = tid; i < n; i += kWThreads) {
      const double dg = sqrt(fmin(fmax(s.Hs[i * kWLD + i], 1e-6), 1e32));
      s.diag[i] = dg;
      s.grad[i] = s.gs[i] / dg;
    }
    __syncthreads();
    if (tid == 0) { s.reuse = 1; s.alpha = -1.0; }  // the Cauchy step length is formed on demand (below)
    for (;;) {
      const double mu = s.mu;
      if (!(mu < 1.0)) { __syncthreads(); return false; }
      for (int i = warp; i < n; i += kWWarps) {
        const int pi = s.pos[i] * kWLD;
        for (int j = lane; j < n; j += 32) {
          double v = s.Hs[i * kWLD + j];
          if (i == j) v += mu * s.diag[i] * s.diag[i];
          s.A[pi + s.pos[j]] = v;
        }
      }
      for (int j = tid; j < n; j += kWThreads) s.A[n * kWLD + s.pos[j]] = s.gs[j];
      __syncthreads();
      FTICK(3)
      const bool solved = cta_chol_solve(s, n, n < 15 ? n : 15, tid);
      FTICK(4)
      if (solved) break;
      __syncthreads();
      if (tid == 0) s.mu = mu * 10.0;
      __syncthreads();
    }
    for (int i = tid; i < n; i += kWThreads) s.gn[i] = -s.diag[i] * s.ysol[i];
    __syncthreads();
  }
  FTICK(6)
  const double gg = wdot(s.grad, s.grad, n, lane);
  const double gnorm = sqrt(gg);
  const double gnn = sqrt(wdot(s.gn, s.gn, n, lane));
  const double radius = s.radius;
  double alpha = s.alpha;
  if (!(gnn <= radius) && alpha < 0.0) {
    // alpha = |g|^2 / (g^T D^-1 H D^-1 g) of the current linearisation (unchanged while steps are rejected)
    for (int i = tid; i < n; i += kWThreads) s.tv[i] = s.grad[i] / s.diag[i];
    __syncthreads();
    cta_matvec(s.Hs, s.tv, s.tt, n, warp, lane);
    __syncthreads();
    alpha = gg / wdot(s.tv, s.tt, n, lane);
    __syncthreads();
    if (tid == 0) s.alpha = alpha;
  }
  double dogleg_norm;
  // the step, component i on lanes i and i + 32 of every warp (warp 0 stores it)
  double st0 = 0, st1 = 0;
  const int i0 = lane, i1 = lane + 32;
  if (gnn <= radius) {
    if (i0 < n) st0 = s.gn[i0];
    if (i1 < n) st1 = s.gn[i1];
    dogleg_norm = gnn;
  } else if (gnorm * alpha >= radius) {
    const double f = -(radius / gnorm);
    if (i0 < n) st0 = f * s.grad[i0];
    if (i1 < n) st1 = f * s.grad[i1];
    dogleg_norm = radius;
  } else {
    double b_dot_a = wdot(s.grad, s.gn, n, lane);
    b_dot_a *= -alpha;
    const double a_sq = (alpha * gnorm) * (alpha * gnorm);
    const double bma = a_sq - 2 * b_dot_a + gnn * gnn;
    const double c = b_dot_a - a_sq;
    const double d = sqrt(c * c + bma * (radius * radius - a_sq));
    const double beta = (c <= 0) ? (d - c) / bma : (radius * radius - a_sq) / (d + c);
    if (i0 < n) st0 = (-alpha * (1.0 - beta)) * s.grad[i0] + beta * s.gn[i0];
    if (i1 < n) st1 = (-alpha * (1.0 - beta)) * s.grad[i1] + beta * s.gn[i1];
    dogleg_norm = sqrt(wsum(st0 * st0 + st1 * st1));
  }
  if (i0 < n) st0 /= s.diag[i0];
  if (i1 < n) st1 /= s.diag[i1];
  if (warp == 0) {
    if (i0 < n) s.step[i0] = st0;
    if (i1 < n) s.step[i1] = st1;
  }
  __syncthreads();
  FTICK(7)
  cta_matvec(s.Hs, s.step, s.tt, n, warp, lane);
  __syncthreads();
  FTICK(8)
  const double sg = wdot(s.step, s.gs, n, lane), sHs = wdot(s.step, s.tt, n, lane);
  const double model_change = -sg - 0.5 * sHs;
  if (!(model_change > 0.0)) { __syncthreads(); return false; }
  double d0 = 0, d1 = 0;
  if (i0 < n) d0 = st0 * s.scale[i0];
  if (i1 < n) d1 = st1 * s.scale[i1];
  const double step_norm = sqrt(wsum(d0 * d0 + d1 * d1));
  if (warp == 0) {
    if (i0 < n) s.x_cand[i0] = s.x[i0] + d0;
    if (i1 < n) s.x_cand[i1] = s.x[i1] + d1;
    if (lane == 0) { s.model_change = model_change; s.step_norm = step_norm; s.dogleg_norm = dogleg_norm; }
  }
  __syncthreads();
  FTICK(9)
  return true;
}

// TrustRegionMinimizer: take the evaluation (cost_new, Hn, gnew) at the current evaluation point and move on
// (DoglegN::feed + advance of the host solver). Returns with s.done set or with the next evaluation point in x_cand.
__device__ void win_feed(WinShared& s, int n, int max_it, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
#ifdef MML_WIN_DEVPROF
  if (tid == 0) s.prof_t0 = clock64();
#endif
  double gm = 0;
  for (int i = lane; i < n; i += 32) gm = fmax(gm, fabs(s.gnew[i]));
  gm = wmax(gm);
  if (tid == 0) {
    const double c = s.cost_new;
    int action = 0, done_after = 0;  // 0 stop, 1 accept (or first evaluation), 2 reject
    if (s.first) {
      s.cost = c; s.min_cost = c;
      action = 1;
      done_after = (!isfinite(c) || gm <= 1e-10);
    } else {
      const double cand = isfinite(c) ? c : DBL_MAX;
      if (!(s.step_norm <= 1e-8 * (s.x_norm + 1e-8))) {
        const double cost_change = s.cost - cand;
        if (!(fabs(cost_change) <= 1e-6 * s.cost)) {
          const double rel = cost_change / s.model_change;
          if (rel > 1e-3) {
            action = 1;
            s.cost = cand;
            if (rel < 0.25) s.radius *= 0.5;
            if (rel > 0.75) s.radius = fmax(s.radius, 3.0 * s.dogleg_norm);
            s.mu = fmax(1e-8, 2.0 * s.mu / 10.0);
            s.reuse = 0;
            done_after = gm <= 1e-10;
          } else {
            action = 2;
            s.radius *= 0.5;
            s.reuse = 1;
          }
        }
      }
    }
    s.action = action;
    s.done_after = done_after;
    if (action == 0) s.done = 1;
  }
  __syncthreads();
  FTICK(0)
  const int action = s.action;
  if (action == 0) return;
  if (action == 1) {
    const bool first = s.first != 0;
    const bool better = first || s.cost < s.min_cost;
    if (first) {
      for (int i = tid; i < n; i += kWThreads) s.scale[i] = 1.0 / (1.0 + sqrt(s.Hn[i * kWLD + i]));
    } else {
      for (int i = tid; i < n; i += kWThreads) s.x[i] = s.x_cand[i];
    }
    __syncthreads();
    for (int i = warp; i < n; i += kWWarps) {
      const double si = s.scale[i];
      for (int j = lane; j < n; j += 32) s.Hs[i * kWLD + j] = s.Hn[i * kWLD + j] * si * s.scale[j];
    }
    for (int i = tid; i < n; i += kWThreads) s.gs[i] = s.gnew[i] * s.scale[i];
    const double xn = sqrt(wdot(s.x, s.x, n, lane));
    if (better) for (int i = tid; i < n; i += kWThreads) s.x_best[i] = s.x[i];
    if (tid == 0) {
      s.x_norm = xn;
      if (better) s.min_cost = s.cost;
      if (s.done_after) s.done = 1;
    }
  }
  __syncthreads();
  if (tid == 0) {
    if (!s.first && s.radius < 1e-32) s.done = 1;
    s.first = 0;
  }
  __syncthreads();
  FTICK(1)
  if (s.done) return;
  for (;;) {
    if (s.it >= max_it) {
      __syncthreads();
      if (tid == 0) s.done = 1;
      __syncthreads();
      return;
    }
    __syncthreads();
    if (tid == 0) { s.it++; s.total_inner++; }
    __syncthreads();
    if (win_compute_step(s, n, tid)) {
      if (tid == 0) s.num_invalid = 0;
      __syncthreads();
      return;
    }
    if (tid == 0) {
      if (++s.num_invalid >= 5) s.done = 1;
      else { s.mu *= 10.0; s.reuse = 0; }
    }
    __syncthreads();
    if (s.done) return;
  }
}

// T_wl = [Q exRbl, Q exPbl + P] of every frame (EST.cpp:1268-1270), stored by physical slot, and the thres_dist of
// the coming association (EST.cpp:1207, 1377-1381)
__device__ inline void win_prepare_assoc(WinDev* wd, int outer_it) {
  for (int f = 0; f < wd->W; f++) {
    const double* sf = wd->states[f];
    double Rq[9];
    quat_to_R(Quat{sf[3], sf[4], sf[5], sf[6]}, Rq);
    double* T = wd->T_wl[wd->slot_of[f]];
    for (int r = 0; r < 3; r++) {
      for (int k = 0; k < 3; k++)
        T[4 * r + k] = Rq[3 * r] * wd->Rbl_raw[k] + Rq[3 * r + 1] * wd->Rbl_raw[3 + k] + Rq[3 * r + 2] * wd->Rbl_raw[6 + k];
      T[4 * r + 3] = Rq[3 * r] * wd->Pbl[0] + Rq[3 * r + 1] * wd->Pbl[1] + Rq[3 * r + 2] * wd->Pbl[2] + sf[r];
    }
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
  }
  wd->thres = (float)wd->thres_sched[outer_it < 2 ? outer_it : 2];
}

__device__ inline void win_reset_control(WinDev* wd) {
  wd->n = (wd->W == 1) ? 6 : 15 * wd->W;  // a lone frame's velocity / bias block has no residual (Ceres drops it)
  wd->done_outer = 0; wd->outer_it = 0;
  wd->total_inner = 0; wd->evals = 0; wd->is_degenerate = 0; wd->n_line_last = 0; wd->n_plane_last = 0;
  wd->final_cost = 0; wd->min_sv = -1;
}

// per-call API: the host has uploaded the head of WinDev
__global__ void k_win_begin(WinDev* wd, int* const* cnt_by_slot) {
  if (threadIdx.x != 0) return;
  win_reset_control(wd);
  win_prepare_assoc(wd, 0);
  // slots that hold no frame of this window must not contribute queries
  bool used[kMaxWindow] = {false, false, false, false};
  for (int f = 0; f < wd->W; f++) used[wd->slot_of[f]] = true;
  for (int p = 0; p < kMaxWindow; p++)
    if (!used[p]) { cnt_by_slot[p][0] = 0; cnt_by_slot[p][1] = 0; }
}

// odometry loop: append the new frame (state predicted from the IMU, PE.cpp:811-820) and drop the oldest beyond
// the window (PE.cpp:830-832); the other frames' states are the ones the previous solve left on the device
__global__ void k_win_push(WinDev* wd, const WinPush* push, int* const* cnt_by_slot) {
  const int tid = threadIdx.x;
  __shared__ int W_new, drop;
  if (tid == 0) {
    drop = wd->W >= push->window ? 1 : 0;
    W_new = wd->W - drop + 1;
  }
  __syncthreads();
  const int W_old = wd->W;
  if (drop) {
    // shift frames 1.. to 0.. (sequential over frames, parallel inside a frame)
    for (int f = 0; f + 1 < W_old; f++) {
      for (int i = tid; i < 16; i += blockDim.x) wd->states[f][i] = wd->states[f + 1][i];
      const unsigned* src = reinterpret_cast<const unsigned*>(&wd->pre[f + 1]);
      unsigned* dst = reinterpret_cast<unsigned*>(&wd->pre[f]);
      for (int i = tid; i < (int)(sizeof(mml_preint) / 4); i += blockDim.x) dst[i] = src[i];
      __syncthreads();
      if (tid == 0) wd->slot_of[f] = wd->slot_of[f + 1];
      __syncthreads();
    }
  }
  const int fn = W_new - 1;
  for (int i = tid; i < 16; i += blockDim.x) wd->states[fn][i] = push->state[i];
  {
    const unsigned* src = reinterpret_cast<const unsigned*>(&push->pre);
    unsigned* dst = reinterpret_cast<unsigned*>(&wd->pre[fn]);
    for (int i = tid; i < (int)(sizeof(mml_preint) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  if (tid == 0) {
    wd->slot_of[fn] = push->slot;
    wd->W = W_new;
    wd->seq = push->seq;
    win_reset_control(wd);
    win_prepare_assoc(wd, 0);
    bool used[kMaxWindow] = {false, false, false, false};
    for (int f = 0; f < W_new; f++) used[wd->slot_of[f]] = true;
    for (int p = 0; p < kMaxWindow; p++)
      if (!used[p]) { cnt_by_slot[p][0] = 0; cnt_by_slot[p][1] = 0; }
  }
}

#ifdef MML_WIN_DEVPROF
#define WTICK(k) { const long long t1_ = clock64(); wtp[k] += t1_ - wt0; wt0 = t1_; }
#else
#define WTICK(k)
#endif

__global__ void __launch_bounds__(kWThreads, 1) k_solve_window(WinSolveArgs A) {
  WinDev* wd = A.wd;
  if (wd->done_outer) return;  // uniform over the cluster
  extern __shared__ __align__(16) unsigned char win_smem[];
  WinShared& s = *reinterpret_cast<WinShared*>(win_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned rank = cluster_ctarank();
  const int W = wd->W, n = wd->n, max_it = wd->max_inner;
  const int nf = W - 1;  // IMU factors
  // ---- set-up: pre-integrations and parameters into shared memory, vector2double (EST.cpp:937-950)
  for (int f = 0; f < nf; f++) {
    const unsigned* src = reinterpret_cast<const unsigned*>(&wd->pre[f + 1]);
    unsigned* dst = reinterpret_cast<unsigned*>(&s.pre[f]);
    for (int i = tid; i < (int)(sizeof(mml_preint) / 4); i += kWThreads) dst[i] = src[i];
  }
  if (tid < 3) { s.gravity[tid] = wd->gravity[tid]; s.Pbl[tid] = wd->Pbl[tid]; }
  if (tid < 9) s.Rbl_q[tid] = wd->Rbl_q[tid];
  if (tid >= 32 && tid < 32 + W) {
    const int f = tid - 32;
    const double* sf = wd->states[f];
    for (int k = 0; k < 3; k++) s.x[6 * f + k] = sf[k];
    so3_log(Quat{sf[3], sf[4], sf[5], sf[6]}, &s.x[6 * f + 3]);
    if (W > 1) for (int k = 0; k < 9; k++) s.x[6 * W + 9 * f + k] = sf[7 + k];
    if (f == W - 1) {
      for (int k = 0; k < 4; k++) s.q_before[k] = sf[3 + k];
      for (int k = 0; k < 3; k++) s.t_before[k] = sf[k];
    }
  }
  for (int c = tid; c < 30 * nf; c += kWThreads) {
    const int fi = c / 30, col = c - 30 * fi;
    const int off[4] = {6 * fi, 6 * W + 9 * fi, 6 * (fi + 1), 6 * W + 9 * (fi + 1)};
    s.gidx[fi][col] = col < 6 ? off[0] + col : col < 15 ? off[1] + col - 6 : col < 21 ? off[2] + col - 15 : off[3] + col - 21;
  }
  for (int c = tid; c < nf * kWinN; c += kWThreads) (&s.lidx[0][0])[c] = -1;
  for (int r = tid; r < 27; r += kWThreads)
    for (int c = 0; c <= r; c++) { s.tri[r * (r + 1) / 2 + c][0] = (unsigned char)r; s.tri[r * (r + 1) / 2 + c][1] = (unsigned char)c; }
  for (int g = tid; g < n; g += kWThreads) {
    int ps = g;
    if (W > 1) ps = g < 6 * W ? 15 * (g / 6) + g % 6 : 15 * ((g - 6 * W) / 9) + 6 + (g - 6 * W) % 9;
    s.pos[g] = ps;
    s.unpos[ps] = g;
  }
  if (tid == 0) {
    s.first = 1; s.done = 0; s.it = 0; s.radius = 1e4; s.mu = 1e-8; s.reuse = 0; s.num_invalid = 0;
    s.total_inner = 0; s.evals = 0; s.alpha = -1.0; s.cost_imu = 0.0; s.nnz = 0;
#ifdef MML_WIN_DEVPROF
    for (int k = 0; k < 16; k++) s.prof[k] = 0;
#endif
  }
  __syncthreads();
  for (int c = tid; c < 30 * nf; c += kWThreads) {
    const int fi = c / 30, col = c - 30 * fi;
    s.lidx[fi][s.gidx[fi][col]] = col;
  }
  __syncthreads();
  // entries of the normal equations that are re-formed by every evaluation: those an IMU factor couples, and the
  // pose blocks the lidar terms add to; everything else stays zero
  for (int i = tid; i < kWinN * kWLD; i += kWThreads) s.Hn[i] = 0.0;
  for (int i = warp; i < n; i += kWWarps)
    for (int j = i + lane; j < n; j += 32) {
      bool hit = i < 6 * W && j < 6 * W && i / 6 == j / 6;
      for (int fi = 0; fi < nf; fi++) hit = hit || (s.lidx[fi][i] >= 0 && s.lidx[fi][j] >= 0);
      if (hit) s.nz[atomicAdd(&s.nnz, 1)] = (unsigned short)((i << 8) | j);
    }
  // the frame whose lidar terms this CTA evaluates
  const int fr = (int)rank % W, sub = (int)rank / W, ncta_f = (kWCluster - fr + W - 1) / W;
  const int slot = wd->slot_of[fr];
  const float4* __restrict__ fl = A.f_line[slot];
  const float4* __restrict__ fp = A.f_plane[slot];
  const int n_line = A.cnt[slot][0], n_all = n_line + A.cnt[slot][1];
  const double s_info = 1.0 / wd->lidar_m, w_tan = wd->w_tan, ha = wd->huber_a;
  __syncthreads();
  int buf = 0;
#ifdef MML_WIN_DEVPROF
  long long wtp[8] = {0, 0, 0, 0, 0, 0, 0, 0}, wt0 = clock64();
#endif
  for (;;) {
    const double* xe = s.first ? s.x : s.x_cand;
    make_pose_split(xe + 6 * fr, s.Rbl_q, s.Pbl, s.L, tid);
    __syncthreads();
    WTICK(0)
    if (warp < kWLidarWarps) {
      double acc[28];
#pragma unroll
      for (int k = 0; k < 28; k++) acc[k] = 0.0;
      for (int i = sub * kWLidarThreads + tid; i < n_all; i += ncta_f * kWLidarThreads) {
        double p[3], a[3], b[3];
        if (i < n_line) {
          if (load_line(fl, i, p, a, b)) eval_line(s.L, p, a, b, s_info, ha, acc);
        } else {
          if (load_plane(fp, i - n_line, p, a, b)) eval_plane(s.L, p, a, b, s_info, w_tan, ha, acc);
        }
      }
      const double v = warp_reduce28(acc, lane);
      if (lane < 28) s.sred[warp][lane] = v;
    } else {
      // IMU factor fi = frames (fi, fi + 1): column `col` of its 15 x 30 Jacobian by forward-mode differentiation
      const int c = tid - kWLidarThreads;
      if (c < 30 * nf) {
        const int fi = c / 30, col = c - 30 * fi;
        Dual1 xd[30], rd[15];
#pragma unroll
        for (int k = 0; k < 30; k++) xd[k] = Dual1(xe[s.gidx[fi][k]], k == col ? 1.0 : 0.0);
        imu_residual<Dual1>(s.pre[fi], s.gravity, xd, xd + 6, xd + 15, xd + 21, rd);
#pragma unroll
        for (int i = 0; i < 15; i++) s.raw[fi][col][i] = rd[i].v;
        if (col == 0) {
#pragma unroll
          for (int i = 0; i < 15; i++) s.raw[fi][30][i] = rd[i].a;
        }
      }
    }
    WTICK(1)
    __syncthreads();
    WTICK(2)
    // all-gather of the lidar sums through distributed shared memory (published by the cluster barrier below)
    for (int t = tid; t < 28 * kWCluster; t += kWThreads) {
      const int r = t / 28, k = t - 28 * r;
      double v = 0;
#pragma unroll
      for (int w8 = 0; w8 < kWLidarWarps; w8++) v += s.sred[w8][k];
      st_dsmem_f64(&s.gather[buf][rank][k], (unsigned)r, v);
    }
    // rows weighted by sqrt_info (EST.cpp:1240-1242): Jw[fi][c][i] = sum_k sqrt_info[i][k] raw[fi][c][k]
    for (int item = tid; item < nf * 465; item += kWThreads) {
      const int fi = item / 465, rem = item - 465 * fi, c = rem / 15, i = rem - 15 * c;
      const double* S = s.pre[fi].sqrt_info + 15 * i;
      const double* rw = s.raw[fi][c];
      double t0 = 0, t1 = 0;
#pragma unroll
      for (int k = 0; k < 14; k += 2) { t0 += S[k] * rw[k]; t1 += S[k + 1] * rw[k + 1]; }
      s.Jw[fi][c][i] = (t0 + S[14] * rw[14]) + t1;
    }
    __syncthreads();
    WTICK(3)
    // the factors' J^T J and J^T r into the dense system (upper triangle formed, mirrored)
    for (int e = tid; e < s.nnz; e += kWThreads) {
      const int gi = s.nz[e] >> 8, gj = s.nz[e] & 255;
      double v = 0;
      for (int fi = 0; fi < nf; fi++) {
        const int li = s.lidx[fi][gi], lj = s.lidx[fi][gj];
        if (li < 0 || lj < 0) continue;
        const double* a = s.Jw[fi][li];
        const double* b = s.Jw[fi][lj];
        double t = 0;
#pragma unroll
        for (int k = 0; k < 15; k++) t += a[k] * b[k];
        v += t;
      }
      s.Hn[gi * kWLD + gj] = v;
      s.Hn[gj * kWLD + gi] = v;
    }
    for (int gi = tid; gi < n; gi += kWThreads) {
      double v = 0;
      for (int fi = 0; fi < nf; fi++) {
        const int li = s.lidx[fi][gi];
        if (li < 0) continue;
        const double* a = s.Jw[fi][li];
        const double* r = s.Jw[fi][30];
        double t = 0;
#pragma unroll
        for (int k = 0; k < 15; k++) t += a[k] * r[k];
        v += t;
      }
      s.gnew[gi] = v;
    }
    if (tid == kWThreads - 1) {
      double cimu = 0;
      for (int fi = 0; fi < nf; fi++)
        for (int k = 0; k < 15; k++) cimu += 0.5 * s.Jw[fi][30][k] * s.Jw[fi][30][k];
      s.cost_imu = cimu;
    }
    WTICK(4)
    cluster_sync_all();
    WTICK(5)
    // lidar blocks: sums over the CTAs of a frame in rank order
    if (tid < 28 * W) {
      const int f = tid / 28, k = tid - 28 * f;
      double v = 0;
      for (int r = f; r < kWCluster; r += W) v += s.gather[buf][r][k];
      s.tot[f][k] = v;
    }
    __syncthreads();
    if (tid < 28 * W) {
      const int f = tid / 28, k = tid - 28 * f;
      const double v = s.tot[f][k];
      if (k >= 1 && k < 7) s.gnew[6 * f + k - 1] += v;
      else if (k >= 7) {
        int i = 0, q = k - 7;
        while (q >= 6 - i) { q -= 6 - i; i++; }
        const int j = i + q;
        s.Hn[(6 * f + i) * kWLD + 6 * f + j] += v;
        if (j != i) s.Hn[(6 * f + j) * kWLD + 6 * f + i] += v;
      }
    } else if (tid == 28 * W) {
      double c = s.cost_imu;
      for (int f = 0; f < W; f++) c += s.tot[f][0];
      s.cost_new = c;
      s.evals++;
    }
    __syncthreads();
    WTICK(6)
    win_feed(s, n, max_it, tid);
    WTICK(7)
    if (s.done) break;
    buf ^= 1;
  }
#ifdef MML_WIN_DEVPROF
  if (rank == 0 && tid == 0)
    printf("  feed: decide=%lld apply=%lld advance=%lld buildA=%lld chol=%lld sync=%lld gn=%lld step=%lld matvec=%lld finish=%lld\n", s.prof[0], s.prof[1],
           s.prof[2], s.prof[3], s.prof[4], s.prof[5], s.prof[6], s.prof[7], s.prof[8], s.prof[9]);
  if (rank == 0 && (tid == 0 || tid == kWLidarThreads))
    printf("win solve tid %d: evals=%d pose=%lld eval=%lld wait=%lld weight=%lld assemble=%lld cluster=%lld totals=%lld feed=%lld cycles\n", tid,
           s.evals, wtp[0], wtp[1], wtp[2], wtp[3], wtp[4], wtp[5], wtp[6], wtp[7]);
#endif
  // every remote store was completed by the last cluster barrier and all CTAs leave the loop in the same iteration
  if (rank != 0) return;
  __syncthreads();
  // ---- double2vector (EST.cpp:952-964), convergence test (EST.cpp:1441-1450), next association
  if (tid < W) {
    const int f = tid;
    double* sf = wd->states[f];
    for (int k = 0; k < 3; k++) sf[k] = s.x_best[6 * f + k];
    const Quat q = so3_exp(&s.x_best[6 * f + 3]);
    sf[3] = q.w; sf[4] = q.x; sf[5] = q.y; sf[6] = q.z;
    if (W > 1) for (int k = 0; k < 9; k++) sf[7 + k] = s.x_best[6 * W + 9 * f + k];
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const int it = wd->outer_it;
    int degenerate = wd->is_degenerate;
    for (int f = 0; f < W; f++) {
      const double* as = A.assoc_stats[wd->slot_of[f]];
      const int* ints = reinterpret_cast<const int*>(as + 16);
      const double sv = as[15];   // localizability value left by the plane association's last CTA (EST.cpp:536-565)
      if (sv < 3.0) degenerate = 1;  // EST.cpp:771-775
      if (f == W - 1) { wd->n_line_last = ints[0]; wd->n_plane_last = ints[1]; wd->min_sv = sv; }
    }
    wd->is_degenerate = degenerate;
    wd->total_inner += s.total_inner;
    wd->evals += s.evals;
    wd->final_cost = s.min_cost;
    const double* sb = wd->states[W - 1];
    const Quat qb = {s.q_before[0], s.q_before[1], s.q_before[2], s.q_before[3]};
    const Quat dq = quat_mul(qb, Quat{sb[3], -sb[4], -sb[5], -sb[6]});
    const double deltaR = 2.0 * atan2(sqrt((dq.x * dq.x + dq.y * dq.y) + dq.z * dq.z), fabs(dq.w)) * 180.0 / 3.14159265358979323846;
    const double d0 = s.t_before[0] - sb[0], d1 = s.t_before[1] - sb[1], d2 = s.t_before[2] - sb[2];
    const double deltaT = sqrt((d0 * d0 + d1 * d1) + d2 * d2);
    const int done = ((deltaR < 0.05 && deltaT < 0.05) || (it + 1) >= wd->max_outer) ? 1 : 0;
    wd->outer_it = it + 1;
    if (!done) win_prepare_assoc(wd, it + 1);
    if (done && A.host_out) {
      for (int f = 0; f < W; f++)
        for (int k = 0; k < 16; k++) A.host_out[16 * f + k] = wd->states[f][k];
      double* st = A.host_out + 16 * kMaxWindow;
      st[0] = it + 1; st[1] = wd->total_inner; st[2] = wd->n_line_last; st[3] = wd->n_plane_last;
      st[4] = wd->final_cost; st[5] = wd->min_sv; st[6] = degenerate; st[7] = wd->evals; st[8] = W;
      __threadfence_system();
      *A.host_seq = wd->seq;
    }
    wd->done_outer = done;
    if (A.use_cond) cudaGraphSetConditional(A.cond, done ? 0u : 1u);
  }
}

}  // namespace mml

using namespace mml;

static int win_lend(mml_ctx* c, WinSlot& s) {
  std::swap(c->q_corner, s.q_corner); std::swap(c->q_surf, s.q_surf);
  std::swap(c->f_line, s.f_line); std::swap(c->f_plane, s.f_plane);
  std::swap(c->assoc_stats, s.assoc_stats);
  std::swap(c->assoc_part[0], s.assoc_part[0]); std::swap(c->assoc_part[1], s.assoc_part[1]);
  return 0;
}

// association of every physical slot (2 kinds x kMaxWindow kernels side by side on captured streams), then the solve
static int capture_outer_iteration(mml_ctx* c, WindowState* w, int cap, cudaGraphConditionalHandle cond, int use_cond) {
  cudaStream_t st = c->stream;
  WinDev* wd = w->dev.as<WinDev>();
  if (cudaEventRecord(w->fork, st) != cudaSuccess) return MML_ERR_CUDA;
  int rc = MML_OK;
  for (int p = 0; p < kMaxWindow && rc == MML_OK; p++) {
    WinSlot& s = w->slot[p];
    win_lend(c, s);
    c->has_perm[0] = c->has_perm[1] = false;
    for (int kind = 1; kind >= 0 && rc == MML_OK; kind--) {
      cudaStream_t fs = w->fstream[p][kind];
      if (cudaStreamWaitEvent(fs, w->fork, 0) != cudaSuccess) { rc = MML_ERR_CUDA; break; }
      c->stream = fs;
      rc = mml_associate_launch(c, kind, nullptr, 0.f, wd->T_wl[p], &wd->thres, &wd->done_outer, s.cnt.as<int>() + kind, cap);
      c->stream = st;
      if (rc == MML_OK && cudaEventRecord(w->fev[p][kind], fs) != cudaSuccess) rc = MML_ERR_CUDA;
    }
    win_lend(c, s);
  }
  if (rc != MML_OK) return rc;
  for (int p = 0; p < kMaxWindow; p++)
    for (int kind = 0; kind < 2; kind++)
      if (cudaStreamWaitEvent(st, w->fev[p][kind], 0) != cudaSuccess) return MML_ERR_CUDA;
  WinSolveArgs SA;
  memset(&SA, 0, sizeof(SA));
  SA.wd = wd;
  for (int p = 0; p < kMaxWindow; p++) {
    SA.f_line[p] = w->slot[p].f_line.as<float4>();
    SA.f_plane[p] = w->slot[p].f_plane.as<float4>();
    SA.cnt[p] = w->slot[p].cnt.as<int>();
    SA.assoc_stats[p] = w->slot[p].assoc_stats.as<double>();
  }
  SA.host_out = w->mapped_dev;
  SA.host_seq = reinterpret_cast<volatile unsigned*>(w->mapped_dev + kMapDoubles);
  SA.cond = cond;
  SA.use_cond = use_cond;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(kWCluster);
  cfg.blockDim = dim3(kWThreads);
  cfg.dynamicSmemBytes = sizeof(WinShared);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kWCluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_solve_window, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return MML_ERR_CUDA;
    if (cudaFuncSetAttribute(k_solve_window, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WinShared)) != cudaSuccess) return MML_ERR_CUDA;
    attr_set = true;
  }
  const bool ok = cudaLaunchKernelEx(&cfg, k_solve_window, SA) == cudaSuccess;
  MML_LAUNCHED(c);
  return ok ? MML_OK : MML_ERR_CUDA;
}

static int* const* slot_cnt_table(mml_ctx* c, WindowState* w) {
  // device table of the slots' count pointers, behind WinDev
  return reinterpret_cast<int* const*>(reinterpret_cast<char*>(w->dev.p) + sizeof(WinDev));
}

// reserve the slots' buffers for `cap` queries per kind and build (or reuse) the graph of the window's solve:
// two outer iterations as plain kernel nodes (gated by done_outer) and a WHILE node holding the same iteration for
// solves that need more (EST.cpp:1211: at most max_outer)
int mml_window_solve_graph(mml_ctx* c, WindowState* w, int cap) {
  cudaStream_t st = c->stream;
  if (cap < 1) cap = 1;
  MML_CUDA(c, w->dev.reserve(sizeof(WinDev) + sizeof(int*) * kMaxWindow + 64));
  MML_CUDA(c, w->push_dev.reserve(sizeof(WinPush) + 64));
  const int grid_max = 4 * kNumSMs;
  for (int p = 0; p < kMaxWindow; p++) {
    WinSlot& s = w->slot[p];
    if (s.cap < cap) s.cap = cap;
    MML_CUDA(c, s.q_corner.reserve(sizeof(float4) * (size_t)(s.cap + 1)));
    MML_CUDA(c, s.q_surf.reserve(sizeof(float4) * (size_t)(s.cap + 1)));
    MML_CUDA(c, s.f_line.reserve(sizeof(float4) * 3 * (size_t)s.cap));
    MML_CUDA(c, s.f_plane.reserve(sizeof(float4) * 3 * (size_t)s.cap));
    MML_CUDA(c, s.assoc_stats.reserve(512));
    const size_t part = sizeof(double) * 8 * (size_t)(std::max(div_up(s.cap, 128), grid_max) + 1) + 64;
    MML_CUDA(c, s.assoc_part[0].reserve(part));
    MML_CUDA(c, s.assoc_part[1].reserve(part));
    MML_CUDA(c, s.cnt.reserve(64));
  }
  long long key = 1469598103934665603ll;
  auto mix = [&](long long v) { key = (key ^ v) * 1099511628211ll; };
  mix(cap); mix((long long)(size_t)w->dev.p); mix((long long)(size_t)w->mapped_dev);
  for (int p = 0; p < kMaxWindow; p++) {
    WinSlot& s = w->slot[p];
    mix(s.cap);
    mix((long long)(size_t)s.q_corner.p); mix((long long)(size_t)s.q_surf.p); mix((long long)(size_t)s.f_line.p);
    mix((long long)(size_t)s.f_plane.p); mix((long long)(size_t)s.assoc_stats.p); mix((long long)(size_t)s.cnt.p);
    mix((long long)(size_t)s.assoc_part[0].p); mix((long long)(size_t)s.assoc_part[1].p);
  }
  for (int k = 0; k < 4; k++) {
    const GridMap& M = c->maps[k];
    mix(M.valid); mix(M.coarse); mix((long long)(size_t)M.pts2.p); mix((long long)(size_t)M.cell_start2.p);
    mix((long long)(size_t)M.pts.p); mix((long long)(size_t)M.cell_start.p); mix(M.m); mix(M.ncell);
    mix(M.dim[0]); mix(M.dim[1]); mix(M.dim[2]); mix((long long)(M.cell * 1e6f));
    mix(M.cube_lo[0]); mix(M.cube_lo[1]); mix(M.cube_lo[2]);
    mix((long long)(M.org_d[0] * 1e6)); mix((long long)(M.org_d[1] * 1e6)); mix((long long)(M.org_d[2] * 1e6));
    mix(M.cen[0]); mix(M.cen[1]); mix(M.cen[2]);
  }
  if (w->graph && w->graph_key == key) return MML_OK;
  if (w->graph) { cudaGraphExecDestroy(w->graph); w->graph = nullptr; }
  // the table of count pointers the begin / push kernels use to silence unused slots
  {
    int* tab[kMaxWindow];
    for (int p = 0; p < kMaxWindow; p++) tab[p] = w->slot[p].cnt.as<int>();
    MML_CUDA(c, cudaMemcpyAsync(reinterpret_cast<char*>(w->dev.p) + sizeof(WinDev), tab, sizeof(tab), cudaMemcpyHostToDevice, st));
    MML_CUDA(c, cudaStreamSynchronize(st));
  }
  const long long launches_before = c->launches;
  cudaGraph_t graph = nullptr;
  MML_CUDA(c, cudaGraphCreate(&graph, 0));
  cudaGraphConditionalHandle cond;
  MML_CUDA(c, cudaGraphConditionalHandleCreate(&cond, graph, 0, cudaGraphCondAssignDefault));
  MML_CUDA(c, cudaStreamBeginCaptureToGraph(st, graph, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  int rc = capture_outer_iteration(c, w, cap, cond, 1);
  if (rc == MML_OK) rc = capture_outer_iteration(c, w, cap, cond, 1);
  std::vector<cudaGraphNode_t> leaves;
  {
    cudaStreamCaptureStatus status;
    const cudaGraphNode_t* deps = nullptr;
    size_t n_deps = 0;
    if (cudaStreamGetCaptureInfo(st, &status, nullptr, nullptr, &deps, &n_deps) == cudaSuccess && deps) leaves.assign(deps, deps + n_deps);
  }
  cudaGraph_t same = nullptr;
  cudaError_t ce = cudaStreamEndCapture(st, &same);
  const long long per_iter = (c->launches - launches_before) / 2;
  if (rc == MML_OK && ce == cudaSuccess && leaves.empty()) rc = mml_fail(c, MML_ERR_CUDA, "window graph: capture left no leaf node");
  if (rc == MML_OK && ce == cudaSuccess) {
    cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
    np.conditional.handle = cond;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t node;
    ce = cudaGraphAddNode(&node, graph, leaves.data(), leaves.size(), &np);
    if (ce == cudaSuccess) {
      cudaGraph_t body = np.conditional.phGraph_out[0];
      ce = cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
      if (ce == cudaSuccess) {
        rc = capture_outer_iteration(c, w, cap, cond, 1);
        ce = cudaStreamEndCapture(st, nullptr);
      }
    }
  }
  c->launches = launches_before;
  w->graph_launches = per_iter;
  if (rc != MML_OK) { cudaGraphDestroy(graph); return rc; }
  MML_CUDA(c, ce);
  MML_CUDA(c, cudaGraphInstantiate(&w->graph, graph, 0));
  cudaGraphDestroy(graph);
  w->graph_key = key;
  return MML_OK;
}

int mml_window_begin_launch(mml_ctx* c, WindowState* w) {
  k_win_begin<<<1, 32, 0, c->stream>>>(w->dev.as<WinDev>(), slot_cnt_table(c, w));
  MML_LAUNCHED(c);
  MML_CUDA(c, cudaGetLastError());
  return MML_OK;
}

int mml_window_push_launch(mml_ctx* c, WindowState* w, const WinPush* push_dev) {
  k_win_push<<<1, 128, 0, c->stream>>>(w->dev.as<WinDev>(), push_dev, slot_cnt_table(c, w));
  MML_LAUNCHED(c);
  MML_CUDA(c, cudaGetLastError());
  return MML_OK;
}
