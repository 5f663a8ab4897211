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
//     warps; the banded Cholesky factorisation (right-looking, three columns per step, the right-hand side carried as
//     an extra row) with warp 0 on the dependent chain and warps 1-10 on panel and update (cta_chol_solve).
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
int mml_grid_table_sync(mml_ctx* ctx);

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

// The solver works in FRAME-MAJOR order ([P, log Q, V, bg, ba] of frame 0, then frame 1, ...): position p of the
// system <-> unknown unpos[p] of the parameter vector x ([P, log Q] x W, then [V, bg, ba] x W, EST.cpp:937-950).
// In that order the normal equations are block tridiagonal with blocks of 15 (an IMU factor couples consecutive
// frames, a lidar factor one frame).
struct WinShared {
  double Hn[kWinN * kWLD];                // normal equations of the newest evaluation
  double Hs[kWinN * kWLD];                // Jacobi-scaled normal equations of the current linearisation
  double A[(kWinN + 1) * kWLD];           // Hs + mu diag^2 (rows 0..n-1) and the right-hand side (row n); factor in place
  double x[kWinN], x_cand[kWinN], x_best[kWinN];                 // parameter order
  double gs[kWinN], gnew[kWinN], scale[kWinN], diag[kWinN], grad[kWinN], gn[kWinN], step[kWinN], tv[kWinN], tt[kWinN],
      ysol[kWinN];                                               // frame-major order
  double raw[kMaxWindow - 1][31][15];     // unweighted IMU residual (column 30) and Jacobian columns
  double Jw[kMaxWindow - 1][31][15];      // weighted by sqrt_info
  double gather[2][kWCluster][28];
  double sred[kWLidarWarps][28];
  double tot[kMaxWindow][28];
  PoseLin L;
  mml_preint pre[kMaxWindow - 1];
  double gravity[3], Rbl_q[9], Pbl[3];
  double q_before[4], t_before[3];
  double cost_new, cost_imu, min_cost_out;
  int gidx[kMaxWindow - 1][30];           // factor column -> unknown (parameter order)
  int lidx[kMaxWindow - 1][kWinN];        // position -> factor column, or -1
  double fac[2][8];                         // factor of a step's 3 x 3 diagonal block: 1 / L[j][j] (3), L[1][0], L[2][0], L[2][1], positive-definite flag
                                            // (two copies by step parity: warp 0 publishes the next step's while the other warps still read this one's)
  double X[28][3];                          // the step's solved panel rows (list order: band rows, then the right-hand side)
  int chol_ok, ntile;                       // (chol_ok: unused since the factor's flag travels with the published block)
  unsigned short tl[(kWinN / 3) * (kWinN / 3 + 1) / 2];  // 3 x 3 tiles of the upper triangle some factor touches: (tile row << 8) | tile column
  unsigned char tri[27 * 28 / 2][2];        // (row, column) of the idx-th entry of a lower triangle, row-major
  int pos[kWinN], unpos[kWinN];
  int total_inner_out, evals_out;
};

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
// Reductions over the n <= 64 components of a vector are formed by EVERY warp (components lane and lane + 32, then
// a butterfly): the same operations in the same order everywhere, so the result is uniform over the CTA - and over
// the cluster - without a broadcast or a barrier.
__device__ __forceinline__ double wdot2(const double* a, const double* b, int n, int lane) {
  double v = lane < n ? a[lane] * b[lane] : 0.0;
  if (lane + 32 < n) v += a[lane + 32] * b[lane + 32];
  return wsum(v);
}
// out = M v (n x n, leading dimension kWLD): warps stride over the rows
__device__ __noinline__ void cta_matvec(const double* M, const double* v, double* out, int n, int warp, int lane) {
  const double v0 = lane < n ? v[lane] : 0.0, v1 = lane + 32 < n ? v[lane + 32] : 0.0;
#pragma unroll 1
  for (int i = warp; i < n; i += kWWarps) {
    double t = lane < n ? M[i * kWLD + lane] * v0 : 0.0;
    if (lane + 32 < n) t += M[i * kWLD + lane + 32] * v1;
    t = wsum(t);
    if (lane == 0) out[i] = t;
  }
}

// Cholesky factorisation and solve of A y = b by the CTA. A holds the matrix (frame-major order) in rows 0..n-1 and
// the right-hand side as row n. The factor has no entries outside the band of the block-tridiagonal system: a column
// in frame block b only reaches the rows of blocks b and b + 1 and the right-hand side row.
// Right-looking, three columns per step, two roles:
//   warp 0 owns the dependent chain. It holds the factor of the step's 3 x 3 diagonal block in registers (every lane
//     the same values), publishes it, and after the step's barrier forms the NEXT diagonal block - its three panel rows
//     solved in registers, this step's update applied - and factors it through the leading minors (three independent
//     reciprocal square roots), while
//   warps 1-10 solve the panel rows against the published block (one thread per row, solved rows into a side buffer),
//     meet at a named barrier of their own and apply the rank-3 update to the band's trailing entries.
// One full barrier per step. What bounds a step is warp 0's chain (~25 dependent float64 operations), not the flops.
// The back substitution runs on warp 0 alone, three unknowns per step through the inverted diagonal blocks, the
// solution in registers (components lane and lane + 32) and exchanged by shuffles: no barrier on the way.
// On success ysol = A^-1 b and the function returns true (uniform over the CTA).
// Replaces chol_solve_n of the reference restatement, which factors the same matrix densely.
#if defined(MML_WIN_DEVPROF) && MML_WIN_DEVPROF >= 2
__device__ long long g_cholprof[8];
#define CTICK(k) if (tid == 0 && blockIdx.x == 0) { const long long t_ = clock64(); g_cholprof[k] += t_ - ct0; ct0 = t_; }
#else
#define CTICK(k)
#endif
// 1 / sqrt(x) for a normal positive x: the hardware seed (rsqrt.approx.f64, relative error < 2^-22) and one third-order
// correction, the sequence rsqrt() itself runs, without its branch to the special-case path: three calls in a row
// overlap instead of queueing behind each other's branch. Non-positive input gives NaN / inf (callers test the sign).
__device__ __forceinline__ double rsqrt_normal(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(x, -(y * y), 1.0);
  return fma(fma(e, 0.375, 0.5), y * e, y);
}

__device__ __noinline__ bool cta_chol_solve(WinShared& s, int n, int B, int tid) {
  double* A = s.A;
#if defined(MML_WIN_DEVPROF) && MML_WIN_DEVPROF >= 2
  long long ct0 = clock64();
#endif
  int par = 0;
  int bnext = B;  // first column of the frame block after the one c0 lies in (no division in the loop)
  // the 3 x 3 diagonal block of the current step, factored one step AHEAD by warp 0 (registers, every lane the same values)
  double d10 = 0, d20 = 0, i00 = 0, i11 = 0, i22 = 0, l21 = 0;
  double q[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  bool pd = true;
  // Factor of a 3 x 3 block through its leading minors: pivots a00, det2 / a00, det3 / det2, so the three reciprocal
  // square roots (of a00, det2, det3) are independent instead of waiting for each other through two reciprocals
  auto factor3 = [&](double a00, double a10, double a11, double a20, double a21, double a22) {
    d10 = a10; d20 = a20;
    const double c21 = a00 * a21 - a20 * a10;
    const double c20 = a11 * a20 - a10 * a21;
    const double det2 = a00 * a11 - a10 * a10;
    const double det3 = a22 * det2 - (a21 * c21 + a20 * c20);
    pd = a00 > 0.0 && det2 > 0.0 && det3 > 0.0;
    const double r0 = rsqrt_normal(a00), r1 = rsqrt_normal(det2), r2 = rsqrt_normal(det3);
    i00 = r0;                  // 1 / L00
    i11 = (a00 * r0) * r1;     // 1 / L11 = sqrt(a00 / det2)
    i22 = (det2 * r1) * r2;    // 1 / L22 = sqrt(det2 / det3)
    l21 = c21 * (r0 * r1);     // L21 = (c21 / a00) / L11
  };
  __syncthreads();  // the caller's A is complete
  if (tid < 32) factor3(A[0], A[kWLD], A[kWLD + 1], A[2 * kWLD], A[2 * kWLD + 1], A[2 * kWLD + 2]);
#pragma unroll 1
  for (int c0 = 0; c0 < n; c0 += 3) {
    if (c0 >= bnext) bnext += B;
    const int rend = min(n, bnext + B);
    const int m = rend - (c0 + 3);  // panel rows below the diagonal block (band only); the right-hand side row is extra
    CTICK(0)
    const double l10 = d10 * i00, l20 = d20 * i00;  // (warp 0 only: the other warps hold zeros)
    if (tid == 31) {
      // warp 0 publishes the block's factor: into the matrix, and the six numbers a panel row needs
      double* D0 = A + c0 * kWLD + c0;
      double* fac = s.fac[par];
      fac[0] = i00; fac[1] = i11; fac[2] = i22; fac[3] = l10; fac[4] = l20; fac[5] = l21; fac[6] = pd ? 1.0 : 0.0;
      D0[kWLD] = l10; D0[2 * kWLD] = l20; D0[2 * kWLD + 1] = l21;  // (the diagonal of the factor lives in tv as reciprocals)
      s.tv[c0] = i00; s.tv[c0 + 1] = i11; s.tv[c0 + 2] = i22;  // 1 / L[j][j] for the back substitution
    }
    CTICK(2)
    __syncthreads();  // factor published, previous update complete
    CTICK(3)
    const double* fac = s.fac[par];
    if (fac[6] == 0.0) return false;  // not positive definite (uniform: the flag was written before the barrier)
    par ^= 1;
    if (tid < 32) {
      // warp 0: the next diagonal block. Its three panel rows are solved here in registers (the other warps solve them
      // too, into the side buffer), then the block takes this step's update and is factored: the chain a step waits for
      if (tid == 0 && c0 > 0) {  // the previous step's solved rows of this block, in place (read by the back substitution only)
        double* Xp = A + c0 * kWLD + c0 - 3;
        Xp[0] = q[0]; Xp[1] = q[1]; Xp[2] = q[2];
        Xp[kWLD] = q[3]; Xp[kWLD + 1] = q[4]; Xp[kWLD + 2] = q[5];
        Xp[2 * kWLD] = q[6]; Xp[2 * kWLD + 1] = q[7]; Xp[2 * kWLD + 2] = q[8];
      }
      if (c0 + 3 < n) {
        const double* X = A + (c0 + 3) * kWLD + c0;        // rows of the next block in this step's columns, not yet solved
        const double* Dn = A + (c0 + 3) * kWLD + c0 + 3;   // its diagonal block before the update
        const double x00 = X[0] * i00, x10 = X[kWLD] * i00, x20 = X[2 * kWLD] * i00;
        const double x01 = (X[1] - x00 * l10) * i11, x11 = (X[kWLD + 1] - x10 * l10) * i11, x21 = (X[2 * kWLD + 1] - x20 * l10) * i11;
        const double x02 = (X[2] - x00 * l20 - x01 * l21) * i22, x12 = (X[kWLD + 2] - x10 * l20 - x11 * l21) * i22,
                     x22 = (X[2 * kWLD + 2] - x20 * l20 - x21 * l21) * i22;
        factor3(Dn[0] - ((x00 * x00 + x01 * x01) + x02 * x02), Dn[kWLD] - ((x10 * x00 + x11 * x01) + x12 * x02),
                Dn[kWLD + 1] - ((x10 * x10 + x11 * x11) + x12 * x12), Dn[2 * kWLD] - ((x20 * x00 + x21 * x01) + x22 * x02),
                Dn[2 * kWLD + 1] - ((x20 * x10 + x21 * x11) + x22 * x12), Dn[2 * kWLD + 2] - ((x20 * x20 + x21 * x21) + x22 * x22));
        // those rows of the factor go into the matrix after the NEXT barrier: the panel threads are still reading them raw
        q[0] = x00; q[1] = x01; q[2] = x02; q[3] = x10; q[4] = x11; q[5] = x12; q[6] = x20; q[7] = x21; q[8] = x22;
      }
    } else {
      // the other warps: one thread per panel row (forward substitution against the published block; solved rows into
      // the side buffer, and in place except the three rows warp 0 is reading), then the rank-3 update of the band's
      // trailing entries, the next diagonal block excepted (the first six entries of the triangle)
      const int t = tid - 32;
      if (t <= m) {
        const int r = t < m ? c0 + 3 + t : n;
        double* Ar = A + r * kWLD + c0;
        const double f0 = fac[0], f1 = fac[1], f2 = fac[2], g10 = fac[3], g20 = fac[4], g21 = fac[5];
        const double x0 = Ar[0] * f0;
        const double x1 = (Ar[1] - x0 * g10) * f1;
        const double x2 = (Ar[2] - x0 * g20 - x1 * g21) * f2;
        s.X[t][0] = x0; s.X[t][1] = x1; s.X[t][2] = x2;
        if (t >= 3 || t == m) { Ar[0] = x0; Ar[1] = x1; Ar[2] = x2; }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(kWThreads - 32) : "memory");  // panel complete (warps 1 .. 10)
      const int ntri = m * (m + 1) / 2;
#pragma unroll 1
      for (int idx = 6 + t; idx < ntri + m; idx += kWThreads - 32) {
        int i, j, r;
        if (idx < ntri) { i = s.tri[idx][0]; j = s.tri[idx][1]; r = c0 + 3 + i; }
        else { i = m; j = idx - ntri; r = n; }
        const double* Xr = s.X[i];
        const double* Xc = s.X[j];
        A[r * kWLD + c0 + 3 + j] -= (Xr[0] * Xc[0] + Xr[1] * Xc[1]) + Xr[2] * Xc[2];
      }
    }
    CTICK(4)
  }
  __syncthreads();
  CTICK(5)
  // row n now holds L^-1 b; back substitution L^T y = (row n), three unknowns per step, last block first
  bool ok = true;
  if (tid < 32) {
    const int lane = tid;
    const double* yrow = A + n * kWLD;
    double y0 = lane < n ? yrow[lane] : 0.0, y1 = lane + 32 < n ? yrow[lane + 32] : 0.0;
    int bstart = n;  // first column of the frame block of c0 (walked down without a division)
#pragma unroll 1
    for (int c0 = n - 3; c0 >= 0; c0 -= 3) {
      const double* D0 = A + c0 * kWLD + c0;
      while (c0 < bstart) bstart -= B;
      const int lo = bstart >= B ? bstart - B : 0;  // rows c0..c0+2 of the factor start at the previous block
      // this lane's entries of those rows and the block's own entries: loads that do not wait for the chain below
      const bool u0 = lane >= lo && lane < c0, u1 = lane + 32 >= lo && lane + 32 < c0;
      const double a00 = u0 ? D0[lane - c0] : 0.0, a01 = u0 ? D0[kWLD + lane - c0] : 0.0, a02 = u0 ? D0[2 * kWLD + lane - c0] : 0.0;
      const double a10 = u1 ? D0[lane + 32 - c0] : 0.0, a11 = u1 ? D0[kWLD + lane + 32 - c0] : 0.0, a12 = u1 ? D0[2 * kWLD + lane + 32 - c0] : 0.0;
      const double L10 = D0[kWLD], L20 = D0[2 * kWLD], L21 = D0[2 * kWLD + 1];
      const double t0 = s.tv[c0], t1 = s.tv[c0 + 1], t2 = s.tv[c0 + 2];
      const double v2 = __shfl_sync(0xffffffffu, ((c0 + 2) >> 5) ? y1 : y0, (c0 + 2) & 31);
      const double v1 = __shfl_sync(0xffffffffu, ((c0 + 1) >> 5) ? y1 : y0, (c0 + 1) & 31);
      const double v0 = __shfl_sync(0xffffffffu, (c0 >> 5) ? y1 : y0, c0 & 31);
      const double x2 = v2 * t2;
      const double x1 = (v1 - L21 * x2) * t1;
      const double x0 = (v0 - L10 * x1 - L20 * x2) * t0;
      y0 -= (a00 * x0 + a01 * x1) + a02 * x2;
      y1 -= (a10 * x0 + a11 * x1) + a12 * x2;
      if (lane == (c0 & 31)) { if (c0 >> 5) y1 = x0; else y0 = x0; }
      if (lane == ((c0 + 1) & 31)) { if ((c0 + 1) >> 5) y1 = x1; else y0 = x1; }
      if (lane == ((c0 + 2) & 31)) { if ((c0 + 2) >> 5) y1 = x2; else y0 = x2; }
    }
    if (lane < n) { s.ysol[lane] = y0; if (!isfinite(y0)) ok = false; }
    if (lane + 32 < n) { s.ysol[lane + 32] = y1; if (!isfinite(y1)) ok = false; }
  }
  CTICK(6)
  const bool res_ = __syncthreads_and(ok) != 0;
  CTICK(7)
  return res_;
}

// The dogleg step outside the Gauss-Newton branch (the Gauss-Newton step leaves the trust region): the Cauchy step
// length, the interpolated step and the model change through products with Hs. DoglegStrategy::ComputeStep of
// Ceres 2.1.0 as restated by DoglegN::compute_step. Rare: the radius starts at 1e4.
// In: grad, gn, gs, diag, Hs. Out: step (scaled space, divided by diag), *model_change; returns the dogleg norm.
__device__ __noinline__ double win_dogleg_branch(WinShared& s, int n, int tid, double gg, double gnorm, double gnn, double radius,
                                                 double* alpha_io, double* model_change) {
  const int lane = tid & 31, warp = tid >> 5;
  double alpha = *alpha_io;
  if (alpha < 0.0) {
    // alpha = |g|^2 / (g^T D^-1 H D^-1 g) of the current linearisation (unchanged while steps are rejected)
    if (tid < n) s.tv[tid] = s.grad[tid] / s.diag[tid];
    __syncthreads();
    cta_matvec(s.Hs, s.tv, s.tt, n, warp, lane);
    __syncthreads();
    alpha = gg / wdot2(s.tv, s.tt, n, lane);
    *alpha_io = alpha;
  }
  double dogleg_norm;
  double st0 = 0, st1 = 0;
  const int i0 = lane, i1 = lane + 32;
  if (gnorm * alpha >= radius) {
    const double f = -(radius / gnorm);
    if (i0 < n) st0 = f * s.grad[i0];
    if (i1 < n) st1 = f * s.grad[i1];
    dogleg_norm = radius;
  } else {
    double b_dot_a = wdot2(s.grad, s.gn, n, lane);
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
  __syncthreads();  // every warp has formed its copy of the step before warp 0 stores it (tv / tt are reused below)
  if (warp == 0) {
    if (i0 < n) s.step[i0] = st0 / s.diag[i0];
    if (i1 < n) s.step[i1] = st1 / s.diag[i1];
  }
  __syncthreads();
  cta_matvec(s.Hs, s.step, s.tt, n, warp, lane);
  __syncthreads();
  const double sg = wdot2(s.step, s.gs, n, lane), sHs = wdot2(s.step, s.tt, n, lane);
  *model_change = -sg - 0.5 * sHs;
  return dogleg_norm;
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
// -DMML_WIN_TIMELINE: %globaltimer stamps of the odometry loop's critical stream, accumulated on the device
// [0] solve kernels (busy), [1] k_win_push, [2] push end -> first solve start (graph launch + association),
// [3] solve end -> next solve start (association), [4] last solve end -> next push start (host turn-around + scan kernels),
// [5] scans, [6] solve launches that did work
#ifdef MML_WIN_TIMELINE
__device__ unsigned long long g_wtl[8], g_wtl_push_end, g_wtl_solve_end, g_wtl_t0;
__device__ __forceinline__ unsigned long long wtl_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
extern "C" int mml_debug_win_timeline(unsigned long long* out8) {
  const int rc = cudaMemcpyFromSymbol(out8, g_wtl, sizeof(unsigned long long) * 8) == cudaSuccess ? 0 : -3;
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(g_wtl, z, sizeof(z));
  return rc;
}
#endif
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
#ifdef MML_WIN_TIMELINE
  if (tid == 0) {
    g_wtl_t0 = wtl_now();
    if (g_wtl_solve_end) g_wtl[4] += g_wtl_t0 - g_wtl_solve_end;
    g_wtl[5]++;
  }
#endif
  if (tid == 0) {
    drop = wd->W >= push->window ? 1 : 0;
    W_new = wd->W - drop + 1;
  }
  __syncthreads();
  const int W_old = wd->W;
  if (drop) {
    // frames 1.. move to 0..: thread i carries word i of every frame down in frame order, so no two threads touch the
    // same word and the shift needs no barrier
    constexpr int kPreWords = (int)(sizeof(mml_preint) / 4);
    for (int i = tid; i < kPreWords; i += blockDim.x)
      for (int f = 0; f + 1 < W_old; f++)
        reinterpret_cast<unsigned*>(&wd->pre[f])[i] = reinterpret_cast<const unsigned*>(&wd->pre[f + 1])[i];
    if (tid < 16)
      for (int f = 0; f + 1 < W_old; f++) wd->states[f][tid] = wd->states[f + 1][tid];
    if (tid == 32)
      for (int f = 0; f + 1 < W_old; f++) wd->slot_of[f] = wd->slot_of[f + 1];
    __syncthreads();
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
#ifdef MML_WIN_TIMELINE
    g_wtl_push_end = wtl_now();
    g_wtl[1] += g_wtl_push_end - g_wtl_t0;
#endif
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
#ifdef MML_WIN_TIMELINE
  unsigned long long wtl_start = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    wtl_start = wtl_now();
    if (wd->outer_it == 0) g_wtl[2] += wtl_start - g_wtl_push_end;
    else g_wtl[3] += wtl_start - g_wtl_solve_end;
    g_wtl[6]++;
  }
#endif
  extern __shared__ __align__(16) unsigned char win_smem[];
  WinShared& s = *reinterpret_cast<WinShared*>(win_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned rank = cluster_ctarank();
  const int W = wd->W, n = wd->n, max_it = wd->max_inner;
  const int nf = W - 1;            // IMU factors
  const int B = n < 15 ? n : 15;   // frame block of the system
  // ---- set-up: pre-integrations and parameters into shared memory, vector2double (EST.cpp:937-950)
  for (int f = 0; f < nf; f++) {
    const unsigned* src = reinterpret_cast<const unsigned*>(&wd->pre[f + 1]);
    unsigned* dst = reinterpret_cast<unsigned*>(&s.pre[f]);
#pragma unroll 1
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
  if (tid < 30 * nf) {
    const int fi = tid / 30, col = tid - 30 * fi;
    const int off[4] = {6 * fi, 6 * W + 9 * fi, 6 * (fi + 1), 6 * W + 9 * (fi + 1)};
    s.gidx[fi][col] = col < 6 ? off[0] + col : col < 15 ? off[1] + col - 6 : col < 21 ? off[2] + col - 15 : off[3] + col - 21;
  }
#pragma unroll 1
  for (int c = tid; c < nf * kWinN; c += kWThreads) (&s.lidx[0][0])[c] = -1;
  if (tid < 27)
    for (int c = 0; c <= tid; c++) { s.tri[tid * (tid + 1) / 2 + c][0] = (unsigned char)tid; s.tri[tid * (tid + 1) / 2 + c][1] = (unsigned char)c; }
  if (tid < n) {
    const int g = tid;
    int ps = g;
    if (W > 1) ps = g < 6 * W ? 15 * (g / 6) + g % 6 : 15 * ((g - 6 * W) / 9) + 6 + (g - 6 * W) % 9;
    s.pos[g] = ps;
    s.unpos[ps] = g;
  }
  if (tid == 0) { s.cost_imu = 0.0; s.ntile = 0; }
#pragma unroll 1
  for (int i = tid; i < kWinN * kWLD; i += kWThreads) s.Hn[i] = 0.0;
  __syncthreads();
  if (tid < 30 * nf) {
    const int fi = tid / 30, col = tid - 30 * fi;
    s.lidx[fi][s.pos[s.gidx[fi][col]]] = col;
  }
  __syncthreads();
  // entries of the normal equations that are re-formed by every evaluation, as 3 x 3 tiles (the parameters come in
  // triples: P, log Q, V, bg, ba of a frame occupy five consecutive tiles): those an IMU factor couples, and the pose
  // tiles the lidar terms add to; everything else stays zero
  {
    const int nt = n / 3;
#pragma unroll 1
    for (int t = tid; t < nt * nt; t += kWThreads) {
      const int ti = t / nt, tj = t - nt * ti;
      if (tj < ti) continue;
      const int i = 3 * ti, j = 3 * tj;
      bool hit = i / B == j / B && i % B < 6 && j % B < 6;
      for (int fi = 0; fi < nf; fi++) hit = hit || (s.lidx[fi][i] >= 0 && s.lidx[fi][j] >= 0);
      if (hit) s.tl[atomicAdd(&s.ntile, 1)] = (unsigned short)((ti << 8) | tj);
    }
  }
  // the frame whose lidar terms this CTA evaluates
  const int fr = (int)rank % W, sub = (int)rank / W, ncta_f = (kWCluster - fr + W - 1) / W;
  const int slot = wd->slot_of[fr];
  const float4* __restrict__ fl = A.f_line[slot];
  const float4* __restrict__ fp = A.f_plane[slot];
  const int n_line = A.cnt[slot][0], n_all = n_line + A.cnt[slot][1];
  const double s_info = 1.0 / wd->lidar_m, w_tan = wd->w_tan, ha = wd->huber_a;
  __syncthreads();
  // trust-region state (Ceres 2.1 TrustRegionMinimizer + DoglegStrategy, as restated by the test oracle): every
  // thread carries its own copy in registers and all copies take the same values
  double cost = 0, min_cost = 0, radius = 1e4, mu = 1e-8, mu_built = -1.0, alpha = -1.0, dogleg_norm = 0, model_change = 0, step_norm = 0, x_norm = 0;
  int first = 1, reuse = 0, it = 0, num_invalid = 0, evals = 0;
  int buf = 0;
#ifdef MML_WIN_DEVPROF
  long long wtp[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, wt0 = clock64();
#endif
#pragma unroll 1
  for (;;) {
    const double* xe = first ? s.x : s.x_cand;
    if (warp < kWLidarWarps) {
      // the pose of this CTA's frame is formed by threads 0, 32, 33 and only the lidar warps wait for it (named
      // barrier 1): the IMU warps start on their factors at once, theirs is the longer evaluation
      make_pose_split(xe + 6 * fr, s.Rbl_q, s.Pbl, s.L, tid);
      asm volatile("bar.sync 1, %0;" ::"n"(kWLidarThreads) : "memory");
      WTICK(0)
      double acc[28];
#pragma unroll
      for (int k = 0; k < 28; k++) acc[k] = 0.0;
#pragma unroll 1
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
#pragma unroll 1
    for (int t = tid; t < 28 * kWCluster; t += kWThreads) {
      const int r = t / 28, k = t - 28 * r;
      double v = 0;
#pragma unroll
      for (int w8 = 0; w8 < kWLidarWarps; w8++) v += s.sred[w8][k];
      st_dsmem_f64(&s.gather[buf][rank][k], (unsigned)r, v);
    }
    // rows weighted by sqrt_info (EST.cpp:1240-1242): Jw[fi][c][i] = sum_k sqrt_info[i][k] raw[fi][c][k]
#pragma unroll 1
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
    // the factors' J^T J and J^T r into the dense system, frame-major positions: one thread per 3 x 3 tile of the upper
    // triangle (six columns of Jw read for nine entries), mirrored; J^T r and the IMU cost on threads from the other end
    if (tid < s.ntile) {
      const int ti = s.tl[tid] >> 8, tj = s.tl[tid] & 255;
      double v00 = 0, v01 = 0, v02 = 0, v10 = 0, v11 = 0, v12 = 0, v20 = 0, v21 = 0, v22 = 0;
      for (int fi = 0; fi < nf; fi++) {
        const int li = s.lidx[fi][3 * ti], lj = s.lidx[fi][3 * tj];
        if (li < 0 || lj < 0) continue;
        const double* a = s.Jw[fi][li];
        const double* b = s.Jw[fi][lj];
#pragma unroll
        for (int k = 0; k < 15; k++) {
          const double a0 = a[k], a1 = a[15 + k], a2 = a[30 + k], b0 = b[k], b1 = b[15 + k], b2 = b[30 + k];
          v00 += a0 * b0; v01 += a0 * b1; v02 += a0 * b2;
          v10 += a1 * b0; v11 += a1 * b1; v12 += a1 * b2;
          v20 += a2 * b0; v21 += a2 * b1; v22 += a2 * b2;
        }
      }
      double* Hu = s.Hn + (3 * ti) * kWLD + 3 * tj;   // tile (ti, tj)
      double* Hl = s.Hn + (3 * tj) * kWLD + 3 * ti;   // its mirror image
      Hu[0] = v00; Hu[1] = v01; Hu[2] = v02;
      Hu[kWLD] = v10; Hu[kWLD + 1] = v11; Hu[kWLD + 2] = v12;
      Hu[2 * kWLD] = v20; Hu[2 * kWLD + 1] = v21; Hu[2 * kWLD + 2] = v22;
      if (ti != tj) {
        Hl[0] = v00; Hl[1] = v10; Hl[2] = v20;
        Hl[kWLD] = v01; Hl[kWLD + 1] = v11; Hl[kWLD + 2] = v21;
        Hl[2 * kWLD] = v02; Hl[2 * kWLD + 1] = v12; Hl[2 * kWLD + 2] = v22;
      }
    }
    const int gt = kWThreads - 2 - tid;
    if (gt >= 0 && gt < n) {
      const int gi = gt;
      double v = 0;
      for (int fi = 0; fi < nf; fi++) {
        const int li = s.lidx[fi][gi];
        if (li < 0) continue;
        const double* a = s.Jw[fi][li];
        const double* r = s.Jw[fi][30];
        double t0 = 0, t1 = 0;
#pragma unroll
        for (int k = 0; k < 14; k += 2) { t0 += a[k] * r[k]; t1 += a[k + 1] * r[k + 1]; }
        v += (t0 + a[14] * r[14]) + t1;
      }
      s.gnew[gi] = v;
    } else if (tid == kWThreads - 1) {
      double cimu = 0;
      for (int fi = 0; fi < nf; fi++)
#pragma unroll 1
        for (int k = 0; k < 15; k++) cimu += 0.5 * s.Jw[fi][30][k] * s.Jw[fi][30][k];
      s.cost_imu = cimu;
    }
    WTICK(4)
    cluster_sync_all();
    WTICK(5)
    // lidar blocks: sums over the CTAs of a frame in rank order, added to the frame's pose block
    if (tid < 28 * W) {
      const int f = tid / 28, k = tid - 28 * f;
      double v = 0;
#pragma unroll 1
      for (int r = f; r < kWCluster; r += W) v += s.gather[buf][r][k];
      const int p0 = B * f;  // position of the frame's pose block
      if (k == 0) s.tot[f][0] = v;
      else if (k < 7) s.gnew[p0 + k - 1] += v;
      else {
        int i = 0, q = k - 7;
        while (q >= 6 - i) { q -= 6 - i; i++; }
        const int j = i + q;
        s.Hn[(p0 + i) * kWLD + p0 + j] += v;
        if (j != i) s.Hn[(p0 + j) * kWLD + p0 + i] += v;
      }
    }
    __syncthreads();
    WTICK(6)
    evals++;
    // ---- TrustRegionMinimizer: take the evaluation (cost, Hn, gnew) at the evaluation point and move on
    double c_new = s.cost_imu;
    for (int f = 0; f < W; f++) c_new += s.tot[f][0];
    double gm = fmax(lane < n ? fabs(s.gnew[lane]) : 0.0, lane + 32 < n ? fabs(s.gnew[lane + 32]) : 0.0);
    gm = wmax(gm);
    int action = 0, done_after = 0;  // 0 stop, 1 accept (or first evaluation), 2 reject
    if (first) {
      cost = c_new; min_cost = c_new;
      action = 1;
      done_after = (!isfinite(c_new) || gm <= 1e-10);
    } else {
      const double cand = isfinite(c_new) ? c_new : DBL_MAX;
      if (!(step_norm <= 1e-8 * (x_norm + 1e-8))) {
        const double cost_change = cost - cand;
        if (!(fabs(cost_change) <= 1e-6 * cost)) {
          const double rel = cost_change / model_change;
          if (rel > 1e-3) {
            action = 1;
            cost = cand;
            if (rel < 0.25) radius *= 0.5;
            if (rel > 0.75) radius = fmax(radius, 3.0 * dogleg_norm);
            mu = fmax(1e-8, 2.0 * mu / 10.0);
            reuse = 0;
            done_after = gm <= 1e-10;
          } else {
            action = 2;
            radius *= 0.5;
            reuse = 1;
          }
        }
      }
    }
    if (action == 0) break;
    if (action == 1) {
      if (first) {
        if (tid < n) s.scale[tid] = 1.0 / (1.0 + sqrt(s.Hn[tid * kWLD + tid]));
        __syncthreads();
      }
      // the new linearisation: Jacobi-scaled copy, and the regularised copy + right-hand side the factorisation works on
      // (the diagonal's square root and division are kept out of the row loop: one lane per row would stall its warp)
#pragma unroll 1
      for (int i = warp; i < n; i += kWWarps) {
        const double si = s.scale[i];
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int j = lane + 32 * h;
          if (j < n && j != i) {
            const double v = s.Hn[i * kWLD + j] * si * s.scale[j];
            s.Hs[i * kWLD + j] = v;
            s.A[i * kWLD + j] = v;
          }
        }
      }
      mu_built = mu;
      const bool better = first || cost < min_cost;
      if (tid < n) {
        const int i = tid;
        const double si = s.scale[i];
        const double v = s.Hn[i * kWLD + i] * si * si;
        const double dg = sqrt(fmin(fmax(v, 1e-6), 1e32));
        const double g = s.gnew[i] * si;
        s.Hs[i * kWLD + i] = v;
        s.diag[i] = dg;
        s.gs[i] = g;
        s.grad[i] = g / dg;
        s.A[n * kWLD + i] = g;
        s.A[i * kWLD + i] = v + mu * dg * dg;
        const double xv = first ? s.x[tid] : s.x_cand[tid];
        s.x[tid] = xv;
        if (better) s.x_best[tid] = xv;
      }
      if (better) min_cost = cost;
      __syncthreads();
      x_norm = sqrt(wdot2(s.x, s.x, n, lane));
      if (done_after) break;
    }
    WTICK(8)
    if (!first && radius < 1e-32) break;
    first = 0;
    // ---- next step (with Ceres' handling of invalid steps)
    bool stop = false;
#pragma unroll 1
    for (;;) {
      if (it >= max_it) { stop = true; break; }
      it++;
      bool valid = true;
      if (!reuse) {
        reuse = 1;
        alpha = -1.0;  // the Cauchy step length is formed on demand
#pragma unroll 1
        for (;;) {
          if (!(mu < 1.0)) { valid = false; break; }
          if (mu_built != mu) {  // retry with a larger regularisation: rebuild the working copy from Hs
            __syncthreads();
#pragma unroll 1
            for (int i = warp; i < n; i += kWWarps)
#pragma unroll
              for (int h = 0; h < 2; h++) {
                const int j = lane + 32 * h;
                if (j < n) s.A[i * kWLD + j] = s.Hs[i * kWLD + j] + (i == j ? mu * s.diag[i] * s.diag[i] : 0.0);
              }
            if (tid < n) s.A[n * kWLD + tid] = s.gs[tid];
            mu_built = mu;
          }
          WTICK(9)
          const bool solved = cta_chol_solve(s, n, B, tid);
          WTICK(10)
          mu_built = -1.0;  // the working copy now holds the factor
          if (solved) break;
          mu *= 10.0;
        }
        if (valid) {
          if (tid < n) s.gn[tid] = -s.diag[tid] * s.ysol[tid];
          __syncthreads();
        }
      }
      if (valid) {
        const double gg = wdot2(s.grad, s.grad, n, lane);
        const double gnorm = sqrt(gg);
        const double gnn = sqrt(wdot2(s.gn, s.gn, n, lane));
        double st0 = 0, st1 = 0;  // the step in scaled space, components lane and lane + 32
        if (gnn <= radius) {
          // Gauss-Newton branch: step = gn / diag = -y with (Hs + mu D^2) y = gs, so Hs step = -gs + mu D^2 y and the
          // model change -step.gs - step.Hs.step / 2 = (y.gs + mu |D y|^2) / 2 needs no product with Hs
          if (lane < n) st0 = s.gn[lane] / s.diag[lane];
          if (lane + 32 < n) st1 = s.gn[lane + 32] / s.diag[lane + 32];
          dogleg_norm = gnn;
          const double ygs = -wsum((lane < n ? st0 * s.gs[lane] : 0.0) + (lane + 32 < n ? st1 * s.gs[lane + 32] : 0.0));
          model_change = 0.5 * (ygs + mu * (gnn * gnn));
        } else {
          dogleg_norm = win_dogleg_branch(s, n, tid, gg, gnorm, gnn, radius, &alpha, &model_change);
          if (lane < n) st0 = s.step[lane];
          if (lane + 32 < n) st1 = s.step[lane + 32];
        }
        if (!(model_change > 0.0)) valid = false;
        else {
          const double d0 = lane < n ? st0 * s.scale[lane] : 0.0, d1 = lane + 32 < n ? st1 * s.scale[lane + 32] : 0.0;
          step_norm = sqrt(wsum(d0 * d0 + d1 * d1));
          if (warp == 0) {
            if (lane < n) { const int g = s.unpos[lane]; s.x_cand[g] = s.x[g] + d0; }
            if (lane + 32 < n) { const int g = s.unpos[lane + 32]; s.x_cand[g] = s.x[g] + d1; }
          }
        }
      }
      if (valid) { num_invalid = 0; break; }
      if (++num_invalid >= 5) { stop = true; break; }
      mu *= 10.0;
      reuse = 0;
    }
    if (stop) break;
    __syncthreads();
    WTICK(7)
    buf ^= 1;
  }
#ifdef MML_WIN_DEVPROF
#if MML_WIN_DEVPROF >= 2
  if (rank == 0 && tid == 0) {
    printf("  chol: topsync=%lld factor=%lld panel=%lld sync=%lld update=%lld endsync=%lld backsub=%lld finalsync=%lld\n", g_cholprof[0], g_cholprof[1],
           g_cholprof[2], g_cholprof[3], g_cholprof[4], g_cholprof[5], g_cholprof[6], g_cholprof[7]);
    for (int k = 0; k < 8; k++) g_cholprof[k] = 0;
  }
#endif
  if (rank == 0 && (tid == 0 || tid == kWLidarThreads))
    printf("win solve tid %d: evals=%d pose=%lld eval=%lld wait=%lld weight=%lld assemble=%lld cluster=%lld totals=%lld | feed: decide+accept=%lld pre-chol=%lld chol=%lld step+rest=%lld cycles\n", tid,
           evals, wtp[0], wtp[1], wtp[2], wtp[3], wtp[4], wtp[5], wtp[6], wtp[8], wtp[9], wtp[10], wtp[7]);
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
    const int oit = wd->outer_it;
    int degenerate = wd->is_degenerate;
    for (int f = 0; f < W; f++) {
      const double* as = A.assoc_stats[wd->slot_of[f]];
      const int* ints = reinterpret_cast<const int*>(as + 16);
      const double sv = as[15];   // localizability value left by the plane association's last CTA (EST.cpp:536-565)
      if (sv < 3.0) degenerate = 1;  // EST.cpp:771-775
      if (f == W - 1) { wd->n_line_last = ints[0]; wd->n_plane_last = ints[1]; wd->min_sv = sv; }
    }
    wd->is_degenerate = degenerate;
    wd->total_inner += it;
    wd->evals += evals;
    wd->final_cost = min_cost;
    const double* sb = wd->states[W - 1];
    const Quat qb = {s.q_before[0], s.q_before[1], s.q_before[2], s.q_before[3]};
    const Quat dq = quat_mul(qb, Quat{sb[3], -sb[4], -sb[5], -sb[6]});
    const double deltaR = 2.0 * atan2(sqrt((dq.x * dq.x + dq.y * dq.y) + dq.z * dq.z), fabs(dq.w)) * 180.0 / 3.14159265358979323846;
    const double d0 = s.t_before[0] - sb[0], d1 = s.t_before[1] - sb[1], d2 = s.t_before[2] - sb[2];
    const double deltaT = sqrt((d0 * d0 + d1 * d1) + d2 * d2);
    const int done = ((deltaR < 0.05 && deltaT < 0.05) || (oit + 1) >= wd->max_outer) ? 1 : 0;
    wd->outer_it = oit + 1;
    if (!done) win_prepare_assoc(wd, oit + 1);
    if (done && A.host_out) {
      for (int f = 0; f < W; f++)
        for (int k = 0; k < 16; k++) A.host_out[16 * f + k] = wd->states[f][k];
      double* st = A.host_out + 16 * kMaxWindow;
      st[0] = oit + 1; st[1] = wd->total_inner; st[2] = wd->n_line_last; st[3] = wd->n_plane_last;
      st[4] = wd->final_cost; st[5] = wd->min_sv; st[6] = degenerate; st[7] = wd->evals; st[8] = W;
      __threadfence_system();
      *A.host_seq = wd->seq;
    }
    wd->done_outer = done;
    if (A.use_cond) cudaGraphSetConditional(A.cond, done ? 0u : 1u);
#ifdef MML_WIN_TIMELINE
    g_wtl_solve_end = wtl_now();
    g_wtl[0] += g_wtl_solve_end - wtl_start;
#endif
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
static int capture_outer_iteration_impl(mml_ctx* c, WindowState* w, int cap, cudaGraphConditionalHandle cond, int use_cond);
static int capture_outer_iteration(mml_ctx* c, WindowState* w, int cap, cudaGraphConditionalHandle cond, int use_cond) {
  c->assoc_table_mode = cap <= 32768;
  const int rc = capture_outer_iteration_impl(c, w, cap, cond, use_cond);
  c->assoc_table_mode = false;
  return rc;
}
static int capture_outer_iteration_impl(mml_ctx* c, WindowState* w, int cap, cudaGraphConditionalHandle cond, int use_cond) {
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
  // scan-sized frames: the association kernels read the maps' descriptors from device memory (ctx->grid_table), so
  // the graph survives map updates; map-sized frames bake the geometry into the launch and are re-captured
  const bool table_mode = cap <= 32768;
  MML_CHECK(mml_grid_table_sync(c));
  mix(table_mode ? 1 : 0); mix((long long)(size_t)c->grid_table.p);
  if (!table_mode)
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
  std::lock_guard<std::recursive_mutex> capture_lock(capture_mutex());
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

// The window's solve as one graph launch, or (MML_WINDOW_DIRECT=1: profilers cannot look inside a graph that holds a
// conditional node) as max_outer plain outer iterations on the stream, the surplus ones gated off by done_outer.
int mml_window_solve_launch(mml_ctx* c, WindowState* w, int cap, int max_outer) {
  static const bool direct = getenv("MML_WINDOW_DIRECT") != nullptr;
  if (!direct) {
    MML_CUDA(c, cudaGraphLaunch(w->graph, c->stream));
    return MML_OK;
  }
  cudaGraphConditionalHandle none;
  memset(&none, 0, sizeof(none));
  const long long before = c->launches;
  for (int it = 0; it < max_outer; it++) MML_CHECK(capture_outer_iteration(c, w, cap, none, 0));
  c->launches = before;
  return MML_OK;
}

int mml_window_begin_launch(mml_ctx* c, WindowState* w) {
  k_win_begin<<<1, 32, 0, c->stream>>>(w->dev.as<WinDev>(), slot_cnt_table(c, w));
  MML_LAUNCHED(c);
  MML_CUDA(c, cudaGetLastError());
  return MML_OK;
}

int mml_window_push_launch(mml_ctx* c, WindowState* w, const WinPush* push_dev) {
  k_win_push<<<1, 512, 0, c->stream>>>(w->dev.as<WinDev>(), push_dev, slot_cnt_table(c, w));
  MML_LAUNCHED(c);
  MML_CUDA(c, cudaGetLastError());
  return MML_OK;
}
