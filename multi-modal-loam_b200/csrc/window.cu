// mmloam_b200: sliding-window Estimate (window sizes 2-4: IMU factors, no marginalisation; BASELINE config 3 is
// window 3). Reference: Estimator::Estimate, src/lio/Estimator.cpp:1143-1581 (IMU blocks 1235-1254, per-frame
// association 1265-1299, lidar blocks 1377-1418, solve 1425-1432, convergence 1441-1450);
// IMUIntegrator::PreIntegration, src/lio/IMUIntegrator.cpp:105-166; Cost_NavState_PRV_Bias,
// include/utils/ceresfunc.h:321-393; pose prediction of process(), src/unionPoseEstimation.cpp:796-835.
//
// Split of the work (SURVEY.md §2 rows 11-12, §8 f F3): the per-point work of every frame — association against the
// resident maps and the residual / Jacobian / Huber / 28-sum reduction — runs on the device, all frames of the
// window in one launch per evaluation (k_accumulate_window). The IMU factors (W-1 blocks of 15 residuals), the
// assembly of the (15 W)^2 normal equations and the dogleg step are a few thousand flops per iteration and stay on
// the host, which reads W x 28 doubles per evaluation.
#include "common.cuh"
#include "smallmath.cuh"
#include "eststate.cuh"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

using namespace mml;

int mml_associate_launch(mml_ctx* ctx, int kind, const double* T16, float thres, const double* T_dev,
                         const float* thres_dev, const int* gate, const int* nq_dev, int cap);
int mml_accumulate_window_launch(mml_ctx* ctx, int W, const float4* const* f_line, const float4* const* f_plane,
                                 const int* n_line, const int* n_plane, const double* x6s, const double* Rbl9,
                                 const double* Pbl3, double lidar_m, double w_tan, double huber_a, double* partials_dev,
                                 unsigned* ticket_dev, double* out_dev, double* host_out_dev, unsigned* host_seq_dev, unsigned seq);
int mml_accumulate_window_grid_max();
int mml_extract_device(mml_ctx* ctx, const float4* xyzi_d, const uint16_t* line_d, const int* scan_off, int n_scans,
                       int n_lines, uint8_t* label_d, bool sequential);
int mml_split_voxel_capacity();
namespace mml { struct SvChain; }
int mml_split_voxel_device(mml_ctx* ctx, const float4* pts_d, const float* s_d, const uint8_t* label_d, int n,
                           const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, float4* corner_out,
                           float4* surf_out, int* counts_d, const mml::SvChain* chain = nullptr);

namespace {

// MML_WIN_PROF=1: wall-clock split of the window loop (debug aid; printed by mml_odom_run_window)
struct WinProf { double push = 0, assoc = 0, launch = 0, imu = 0, wait = 0, solve = 0, other = 0; long evals = 0, scans = 0; };
WinProf g_prof;
const bool g_prof_on = getenv("MML_WIN_PROF") != nullptr;
inline double now_us() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

constexpr int kMaxWindow = 4;

struct WinSlot {
  DevBuf q_corner, q_surf, f_line, f_plane;
  DevBuf assoc_stats, assoc_part[2];  // per-frame association statistics: frames are associated concurrently
  int n_corner = 0, n_surf = 0;
};
struct WindowState {
  WinSlot slot[kMaxWindow];
  int n_slots = 0;
  DevBuf partials, out;   // out: [W][28] sums
  // zero-copy hand-over of an evaluation to the host solver: mapped pinned memory written by the kernel
  // [0, 28 W) sums | [28 W, 28 W + 20 W) association statistics (copied) | sequence words
  void* mapped = nullptr;
  double* mapped_dev = nullptr;
  unsigned seq = 0;
  cudaStream_t fstream[kMaxWindow][2] = {};
  cudaEvent_t fev[kMaxWindow][2] = {};
  cudaEvent_t fork = nullptr;
  bool streams_ok = false;
  // odometry loop: the next scan is copied (host buffers) and labelled on its own stream while the window is solved
  cudaStream_t xstream = nullptr;
  cudaEvent_t xev[2] = {nullptr, nullptr}, xfree[2] = {nullptr, nullptr};
  DevBuf x_xyzi[2], x_line[2], x_s[2], x_label[2], x_counters[2];
};
constexpr int kMapDoubles = (28 + 20) * kMaxWindow;
constexpr size_t kMapBytes = sizeof(double) * kMapDoubles + 64;

WindowState* win(mml_ctx* c) {
  if (!c->window) c->window = new WindowState();
  return static_cast<WindowState*>(c->window);
}
int win_prepare(mml_ctx* c, WindowState* w) {
  if (!w->mapped) {
    MML_CUDA(c, cudaHostAlloc(&w->mapped, kMapBytes, cudaHostAllocMapped));
    memset(w->mapped, 0, kMapBytes);
    void* d = nullptr;
    MML_CUDA(c, cudaHostGetDevicePointer(&d, w->mapped, 0));
    w->mapped_dev = static_cast<double*>(d);
  }
  if (!w->streams_ok) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    for (int f = 0; f < kMaxWindow; f++) for (int k = 0; k < 2; k++) {
      MML_CUDA(c, cudaStreamCreateWithPriority(&w->fstream[f][k], cudaStreamNonBlocking, hi));
      MML_CUDA(c, cudaEventCreateWithFlags(&w->fev[f][k], cudaEventDisableTiming));
    }
    MML_CUDA(c, cudaEventCreateWithFlags(&w->fork, cudaEventDisableTiming));
    MML_CUDA(c, cudaStreamCreateWithPriority(&w->xstream, cudaStreamNonBlocking, lo));
    for (int k = 0; k < 2; k++) {
      MML_CUDA(c, cudaEventCreateWithFlags(&w->xev[k], cudaEventDisableTiming));
      MML_CUDA(c, cudaEventCreateWithFlags(&w->xfree[k], cudaEventDisableTiming));
    }
    w->streams_ok = true;
  }
  return MML_OK;
}

// ---- scalar overloads used by the functor text below when it is instantiated for plain doubles
__host__ __device__ inline double dsqrt(double f) { return sqrt(f); }
__host__ __device__ inline double dsin(double f) { return sin(f); }
__host__ __device__ inline double dcos(double f) { return cos(f); }
__host__ __device__ inline double datan(double f) { return atan(f); }
__host__ __device__ inline double val(double x) { return x; }

template <class T> struct Q4 { T w, x, y, z; };
template <class T> __host__ __device__ inline Q4<T> qmul(const Q4<T>& a, const Q4<T>& b) {  // sophus/so3.hpp:326-340
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
template <class T> __host__ __device__ inline Q4<T> qconj(const Q4<T>& q) { return {q.w, -q.x, -q.y, -q.z}; }
template <class T> __host__ __device__ inline void qrot(const Q4<T>& q, const T* p, T* out) {  // so3.hpp:358-371
  T uv[3] = {q.y * p[2] - q.z * p[1], q.z * p[0] - q.x * p[2], q.x * p[1] - q.y * p[0]};
  for (int k = 0; k < 3; k++) uv[k] = uv[k] + uv[k];
  T c[3] = {q.y * uv[2] - q.z * uv[1], q.z * uv[0] - q.x * uv[2], q.x * uv[1] - q.y * uv[0]};
  for (int k = 0; k < 3; k++) out[k] = p[k] + q.w * uv[k] + c[k];
}
template <class T> __host__ __device__ inline Q4<T> qexp(const T* om) {  // so3.hpp:585-623, epsilon 1e-10
  T theta_sq = (om[0] * om[0] + om[1] * om[1]) + om[2] * om[2];
  T imag, real;
  if (val(theta_sq) < 1e-20) {
    T theta_po4 = theta_sq * theta_sq;
    imag = T(0.5) - T(1.0 / 48.0) * theta_sq + T(1.0 / 3840.0) * theta_po4;
    real = T(1.0) - T(1.0 / 8.0) * theta_sq + T(1.0 / 384.0) * theta_po4;
  } else {
    T theta = dsqrt(theta_sq);
    T half = T(0.5) * theta;
    imag = dsin(half) / theta;
    real = dcos(half);
  }
  return {real, imag * om[0], imag * om[1], imag * om[2]};
}
template <class T> __host__ __device__ inline void qlog(const Q4<T>& q, T* out) {  // so3.hpp:247-292
  T squared_n = (q.x * q.x + q.y * q.y) + q.z * q.z;
  T w = q.w;
  T f;
  if (val(squared_n) < 1e-20) {
    T squared_w = w * w;
    f = T(2.0) / w - T(2.0 / 3.0) * squared_n / (w * squared_w);
  } else {
    T n = dsqrt(squared_n);
    if (fabs(val(w)) < 1e-10) f = T(val(w) > 0 ? 3.14159265358979323846 : -3.14159265358979323846) / n;
    else f = T(2.0) * datan(n / w) / n;
  }
  out[0] = f * q.x; out[1] = f * q.y; out[2] = f * q.z;
}

// Cost_NavState_PRV_Bias::operator(), CF.h:331-377, before the multiplication by sqrt_information
template <class T>
__host__ __device__ void imu_residual(const mml_preint& m, const double* g, const T* pri, const T* vbi, const T* prj,
                                      const T* vbj, T* r) {
  const Q4<T> Ri = qexp(pri + 3), Rj = qexp(prj + 3);
  T dbg[3], dba[3];
  for (int k = 0; k < 3; k++) { dbg[k] = vbi[3 + k] - T(m.bg[k]); dba[k] = vbi[6 + k] - T(m.ba[k]); }
  const double dT = m.dt, dT2 = m.dt * m.dt;
  Q4<T> dRij;
  {  // Sophus::SO3<T>(quaternion) normalises, so3.hpp:487-494
    const double n = sqrt(((m.dq[1] * m.dq[1] + m.dq[2] * m.dq[2]) + m.dq[3] * m.dq[3]) + m.dq[0] * m.dq[0]);
    dRij = {T(m.dq[0] / n), T(m.dq[1] / n), T(m.dq[2] / n), T(m.dq[3] / n)};
  }
  const Q4<T> RiT = qconj(Ri);
#define MML_J(r0, c0, r, c) m.jac[((r0) + (r)) * 15 + (c0) + (c)]
  T a[3], ra[3];
  for (int k = 0; k < 3; k++) a[k] = prj[k] - pri[k] - vbi[k] * T(dT) - T(0.5 * g[k]) * T(dT2);
  qrot(RiT, a, ra);
  for (int k = 0; k < 3; k++) {
    T c = T(m.dp[k]) + ((T(MML_J(0, 9, k, 0)) * dbg[0] + T(MML_J(0, 9, k, 1)) * dbg[1]) + T(MML_J(0, 9, k, 2)) * dbg[2]) +
          ((T(MML_J(0, 12, k, 0)) * dba[0] + T(MML_J(0, 12, k, 1)) * dba[1]) + T(MML_J(0, 12, k, 2)) * dba[2]);
    r[k] = ra[k] - c;
  }
  T w[3];
  for (int k = 0; k < 3; k++) w[k] = (T(MML_J(3, 9, k, 0)) * dbg[0] + T(MML_J(3, 9, k, 1)) * dbg[1]) + T(MML_J(3, 9, k, 2)) * dbg[2];
  const Q4<T> dR_dbg = qexp(w);
  const Q4<T> rR = qmul(qmul(qconj(qmul(dRij, dR_dbg)), RiT), Rj);
  qlog(rR, r + 3);
  for (int k = 0; k < 3; k++) a[k] = vbj[k] - vbi[k] - T(g[k]) * T(dT);
  qrot(RiT, a, ra);
  for (int k = 0; k < 3; k++) {
    T c = T(m.dv[k]) + ((T(MML_J(6, 9, k, 0)) * dbg[0] + T(MML_J(6, 9, k, 1)) * dbg[1]) + T(MML_J(6, 9, k, 2)) * dbg[2]) +
          ((T(MML_J(6, 12, k, 0)) * dba[0] + T(MML_J(6, 12, k, 1)) * dba[1]) + T(MML_J(6, 12, k, 2)) * dba[2]);
    r[6 + k] = ra[k] - c;
  }
#undef MML_J
  for (int k = 0; k < 6; k++) r[9 + k] = vbj[3 + k] - vbi[3 + k];
}

// Forward-mode dual number over the 30 parameters of the IMU factor: what Ceres' Jet<double, 30> is. The functor is
// differentiated automatically, exactly as the reference does (same derivative values, no hand-derived Jacobian).
struct Dual30 {
  double a, v[30];
  __host__ __device__ Dual30() : a(0) { for (int i = 0; i < 30; i++) v[i] = 0; }
  __host__ __device__ Dual30(double s) : a(s) { for (int i = 0; i < 30; i++) v[i] = 0; }
};
__host__ __device__ inline Dual30 operator+(const Dual30& f, const Dual30& g) { Dual30 h; h.a = f.a + g.a; for (int i = 0; i < 30; i++) h.v[i] = f.v[i] + g.v[i]; return h; }
__host__ __device__ inline Dual30 operator-(const Dual30& f, const Dual30& g) { Dual30 h; h.a = f.a - g.a; for (int i = 0; i < 30; i++) h.v[i] = f.v[i] - g.v[i]; return h; }
__host__ __device__ inline Dual30 operator-(const Dual30& f) { Dual30 h; h.a = -f.a; for (int i = 0; i < 30; i++) h.v[i] = -f.v[i]; return h; }
__host__ __device__ inline Dual30 operator*(const Dual30& f, const Dual30& g) { Dual30 h; h.a = f.a * g.a; for (int i = 0; i < 30; i++) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
__host__ __device__ inline Dual30 operator/(const Dual30& f, const Dual30& g) { Dual30 h; const double gi = 1.0 / g.a, q = f.a * gi; h.a = q; for (int i = 0; i < 30; i++) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h; }
__host__ __device__ inline Dual30 chain30(double val_, double d, const Dual30& f) { Dual30 h; h.a = val_; for (int i = 0; i < 30; i++) h.v[i] = d * f.v[i]; return h; }
__host__ __device__ inline Dual30 dsqrt(const Dual30& f) { const double t = sqrt(f.a); return chain30(t, 1.0 / (2.0 * t), f); }
__host__ __device__ inline Dual30 dsin(const Dual30& f) { return chain30(sin(f.a), cos(f.a), f); }
__host__ __device__ inline Dual30 dcos(const Dual30& f) { return chain30(cos(f.a), -sin(f.a), f); }
__host__ __device__ inline Dual30 datan(const Dual30& f) { return chain30(atan(f.a), 1.0 / (1.0 + f.a * f.a), f); }
__host__ __device__ inline double val(const Dual30& x) { return x.a; }

// weighted residual r15 = sqrt_info * r and Jacobian J (15 x 30 row-major, columns [PR_i | VBias_i | PR_j | VBias_j])
void imu_factor_eval(const mml_preint& m, const double* g, const double* pri, const double* vbi, const double* prj,
                     const double* vbj, double* r15, double* J450) {
  const double* src[4] = {pri, vbi, prj, vbj};
  const int sz[4] = {6, 9, 6, 9}, off[4] = {0, 6, 15, 21};
  double x[30];
  for (int b = 0; b < 4; b++) for (int k = 0; k < sz[b]; k++) x[off[b] + k] = src[b][k];
  double r[15], rw[15 * 31];
  if (!J450) {
    imu_residual<double>(m, g, x, x + 6, x + 15, x + 21, r);
    for (int i = 0; i < 15; i++) rw[i] = r[i];
  } else {
    Dual30 xd[30], rd[15];
    for (int k = 0; k < 30; k++) { xd[k] = Dual30(x[k]); xd[k].v[k] = 1.0; }
    imu_residual<Dual30>(m, g, xd, xd + 6, xd + 15, xd + 21, rd);
    for (int i = 0; i < 15; i++) { rw[i] = rd[i].a; for (int c = 0; c < 30; c++) rw[15 * (1 + c) + i] = rd[i].v[c]; }
  }
  for (int i = 0; i < 15; i++) {
    double s = 0;
    for (int k = 0; k < 15; k++) s += m.sqrt_info[i * 15 + k] * rw[k];
    r15[i] = s;
    if (J450) for (int c = 0; c < 30; c++) {
      double t = 0;
      for (int k = 0; k < 15; k++) t += m.sqrt_info[i * 15 + k] * rw[15 * (1 + c) + k];
      J450[i * 30 + c] = t;
    }
  }
}

void mat_mul(int n, int k, int m, const double* A, const double* B, double* C) {
  for (int i = 0; i < n; i++) for (int j = 0; j < m; j++) {
    double s = 0;
    for (int t = 0; t < k; t++) s += A[i * k + t] * B[t * m + j];
    C[i * m + j] = s;
  }
}
void hat3(const double* v, double* K) { K[0] = 0; K[1] = -v[2]; K[2] = v[1]; K[3] = v[2]; K[4] = 0; K[5] = -v[0]; K[6] = -v[1]; K[7] = v[0]; K[8] = 0; }
bool invert_n(int n, const double* A, double* inv) {  // Gauss-Jordan, partial pivoting
  std::vector<double> a((size_t)n * 2 * n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { a[(size_t)i * 2 * n + j] = A[i * n + j]; a[(size_t)i * 2 * n + n + j] = (i == j); }
  for (int c = 0; c < n; c++) {
    int p = c;
    for (int i = c + 1; i < n; i++) if (fabs(a[(size_t)i * 2 * n + c]) > fabs(a[(size_t)p * 2 * n + c])) p = i;
    if (a[(size_t)p * 2 * n + c] == 0.0) return false;
    if (p != c) for (int j = 0; j < 2 * n; j++) std::swap(a[(size_t)p * 2 * n + j], a[(size_t)c * 2 * n + j]);
    const double d = a[(size_t)c * 2 * n + c];
    for (int j = 0; j < 2 * n; j++) a[(size_t)c * 2 * n + j] /= d;
    for (int i = 0; i < n; i++) if (i != c) {
      const double f = a[(size_t)i * 2 * n + c];
      if (f == 0.0) continue;
      for (int j = 0; j < 2 * n; j++) a[(size_t)i * 2 * n + j] -= f * a[(size_t)c * 2 * n + j];
    }
  }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) inv[i * n + j] = a[(size_t)i * 2 * n + n + j];
  return true;
}
bool chol_solve_n(int n, const double* A, const double* b, double* x, std::vector<double>& L, std::vector<double>& y) {
  L.resize((size_t)n * n);
  y.resize(n);
  for (int i = 0; i < n; i++) {
    double* Li = &L[(size_t)i * n];
    for (int j = 0; j <= i; j++) {
      const double* Lj = &L[(size_t)j * n];
      double s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= Li[k] * Lj[k];
      if (i == j) { if (!(s > 0)) return false; Li[i] = sqrt(s); }
      else Li[j] = s / Lj[j];
    }
  }
  for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[i * n + k] * y[k]; y[i] = s / L[i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k]; x[i] = s / L[i * n + i]; }
  return true;
}

// ceres::Solve as EST.cpp:1425-1432 configures it (TrustRegionMinimizer, traditional dogleg, Jacobi scaling,
// Ceres 2.1.0 defaults otherwise) on dense normal equations of any size: the same state machine as the 6-dim
// device version (accumulate.cu dogleg_update), driven by evaluations the caller supplies.
struct DoglegN {
  int n = 0, max_it = 10, it = 0;
  std::vector<double> x, x_cand, x_best, H, g, scale, Hs, gs, diag, grad, gn, step, A, y, L, ytmp, vtmp;
  double cost = 0, min_cost = 0, radius = 1e4, mu = 1e-8, alpha = 0, dogleg_norm = 0, model_change = 0, step_norm = 0, x_norm = 0;
  bool reuse = false, first = true, done = false;
  int num_invalid = 0, iterations = 0;
  void begin(int n_, const double* x0, int max_iterations) {
    n = n_; max_it = max_iterations; it = 0; iterations = 0;
    x.assign(x0, x0 + n); x_cand = x; x_best = x;
    H.assign((size_t)n * n, 0); g.assign(n, 0); scale.assign(n, 1); Hs = H; gs = g; diag = g; grad = g; gn = g; step = g; y = g;
    radius = 1e4; mu = 1e-8; reuse = false; first = true; done = false; num_invalid = 0;
  }
  const double* eval_point() const { return first ? x.data() : x_cand.data(); }
  double gmax(const std::vector<double>& v) const { double m = 0; for (double e : v) m = std::max(m, fabs(e)); return m; }
  void apply_scale() {
    for (int i = 0; i < n; i++) { gs[i] = g[i] * scale[i]; for (int j = 0; j < n; j++) Hs[(size_t)i * n + j] = H[(size_t)i * n + j] * scale[i] * scale[j]; }
  }
  bool compute_step() {
    if (!reuse) {
      reuse = true;
      for (int i = 0; i < n; i++) diag[i] = sqrt(std::min(std::max(Hs[(size_t)i * n + i], 1e-6), 1e32));
      for (int i = 0; i < n; i++) grad[i] = gs[i] / diag[i];
      double gg = 0, vHv = 0;
      std::vector<double>& v = vtmp;
      v.resize(n);
      for (int i = 0; i < n; i++) { v[i] = grad[i] / diag[i]; gg += grad[i] * grad[i]; }
      for (int i = 0; i < n; i++) { double s = 0; for (int j = 0; j < n; j++) s += Hs[(size_t)i * n + j] * v[j]; vHv += v[i] * s; }
      alpha = gg / vHv;
      bool ok = false;
      while (mu < 1.0) {
        A = Hs;
        for (int i = 0; i < n; i++) A[(size_t)i * n + i] += mu * diag[i] * diag[i];
        bool s_ok = chol_solve_n(n, A.data(), gs.data(), y.data(), L, ytmp);
        if (s_ok) for (int i = 0; i < n; i++) if (!std::isfinite(y[i])) s_ok = false;
        if (!s_ok) { mu *= 10.0; continue; }
        for (int i = 0; i < n; i++) gn[i] = -diag[i] * y[i];
        ok = true;
        break;
      }
      if (!ok) return false;
    }
    double gnorm = 0, gnn = 0;
    for (int i = 0; i < n; i++) { gnorm += grad[i] * grad[i]; gnn += gn[i] * gn[i]; }
    gnorm = sqrt(gnorm); gnn = sqrt(gnn);
    if (gnn <= radius) { step = gn; dogleg_norm = gnn; }
    else if (gnorm * alpha >= radius) { for (int i = 0; i < n; i++) step[i] = -(radius / gnorm) * grad[i]; dogleg_norm = radius; }
    else {
      double b_dot_a = 0;
      for (int i = 0; i < n; i++) b_dot_a += grad[i] * gn[i];
      b_dot_a *= -alpha;
      const double a_sq = (alpha * gnorm) * (alpha * gnorm);
      const double bma = a_sq - 2 * b_dot_a + gnn * gnn;
      const double c = b_dot_a - a_sq;
      const double d = sqrt(c * c + bma * (radius * radius - a_sq));
      const double beta = (c <= 0) ? (d - c) / bma : (radius * radius - a_sq) / (d + c);
      double sn = 0;
      for (int i = 0; i < n; i++) { step[i] = (-alpha * (1.0 - beta)) * grad[i] + beta * gn[i]; sn += step[i] * step[i]; }
      dogleg_norm = sqrt(sn);
    }
    for (int i = 0; i < n; i++) step[i] /= diag[i];
    double sg = 0, sHs = 0;
    for (int i = 0; i < n; i++) { double t = 0; for (int j = 0; j < n; j++) t += Hs[(size_t)i * n + j] * step[j]; sHs += step[i] * t; sg += step[i] * gs[i]; }
    model_change = -sg - 0.5 * sHs;
    if (!(model_change > 0.0)) return false;
    double sn = 0;
    for (int i = 0; i < n; i++) { const double d = step[i] * scale[i]; x_cand[i] = x[i] + d; sn += d * d; }
    step_norm = sqrt(sn);
    return true;
  }
  // next step (with Ceres' handling of invalid steps); sets done when the iteration budget or the retries run out
  void advance() {
    while (true) {
      if (it >= max_it) { done = true; return; }
      it++; iterations = it;
      if (compute_step()) { num_invalid = 0; return; }
      if (++num_invalid >= 5) { done = true; return; }
      mu *= 10.0;
      reuse = false;
    }
  }
  // feed the evaluation at eval_point(): cost, H (n x n), g
  void feed(double c, const double* Hn, const double* gn_) {
    auto xnorm = [&]() { double s = 0; for (double e : x) s += e * e; return sqrt(s); };
    if (first) {
      first = false;
      cost = c; min_cost = c;
      H.assign(Hn, Hn + (size_t)n * n); g.assign(gn_, gn_ + n);
      for (int i = 0; i < n; i++) scale[i] = 1.0 / (1.0 + sqrt(H[(size_t)i * n + i]));
      apply_scale();
      x_norm = xnorm();
      if (!std::isfinite(c) || gmax(g) <= 1e-10) { done = true; return; }
      advance();
      return;
    }
    const double cand_cost = std::isfinite(c) ? c : DBL_MAX;
    if (step_norm <= 1e-8 * (x_norm + 1e-8)) { done = true; return; }
    const double cost_change = cost - cand_cost;
    if (fabs(cost_change) <= 1e-6 * cost) { done = true; return; }
    const double rel = cost_change / model_change;
    if (rel > 1e-3) {
      x = x_cand; cost = cand_cost;
      H.assign(Hn, Hn + (size_t)n * n); g.assign(gn_, gn_ + n);
      apply_scale();
      x_norm = xnorm();
      if (rel < 0.25) radius *= 0.5;
      if (rel > 0.75) radius = std::max(radius, 3.0 * dogleg_norm);
      mu = std::max(1e-8, 2.0 * mu / 10.0);
      reuse = false;
      if (cost < min_cost) { min_cost = cost; x_best = x; }
      if (gmax(g) <= 1e-10) { done = true; return; }
    } else {
      radius *= 0.5;
      reuse = true;
    }
    if (radius < 1e-32) { done = true; return; }
    advance();
  }
};

}  // namespace

void mml_window_destroy(mml_ctx* c) {
  if (!c->window) return;
  WindowState* w = static_cast<WindowState*>(c->window);
  for (auto& s : w->slot) {
    s.q_corner.release(); s.q_surf.release(); s.f_line.release(); s.f_plane.release();
    s.assoc_stats.release(); s.assoc_part[0].release(); s.assoc_part[1].release();
  }
  w->partials.release(); w->out.release();
  if (w->mapped) cudaFreeHost(w->mapped);
  if (w->streams_ok) {
    for (int f = 0; f < kMaxWindow; f++) for (int k = 0; k < 2; k++) { cudaStreamDestroy(w->fstream[f][k]); cudaEventDestroy(w->fev[f][k]); }
    cudaEventDestroy(w->fork);
    cudaStreamDestroy(w->xstream);
    for (int k = 0; k < 2; k++) { cudaEventDestroy(w->xev[k]); cudaEventDestroy(w->xfree[k]); }
  }
  for (int k = 0; k < 2; k++) { w->x_xyzi[k].release(); w->x_line[k].release(); w->x_s[k].release(); w->x_label[k].release(); w->x_counters[k].release(); }
  delete w;
  c->window = nullptr;
}

extern "C" {

// IMUIntegrator::PreIntegration, IMU.cpp:105-166 (+ sqrt_information of EST.cpp:1240-1242). Host side: ~20 samples
// of 15x15 algebra per scan.
int mml_imu_preintegrate(const double* t, const double* gyr, const double* acc, int n, double last_time,
                         const double* bg3, const double* ba3, mml_preint* out) {
  if (!out || n < 0 || (n && (!t || !gyr || !acc)) || !bg3 || !ba3) return MML_ERR_INVALID;
  const double acc_n = 0.08, gyr_n = 0.004, acc_w = 2.0e-4, gyr_w = 2.0e-5, gnorm = 9.805;  // IMU.h:79-84
  Quat dq = {1, 0, 0, 0};
  double dp[3] = {0, 0, 0}, dv[3] = {0, 0, 0}, dtime = 0;
  std::vector<double> cov(225, 0.0), jac(225, 0.0), noise(144, 0.0), tmp(225), tmp2(225), AT(225), BN(180), BT(180), BNB(225);
  for (int i = 0; i < 15; i++) jac[i * 15 + i] = 1.0;
  for (int i = 0; i < 3; i++) {
    noise[i * 12 + i] = gyr_n * gyr_n; noise[(3 + i) * 12 + 3 + i] = acc_n * acc_n;
    noise[(6 + i) * 12 + 6 + i] = gyr_w * gyr_w; noise[(9 + i) * 12 + 9 + i] = acc_w * acc_w;
  }
  double current_time = last_time;
  for (int s = 0; s < n; s++) {
    const double g3[3] = {gyr[3 * s] - bg3[0], gyr[3 * s + 1] - bg3[1], gyr[3 * s + 2] - bg3[2]};
    const double a3[3] = {acc[3 * s] * gnorm - ba3[0], acc[3 * s + 1] * gnorm - ba3[1], acc[3 * s + 2] * gnorm - ba3[2]};
    const double dt = t[s] - current_time, dt2 = dt * dt;
    const double gdt[3] = {g3[0] * dt, g3[1] * dt, g3[2] * dt};
    double dR[9];
    quat_to_R(so3_exp(gdt), dR);
    double Jr[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double nrm = sqrt((gdt[0] * gdt[0] + gdt[1] * gdt[1]) + gdt[2] * gdt[2]);
    if (nrm > 0.00001) {
      const double k[3] = {gdt[0] / nrm, gdt[1] / nrm, gdt[2] / nrm};
      double K[9], KK[9];
      hat3(k, K);
      mat_mul(3, 3, 3, K, K, KK);
      const double c1 = (1 - cos(nrm)) / nrm, c2 = 1 - sin(nrm) / nrm;
      for (int i = 0; i < 9; i++) Jr[i] = (i % 4 == 0 ? 1.0 : 0.0) - c1 * K[i] + c2 * KK[i];
    }
    double Rq[9], Ha[9], RH[9];
    quat_to_R(dq, Rq);
    hat3(a3, Ha);
    mat_mul(3, 3, 3, Rq, Ha, RH);
    double A[225], B[180];
    memset(A, 0, sizeof(A)); memset(B, 0, sizeof(B));
    for (int i = 0; i < 15; i++) A[i * 15 + i] = 1.0;
    auto setA = [&](int r0, int c0, const double* M, double f) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) A[(r0 + r) * 15 + c0 + c] = f * M[3 * r + c]; };
    auto setB = [&](int r0, int c0, const double* M, double f) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) B[(r0 + r) * 12 + c0 + c] = f * M[3 * r + c]; };
    const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double dRT[9] = {dR[0], dR[3], dR[6], dR[1], dR[4], dR[7], dR[2], dR[5], dR[8]};
    setA(0, 3, RH, -0.5 * dt2); setA(0, 6, I3, dt); setA(0, 12, Rq, -0.5 * dt2);
    setA(3, 3, dRT, 1.0); setA(3, 9, Jr, -dt);
    setA(6, 3, RH, -dt); setA(6, 12, Rq, -dt);
    setB(0, 3, Rq, 0.5 * dt2); setB(3, 0, Jr, dt); setB(6, 3, Rq, dt); setB(9, 6, I3, dt); setB(12, 9, I3, dt);
    mat_mul(15, 15, 15, A, jac.data(), tmp.data());
    jac = tmp;
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) AT[i * 15 + j] = A[j * 15 + i];
    mat_mul(15, 15, 15, A, cov.data(), tmp.data());
    mat_mul(15, 15, 15, tmp.data(), AT.data(), tmp2.data());
    mat_mul(15, 12, 12, B, noise.data(), BN.data());
    for (int i = 0; i < 12; i++) for (int j = 0; j < 15; j++) BT[i * 15 + j] = B[j * 12 + i];
    mat_mul(15, 12, 15, BN.data(), BT.data(), BNB.data());
    for (int i = 0; i < 225; i++) cov[i] = tmp2[i] + BNB[i];
    const double Ra[3] = {Rq[0] * a3[0] + Rq[1] * a3[1] + Rq[2] * a3[2], Rq[3] * a3[0] + Rq[4] * a3[1] + Rq[5] * a3[2],
                          Rq[6] * a3[0] + Rq[7] * a3[1] + Rq[8] * a3[2]};
    for (int k = 0; k < 3; k++) dp[k] += dv[k] * dt + 0.5 * Ra[k] * dt2;
    for (int k = 0; k < 3; k++) dv[k] += Ra[k] * dt;
    double m3[9];
    mat_mul(3, 3, 3, Rq, dR, m3);
    Quat qt = quat_from_R9(m3);
    if (qt.w < 0) { qt.w = -qt.w; qt.x = -qt.x; qt.y = -qt.y; qt.z = -qt.z; }
    const double qn = sqrt(((qt.x * qt.x + qt.y * qt.y) + qt.z * qt.z) + qt.w * qt.w);
    dq = {qt.w / qn, qt.x / qn, qt.y / qn, qt.z / qn};
    dtime += dt;
    current_time = t[s];
  }
  out->dq[0] = dq.w; out->dq[1] = dq.x; out->dq[2] = dq.y; out->dq[3] = dq.z;
  for (int k = 0; k < 3; k++) { out->dp[k] = dp[k]; out->dv[k] = dv[k]; out->bg[k] = bg3[k]; out->ba[k] = ba3[k]; }
  out->dt = dtime;
  memcpy(out->cov, cov.data(), sizeof(out->cov));
  memcpy(out->jac, jac.data(), sizeof(out->jac));
  double inv[225], L[225];
  memset(L, 0, sizeof(L));
  memset(out->sqrt_info, 0, sizeof(out->sqrt_info));
  if (n > 0 && invert_n(15, out->cov, inv)) {
    for (int i = 0; i < 15; i++) for (int j = 0; j <= i; j++) {
      double s = inv[i * 15 + j];
      for (int k = 0; k < j; k++) s -= L[i * 15 + k] * L[j * 15 + k];
      L[i * 15 + j] = (i == j) ? sqrt(s) : s / L[j * 15 + j];
    }
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) out->sqrt_info[i * 15 + j] = L[j * 15 + i];
  }
  return MML_OK;
}

int mml_imu_factor(const mml_preint* pre, const double* gravity3, const double* pri6, const double* vbi9,
                   const double* prj6, const double* vbj9, double* r15, double* J450) {
  if (!pre || !gravity3 || !pri6 || !vbi9 || !prj6 || !vbj9 || !r15) return MML_ERR_INVALID;
  imu_factor_eval(*pre, gravity3, pri6, vbi9, prj6, vbj9, r15, J450);
  return MML_OK;
}

// PE.cpp:812-829. state = P(3) q_wxyz(4) V(3) bg(3) ba(3)
int mml_imu_predict(const double* prev16, const mml_preint* pre, double* next16) {
  if (!prev16 || !pre || !next16) return MML_ERR_INVALID;
  auto rot = [](const double* q, const double* v, double* o) {  // Eigen 3.3 _transformVector
    const double uv0[3] = {q[2] * v[2] - q[3] * v[1], q[3] * v[0] - q[1] * v[2], q[1] * v[1] - q[2] * v[0]};
    const double uv[3] = {uv0[0] + uv0[0], uv0[1] + uv0[1], uv0[2] + uv0[2]};
    const double c[3] = {q[2] * uv[2] - q[3] * uv[1], q[3] * uv[0] - q[1] * uv[2], q[1] * uv[1] - q[2] * uv[0]};
    for (int k = 0; k < 3; k++) o[k] = v[k] + q[0] * uv[k] + c[k];
  };
  const double* Qp = prev16 + 3;
  const Quat Q = quat_mul(Quat{Qp[0], Qp[1], Qp[2], Qp[3]}, Quat{pre->dq[0], pre->dq[1], pre->dq[2], pre->dq[3]});
  double rp[3], rv[3];
  rot(Qp, pre->dp, rp);
  rot(Qp, pre->dv, rv);
  for (int k = 0; k < 3; k++) { next16[k] = prev16[k] + rp[k]; next16[7 + k] = prev16[7 + k] + rv[k]; }
  next16[3] = Q.w; next16[4] = Q.x; next16[5] = Q.y; next16[6] = Q.z;
  for (int k = 0; k < 6; k++) next16[10 + k] = prev16[10 + k];
  return MML_OK;
}

// ---- window slots: the downsampled corner / surf clouds of the frames in the window stay in HBM -----------------
int mml_window_reset(mml_ctx* c) {
  if (!c) return MML_ERR_INVALID;
  win(c)->n_slots = 0;
  return MML_OK;
}
int mml_window_size(mml_ctx* c) { return c ? win(c)->n_slots : 0; }

static int window_make_room(mml_ctx* c, int max_frames) {
  WindowState* w = win(c);
  if (max_frames < 1 || max_frames > kMaxWindow) return mml_fail(c, MML_ERR_INVALID, "window size must be 1..4");
  while (w->n_slots >= max_frames) {  // drop the oldest frame (PE.cpp:830-832), buffers rotate to the back
    WinSlot first = w->slot[0];
    for (int f = 0; f + 1 < w->n_slots; f++) w->slot[f] = w->slot[f + 1];
    w->slot[w->n_slots - 1] = first;
    w->n_slots--;
  }
  return MML_OK;
}

// push a frame given as host clouds (already undistorted and voxel-filtered: the A6 outputs)
int mml_window_push_frame(mml_ctx* c, const float* corner_xyzi, int n_corner, const float* surf_xyzi, int n_surf, int max_frames) {
  if (!c || n_corner < 0 || n_surf < 0) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  MML_CHECK(window_make_room(c, max_frames));
  WindowState* w = win(c);
  WinSlot& s = w->slot[w->n_slots];
  MML_CUDA(c, s.q_corner.reserve(sizeof(float4) * (size_t)(n_corner + 1)));
  MML_CUDA(c, s.q_surf.reserve(sizeof(float4) * (size_t)(n_surf + 1)));
  if (n_corner) MML_CUDA(c, cudaMemcpyAsync(s.q_corner.p, corner_xyzi, sizeof(float4) * (size_t)n_corner, cudaMemcpyHostToDevice, c->stream));
  if (n_surf) MML_CUDA(c, cudaMemcpyAsync(s.q_surf.p, surf_xyzi, sizeof(float4) * (size_t)n_surf, cudaMemcpyHostToDevice, c->stream));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  s.n_corner = n_corner; s.n_surf = n_surf;
  w->n_slots++;
  return MML_OK;
}

// push a raw scan resident in HBM: extraction (A1) -> undistortion + label split + voxel filter (A4, A6) on the
// device, the downsampled clouds become the newest window frame. out_counts (may be NULL): n_sharp, n_flat,
// n_corner_ds, n_surf_ds.
static int window_push_scan(mml_ctx* c, const void* xyzi_dev, const void* line_id_dev, const void* s_dev, int n, int n_lines,
                            const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, int max_frames,
                            int* out_counts, const uint8_t* pre_label, const int* pre_counters);

int mml_window_push_scan_dev(mml_ctx* c, const void* xyzi_dev, const void* line_id_dev, const void* s_dev, int n, int n_lines,
                             const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, int max_frames,
                             int* out_counts) {
  return window_push_scan(c, xyzi_dev, line_id_dev, s_dev, n, n_lines, dR9, dt3, leaf_corner, leaf_surf, max_frames, out_counts,
                          nullptr, nullptr);
}

// pre_label / pre_counters: labels and extractor counters already produced for this scan (the loop labels scan k+1
// on its own stream while the window of scan k is solved; the caller has made c->stream wait for them)
static int window_push_scan(mml_ctx* c, const void* xyzi_dev, const void* line_id_dev, const void* s_dev, int n, int n_lines,
                            const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, int max_frames,
                            int* out_counts, const uint8_t* pre_label, const int* pre_counters) {
  if (!c || n < 0) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  MML_CHECK(window_make_room(c, max_frames));
  WindowState* w = win(c);
  WinSlot& s = w->slot[w->n_slots];
  cudaStream_t st = c->stream;
  const int cap = mml_split_voxel_capacity();
  MML_CUDA(c, c->in_label.reserve((size_t)n + 16));
  MML_CUDA(c, c->frame_cnt.reserve(64));
  MML_CUDA(c, s.q_corner.reserve(sizeof(float4) * (size_t)cap));
  MML_CUDA(c, s.q_surf.reserve(sizeof(float4) * (size_t)cap));
  MML_CUDA(c, c->pin_flags.reserve(64));
  const int off[2] = {0, n};
  const uint8_t* label_d = pre_label;
  const int* counters_d = pre_counters;
  if (!label_d) {
    MML_CHECK(mml_extract_device(c, (const float4*)xyzi_dev, (const uint16_t*)line_id_dev, off, 1, n_lines, c->in_label.as<uint8_t>(), false));
    label_d = c->in_label.as<uint8_t>();
    counters_d = c->counters.as<int>();
  }
  int* cnt = c->frame_cnt.as<int>();
  const bool undist = dR9 && dt3 && s_dev;
  MML_CHECK(mml_split_voxel_device(c, (const float4*)xyzi_dev, undist ? (const float*)s_dev : nullptr, label_d, n,
                                   dR9, dt3, leaf_corner, leaf_surf, s.q_corner.as<float4>(), s.q_surf.as<float4>(), cnt));
  int* hf = c->pin_flags.as<int>();
  MML_CUDA(c, cudaMemcpyAsync(hf, counters_d, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
  MML_CUDA(c, cudaMemcpyAsync(hf + 4, cnt, 5 * sizeof(int), cudaMemcpyDeviceToHost, st));
  MML_CUDA(c, cudaStreamSynchronize(st));
  if (hf[2] || hf[8]) return mml_fail(c, MML_ERR_CAPACITY, "scan exceeds the fused extraction / split-voxel capacities (window path)");
  s.n_corner = hf[4]; s.n_surf = hf[5];
  if (out_counts) { out_counts[0] = hf[0]; out_counts[1] = hf[1]; out_counts[2] = hf[4]; out_counts[3] = hf[5]; }
  w->n_slots++;
  return MML_OK;
}

int mml_window_get_frame(mml_ctx* c, int f, int kind, float* out_xyzi, int cap, int* n_out) {
  if (!c || f < 0 || f >= win(c)->n_slots || kind < 0 || kind > 1) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  WinSlot& s = win(c)->slot[f];
  const int n = kind == 0 ? s.n_corner : s.n_surf;
  if (n_out) *n_out = n;
  if (out_xyzi && n) {
    if (cap < n) return MML_ERR_CAPACITY;
    MML_CUDA(c, cudaMemcpyAsync(out_xyzi, (kind == 0 ? s.q_corner : s.q_surf).p, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    MML_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return MML_OK;
}

// Estimator::Estimate on the frames currently in the window (W = mml_window_size). states: W x 16 doubles
// (P, q_wxyz, V, bg, ba), in place. preints[f] (f >= 1) links frame f-1 to f. stats (optional, 16 doubles):
// [outer, inner_total, n_line(last frame), n_plane(last frame), final_cost, min_sv(last frame), degenerate, evals].
int mml_estimate_window(mml_ctx* c, double* states, const mml_preint* const* preints, const double* exTlb16,
                        const double* gravity3, const mml_est_params* prm, double* stats) {
  if (!c || !states || !exTlb16 || !gravity3) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  WindowState* w = win(c);
  const int W = w->n_slots;
  if (W < 1) return mml_fail(c, MML_ERR_STATE, "estimate_window: no frame in the window");
  for (int f = 1; f < W; f++) if (!preints || !preints[f]) return mml_fail(c, MML_ERR_INVALID, "estimate_window: missing pre-integration");
  bool any_map = false;
  for (int k = 0; k < 4; k++) any_map = any_map || c->maps[k].valid;
  if (!any_map) return mml_fail(c, MML_ERR_STATE, "estimate_window: no feature map set");
  mml_est_params def;
  mml_est_params_default(&def);
  if (!prm) prm = &def;
  cudaStream_t st = c->stream;

  // exRbl = R^T, exPbl = -R^T t (EST.cpp:1155-1156); the functors re-normalise the rotation through a quaternion (CF.h:405-408)
  double Rbl_raw[9], Pbl[3];
  for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) Rbl_raw[3 * r + k] = exTlb16[4 * k + r];
  for (int r = 0; r < 3; r++) Pbl[r] = -1.0 * (Rbl_raw[3 * r] * exTlb16[3] + Rbl_raw[3 * r + 1] * exTlb16[7] + Rbl_raw[3 * r + 2] * exTlb16[11]);
  double Rbl_q[9];
  quat_to_R(quat_from_R9(Rbl_raw), Rbl_q);

  const int gmax = mml_accumulate_window_grid_max();
  MML_CHECK(win_prepare(c, w));
  MML_CUDA(c, w->partials.reserve(sizeof(double) * 28 * (size_t)gmax * kMaxWindow + 256));
  MML_CUDA(c, w->out.reserve(sizeof(double) * 28 * kMaxWindow + 64));
  double* out_dev = w->out.as<double>();
  unsigned* ticket_dev = reinterpret_cast<unsigned*>(w->partials.as<double>() + 28 * (size_t)gmax * kMaxWindow);
  double* host = static_cast<double*>(w->mapped);                 // [0, 28 W) sums, [28 kMaxWindow, ...) statistics
  volatile unsigned* host_seq = reinterpret_cast<volatile unsigned*>(host + kMapDoubles);
  unsigned* host_seq_dev = reinterpret_cast<unsigned*>(w->mapped_dev + kMapDoubles);

  const int n = (W == 1) ? 6 : 15 * W;  // a lone frame's velocity / bias block has no residual (Ceres drops it)
  double thres = prm->thres0;
  const double huber_a = prm->use_huber ? 0.1 / prm->lidar_m : 0.0;
  int outer_done = 0, inner_total = 0, evals = 0, n_line_last = 0, n_plane_last = 0, is_degenerate = 0;
  double final_cost = 0, min_sv = -1;
  std::vector<double> x(15 * W), H((size_t)n * n), g(n), Himu((size_t)n * n), gimu(n);
  DoglegN D;
  const float4* fl[kMaxWindow];
  const float4* fp[kMaxWindow];
  int nl[kMaxWindow], np[kMaxWindow];

  for (int it = 0; it < prm->max_outer; ++it) {
    // vector2double, EST.cpp:937-950
    for (int f = 0; f < W; f++) {
      const double* s = states + 16 * f;
      for (int k = 0; k < 3; k++) x[6 * f + k] = s[k];
      so3_log(Quat{s[3], s[4], s[5], s[6]}, &x[6 * f + 3]);
      for (int k = 0; k < 9; k++) x[6 * W + 9 * f + k] = s[7 + k];
    }
    double* sb = states + 16 * (W - 1);
    const Quat q_before = {sb[3], sb[4], sb[5], sb[6]};
    const double t_before[3] = {sb[0], sb[1], sb[2]};
    // association of every frame (EST.cpp:1265-1299). The reference joins its two association threads frame by
    // frame; the frames do not depend on each other, so here all 2 W kernels run side by side on their own streams.
    const double tp0 = g_prof_on ? now_us() : 0;
    MML_CUDA(c, cudaEventRecord(w->fork, st));
    for (int f = 0; f < W; f++) {
      WinSlot& s = w->slot[f];
      const double* sf = states + 16 * f;
      double Rq[9], T[16] = {0};
      quat_to_R(Quat{sf[3], sf[4], sf[5], sf[6]}, Rq);
      for (int r = 0; r < 3; r++) {
        for (int k = 0; k < 3; k++) T[4 * r + k] = Rq[3 * r] * Rbl_raw[k] + Rq[3 * r + 1] * Rbl_raw[3 + k] + Rq[3 * r + 2] * Rbl_raw[6 + k];
        T[4 * r + 3] = Rq[3 * r] * Pbl[0] + Rq[3 * r + 1] * Pbl[1] + Rq[3 * r + 2] * Pbl[2] + sf[r];
      }
      T[15] = 1;
      // the association works on the context's frame slot and statistics: lend it this frame's buffers
      auto lend = [&]() {
        std::swap(c->q_corner, s.q_corner); std::swap(c->q_surf, s.q_surf);
        std::swap(c->f_line, s.f_line); std::swap(c->f_plane, s.f_plane);
        std::swap(c->assoc_stats, s.assoc_stats);
        std::swap(c->assoc_part[0], s.assoc_part[0]); std::swap(c->assoc_part[1], s.assoc_part[1]);
      };
      lend();
      c->has_perm[0] = c->has_perm[1] = false;
      int rc = MML_OK;
      for (int kind = 1; kind >= 0 && rc == MML_OK; kind--) {
        cudaStream_t fs = w->fstream[f][kind];
        if (cudaStreamWaitEvent(fs, w->fork, 0) != cudaSuccess) { rc = MML_ERR_CUDA; break; }
        c->stream = fs;
        rc = mml_associate_launch(c, kind, T, (float)thres, nullptr, nullptr, nullptr, nullptr, kind ? s.n_surf : s.n_corner);
        c->stream = st;
        if (rc == MML_OK && cudaEventRecord(w->fev[f][kind], fs) != cudaSuccess) rc = MML_ERR_CUDA;
      }
      lend();
      MML_CHECK(rc);
      fl[f] = s.f_line.as<float4>(); fp[f] = s.f_plane.as<float4>();
      nl[f] = s.n_corner; np[f] = s.n_surf;
    }
    for (int f = 0; f < W; f++) for (int kind = 0; kind < 2; kind++) MML_CUDA(c, cudaStreamWaitEvent(st, w->fev[f][kind], 0));
    for (int f = 0; f < W; f++)  // both kinds of a frame write into the same 160-byte block (line: [0, 8), plane: [8, 16), counts)
      MML_CUDA(c, cudaMemcpyAsync(host + 28 * kMaxWindow + 20 * f, w->slot[f].assoc_stats.p, sizeof(double) * 20, cudaMemcpyDeviceToHost, st));
    thres = (it == 0) ? prm->thres1 : prm->thres2;  // EST.cpp:1377-1381
    if (g_prof_on) g_prof.assoc += now_us() - tp0;

    D.begin(n, x.data(), prm->max_inner);
    bool stats_read = false;
    while (!D.done) {
      const double* xe = D.eval_point();
      const unsigned seq = ++w->seq;
      const double te0 = g_prof_on ? now_us() : 0;
      MML_CHECK(mml_accumulate_window_launch(c, W, fl, fp, nl, np, xe, Rbl_q, Pbl, prm->lidar_m, prm->plan_weight_tan, huber_a,
                                             w->partials.as<double>(), ticket_dev, out_dev, w->mapped_dev, host_seq_dev, seq));
      // while the device evaluates the lidar terms, the host evaluates the IMU factors at the same point
      const double te1 = g_prof_on ? now_us() : 0;
      double cost_imu = 0;
      std::fill(Himu.begin(), Himu.end(), 0.0);
      std::fill(gimu.begin(), gimu.end(), 0.0);
      for (int f = 1; f < W; f++) {
        double r[15], J[450];
        imu_factor_eval(*preints[f], gravity3, xe + 6 * (f - 1), xe + 6 * W + 9 * (f - 1), xe + 6 * f, xe + 6 * W + 9 * f, r, J);
        for (int k = 0; k < 15; k++) cost_imu += 0.5 * r[k] * r[k];
        int gidx[30];
        {
          const int off[4] = {6 * (f - 1), 6 * W + 9 * (f - 1), 6 * f, 6 * W + 9 * f};
          const int sz[4] = {6, 9, 6, 9};
          int q = 0;
          for (int b = 0; b < 4; b++) for (int i = 0; i < sz[b]; i++) gidx[q++] = off[b] + i;
        }
        double JtJ[900];
        for (int i = 0; i < 30; i++) {
          double sg = 0;
          for (int k = 0; k < 15; k++) sg += J[k * 30 + i] * r[k];
          gimu[gidx[i]] += sg;
          for (int j = i; j < 30; j++) {
            double sh = 0;
            for (int k = 0; k < 15; k++) sh += J[k * 30 + i] * J[k * 30 + j];
            JtJ[i * 30 + j] = sh;
          }
        }
        for (int i = 0; i < 30; i++) for (int j = i; j < 30; j++) {
          Himu[(size_t)gidx[i] * n + gidx[j]] += JtJ[i * 30 + j];
          if (j != i) Himu[(size_t)gidx[j] * n + gidx[i]] += JtJ[i * 30 + j];
        }
      }
      // wait for the device: the kernel's last CTAs publish `seq` after their sums (system-scope fence)
      const double te2 = g_prof_on ? now_us() : 0;
      {
        long spins = 0;
        bool ready = false;
        while (!ready) {
          ready = true;
          for (int f = 0; f < W; f++) if (host_seq[f] != seq) { ready = false; break; }
          if (!ready && (++spins & 0xFFFF) == 0) {
            cudaError_t e = cudaStreamQuery(st);
            if (e != cudaSuccess && e != cudaErrorNotReady) { c->err = std::string("window evaluation: ") + cudaGetErrorString(e); return MML_ERR_CUDA; }
            if (e == cudaSuccess) {  // stream drained: the flags must be there now
              bool all = true;
              for (int f = 0; f < W; f++) all = all && host_seq[f] == seq;
              if (!all) return mml_fail(c, MML_ERR_CUDA, "window evaluation finished without publishing its result");
            }
          }
        }
        __sync_synchronize();
      }
      const double te3 = g_prof_on ? now_us() : 0;
      evals++;
      if (!stats_read) {
        stats_read = true;
        for (int f = 0; f < W; f++) {
          const double* a = host + 28 * kMaxWindow + 20 * f;
          const int* ints = reinterpret_cast<const int*>(a + 16);
          const double* m = a + 8;
          const double M[9] = {m[0], m[1], m[2], m[1], m[3], m[4], m[2], m[4], m[5]};
          double sv = -1.0;
          if (ints[1] > 10) sv = sqrt(fmax(eig3_sym_min(M), 0.0));  // checkLocalizability, EST.cpp:536-565
          if (sv < 3.0) is_degenerate = 1;                         // EST.cpp:771-775
          if (f == W - 1) { n_line_last = ints[0]; n_plane_last = ints[1]; min_sv = sv; }
        }
      }
      // assemble: lidar blocks from the device sums + the IMU part
      double cost = cost_imu;
      H = Himu;
      g = gimu;
      for (int f = 0; f < W; f++) {
        const double* o = host + 28 * f;
        cost += o[0];
        for (int i = 0; i < 6; i++) g[6 * f + i] += o[1 + i];
        int k = 7;
        for (int i = 0; i < 6; i++) for (int j = i; j < 6; j++, k++) {
          H[(size_t)(6 * f + i) * n + 6 * f + j] += o[k];
          if (j != i) H[(size_t)(6 * f + j) * n + 6 * f + i] += o[k];
        }
      }
      D.feed(cost, H.data(), g.data());
      if (g_prof_on) {
        const double te4 = now_us();
        g_prof.launch += te1 - te0; g_prof.imu += te2 - te1; g_prof.wait += te3 - te2; g_prof.solve += te4 - te3; g_prof.evals++;
      }
    }
    inner_total += D.iterations;
    final_cost = D.min_cost;
    // double2vector, EST.cpp:952-964
    for (int f = 0; f < W; f++) {
      double* s = states + 16 * f;
      for (int k = 0; k < 3; k++) s[k] = D.x_best[6 * f + k];
      const Quat q = so3_exp(&D.x_best[6 * f + 3]);
      s[3] = q.w; s[4] = q.x; s[5] = q.y; s[6] = q.z;
      if (W > 1) for (int k = 0; k < 9; k++) s[7 + k] = D.x_best[6 * W + 9 * f + k];
    }
    outer_done = it + 1;
    const Quat Q = {sb[3], sb[4], sb[5], sb[6]};
    const Quat dq = quat_mul(q_before, Quat{Q.w, -Q.x, -Q.y, -Q.z});
    const double deltaR = 2.0 * atan2(sqrt((dq.x * dq.x + dq.y * dq.y) + dq.z * dq.z), fabs(dq.w)) * 180.0 / 3.14159265358979323846;
    const double d0 = t_before[0] - sb[0], d1 = t_before[1] - sb[1], d2 = t_before[2] - sb[2];
    const double deltaT = sqrt((d0 * d0 + d1 * d1) + d2 * d2);
    if ((deltaR < 0.05 && deltaT < 0.05) || (it + 1) == prm->max_outer) break;
  }
  if (stats) {
    stats[0] = outer_done; stats[1] = inner_total; stats[2] = n_line_last; stats[3] = n_plane_last;
    stats[4] = final_cost; stats[5] = min_sv; stats[6] = is_degenerate; stats[7] = evals;
  }
  return MML_OK;
}

// ---- odometry loop with an IMU-initialised sliding window: the per-scan body of process() in its
// LidarIMUInited branch, src/unionPoseEstimation.cpp:796-891, with WINDOWSIZE frames kept (PE.cpp:830-832):
//   pre-integrate the IMU samples of (t_{k-1}, t_k] at the previous frame's biases      PE.cpp:807-809
//   predict the new frame's state from the previous (optimised) one                     PE.cpp:811-820
//   motion of the LiDAR over the sweep -> RemoveLidarDistortion                         PE.cpp:822-829, 862
//   push the frame, drop the oldest beyond the window                                   PE.cpp:830-832
//   EstimateLidarPose on the window                                                     PE.cpp:872
//   the odometry output is the OLDEST frame of the window (EST.cpp:1043-1049, PE.cpp:879-880)
// xyzi / line / s: per-scan pointers (device when host_buffers == 0). imu_t / imu_gyr / imu_acc: all samples
// concatenated, scan k owns imu_n[k] of them. state0: state of the frame before the first scan (t = stamp0).
// poses_front / poses_newest (n_scans x 16 row-major T_wb, either may be NULL), states_out (n_scans x 16, newest
// frame's state after each solve, may be NULL), stats_out (n_scans x 8, may be NULL). The feature maps are the
// ones set on the context (they are not updated by this call).
int mml_odom_run_window(mml_ctx* c, const void* const* xyzi, const void* const* line, const void* const* s,
                        const int* n_pts, int n_scans, int n_lines, int host_buffers, int window, const double* stamps,
                        double stamp0, const double* imu_t, const double* imu_gyr, const double* imu_acc, const int* imu_n,
                        const double* state0, const double* exTlb16, const double* gravity3, float leaf_corner,
                        float leaf_surf, const mml_est_params* prm, double* poses_front, double* poses_newest,
                        double* states_out, double* stats_out, float* total_ms) {
  if (!c || n_scans < 0 || window < 1 || window > kMaxWindow || !state0 || !exTlb16 || !gravity3 || !stamps) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  MML_CHECK(mml_window_reset(c));
  double Rbl[9], Pbl[3];
  for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) Rbl[3 * r + k] = exTlb16[4 * k + r];
  for (int r = 0; r < 3; r++) Pbl[r] = -1.0 * (Rbl[3 * r] * exTlb16[3] + Rbl[3 * r + 1] * exTlb16[7] + Rbl[3 * r + 2] * exTlb16[11]);
  std::vector<double> states;           // frames in the window, 16 doubles each
  std::vector<mml_preint> pre_store;    // same indexing
  double prev[16];
  memcpy(prev, state0, sizeof(prev));
  double t_prev = stamp0;
  size_t imu_off = 0;
  WindowState* w = win(c);
  MML_CHECK(win_prepare(c, w));
  if (total_ms) { MML_CUDA(c, cudaEventRecord(c->ev0, c->stream)); }
  // labels do not depend on the pose (the extraction node of the reference runs ahead of the estimator, FE.cpp ->
  // /union_feature_cloud -> PE.cpp): scan k+1 is copied and labelled on its own stream while scan k is solved
  const void* xdv[2] = {nullptr, nullptr};
  const void* ldv[2] = {nullptr, nullptr};
  const void* sdv[2] = {nullptr, nullptr};
  auto prefetch = [&](int k) -> int {
    const int b = k & 1;
    const size_t n = (size_t)n_pts[k];
    cudaStream_t xs = w->xstream;
    MML_CUDA(c, cudaStreamWaitEvent(xs, w->xfree[b], 0));  // the previous user of this buffer pair (scan k-2) is done with it
    xdv[b] = xyzi[k]; ldv[b] = line[k]; sdv[b] = s ? s[k] : nullptr;
    if (host_buffers) {
      MML_CUDA(c, w->x_xyzi[b].reserve(sizeof(float4) * (n + 1)));
      MML_CUDA(c, w->x_line[b].reserve(sizeof(uint16_t) * (n + 1)));
      MML_CUDA(c, w->x_s[b].reserve(sizeof(float) * (n + 1)));
      MML_CUDA(c, cudaMemcpyAsync(w->x_xyzi[b].p, xyzi[k], sizeof(float4) * n, cudaMemcpyHostToDevice, xs));
      MML_CUDA(c, cudaMemcpyAsync(w->x_line[b].p, line[k], sizeof(uint16_t) * n, cudaMemcpyHostToDevice, xs));
      if (sdv[b]) MML_CUDA(c, cudaMemcpyAsync(w->x_s[b].p, s[k], sizeof(float) * n, cudaMemcpyHostToDevice, xs));
      xdv[b] = w->x_xyzi[b].p; ldv[b] = w->x_line[b].p; sdv[b] = sdv[b] ? w->x_s[b].p : nullptr;
    }
    MML_CUDA(c, w->x_label[b].reserve(n + 16));
    MML_CUDA(c, w->x_counters[b].reserve(sizeof(int) * 80));
    const int off[2] = {0, (int)n};
    cudaStream_t keep = c->stream;
    c->stream = xs;
    c->counters_alt = w->x_counters[b].as<int>();
    const int rc = mml_extract_device(c, (const float4*)xdv[b], (const uint16_t*)ldv[b], off, 1, n_lines, w->x_label[b].as<uint8_t>(), false);
    c->counters_alt = nullptr;
    c->stream = keep;
    MML_CHECK(rc);
    MML_CUDA(c, cudaEventRecord(w->xev[b], xs));
    return MML_OK;
  };
  if (n_scans > 0) MML_CHECK(prefetch(0));
  auto pose16 = [](const double* st, double* T) {
    double R[9];
    quat_to_R(Quat{st[3], st[4], st[5], st[6]}, R);
    const double Tn[16] = {R[0], R[1], R[2], st[0], R[3], R[4], R[5], st[1], R[6], R[7], R[8], st[2], 0, 0, 0, 1};
    memcpy(T, Tn, sizeof(Tn));
  };
  for (int k = 0; k < n_scans; k++) {
    mml_preint pre;
    MML_CHECK(mml_imu_preintegrate(imu_t + imu_off, imu_gyr + 3 * imu_off, imu_acc + 3 * imu_off, imu_n[k], t_prev, prev + 10, prev + 13, &pre));
    imu_off += imu_n[k];
    double next[16];
    MML_CHECK(mml_imu_predict(prev, &pre, next));
    // LiDAR motion over the sweep, PE.cpp:822-829: delta = T_wl(prev)^-1 T_wl(predicted)
    double Tp[16], Tn[16], Twl_p[16], Twl_n[16], Tbl_h[16] = {0}, inv[16], dT[16];
    pose16(prev, Tp); pose16(next, Tn);
    for (int r = 0; r < 3; r++) { for (int q = 0; q < 3; q++) Tbl_h[4 * r + q] = Rbl[3 * r + q]; Tbl_h[4 * r + 3] = Pbl[r]; }
    Tbl_h[15] = 1;
    mat4_mul(Tp, Tbl_h, Twl_p); mat4_mul(Tn, Tbl_h, Twl_n);
    rigid_inv(Twl_p, inv);
    mat4_mul(inv, Twl_n, dT);
    const double dR9[9] = {dT[0], dT[1], dT[2], dT[4], dT[5], dT[6], dT[8], dT[9], dT[10]};
    const double dt3[3] = {dT[3], dT[7], dT[11]};
    const int b = k & 1;
    const void *xd = xdv[b], *ld = ldv[b], *sd = sdv[b];
    MML_CUDA(c, cudaStreamWaitEvent(c->stream, w->xev[b], 0));
    const double tq0 = g_prof_on ? now_us() : 0;
    if ((int)(states.size() / 16) >= window) {  // PE.cpp:830-832
      states.erase(states.begin(), states.begin() + 16);
      pre_store.erase(pre_store.begin());
    }
    MML_CHECK(window_push_scan(c, xd, ld, sd, n_pts[k], n_lines, dR9, dt3, leaf_corner, leaf_surf, window, nullptr,
                               w->x_label[b].as<uint8_t>(), w->x_counters[b].as<int>()));
    MML_CUDA(c, cudaEventRecord(w->xfree[b], c->stream));  // raw scan, labels and counters of this buffer pair are consumed
    if (k + 1 < n_scans) MML_CHECK(prefetch(k + 1));
    if (g_prof_on) { g_prof.push += now_us() - tq0; g_prof.scans++; }
    states.insert(states.end(), next, next + 16);
    pre_store.push_back(pre);
    const int W = (int)(states.size() / 16);
    const mml_preint* pp[kMaxWindow] = {nullptr, nullptr, nullptr, nullptr};
    for (int f = 1; f < W; f++) pp[f] = &pre_store[f];
    double st8[16];
    MML_CHECK(mml_estimate_window(c, states.data(), pp, exTlb16, gravity3, prm, st8));
    memcpy(prev, &states[16 * (W - 1)], sizeof(prev));
    t_prev = stamps[k];
    if (poses_front) pose16(&states[0], poses_front + 16 * (size_t)k);
    if (poses_newest) pose16(prev, poses_newest + 16 * (size_t)k);
    if (states_out) memcpy(states_out + 16 * (size_t)k, prev, sizeof(prev));
    if (stats_out) memcpy(stats_out + 8 * (size_t)k, st8, sizeof(double) * 8);
  }
  if (total_ms) {
    MML_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    MML_CUDA(c, cudaEventSynchronize(c->ev1));
    MML_CUDA(c, cudaEventElapsedTime(total_ms, c->ev0, c->ev1));
  }
  if (g_prof_on && g_prof.scans) {
    const double ns = (double)g_prof.scans, ne = (double)(g_prof.evals ? g_prof.evals : 1);
    fprintf(stderr, "[mml window prof] per scan: push %.1f us, association enqueue %.1f us, evaluations %.1f; per evaluation: launch %.1f, imu %.1f, wait %.1f, assemble+dogleg %.1f us\n",
            g_prof.push / ns, g_prof.assoc / ns, g_prof.evals / ns, g_prof.launch / ne, g_prof.imu / ne, g_prof.wait / ne, g_prof.solve / ne);
    g_prof = WinProf();
  }
  return MML_OK;
}

}  // extern "C"
