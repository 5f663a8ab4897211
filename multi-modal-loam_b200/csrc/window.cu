// mmloam_b200: sliding-window Estimate (window sizes 2-4: IMU factors, no marginalisation; BASELINE config 3 is
// window 3). Reference: Estimator::Estimate, src/lio/Estimator.cpp:1143-1581 (IMU blocks 1235-1254, per-frame
// association 1265-1299, lidar blocks 1377-1418, solve 1425-1432, convergence 1441-1450);
// IMUIntegrator::PreIntegration, src/lio/IMUIntegrator.cpp:105-166; Cost_NavState_PRV_Bias,
// include/utils/ceresfunc.h:321-393; pose prediction of process(), src/unionPoseEstimation.cpp:796-835.
//
// Split of the work (SURVEY.md §2 rows 11-12, §8 f F3): the whole solve of a window runs on the device
// (csrc/windowsolve.cu: association of every frame against the resident maps, lidar residuals / Jacobians / Huber /
// 28-sum reduction, the W - 1 IMU factors under forward-mode differentiation, the (15 W)-dim dogleg). This file is the
// host side: IMU pre-integration, prediction and the factor as host API (mml_imu_*), the frame slots of the window,
// the per-call entry points (mml_window_push_frame, mml_estimate_window) and the odometry loop (mml_odom_run_window:
// per scan one pre-integration, one graph launch, one wait on the mapped result, the map update at its gate).
#include "common.cuh"
#include "smallmath.cuh"
#include "eststate.cuh"
#include "imufactor.cuh"
#include "windowstate.cuh"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

using namespace mml;

int mml_associate_launch(mml_ctx* ctx, int kind, const double* T16, float thres, const double* T_dev,
                         const float* thres_dev, const int* gate, const int* nq_dev, int cap);
int mml_extract_device(mml_ctx* ctx, const float4* xyzi_d, const uint16_t* line_d, const int* scan_off, int n_scans,
                       int n_lines, uint8_t* label_d, bool sequential);
int mml_split_voxel_capacity();
int mml_label_compact_device(mml_ctx* ctx, const uint8_t* label_d, int n, int* idx0, int* idx1, int* cnt_d);
extern "C" int mml_local_map_push_dev(mml_ctx* c, const void* corner_dev, int n_corner, const void* surf_dev, int n_surf, const double* T_wl16,
                                      float leaf_corner, float leaf_surf, int clear_first, int* n_corner_map, int* n_surf_map);
int mml_split_voxel_device(mml_ctx* ctx, const float4* pts_d, const float* s_d, const uint8_t* label_d, int n,
                           const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, float4* corner_out,
                           float4* surf_out, int* counts_d, const mml::SvChain* chain = nullptr);

namespace {

// MML_WIN_PROF=1: wall-clock split of the window loop (debug aid; printed by mml_odom_run_window)
struct WinProf { double push = 0, assoc = 0, launch = 0, imu = 0, wait = 0, solve = 0, other = 0; long evals = 0, scans = 0; };
WinProf g_prof;
#ifdef MML_WIN_TIMELINE
extern "C" int mml_debug_win_timeline(unsigned long long* out8);
#endif
const bool g_prof_on = getenv("MML_WIN_PROF") != nullptr;
inline double now_us() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

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
  if (!w->pin_up) MML_CUDA(c, cudaHostAlloc(&w->pin_up, kPinUpBytes, cudaHostAllocDefault));
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
}  // namespace

void mml_window_destroy(mml_ctx* c) {
  if (!c->window) return;
  WindowState* w = static_cast<WindowState*>(c->window);
  for (auto& s : w->slot) {
    s.q_corner.release(); s.q_surf.release(); s.f_line.release(); s.f_plane.release();
    s.assoc_stats.release(); s.assoc_part[0].release(); s.assoc_part[1].release(); s.cnt.release();
  }
  w->dev.release(); w->push_dev.release();
  if (w->graph) cudaGraphExecDestroy(w->graph);
  if (w->mapped) cudaFreeHost(w->mapped);
  if (w->pin_up) cudaFreeHost(w->pin_up);
  if (w->streams_ok) {
    for (int f = 0; f < kMaxWindow; f++) for (int k = 0; k < 2; k++) { cudaStreamDestroy(w->fstream[f][k]); cudaEventDestroy(w->fev[f][k]); }
    cudaEventDestroy(w->fork);
    cudaStreamDestroy(w->xstream);
    for (int k = 0; k < 2; k++) { cudaEventDestroy(w->xev[k]); cudaEventDestroy(w->xfree[k]); }
  }
  for (int k = 0; k < 2; k++) { w->x_xyzi[k].release(); w->x_line[k].release(); w->x_s[k].release(); w->x_label[k].release(); w->x_counters[k].release(); w->x_idx[k].release(); }
  delete w;
  c->window = nullptr;
}

extern "C" {

// IMUIntegrator::PreIntegration, IMU.cpp:105-166 (+ sqrt_information of EST.cpp:1240-1242). Host side: ~20 samples
// of 15x15 algebra per scan.
}  // extern "C" (helpers with internal linkage follow)

// The mean of the pre-integration (dq, dp, dv, dt) does not depend on its covariance / Jacobian. The odometry loop
// needs the mean first (it predicts the new frame's pose, which the undistortion of the scan waits for) and the rest
// only when the window is pushed, so the mean can be formed alone (mml_imu_preintegrate_mean) and the full pre-integration afterwards,
// while the device already works on the scan. Both run the SAME compiled code for the mean (the helpers below are
// not inlined), so the two results are identical to the bit.
namespace {
__attribute__((noinline)) void imu_sample_terms(const double* gyr, const double* acc, const double* t, int s, double current_time,
                                                const double* bg3, const double* ba3, double* g3, double* a3, double* dt_out,
                                                double* gdt, double* dR) {
  const double gnorm = 9.805;  // IMU.h:84
  g3[0] = gyr[3 * s] - bg3[0]; g3[1] = gyr[3 * s + 1] - bg3[1]; g3[2] = gyr[3 * s + 2] - bg3[2];
  a3[0] = acc[3 * s] * gnorm - ba3[0]; a3[1] = acc[3 * s + 1] * gnorm - ba3[1]; a3[2] = acc[3 * s + 2] * gnorm - ba3[2];
  const double dt = t[s] - current_time;
  *dt_out = dt;
  gdt[0] = g3[0] * dt; gdt[1] = g3[1] * dt; gdt[2] = g3[2] * dt;
  quat_to_R(so3_exp(gdt), dR);
}
__attribute__((noinline)) void imu_rotation_of(const Quat& dq, double* Rq) { quat_to_R(dq, Rq); }
__attribute__((noinline)) void imu_mean_update(Quat& dq, double* dp, double* dv, const double* Rq, const double* a3, double dt,
                                               const double* dR) {
  const double dt2 = dt * dt;
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
}
}  // namespace

extern "C" {

int mml_imu_preintegrate(const double* t, const double* gyr, const double* acc, int n, double last_time,
                         const double* bg3, const double* ba3, mml_preint* out) {
  if (!out || n < 0 || (n && (!t || !gyr || !acc)) || !bg3 || !ba3) return MML_ERR_INVALID;
  const double acc_n = 0.08, gyr_n = 0.004, acc_w = 2.0e-4, gyr_w = 2.0e-5;  // IMU.h:79-83
  Quat dq = {1, 0, 0, 0};
  double dp[3] = {0, 0, 0}, dv[3] = {0, 0, 0}, dtime = 0;
  std::vector<double> cov(225, 0.0), jac(225, 0.0), noise(144, 0.0), tmp(225), tmp2(225), AT(225), BNB(225);
  for (int i = 0; i < 15; i++) jac[i * 15 + i] = 1.0;
  for (int i = 0; i < 3; i++) {
    noise[i * 12 + i] = gyr_n * gyr_n; noise[(3 + i) * 12 + 3 + i] = acc_n * acc_n;
    noise[(6 + i) * 12 + 6 + i] = gyr_w * gyr_w; noise[(9 + i) * 12 + 9 + i] = acc_w * acc_w;
  }
  double current_time = last_time;
  for (int s = 0; s < n; s++) {
    double g3[3], a3[3], dt, gdt[3], dR[9], Rq[9];
    imu_sample_terms(gyr, acc, t, s, current_time, bg3, ba3, g3, a3, &dt, gdt, dR);
    imu_rotation_of(dq, Rq);
    const double dt2 = dt * dt;
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
    double Ha[9], RH[9];
    hat3(a3, Ha);
    mat_mul(3, 3, 3, Rq, Ha, RH);
    // A = d(state_{k+1}) / d(state_k) and B = d(state_{k+1}) / d(noise) (IMU.cpp:131-156) are identity plus a few
    // 3 x 3 blocks: the products A jac, A cov A^T and B N B^T are formed block-wise, every sum in the order of the
    // dense products (ascending inner index; the skipped terms are exact zeros)
    const double dRT[9] = {dR[0], dR[3], dR[6], dR[1], dR[4], dR[7], dR[2], dR[5], dR[8]};
    double A03[9], A012[9], A39[9], A63[9], A612[9];
    for (int i = 0; i < 9; i++) {
      A03[i] = -0.5 * dt2 * RH[i]; A012[i] = -0.5 * dt2 * Rq[i]; A39[i] = -dt * Jr[i];
      A63[i] = -dt * RH[i]; A612[i] = -dt * Rq[i];
    }
    // out = A M (rows of M in, rows of out out); M and out are 15 x 15 row-major and distinct
    auto applyA = [&](const double* M, double* out) {
      for (int j = 0; j < 15; j++) {
        for (int r = 0; r < 3; r++) {
          double v = M[r * 15 + j];
          for (int c = 0; c < 3; c++) v += A03[3 * r + c] * M[(3 + c) * 15 + j];
          v += dt * M[(6 + r) * 15 + j];
          for (int c = 0; c < 3; c++) v += A012[3 * r + c] * M[(12 + c) * 15 + j];
          out[r * 15 + j] = v;
          double w = 0;
          for (int c = 0; c < 3; c++) w += dRT[3 * r + c] * M[(3 + c) * 15 + j];
          for (int c = 0; c < 3; c++) w += A39[3 * r + c] * M[(9 + c) * 15 + j];
          out[(3 + r) * 15 + j] = w;
          double u = 0;
          for (int c = 0; c < 3; c++) u += A63[3 * r + c] * M[(3 + c) * 15 + j];
          u += M[(6 + r) * 15 + j];
          for (int c = 0; c < 3; c++) u += A612[3 * r + c] * M[(12 + c) * 15 + j];
          out[(6 + r) * 15 + j] = u;
        }
        for (int r = 9; r < 15; r++) out[r * 15 + j] = M[r * 15 + j];
      }
    };
    applyA(jac.data(), tmp.data());
    jac = tmp;
    // cov <- A cov A^T + B N B^T: T = A cov, then (T A^T)[i][j] = sum_t T[i][t] A[j][t] = (A T^T)[j][i]
    applyA(cov.data(), tmp.data());
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) AT[i * 15 + j] = tmp[j * 15 + i];
    applyA(AT.data(), tmp2.data());  // tmp2 = A T^T = (T A^T)^T
    // B N B^T: B has blocks (0,3) = Rq dt2/2, (3,0) = Jr dt, (6,3) = Rq dt, (9,6) = (12,9) = I dt; N is diagonal
    const double ng = noise[0], na = noise[3 * 12 + 3], nwg = noise[6 * 12 + 6], nwa = noise[9 * 12 + 9];
    std::fill(BNB.begin(), BNB.end(), 0.0);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
      double pp = 0, pv = 0, vp = 0, vv = 0, rr = 0;
      for (int k = 0; k < 3; k++) {
        const double bpi = 0.5 * dt2 * Rq[3 * i + k], bpj = 0.5 * dt2 * Rq[3 * j + k], bvi = dt * Rq[3 * i + k], bvj = dt * Rq[3 * j + k];
        pp += (bpi * na) * bpj; pv += (bpi * na) * bvj; vp += (bvi * na) * bpj; vv += (bvi * na) * bvj;
        rr += ((dt * Jr[3 * i + k]) * ng) * (dt * Jr[3 * j + k]);
      }
      BNB[i * 15 + j] = pp; BNB[i * 15 + 6 + j] = pv; BNB[(6 + i) * 15 + j] = vp; BNB[(6 + i) * 15 + 6 + j] = vv;
      BNB[(3 + i) * 15 + 3 + j] = rr;
    }
    for (int i = 0; i < 3; i++) { BNB[(9 + i) * 15 + 9 + i] = (dt * nwg) * dt; BNB[(12 + i) * 15 + 12 + i] = (dt * nwa) * dt; }
    for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) cov[i * 15 + j] = tmp2[j * 15 + i] + BNB[i * 15 + j];
    imu_mean_update(dq, dp, dv, Rq, a3, dt, dR);
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
int mml_imu_preintegrate_mean(const double* t, const double* gyr, const double* acc, int n, double last_time,
                              const double* bg3, const double* ba3, mml_preint* out) {
  if (!out || n < 0 || (n && (!t || !gyr || !acc)) || !bg3 || !ba3) return MML_ERR_INVALID;
  // the mean steps of mml_imu_preintegrate alone (the same non-inlined helpers: the same bits), without its work arrays
  Quat dq = {1, 0, 0, 0};
  double dp[3] = {0, 0, 0}, dv[3] = {0, 0, 0}, dtime = 0, current_time = last_time;
  for (int s = 0; s < n; s++) {
    double g3[3], a3[3], dt, gdt[3], dR[9], Rq[9];
    imu_sample_terms(gyr, acc, t, s, current_time, bg3, ba3, g3, a3, &dt, gdt, dR);
    imu_rotation_of(dq, Rq);
    imu_mean_update(dq, dp, dv, Rq, a3, dt, dR);
    dtime += dt;
    current_time = t[s];
  }
  out->dq[0] = dq.w; out->dq[1] = dq.x; out->dq[2] = dq.y; out->dq[3] = dq.z;
  for (int k = 0; k < 3; k++) { out->dp[k] = dp[k]; out->dv[k] = dv[k]; out->bg[k] = bg3[k]; out->ba[k] = ba3[k]; }
  out->dt = dtime;
  return MML_OK;
}

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
// Frames live in PHYSICAL slots (the association nodes of the solve graph are bound to them); order[f] maps frame f
// (0 = oldest) to its slot.
int mml_window_reset(mml_ctx* c) {
  if (!c) return MML_ERR_INVALID;
  win(c)->n_slots = 0;
  return MML_OK;
}
int mml_window_size(mml_ctx* c) { return c ? win(c)->n_slots : 0; }

static int window_make_room(mml_ctx* c, int max_frames) {
  WindowState* w = win(c);
  if (max_frames < 1 || max_frames > kMaxWindow) return mml_fail(c, MML_ERR_INVALID, "window size must be 1..4");
  while (w->n_slots >= max_frames) {  // drop the oldest frame (PE.cpp:830-832), its slot rotates to the back
    const int first = w->order[0];
    for (int f = 0; f + 1 < w->n_slots; f++) w->order[f] = w->order[f + 1];
    w->order[w->n_slots - 1] = first;
    w->n_slots--;
  }
  return MML_OK;
}

static int slot_reserve_queries(mml_ctx* c, WinSlot& s, int need) {
  if (s.cap < need) s.cap = (need + 4095) / 4096 * 4096;  // coarse steps: the solve graph is keyed on the capacities
  MML_CUDA(c, s.q_corner.reserve(sizeof(float4) * (size_t)(s.cap + 1)));
  MML_CUDA(c, s.q_surf.reserve(sizeof(float4) * (size_t)(s.cap + 1)));
  MML_CUDA(c, s.cnt.reserve(64));
  return MML_OK;
}

// push a frame given as host clouds (already undistorted and voxel-filtered: the A6 outputs)
int mml_window_push_frame(mml_ctx* c, const float* corner_xyzi, int n_corner, const float* surf_xyzi, int n_surf, int max_frames) {
  if (!c || n_corner < 0 || n_surf < 0) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  MML_CHECK(window_make_room(c, max_frames));
  WindowState* w = win(c);
  WinSlot& s = w->slot[w->order[w->n_slots]];
  MML_CHECK(slot_reserve_queries(c, s, (n_corner > n_surf ? n_corner : n_surf) + 1));
  if (n_corner) MML_CUDA(c, cudaMemcpyAsync(s.q_corner.p, corner_xyzi, sizeof(float4) * (size_t)n_corner, cudaMemcpyHostToDevice, c->stream));
  if (n_surf) MML_CUDA(c, cudaMemcpyAsync(s.q_surf.p, surf_xyzi, sizeof(float4) * (size_t)n_surf, cudaMemcpyHostToDevice, c->stream));
  const int cnt[8] = {n_corner, n_surf, 0, 0, 0, 0, 0, 0};
  MML_CUDA(c, cudaMemcpyAsync(s.cnt.p, cnt, sizeof(cnt), cudaMemcpyHostToDevice, c->stream));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  s.n_corner = n_corner; s.n_surf = n_surf;
  w->n_slots++;
  return MML_OK;
}

// push a raw scan resident in HBM: extraction (A1) -> undistortion + label split + voxel filter (A4, A6) on the
// device, the downsampled clouds become the newest window frame. out_counts (may be NULL): n_sharp, n_flat,
// n_corner_ds, n_surf_ds.
// pre_label / pre_counters: labels and extractor counters already produced for this scan (the loop labels scan k+1
// on its own stream while the window of scan k is solved; the caller has made c->stream wait for them).
// flags_async: pinned int[16] that receives the extractor counters [0, 3) and the split / voxel counts [4, 9) with
// an asynchronous copy instead of a synchronisation (the loop reads them after the scan's solve has been awaited).
static int window_push_scan(mml_ctx* c, const void* xyzi_dev, const void* line_id_dev, const void* s_dev, int n, int n_lines,
                            const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, int max_frames,
                            int* out_counts, const uint8_t* pre_label, const int* pre_counters, int* flags_async,
                            const int* pre_idx = nullptr) {
  if (!c || n < 0) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  MML_CHECK(window_make_room(c, max_frames));
  WindowState* w = win(c);
  WinSlot& s = w->slot[w->order[w->n_slots]];
  cudaStream_t st = c->stream;
  MML_CUDA(c, c->in_label.reserve((size_t)n + 16));
  MML_CHECK(slot_reserve_queries(c, s, mml_split_voxel_capacity()));
  MML_CUDA(c, c->pin_flags.reserve(64));
  const int off[2] = {0, n};
  const uint8_t* label_d = pre_label;
  const int* counters_d = pre_counters;
  if (!label_d) {
    MML_CHECK(mml_extract_device(c, (const float4*)xyzi_dev, (const uint16_t*)line_id_dev, off, 1, n_lines, c->in_label.as<uint8_t>(), false));
    label_d = c->in_label.as<uint8_t>();
    counters_d = c->counters.as<int>();
  }
  int* cnt = s.cnt.as<int>();  // [n_corner, n_surf, ., ., overflow]: the association and the solve read the counts here
  const bool undist = dR9 && dt3 && s_dev;
  if (pre_idx) {
    // the label split ran behind the extraction (k_label_compact): the cluster form of the fused undistortion + voxel
    // filter starts from the compacted indices, eight CTAs per label sharing the float64 slerp
    const int cap = mml_split_voxel_capacity();
    mml::SvChain ch;
    memset(&ch, 0, sizeof(ch));
    ch.pre_idx[0] = pre_idx; ch.pre_idx[1] = pre_idx + cap; ch.pre_cnt = pre_idx + 2 * cap;
    MML_CUDA(c, cudaMemsetAsync(cnt + 4, 0, sizeof(int), st));
    MML_CHECK(mml_split_voxel_device(c, (const float4*)xyzi_dev, undist ? (const float*)s_dev : nullptr, label_d, n,
                                     dR9, dt3, leaf_corner, leaf_surf, s.q_corner.as<float4>(), s.q_surf.as<float4>(), cnt, &ch));
  } else {
    MML_CHECK(mml_split_voxel_device(c, (const float4*)xyzi_dev, undist ? (const float*)s_dev : nullptr, label_d, n,
                                     dR9, dt3, leaf_corner, leaf_surf, s.q_corner.as<float4>(), s.q_surf.as<float4>(), cnt));
  }
  int* hf = flags_async ? flags_async : c->pin_flags.as<int>();
  MML_CUDA(c, cudaMemcpyAsync(hf, counters_d, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
  MML_CUDA(c, cudaMemcpyAsync(hf + 4, cnt, 5 * sizeof(int), cudaMemcpyDeviceToHost, st));
  s.n_corner = s.n_surf = -1;
  if (!flags_async) {
    MML_CUDA(c, cudaStreamSynchronize(st));
    if (hf[2] || hf[8]) return mml_fail(c, MML_ERR_CAPACITY, "scan exceeds the fused extraction / split-voxel capacities (window path)");
    s.n_corner = hf[4]; s.n_surf = hf[5];
    if (out_counts) { out_counts[0] = hf[0]; out_counts[1] = hf[1]; out_counts[2] = hf[4]; out_counts[3] = hf[5]; }
  }
  w->n_slots++;
  return MML_OK;
}

int mml_window_push_scan_dev(mml_ctx* c, const void* xyzi_dev, const void* line_id_dev, const void* s_dev, int n, int n_lines,
                             const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, int max_frames,
                             int* out_counts) {
  return window_push_scan(c, xyzi_dev, line_id_dev, s_dev, n, n_lines, dR9, dt3, leaf_corner, leaf_surf, max_frames, out_counts,
                          nullptr, nullptr, nullptr);
}

int mml_window_get_frame(mml_ctx* c, int f, int kind, float* out_xyzi, int cap, int* n_out) {
  if (!c || f < 0 || f >= win(c)->n_slots || kind < 0 || kind > 1) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  WinSlot& s = win(c)->frame(f);
  int cnt[2] = {0, 0};
  MML_CUDA(c, cudaMemcpyAsync(cnt, s.cnt.p, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  const int n = cnt[kind];
  if (n_out) *n_out = n;
  if (out_xyzi && n) {
    if (cap < n) return MML_ERR_CAPACITY;
    MML_CUDA(c, cudaMemcpyAsync(out_xyzi, (kind == 0 ? s.q_corner : s.q_surf).p, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    MML_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return MML_OK;
}

// solve parameters and extrinsics of a window solve into the head of a WinDev image
static void fill_win_params(WinDev* h, const double* exTlb16, const double* gravity3, const mml_est_params* prm) {
  // exRbl = R^T, exPbl = -R^T t (EST.cpp:1155-1156); the functors re-normalise the rotation through a quaternion (CF.h:405-408)
  for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) h->Rbl_raw[3 * r + k] = exTlb16[4 * k + r];
  for (int r = 0; r < 3; r++)
    h->Pbl[r] = -1.0 * (h->Rbl_raw[3 * r] * exTlb16[3] + h->Rbl_raw[3 * r + 1] * exTlb16[7] + h->Rbl_raw[3 * r + 2] * exTlb16[11]);
  quat_to_R(quat_from_R9(h->Rbl_raw), h->Rbl_q);
  for (int k = 0; k < 3; k++) h->gravity[k] = gravity3[k];
  h->max_outer = prm->max_outer; h->max_inner = prm->max_inner;
  h->lidar_m = prm->lidar_m; h->w_tan = prm->plan_weight_tan;
  h->huber_a = prm->use_huber ? 0.1 / prm->lidar_m : 0.0;
  h->thres_sched[0] = prm->thres0; h->thres_sched[1] = prm->thres1; h->thres_sched[2] = prm->thres2;  // EST.cpp:1207, 1377-1381
}

// wait for the solve that publishes `seq` in the mapped result block (the kernel's tail stores the result with a
// system-scope fence before the sequence word: no stream synchronisation on the way)
static int wait_window_result(mml_ctx* c, WindowState* w, unsigned seq) {
  volatile unsigned* host_seq = reinterpret_cast<volatile unsigned*>(static_cast<double*>(w->mapped) + kMapDoubles);
  long spins = 0;
  while (*host_seq != seq) {
    if ((++spins & 0x3FFF) == 0) {
      cudaError_t e = cudaStreamQuery(c->stream);
      if (e != cudaSuccess && e != cudaErrorNotReady) { c->err = std::string("window solve: ") + cudaGetErrorString(e); return MML_ERR_CUDA; }
      if (e == cudaSuccess && *host_seq != seq) return mml_fail(c, MML_ERR_CUDA, "window solve finished without publishing its result");
    }
  }
  __sync_synchronize();
  return MML_OK;
}

static int window_cap(WindowState* w) {
  int cap = 1;
  for (int p = 0; p < kMaxWindow; p++) cap = cap > w->slot[p].cap ? cap : w->slot[p].cap;
  return cap;
}

// Estimator::Estimate on the frames currently in the window (W = mml_window_size). states: W x 16 doubles
// (P, q_wxyz, V, bg, ba), in place. preints[f] (f >= 1) links frame f-1 to f. stats (optional, 16 doubles):
// [outer, inner_total, n_line(last frame), n_plane(last frame), final_cost, min_sv(last frame), degenerate, evals].
// The whole solve runs on the device (windowsolve.cu): one graph launch, one wait.
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
  MML_CHECK(win_prepare(c, w));
  MML_CHECK(mml_window_solve_graph(c, w, window_cap(w)));
  WinDev* h = static_cast<WinDev*>(w->pin_up);
  memset(h, 0, offsetof(WinDev, upload_end));
  h->W = W;
  for (int f = 0; f < kMaxWindow; f++) h->slot_of[f] = w->order[f];
  memcpy(h->states, states, sizeof(double) * 16 * (size_t)W);
  for (int f = 1; f < W; f++) h->pre[f] = *preints[f];
  fill_win_params(h, exTlb16, gravity3, prm);
  const unsigned seq = ++w->seq;
  h->seq = seq;
  MML_CUDA(c, cudaMemcpyAsync(w->dev.p, h, offsetof(WinDev, upload_end), cudaMemcpyHostToDevice, st));
  MML_CHECK(mml_window_begin_launch(c, w));
  MML_CHECK(mml_window_solve_launch(c, w, window_cap(w), prm->max_outer));
  MML_CHECK(wait_window_result(c, w, seq));
  const double* out = static_cast<const double*>(w->mapped);
  memcpy(states, out, sizeof(double) * 16 * (size_t)W);
  const double* so = out + 16 * kMaxWindow;
  c->launches += w->graph_launches * (long long)so[0];
  if (stats) for (int k = 0; k < 8; k++) stats[k] = so[k];
  MML_CUDA(c, cudaStreamSynchronize(st));  // per-call API: leave the stream idle (tail of the graph: gated no-op nodes)
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
// The window (states, pre-integrations, clouds) lives on the device between scans; per scan the host pre-integrates
// the IMU samples at the biases the previous solve returned (~20 samples of 15 x 15 algebra), enqueues
// [undistortion + voxel filter of the new frame -> window shift -> solve graph] and waits once, on the mapped
// result of the solve. Extraction of scan k+1 runs on its own stream meanwhile.
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
  bool any_map = false;
  for (int k = 0; k < 4; k++) any_map = any_map || c->maps[k].valid;
  if (!any_map) return mml_fail(c, MML_ERR_STATE, "odom_run_window: no feature map set");
  mml_est_params def;
  mml_est_params_default(&def);
  if (!prm) prm = &def;
  MML_CHECK(mml_window_reset(c));
  WindowState* w = win(c);
  MML_CHECK(win_prepare(c, w));
  cudaStream_t st = c->stream;
  for (int p = 0; p < kMaxWindow; p++) MML_CHECK(slot_reserve_queries(c, w->slot[p], mml_split_voxel_capacity()));
  MML_CHECK(mml_window_solve_graph(c, w, window_cap(w)));
  MML_CUDA(c, c->pin_flags.reserve(sizeof(int) * 64));
  // an empty window with this run's parameters
  WinDev* h = static_cast<WinDev*>(w->pin_up);
  memset(h, 0, offsetof(WinDev, upload_end));
  fill_win_params(h, exTlb16, gravity3, prm);
  MML_CUDA(c, cudaMemcpyAsync(w->dev.p, h, offsetof(WinDev, upload_end), cudaMemcpyHostToDevice, st));
  WinPush* ring = reinterpret_cast<WinPush*>(static_cast<char*>(w->pin_up) + sizeof(WinDev));
  double Rbl[9], Pbl[3];
  memcpy(Rbl, h->Rbl_raw, sizeof(Rbl));
  memcpy(Pbl, h->Pbl, sizeof(Pbl));
  double prev[16];
  memcpy(prev, state0, sizeof(prev));
  double t_prev = stamp0;
  size_t imu_off = 0;
  // map update of EstimateLidarPose (prm->map_update): counts of the frame in every physical slot, the transform the
  // reference carries from call to call (transformTobeMapped) and the position of the last update (EST.h:339)
  int slot_counts[kMaxWindow][3] = {};  // n_corner_ds, n_surf_ds, n_sharp
  double T_map[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  double last_update_pose[3] = {-1.0, -1.0, -1.0};
  long map_updates = 0;
  if (total_ms) { MML_CUDA(c, cudaEventRecord(c->ev0, st)); }
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
    int rc2 = MML_OK;
    if (rc == MML_OK) {
      const int cap = mml_split_voxel_capacity();
      if (w->x_idx[b].reserve(sizeof(int) * (2 * (size_t)cap + 4)) != cudaSuccess) rc2 = MML_ERR_CUDA;
      int* ix = w->x_idx[b].as<int>();
      if (rc2 == MML_OK) rc2 = mml_label_compact_device(c, w->x_label[b].as<uint8_t>(), (int)n, ix, ix + cap, ix + 2 * cap);
    }
    c->stream = keep;
    MML_CHECK(rc);
    MML_CHECK(rc2);
    MML_CUDA(c, cudaEventRecord(w->xev[b], xs));
    return MML_OK;
  };
  if (n_scans > 0) MML_CHECK(prefetch(0));
  auto pose16 = [](const double* st16, double* T) {
    double R[9];
    quat_to_R(Quat{st16[3], st16[4], st16[5], st16[6]}, R);
    const double Tn[16] = {R[0], R[1], R[2], st16[0], R[3], R[4], R[5], st16[1], R[6], R[7], R[8], st16[2], 0, 0, 0, 1};
    memcpy(T, Tn, sizeof(Tn));
  };
  for (int k = 0; k < n_scans; k++) {
    const double tq0 = g_prof_on ? now_us() : 0;
    WinPush* hp = ring + (k & 7);
    // the mean of the pre-integration first: it gives the predicted pose the scan's undistortion waits for; covariance,
    // Jacobian and sqrt-information (most of the host's time) follow below, while the device works on the scan
    MML_CHECK(mml_imu_preintegrate_mean(imu_t + imu_off, imu_gyr + 3 * imu_off, imu_acc + 3 * imu_off, imu_n[k], t_prev, prev + 10, prev + 13, &hp->pre));
    double next[16];
    MML_CHECK(mml_imu_predict(prev, &hp->pre, next));
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
    const double tq1 = g_prof_on ? now_us() : 0;
    MML_CUDA(c, cudaStreamWaitEvent(st, w->xev[b], 0));
    int* flags = c->pin_flags.as<int>() + 16 * b;
    MML_CHECK(window_push_scan(c, xd, ld, sd, n_pts[k], n_lines, dR9, dt3, leaf_corner, leaf_surf, window, nullptr,
                               w->x_label[b].as<uint8_t>(), w->x_counters[b].as<int>(), flags, w->x_idx[b].as<int>()));
    MML_CUDA(c, cudaEventRecord(w->xfree[b], st));  // raw scan, labels and counters of this buffer pair are consumed
    {  // the full pre-integration (same mean, to the bit: shared code) for the window's IMU factor
      const double dq_mean[4] = {hp->pre.dq[0], hp->pre.dq[1], hp->pre.dq[2], hp->pre.dq[3]};
      MML_CHECK(mml_imu_preintegrate(imu_t + imu_off, imu_gyr + 3 * imu_off, imu_acc + 3 * imu_off, imu_n[k], t_prev, prev + 10, prev + 13, &hp->pre));
      if (memcmp(dq_mean, hp->pre.dq, sizeof(dq_mean)) != 0) return mml_fail(c, MML_ERR_STATE, "pre-integration: the mean differs between its two passes");
      imu_off += imu_n[k];
    }
    const int W = w->n_slots;
    memcpy(hp->state, next, sizeof(next));
    hp->slot = w->order[W - 1];
    hp->window = window;
    const unsigned seq = ++w->seq;
    hp->seq = seq;
    MML_CUDA(c, cudaMemcpyAsync(w->push_dev.p, hp, sizeof(WinPush), cudaMemcpyHostToDevice, st));
    MML_CHECK(mml_window_push_launch(c, w, w->push_dev.as<WinPush>()));
    MML_CHECK(mml_window_solve_launch(c, w, window_cap(w), prm->max_outer));
    if (k + 1 < n_scans) MML_CHECK(prefetch(k + 1));
    const double tq2 = g_prof_on ? now_us() : 0;
    MML_CHECK(wait_window_result(c, w, seq));
    const double tq3 = g_prof_on ? now_us() : 0;
    if (flags[2] || flags[8]) return mml_fail(c, MML_ERR_CAPACITY, "scan exceeds the fused extraction / split-voxel capacities (window path)");
    const double* out = static_cast<const double*>(w->mapped);
    const double* so = out + 16 * kMaxWindow;
    c->launches += w->graph_launches * (long long)so[0] + 1;
    memcpy(prev, out + 16 * (W - 1), sizeof(prev));
    t_prev = stamps[k];
    if (poses_front) pose16(out, poses_front + 16 * (size_t)k);
    if (poses_newest) pose16(prev, poses_newest + 16 * (size_t)k);
    if (states_out) memcpy(states_out + 16 * (size_t)k, prev, sizeof(prev));
    if (stats_out) memcpy(stats_out + 8 * (size_t)k, so, sizeof(double) * 8);
    slot_counts[hp->slot][0] = flags[4]; slot_counts[hp->slot][1] = flags[5]; slot_counts[hp->slot][2] = flags[0];
    if (prm->map_update) {
      // EST.cpp:973-977: transformTobeMapped starts from the newest frame; 1041-1062: the oldest frame's pose replaces
      // it when enough corner points were seen (lidarMode 2: corner_cnt > 50), else only x and y are taken over
      auto twl = [&](const double* st16, double* T) {
        double Rq[9];
        quat_to_R(Quat{st16[3], st16[4], st16[5], st16[6]}, Rq);
        for (int r = 0; r < 3; r++) {
          for (int q = 0; q < 3; q++) T[4 * r + q] = Rq[3 * r] * Rbl[q] + Rq[3 * r + 1] * Rbl[3 + q] + Rq[3 * r + 2] * Rbl[6 + q];
          T[4 * r + 3] = Rq[3 * r] * Pbl[0] + Rq[3 * r + 1] * Pbl[1] + Rq[3 * r + 2] * Pbl[2] + st16[r];
        }
        T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
      };
      int corner_cnt = 0;
      for (int f = 0; f < W; f++) corner_cnt += slot_counts[w->order[f]][2];
      const bool degenerate = so[6] != 0.0;
      if (corner_cnt > 50) twl(out, T_map);
      else {
        double Tb[16];
        twl(next, Tb);  // the pose the call started from (the predicted newest frame)
        memcpy(T_map, Tb, sizeof(Tb));
        T_map[3] = out[0]; T_map[7] = out[1];
      }
      if (!degenerate) {  // EST.cpp:1066-1135
        const double dx = last_update_pose[0] - T_map[3], dy = last_update_pose[1] - T_map[7], dz = last_update_pose[2] - T_map[11];
        const float dis = (float)(dx * dx + dy * dy + dz * dz);
        if (dis >= 0.5f) {
          WinSlot& f0 = w->frame(0);
          const double tm0 = g_prof_on ? now_us() : 0;
          MML_CHECK(mml_local_map_push_dev(c, f0.q_corner.p, slot_counts[w->order[0]][0], f0.q_surf.p, slot_counts[w->order[0]][1], T_map,
                                           leaf_corner, leaf_surf, 1, nullptr, nullptr));
          if (g_prof_on) g_prof.other += now_us() - tm0;
          last_update_pose[0] = T_map[3]; last_update_pose[1] = T_map[7]; last_update_pose[2] = T_map[11];
          map_updates++;
        }
      }
    }
    if (g_prof_on) { g_prof.imu += tq1 - tq0; g_prof.launch += tq2 - tq1; g_prof.wait += tq3 - tq2; g_prof.evals += (long)so[7]; g_prof.scans++; }
  }
  if (total_ms) {
    MML_CUDA(c, cudaEventRecord(c->ev1, st));
    MML_CUDA(c, cudaEventSynchronize(c->ev1));
    MML_CUDA(c, cudaEventElapsedTime(total_ms, c->ev0, c->ev1));
  } else {
    MML_CUDA(c, cudaStreamSynchronize(st));
  }
#ifdef MML_WIN_TIMELINE
  if (g_prof_on) {
    unsigned long long t8[8];
    if (mml_debug_win_timeline(t8) == 0 && t8[5]) {
      const double ns = (double)t8[5];
      fprintf(stderr, "[mml window timeline] per scan (device clock): solve kernels %.1f us (%.2f launches), k_win_push %.1f us, push end -> first solve %.1f us, "
                      "solve -> next solve %.1f us, last solve end -> next push start %.1f us\n",
              t8[0] / ns * 1e-3, t8[6] / ns, t8[1] / ns * 1e-3, t8[2] / ns * 1e-3, t8[3] / ns * 1e-3, t8[4] / ns * 1e-3);
    }
  }
#endif
  if (g_prof_on && g_prof.scans) {
    const double ns = (double)g_prof.scans;
    fprintf(stderr, "[mml window prof] per scan: host pre-integration + prediction %.1f us, enqueue %.1f us, wait for the solve %.1f us, evaluations %.1f; map updates %ld\n",
            g_prof.imu / ns, g_prof.launch / ns, g_prof.wait / ns, g_prof.evals / ns, map_updates);
    if (map_updates) fprintf(stderr, "[mml window prof] map update: %.1f us each on the host (enqueue + the one wait)\n", g_prof.other / (double)map_updates);
    g_prof = WinProf();
  }
  return MML_OK;
}

}  // extern "C"
