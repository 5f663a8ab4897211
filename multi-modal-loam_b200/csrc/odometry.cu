// Native odometry loop: the per-scan body of process() (src/unionPoseEstimation.cpp:650-906) for a
// sequence of scans, with the reference's node pipeline kept: feature extraction does not depend on
// the pose (it is a separate ROS node in the reference, mm_scanRegistration -> mm_PoseEstimation),
// so scan k+1 is labelled on its own stream while scan k is being matched.
//
//   stream_fe : [H2D scan k+1] -> extraction k+1 (labels + counters into slot (k+1)%2)
//   stream    : wait(slot k) -> fused split + undistort + voxel -> Estimate graphs -> pose k
//
// Pose prediction is the constant-velocity model the reference uses before IMU initialisation
// (delta of the last two poses, PE.cpp:847-852, 882-890); the same delta drives undistortion.
#include "common.cuh"
#include "smallmath.cuh"
#include <math.h>

int mml_extract_device(mml_ctx* ctx, const float4* xyzi_d, const uint16_t* line_d, const int* scan_off, int n_scans,
                       int n_lines, uint8_t* label_d, bool force_sequential);
int mml_split_voxel_capacity();
int mml_split_voxel_device(mml_ctx* ctx, const float4* pts_d, const float* s_d, const uint8_t* label_d, int n,
                           const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, float4* corner_out,
                           float4* surf_out, int* counts_d);

namespace {

struct Slot {
  mml::DevBuf label, counters, in_xyzi, in_line, in_s;
  cudaEvent_t done = nullptr;
};

struct Odom {
  Slot slot[2];
};

void mat4_mul(const double* A, const double* B, double* C) {
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += A[4 * r + k] * B[4 * k + c];
      C[4 * r + c] = s;
    }
}
void rigid_inv(const double* T, double* Ti) {
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) Ti[4 * r + c] = T[4 * c + r];
    Ti[4 * r + 3] = -(T[0 * 4 + r] * T[3] + T[1 * 4 + r] * T[7] + T[2 * 4 + r] * T[11]);
  }
  Ti[12] = Ti[13] = Ti[14] = 0;
  Ti[15] = 1;
}

Odom* get_odom(mml_ctx* c) {
  if (!c->odom) {
    Odom* o = new Odom();
    for (int k = 0; k < 2; k++) cudaEventCreateWithFlags(&o->slot[k].done, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&c->stream_fe, cudaStreamNonBlocking);
    c->odom = o;
  }
  return static_cast<Odom*>(c->odom);
}

// enqueue (optional H2D +) extraction of one scan on the FE stream
int submit(mml_ctx* c, Odom* o, int k, const void* xyzi, const void* line, const void* s, int n, int n_lines, bool host,
           const void** xyzi_dev, const void** s_dev) {
  Slot& S = o->slot[k & 1];
  MML_CUDA(c, S.label.reserve((size_t)n + 16));
  MML_CUDA(c, S.counters.reserve(64));
  const void* xd = xyzi;
  const void* ld = line;
  const void* sd = s;
  if (host) {
    MML_CUDA(c, S.in_xyzi.reserve(sizeof(float4) * (size_t)n));
    MML_CUDA(c, S.in_line.reserve(sizeof(uint16_t) * (size_t)n));
    MML_CUDA(c, S.in_s.reserve(sizeof(float) * (size_t)n));
    MML_CUDA(c, cudaMemcpyAsync(S.in_xyzi.p, xyzi, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, c->stream_fe));
    MML_CUDA(c, cudaMemcpyAsync(S.in_line.p, line, sizeof(uint16_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream_fe));
    if (s) MML_CUDA(c, cudaMemcpyAsync(S.in_s.p, s, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, c->stream_fe));
    xd = S.in_xyzi.p;
    ld = S.in_line.p;
    sd = s ? S.in_s.p : nullptr;
  }
  *xyzi_dev = xd;
  *s_dev = sd;
  const int off[2] = {0, n};
  cudaStream_t main_stream = c->stream;
  c->stream = c->stream_fe;
  c->counters_alt = S.counters.as<int>();
  const int rc = mml_extract_device(c, (const float4*)xd, (const uint16_t*)ld, off, 1, n_lines, S.label.as<uint8_t>(), false);
  c->counters_alt = nullptr;
  c->stream = main_stream;
  MML_CHECK(rc);
  MML_CUDA(c, cudaEventRecord(S.done, c->stream_fe));
  return MML_OK;
}

}  // namespace

extern "C" {

int mml_scan_to_pose_dev(mml_ctx* c, const void* xyzi_dev, const void* line_id_dev, const void* s_dev, int n, int n_lines,
                         const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, const double* exTlb16,
                         double* P3, double* q_wxyz4, const mml_est_params* prm, double* stats, int* out_counts);
int mml_scan_to_pose(mml_ctx* c, const float* xyzi, const uint16_t* line_id, const float* s, int n, int n_lines,
                     const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, const double* exTlb16,
                     double* P3, double* q_wxyz4, const mml_est_params* prm, double* stats, int* out_counts);

// Run the odometry loop over n_scans scans. xyzi/line/s are arrays of per-scan pointers (device pointers when
// host_buffers == 0, host — ideally pinned — pointers otherwise). T_init16 = pose of the frame before the first
// scan, T_prev16 the one before that (constant-velocity seed). poses_out: n_scans x 16 (row-major T_wb).
// total_ms (optional): CUDA-event time of the whole run on the context's stream.
int mml_odom_run(mml_ctx* c, const void* const* xyzi, const void* const* line, const void* const* s, const int* n_pts,
                 int n_scans, int n_lines, int host_buffers, const double* T_init16, const double* T_prev16,
                 const double* exTlb16, float leaf_corner, float leaf_surf, const mml_est_params* prm, double* poses_out,
                 float* total_ms, int* counts_out /* n_scans x 4, optional */) {
  if (!c || n_scans < 0 || !T_init16 || !T_prev16 || !exTlb16 || !poses_out) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  mml_est_params def;
  mml_est_params_default(&def);
  if (!prm) prm = &def;
  Odom* o = get_odom(c);
  cudaStream_t st = c->stream;
  const int cap = mml_split_voxel_capacity();
  MML_CUDA(c, c->frame_cnt.reserve(64));
  MML_CUDA(c, c->q_corner.reserve(sizeof(float4) * (size_t)cap));
  MML_CUDA(c, c->q_surf.reserve(sizeof(float4) * (size_t)cap));
  MML_CUDA(c, c->pin_flags.reserve(64));
  c->has_perm[0] = c->has_perm[1] = false;
  double T_last[16], T_before[16];
  memcpy(T_last, T_init16, sizeof(T_last));
  memcpy(T_before, T_prev16, sizeof(T_before));
  MML_CUDA(c, cudaStreamSynchronize(st));
  MML_CUDA(c, cudaStreamSynchronize(c->stream_fe));
  if (total_ms) MML_CUDA(c, cudaEventRecord(c->ev0, st));
  const void* xd[2] = {nullptr, nullptr};
  const void* sd[2] = {nullptr, nullptr};
  if (n_scans > 0)
    MML_CHECK(submit(c, o, 0, xyzi[0], line[0], s ? s[0] : nullptr, n_pts[0], n_lines, host_buffers != 0, &xd[0], &sd[0]));
  for (int k = 0; k < n_scans; k++) {
    // the reference's pipeline: the extractor node works on the next scan meanwhile. Its launches are issued from
    // the hook below, after this scan's critical-path work is already in the stream.
    struct Next {
      mml_ctx* c; Odom* o; int k; const void* xyzi; const void* line; const void* s; int n, n_lines; bool host;
      const void** xd; const void** sd;
    } nx = {c, o, k + 1, nullptr, nullptr, nullptr, 0, n_lines, host_buffers != 0, &xd[(k + 1) & 1], &sd[(k + 1) & 1]};
    if (k + 1 < n_scans) { nx.xyzi = xyzi[k + 1]; nx.line = line[k + 1]; nx.s = s ? s[k + 1] : nullptr; nx.n = n_pts[k + 1]; }
    auto submit_next = [](void* a) -> int {
      Next* x = static_cast<Next*>(a);
      return submit(x->c, x->o, x->k, x->xyzi, x->line, x->s, x->n, x->n_lines, x->host, x->xd, x->sd);
    };
    Slot& S = o->slot[k & 1];
    // constant-velocity prediction and the motion used for undistortion
    double Tinv[16], delta[16], Tp[16];
    rigid_inv(T_before, Tinv);
    mat4_mul(Tinv, T_last, delta);
    mat4_mul(T_last, delta, Tp);
    const double dR[9] = {delta[0], delta[1], delta[2], delta[4], delta[5], delta[6], delta[8], delta[9], delta[10]};
    const double dt[3] = {delta[3], delta[7], delta[11]};
    const double Rp[9] = {Tp[0], Tp[1], Tp[2], Tp[4], Tp[5], Tp[6], Tp[8], Tp[9], Tp[10]};
    const mml::Quat qp = mml::quat_from_R9(Rp);
    double P[3] = {Tp[3], Tp[7], Tp[11]}, q[4] = {qp.w, qp.x, qp.y, qp.z};
    double stats[16];
    const int n = n_pts[k];
    int* cnt = c->frame_cnt.as<int>();
    MML_CUDA(c, cudaStreamWaitEvent(st, S.done, 0));
    MML_CHECK(mml_split_voxel_device(c, (const float4*)xd[k & 1], (const float*)sd[k & 1], S.label.as<uint8_t>(), n, dR, dt,
                                     leaf_corner, leaf_surf, c->q_corner.as<float4>(), c->q_surf.as<float4>(), cnt));
    int* hf = c->pin_flags.as<int>();
    MML_CUDA(c, cudaMemcpyAsync(hf, S.counters.p, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
    MML_CUDA(c, cudaMemcpyAsync(hf + 4, cnt, 5 * sizeof(int), cudaMemcpyDeviceToHost, st));
    MML_CHECK(mml_estimate_device(c, cnt, cap, cap, exTlb16, P, q, prm, stats, k + 1 < n_scans ? +submit_next : nullptr,
                                  &nx));  // synchronises `st`
    if (hf[2] || hf[8]) {
      // capacity overflow of a fused kernel: this scan goes through the general (unpipelined) path
      MML_CUDA(c, cudaStreamSynchronize(c->stream_fe));
      P[0] = Tp[3]; P[1] = Tp[7]; P[2] = Tp[11];
      q[0] = qp.w; q[1] = qp.x; q[2] = qp.y; q[3] = qp.z;
      int oc[4];
      if (host_buffers)
        MML_CHECK(mml_scan_to_pose(c, (const float*)xyzi[k], (const uint16_t*)line[k], s ? (const float*)s[k] : nullptr, n,
                                   n_lines, dR, dt, leaf_corner, leaf_surf, exTlb16, P, q, prm, stats, oc));
      else
        MML_CHECK(mml_scan_to_pose_dev(c, xyzi[k], line[k], s ? s[k] : nullptr, n, n_lines, dR, dt, leaf_corner, leaf_surf,
                                       exTlb16, P, q, prm, stats, oc));
      if (counts_out) memcpy(counts_out + 4 * k, oc, sizeof(oc));
    } else if (counts_out) {
      counts_out[4 * k] = hf[0]; counts_out[4 * k + 1] = hf[1]; counts_out[4 * k + 2] = hf[4]; counts_out[4 * k + 3] = hf[5];
    }
    double R[9];
    mml::quat_to_R(mml::Quat{q[0], q[1], q[2], q[3]}, R);
    double Tn[16] = {R[0], R[1], R[2], P[0], R[3], R[4], R[5], P[1], R[6], R[7], R[8], P[2], 0, 0, 0, 1};
    memcpy(poses_out + 16 * (size_t)k, Tn, sizeof(Tn));
    memcpy(T_before, T_last, sizeof(T_last));
    memcpy(T_last, Tn, sizeof(Tn));
  }
  if (total_ms) {
    MML_CUDA(c, cudaEventRecord(c->ev1, st));
    MML_CUDA(c, cudaEventSynchronize(c->ev1));
    MML_CUDA(c, cudaEventElapsedTime(total_ms, c->ev0, c->ev1));
  }
  return MML_OK;
}

}  // extern "C"
