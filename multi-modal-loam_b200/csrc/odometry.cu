// Native odometry loop: the per-scan body of process() (src/unionPoseEstimation.cpp:650-906) for a
// sequence of scans, with the reference's node pipeline kept: feature extraction does not depend on
// the pose (it is a separate ROS node in the reference, mm_scanRegistration -> mm_PoseEstimation),
// so scan k+1 is labelled on its own stream while scan k is being matched.
//
//   stream_fe : [H2D scan k+1] -> extraction k+1 (labels + counters into slot (k+1)%2)
//   stream    : wait(slot k) -> fused split + undistort + voxel -> Estimate -> pose k
//
// Pose prediction is the constant-velocity model the reference uses before IMU initialisation
// (delta of the last two poses, PE.cpp:847-852, 882-890); the same delta drives undistortion.
//
// Two drivers share the kernels:
//   chained (default): the last two poses live on the device (OdomDev); the split/voxel launch derives the
//     prediction from them and starts the solve, and the solve is one graph whose WHILE node repeats the outer
//     iteration until the device-side convergence test passes. The host only enqueues - it never waits on a
//     scan - and reads all poses back once at the end.
//   classic (MML_ODOM_CLASSIC=1, and for scans whose fused kernels overflow): host-driven, one short
//     synchronisation per outer iteration.
#include "common.cuh"
#include "smallmath.cuh"
#include "eststate.cuh"
#include <math.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include <chrono>
#include <utility>

int mml_extract_device(mml_ctx* ctx, const float4* xyzi_d, const uint16_t* line_d, const int* scan_off, int n_scans,
                       int n_lines, uint8_t* label_d, bool force_sequential);
int mml_split_voxel_capacity();
int mml_split_voxel_device(mml_ctx* ctx, const float4* pts_d, const float* s_d, const uint8_t* label_d, int n,
                           const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, float4* corner_out,
                           float4* surf_out, int* counts_d, const mml::SvChain* chain = nullptr);
int mml_label_compact_device(mml_ctx* ctx, const uint8_t* label_d, int n, int* idx0, int* idx1, int* cnt_d);
int mml_chain_prepare(mml_ctx* ctx, int cap, mml::EstState** S_out);
mml::EstInit mml_make_est_init(const double* exTlb16, const mml_est_params* prm);
int mml_chain_solve_launch(mml_ctx* ctx, const int* cnt_dev, int cap, mml::OdomDev* od, mml::ChainOut out, cudaEvent_t wait_ev);

namespace {

using mml::mat4_mul;
using mml::rigid_inv;

constexpr int kSlots = 4;   // scans in flight: being copied, being labelled (x2), being matched
constexpr int kLanes = 2;   // extraction lanes (streams + scratch sets): scans k and k+1 are labelled concurrently

struct Slot {
  mml::DevBuf label, counters, in_xyzi, in_line, in_s;
  mml::DevBuf idx;                 // chained loop: [2][cap] compacted labelled indices + int[2] counts (k_label_compact)
  cudaEvent_t copied = nullptr;    // host -> device copy of the scan has finished (copy stream)
  cudaEvent_t done = nullptr;      // extraction of the scan in this slot has finished (extraction lane)
  cudaEvent_t consumed = nullptr;  // the matcher has read the slot (main stream): the next scan may overwrite it
  bool used = false;
  const void* xd = nullptr;        // device pointers of the scan currently in the slot
  const void* ld = nullptr;
  const void* sd = nullptr;
};

// the extraction's working buffers (mml_ctx members), one set per lane; swapped into the context around a launch
struct FeScratch {
  mml::DevBuf srt_xyzi, srt_src, srt_line, chunk_tab, chunk_hist, line_start, line_count, curv, refl, attr, sort_ind, refl_ind;
  mml::PinBuf pin_small;
  std::vector<int> last_scan_off;
};
void swap_scratch(mml_ctx* c, FeScratch& f) {
  std::swap(c->srt_xyzi, f.srt_xyzi); std::swap(c->srt_src, f.srt_src); std::swap(c->srt_line, f.srt_line);
  std::swap(c->chunk_tab, f.chunk_tab); std::swap(c->chunk_hist, f.chunk_hist); std::swap(c->line_start, f.line_start);
  std::swap(c->line_count, f.line_count); std::swap(c->curv, f.curv); std::swap(c->refl, f.refl); std::swap(c->attr, f.attr);
  std::swap(c->sort_ind, f.sort_ind); std::swap(c->refl_ind, f.refl_ind); std::swap(c->pin_small, f.pin_small);
  std::swap(c->last_scan_off, f.last_scan_off);
}

struct Odom {
  Slot slot[kSlots];
  cudaStream_t lane[kLanes] = {nullptr, nullptr};  // lane[0] == ctx->stream_fe
  cudaStream_t copy = nullptr;
  cudaStream_t sv = nullptr;        // chained loop: the split / voxel launch of scan k (beside the graph launch of scan k)
  cudaEvent_t sv_done = nullptr;    // ... finished (waited for inside the scan's graph)
  cudaEvent_t scan_done = nullptr;  // the solve of the previous scan has finished (matcher stream)
  bool big_scans = false;           // sticky: scans label more points than the fused split / voxel kernel holds ->
                                    // straight to the general path until a scan fits comfortably again
  FeScratch scratch[kLanes];  // scratch[0] stays empty: lane 0 works in the context's own buffers
  mml::DevBuf state;  // OdomDev
  mml::DevBuf out;    // ChainOut arrays
  mml::PinBuf host;   // staging: OdomDev upload + ChainOut read-back
};

Odom* get_odom(mml_ctx* c) {
  if (!c->odom) {
    Odom* o = new Odom();
    for (int k = 0; k < kSlots; k++) {
      cudaEventCreateWithFlags(&o->slot[k].copied, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&o->slot[k].done, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&o->slot[k].consumed, cudaEventDisableTiming);
    }
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    cudaStreamCreateWithPriority(&c->stream_fe, cudaStreamNonBlocking, prio_lo);
    o->lane[0] = c->stream_fe;
    for (int k = 1; k < kLanes; k++) cudaStreamCreateWithPriority(&o->lane[k], cudaStreamNonBlocking, prio_lo);
    cudaStreamCreateWithPriority(&o->copy, cudaStreamNonBlocking, prio_lo);
    cudaStreamCreateWithPriority(&o->sv, cudaStreamNonBlocking, prio_hi);
    cudaEventCreateWithFlags(&o->sv_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&o->scan_done, cudaEventDisableTiming);
    c->odom = o;
  }
  return static_cast<Odom*>(c->odom);
}

struct Trace;
void trace_rec(Trace* t, int k, int j, cudaStream_t st);

// Stage 1 of a scan: make it resident. Host buffers are copied into the scan's slot - on the copy stream when
// `own_stream` (chained loop: the copy of scan k+3 runs beside the labelling of k+1, k+2), else on lane 0.
int submit_copy(mml_ctx* c, Odom* o, int k, const void* xyzi, const void* line, const void* s, int n, bool host, bool own_stream) {
  Slot& S = o->slot[k % kSlots];
  cudaStream_t sc = own_stream ? o->copy : o->lane[0];
  if (S.used) MML_CUDA(c, cudaStreamWaitEvent(sc, S.consumed, 0));
  S.xd = xyzi;
  S.ld = line;
  S.sd = s;
  if (host) {
    MML_CUDA(c, S.in_xyzi.reserve(sizeof(float4) * (size_t)n));
    MML_CUDA(c, S.in_line.reserve(sizeof(uint16_t) * (size_t)n));
    MML_CUDA(c, S.in_s.reserve(sizeof(float) * (size_t)n));
    MML_CUDA(c, cudaMemcpyAsync(S.in_xyzi.p, xyzi, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, sc));
    MML_CUDA(c, cudaMemcpyAsync(S.in_line.p, line, sizeof(uint16_t) * (size_t)n, cudaMemcpyHostToDevice, sc));
    if (s) MML_CUDA(c, cudaMemcpyAsync(S.in_s.p, s, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, sc));
    S.xd = S.in_xyzi.p;
    S.ld = S.in_line.p;
    S.sd = s ? S.in_s.p : nullptr;
  }
  MML_CUDA(c, cudaEventRecord(S.copied, sc));
  return MML_OK;
}

// Stage 2: label the scan (and, for the chained loop, compact the labelled indices) on extraction lane `ln`.
int submit_extract(mml_ctx* c, Odom* o, int k, int n, int n_lines, int ln, bool compact, Trace* tr) {
  Slot& S = o->slot[k % kSlots];
  cudaStream_t sfe = o->lane[ln];
  MML_CUDA(c, S.label.reserve((size_t)n + 16));
  MML_CUDA(c, S.counters.reserve(64));
  MML_CUDA(c, cudaStreamWaitEvent(sfe, S.copied, 0));  // also orders the lane after the slot's previous consumer
  if (tr) trace_rec(tr, k, 3, sfe);
  const int off[2] = {0, n};
  cudaStream_t main_stream = c->stream;
  c->stream = sfe;
  c->counters_alt = S.counters.as<int>();
  if (ln > 0) swap_scratch(c, o->scratch[ln]);
  int rc = mml_extract_device(c, (const float4*)S.xd, (const uint16_t*)S.ld, off, 1, n_lines, S.label.as<uint8_t>(), false);
  if (rc == MML_OK && compact) {
    const int cap = mml_split_voxel_capacity();
    if (S.idx.reserve(sizeof(int) * (2 * (size_t)cap + 4)) != cudaSuccess) rc = MML_ERR_CUDA;
    int* ix = S.idx.as<int>();
    if (rc == MML_OK) rc = mml_label_compact_device(c, S.label.as<uint8_t>(), n, ix, ix + cap, ix + 2 * cap);
  }
  if (ln > 0) swap_scratch(c, o->scratch[ln]);
  c->counters_alt = nullptr;
  c->stream = main_stream;
  MML_CHECK(rc);
  if (tr) trace_rec(tr, k, 4, sfe);
  MML_CUDA(c, cudaEventRecord(S.done, sfe));
  S.used = true;
  return MML_OK;
}

// classic driver: copy + labelling of one scan on lane 0
int submit(mml_ctx* c, Odom* o, int k, const void* xyzi, const void* line, const void* s, int n, int n_lines, bool host,
           const void** xyzi_dev, const void** s_dev) {
  MML_CHECK(submit_copy(c, o, k, xyzi, line, s, n, host, false));
  MML_CHECK(submit_extract(c, o, k, n, n_lines, 0, false, nullptr));
  *xyzi_dev = o->slot[k % kSlots].xd;
  *s_dev = o->slot[k % kSlots].sd;
  return MML_OK;
}

// MML_ODOM_TRACE=1: CUDA-event timeline of the chained loop (per-scan stage durations on both streams), printed
// to stderr. Events perturb the pipeline a little; never enabled in bench numbers.
struct Trace {
  bool on = false;
  std::vector<cudaEvent_t> ev;  // per scan: main after-wait, after split/voxel, after solve; FE start, FE end
  cudaEvent_t at(int k, int j) { return ev[5 * (size_t)k + j]; }
  void init(int n) {
    on = getenv("MML_ODOM_TRACE") && atoi(getenv("MML_ODOM_TRACE")) != 0;
    if (!on) return;
    ev.resize(5 * (size_t)n);
    for (auto& e : ev) cudaEventCreate(&e);
  }
  void rec(int k, int j, cudaStream_t st) { if (on) cudaEventRecord(at(k, j), st); }
  void report(int n) {
    if (!on) return;
    double sv = 0, solve = 0, fe = 0, wait = 0, period = 0;
    float ms;
    for (int k = 1; k < n; k++) {
      cudaEventElapsedTime(&ms, at(k, 0), at(k, 1)); sv += ms;
      cudaEventElapsedTime(&ms, at(k, 1), at(k, 2)); solve += ms;
      cudaEventElapsedTime(&ms, at(k, 3), at(k, 4)); fe += ms;
      cudaEventElapsedTime(&ms, at(k - 1, 2), at(k, 0)); wait += ms;
      cudaEventElapsedTime(&ms, at(k - 1, 2), at(k, 2)); period += ms;
    }
    const double d = 1e3 / (n - 1);
    fprintf(stderr, "odom trace (us/scan over %d scans): period %.1f = wait-for-extraction %.1f + split/voxel %.1f + solve %.1f; "
            "extraction stream busy %.1f\n", n - 1, period * d, wait * d, sv * d, solve * d, fe * d);
    for (auto& e : ev) cudaEventDestroy(e);
    ev.clear();
  }
};

struct RunArgs {
  const void* const* xyzi; const void* const* line; const void* const* s; const int* n_pts;
  int n_scans, n_lines, host_buffers;
  const double* exTlb16; float leaf_corner, leaf_surf; const mml_est_params* prm;
  double* poses_out; int* counts_out;
};

void trace_rec(Trace* t, int k, int j, cudaStream_t st) { t->rec(k, j, st); }

}  // namespace

extern "C" {
int mml_scan_to_pose_dev(mml_ctx* c, const void* xyzi_dev, const void* line_id_dev, const void* s_dev, int n, int n_lines,
                         const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, const double* exTlb16,
                         double* P3, double* q_wxyz4, const mml_est_params* prm, double* stats, int* out_counts);
int mml_scan_to_pose(mml_ctx* c, const float* xyzi, const uint16_t* line_id, const float* s, int n, int n_lines,
                     const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, const double* exTlb16,
                     double* P3, double* q_wxyz4, const mml_est_params* prm, double* stats, int* out_counts);
}

// ---------------------------------------------------------------- classic driver: scans [first, n_scans)
static int run_classic(mml_ctx* c, Odom* o, const RunArgs& R, int first, const double* T_init16, const double* T_prev16) {
  cudaStream_t st = c->stream;
  const int cap = mml_split_voxel_capacity();
  const int n_scans = R.n_scans, n_lines = R.n_lines;
  const bool host = R.host_buffers != 0;
  double T_last[16], T_before[16];
  memcpy(T_last, T_init16, sizeof(T_last));
  memcpy(T_before, T_prev16, sizeof(T_before));
  const void* xd[2] = {nullptr, nullptr};
  const void* sd[2] = {nullptr, nullptr};
  int next_submit = first;  // first scan whose extraction has not been enqueued yet
  for (int k = first; k < n_scans; k++) {
    // the reference's pipeline: the extractor node works on the next scan meanwhile. Its launches are issued from
    // the hook below, after this scan's critical-path work is already in the stream.
    struct Next {
      mml_ctx* c; Odom* o; int k; const void* xyzi; const void* line; const void* s; int n, n_lines; bool host;
      const void** xd; const void** sd;
    } nx = {c, o, k + 1, nullptr, nullptr, nullptr, 0, n_lines, host, &xd[(k + 1) & 1], &sd[(k + 1) & 1]};
    if (k + 1 < n_scans) { nx.xyzi = R.xyzi[k + 1]; nx.line = R.line[k + 1]; nx.s = R.s ? R.s[k + 1] : nullptr; nx.n = R.n_pts[k + 1]; }
    auto submit_next = [](void* a) -> int {
      Next* x = static_cast<Next*>(a);
      return submit(x->c, x->o, x->k, x->xyzi, x->line, x->s, x->n, x->n_lines, x->host, x->xd, x->sd);
    };
    const bool fused = !o->big_scans;
    if (fused && next_submit <= k) {
      MML_CHECK(submit(c, o, k, R.xyzi[k], R.line[k], R.s ? R.s[k] : nullptr, R.n_pts[k], n_lines, host, &xd[k & 1], &sd[k & 1]));
      next_submit = k + 1;
    }
    // general mode (scans beyond the fused kernels' capacity): scans k and k+1 are copied / labelled on the copy stream
    // and extraction lane 1 (its own scratch set: the general path below works in the context's), so that the
    // labelling of scan k+1 overlaps the matching of scan k
    auto prefetch = [&](int j) -> int {
      MML_CHECK(submit_copy(c, o, j, R.xyzi[j], R.line[j], R.s ? R.s[j] : nullptr, R.n_pts[j], host, true));
      return submit_extract(c, o, j, R.n_pts[j], n_lines, 1, false, nullptr);
    };
    bool prelabelled = false;
    if (!fused) {
      while (next_submit <= k + 1 && next_submit < n_scans) { MML_CHECK(prefetch(next_submit)); next_submit++; }
      prelabelled = true;
    }
    Slot& S = o->slot[k % kSlots];
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
    const int n = R.n_pts[k];
    int* cnt = c->frame_cnt.as<int>();
    int* hf = c->pin_flags.as<int>();
    if (fused) {
      MML_CUDA(c, cudaStreamWaitEvent(st, S.done, 0));
      MML_CHECK(mml_split_voxel_device(c, (const float4*)xd[k & 1], (const float*)sd[k & 1], S.label.as<uint8_t>(), n, dR, dt,
                                       R.leaf_corner, R.leaf_surf, c->q_corner.as<float4>(), c->q_surf.as<float4>(), cnt));
      MML_CUDA(c, cudaMemcpyAsync(hf, S.counters.p, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
      MML_CUDA(c, cudaMemcpyAsync(hf + 4, cnt, 5 * sizeof(int), cudaMemcpyDeviceToHost, st));
      MML_CUDA(c, cudaEventRecord(S.consumed, st));
      const bool pipeline_next = k + 1 < n_scans && next_submit == k + 1;
      MML_CHECK(mml_estimate_device(c, cnt, cap, cap, R.exTlb16, P, q, R.prm, stats, pipeline_next ? +submit_next : nullptr,
                                    &nx));  // synchronises `st`
      if (pipeline_next) next_submit = k + 2;
      if (hf[8]) o->big_scans = true;  // the labelled points do not fit the fused kernel: stop trying for now
    }
    if (!fused || hf[2] || hf[8]) {
      // capacity overflow of a fused kernel: this scan goes through the general (unpipelined) path
      MML_CUDA(c, cudaStreamSynchronize(c->stream_fe));
      P[0] = Tp[3]; P[1] = Tp[7]; P[2] = Tp[11];
      q[0] = qp.w; q[1] = qp.x; q[2] = qp.y; q[3] = qp.z;
      int oc[4];
      c->prefer_general = !fused;  // a scan already known not to fit goes straight to the general path
      struct Reset { mml_ctx* c; ~Reset() { c->prefer_general = false; c->pre_label = nullptr; c->pre_counters = nullptr; } } reset{c};
      if (prelabelled) {
        // labels, counters and the device copy of the scan are in its slot (lane 1); the matcher stream waits for them
        MML_CUDA(c, cudaStreamWaitEvent(st, S.done, 0));
        c->pre_label = S.label.as<uint8_t>();
        c->pre_counters = S.counters.as<int>();
        MML_CHECK(mml_scan_to_pose_dev(c, S.xd, S.ld, S.sd, n, n_lines, dR, dt, R.leaf_corner, R.leaf_surf, R.exTlb16, P, q, R.prm,
                                       stats, oc));
        MML_CUDA(c, cudaEventRecord(S.consumed, st));
      } else if (host) {
        MML_CHECK(mml_scan_to_pose(c, (const float*)R.xyzi[k], (const uint16_t*)R.line[k], R.s ? (const float*)R.s[k] : nullptr,
                                   n, n_lines, dR, dt, R.leaf_corner, R.leaf_surf, R.exTlb16, P, q, R.prm, stats, oc));
      } else {
        MML_CHECK(mml_scan_to_pose_dev(c, R.xyzi[k], R.line[k], R.s ? R.s[k] : nullptr, n, n_lines, dR, dt, R.leaf_corner,
                                       R.leaf_surf, R.exTlb16, P, q, R.prm, stats, oc));
      }
      if (R.counts_out) memcpy(R.counts_out + 4 * k, oc, sizeof(oc));
      if (o->big_scans && oc[0] < cap / 2 && oc[1] < cap / 2) o->big_scans = false;
    } else if (R.counts_out) {
      R.counts_out[4 * k] = hf[0]; R.counts_out[4 * k + 1] = hf[1]; R.counts_out[4 * k + 2] = hf[4]; R.counts_out[4 * k + 3] = hf[5];
    }
    double Rm[9];
    mml::quat_to_R(mml::Quat{q[0], q[1], q[2], q[3]}, Rm);
    double Tn[16] = {Rm[0], Rm[1], Rm[2], P[0], Rm[3], Rm[4], Rm[5], P[1], Rm[6], Rm[7], Rm[8], P[2], 0, 0, 0, 1};
    memcpy(R.poses_out + 16 * (size_t)k, Tn, sizeof(Tn));
    memcpy(T_before, T_last, sizeof(T_last));
    memcpy(T_last, Tn, sizeof(Tn));
  }
  return MML_OK;
}

// ---------------------------------------------------------------- chained driver
// Returns the index of the first scan whose results are not valid (a fused kernel overflowed its capacity), or
// n_scans when every scan went through.
static int run_chained(mml_ctx* c, Odom* o, const RunArgs& R, const double* T_init16, const double* T_prev16, int* first_bad) {
  cudaStream_t st = c->stream;
  const int cap = mml_split_voxel_capacity();
  const int n_scans = R.n_scans, n_lines = R.n_lines;
  const bool host = R.host_buffers != 0;
  *first_bad = n_scans;
  if (n_scans == 0) return MML_OK;
  mml::EstState* S = nullptr;
  MML_CHECK(mml_chain_prepare(c, cap, &S));
#ifdef MML_TIMELINE
  MML_CUDA(c, c->timeline.reserve(8 * (16 + 8 * 4000) + 4 * 32768 * 4));
  MML_CUDA(c, cudaMemsetAsync(c->timeline.p, 0, 8 * 16, st));
#endif
  const size_t out_doubles = (size_t)n_scans * 24, out_ints = (size_t)n_scans * 8;
  const size_t out_bytes = out_doubles * sizeof(double) + out_ints * sizeof(int);
  MML_CUDA(c, o->state.reserve(sizeof(mml::OdomDev) + 64));
  MML_CUDA(c, o->out.reserve(out_bytes + 64));
  MML_CUDA(c, o->host.reserve(out_bytes + sizeof(mml::OdomDev) + 64));
  mml::OdomDev* od = o->state.as<mml::OdomDev>();
  mml::ChainOut out;
  out.poses = o->out.as<double>();
  out.stats = out.poses + (size_t)n_scans * 16;
  out.counts = reinterpret_cast<int*>(out.poses + out_doubles);
  // pose history on the device
  char* hp = o->host.as<char>();
  mml::OdomDev* h_od = reinterpret_cast<mml::OdomDev*>(hp + out_bytes);
  memset(h_od, 0, sizeof(*h_od));
  memcpy(h_od->T_last, T_init16, sizeof(h_od->T_last));
  memcpy(h_od->T_before, T_prev16, sizeof(h_od->T_before));
  MML_CUDA(c, cudaMemcpyAsync(od, h_od, sizeof(*h_od), cudaMemcpyHostToDevice, st));
  MML_CUDA(c, cudaMemsetAsync(out.counts, 0, out_ints * sizeof(int), st));
  mml::SvChain ch;
  ch.od = od;
  ch.est = S;
  ch.I = mml_make_est_init(R.exTlb16, R.prm);
  Trace tr;
  tr.init(n_scans);
  MML_CUDA(c, cudaEventRecord(o->scan_done, st));  // the pose history upload above precedes the first split / voxel launch
  // pipeline depth: copies run three scans ahead of the matcher, labelling two (one scan per extraction lane)
  auto copy_of = [&](int k) { return submit_copy(c, o, k, R.xyzi[k], R.line[k], R.s ? R.s[k] : nullptr, R.n_pts[k], host, true); };
  auto extract_of = [&](int k) { return submit_extract(c, o, k, R.n_pts[k], n_lines, k % kLanes, true, &tr); };
  for (int k = 0; k < 3 && k < n_scans; k++) MML_CHECK(copy_of(k));
  for (int k = 0; k < 2 && k < n_scans; k++) MML_CHECK(extract_of(k));
  int* cnt = c->frame_cnt.as<int>();
  const auto host_t0 = std::chrono::steady_clock::now();
  for (int k = 0; k < n_scans; k++) {
    Slot& SL = o->slot[k % kSlots];
    // split / voxel of scan k on its own stream: after the labelling of scan k and the solve of scan k-1 ...
    MML_CUDA(c, cudaStreamWaitEvent(o->sv, SL.done, 0));
    MML_CUDA(c, cudaStreamWaitEvent(o->sv, o->scan_done, 0));
    tr.rec(k, 0, o->sv);
    ch.fe_counters = SL.counters.as<int>();
    ch.counts_out = out.counts + 8 * (size_t)k;
    ch.pre_idx[0] = SL.idx.as<int>();
    ch.pre_idx[1] = SL.idx.as<int>() + cap;
    ch.pre_cnt = SL.idx.as<int>() + 2 * cap;
    c->stream = o->sv;
    const int rc_sv = mml_split_voxel_device(c, (const float4*)SL.xd, (const float*)SL.sd, SL.label.as<uint8_t>(), R.n_pts[k], nullptr,
                                             nullptr, R.leaf_corner, R.leaf_surf, c->q_corner.as<float4>(), c->q_surf.as<float4>(),
                                             cnt, &ch);
    c->stream = st;
    MML_CHECK(rc_sv);
    MML_CUDA(c, cudaEventRecord(SL.consumed, o->sv));
    MML_CUDA(c, cudaEventRecord(o->sv_done, o->sv));
    tr.rec(k, 1, o->sv);
    // ... and the scan's graph on the matcher stream: it waits for sv_done inside (event-wait node), so its launch
    // latency overlaps the split / voxel kernel
    MML_CHECK(mml_chain_solve_launch(c, cnt, cap, od, out, o->sv_done));
    MML_CUDA(c, cudaEventRecord(o->scan_done, st));
    tr.rec(k, 2, st);
    if (k + 3 < n_scans) MML_CHECK(copy_of(k + 3));
    if (k + 2 < n_scans) MML_CHECK(extract_of(k + 2));
  }
  const double host_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - host_t0).count();
  if (tr.on) fprintf(stderr, "odom trace: host enqueue time %.1f us per scan (%d scans)\n", host_us / n_scans, n_scans);
  MML_CUDA(c, cudaMemcpyAsync(hp, out.poses, out_bytes, cudaMemcpyDeviceToHost, st));
  MML_CUDA(c, cudaStreamSynchronize(st));
  for (int k = 0; k < kLanes; k++) MML_CUDA(c, cudaStreamSynchronize(o->lane[k]));
  MML_CUDA(c, cudaStreamSynchronize(o->copy));
  MML_CUDA(c, cudaStreamSynchronize(o->sv));
  tr.report(n_scans);
#ifdef MML_TIMELINE
  {  // device-side stamps of every kernel on the matcher's path (debug build only)
    std::vector<unsigned long long> tl(16 + 8 * 4000);
    cudaMemcpy(tl.data(), c->timeline.p, tl.size() * 8, cudaMemcpyDeviceToHost);
    const int n = (int)tl[9] < 4000 ? (int)tl[9] : 4000;
    double gap = 0, sv = 0, entry = 0, a1 = 0, a0 = 0, join = 0, solve = 0;
    int m = 0;
    for (int i = 1; i < n; i++) {
      const unsigned long long* t = &tl[16 + 8 * i];
      const unsigned long long* p = &tl[16 + 8 * (i - 1)];
      if (t[0] == p[0]) continue;  // second outer iteration of the same scan
      gap += (double)(t[0] - p[7]); sv += (double)(t[1] - t[0]);
      const unsigned long long as = t[2] < t[4] ? t[2] : t[4], ae = t[3] > t[5] ? t[3] : t[5];
      entry += (double)(as - t[1]); a1 += (double)(t[3] - t[2]); a0 += (double)(t[5] - t[4]); join += (double)(t[6] - ae);
      solve += (double)(t[7] - t[6]);
      m++;
    }
    std::vector<unsigned> qc(32768 * 4);
    cudaMemcpy(qc.data(), (char*)c->timeline.p + 8 * (16 + 8 * 4000), 4 * 32768 * 4, cudaMemcpyDeviceToHost);
    for (int kind = 0; kind < 2; kind++) {  // the five slowest queries: total = setup + search + fit + store
      std::vector<std::pair<unsigned, int>> v;
      for (int i = 0; i < 16384; i++) if (qc[kind * 16384 + i]) v.push_back({qc[kind * 16384 + i], i});
      std::sort(v.begin(), v.end());
      for (size_t j = v.size() > 5 ? v.size() - 5 : 0; j < v.size(); j++) {
        const int i = kind * 16384 + v[j].second;
        fprintf(stderr, "  kind %d slow query %d: total %u setup %u search %u through-fit %u\n", kind, v[j].second, qc[i], qc[32768 + i],
                qc[65536 + i], qc[98304 + i]);
      }
    }
    for (int kind = 0; kind < 2; kind++) {
      double su = 0, se = 0; int nn = 0;
      for (int i = 0; i < 16384; i++) if (qc[65536 + kind * 16384 + i]) { su += qc[32768 + kind * 16384 + i]; se += qc[65536 + kind * 16384 + i]; nn++; }
      if (nn) fprintf(stderr, "association kind %d: mean setup %.0f, search %.0f cycles\n", kind, su / nn, se / nn);
      std::vector<unsigned> v;
      for (int i = 0; i < 16384; i++) if (qc[kind * 16384 + i]) v.push_back(qc[kind * 16384 + i]);
      if (v.empty()) continue;
      std::sort(v.begin(), v.end());
      double sum = 0; for (unsigned x : v) sum += x;
      fprintf(stderr, "association kind %d, last scan: %zu queries, cycles per query mean %.0f p50 %u p90 %u p99 %u max %u\n", kind, v.size(),
              sum / v.size(), v[v.size() / 2], v[v.size() * 9 / 10], v[v.size() * 99 / 100], v.back());
    }
    if (m) fprintf(stderr, "device timeline (us, %d scans, %d outer iterations in total): prev solve end -> split/voxel start %.1f | split/voxel %.1f | -> first association start %.1f | "
                   "plane association %.1f, line association %.1f | -> solve start %.1f | solve %.1f\n", m, n, gap / m / 1e3, sv / m / 1e3,
                   entry / m / 1e3, a1 / m / 1e3, a0 / m / 1e3, join / m / 1e3, solve / m / 1e3);
  }
#endif
  const double* h_poses = reinterpret_cast<const double*>(hp);
  const double* h_stats = h_poses + (size_t)n_scans * 16;
  const int* h_counts = reinterpret_cast<const int*>(h_poses + out_doubles);
  for (int k = 0; k < n_scans; k++) {
    if (h_counts[8 * k + 4] || h_counts[8 * k + 5]) { *first_bad = k; break; }
    memcpy(R.poses_out + 16 * (size_t)k, h_poses + 16 * (size_t)k, 16 * sizeof(double));
    if (R.counts_out) memcpy(R.counts_out + 4 * (size_t)k, h_counts + 8 * (size_t)k, 4 * sizeof(int));
    c->launches += c->chain_launches_per_iter * (long long)h_stats[8 * (size_t)k];  // launches inside the WHILE body
  }
  return MML_OK;
}

// called by mml_ctx_destroy: streams, events and buffers of the pipelined runner
void mml_odom_destroy(mml_ctx* c) {
  if (!c->odom) return;
  Odom* o = static_cast<Odom*>(c->odom);
  for (int k = 0; k < kLanes; k++)
    if (o->lane[k]) { cudaStreamSynchronize(o->lane[k]); if (k > 0) cudaStreamDestroy(o->lane[k]); }
  if (o->copy) { cudaStreamSynchronize(o->copy); cudaStreamDestroy(o->copy); }
  if (o->sv) { cudaStreamSynchronize(o->sv); cudaStreamDestroy(o->sv); }
  if (o->sv_done) cudaEventDestroy(o->sv_done);
  if (o->scan_done) cudaEventDestroy(o->scan_done);
  for (int k = 0; k < kSlots; k++) {
    Slot& S = o->slot[k];
    cudaEventDestroy(S.copied); cudaEventDestroy(S.done); cudaEventDestroy(S.consumed);
    mml::DevBuf* b[] = {&S.label, &S.counters, &S.in_xyzi, &S.in_line, &S.in_s, &S.idx};
    for (auto* x : b) x->release();
  }
  for (int k = 0; k < kLanes; k++) {
    FeScratch& f = o->scratch[k];
    mml::DevBuf* b[] = {&f.srt_xyzi, &f.srt_src, &f.srt_line, &f.chunk_tab, &f.chunk_hist, &f.line_start, &f.line_count,
                        &f.curv, &f.refl, &f.attr, &f.sort_ind, &f.refl_ind};
    for (auto* x : b) x->release();
    f.pin_small.release();
  }
  o->state.release();
  o->out.release();
  o->host.release();
  delete o;
  c->odom = nullptr;
}

extern "C" {

// Run the odometry loop over n_scans scans. xyzi/line/s are arrays of per-scan pointers (device pointers when
// host_buffers == 0, host — ideally pinned — pointers otherwise). T_init16 = pose of the frame before the first
// scan, T_prev16 the one before that (constant-velocity seed). poses_out: n_scans x 16 (row-major T_wb).
// total_ms (optional): CUDA-event time of the whole run on the context's stream.
int mml_odom_run(mml_ctx* c, const void* const* xyzi, const void* const* line, const void* const* s, const int* n_pts,
                 int n_scans, int n_lines, int host_buffers, const double* T_init16, const double* T_prev16,
                 const double* exTlb16, float leaf_corner, float leaf_surf, const mml_est_params* prm, double* poses_out,
                 float* total_ms, int* counts_out /* n_scans x 4, optional */) {
  if (!c || n_scans < 0 || !T_init16 || !T_prev16 || !exTlb16 || !poses_out) return MML_ERR_INVALID;
  if (n_scans > 0 && (!xyzi || !line || !n_pts)) return MML_ERR_INVALID;
  {
    bool any_map = false;
    for (int k = 0; k < 4; k++) any_map = any_map || c->maps[k].valid;
    if (!any_map) return mml_fail(c, MML_ERR_STATE, "mml_odom_run: no feature map set");
  }
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
  MML_CUDA(c, cudaStreamSynchronize(st));
  for (int k = 0; k < kLanes; k++) MML_CUDA(c, cudaStreamSynchronize(o->lane[k]));
  MML_CUDA(c, cudaStreamSynchronize(o->copy));
  for (int k = 0; k < kSlots; k++) o->slot[k].used = false;
  if (total_ms) MML_CUDA(c, cudaEventRecord(c->ev0, st));
  RunArgs R = {xyzi, line, s, n_pts, n_scans, n_lines, host_buffers, exTlb16, leaf_corner, leaf_surf, prm, poses_out, counts_out};
  const bool classic = getenv("MML_ODOM_CLASSIC") && atoi(getenv("MML_ODOM_CLASSIC")) != 0;
  int first = 0;
  if (!classic && !o->big_scans) MML_CHECK(run_chained(c, o, R, T_init16, T_prev16, &first));
  if (first < n_scans) {
    // classic driver from the first scan the chained one could not finish, seeded with the poses before it
    const double* Ti = first >= 1 ? poses_out + 16 * (size_t)(first - 1) : T_init16;
    const double* Tp = first >= 2 ? poses_out + 16 * (size_t)(first - 2) : (first == 1 ? T_init16 : T_prev16);
    for (int k = 0; k < kSlots; k++) o->slot[k].used = false;
    MML_CHECK(run_classic(c, o, R, first, Ti, Tp));
  }
  if (total_ms) {
    MML_CUDA(c, cudaEventRecord(c->ev1, st));
    MML_CUDA(c, cudaEventSynchronize(c->ev1));
    MML_CUDA(c, cudaEventElapsedTime(total_ms, c->ev0, c->ev1));
  }
  return MML_OK;
}

}  // extern "C"
