// C-ABI of libmmloam_b200.so (include/mmloam_b200.h): host-buffer entry points stage through
// pinned memory, run the device pipeline on the context's stream and copy results back.
#include "common.cuh"
#include "eststate.cuh"
#include <math.h>
#include <stdlib.h>

// device-resident stages (extract.cu, geometry.cu, associate.cu, accumulate.cu)
namespace mml { struct EstState; }
int mml_extract_device(mml_ctx* ctx, const float4* xyzi_d, const uint16_t* line_d, const int* scan_off, int n_scans,
                       int n_lines, uint8_t* label_d, bool force_sequential);
int mml_undistort_device(mml_ctx* ctx, float4* pts_d, const float* s_d, int n, const double* dR9, const double* dt3);
int mml_label_split_device(mml_ctx* ctx, const float4* pts_d, const uint8_t* label_d, int n, float4* corner_d,
                           float4* surf_d, int* cnt_d);
int mml_voxel_device(mml_ctx* ctx, const float4* pts_d, const int* n_dev, int n_max, float leaf, float4* out_d, int* m_dev);
int mml_velo_ring_time_device(mml_ctx* ctx, const float4* pts_d, int n, const float* first_last_xy, int16_t* ring_d,
                              float* reltime_d);
int mml_hori_filter_device(mml_ctx* ctx, const uint32_t* off_d, const float* xyz_d, const uint8_t* line_d, int n,
                           uint32_t last_offset, uint8_t* keep_d, float* reltime_d);
int mml_map_set_device(mml_ctx* ctx, int kind, const float4* pts_d, int m, const int* cen3, float cell_hint, const float* bbox6 = nullptr);
int mml_associate_launch(mml_ctx* ctx, int kind, const double* T16, float thres, const double* T_dev,
                         const float* thres_dev, const int* gate, const int* nq_dev, int cap);
int mml_export_features(mml_ctx* ctx, int kind, int nq, double* out_dev);
int mml_accumulate_launch(mml_ctx* ctx, const double* x6, const double* T_bl16, double lidar_m, double w_tan,
                          double huber_a, mml::EstState* st_dev, const int* n_dev, int cap_line, int cap_plane,
                          const double* wide_line_dev, const double* wide_plane_dev);
int mml_sort_queries_device(mml_ctx* ctx, float4* q_d, int n, mml::DevBuf& perm_buf);
int mml_split_voxel_capacity();
namespace mml { struct SvChain; }
void mml_odom_destroy(mml_ctx* c);
void mml_local_map_destroy(mml_ctx* c);
void mml_global_map_destroy(mml_ctx* c);
void mml_window_destroy(mml_ctx* c);
int mml_split_voxel_device(mml_ctx* ctx, const float4* pts_d, const float* s_d, const uint8_t* label_d, int n,
                           const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, float4* corner_out,
                           float4* surf_out, int* counts_d, const mml::SvChain* chain = nullptr);

using namespace mml;

static int round_cap(int n) {  // capacity classes keep the captured graph reusable across scans
  int c = 1024;
  while (c < n) c += c / 2 > 4096 ? 4096 : c;
  return c;
}

extern "C" {

int mml_version(void) { return 100; }

int mml_ctx_create(int device, int stream_count, mml_ctx** out) {
  if (!out) return MML_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return MML_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return MML_ERR_NO_DEVICE;
  mml_ctx* c = new mml_ctx();
  c->device = device;
  // the matcher's streams outrank the extraction stream of the pipelined loop (odometry.cu): when both have CTAs
  // waiting, the critical path gets the SMs first
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
    delete c;
    return MML_ERR_CUDA;
  }
  for (int i = 1; i < stream_count; i++) {
    cudaStream_t s;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess) c->extra_streams.push_back(s);
  }
  cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_hi);
  cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
  cudaEventCreate(&c->ev0);
  cudaEventCreate(&c->ev1);
  *out = c;
  return MML_OK;
}

int mml_ctx_destroy(mml_ctx* c) {
  if (!c) return MML_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->est_graph) cudaGraphExecDestroy(c->est_graph);
  if (c->chain_graph) cudaGraphExecDestroy(c->chain_graph);
  mml_odom_destroy(c);
  mml_local_map_destroy(c);
  mml_global_map_destroy(c);
  mml_window_destroy(c);
  if (c->stream_fe) { cudaStreamSynchronize(c->stream_fe); cudaStreamDestroy(c->stream_fe); }
  mml::DevBuf* bufs[] = {&c->in_xyzi, &c->in_line, &c->in_s, &c->in_label, &c->srt_xyzi, &c->srt_src, &c->srt_line,
                         &c->chunk_tab, &c->chunk_hist, &c->line_start, &c->line_count, &c->curv, &c->refl, &c->attr,
                         &c->sort_ind, &c->refl_ind, &c->counters, &c->tmp_a, &c->tmp_b, &c->tmp_c, &c->tmp_d, &c->tmp_e, &c->scan_state, &c->msg_raw,
                         &c->vox_keys[0], &c->vox_keys[1], &c->vox_vals[0], &c->vox_vals[1], &c->vox_hist, &c->vox_bbox,
                         &c->corner_raw, &c->surf_raw, &c->sv_bbox, &c->q_corner, &c->q_surf, &c->f_line, &c->f_plane,
                         &c->acc_partials, &c->acc_out, &c->est_state, &c->assoc_stats, &c->frame_cnt, &c->export_buf};
  for (auto* b : bufs) b->release();
  for (int k = 0; k < 4; k++) {
    c->maps[k].pts.release();
    c->maps[k].cell_start.release();
    c->maps[k].cube_count.release();
    c->maps[k].pts2.release();
    c->maps[k].cell_start2.release();
  }
  c->grid_table.release();
  mml_shard_close(c);
  c->grid_table_pin.release();
  c->pin_in.release();
  c->pin_out.release();
  c->pin_small.release();
  c->pin_flags.release();
  for (auto s : c->extra_streams) cudaStreamDestroy(s);
  c->pre_knn[0].release();
  c->pre_knn[1].release();
  c->perm[0].release();
  c->perm[1].release();
  c->assoc_part[0].release();
  c->assoc_part[1].release();
  cudaEventDestroy(c->ev_fork);
  cudaEventDestroy(c->ev_join);
  cudaStreamDestroy(c->stream2);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaStreamDestroy(c->stream);
  delete c;
  return MML_OK;
}

const char* mml_last_error(const mml_ctx* c) { return c ? c->err.c_str() : "null context"; }
long long mml_launch_count(const mml_ctx* c) { return c ? c->launches : 0; }
void* mml_stream_handle(mml_ctx* c) { return c ? (void*)c->stream : nullptr; }

int mml_sync(mml_ctx* c) {
  if (!c) return MML_ERR_INVALID;
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  return MML_OK;
}

int mml_timer_start(mml_ctx* c) {
  if (!c) return MML_ERR_INVALID;
  MML_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  return MML_OK;
}
int mml_timer_stop_ms(mml_ctx* c, float* ms) {
  if (!c || !ms) return MML_ERR_INVALID;
  MML_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  MML_CUDA(c, cudaEventSynchronize(c->ev1));
  MML_CUDA(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return MML_OK;
}

void mml_est_params_default(mml_est_params* p) {
  p->max_outer = 5;
  p->max_inner = 10;
  p->lidar_m = 1.5e-3;
  p->plan_weight_tan = 0.0;
  p->thres0 = 25.0;
  p->thres1 = 10.0;
  p->thres2 = 1.0;
  p->use_huber = 1;
  p->map_update = 0;
}

// association / estimation without any valid feature map: MML_ERR_STATE (the header documents it)
static int require_map(mml_ctx* c, const char* who) {
  for (int k = 0; k < 4; k++)
    if (c->maps[k].valid) return MML_OK;
  return mml_fail(c, MML_ERR_STATE, (std::string(who) + ": no feature map set (mml_map_set / mml_local_map_push first)").c_str());
}

// ---- staging helpers -----------------------------------------------------------------
static int upload(mml_ctx* c, mml::DevBuf& dst, const void* src, size_t bytes) {
  MML_CUDA(c, dst.reserve(bytes ? bytes : 16));
  if (bytes) MML_CUDA(c, cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return MML_OK;
}
static int download(mml_ctx* c, void* dst, const void* src_d, size_t bytes) {
  if (bytes) MML_CUDA(c, cudaMemcpyAsync(dst, src_d, bytes, cudaMemcpyDeviceToHost, c->stream));
  return MML_OK;
}

int mml_extract_features_batch(mml_ctx* c, const float* xyzi, const uint16_t* line_id, const int* scan_offsets,
                               int n_scans, int n_lines, uint8_t* out_label, int* out_n_sharp, int* out_n_flat) {
  if (!c || !scan_offsets || n_scans < 0) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  const int n = n_scans > 0 ? scan_offsets[n_scans] : 0;
  if (n > 0 && (!xyzi || !line_id || !out_label)) return MML_ERR_INVALID;
  MML_CHECK(upload(c, c->in_xyzi, xyzi, sizeof(float) * 4 * (size_t)n));
  MML_CHECK(upload(c, c->in_line, line_id, sizeof(uint16_t) * (size_t)n));
  MML_CUDA(c, c->in_label.reserve((size_t)n + 16));
  std::vector<int> cnt(2 * (size_t)n_scans + 2, 0);
  for (int attempt = 0; attempt < 3; attempt++) {
    MML_CHECK(mml_extract_device(c, c->in_xyzi.as<float4>(), c->in_line.as<uint16_t>(), scan_offsets, n_scans, n_lines,
                                 c->in_label.as<uint8_t>(), c->sel_tier >= 2));
    MML_CHECK(download(c, cnt.data(), c->counters.p, sizeof(int) * (2 * (size_t)n_scans + 1)));
    MML_CUDA(c, cudaStreamSynchronize(c->stream));
    if (!cnt[2 * (size_t)n_scans]) break;  // no line overflowed the selection kernel's shared-memory tier
    if (c->sel_tier < 2) c->sel_tier++;
  }
  MML_CHECK(download(c, out_label, c->in_label.p, (size_t)n));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int s = 0; s < n_scans; s++) {
    if (out_n_sharp) out_n_sharp[s] = cnt[2 * s];
    if (out_n_flat) out_n_flat[s] = cnt[2 * s + 1];
  }
  return MML_OK;
}

int mml_extract_features(mml_ctx* c, const float* xyzi, const uint16_t* line_id, int n, int n_lines,
                         uint8_t* out_label, int* out_n_sharp, int* out_n_flat) {
  if (n < 0) return MML_ERR_INVALID;
  const int off[2] = {0, n};
  return mml_extract_features_batch(c, xyzi, line_id, off, 1, n_lines, out_label, out_n_sharp, out_n_flat);
}

int mml_velo_ring_time(mml_ctx* c, const float* xyzi, int n, int16_t* line_out, float* reltime_out) {
  if (!c || n < 0) return MML_ERR_INVALID;
  if (n == 0) return MML_OK;
  cudaSetDevice(c->device);
  MML_CHECK(upload(c, c->in_xyzi, xyzi, sizeof(float) * 4 * (size_t)n));
  MML_CUDA(c, c->tmp_a.reserve(sizeof(int16_t) * (size_t)n));
  MML_CUDA(c, c->tmp_b.reserve(sizeof(float) * (size_t)n));
  const float fl[4] = {xyzi[0], xyzi[1], xyzi[4 * (size_t)(n - 1)], xyzi[4 * (size_t)(n - 1) + 1]};
  MML_CHECK(mml_velo_ring_time_device(c, c->in_xyzi.as<float4>(), n, fl, c->tmp_a.as<int16_t>(), c->tmp_b.as<float>()));
  MML_CHECK(download(c, line_out, c->tmp_a.p, sizeof(int16_t) * (size_t)n));
  MML_CHECK(download(c, reltime_out, c->tmp_b.p, sizeof(float) * (size_t)n));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  return MML_OK;
}

int mml_hori_filter(mml_ctx* c, const uint32_t* offset_time, const float* xyz3, const uint8_t* line, int n, uint8_t* keep,
                    float* reltime_out) {
  if (!c || n < 0) return MML_ERR_INVALID;
  if (n == 0) return MML_OK;
  cudaSetDevice(c->device);
  MML_CHECK(upload(c, c->tmp_a, offset_time, sizeof(uint32_t) * (size_t)n));
  MML_CHECK(upload(c, c->tmp_b, xyz3, sizeof(float) * 3 * (size_t)n));
  MML_CHECK(upload(c, c->in_line, line, (size_t)n));
  MML_CUDA(c, c->in_label.reserve((size_t)n + 16));
  MML_CUDA(c, c->in_s.reserve(sizeof(float) * (size_t)n));
  MML_CHECK(mml_hori_filter_device(c, c->tmp_a.as<uint32_t>(), c->tmp_b.as<float>(), c->in_line.as<uint8_t>(), n,
                                   offset_time[n - 1], c->in_label.as<uint8_t>(), c->in_s.as<float>()));
  MML_CHECK(download(c, keep, c->in_label.p, (size_t)n));
  MML_CHECK(download(c, reltime_out, c->in_s.p, sizeof(float) * (size_t)n));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  return MML_OK;
}

int mml_undistort(mml_ctx* c, float* xyzi, const float* s, int n, const double* dR9, const double* dt3) {
  if (!c || n < 0 || !dR9 || !dt3) return MML_ERR_INVALID;
  if (n == 0) return MML_OK;
  cudaSetDevice(c->device);
  MML_CHECK(upload(c, c->in_xyzi, xyzi, sizeof(float) * 4 * (size_t)n));
  MML_CHECK(upload(c, c->in_s, s, sizeof(float) * (size_t)n));
  MML_CHECK(mml_undistort_device(c, c->in_xyzi.as<float4>(), c->in_s.as<float>(), n, dR9, dt3));
  MML_CHECK(download(c, xyzi, c->in_xyzi.p, sizeof(float) * 4 * (size_t)n));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  return MML_OK;
}

int mml_voxel_downsample(mml_ctx* c, const float* xyzi, int n, float leaf, float* out, int* m_out) {
  if (!c || n < 0 || !(leaf > 0.f) || !m_out) return MML_ERR_INVALID;
  *m_out = 0;
  if (n == 0) return MML_OK;
  cudaSetDevice(c->device);
  MML_CHECK(upload(c, c->corner_raw, xyzi, sizeof(float) * 4 * (size_t)n));
  MML_CUDA(c, c->q_corner.reserve(sizeof(float4) * (size_t)n));
  MML_CUDA(c, c->frame_cnt.reserve(64));
  int* cnt = c->frame_cnt.as<int>();
  MML_CUDA(c, cudaMemcpyAsync(cnt + 4, &n, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  MML_CHECK(mml_voxel_device(c, c->corner_raw.as<float4>(), cnt + 4, n, leaf, c->q_corner.as<float4>(), cnt + 5));
  int m = 0;
  MML_CHECK(download(c, &m, cnt + 5, sizeof(int)));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  MML_CHECK(download(c, out, c->q_corner.p, sizeof(float) * 4 * (size_t)m));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  *m_out = m;
  return MML_OK;
}

int mml_map_set(mml_ctx* c, int kind, const float* xyzi, int m, const int* cube_centre3) {
  if (!c || m < 0) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  MML_CHECK(upload(c, c->tmp_a, xyzi, sizeof(float) * 4 * (size_t)m));
  return mml_map_set_device(c, kind, c->tmp_a.as<float4>(), m, cube_centre3, 0.f);
}

// the same with the points already resident in HBM (float4 xyzi): the kernel-only cost of a map build
int mml_map_set_dev(mml_ctx* c, int kind, const void* xyzi_dev, int m, const int* cube_centre3) {
  if (!c || m < 0 || (m && !xyzi_dev)) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  return mml_map_set_device(c, kind, static_cast<const float4*>(xyzi_dev), m, cube_centre3, 0.f);
}

// like mml_map_set with an explicit cell edge (0 = automatic); used by the roofline sweep
int mml_map_set_ex(mml_ctx* c, int kind, const float* xyzi, int m, const int* cube_centre3, float cell) {
  if (!c || m < 0) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  MML_CHECK(upload(c, c->tmp_a, xyzi, sizeof(float) * 4 * (size_t)m));
  return mml_map_set_device(c, kind, c->tmp_a.as<float4>(), m, cube_centre3, cell);
}

int mml_map_info(mml_ctx* c, int kind, double* info8) {
  if (!c || kind < 0 || kind > 3 || !info8) return MML_ERR_INVALID;
  const mml::GridMap& M = c->maps[kind];
  info8[0] = M.valid; info8[1] = M.m; info8[2] = M.cell; info8[3] = M.dim[0]; info8[4] = M.dim[1]; info8[5] = M.dim[2];
  info8[6] = (double)M.ncell; info8[7] = M.k_per_cube;
  return MML_OK;
}

int mml_frame_set(mml_ctx* c, const float* corner_xyzi, int n_corner, const float* surf_xyzi, int n_surf) {
  if (!c || n_corner < 0 || n_surf < 0) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  MML_CUDA(c, c->q_corner.reserve(sizeof(float4) * (size_t)round_cap(n_corner)));
  MML_CUDA(c, c->q_surf.reserve(sizeof(float4) * (size_t)round_cap(n_surf)));
  if (n_corner) MML_CUDA(c, cudaMemcpyAsync(c->q_corner.p, corner_xyzi, sizeof(float4) * (size_t)n_corner, cudaMemcpyHostToDevice, c->stream));
  if (n_surf) MML_CUDA(c, cudaMemcpyAsync(c->q_surf.p, surf_xyzi, sizeof(float4) * (size_t)n_surf, cudaMemcpyHostToDevice, c->stream));
  c->n_corner = n_corner;
  c->n_surf = n_surf;
  // map-sized query sets are put into spatial (Morton) order once per frame; scans already are coherent
  c->has_perm[0] = c->has_perm[1] = false;
  static const int sort_min = getenv("MML_SORT_MIN") ? atoi(getenv("MML_SORT_MIN")) : 32768;
  if (n_corner > sort_min) { MML_CHECK(mml_sort_queries_device(c, c->q_corner.as<float4>(), n_corner, c->perm[0])); c->has_perm[0] = true; }
  if (n_surf > sort_min) { MML_CHECK(mml_sort_queries_device(c, c->q_surf.as<float4>(), n_surf, c->perm[1])); c->has_perm[1] = true; }
  MML_CUDA(c, c->frame_cnt.reserve(64));
  const int cnt[2] = {n_corner, n_surf};
  MML_CUDA(c, cudaMemcpyAsync(c->frame_cnt.p, cnt, sizeof(cnt), cudaMemcpyHostToDevice, c->stream));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  return MML_OK;
}

static int read_assoc_stats(mml_ctx* c, int* n_line, int* n_plane, double* moment9, int* n_normals) {
  double h[20];
  MML_CHECK(download(c, h, c->assoc_stats.p, sizeof(h)));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  const int* ints = reinterpret_cast<const int*>(h + 16);
  if (n_line) *n_line = ints[0];
  if (n_plane) *n_plane = ints[1];
  if (moment9) {
    const double* m = h + 8;
    const double M[9] = {m[0], m[1], m[2], m[1], m[3], m[4], m[2], m[4], m[5]};
    for (int i = 0; i < 9; i++) moment9[i] = M[i];
  }
  if (n_normals) *n_normals = ints[1];
  return MML_OK;
}

int mml_frame_associate_async(mml_ctx* c, const double* T_wl16, double thres_dist, int repeat) {
  if (!c || !T_wl16) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  // line || plane like the reference's two association threads (EST.cpp:1265-1299): the plane kind runs on the
  // context's second stream, forked from and joined back into the main one
  cudaStream_t st = c->stream, st2 = c->stream2;
  for (int r = 0; r < repeat; r++) {
    MML_CUDA(c, cudaEventRecord(c->ev_fork, st));
    MML_CUDA(c, cudaStreamWaitEvent(st2, c->ev_fork, 0));
    c->stream = st2;
    const int rc1 = mml_associate_launch(c, 1, T_wl16, (float)thres_dist, nullptr, nullptr, nullptr, nullptr, c->n_surf);
    c->stream = st;
    MML_CHECK(rc1);
    MML_CHECK(mml_associate_launch(c, 0, T_wl16, (float)thres_dist, nullptr, nullptr, nullptr, nullptr, c->n_corner));
    MML_CUDA(c, cudaEventRecord(c->ev_join, st2));
    MML_CUDA(c, cudaStreamWaitEvent(st, c->ev_join, 0));
  }
  return MML_OK;
}

// one kind only (0 line / 1 plane), for timing a single association kernel
int mml_frame_associate_kind_async(mml_ctx* c, int kind, const double* T_wl16, double thres_dist, int repeat) {
  if (!c || !T_wl16 || kind < 0 || kind > 1) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  for (int r = 0; r < repeat; r++)
    MML_CHECK(mml_associate_launch(c, kind, T_wl16, (float)thres_dist, nullptr, nullptr, nullptr, nullptr, kind ? c->n_surf : c->n_corner));
  return MML_OK;
}

int mml_frame_associate(mml_ctx* c, const double* T_wl16, double thres_dist, int* n_line, int* n_plane,
                        double* normal_moment9, int* n_normals) {
  if (!c || !T_wl16) return MML_ERR_INVALID;
  MML_CHECK(require_map(c, "mml_frame_associate"));
  MML_CHECK(mml_frame_associate_async(c, T_wl16, thres_dist, 1));
  return read_assoc_stats(c, n_line, n_plane, normal_moment9, n_normals);
}

int mml_frame_accumulate_async(mml_ctx* c, const double* x6, const double* T_bl16, double plan_weight_tan,
                               double huber_a, int repeat) {
  if (!c || !x6 || !T_bl16) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  for (int r = 0; r < repeat; r++)
    MML_CHECK(mml_accumulate_launch(c, x6, T_bl16, 1.5e-3, plan_weight_tan, huber_a, nullptr, nullptr, c->n_corner, c->n_surf,
                                    nullptr, nullptr));
  return MML_OK;
}

static void unpack28_host(const double* o, double* H36, double* g6, double* cost) {
  if (cost) *cost = o[0];
  if (g6) for (int i = 0; i < 6; i++) g6[i] = o[1 + i];
  if (H36) {
    int k = 7;
    for (int i = 0; i < 6; i++)
      for (int j = i; j < 6; j++) {
        H36[6 * i + j] = o[k];
        H36[6 * j + i] = o[k];
        k++;
      }
  }
}

int mml_frame_accumulate(mml_ctx* c, const double* x6, const double* T_bl16, double plan_weight_tan, double huber_a,
                         double* H36, double* g6, double* cost) {
  MML_CHECK(mml_frame_accumulate_async(c, x6, T_bl16, plan_weight_tan, huber_a, 1));
  double o[28];
  MML_CHECK(download(c, o, c->acc_out.p, sizeof(o)));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  unpack28_host(o, H36, g6, cost);
  return MML_OK;
}

int mml_frame_accumulate_partial_dev(mml_ctx* c, const double* x6, const double* T_bl16, double plan_weight_tan,
                                     double huber_a, void** partial28_dev) {
  MML_CHECK(mml_frame_accumulate_async(c, x6, T_bl16, plan_weight_tan, huber_a, 1));
  if (partial28_dev) *partial28_dev = c->acc_out.p;
  return MML_OK;
}

int mml_frame_get_features(mml_ctx* c, int kind, double* out_feat) {
  if (!c || (kind != 0 && kind != 1)) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  const int nq = kind == 0 ? c->n_corner : c->n_surf;
  if (nq <= 0) return MML_OK;
  if (!out_feat) return MML_ERR_INVALID;
  MML_CUDA(c, c->export_buf.reserve(sizeof(double) * 12 * (size_t)nq));
  MML_CHECK(mml_export_features(c, kind, nq, c->export_buf.as<double>()));
  MML_CHECK(download(c, out_feat, c->export_buf.p, sizeof(double) * 12 * (size_t)nq));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  return MML_OK;
}

int mml_associate(mml_ctx* c, int kind, const float* q_xyzi, int nq, const double* T_wl16, double thres_dist,
                  double* out_feat, int* n_feat, double* normal_moment9, int* n_normals) {
  if (!c || (kind != 0 && kind != 1) || nq < 0 || !T_wl16) return MML_ERR_INVALID;
  MML_CHECK(require_map(c, "mml_associate"));
  if (kind == 0) MML_CHECK(mml_frame_set(c, q_xyzi, nq, nullptr, 0));
  else MML_CHECK(mml_frame_set(c, nullptr, 0, q_xyzi, nq));
  MML_CHECK(mml_associate_launch(c, kind, T_wl16, (float)thres_dist, nullptr, nullptr, nullptr, nullptr, nq));
  int nl = 0, np = 0;
  MML_CHECK(read_assoc_stats(c, &nl, &np, kind == 1 ? normal_moment9 : nullptr, kind == 1 ? n_normals : nullptr));
  if (n_feat) *n_feat = kind == 0 ? nl : np;
  if (out_feat) MML_CHECK(mml_frame_get_features(c, kind, out_feat));
  return MML_OK;
}

// stateless form: features arrive as host 12-double records and are evaluated as such
int mml_accumulate(mml_ctx* c, const double* line_feat, int n_line, const double* plane_feat, int n_plane,
                   const double* x6, const double* T_bl16, double plan_weight_tan, double huber_a, double* H36,
                   double* g6, double* cost) {
  if (!c || n_line < 0 || n_plane < 0 || !x6 || !T_bl16) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  MML_CHECK(upload(c, c->tmp_a, line_feat, sizeof(double) * 12 * (size_t)n_line));
  MML_CHECK(upload(c, c->tmp_b, plane_feat, sizeof(double) * 12 * (size_t)n_plane));
  MML_CHECK(mml_accumulate_launch(c, x6, T_bl16, 1.5e-3, plan_weight_tan, huber_a, nullptr, nullptr, n_line, n_plane,
                                  c->tmp_a.as<double>(), c->tmp_b.as<double>()));
  double o[28];
  MML_CHECK(download(c, o, c->acc_out.p, sizeof(o)));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  unpack28_host(o, H36, g6, cost);
  return MML_OK;
}

// ---- cube-sharded global map over several GPUs (SURVEY.md 8 e): exchange through peer memory (eststate.cuh) ----------
static int shard_upload_desc(mml_ctx* c, double* const* peers) {
  mml::ShardDev h;
  memset(&h, 0, sizeof(h));
  h.rank = c->shard_rank; h.world = c->shard_world;
  for (int r = 0; r < c->shard_world; r++) h.peer[r] = peers[r];
  MML_CUDA(c, c->shard_dev.reserve(sizeof(mml::ShardDev)));
  MML_CUDA(c, cudaMemcpy(c->shard_dev.p, &h, sizeof(h), cudaMemcpyHostToDevice));
  return MML_OK;
}

int mml_shard_init(mml_ctx* c, int rank, int world, void* ipc_handle64_out) {
  if (!c || world < 1 || world > mml::kShardMaxWorld || rank < 0 || rank >= world) return MML_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaSetDevice(c->device);
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  c->shard_rank = rank; c->shard_world = world;
  MML_CUDA(c, c->shard_buf.reserve(sizeof(double) * mml::kShardBufDoubles));
  MML_CUDA(c, cudaMemset(c->shard_buf.p, 0, sizeof(double) * mml::kShardBufDoubles));
  if (ipc_handle64_out) {
    cudaIpcMemHandle_t hd;
    MML_CUDA(c, cudaIpcGetMemHandle(&hd, c->shard_buf.p));
    memcpy(ipc_handle64_out, &hd, 64);
  }
  double* peers[mml::kShardMaxWorld] = {};
  peers[rank] = c->shard_buf.as<double>();
  if (world == 1) return shard_upload_desc(c, peers);
  return MML_OK;
}

int mml_shard_local_ptr(mml_ctx* c, void** out) {
  if (!c || !out || !c->shard_buf.p) return MML_ERR_INVALID;
  *out = c->shard_buf.p;
  return MML_OK;
}

// every rank's handle (world x 64 bytes, rank order): other processes' buffers are opened for peer access
int mml_shard_connect_ipc(mml_ctx* c, const void* handles) {
  if (!c || !handles || !c->shard_buf.p) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  double* peers[mml::kShardMaxWorld] = {};
  for (int r = 0; r < c->shard_world; r++) {
    if (r == c->shard_rank) { peers[r] = c->shard_buf.as<double>(); continue; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, static_cast<const char*>(handles) + 64 * (size_t)r, 64);
    void* p = nullptr;
    MML_CUDA(c, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    c->shard_peer_opened[r] = p;
    peers[r] = static_cast<double*>(p);
  }
  return shard_upload_desc(c, peers);
}

// the same for ranks that live in ONE process (several contexts): plain device pointers, peer access enabled here
int mml_shard_connect_ptrs(mml_ctx* c, void* const* bufs, const int* devices) {
  if (!c || !bufs || !c->shard_buf.p) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  double* peers[mml::kShardMaxWorld] = {};
  for (int r = 0; r < c->shard_world; r++) {
    if (!bufs[r]) return MML_ERR_INVALID;
    peers[r] = static_cast<double*>(bufs[r]);
    if (devices && devices[r] != c->device) {
      const cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); MML_CUDA(c, e); }
      cudaGetLastError();
    }
  }
  return shard_upload_desc(c, peers);
}

int mml_shard_close(mml_ctx* c) {
  if (!c) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (int r = 0; r < mml::kShardMaxWorld; r++)
    if (c->shard_peer_opened[r]) { cudaIpcCloseMemHandle(c->shard_peer_opened[r]); c->shard_peer_opened[r] = nullptr; }
  c->shard_dev.release();
  c->shard_buf.release();
  c->shard_world = 1; c->shard_rank = 0;
  return MML_OK;
}

int mml_estimate(mml_ctx* c, const float* corner_xyzi, int n_corner, const float* surf_xyzi, int n_surf,
                 const double* exTlb16, double* P3, double* q_wxyz4, const mml_est_params* prm, double* stats);

// Estimator::Estimate (window size 1) against a global map sharded by 50 m cube over the ranks of mml_shard_init:
// a collective - every rank calls it with the same queries and the same start pose; each rank matches the queries
// that fall into the cubes it holds, the partial sums are exchanged through peer memory inside the kernels, and every
// rank returns the same pose (bit for bit). One host wait per outer iteration.
int mml_estimate_sharded(mml_ctx* c, const float* corner_xyzi, int n_corner, const float* surf_xyzi, int n_surf,
                         const double* exTlb16, double* P3, double* q_wxyz4, const mml_est_params* prm, double* stats) {
  if (!c || !exTlb16 || !P3 || !q_wxyz4) return MML_ERR_INVALID;
  if (!c->shard_dev.p) return mml_fail(c, MML_ERR_STATE, "mml_estimate_sharded: mml_shard_init / mml_shard_connect_* first");
  // a query whose cube lives on another rank must find NO map here: the local-map fallback (EST.cpp:283, 702) would
  // match it on every rank and the exchange would count it world times
  if (c->maps[MML_MAP_CORNER_LOCAL].valid || c->maps[MML_MAP_SURF_LOCAL].valid)
    return mml_fail(c, MML_ERR_STATE, "mml_estimate_sharded: local maps must not be set on a shard context");
  c->shard_active = true;
  const int rc = mml_estimate(c, corner_xyzi, n_corner, surf_xyzi, n_surf, exTlb16, P3, q_wxyz4, prm, stats);
  c->shard_active = false;
  if (rc != MML_OK) return rc;
  mml::ShardDev h;
  MML_CUDA(c, cudaMemcpy(&h, c->shard_dev.p, sizeof(h), cudaMemcpyDeviceToHost));
  if (h.pad) return mml_fail(c, MML_ERR_CUDA, "mml_estimate_sharded: a peer rank did not arrive (exchange timed out)");
  return MML_OK;
}

int mml_estimate(mml_ctx* c, const float* corner_xyzi, int n_corner, const float* surf_xyzi, int n_surf,
                 const double* exTlb16, double* P3, double* q_wxyz4, const mml_est_params* prm, double* stats) {
  if (!c || !exTlb16 || !P3 || !q_wxyz4) return MML_ERR_INVALID;
  if (!c->shard_active) MML_CHECK(require_map(c, "mml_estimate"));  // a shard may hold no cube at all
  mml_est_params def;
  mml_est_params_default(&def);
  if (!prm) prm = &def;
  if (n_corner < 0 && n_surf < 0 && !corner_xyzi && !surf_xyzi) {
    // the frame mml_frame_set left in HBM (queries uploaded and spatially sorted once, solved several times)
    n_corner = c->n_corner; n_surf = c->n_surf;
  } else {
    MML_CHECK(mml_frame_set(c, corner_xyzi, n_corner, surf_xyzi, n_surf));
  }
  return mml_estimate_device(c, c->frame_cnt.as<int>(), round_cap(n_corner), round_cap(n_surf), exTlb16, P3, q_wxyz4, prm,
                             stats);
}

// general path: any number of labelled points (multi-kernel split + radix-sort voxel filter)
static int scan_to_pose_general(mml_ctx* c, const void* xyzi_dev, const void* line_id_dev, const void* s_dev, int n, int n_lines,
                         const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, const double* exTlb16,
                         double* P3, double* q_wxyz4, const mml_est_params* prm, double* stats, int* out_counts) {
  if (!c || n < 0 || !exTlb16 || !P3 || !q_wxyz4) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  mml_est_params def;
  mml_est_params_default(&def);
  if (!prm) prm = &def;
  cudaStream_t st = c->stream;
  MML_CUDA(c, c->in_label.reserve((size_t)n + 16));
  MML_CUDA(c, c->tmp_e.reserve(sizeof(float4) * (size_t)(n > 0 ? n : 1)));
  const int off[2] = {0, n};
  if (c->profile) MML_CUDA(c, cudaEventRecord(c->pev[0], st));
  // labels: extracted here, or already there (the odometry loop labels the next scan on an extraction lane while this
  // one is matched and hands over the slot's label / counter buffers; the stream already waits for that lane)
  uint8_t* label_d = c->in_label.as<uint8_t>();
  const int* counters_d = c->counters.as<int>();
  if (c->pre_label) {
    label_d = c->pre_label;
    counters_d = c->pre_counters;
  } else {
    MML_CHECK(mml_extract_device(c, (const float4*)xyzi_dev, (const uint16_t*)line_id_dev, off, 1, n_lines, label_d, c->sel_tier >= 2));
  }
  MML_CUDA(c, c->frame_cnt.reserve(64));
  int* cnt = c->frame_cnt.as<int>();  // [0..1] voxel outputs, [2..3] raw split counts
  int hc[3] = {0, 0, 0};
  MML_CHECK(download(c, hc, counters_d, sizeof(hc)));
  MML_CUDA(c, cudaStreamSynchronize(st));
  if (hc[2] && c->pre_label) {  // the pre-labelled scan overflowed the tier it was labelled with: label it again here
    label_d = c->in_label.as<uint8_t>();
    counters_d = c->counters.as<int>();
    MML_CHECK(mml_extract_device(c, (const float4*)xyzi_dev, (const uint16_t*)line_id_dev, off, 1, n_lines, label_d, c->sel_tier >= 2));
    MML_CHECK(download(c, hc, counters_d, sizeof(hc)));
    MML_CUDA(c, cudaStreamSynchronize(st));
  }
  while (hc[2] && c->sel_tier < 2) {  // a line overflowed the selection kernel's tier: next tier (2 = sequential)
    c->sel_tier++;
    MML_CHECK(mml_extract_device(c, (const float4*)xyzi_dev, (const uint16_t*)line_id_dev, off, 1, n_lines, label_d, c->sel_tier >= 2));
    MML_CHECK(download(c, hc, counters_d, sizeof(hc)));
    MML_CUDA(c, cudaStreamSynchronize(st));
  }
  // undistort a copy of the scan (the caller's buffer stays untouched). Only now: the copy lives in the storage of
  // the extraction's line-sorted cloud, which every (re-)run of the extraction above overwrites.
  mml::DevBuf& wbuf = c->srt_xyzi;
  MML_CUDA(c, wbuf.reserve(sizeof(float4) * (size_t)(n > 0 ? n : 1)));
  float4* work = wbuf.as<float4>();
  if (c->profile) MML_CUDA(c, cudaEventRecord(c->pev[1], st));
  if (n) MML_CUDA(c, cudaMemcpyAsync(work, xyzi_dev, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  if (dR9 && dt3 && s_dev) MML_CHECK(mml_undistort_device(c, work, (const float*)s_dev, n, dR9, dt3));
  // label split (EST.cpp:992-1011); capacities from the counts of the extractor
  const int n_sharp = hc[0], n_flat = hc[1];
  MML_CUDA(c, c->corner_raw.reserve(sizeof(float4) * (size_t)(n_sharp + 1)));
  MML_CUDA(c, c->surf_raw.reserve(sizeof(float4) * (size_t)(n_flat + 1)));
  const int cap_c = round_cap(n_sharp), cap_s = round_cap(n_flat);
  MML_CUDA(c, c->q_corner.reserve(sizeof(float4) * (size_t)cap_c));
  MML_CUDA(c, c->q_surf.reserve(sizeof(float4) * (size_t)cap_s));
  MML_CHECK(mml_label_split_device(c, work, label_d, n, c->corner_raw.as<float4>(), c->surf_raw.as<float4>(), cnt + 2));
  MML_CHECK(mml_voxel_device(c, c->corner_raw.as<float4>(), cnt + 2, n_sharp, leaf_corner, c->q_corner.as<float4>(), cnt));
  MML_CHECK(mml_voxel_device(c, c->surf_raw.as<float4>(), cnt + 3, n_flat, leaf_surf, c->q_surf.as<float4>(), cnt + 1));
  if (c->profile) MML_CUDA(c, cudaEventRecord(c->pev[2], st));
  MML_CHECK(mml_estimate_device(c, cnt, cap_c, cap_s, exTlb16, P3, q_wxyz4, prm, stats));
  if (c->profile) {
    MML_CUDA(c, cudaEventRecord(c->pev[3], st));
    MML_CUDA(c, cudaEventSynchronize(c->pev[3]));
    for (int k = 0; k < 3; k++) {
      float ms = 0.f;
      MML_CUDA(c, cudaEventElapsedTime(&ms, c->pev[k], c->pev[k + 1]));
      c->stage_ms[k] += ms;
    }
    c->stage_n++;
  }
  if (out_counts) {
    int hv[2];
    MML_CHECK(download(c, hv, cnt, sizeof(hv)));
    MML_CUDA(c, cudaStreamSynchronize(st));
    out_counts[0] = n_sharp; out_counts[1] = n_flat; out_counts[2] = hv[0]; out_counts[3] = hv[1];
    c->n_corner = hv[0];
    c->n_surf = hv[1];
  }
  return MML_OK;
}

// extract -> fused (split + undistort + voxel) -> estimate, everything resident between stages and ONE host
// synchronisation per scan. Falls back to the general path when a line or a label class exceeds the
// shared-memory capacities of the fused kernels.
int mml_scan_to_pose_dev(mml_ctx* c, const void* xyzi_dev, const void* line_id_dev, const void* s_dev, int n, int n_lines,
                         const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, const double* exTlb16,
                         double* P3, double* q_wxyz4, const mml_est_params* prm, double* stats, int* out_counts) {
  if (!c || n < 0 || !exTlb16 || !P3 || !q_wxyz4) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  mml_est_params def;
  mml_est_params_default(&def);
  if (!prm) prm = &def;
  cudaStream_t st = c->stream;
  const int cap = mml_split_voxel_capacity();
  MML_CUDA(c, c->in_label.reserve((size_t)n + 16));
  MML_CUDA(c, c->frame_cnt.reserve(64));
  MML_CUDA(c, c->q_corner.reserve(sizeof(float4) * (size_t)cap));
  MML_CUDA(c, c->q_surf.reserve(sizeof(float4) * (size_t)cap));
  MML_CUDA(c, c->pin_flags.reserve(64));
  if (c->prefer_general)  // the caller knows the scan does not fit the fused kernels (odometry loop, big scans)
    return scan_to_pose_general(c, xyzi_dev, line_id_dev, s_dev, n, n_lines, dR9, dt3, leaf_corner, leaf_surf, exTlb16, P3,
                                q_wxyz4, prm, stats, out_counts);
  const double P_in[3] = {P3[0], P3[1], P3[2]}, q_in[4] = {q_wxyz4[0], q_wxyz4[1], q_wxyz4[2], q_wxyz4[3]};
  const int off[2] = {0, n};
  if (c->profile) MML_CUDA(c, cudaEventRecord(c->pev[0], st));
  MML_CHECK(mml_extract_device(c, (const float4*)xyzi_dev, (const uint16_t*)line_id_dev, off, 1, n_lines, c->in_label.as<uint8_t>(), false));
  if (c->profile) MML_CUDA(c, cudaEventRecord(c->pev[1], st));
  c->has_perm[0] = c->has_perm[1] = false;
  int* cnt = c->frame_cnt.as<int>();  // [0..1] voxel outputs, [2..3] raw split counts, [4] overflow
  const bool undist = dR9 && dt3 && s_dev;
  MML_CHECK(mml_split_voxel_device(c, (const float4*)xyzi_dev, undist ? (const float*)s_dev : nullptr, c->in_label.as<uint8_t>(), n,
                                   dR9, dt3, leaf_corner, leaf_surf, c->q_corner.as<float4>(), c->q_surf.as<float4>(), cnt));
  int* hf = c->pin_flags.as<int>();  // [0..2] extractor counters (+ overflow), [4..8] split/voxel counts
  MML_CUDA(c, cudaMemcpyAsync(hf, c->counters.p, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
  MML_CUDA(c, cudaMemcpyAsync(hf + 4, cnt, 5 * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (c->profile) MML_CUDA(c, cudaEventRecord(c->pev[2], st));
  MML_CHECK(mml_estimate_device(c, cnt, cap, cap, exTlb16, P3, q_wxyz4, prm, stats));  // synchronises the stream
  if (hf[2] || hf[8]) {
    // a scan line or a label class did not fit the fused kernels: redo on the general path
    for (int i = 0; i < 3; i++) P3[i] = P_in[i];
    for (int i = 0; i < 4; i++) q_wxyz4[i] = q_in[i];
    return scan_to_pose_general(c, xyzi_dev, line_id_dev, s_dev, n, n_lines, dR9, dt3, leaf_corner, leaf_surf, exTlb16, P3,
                                q_wxyz4, prm, stats, out_counts);
  }
  if (c->profile) {
    MML_CUDA(c, cudaEventRecord(c->pev[3], st));
    MML_CUDA(c, cudaEventSynchronize(c->pev[3]));
    for (int k = 0; k < 3; k++) {
      float ms = 0.f;
      MML_CUDA(c, cudaEventElapsedTime(&ms, c->pev[k], c->pev[k + 1]));
      c->stage_ms[k] += ms;
    }
    c->stage_n++;
  }
  c->n_corner = hf[4];
  c->n_surf = hf[5];
  if (out_counts) {
    out_counts[0] = hf[0]; out_counts[1] = hf[1]; out_counts[2] = hf[4]; out_counts[3] = hf[5];
  }
  return MML_OK;
}

int mml_scan_to_pose(mml_ctx* c, const float* xyzi, const uint16_t* line_id, const float* s, int n, int n_lines,
                     const double* dR9, const double* dt3, float leaf_corner, float leaf_surf, const double* exTlb16,
                     double* P3, double* q_wxyz4, const mml_est_params* prm, double* stats, int* out_counts) {
  if (!c || n < 0) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  MML_CHECK(upload(c, c->in_xyzi, xyzi, sizeof(float) * 4 * (size_t)n));
  MML_CHECK(upload(c, c->in_line, line_id, sizeof(uint16_t) * (size_t)n));
  if (s) MML_CHECK(upload(c, c->in_s, s, sizeof(float) * (size_t)n));
  return mml_scan_to_pose_dev(c, c->in_xyzi.p, c->in_line.p, s ? c->in_s.p : nullptr, n, n_lines, dR9, dt3, leaf_corner,
                              leaf_surf, exTlb16, P3, q_wxyz4, prm, stats, out_counts);
}

// per-stage CUDA-event timing of mml_scan_to_pose[_dev]: [extract, undistort+split+voxel, estimate] in ms
int mml_profile_enable(mml_ctx* c, int on) {
  if (!c) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  if (on && !c->pev[0])
    for (int k = 0; k < 4; k++) MML_CUDA(c, cudaEventCreate(&c->pev[k]));
  c->profile = on != 0;
  for (int k = 0; k < 4; k++) c->stage_ms[k] = 0.0;
  c->stage_n = 0;
  return MML_OK;
}
int mml_profile_read(mml_ctx* c, double* stage_ms3, long long* n_scans) {
  if (!c || !stage_ms3) return MML_ERR_INVALID;
  for (int k = 0; k < 3; k++) stage_ms3[k] = c->stage_ms[k];
  if (n_scans) *n_scans = c->stage_n;
  return MML_OK;
}

// device allocation helpers for callers that keep scans resident (bench.py)
int mml_dev_alloc(mml_ctx* c, size_t bytes, void** out) {
  if (!c || !out) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  MML_CUDA(c, cudaMalloc(out, bytes ? bytes : 16));
  return MML_OK;
}
int mml_dev_free(mml_ctx* c, void* p) {
  if (!c) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  MML_CUDA(c, cudaFree(p));
  return MML_OK;
}
int mml_dev_upload(mml_ctx* c, void* dst_dev, const void* src, size_t bytes) {
  if (!c) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  MML_CUDA(c, cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, c->stream));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  return MML_OK;
}

}  // extern "C"

// test / debug view of a built map level: cell-sorted points (xyz + original index) and the cell table
extern "C" int mml_map_dump(mml_ctx* c, int kind, int level, float* pts_out /*m x 4*/, int* cell_start_out /*ncell+1*/) {
  if (!c || kind < 0 || kind > 3) return MML_ERR_INVALID;
  const mml::GridMap& M = c->maps[kind];
  if (!M.valid) return MML_ERR_STATE;
  cudaSetDevice(c->device);
  const long long ncell = level == 0 ? M.ncell : (long long)M.dim2[0] * M.dim2[1] * M.dim2[2];
  const void* p = level == 0 ? M.pts.p : M.pts2.p;
  const void* cs = level == 0 ? M.cell_start.p : M.cell_start2.p;
  if (level == 1 && M.coarse <= 1) return MML_ERR_STATE;
  MML_CUDA(c, cudaMemcpyAsync(pts_out, p, sizeof(float) * 4 * (size_t)M.m, cudaMemcpyDeviceToHost, c->stream));
  MML_CUDA(c, cudaMemcpyAsync(cell_start_out, cs, sizeof(int) * ((size_t)ncell + 1), cudaMemcpyDeviceToHost, c->stream));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  return MML_OK;
}
extern "C" int mml_map_dims(mml_ctx* c, int kind, int* dims7 /*dim xyz, dim2 xyz, coarse*/, double* org3) {
  if (!c || kind < 0 || kind > 3) return MML_ERR_INVALID;
  const mml::GridMap& M = c->maps[kind];
  for (int a = 0; a < 3; a++) { dims7[a] = M.dim[a]; dims7[3 + a] = M.dim2[a]; org3[a] = M.org_d[a]; }
  dims7[6] = M.coarse;
  return MML_OK;
}

// Device-resident batched extraction (kernels only, asynchronous): scans concatenated in xyzi_dev / line_dev,
// scan_offsets on the host. Labels go to label_dev. Used for the batched roofline figure (SURVEY.md §8 d).
extern "C" int mml_extract_features_batch_dev(mml_ctx* c, const void* xyzi_dev, const void* line_dev, const int* scan_offsets,
                                              int n_scans, int n_lines, void* label_dev) {
  if (!c || !scan_offsets || n_scans < 0) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  return mml_extract_device(c, (const float4*)xyzi_dev, (const uint16_t*)line_dev, scan_offsets, n_scans, n_lines,
                            (uint8_t*)label_dev, false);
}
