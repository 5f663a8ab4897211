// mmloam_b200 internal: context, device buffers, launch helpers. sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/mmloam_b200.h"

namespace mml {

constexpr int kNumSMs = 148;            // B200: 2 dies x 74 SMs
constexpr int kParts = 50;              // thPartNum, FE.cpp:356
constexpr int kCubeW = 21, kCubeH = 11, kCubeD = 21;  // MM.h:117-119
constexpr int kNumCubes = kCubeW * kCubeH * kCubeD;
constexpr int kCubeNone = 5000;         // MM.cpp:601

// Several contexts may be driven from several host threads (the reference calls the extractor from six threads and
// matches on two). A stream capture must not overlap another thread's allocation + legacy-stream memset or another
// capture: both take this process-wide lock (recursive: a capture may find a buffer to grow).
inline std::recursive_mutex& capture_mutex() {
  static std::recursive_mutex m;
  return m;
}

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    std::lock_guard<std::recursive_mutex> lk(capture_mutex());
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return e;
    cap = want;
    // fresh scratch is zero-filled (tickets and counters rely on it); growth is rare
    e = cudaMemset(p, 0, want);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// One spatial-hash map (a "kind" of mml_map_set): points sorted by cell.
struct GridMap {
  int m = 0;                 // points
  float cell = 0.f;          // cell edge
  float inv_cell = 0.f;
  double org_d[3] = {0, 0, 0};  // lower corner of cell (0,0,0)
  int dim[3] = {0, 0, 0};
  int k_per_cube = 0;        // cells per 50 m cube edge (global kinds)
  int cube_lo[3] = {0, 0, 0};
  long long ncell = 0;
  bool global = false;       // true: honour the 50 m cube rule
  int cen[3] = {10, 5, 10};
  DevBuf pts;                // float4[m] cell-sorted (w = original index bits)
  DevBuf cell_start;         // int[ncell+1]
  DevBuf cube_count;         // int[kNumCubes] points per 50 m cube (global kinds)
  int coarse = 1;            // coarse level: cells `coarse` times larger (1 = none)
  int dim2[3] = {0, 0, 0};
  DevBuf pts2, cell_start2;
  bool valid = false;
  // cell edge the occupancy trial chose last time and the point count it chose it for: a rebuild of a map of similar
  // size (the local map follows the trajectory update by update) skips the trial pass and its host synchronisation
  float auto_cell = 0.f;
  int auto_m = 0;
};

struct EstDev;  // device-side solver state (accumulate.cu)

}  // namespace mml

struct mml_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;       // second captured stream (plane association runs beside line association)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<cudaStream_t> extra_streams;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  long long launches = 0;

  // scratch (grow-only)
  mml::DevBuf in_xyzi, in_line, in_s, in_label;          // staged scan
  mml::DevBuf msg_raw;                                   // raw message bytes / packed output clouds (msgpack.cu)
  mml::DevBuf srt_xyzi, srt_src, srt_line;               // line-sorted scan
  mml::DevBuf chunk_tab, chunk_hist, line_start, line_count;
  mml::DevBuf curv, refl, attr, sort_ind, refl_ind;      // per-point extraction state
  mml::DevBuf counters;                                  // small int scratch
  int sel_tier = 0;                                      // shared-memory tier of the selection kernel (extract.cu)
  int* counters_alt = nullptr;                           // when set, extraction writes its counters here
  mml::DevBuf tmp_a, tmp_b, tmp_c, tmp_d, tmp_e;         // generic
  mml::DevBuf scan_state;                                // tile status words of the look-back scan (sort.cuh)
  mml::DevBuf vox_keys[2], vox_vals[2], vox_hist, vox_bbox;
  mml::DevBuf corner_raw, surf_raw;                      // label-split clouds
  mml::DevBuf timeline;                                  // MML_TIMELINE debug stamps
  mml::DevBuf sv_bbox;                                   // per-CTA boxes of the clustered split/voxel launch
  mml::PinBuf pin_in, pin_out, pin_small, pin_flags;

  // resident maps
  mml::GridMap maps[4];
  // device-resident copy of the maps' descriptors for captured association launches (associate.cu, windowsolve.cu)
  mml::DevBuf grid_table;
  mml::PinBuf grid_table_pin;
  bool grid_table_dirty = true;
  unsigned grid_table_gen = 0;
  bool assoc_table_mode = false;
  // cube-sharded map over several GPUs: exchange buffer of this rank, device-side descriptor (eststate.cuh ShardDev)
  mml::DevBuf shard_buf, shard_dev;
  bool shard_active = false;
  int shard_rank = 0, shard_world = 1;
  void* shard_peer_opened[8] = {};

  // frame slot (queries + features)
  mml::DevBuf q_corner, q_surf;   // float4
  mml::DevBuf perm[2];            // spatial sort permutation of large query sets (framesort.cu)
  bool has_perm[2] = {false, false};
  mml::DevBuf pre_knn[2];         // map-sized sets: neighbour positions + status left by k_knn_walk for the fit kernel
  int n_corner = 0, n_surf = 0;
  mml::DevBuf f_line, f_plane;    // compact features (see associate.cu)
  mml::DevBuf acc_partials, acc_out, est_state;
  mml::DevBuf assoc_stats;        // ints + doubles
  mml::DevBuf assoc_part[2];      // per-CTA moment partials of the line / plane association
  cudaGraphExec_t est_graph = nullptr;
  long long est_graph_key = 0;
  uint8_t* pre_label = nullptr;                          // general path: labels / counters of the scan already extracted by the caller
  const int* pre_counters = nullptr;
  bool prefer_general = false;                           // mml_scan_to_pose[_dev]: skip the fused attempt (set by the odometry loop for big scans)
  int solve_small = 1;                                   // sticky: scan-sized frames use the one-CTA solve (accumulate.cu)
  long long est_launches_per_graph = 0;
  cudaGraphExec_t chain_graph = nullptr;   // chained odometry loop: WHILE graph of one scan's solve (accumulate.cu)
  long long chain_graph_key = 0;
  long long chain_launches_per_iter = 0;
  std::vector<int> last_scan_off;  // scan offsets the resident chunk table was built for
  void* odom = nullptr;            // pipelined odometry runner state (odometry.cu)
  void* local_map = nullptr;       // device-side local feature map (localmap.cu)
  void* global_map = nullptr;      // global cube map kept on the device (globalmap.cu)
  void* window = nullptr;          // sliding-window frame slots and solver scratch (window.cu)
  cudaStream_t stream_fe = nullptr;  // feature-extraction stream of the pipelined runner
  bool profile = false;
  cudaEvent_t pev[4] = {nullptr, nullptr, nullptr, nullptr};
  double stage_ms[4] = {0, 0, 0, 0};
  long long stage_n = 0;
  mml::DevBuf frame_cnt;          // int[2] device-side query counts
  mml::DevBuf export_buf;
};

#define MML_CUDA(ctx, call)                                                         \
  do {                                                                              \
    cudaError_t _e = (call);                                                        \
    if (_e != cudaSuccess) {                                                        \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(_e);              \
      return MML_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define MML_CHECK(expr)      \
  do {                       \
    int _r = (expr);         \
    if (_r != MML_OK) return _r; \
  } while (0)

#define MML_LAUNCHED(ctx) ((ctx)->launches++)

// MML_TIMELINE builds: every kernel of the chained loop stamps %globaltimer into a small device array so that the
// last kernel of a scan can print where the scan's time went (debug only; never in bench numbers).
#ifdef MML_TIMELINE
__device__ __forceinline__ void tl_stamp(unsigned long long* tl, int k) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  tl[k] = t;
}
#define MML_TL(tl, k) do { if (tl) mml_tl_stamp_once(tl, k); } while (0)
__device__ __forceinline__ void mml_tl_stamp_once(unsigned long long* tl, int k) { tl_stamp(tl, k); }
#else
#define MML_TL(tl, k) do { } while (0)
#endif

static inline int mml_fail(mml_ctx* ctx, int code, const char* msg) {
  if (ctx) ctx->err = msg;
  return code;
}

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// Estimate loop of one frame on the device (accumulate.cu). after_first_launch (optional) runs on the host right
// after the first outer iteration has been enqueued, i.e. overlapped with it.
int mml_estimate_device(mml_ctx* ctx, const int* cnt_dev, int cap_corner, int cap_surf, const double* exTlb16,
                        double* P3, double* q4, const mml_est_params* prm, double* stats,
                        int (*after_first_launch)(void*) = nullptr, void* hook_arg = nullptr);
