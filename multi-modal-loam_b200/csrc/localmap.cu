// Device-side local feature map: Estimator::MapIncrementLocal (src/lio/Estimator.cpp:1585-1643), the first of the
// "next" rows of SURVEY.md §8 (f) F1. The last 50 frames' corner / surf clouds are kept in the world frame in HBM;
// an update transforms the new frame (MAP_MANAGER::pointAssociateToMap, MM.cpp:75-89), concatenates the previous
// filtered map and the 50 ring entries (EST.cpp:1620-1624: the previous result is NOT cleared first), runs the voxel
// filter (EST.cpp:1630-1635, pcl::VoxelGrid restated in geometry.cu) and rebuilds the spatial hash of the local map
// kind - the map the association searches never leaves the device and is never re-uploaded.
// Checked against the CPU restatement in tests/test_gpu_parity.py::test_local_map_increment_matches_oracle.
#include "common.cuh"

int mml_voxel_device(mml_ctx* ctx, const float4* pts_d, const int* n_dev, int n_max, float leaf, float4* out_d, int* m_dev);
int mml_map_set_device(mml_ctx* ctx, int kind, const float4* pts_d, int m, const int* cen3, float cell_hint, const float* bbox6 = nullptr);

namespace {

constexpr int kLocalWindow = 50;  // localMapWindowSize, include/Estimator/Estimator.h:326

struct Pose16 { double T[16]; };

// p_w = R p + t in float64, summed left to right, rounded to float32 (MM.cpp:75-89); intensity kept
__global__ void __launch_bounds__(256) k_point_to_map(const float4* __restrict__ in, int n, Pose16 P, float4* __restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float4 p = in[i];
  const double x = (double)p.x, y = (double)p.y, z = (double)p.z;
  float4 o;
  o.x = (float)(((P.T[0] * x + P.T[1] * y) + P.T[2] * z) + P.T[3]);
  o.y = (float)(((P.T[4] * x + P.T[5] * y) + P.T[6] * z) + P.T[7]);
  o.z = (float)(((P.T[8] * x + P.T[9] * y) + P.T[10] * z) + P.T[11]);
  o.w = p.w;
  out[i] = o;
}

// The filtered map's input is the previous map (unless cleared) followed by the 50 ring entries in slot order: one
// gather launch instead of up to 51 copy nodes per kind (the copies' launch overhead was most of an update's time).
struct RingSrc {
  const float4* p[kLocalWindow + 1];
  int off[kLocalWindow + 2];  // off[j] = first output index of source j, off[n] = total
  int n;
};
__global__ void __launch_bounds__(256) k_ring_concat(RingSrc R, float4* __restrict__ dst) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= R.off[R.n]) return;
  int lo = 0, hi = R.n;  // the source j with off[j] <= i < off[j + 1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (R.off[mid] <= i) lo = mid;
    else hi = mid;
  }
  dst[i] = R.p[lo][i - R.off[lo]];
}

struct LocalMapDev {
  mml::DevBuf ring[2][kLocalWindow];
  int ring_n[2][kLocalWindow];
  mml::DevBuf from_local[2];
  int from_n[2] = {0, 0};
  mml::DevBuf stage, concat, cnt;
  mml::DevBuf pending[2], result[2];  // an update is prepared here and committed only when both kinds succeeded
  mml::PinBuf pin;                    // sizes and bounding boxes read back once per update
  long long id = 0;  // localMapID
  LocalMapDev() { memset(ring_n, 0, sizeof(ring_n)); }
};

LocalMapDev* get(mml_ctx* c) {
  if (!c->local_map) c->local_map = new LocalMapDev();
  return static_cast<LocalMapDev*>(c->local_map);
}

// the local maps of the association become invalid (nothing is matched against them) until the next update
int drop_local_maps(mml_ctx* c) {
  int rc = mml_map_set_device(c, MML_MAP_CORNER_LOCAL, nullptr, 0, nullptr, 0.f);
  const int rc2 = mml_map_set_device(c, MML_MAP_SURF_LOCAL, nullptr, 0, nullptr, 0.f);
  return rc != MML_OK ? rc : rc2;
}

// One MapIncrementLocal (EST.cpp:1585-1643). src_device: the clouds are device pointers. clear_first: the caller
// cleared laserCloud*FromLocal beforehand (EST.cpp:1085-1087, 1127-1129: what EstimateLidarPose does), so the new
// map is the filtered concatenation of the ring alone.
// Transactional: the new ring entry and the new filtered maps of BOTH kinds are prepared in side buffers; ring,
// sizes, localMapID and the association's search structures change only after every fallible step has succeeded.
int local_map_push_impl(mml_ctx* c, const void* corner, int n_corner, const void* surf, int n_surf, bool src_device,
                        const double* T_wl16, float leaf_corner, float leaf_surf, bool clear_first, int* n_corner_map,
                        int* n_surf_map) {
  cudaStream_t st = c->stream;
  LocalMapDev* L = get(c);
  const int slot = (int)(L->id % kLocalWindow);  // EST.cpp:1597
  Pose16 P;
  for (int i = 0; i < 16; i++) P.T[i] = T_wl16[i];
  MML_CUDA(c, L->cnt.reserve(64));
  MML_CUDA(c, L->pin.reserve(sizeof(int) * 16));
  int* hp = L->pin.as<int>();
  bool have[2] = {false, false};
  int m_new[2] = {0, 0};
  for (int k = 0; k < 2; k++) {
    const void* src = k == 0 ? corner : surf;
    const int n = k == 0 ? n_corner : n_surf;
    const float leaf = k == 0 ? leaf_corner : leaf_surf;
    // new frame -> world frame -> pending ring entry (EST.cpp:1600-1612)
    if (n > 0) {
      MML_CUDA(c, L->pending[k].reserve(sizeof(float4) * (size_t)n));
      const float4* in = static_cast<const float4*>(src);
      if (!src_device) {
        MML_CUDA(c, L->stage.reserve(sizeof(float4) * (size_t)n));
        MML_CUDA(c, cudaMemcpyAsync(L->stage.p, src, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
        in = L->stage.as<float4>();
      }
      k_point_to_map<<<div_up(n, 256), 256, 0, st>>>(in, n, P, L->pending[k].as<float4>());
      MML_LAUNCHED(c);
    }
    // previous filtered map (unless cleared) + the 50 ring entries, in that order (EST.cpp:1620-1624)
    const int keep = clear_first ? 0 : L->from_n[k];
    long long total = keep;
    for (int i = 0; i < kLocalWindow; i++) total += (i == slot) ? n : L->ring_n[k][i];
    if (total > 0x7fffffffLL / 32) return mml_fail(c, MML_ERR_CAPACITY, "local map too large");
    if (total > 0) {
      MML_CUDA(c, L->concat.reserve(sizeof(float4) * (size_t)total));
      MML_CUDA(c, L->result[k].reserve(sizeof(float4) * (size_t)total));
      float4* dst = L->concat.as<float4>();
      RingSrc R;
      R.n = 0;
      int at = 0;
      auto add = [&](const void* from, int cnt_pts) {
        if (!cnt_pts) return;
        R.p[R.n] = static_cast<const float4*>(from);
        R.off[R.n] = at;
        R.n++;
        at += cnt_pts;
      };
      if (keep) add(L->from_local[k].p, keep);
      for (int i = 0; i < kLocalWindow; i++) add((i == slot) ? L->pending[k].p : L->ring[k][i].p, (i == slot) ? n : L->ring_n[k][i]);
      R.off[R.n] = at;
      k_ring_concat<<<div_up(at, 256), 256, 0, st>>>(R, dst);
      MML_LAUNCHED(c);
      // voxel filter (EST.cpp:1630-1635)
      int* cnt = L->cnt.as<int>() + 4 * k;
      const int tot = (int)total;
      MML_CUDA(c, cudaMemcpyAsync(cnt, &tot, sizeof(int), cudaMemcpyHostToDevice, st));
      MML_CHECK(mml_voxel_device(c, dst, cnt, tot, leaf, L->result[k].as<float4>(), cnt + 1));
      // the size of the filtered map and the filter's bounding box of its input (the centroids lie inside it) come
      // back together after both kinds have been enqueued: one host wait per update
      MML_CUDA(c, cudaMemcpyAsync(hp + 8 * k, cnt + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
      MML_CUDA(c, cudaMemcpyAsync(hp + 8 * k + 1, c->vox_bbox.p, 6 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
      have[k] = true;
    }
  }
  MML_CUDA(c, cudaStreamSynchronize(st));
  float bbox[2][6];
  for (int k = 0; k < 2; k++) {
    m_new[k] = have[k] ? hp[8 * k] : 0;
    for (int a = 0; a < 6; a++) {  // order-preserving unsigned encoding of float (geometry.cu f2ord)
      const unsigned u = (unsigned)hp[8 * k + 1 + a];
      const unsigned b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
      memcpy(&bbox[k][a], &b, 4);
    }
  }
  // ---- commit
  for (int k = 0; k < 2; k++) {
    std::swap(L->ring[k][slot], L->pending[k]);
    L->ring_n[k][slot] = k == 0 ? n_corner : n_surf;
    std::swap(L->from_local[k], L->result[k]);
    L->from_n[k] = m_new[k];
  }
  L->id++;  // EST.cpp:1640
  // the association's search structures (replace the kd-tree rebuilds at EST.cpp:1159-1167)
  for (int k = 0; k < 2; k++) {
    const int rc = mml_map_set_device(c, k == 0 ? MML_MAP_CORNER_LOCAL : MML_MAP_SURF_LOCAL, L->from_local[k].as<float4>(), L->from_n[k], nullptr, 0.f,
                                      have[k] ? bbox[k] : nullptr);
    if (rc != MML_OK) { drop_local_maps(c); return rc; }  // never leave one kind new and the other stale
  }
  if (n_corner_map) *n_corner_map = m_new[0];
  if (n_surf_map) *n_surf_map = m_new[1];
  return MML_OK;  // what follows on the context's stream is ordered behind the update
}

}  // namespace

void mml_local_map_destroy(mml_ctx* c) {
  if (!c->local_map) return;
  LocalMapDev* L = static_cast<LocalMapDev*>(c->local_map);
  for (int k = 0; k < 2; k++) {
    for (int i = 0; i < kLocalWindow; i++) L->ring[k][i].release();
    L->from_local[k].release(); L->pending[k].release(); L->result[k].release();
  }
  L->stage.release(); L->concat.release(); L->cnt.release(); L->pin.release();
  delete L;
  c->local_map = nullptr;
}

extern "C" {

// forget the ring and the filtered maps; the association's local maps become invalid with them
int mml_local_map_reset(mml_ctx* c) {
  if (!c) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  mml_local_map_destroy(c);
  return drop_local_maps(c);
}

// One MapIncrementLocal. corner / surf: the frame's (down-sampled) feature clouds in the LiDAR frame, T_wl16 the
// LiDAR-to-world transform (transformTobeMapped). On return the local corner / surf maps (kinds 2 / 3) are the new
// filtered clouds; n_corner_map / n_surf_map (may be NULL) receive their sizes.
int mml_local_map_push(mml_ctx* c, const float* corner_xyzi, int n_corner, const float* surf_xyzi, int n_surf, const double* T_wl16,
                       float leaf_corner, float leaf_surf, int* n_corner_map, int* n_surf_map) {
  if (!c || !T_wl16 || n_corner < 0 || n_surf < 0 || (n_corner && !corner_xyzi) || (n_surf && !surf_xyzi) || !(leaf_corner > 0.f) ||
      !(leaf_surf > 0.f))
    return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  return local_map_push_impl(c, corner_xyzi, n_corner, surf_xyzi, n_surf, false, T_wl16, leaf_corner, leaf_surf, false, n_corner_map, n_surf_map);
}

// The same with the clouds resident in HBM (float4 xyzi) and the caller's clear of laserCloud*FromLocal
// (clear_first != 0: the local map becomes the filtered concatenation of the ring, as in EstimateLidarPose).
int mml_local_map_push_dev(mml_ctx* c, const void* corner_dev, int n_corner, const void* surf_dev, int n_surf, const double* T_wl16,
                           float leaf_corner, float leaf_surf, int clear_first, int* n_corner_map, int* n_surf_map) {
  if (!c || !T_wl16 || n_corner < 0 || n_surf < 0 || (n_corner && !corner_dev) || (n_surf && !surf_dev) || !(leaf_corner > 0.f) ||
      !(leaf_surf > 0.f))
    return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  return local_map_push_impl(c, corner_dev, n_corner, surf_dev, n_surf, true, T_wl16, leaf_corner, leaf_surf, clear_first != 0, n_corner_map, n_surf_map);
}

// Place a world-frame cloud into ring entry `slot` of one kind (0 corner / 1 surf): the state after earlier frames
// had been pushed (tests and benchmarks start from a map instead of building it up scan by scan). The filtered maps
// change with the next push.
int mml_local_map_seed(mml_ctx* c, int kind, int slot, const float* xyzi_world, int n) {
  if (!c || kind < 0 || kind > 1 || slot < 0 || slot >= kLocalWindow || n < 0 || (n && !xyzi_world)) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  LocalMapDev* L = get(c);
  if (n > 0) {
    MML_CUDA(c, L->ring[kind][slot].reserve(sizeof(float4) * (size_t)n));
    MML_CUDA(c, cudaMemcpyAsync(L->ring[kind][slot].p, xyzi_world, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    MML_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  L->ring_n[kind][slot] = n;
  return MML_OK;
}

// Copy of the current local map of one kind (0 corner / 1 surf): out has room for cap points; *n_out = its size.
int mml_local_map_get(mml_ctx* c, int kind, float* out_xyzi, int cap, int* n_out) {
  if (!c || kind < 0 || kind > 1 || !n_out) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  LocalMapDev* L = get(c);
  *n_out = L->from_n[kind];
  if (!out_xyzi) return MML_OK;
  if (cap < L->from_n[kind]) return mml_fail(c, MML_ERR_CAPACITY, "output buffer too small for the local map");
  if (L->from_n[kind]) {
    MML_CUDA(c, cudaMemcpyAsync(out_xyzi, L->from_local[kind].p, sizeof(float4) * (size_t)L->from_n[kind], cudaMemcpyDeviceToHost, c->stream));
    MML_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return MML_OK;
}

}  // extern "C"
