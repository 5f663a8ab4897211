// Device-side global feature map: MAP_MANAGER::MapIncrement with MapMove (src/lio/Map_Manager.cpp:125-281, 288-581),
// the second half of SURVEY.md §8 (f) F1. The map is the reference's grid of 21 x 11 x 21 cubes of 50 m; every
// populated cube keeps its points in HBM in insertion order. One update:
//   snapshot for the matcher     MM.cpp:133-146   the cubes as they are BEFORE the update become the k-NN target of
//                                                 Estimate (the association's spatial hash of the global kinds is
//                                                 rebuilt from them, with the cube centre of that moment)
//   MapMove                      MM.cpp:288-581   the cube grid follows the sensor: cubes change index (no data moves,
//                                                 the host only re-keys its table), cubes pushed over the edge are dropped
//   insert                       MM.cpp:159-175   new world-frame points -> their cubes (reference cube rule), appended
//   per-cube voxel filter        MM.cpp:219-257   touched cubes holding more than 300 points are voxel-filtered (leaf 0.4
//                                                 for both kinds, MM.cpp:56-58); the touched cubes form laserCloud*FromMap
// The per-point work (cube rule, gather, voxel filter, concatenation, spatial-hash build) runs on the device; the host
// keeps the table of populated cubes. Bit-identical to the oracle (tests/test_gpu_parity.py).
#include "common.cuh"
#include <algorithm>
#include <map>
#include <vector>

int mml_voxel_device(mml_ctx* ctx, const float4* pts_d, const int* n_dev, int n_max, float leaf, float4* out_d, int* m_dev);
int mml_map_set_device(mml_ctx* ctx, int kind, const float4* pts_d, int m, const int* cen3, float cell_hint, const float* bbox6 = nullptr);

namespace {

constexpr int kW = 21, kH = 11, kD = 21;      // laserCloudWidth / Height / Depth, Map_Manager.h:117-119
constexpr int kDownsampleOver = 300;          // MM.cpp:222
constexpr float kLeaf = 0.4f;                 // MM.cpp:56-58

// MM.cpp:161-173 (= 583-605): cube index of a world-frame point, 5000 outside the grid
__global__ void __launch_bounds__(256) k_cube_ids(const float4* __restrict__ p, int n, int cen_w, int cen_h, int cen_d,
                                                  int* __restrict__ ids) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float4 q = p[i];
  int cI = int(((double)q.x + 25.0) / 50.0) + cen_d;
  int cJ = int(((double)q.y + 25.0) / 50.0) + cen_w;
  int cK = int(((double)q.z + 25.0) / 50.0) + cen_h;
  if ((double)q.x + 25.0 < 0) cI--;
  if ((double)q.y + 25.0 < 0) cJ--;
  if ((double)q.z + 25.0 < 0) cK--;
  const bool ok = cI >= 0 && cI < kD && cJ >= 0 && cJ < kW && cK >= 0 && cK < kH;
  ids[i] = ok ? cI + kD * cJ + kD * kW * cK : 5000;  // MAP_MANAGER::ToIndex
}
__global__ void __launch_bounds__(256) k_gather_points(const float4* __restrict__ src, const int* __restrict__ idx, int n,
                                                       float4* __restrict__ dst) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

struct Cube {
  mml::DevBuf pts;
  int n = 0;
};
struct GlobalMapDev {
  std::map<int, Cube> cubes[2];            // linear cube index -> points (insertion order)
  int cen[3] = {10, 5, 10};                // CenWidth, CenHeight, CenDepth
  int cen_last[3] = {10, 5, 10};
  mml::DevBuf match[2], from_map[2], stage, ids, lists, vox_out, cnt;
  int match_n[2] = {0, 0}, from_n[2] = {0, 0};
};
GlobalMapDev* get(mml_ctx* c) {
  if (!c->global_map) c->global_map = new GlobalMapDev();
  return static_cast<GlobalMapDev*>(c->global_map);
}
void release_all(GlobalMapDev* G) {
  for (int k = 0; k < 2; k++) {
    for (auto& kv : G->cubes[k]) kv.second.pts.release();
    G->cubes[k].clear();
    G->match[k].release(); G->from_map[k].release();
  }
  G->stage.release(); G->ids.release(); G->lists.release(); G->vox_out.release(); G->cnt.release();
}

// room for `need` points, contents kept
int grow_keep(mml_ctx* c, Cube& q, int need) {
  if (sizeof(float4) * (size_t)need <= q.pts.cap) return MML_OK;
  mml::DevBuf bigger;
  MML_CUDA(c, bigger.reserve(sizeof(float4) * (size_t)(need + need / 2 + 64)));
  if (q.n) MML_CUDA(c, cudaMemcpyAsync(bigger.p, q.pts.p, sizeof(float4) * (size_t)q.n, cudaMemcpyDeviceToDevice, c->stream));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  q.pts.release();
  q.pts = bigger;
  return MML_OK;
}

// all cubes of one kind in cube-index order into `dst`
int concat_cubes(mml_ctx* c, std::map<int, Cube>& cubes, const std::vector<int>* only, mml::DevBuf& dst, int* n_out) {
  long long total = 0;
  if (only) for (int ci : *only) total += cubes[ci].n;
  else for (auto& kv : cubes) total += kv.second.n;
  if (total > 0x7fffffffLL / 32) return mml_fail(c, MML_ERR_CAPACITY, "global map too large");
  MML_CUDA(c, dst.reserve(sizeof(float4) * (size_t)(total > 0 ? total : 1)));
  size_t at = 0;
  auto put = [&](Cube& q) -> int {
    if (!q.n) return MML_OK;
    MML_CUDA(c, cudaMemcpyAsync(dst.as<float4>() + at, q.pts.p, sizeof(float4) * (size_t)q.n, cudaMemcpyDeviceToDevice, c->stream));
    at += (size_t)q.n;
    return MML_OK;
  };
  if (only) { for (int ci : *only) MML_CHECK(put(cubes[ci])); }
  else { for (auto& kv : cubes) MML_CHECK(put(kv.second)); }
  *n_out = (int)total;
  return MML_OK;
}

// one pass of a MapMove while-loop body (MM.cpp:307-579): every cube moves by `step` along `axis` (0 = I depth,
// 1 = J width, 2 = K height); what leaves the grid is dropped
void shift_cubes(GlobalMapDev* G, int axis, int step) {
  const int size[3] = {kD, kW, kH};
  for (int kind = 0; kind < 2; kind++) {
    std::map<int, Cube> moved;
    for (auto& kv : G->cubes[kind]) {
      int ijk[3] = {kv.first % kD, (kv.first / kD) % kW, kv.first / (kD * kW)};
      ijk[axis] += step;
      if (ijk[axis] >= 0 && ijk[axis] < size[axis] && kv.second.n > 0) moved[ijk[0] + kD * ijk[1] + kD * kW * ijk[2]] = kv.second;
      else kv.second.pts.release();
    }
    G->cubes[kind].swap(moved);
  }
}

void map_move(GlobalMapDev* G, const double* T_wl16) {
  const double t[3] = {T_wl16[3], T_wl16[7], T_wl16[11]};
  // centre cube of the sensor, MM.cpp:299-305 (I from x with CenDepth, J from y with CenWidth, K from z with CenHeight)
  int cen_axis[3] = {G->cen[2], G->cen[0], G->cen[1]};
  int cc[3];
  for (int a = 0; a < 3; a++) {
    cc[a] = int((t[a] + 25.0) / 50.0) + cen_axis[a];
    if (t[a] + 25.0 < 0) cc[a]--;
  }
  const int size[3] = {kD, kW, kH};
  for (int a = 0; a < 3; a++) {
    while (cc[a] < 8) { shift_cubes(G, a, +1); cc[a]++; cen_axis[a]++; }
    while (cc[a] >= size[a] - 8) { shift_cubes(G, a, -1); cc[a]--; cen_axis[a]--; }
  }
  G->cen[0] = cen_axis[1]; G->cen[1] = cen_axis[2]; G->cen[2] = cen_axis[0];
}

}  // namespace

void mml_global_map_destroy(mml_ctx* c) {
  if (!c->global_map) return;
  GlobalMapDev* G = static_cast<GlobalMapDev*>(c->global_map);
  release_all(G);
  delete G;
  c->global_map = nullptr;
}

extern "C" {

int mml_global_map_reset(mml_ctx* c) {
  if (!c) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  mml_global_map_destroy(c);
  int rc = mml_map_set_device(c, MML_MAP_CORNER_GLOBAL, nullptr, 0, nullptr, 0.f);
  const int rc2 = mml_map_set_device(c, MML_MAP_SURF_GLOBAL, nullptr, 0, nullptr, 0.f);
  return rc != MML_OK ? rc : rc2;
}

// One MAP_MANAGER::MapIncrement. corner_w / surf_w: the frame's feature clouds in the WORLD frame (host, MM.cpp:159 copies
// the stack as it is); T_wl16 (may be NULL: no MapMove) the LiDAR pose that drives MapMove. On return the global map kinds
// of the association (0 / 1) hold the cubes as they were BEFORE this update. n_from_map2 (may be NULL): sizes of
// laserCloudCornerFromMap / SurfFromMap (the touched cubes).
int mml_global_map_push(mml_ctx* c, const float* corner_w, int n_corner, const float* surf_w, int n_surf, const double* T_wl16,
                        int* n_from_map2) {
  if (!c || n_corner < 0 || n_surf < 0 || (n_corner && !corner_w) || (n_surf && !surf_w)) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  GlobalMapDev* G = get(c);
  // ---- snapshot for the matcher (MM.cpp:133-146)
  for (int kind = 0; kind < 2; kind++) {
    MML_CHECK(concat_cubes(c, G->cubes[kind], nullptr, G->match[kind], &G->match_n[kind]));
    MML_CHECK(mml_map_set_device(c, kind == 0 ? MML_MAP_CORNER_GLOBAL : MML_MAP_SURF_GLOBAL, G->match[kind].as<float4>(), G->match_n[kind],
                                 G->cen, 0.f));
  }
  for (int a = 0; a < 3; a++) G->cen_last[a] = G->cen[a];
  if (T_wl16) map_move(G, T_wl16);  // MM.cpp:149
  MML_CUDA(c, G->cnt.reserve(64));
  for (int kind = 0; kind < 2; kind++) {
    const float* src = kind == 0 ? corner_w : surf_w;
    const int n = kind == 0 ? n_corner : n_surf;
    std::vector<int> touched;
    if (n > 0) {
      MML_CUDA(c, G->stage.reserve(sizeof(float4) * (size_t)n));
      MML_CUDA(c, G->ids.reserve(sizeof(int) * (size_t)n));
      MML_CUDA(c, G->lists.reserve(sizeof(int) * (size_t)n));
      MML_CUDA(c, cudaMemcpyAsync(G->stage.p, src, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
      k_cube_ids<<<div_up(n, 256), 256, 0, st>>>(G->stage.as<float4>(), n, G->cen[0], G->cen[1], G->cen[2], G->ids.as<int>());
      MML_LAUNCHED(c);
      std::vector<int> ids((size_t)n);
      MML_CUDA(c, cudaMemcpyAsync(ids.data(), G->ids.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
      MML_CUDA(c, cudaStreamSynchronize(st));
      // per touched cube: the indices of its new points, in input order (push_back, MM.cpp:175)
      std::map<int, std::vector<int>> lists;
      for (int i = 0; i < n; i++) if (ids[(size_t)i] != 5000) lists[ids[(size_t)i]].push_back(i);
      std::vector<int> flat;
      flat.reserve((size_t)n);
      for (auto& kv : lists) flat.insert(flat.end(), kv.second.begin(), kv.second.end());
      if (!flat.empty()) MML_CUDA(c, cudaMemcpyAsync(G->lists.p, flat.data(), sizeof(int) * flat.size(), cudaMemcpyHostToDevice, st));
      size_t at = 0;
      for (auto& kv : lists) {
        Cube& q = G->cubes[kind][kv.first];
        const int add = (int)kv.second.size();
        MML_CHECK(grow_keep(c, q, q.n + add));
        k_gather_points<<<div_up(add, 256), 256, 0, st>>>(G->stage.as<float4>(), G->lists.as<int>() + at, add, q.pts.as<float4>() + q.n);
        MML_LAUNCHED(c);
        q.n += add;
        at += (size_t)add;
        touched.push_back(kv.first);  // std::map iterates in cube-index order (MM.cpp:219)
      }
      MML_CUDA(c, cudaStreamSynchronize(st));  // `flat` leaves scope
    }
    // per-cube voxel filter of the touched cubes holding more than 300 points (MM.cpp:219-257)
    for (int ci : touched) {
      Cube& q = G->cubes[kind][ci];
      if (q.n <= kDownsampleOver) continue;
      MML_CUDA(c, G->vox_out.reserve(sizeof(float4) * (size_t)q.n));
      int* cnt = G->cnt.as<int>();
      MML_CUDA(c, cudaMemcpyAsync(cnt, &q.n, sizeof(int), cudaMemcpyHostToDevice, st));
      MML_CHECK(mml_voxel_device(c, q.pts.as<float4>(), cnt, q.n, kLeaf, G->vox_out.as<float4>(), cnt + 1));
      int m = 0;
      MML_CUDA(c, cudaMemcpyAsync(&m, cnt + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
      MML_CUDA(c, cudaStreamSynchronize(st));
      if (m > 0) MML_CUDA(c, cudaMemcpyAsync(q.pts.p, G->vox_out.p, sizeof(float4) * (size_t)m, cudaMemcpyDeviceToDevice, st));
      q.n = m;
    }
    // laserCloud*FromMap: the cubes this update touched (MM.cpp:215-217, 233)
    MML_CHECK(concat_cubes(c, G->cubes[kind], &touched, G->from_map[kind], &G->from_n[kind]));
    if (n_from_map2) n_from_map2[kind] = G->from_n[kind];
  }
  MML_CUDA(c, cudaStreamSynchronize(st));
  return MML_OK;
}

// which: 0 = all cubes now (cube-index order), 1 = the snapshot the matcher sees, 2 = laserCloud*FromMap of the last update.
// cen3_out (may be NULL): the cube centre that belongs to it (CenWidth, CenHeight, CenDepth).
int mml_global_map_get(mml_ctx* c, int kind, int which, float* out_xyzi, int cap, int* n_out, int* cen3_out) {
  if (!c || kind < 0 || kind > 1 || which < 0 || which > 2 || !n_out) return MML_ERR_INVALID;
  cudaSetDevice(c->device);
  GlobalMapDev* G = get(c);
  const mml::DevBuf* src = nullptr;
  mml::DevBuf tmp;
  int n = 0;
  if (which == 0) {
    MML_CHECK(concat_cubes(c, G->cubes[kind], nullptr, tmp, &n));
    src = &tmp;
  } else if (which == 1) { src = &G->match[kind]; n = G->match_n[kind]; }
  else { src = &G->from_map[kind]; n = G->from_n[kind]; }
  *n_out = n;
  if (cen3_out) for (int a = 0; a < 3; a++) cen3_out[a] = which == 1 ? G->cen_last[a] : G->cen[a];
  int rc = MML_OK;
  if (out_xyzi && n) {
    if (cap < n) rc = mml_fail(c, MML_ERR_CAPACITY, "output buffer too small for the global map");
    else {
      cudaError_t e = cudaMemcpyAsync(out_xyzi, src->p, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess) { c->err = cudaGetErrorString(e); rc = MML_ERR_CUDA; }
    }
  } else {
    cudaStreamSynchronize(c->stream);
  }
  tmp.release();
  return rc;
}

}  // extern "C"
