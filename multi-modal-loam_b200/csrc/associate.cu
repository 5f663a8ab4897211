// Spatial-hash feature map + correspondence search + local geometry fit on the device.
//
// Replaces, for the hot path:
//   pcl::KdTreeFLANN::setInputCloud / nearestKSearch   EST.cpp:1159-1179, 199, 284, 630, 704
//   MAP_MANAGER::pointAssociateToMap + cube rule       MM.cpp:75-89, 583-629
//   Estimator::processPointToLine                      EST.cpp:148-365
//   Estimator::processPointToPlanVec                   EST.cpp:573-777 (+ 536-565 moments)
//
// Map layout in HBM: points are counting-sorted into a dense grid of cubic cells
// (float4 xyz + original index in .w, cell_start[ncell+1]); x is the fastest cell axis so a
// row of neighbouring cells is ONE contiguous point range. Global maps keep the reference's
// "own 50 m cube only" rule exactly: a point is binned by the reference's cube formula and
// its cell is clamped into that cube's block of cells; a query only walks cells of its cube.
//
// Search: exact 5-NN inside radius sqrt(thres_dist) by growing Chebyshev shells; shell r is
// final once the 5th best squared distance <= (r*cell - margin)^2. Distances are
// ((dx*dx)+dy*dy)+dz*dz in float32 without FMA (FLANN L2_Simple<float>), ties broken by the
// lower original index, so neighbour sets equal the CPU oracle's bit for bit.
#include "common.cuh"
#include "sort.cuh"
#include "smallmath.cuh"
#include <math.h>
#include <stdlib.h>

namespace mml {

struct GridDev {
  const float4* pts;
  const int* cell_start;
  const int* cube_count;  // global kinds only
  const float4* pts2;     // coarse level (cells `coarse` times larger), null when absent
  const int* cell_start2;
  int dim2[3];
  int coarse;
  double org[3];          // lower corner of cell (0,0,0)
  double inv_cell;
  float cell;
  int dim[3];
  int m;
  int global;             // 1: cube rule
  int k_per_cube;         // cells per cube edge (global)
  int cube_lo[3];         // lowest cube index (I,J,K) covered by the grid
  int cen[3];             // (CenWidth, CenHeight, CenDepth)
  int min_cube_pts;       // > this many points in the cube (100 corner / 50 surf)  EST.cpp:198,627
  int min_local_pts;      // > 20 for local maps                                    EST.cpp:283,702
  int valid;
};

// reference cube rule, MM.cpp:583-605. Returns false when outside the 21x11x21 grid.
__host__ __device__ inline bool cube_of(float x, float y, float z, const int* cen, int& cI, int& cJ, int& cK) {
  cI = int(((double)x + 25.0) / 50.0) + cen[2];
  cJ = int(((double)y + 25.0) / 50.0) + cen[0];
  cK = int(((double)z + 25.0) / 50.0) + cen[1];
  if ((double)x + 25.0 < 0) cI--;
  if ((double)y + 25.0 < 0) cJ--;
  if ((double)z + 25.0 < 0) cK--;
  return cI >= 0 && cI < kCubeD && cJ >= 0 && cJ < kCubeW && cK >= 0 && cK < kCubeH;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// cell coordinates of a point + the inclusive cell range it may search / be binned in
__device__ __forceinline__ bool locate(const GridDev& G, float x, float y, float z, int c[3], int lo[3], int hi[3],
                                       int& cube_id) {
  cube_id = -1;
  if (G.global) {
    int cI, cJ, cK;
    if (!cube_of(x, y, z, G.cen, cI, cJ, cK)) return false;
    cube_id = cI + kCubeD * cJ + kCubeD * kCubeW * cK;
    const int rel[3] = {cI - G.cube_lo[0], cJ - G.cube_lo[1], cK - G.cube_lo[2]};
    for (int a = 0; a < 3; a++) {
      lo[a] = rel[a] * G.k_per_cube;
      hi[a] = lo[a] + G.k_per_cube - 1;
      if (lo[a] < 0 || hi[a] >= G.dim[a]) return false;  // cube not covered by the map
    }
  } else {
    for (int a = 0; a < 3; a++) { lo[a] = 0; hi[a] = G.dim[a] - 1; }
  }
  const double p[3] = {(double)x, (double)y, (double)z};
  for (int a = 0; a < 3; a++) {
    double f = floor((p[a] - G.org[a]) * G.inv_cell);
    int ci = f < -1e9 ? -1000000000 : (f > 1e9 ? 1000000000 : (int)f);
    c[a] = clampi(ci, lo[a], hi[a]);
  }
  return true;
}

// ---------------------------------------------------------------- map build
__global__ void __launch_bounds__(256) k_map_count(const float4* __restrict__ pts, int m, GridDev G, int* __restrict__ cell_cnt,
                                                   int* __restrict__ cell_of, int* __restrict__ cube_cnt) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= m) return;
  const float4 p = pts[i];
  int c[3], lo[3], hi[3], cube;
  if (!locate(G, p.x, p.y, p.z, c, lo, hi, cube)) {  // outside the cube grid: dropped like MM.cpp:168-175
    cell_of[i] = -1;
    return;
  }
  const int cell = (c[2] * G.dim[1] + c[1]) * G.dim[0] + c[0];
  cell_of[i] = cell;
  atomicAdd(&cell_cnt[cell], 1);
  if (G.global) {
    // a map has a handful of populated cubes and millions of points: one atomic per (warp, cube) instead of one per point
    const unsigned peers = __match_any_sync(__activemask(), cube);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cube_cnt[cube], __popc(peers));
  }
}

__global__ void __launch_bounds__(256) k_map_scatter(const float4* __restrict__ pts, int m, const int* __restrict__ cell_of,
                                                     const int* __restrict__ cell_start, int* __restrict__ cell_fill,
                                                     float4* __restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= m) return;
  const int cell = cell_of[i];
  if (cell < 0) return;
  const int pos = cell_start[cell] + atomicAdd(&cell_fill[cell], 1);
  float4 p = pts[i];
  p.w = __int_as_float(i);
  out[pos] = p;
}

// coarse level: cell of every point from its fine cell (integer division keeps cube blocks aligned)
__global__ void __launch_bounds__(256) k_map_coarse_count(const int* __restrict__ cell_of, int m, int dx, int dy, int f,
                                                          int dx2, int dy2, int* __restrict__ cell_of2,
                                                          int* __restrict__ cell_cnt2) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= m) return;
  const int cell = cell_of[i];
  if (cell < 0) { cell_of2[i] = -1; return; }
  const int x = cell % dx, y = (cell / dx) % dy, z = cell / (dx * dy);
  const int c2 = ((z / f) * dy2 + (y / f)) * dx2 + (x / f);
  cell_of2[i] = c2;
  atomicAdd(&cell_cnt2[c2], 1);
}

// occupied-cell count at a trial resolution (to pick the final cell size)
__global__ void __launch_bounds__(256) k_count_nonzero(const int* __restrict__ a, long long n, unsigned long long* out) {
  long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  int v = 0;
  for (; i < n; i += (long long)gridDim.x * 256) v += a[i] != 0;
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, (unsigned long long)v);
}

// distance (in cells, conservative) from a point at fraction f of its cell to the cell at offset o along one axis
__device__ __forceinline__ float axis_gap(float f, int o) {
  const float g = o > 0 ? (float)o - f : (o < 0 ? f - (float)(o + 1) : 0.f);
  return fmaxf(g - 1e-3f, 0.f);
}

// ---------------------------------------------------------------- k-NN
struct Knn5 {
  float d[5];
  int id[5];
  int loc[5];  // position in the cell-sorted array
  int cnt;
};

__device__ __forceinline__ void knn_init(Knn5& r) {
#pragma unroll
  for (int k = 0; k < 5; k++) { r.d[k] = INFINITY; r.id[k] = 0x7fffffff; r.loc[k] = -1; }
  r.cnt = 0;
}
__device__ __forceinline__ bool knn_better(float d, int id, float d2, int id2) { return d < d2 || (d == d2 && id < id2); }
__device__ __forceinline__ void knn_push(Knn5& r, float d, int id, int loc) {
  if (!knn_better(d, id, r.d[4], r.id[4])) return;
  // insertion into the sorted 5-list (fully unrolled, registers only)
#pragma unroll
  for (int k = 4; k >= 0; k--) {
    const bool up = (k > 0) && knn_better(d, id, r.d[k - 1], r.id[k - 1]);
    if (up) {
      r.d[k] = r.d[k - 1]; r.id[k] = r.id[k - 1]; r.loc[k] = r.loc[k - 1];
    } else {
      r.d[k] = d; r.id[k] = id; r.loc[k] = loc;
      break;
    }
  }
  if (r.cnt < 5) r.cnt++;
}

__device__ __forceinline__ void scan_range(const GridDev& G, int c0, int c1, float qx, float qy, float qz, Knn5& r) {
  const int s = __ldg(G.cell_start + c0), e = __ldg(G.cell_start + c1 + 1);
  for (int k = s; k < e; k++) {
    const float4 p = __ldg(G.pts + k);
    const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
    const float d = (dx * dx + dy * dy) + dz * dz;
    knn_push(r, d, __float_as_int(p.w), k);
  }
}

// exact 5-NN within squared radius `thres`; true iff 5 neighbours found and d2[4] < thres
__device__ bool knn5_grid(const GridDev& G, float qx, float qy, float qz, float thres, Knn5& r) {
  knn_init(r);
  int c[3], lo[3], hi[3], cube;
  if (!locate(G, qx, qy, qz, c, lo, hi, cube)) return false;
  if (G.global) {
    if (!(__ldg(G.cube_count + cube) > G.min_cube_pts)) return false;
  } else {
    if (!(G.m > G.min_local_pts)) return false;
  }
  const float cellf = G.cell;
  const int rmax = (int)ceilf(sqrtf(thres) / cellf) + 1;
  for (int rr = 1; rr <= rmax; rr++) {
    // shell rr (rr == 1 also covers the centre cell)
    for (int dz = -rr; dz <= rr; dz++) {
      const int z = c[2] + dz;
      if (z < lo[2] || z > hi[2]) continue;
      for (int dy = -rr; dy <= rr; dy++) {
        const int y = c[1] + dy;
        if (y < lo[1] || y > hi[1]) continue;
        const int row = (z * G.dim[1] + y) * G.dim[0];
        const bool full = (rr == 1) || dz == -rr || dz == rr || dy == -rr || dy == rr;
        if (full) {
          const int x0 = max(c[0] - rr, lo[0]), x1 = min(c[0] + rr, hi[0]);
          if (x0 <= x1) scan_range(G, row + x0, row + x1, qx, qy, qz, r);
        } else {
          const int xa = c[0] - rr, xb = c[0] + rr;
          if (xa >= lo[0]) scan_range(G, row + xa, row + xa, qx, qy, qz, r);
          if (xb <= hi[0]) scan_range(G, row + xb, row + xb, qx, qy, qz, r);
        }
      }
    }
    // nothing left inside the searchable block of cells
    if (c[0] - rr < lo[0] && c[0] + rr > hi[0] && c[1] - rr < lo[1] && c[1] + rr > hi[1] && c[2] - rr < lo[2] &&
        c[2] + rr > hi[2])
      break;
    const float reach = fmaxf((float)rr * cellf - 1e-3f, 0.f);
    const float reach2 = reach * reach;
    if (r.cnt == 5 && r.d[4] <= reach2) break;
    if (reach2 >= thres) break;
  }
  return r.cnt == 5 && r.d[4] < thres;
}

// ---- thread-per-query search (map-sized query sets): packed keys + box pruning
// The 5-list is kept as 64-bit keys (distance bits << 32 | original index): squared distances are >= +0, so their bit
// patterns order like unsigned integers and one 64-bit compare implements "closer, ties to the lower index". The
// insertion is branch-free (new_i = p_{i-1} ? old_{i-1} : p_i ? x : old_i with p_i = x < old_i).
struct KnnP {
  unsigned long long key[5];
  int loc[5];
};
__device__ __forceinline__ void knnp_init(KnnP& r) {
#pragma unroll
  for (int k = 0; k < 5; k++) { r.key[k] = 0x7f800000ffffffffull; r.loc[k] = -1; }  // +inf distance
}
__device__ __forceinline__ void knnp_push(KnnP& r, float d, int id, int loc) {
  const unsigned long long x = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)id;
  if (x >= r.key[4]) return;
  const bool p0 = x < r.key[0], p1 = x < r.key[1], p2 = x < r.key[2], p3 = x < r.key[3];
  r.key[4] = p3 ? r.key[3] : x;            r.loc[4] = p3 ? r.loc[3] : loc;
  r.key[3] = p2 ? r.key[2] : (p3 ? x : r.key[3]); r.loc[3] = p2 ? r.loc[2] : (p3 ? loc : r.loc[3]);
  r.key[2] = p1 ? r.key[1] : (p2 ? x : r.key[2]); r.loc[2] = p1 ? r.loc[1] : (p2 ? loc : r.loc[2]);
  r.key[1] = p0 ? r.key[0] : (p1 ? x : r.key[1]); r.loc[1] = p0 ? r.loc[0] : (p1 ? loc : r.loc[1]);
  r.key[0] = p0 ? x : r.key[0];            r.loc[0] = p0 ? loc : r.loc[0];
}
__device__ __forceinline__ float knnp_d(const KnnP& r, int k) { return __uint_as_float((unsigned)(r.key[k] >> 32)); }

__device__ __forceinline__ void scan_range_p(const GridDev& G, int c0, int c1, float qx, float qy, float qz, KnnP& r) {
  const int s = __ldg(G.cell_start + c0), e = __ldg(G.cell_start + c1 + 1);
  // a candidate farther than the current 5th neighbour is dropped on one float compare (equal distances take the
  // 64-bit key path: ties go to the lower index); two candidates per trip so that both loads are in flight together
  int k = s;
  for (; k + 1 < e; k += 2) {
    const float4 p = __ldg(G.pts + k), p2 = __ldg(G.pts + k + 1);
    const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
    const float d = (dx * dx + dy * dy) + dz * dz;
    const float ex = qx - p2.x, ey = qy - p2.y, ez = qz - p2.z;
    const float d2 = (ex * ex + ey * ey) + ez * ez;
    if (!(d > knnp_d(r, 4))) knnp_push(r, d, __float_as_int(p.w), k);
    if (!(d2 > knnp_d(r, 4))) knnp_push(r, d2, __float_as_int(p2.w), k + 1);
  }
  if (k < e) {
    const float4 p = __ldg(G.pts + k);
    const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
    const float d = (dx * dx + dy * dy) + dz * dz;
    if (!(d > knnp_d(r, 4))) knnp_push(r, d, __float_as_int(p.w), k);
  }
}

// exact 5-NN within squared radius `thres` (same result as knn5_grid). Rows of cells are visited nearest first
// inside each shell and a row (or end cell) is skipped when its box lies farther than the current 5th neighbour.
__device__ bool knn5_grid_packed(const GridDev& G, float qx, float qy, float qz, float thres, KnnP& r) {
  knnp_init(r);
  int c[3], lo[3], hi[3], cube;
  if (!locate(G, qx, qy, qz, c, lo, hi, cube)) return false;
  if (G.global) {
    if (!(__ldg(G.cube_count + cube) > G.min_cube_pts)) return false;
  } else {
    if (!(G.m > G.min_local_pts)) return false;
  }
  const float cellf = G.cell;
  const float cell2 = cellf * cellf;
  // position of the query inside its (possibly clamped) cell, in cells
  const float fx = (float)(((double)qx - G.org[0]) * G.inv_cell - (double)c[0]);
  const float fy = (float)(((double)qy - G.org[1]) * G.inv_cell - (double)c[1]);
  const float fz = (float)(((double)qz - G.org[2]) * G.inv_cell - (double)c[2]);
  const int rmax = (int)ceilf(sqrtf(thres) / cellf) + 1;
  for (int rr = 1; rr <= rmax; rr++) {
    // ring order: |dz| + |dy| ascending puts the nearest rows first, so the bound tightens before the far rows are tested
    for (int sum = 0; sum <= 2 * rr; sum++) {
      for (int adz = 0; adz <= min(sum, rr); adz++) {
        const int ady = sum - adz;
        if (ady > rr) continue;
        for (int sz = (adz ? -1 : 1); sz <= 1; sz += 2) {
          for (int sy = (ady ? -1 : 1); sy <= 1; sy += 2) {
            const int dz = sz * adz, dy = sy * ady;
            const int z = c[2] + dz, y = c[1] + dy;
            if (z < lo[2] || z > hi[2] || y < lo[1] || y > hi[1]) continue;
            const float gy = axis_gap(fy, dy), gz = axis_gap(fz, dz);
            const float row2 = (gy * gy + gz * gz) * cell2;
            const float bound = knnp_d(r, 4);  // +inf until five candidates have been seen
            if (row2 > bound) continue;
            const int row = (z * G.dim[1] + y) * G.dim[0];
            const bool full = (rr == 1) || adz == rr || ady == rr;
            if (full) {
              const int x0 = max(c[0] - rr, lo[0]), x1 = min(c[0] + rr, hi[0]);
              if (x0 <= x1) scan_range_p(G, row + x0, row + x1, qx, qy, qz, r);
            } else {
              const int xa = c[0] - rr, xb = c[0] + rr;
              const float ga = axis_gap(fx, -rr), gb = axis_gap(fx, rr);
              if (xa >= lo[0] && row2 + ga * ga * cell2 <= bound) scan_range_p(G, row + xa, row + xa, qx, qy, qz, r);
              if (xb <= hi[0] && row2 + gb * gb * cell2 <= knnp_d(r, 4)) scan_range_p(G, row + xb, row + xb, qx, qy, qz, r);
            }
          }
        }
      }
    }
    if (c[0] - rr < lo[0] && c[0] + rr > hi[0] && c[1] - rr < lo[1] && c[1] + rr > hi[1] && c[2] - rr < lo[2] &&
        c[2] + rr > hi[2])
      break;
    const float reach = fmaxf((float)rr * cellf - 1e-3f, 0.f);
    const float reach2 = reach * reach;
    if (r.loc[4] >= 0 && knnp_d(r, 4) <= reach2) break;
    if (reach2 >= thres) break;
  }
  return r.loc[4] >= 0 && knnp_d(r, 4) < thres;
}

__device__ __forceinline__ void scan_range_p2(const float4* __restrict__ pts, const int* __restrict__ cell_start, int c0, int c1, float qx,
                                             float qy, float qz, KnnP& r) {
  const int s = __ldg(cell_start + c0), e = __ldg(cell_start + c1 + 1);
  for (int k = s; k < e; k++) {
    const float4 p = __ldg(pts + k);
    const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
    const float d = (dx * dx + dy * dy) + dz * dz;
    if (!(d > knnp_d(r, 4))) knnp_push(r, d, __float_as_int(p.w), k);
  }
}

// knn5_grid_packed for search radii that span many fine cells (MML_TWO_LEVEL_MIN_SHELLS shells or more; chosen by the
// host, k_knn_walk<true>): two levels like the group search. The first
// kFineShellsP shells on the fine cells; a query still open (displaced from the map, or with no neighbours in reach)
// starts over on the coarse cells, whose shells cover `coarse` times the distance per row visited, bounded by the 5th
// distance the fine level found. One copy of the shell walk in the instruction stream, run once per level.
// Returns 0: no acceptable list, 1: the list in `r` indexes the fine copy of the points (G.pts), 3: the coarse copy
// (G.pts2).
#ifndef MML_FINE_SHELLS_P
#define MML_FINE_SHELLS_P 2
#endif
#ifndef MML_TWO_LEVEL_MIN_SHELLS
#define MML_TWO_LEVEL_MIN_SHELLS 20
#endif
__device__ int knn5_grid_packed2(const GridDev& G, float qx, float qy, float qz, float thres, KnnP& r) {
  knnp_init(r);
  int c[3], lo[3], hi[3], cube;
  if (!locate(G, qx, qy, qz, c, lo, hi, cube)) return 0;
  if (G.global) {
    if (!(__ldg(G.cube_count + cube) > G.min_cube_pts)) return 0;
  } else {
    if (!(G.m > G.min_local_pts)) return 0;
  }
  const int rmax = (int)ceilf(sqrtf(thres) / G.cell) + 1;
  const bool two_level = G.pts2 != nullptr && rmax >= MML_TWO_LEVEL_MIN_SHELLS;
  float prune = INFINITY;
  int lvl = 0;
#pragma unroll 1
  for (;; lvl++) {
    const int f = lvl == 0 ? 1 : G.coarse;
    const float4* __restrict__ pts = lvl == 0 ? G.pts : G.pts2;
    const int* __restrict__ cell_start = lvl == 0 ? G.cell_start : G.cell_start2;
    const int dim0 = lvl == 0 ? G.dim[0] : G.dim2[0], dim1 = lvl == 0 ? G.dim[1] : G.dim2[1];
    const float cellf = G.cell * (float)f;
    const float cell2 = cellf * cellf;
    const int cx = c[0] / f, cy = c[1] / f, cz = c[2] / f;
    const int lox = lo[0] / f, loy = lo[1] / f, loz = lo[2] / f, hix = hi[0] / f, hiy = hi[1] / f, hiz = hi[2] / f;
    const int last = lvl == 0 ? (two_level ? MML_FINE_SHELLS_P : rmax) : (int)ceilf(sqrtf(thres) / cellf) + 1;
    // position of the query inside its (possibly clamped) cell of this level, in cells
    const double inv_l = G.inv_cell / (double)f;
    const float fx = (float)(((double)qx - G.org[0]) * inv_l - (double)cx);
    const float fy = (float)(((double)qy - G.org[1]) * inv_l - (double)cy);
    const float fz = (float)(((double)qz - G.org[2]) * inv_l - (double)cz);
    bool done = false;
    for (int rr = 1; rr <= last; rr++) {
      // ring order: |dz| + |dy| ascending puts the nearest rows first, so the bound tightens before the far rows are tested
      for (int sum = 0; sum <= 2 * rr; sum++) {
        for (int adz = 0; adz <= min(sum, rr); adz++) {
          const int ady = sum - adz;
          if (ady > rr) continue;
          for (int sz = (adz ? -1 : 1); sz <= 1; sz += 2) {
            for (int sy = (ady ? -1 : 1); sy <= 1; sy += 2) {
              const int dz = sz * adz, dy = sy * ady;
              const int z = cz + dz, y = cy + dy;
              if (z < loz || z > hiz || y < loy || y > hiy) continue;
              const float gy = axis_gap(fy, dy), gz = axis_gap(fz, dz);
              const float row2 = (gy * gy + gz * gz) * cell2;
              const float bound = fminf(knnp_d(r, 4), prune);  // +inf until five candidates have been seen
              if (row2 > bound) continue;
              const int row = (z * dim1 + y) * dim0;
              const bool full = (rr == 1) || adz == rr || ady == rr;
              if (full) {
                const int x0 = max(cx - rr, lox), x1 = min(cx + rr, hix);
                if (x0 <= x1) scan_range_p2(pts, cell_start, row + x0, row + x1, qx, qy, qz, r);
              } else {
                const int xa = cx - rr, xb = cx + rr;
                const float ga = axis_gap(fx, -rr), gb = axis_gap(fx, rr);
                if (xa >= lox && row2 + ga * ga * cell2 <= bound) scan_range_p2(pts, cell_start, row + xa, row + xa, qx, qy, qz, r);
                if (xb <= hix && row2 + gb * gb * cell2 <= fminf(knnp_d(r, 4), prune))
                  scan_range_p2(pts, cell_start, row + xb, row + xb, qx, qy, qz, r);
              }
            }
          }
        }
      }
      // nothing left inside the searchable block of cells
      if (cx - rr < lox && cx + rr > hix && cy - rr < loy && cy + rr > hiy && cz - rr < loz && cz + rr > hiz) { done = true; break; }
      const float reach = fmaxf((float)rr * cellf - 1e-3f, 0.f);
      const float reach2 = reach * reach;
      if (r.loc[4] >= 0 && knnp_d(r, 4) <= reach2) { done = true; break; }
      if (reach2 >= thres) { done = true; break; }
    }
    if (done || !two_level || lvl == 1) break;
    // the coarse level starts over (its list indexes another copy of the points); five points found on the fine level
    // already bound the 5th distance
    if (r.loc[4] >= 0) prune = knnp_d(r, 4);
    knnp_init(r);
  }
  if (!(r.loc[4] >= 0 && knnp_d(r, 4) < thres)) return 0;
  return lvl == 0 ? 1 : 3;
}

// checkLocalizability (EST.cpp:536-565) on the plane association's own statistics: the last CTA leaves the smallest
// singular value of the stacked normals next to the moments (slot 7 of the plane block of assoc_stats), so the
// solve does not spend its tail on a serial 3x3 eigen-solve. Called by one thread after the moments are final.
__device__ inline void publish_localizability(double* moment_out /* [8]: 6 moments, count, value */) {
  double sv = -1.0;
  if ((int)moment_out[6] > 10) {
    const double* mo = moment_out;
    const double M[9] = {mo[0], mo[1], mo[2], mo[1], mo[3], mo[4], mo[2], mo[4], mo[5]};
    sv = sqrt(fmax(eig3_sym_min(M), 0.0));
  }
  moment_out[7] = sv;
}

// ---------------------------------------------------------------- features
// compact feature record: 3 x float4 per query slot
//   line : f0 = (p.xyz, valid) f1 = (a.xyz, b.x) f2 = (b.y, b.z, 0, 0)
//   plane: f0 = (p.xyz, valid) f1 = (sel.xyz, dist) f2 = (n.xyz, 0)       proj = sel - dist * n in double
// valid: -1 none, 0 |error| <= 1e-5, 1 used.
struct AssocArgs {
  const float4* q;
  int nq;
  const int* nq_dev;        // optional: query count read from device memory (<= nq)
  int* overflow;            // set to 1 when *nq_dev exceeds the launch capacity nq
  const unsigned* perm;     // optional: queries were spatially sorted; perm[i] = original index of query i
  const int* qlist;         // optional: process queries qlist[0 .. *nq_dev) instead of 0 .. nq
  int accumulate_out;       // add to moment_out / n_feat_out instead of overwriting
  // map-sized sets: the neighbour search runs in its own kernel (k_knn_walk) and leaves, per sorted query, the five
  // neighbours' positions in the cell-sorted point array and a status; the fit kernel (k_associate<KIND, true>)
  // picks them up. pre_map = index (0 global / 1 local) of the map that was searched.
  int* pre_loc;             // [nq][5]
  int* pre_status;          // [nq]: -1 no search (query outside the grid / NaN), 0 no 5 neighbours inside thres, 1 / 3 found (pre_loc indexes the fine / coarse copy),
                            //        1 found, 2 undecided (the fit kernel searches itself)
  int pre_map;
  double T[16];
  float thres;
  GridDev G[2];  // [0] global, [1] local
  const GridDev* G_tab[2];  // TAB kernels: the same two descriptors read from device memory (ctx->grid_table), so that a
                            // captured launch stays valid when a map is rebuilt (mml_map_set / mml_local_map_push)
  float4* feat;
  double* moment_partials;  // [grid][8]: 6 moments, count, pad  (plane only)
  unsigned* ticket;
  double* moment_out;       // [8]
  int* n_feat_out;
  const int* gate;          // device flag: nonzero -> skip (used by the estimate graph)
  const double* T_dev;      // optional: T and thres read from device state
  const float* thres_dev;
  unsigned long long* tl;   // MML_TIMELINE
  int tl_slot;
};

template <class KnnT>
__device__ bool fit_line(const float4* pts, const KnnT& r, float* a, float* b) {
  float px[5], py[5], pz[5];
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const float4 p = pts[r.loc[j]];
    px[j] = p.x; py[j] = p.y; pz[j] = p.z;
  }
  float cx = 0, cy = 0, cz = 0;
#pragma unroll
  for (int j = 0; j < 5; j++) { cx += px[j]; cy += py[j]; cz += pz[j]; }
  cx /= 5; cy /= 5; cz /= 5;
  float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const float ax = px[j] - cx, ay = py[j] - cy, az = pz[j] - cz;
    a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
    a22 += ay * ay; a23 += ay * az; a33 += az * az;
  }
  a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
  const double A[9] = {a11, a12, a13, a12, a22, a23, a13, a23, a33};
  double ev[3], V[9];
  eig3_sym(A, ev, V);
  if (!(ev[2] > 3 * ev[1])) return false;
  const double u0 = V[2], u1 = V[5], u2 = V[8];
  a[0] = (float)((double)cx + 0.1 * u0); a[1] = (float)((double)cy + 0.1 * u1); a[2] = (float)((double)cz + 0.1 * u2);
  b[0] = (float)((double)cx - 0.1 * u0); b[1] = (float)((double)cy - 0.1 * u1); b[2] = (float)((double)cz - 0.1 * u2);
  return true;
}

template <class KnnT>
__device__ bool fit_plane(const float4* pts, const KnnT& r, float sx, float sy, float sz, float* nrm, float* dist_out) {
  double A[5][3], bb[5];
  float px[5], py[5], pz[5];
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const float4 p = pts[r.loc[j]];
    px[j] = p.x; py[j] = p.y; pz[j] = p.z;
    A[j][0] = p.x; A[j][1] = p.y; A[j][2] = p.z;
    bb[j] = -1.0;
  }
  double X[3];
  qr5x3_solve(A, bb, X);
  float pa = (float)X[0], pb = (float)X[1], pc = (float)X[2], pd = 1.f;
  const float ps = sqrtf(pa * pa + pb * pb + pc * pc);
  pa /= ps; pb /= ps; pc /= ps; pd /= ps;
#pragma unroll
  for (int j = 0; j < 5; j++)
    if ((double)fabsf(pa * px[j] + pb * py[j] + pc * pz[j] + pd) > 0.2) return false;
  nrm[0] = pa; nrm[1] = pb; nrm[2] = pc;
  *dist_out = pa * sx + pb * sy + pc * sz + pd;
  return true;
}

// ---------------------------------------------------------------- map-sized sets: search kernel
// Map-sized query sets (S4 / S5) run the neighbour search and the fit as two kernels: the search keeps 56 registers
// (the fused kernel: 72 + 248 B of stack for the float64 QR), the fit runs with every lane busy, and the pair is
// 12 % faster than the fused kernel at 1.05 M queries (0.685 vs 0.776 ms, profiles/r2_s4_knn_experiments.txt).
template <bool TWO>
__global__ void __launch_bounds__(128) k_knn_walk(AssocArgs A) {
  if (A.gate && *A.gate) return;
  const int i = blockIdx.x * 128 + threadIdx.x;
  int nq = A.nq_dev ? *A.nq_dev : A.nq;
  if (nq > A.nq) {  // more queries than this launch was sized for: flag it, the host re-launches
    if (i == 0 && A.overflow) atomicExch(A.overflow, 1);
    nq = A.nq;
  }
  if (i >= nq) return;
  const GridDev& G = A.G[A.pre_map];
  const float thres = A.thres_dev ? *A.thres_dev : A.thres;
  // both variants are launched when the radius is only known on the device: the one it does not call for leaves
  if ((G.pts2 != nullptr && (int)ceilf(sqrtf(thres) / G.cell) + 1 >= MML_TWO_LEVEL_MIN_SHELLS) != TWO) return;
  const double* __restrict__ Tq = A.T_dev ? A.T_dev : A.T;
  const float4 q = A.q[i];
  const double pin[3] = {(double)q.x, (double)q.y, (double)q.z};
  float sel[3];
#pragma unroll
  for (int rr = 0; rr < 3; rr++)
    sel[rr] = (float)(((Tq[4 * rr] * pin[0] + Tq[4 * rr + 1] * pin[1]) + Tq[4 * rr + 2] * pin[2]) + Tq[4 * rr + 3]);
  int cI, cJ, cK;
  const bool in_grid = cube_of(sel[0], sel[1], sel[2], A.G[0].cen, cI, cJ, cK);
  const bool finite = !(isnan(sel[0]) || isnan(sel[1]) || isnan(sel[2]));
  int status = -1;
  if (in_grid && finite) {
    KnnP r;
    status = TWO ? knn5_grid_packed2(G, sel[0], sel[1], sel[2], thres, r) : (knn5_grid_packed(G, sel[0], sel[1], sel[2], thres, r) ? 1 : 0);
    if (status != 0) {
#pragma unroll
      for (int k = 0; k < 5; k++) A.pre_loc[5 * (size_t)i + k] = r.loc[k];
    }
  }
  A.pre_status[i] = status;
}

// Sum of the per-CTA moment partials by the last CTA (128 threads): every thread adds a strided share, then a fixed
// shuffle / shared-memory tree - deterministic, and ~60 dependent loads per thread instead of thousands on seven.
__device__ __forceinline__ void reduce_moment_partials(const double* __restrict__ partials, unsigned nblk, double (*sred)[7], double* out7) {
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (unsigned b = threadIdx.x; b < nblk; b += 128) {
#pragma unroll
    for (int k = 0; k < 7; k++) acc[k] += __ldcg(partials + (size_t)b * 8 + k);
  }
  __syncthreads();  // sred is reused
#pragma unroll
  for (int k = 0; k < 7; k++) {
    double v = acc[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) out7[threadIdx.x] = ((sred[0][threadIdx.x] + sred[1][threadIdx.x]) + sred[2][threadIdx.x]) + sred[3][threadIdx.x];
  __syncthreads();
}

// PRE: the neighbour search of map A.pre_map was done by k_knn_walk
template <int KIND, bool PRE>
__global__ void __launch_bounds__(128) k_associate(AssocArgs A) {
  if (A.gate && *A.gate) return;
  const int i = blockIdx.x * 128 + threadIdx.x;
  double T[16];
  float thres = A.thres;
  if (A.T_dev) {
#pragma unroll
    for (int k = 0; k < 16; k++) T[k] = A.T_dev[k];
    thres = *A.thres_dev;
  } else {
#pragma unroll
    for (int k = 0; k < 16; k++) T[k] = A.T[k];
  }
  double mom[7] = {0, 0, 0, 0, 0, 0, 0};
  int found = 0;
  int nq = A.nq_dev ? *A.nq_dev : A.nq;
  if (nq > A.nq) {  // more queries than this launch was sized for: flag it, the host re-launches
    if (i == 0 && A.overflow) atomicExch(A.overflow, 1);
    nq = A.nq;
  }
  if (i < nq) {
    const float4 q = A.q[i];
    const double pin[3] = {(double)q.x, (double)q.y, (double)q.z};
    float sel[3];
#pragma unroll
    for (int rr = 0; rr < 3; rr++)
      sel[rr] = (float)(((T[4 * rr] * pin[0] + T[4 * rr + 1] * pin[1]) + T[4 * rr + 2] * pin[2]) + T[4 * rr + 3]);
    float4 f0 = make_float4(q.x, q.y, q.z, -1.f), f1 = make_float4(0, 0, 0, 0), f2 = make_float4(0, 0, 0, 0);
    // EST.cpp:192-196: out of the cube grid or NaN -> no feature at all
    int cI, cJ, cK;
    const bool in_grid = cube_of(sel[0], sel[1], sel[2], A.G[0].cen, cI, cJ, cK);
    const bool finite = !(isnan(sel[0]) || isnan(sel[1]) || isnan(sel[2]));
    if (in_grid && finite) {
      KnnP r;
      for (int mp = 0; mp < 2 && !found; mp++) {
        const GridDev& G = A.G[mp];
        if (!G.valid) continue;
        bool have;
        int pre = 2;  // 2: not searched yet; else the walk's verdict (0 none, 1 fine copy, 3 coarse copy; -1 outside)
        if (PRE && mp == A.pre_map) pre = A.pre_status[i];
        if (pre == 2) {
          have = knn5_grid_packed(G, sel[0], sel[1], sel[2], thres, r);
          pre = have ? 1 : 0;
        } else {
          have = pre == 1 || pre == 3;
          if (have) {
#pragma unroll
            for (int k = 0; k < 5; k++) r.loc[k] = A.pre_loc[5 * (size_t)i + k];
          }
        }
        if (!have) continue;
        if (KIND == 0) {
          float a[3], b[3];
          if (!fit_line(pre == 3 ? G.pts2 : G.pts, r, a, b)) continue;
          f1 = make_float4(a[0], a[1], a[2], b[0]);
          f2 = make_float4(b[1], b[2], 0.f, 0.f);
          // Estimator.h:71-83 FeatureLine::ComputeError at the association pose
          double P[3];
#pragma unroll
          for (int rr = 0; rr < 3; rr++)
            P[rr] = ((T[4 * rr] * pin[0] + T[4 * rr + 1] * pin[1]) + T[4 * rr + 2] * pin[2]) + T[4 * rr + 3];
          const double da[3] = {a[0], a[1], a[2]}, db[3] = {b[0], b[1], b[2]};
          const double l12 = sqrt((da[0] - db[0]) * (da[0] - db[0]) + (da[1] - db[1]) * (da[1] - db[1]) +
                                  (da[2] - db[2]) * (da[2] - db[2]));
          const double c0 = (P[0] - da[0]) * (P[1] - db[1]) - (P[0] - db[0]) * (P[1] - da[1]);
          const double c1 = (P[0] - da[0]) * (P[2] - db[2]) - (P[0] - db[0]) * (P[2] - da[2]);
          const double c2 = (P[1] - da[1]) * (P[2] - db[2]) - (P[1] - db[1]) * (P[2] - da[2]);
          const double err = sqrt(c0 * c0 + c1 * c1 + c2 * c2) / l12;
          f0.w = (fabs(err) > 1e-5) ? 1.f : 0.f;
          f2.z = (float)err;
          found = 1;
        } else {
          float nrm[3], dist;
          if (!fit_plane(pre == 3 ? G.pts2 : G.pts, r, sel[0], sel[1], sel[2], nrm, &dist)) continue;
          f1 = make_float4(sel[0], sel[1], sel[2], dist);
          f2 = make_float4(nrm[0], nrm[1], nrm[2], 0.f);
          double e[3];
#pragma unroll
          for (int rr = 0; rr < 3; rr++) {
            const double P = ((T[4 * rr] * pin[0] + T[4 * rr + 1] * pin[1]) + T[4 * rr + 2] * pin[2]) + T[4 * rr + 3];
            const double proj = (double)sel[rr] - (double)dist * (double)nrm[rr];
            e[rr] = P - proj;
          }
          const double err = sqrt((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]);
          f0.w = (fabs(err) > 1e-5) ? 1.f : 0.f;
          f2.w = (float)err;
          const double n0 = nrm[0], n1 = nrm[1], n2 = nrm[2];
          mom[0] = n0 * n0; mom[1] = n0 * n1; mom[2] = n0 * n2; mom[3] = n1 * n1; mom[4] = n1 * n2; mom[5] = n2 * n2;
          found = 1;
        }
      }
    }
    const size_t slot = A.perm ? (size_t)A.perm[i] : (size_t)i;  // features live in the caller's query order
    A.feat[3 * slot] = f0;
    A.feat[3 * slot + 1] = f1;
    A.feat[3 * slot + 2] = f2;
  }
  // ---- block reduction of (moments, count); last block sums the partials in fixed order
  mom[6] = (double)found;
  __shared__ double sred[4][7];
  __shared__ bool is_last;
#pragma unroll
  for (int k = 0; k < 7; k++) {
    double v = mom[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    const double v = ((sred[0][threadIdx.x] + sred[1][threadIdx.x]) + sred[2][threadIdx.x]) + sred[3][threadIdx.x];
    A.moment_partials[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(A.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {  // block-uniform
    __threadfence();
    __shared__ double tot7[7];
    reduce_moment_partials(A.moment_partials, gridDim.x, sred, tot7);
    if (threadIdx.x < 7) {
      const double s = tot7[threadIdx.x];
      A.moment_out[threadIdx.x] = s;
      if (threadIdx.x == 6) *A.n_feat_out = (int)s;
    }
    __syncwarp();
    if (KIND == 1 && threadIdx.x == 0) publish_localizability(A.moment_out);
    if (threadIdx.x == 0) *A.ticket = 0;
  }
}

// ---------------------------------------------------------------- group-per-query search
// G lanes cooperate on one query: the cells of a shell are dealt out to the lanes (ring 1: the 27
// cells individually, outer shells: x-rows / end cells), every lane keeps its own sorted 5-list, and
// after each shell the lists are merged with shuffle arg-min rounds into the exact global 5-list
// (same (d2, index) order as the one-thread search, so results are identical). This removes the
// serial chain of dependent L2 loads that bounds the latency of the thread-per-query kernel.
template <int G>
__device__ __forceinline__ unsigned group_mask() {
  if constexpr (G == 32) {
    return 0xffffffffu;
  } else {
    const unsigned lane = threadIdx.x & 31u;
    return ((1u << G) - 1u) << (lane & ~(unsigned)(G - 1));
  }
}

template <int G>
__device__ __forceinline__ void group_merge(const Knn5& r, Knn5& m, unsigned mask, int lg) {
  // five arg-min rounds over the heads of the lanes' sorted lists. Squared distances are >= +0, so their
  // bit patterns order like unsigned integers: one redux for the distance, one for the index tie-break.
  int ptr = 0;
  m.cnt = 0;
  const int my_lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const float d = ptr == 0 ? r.d[0] : ptr == 1 ? r.d[1] : ptr == 2 ? r.d[2] : ptr == 3 ? r.d[3] : ptr == 4 ? r.d[4] : INFINITY;
    const int id = ptr == 0 ? r.id[0] : ptr == 1 ? r.id[1] : ptr == 2 ? r.id[2] : ptr == 3 ? r.id[3] : ptr == 4 ? r.id[4] : 0x7fffffff;
    const int loc = ptr == 0 ? r.loc[0] : ptr == 1 ? r.loc[1] : ptr == 2 ? r.loc[2] : ptr == 3 ? r.loc[3] : ptr == 4 ? r.loc[4] : -1;
    const unsigned db = __float_as_uint(d);
    const unsigned dmin = __reduce_min_sync(mask, db);
    const unsigned idc = db == dmin ? (unsigned)id : 0xffffffffu;
    const unsigned idmin = __reduce_min_sync(mask, idc);
    const unsigned win = __ballot_sync(mask, db == dmin && (unsigned)id == idmin);
    const int src = __ffs(win) - 1;  // lowest winning lane (several only for the INF padding entries)
    m.d[k] = __uint_as_float(dmin);
    m.id[k] = (int)idmin;
    m.loc[k] = __shfl_sync(mask, loc, src);
    if (dmin < 0x7f800000u) m.cnt++;
    if (my_lane == src) ptr++;
  }
}

// one level of the hash (fine or coarse): its sorted points, cell table and geometry
struct GridLevel {
  const float4* pts;
  const int* cell_start;
  int dim[3];
  float cell;
};

__device__ __forceinline__ void scan_cells(const GridLevel& L, int c0, int c1, float qx, float qy, float qz, Knn5& r) {
  const int s = __ldg(L.cell_start + c0), e = __ldg(L.cell_start + c1 + 1);
  for (int k = s; k < e; k++) {
    const float4 p = __ldg(L.pts + k);
    const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
    const float d = (dx * dx + dy * dy) + dz * dz;
    knn_push(r, d, __float_as_int(p.w), k);
  }
}

// shells rr0..rr1 of one level. Returns true when the search is finished (5-list final, or proven that no
// acceptable 5-list exists); false when rr1 was exhausted without a verdict.
template <int G>
__device__ bool search_shells(const GridLevel& L, const int* c, const int* lo, const int* hi, float qx, float qy, float qz,
                              float thres, int rr0, int rr1, Knn5& r, Knn5& m, unsigned mask, int lg, int* lst,
                              const float* fpos, float prune_d) {
  constexpr int kListCap = 8 * G;  // candidate indices staged per group
  const float cell2 = L.cell * L.cell;
  for (int rr = rr0; rr <= rr1; rr++) {
    const int side = 2 * rr + 1;
    const int nseg = rr == 1 ? 27 : 2 * side * side;
    // every lane resolves one segment (cell or x-row) to its point range; the non-empty ranges are then scanned
    // by the whole group together (lane-strided, coalesced), so a long row - e.g. along a map edge - is not
    // left to a single lane
    for (int base = 0; base < nseg; base += G) {
      const int s = base + lg;
      int p0 = 0, p1 = 0;
      if (s < nseg) {
        int y, z, x0, x1;
        bool ok = true;
        if (rr == 1) {
          const int dz = s / 9 - 1, dy = (s / 3) % 3 - 1, dx = s % 3 - 1;
          z = c[2] + dz; y = c[1] + dy; x0 = x1 = c[0] + dx;
          ok = !(x0 < lo[0] || x0 > hi[0]);
        } else {
          const int row = s >> 1, which = s & 1;
          const int dz = row / side - rr, dy = row % side - rr;
          z = c[2] + dz; y = c[1] + dy;
          const bool full = dz == -rr || dz == rr || dy == -rr || dy == rr;
          if (full) {
            x0 = max(c[0] - rr, lo[0]); x1 = min(c[0] + rr, hi[0]);
            ok = !which && x0 <= x1;
          } else {
            x0 = x1 = which ? c[0] + rr : c[0] - rr;
            ok = !(x0 < lo[0] || x0 > hi[0]);
          }
        }
        ok = ok && !(z < lo[2] || z > hi[2] || y < lo[1] || y > hi[1]);
        const float bound = m.cnt == 5 ? fminf(m.d[4], prune_d) : prune_d;  // prune_d: 5th distance of the finer level
        if (ok && bound < INFINITY) {
          // outer shells of a far query: a row / end cell whose box lies farther than the current 5th neighbour
          // (list merged after the previous shell) cannot change the result and is not read at all
          const float gy = axis_gap(fpos[1], y - c[1]), gz = axis_gap(fpos[2], z - c[2]);
          float g2 = gy * gy + gz * gz;
          if (x0 == x1) { const float gx = axis_gap(fpos[0], x0 - c[0]); g2 += gx * gx; }
          ok = !(g2 * cell2 > bound);
        }
        if (ok) {
          const int rowb = (z * L.dim[1] + y) * L.dim[0];
          p0 = __ldg(L.cell_start + rowb + x0);
          p1 = __ldg(L.cell_start + rowb + x1 + 1);
        }
      }
      // The non-empty ranges are flattened into one candidate list in shared memory (group prefix sum of the range
      // lengths), so that the point loads of a whole batch of cells are issued back to back - one memory round trip
      // per batch instead of one per cell range, which is what a lone query's latency is made of.
      const int len = p1 - p0;
      int incl = len;
#pragma unroll
      for (int d = 1; d < G; d <<= 1) {
        const int y = __shfl_up_sync(mask, incl, d, G);
        if (lg >= d) incl += y;
      }
      const int excl = incl - len;
      const int total = __shfl_sync(mask, incl, G - 1, G);
      for (int cbase = 0; cbase < total; cbase += kListCap) {
        const int s0 = max(excl, cbase), s1 = min(incl, cbase + kListCap);
        for (int c = s0; c < s1; c++) lst[c - cbase] = p0 + (c - excl);
        __syncwarp(mask);
        const int ncand = min(total - cbase, kListCap);
        for (int c0 = 0; c0 < ncand; c0 += 4 * G) {
          int kk[4];
          float4 pp[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int c = c0 + u * G + lg;
            kk[u] = c < ncand ? lst[c] : -1;
            if (kk[u] >= 0) pp[u] = __ldg(L.pts + kk[u]);
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            if (kk[u] < 0) continue;
            const float dx = qx - pp[u].x, dy = qy - pp[u].y, dz = qz - pp[u].z;
            const float d = (dx * dx + dy * dy) + dz * dz;
            knn_push(r, d, __float_as_int(pp[u].w), kk[u]);
          }
        }
        __syncwarp(mask);
      }
    }
    group_merge<G>(r, m, mask, lg);
    // lane 0 carries the merged list forward, the others start the next shell empty (no duplicates)
    if (lg == 0) r = m;
    else knn_init(r);
    if (c[0] - rr < lo[0] && c[0] + rr > hi[0] && c[1] - rr < lo[1] && c[1] + rr > hi[1] && c[2] - rr < lo[2] &&
        c[2] + rr > hi[2])
      return true;  // nothing left inside the searchable block of cells
    const float reach = fmaxf((float)rr * L.cell - 1e-3f, 0.f);
    const float reach2 = reach * reach;
    if (m.cnt == 5 && m.d[4] <= reach2) return true;
    if (reach2 >= thres) return true;
  }
  return false;
}

// Two-level search: the first kFineShells shells on the fine cells (where almost every query ends); queries
// that are still open restart on the coarse level (cells kCoarse times larger), which bounds the number of
// cell rows a far or hopeless query has to walk by kCoarse^2 per shell and kCoarse fewer shells.
__device__ int g_fine_shells = 2;
#define kFineShells g_fine_shells

template <int G>
__device__ bool knn5_grid_group(const GridDev& Gd, float qx, float qy, float qz, float thres, Knn5& m, unsigned mask, int lg,
                                const float4*& base, int* lst) {
  base = Gd.pts;
  Knn5 r;
  knn_init(r);
  knn_init(m);
  int c[3], lo[3], hi[3], cube;
  if (!locate(Gd, qx, qy, qz, c, lo, hi, cube)) return false;
  if (Gd.global) {
    if (!(__ldg(Gd.cube_count + cube) > Gd.min_cube_pts)) return false;
  } else {
    if (!(Gd.m > Gd.min_local_pts)) return false;
  }
  const int rmax = (int)ceilf(sqrtf(thres) / Gd.cell) + 1;
  const bool two_level = Gd.pts2 != nullptr && rmax > kFineShells;
  // one copy of the shell search in the instruction stream, run once per level (a query's warp executes this kernel
  // exactly once, so every extra inlined copy is paid for in instruction fetch, ncu's `no_instruction` stall)
  float prune_d = INFINITY;
#pragma unroll 1
  for (int lvl = 0; lvl < 2; lvl++) {
    GridLevel L;
    int cc[3], llo[3], hhi[3];
    const int f = lvl == 0 ? 1 : Gd.coarse;
    L.pts = lvl == 0 ? Gd.pts : Gd.pts2;
    L.cell_start = lvl == 0 ? Gd.cell_start : Gd.cell_start2;
    L.cell = Gd.cell * (float)f;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      L.dim[a] = lvl == 0 ? Gd.dim[a] : Gd.dim2[a];
      cc[a] = c[a] / f; llo[a] = lo[a] / f; hhi[a] = hi[a] / f;
    }
    const int last = lvl == 0 ? (two_level ? kFineShells : rmax) : (int)ceilf(sqrtf(thres) / L.cell) + 1;
    // position of the query inside its (possibly clamped) cell of this level, in cells
    const double inv_l = Gd.inv_cell / (double)f;
    const float fpos[3] = {(float)(((double)qx - Gd.org[0]) * inv_l - (double)cc[0]),
                           (float)(((double)qy - Gd.org[1]) * inv_l - (double)cc[1]),
                           (float)(((double)qz - Gd.org[2]) * inv_l - (double)cc[2])};
    const bool done = search_shells<G>(L, cc, llo, hhi, qx, qy, qz, thres, 1, last, r, m, mask, lg, lst, fpos, prune_d);
    base = L.pts;  // the 5-list indexes this level's sorted copy
    if (done || !two_level) break;
    // the coarse level starts over (its list indexes another copy of the points), but five points found on the
    // fine level already bound the 5th distance: cells farther than that are skipped from the first coarse shell on
    if (m.cnt == 5) prune_d = m.d[4];
    knn_init(r);
    knn_init(m);
  }
  return m.cnt == 5 && m.d[4] < thres;
}

template <int KIND, int G, bool TAB = false>
__global__ void __launch_bounds__(128) k_associate_g(AssocArgs A) {
  if (A.gate && *A.gate) return;
  constexpr int QPB = 128 / G;  // queries per block
  __shared__ int s_lst[QPB][8 * G];
  __shared__ GridDev s_G[TAB ? 2 : 1];
  if (TAB) {
    constexpr int kWords = (int)(sizeof(GridDev) / 4);
    for (int i = threadIdx.x; i < 2 * kWords; i += 128)
      reinterpret_cast<unsigned*>(s_G)[i] = reinterpret_cast<const unsigned*>(A.G_tab[i / kWords])[i % kWords];
    __syncthreads();
  }
  int* lst = s_lst[threadIdx.x / G];
  if (blockIdx.x == 0 && threadIdx.x == 0) MML_TL(A.tl, A.tl_slot);
  const int lg = threadIdx.x % G;
  const unsigned mask = group_mask<G>();
  double T[16];
  float thres = A.thres;
  if (A.T_dev) {
#pragma unroll
    for (int k = 0; k < 16; k++) T[k] = A.T_dev[k];
    thres = *A.thres_dev;
  } else {
#pragma unroll
    for (int k = 0; k < 16; k++) T[k] = A.T[k];
  }
  double mom[7] = {0, 0, 0, 0, 0, 0, 0};
  int nq = A.nq_dev ? *A.nq_dev : A.nq;
  if (nq > A.nq) {  // more queries than this launch was sized for: flag it, the host re-launches
    if (blockIdx.x == 0 && threadIdx.x == 0 && A.overflow) atomicExch(A.overflow, 1);
    nq = A.nq;
  }
  // The grid is sized for the machine, not for the capacity of the query buffer: groups stride over the queries.
  // CTAs beyond the last query leave at once; the reduction below only spans the active ones
  const unsigned n_need = nq > 0 ? (unsigned)((nq + QPB - 1) / QPB) : 1u;
  const unsigned n_active = n_need < gridDim.x ? n_need : gridDim.x;
  if (blockIdx.x >= n_active) return;
  for (int slot_i = blockIdx.x * QPB + threadIdx.x / G; slot_i < nq; slot_i += gridDim.x * QPB) {
    int found = 0;
#ifdef MML_TIMELINE
    const long long qc0 = clock64();
#endif
    const int i = A.qlist ? A.qlist[slot_i] : slot_i;
    const float4 q = A.q[i];
    const double pin[3] = {(double)q.x, (double)q.y, (double)q.z};
    float sel[3];
#pragma unroll
    for (int rr = 0; rr < 3; rr++)
      sel[rr] = (float)(((T[4 * rr] * pin[0] + T[4 * rr + 1] * pin[1]) + T[4 * rr + 2] * pin[2]) + T[4 * rr + 3]);
    float4 f0 = make_float4(q.x, q.y, q.z, -1.f), f1 = make_float4(0, 0, 0, 0), f2 = make_float4(0, 0, 0, 0);
    int cI, cJ, cK;
    const bool in_grid = cube_of(sel[0], sel[1], sel[2], (TAB ? s_G[0] : A.G[0]).cen, cI, cJ, cK);
    const bool finite = !(isnan(sel[0]) || isnan(sel[1]) || isnan(sel[2]));
    if (in_grid && finite) {
      Knn5 r;
#pragma unroll 1
      for (int mp = 0; mp < 2 && !found; mp++) {
        const GridDev& Gd = TAB ? s_G[mp] : A.G[mp];
        if (!Gd.valid) continue;
        const float4* base;
#ifdef MML_TIMELINE
        const long long qc1 = clock64();
#endif
        const bool knn_ok = knn5_grid_group<G>(Gd, sel[0], sel[1], sel[2], thres, r, mask, lg, base, lst);
#ifdef MML_TIMELINE
        if (A.tl && lg == 0) {
          unsigned* dbg = reinterpret_cast<unsigned*>(A.tl + 16 + 8 * 4000) + 32768;
          const int sl = (KIND ? 16384 : 0) + (slot_i < 16384 ? slot_i : 16383);
          dbg[sl] = (unsigned)(qc1 - qc0);                 // setup
          dbg[32768 + sl] = (unsigned)(clock64() - qc1);   // search
        }
#endif
        if (!knn_ok) continue;  // group-uniform
        int ok = 0;
        if (lg == 0) {
          if (KIND == 0) {
            float a[3], b[3];
            if (fit_line(base, r, a, b)) {
              f1 = make_float4(a[0], a[1], a[2], b[0]);
              f2 = make_float4(b[1], b[2], 0.f, 0.f);
              double P[3];
#pragma unroll
              for (int rr = 0; rr < 3; rr++)
                P[rr] = ((T[4 * rr] * pin[0] + T[4 * rr + 1] * pin[1]) + T[4 * rr + 2] * pin[2]) + T[4 * rr + 3];
              const double da[3] = {a[0], a[1], a[2]}, db[3] = {b[0], b[1], b[2]};
              const double l12 = sqrt((da[0] - db[0]) * (da[0] - db[0]) + (da[1] - db[1]) * (da[1] - db[1]) +
                                      (da[2] - db[2]) * (da[2] - db[2]));
              const double c0 = (P[0] - da[0]) * (P[1] - db[1]) - (P[0] - db[0]) * (P[1] - da[1]);
              const double c1 = (P[0] - da[0]) * (P[2] - db[2]) - (P[0] - db[0]) * (P[2] - da[2]);
              const double c2 = (P[1] - da[1]) * (P[2] - db[2]) - (P[1] - db[1]) * (P[2] - da[2]);
              const double err = sqrt(c0 * c0 + c1 * c1 + c2 * c2) / l12;
              f0.w = (fabs(err) > 1e-5) ? 1.f : 0.f;
              f2.z = (float)err;
              ok = 1;
            }
          } else {
            float nrm[3], dist;
            if (fit_plane(base, r, sel[0], sel[1], sel[2], nrm, &dist)) {
              f1 = make_float4(sel[0], sel[1], sel[2], dist);
              f2 = make_float4(nrm[0], nrm[1], nrm[2], 0.f);
              double e[3];
#pragma unroll
              for (int rr = 0; rr < 3; rr++) {
                const double P = ((T[4 * rr] * pin[0] + T[4 * rr + 1] * pin[1]) + T[4 * rr + 2] * pin[2]) + T[4 * rr + 3];
                const double proj = (double)sel[rr] - (double)dist * (double)nrm[rr];
                e[rr] = P - proj;
              }
              const double err = sqrt((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]);
              f0.w = (fabs(err) > 1e-5) ? 1.f : 0.f;
              f2.w = (float)err;
              const double n0 = nrm[0], n1 = nrm[1], n2 = nrm[2];
              mom[0] += n0 * n0; mom[1] += n0 * n1; mom[2] += n0 * n2; mom[3] += n1 * n1; mom[4] += n1 * n2; mom[5] += n2 * n2;
              ok = 1;
            }
          }
        }
#ifdef MML_TIMELINE
        if (A.tl && lg == 0) {
          unsigned* dbg = reinterpret_cast<unsigned*>(A.tl + 16 + 8 * 4000) + 32768 * 3;
          const int sl = (KIND ? 16384 : 0) + (slot_i < 16384 ? slot_i : 16383);
          dbg[sl] = (unsigned)(clock64() - qc0);  // up to and including the fit
        }
#endif
        found = __shfl_sync(mask, ok, 0, G);
      }
    }
    if (lg == 0) {
      const size_t slot = A.perm ? (size_t)A.perm[i] : (size_t)i;
      A.feat[3 * slot] = f0;
      A.feat[3 * slot + 1] = f1;
      A.feat[3 * slot + 2] = f2;
      mom[6] += (double)found;
#ifdef MML_TIMELINE
      if (A.tl) reinterpret_cast<unsigned*>(A.tl + 16 + 8 * 4000)[(KIND ? 16384 : 0) + (slot_i < 16384 ? slot_i : 16383)] = (unsigned)(clock64() - qc0);
#endif
    }
  }
  __shared__ double sred[4][7];
  __shared__ bool is_last;
#pragma unroll
  for (int k = 0; k < 7; k++) {
    double v = mom[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    const double v = ((sred[0][threadIdx.x] + sred[1][threadIdx.x]) + sred[2][threadIdx.x]) + sred[3][threadIdx.x];
    A.moment_partials[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(A.ticket, 1u) == n_active - 1);
  __syncthreads();
  if (is_last) {  // block-uniform
    __threadfence();
    __shared__ double tot7[7];
    reduce_moment_partials(A.moment_partials, n_active, sred, tot7);
    if (threadIdx.x < 7) {
      double s = tot7[threadIdx.x];
      if (A.accumulate_out) s += A.moment_out[threadIdx.x];
      A.moment_out[threadIdx.x] = s;
      if (threadIdx.x == 6) *A.n_feat_out = (int)s;
    }
    __syncwarp();
    if (KIND == 1 && threadIdx.x == 0) publish_localizability(A.moment_out);
    if (threadIdx.x == 0) *A.ticket = 0;
    if (threadIdx.x == 0) MML_TL(A.tl, A.tl_slot + 1);
  }
}

// expand compact features to the host-visible 12-double records of include/mmloam_b200.h
template <int KIND>
__global__ void __launch_bounds__(256) k_export_features(const float4* __restrict__ feat, int nq, double* __restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= nq) return;
  const float4 f0 = feat[3 * (size_t)i], f1 = feat[3 * (size_t)i + 1], f2 = feat[3 * (size_t)i + 2];
  double* o = out + 12 * (size_t)i;
  o[0] = f0.x; o[1] = f0.y; o[2] = f0.z;
  if (f0.w < 0.f) {
    for (int k = 3; k < 10; k++) o[k] = 0;
  } else if (KIND == 0) {
    o[3] = f1.x; o[4] = f1.y; o[5] = f1.z;
    o[6] = f1.w; o[7] = f2.x; o[8] = f2.y;
    o[9] = f2.z;
  } else {
    o[3] = (double)f1.x - (double)f1.w * (double)f2.x;
    o[4] = (double)f1.y - (double)f1.w * (double)f2.y;
    o[5] = (double)f1.z - (double)f1.w * (double)f2.z;
    o[6] = f2.x; o[7] = f2.y; o[8] = f2.z;
    o[9] = f2.w;
  }
  o[10] = (double)f0.w;
  o[11] = (double)i;
}

}  // namespace mml

using namespace mml;

// ---------------------------------------------------------------- host side
constexpr int kCoarseFactor = 4;
static int build_grid(mml_ctx* ctx, GridMap& M, const float4* pts_d, int m, float cell_hint, const float* bbox6);
int mml_grid_table_sync(mml_ctx* ctx);

// bbox6 (may be NULL): a bounding box of the points the caller already has (min xyz, max xyz; it may be loose): the
// build then needs no pass over the points and no host synchronisation of its own to size the grid
int mml_map_set_device(mml_ctx* ctx, int kind, const float4* pts_d, int m, const int* cen3, float cell_hint, const float* bbox6) {
  if (kind < 0 || kind > 3) return mml_fail(ctx, MML_ERR_INVALID, "map kind must be 0..3");
  GridMap& M = ctx->maps[kind];
  M.global = (kind == MML_MAP_CORNER_GLOBAL || kind == MML_MAP_SURF_GLOBAL);
  if (cen3) { M.cen[0] = cen3[0]; M.cen[1] = cen3[1]; M.cen[2] = cen3[2]; }
  M.valid = false;
  M.m = m;
  ctx->grid_table_dirty = true;
  int rc = MML_OK;
  if (m > 0) rc = build_grid(ctx, M, pts_d, m, cell_hint, bbox6);
  // captured association launches read the descriptors from device memory: keep that copy current
  if (ctx->grid_table.p) { const int rc2 = mml_grid_table_sync(ctx); if (rc == MML_OK) rc = rc2; }
  return rc;
}

extern int mml_bbox_device(mml_ctx* ctx, const float4* pts_d, int n, float* mn3, float* mx3);

static GridDev grid_dev(const GridMap& M, int kind) {
  GridDev G;
  memset(&G, 0, sizeof(G));
  G.valid = M.valid ? 1 : 0;
  if (!M.valid) return G;
  G.pts = M.pts.as<float4>();
  G.cell_start = M.cell_start.as<int>();
  G.cube_count = M.cube_count.as<int>();
  G.pts2 = M.coarse > 1 ? M.pts2.as<float4>() : nullptr;
  G.cell_start2 = M.cell_start2.as<int>();
  G.coarse = M.coarse;
  for (int a = 0; a < 3; a++) { G.org[a] = M.org_d[a]; G.dim[a] = M.dim[a]; G.cen[a] = M.cen[a]; G.cube_lo[a] = M.cube_lo[a]; G.dim2[a] = M.dim2[a]; }
  G.inv_cell = 1.0 / (double)M.cell;
  G.cell = M.cell;
  G.m = M.m;
  G.global = M.global ? 1 : 0;
  G.k_per_cube = M.k_per_cube;
  G.min_cube_pts = (kind == MML_MAP_CORNER_GLOBAL) ? 100 : 50;
  G.min_local_pts = 20;
  return G;
}

static int build_grid(mml_ctx* ctx, GridMap& M, const float4* pts_d, int m, float cell_hint, const float* bbox6) {
  cudaStream_t st = ctx->stream;
  float mn[3], mx[3];
  if (bbox6) {
    for (int a = 0; a < 3; a++) { mn[a] = bbox6[a]; mx[a] = bbox6[3 + a]; }
  } else {
    MML_CHECK(mml_bbox_device(ctx, pts_d, m, mn, mx));
  }
  const long long kMaxCells = 1ll << 28;

  auto layout = [&](float cell) -> bool {  // fills M.{cell,org_d,dim,ncell,k_per_cube,cube_lo}
    if (M.global) {
      int k = (int)floor(50.0 / (double)cell + 0.5);
      if (k < 1) k = 1;
      if (k >= 2 * kCoarseFactor) k = (k / kCoarseFactor) * kCoarseFactor;  // cube blocks stay aligned on the coarse level
      M.k_per_cube = k;
      M.cell = (float)(50.0 / k);
      int lo[3], hi[3];
      // cube range of the bounding box corners (cube_of is monotone per axis)
      int a0, a1, a2, b0, b1, b2;
      cube_of(mn[0], mn[1], mn[2], M.cen, a0, a1, a2);
      cube_of(mx[0], mx[1], mx[2], M.cen, b0, b1, b2);
      lo[0] = a0; lo[1] = a1; lo[2] = a2; hi[0] = b0; hi[1] = b1; hi[2] = b2;
      const int lim[3] = {kCubeD, kCubeW, kCubeH};
      for (int a = 0; a < 3; a++) {
        if (lo[a] < 0) lo[a] = 0;
        if (hi[a] > lim[a] - 1) hi[a] = lim[a] - 1;
        if (hi[a] < lo[a]) hi[a] = lo[a];
        M.cube_lo[a] = lo[a];
        M.dim[a] = (hi[a] - lo[a] + 1) * k;
      }
      // lower corner of cube index c along x is -25 + 50*(c - cenDepth); y uses cenWidth, z cenHeight
      M.org_d[0] = -25.0 + 50.0 * (lo[0] - M.cen[2]);
      M.org_d[1] = -25.0 + 50.0 * (lo[1] - M.cen[0]);
      M.org_d[2] = -25.0 + 50.0 * (lo[2] - M.cen[1]);
    } else {
      M.k_per_cube = 0;
      M.cell = cell;
      for (int a = 0; a < 3; a++) {
        M.cube_lo[a] = 0;
        M.org_d[a] = (double)mn[a];
        M.dim[a] = (int)floor(((double)mx[a] - (double)mn[a]) / (double)cell) + 1;
        if (M.dim[a] < 1) M.dim[a] = 1;
      }
    }
    M.ncell = (long long)M.dim[0] * M.dim[1] * M.dim[2];
    return M.ncell <= kMaxCells;
  };

  MML_CUDA(ctx, ctx->tmp_d.reserve(sizeof(int) * (size_t)m));        // cell_of
  MML_CUDA(ctx, M.cube_count.reserve(sizeof(int) * kNumCubes));
  int* cell_of = ctx->tmp_d.as<int>();

  auto count_pass = [&](int* cnt) -> int {
    MML_CUDA(ctx, cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)M.ncell + 1), st));
    MML_CUDA(ctx, cudaMemsetAsync(M.cube_count.p, 0, sizeof(int) * kNumCubes, st));
    GridDev G = grid_dev(M, 0);
    G.valid = 1;
    G.pts = nullptr; G.cell_start = nullptr; G.cube_count = nullptr;
    for (int a = 0; a < 3; a++) { G.org[a] = M.org_d[a]; G.dim[a] = M.dim[a]; G.cen[a] = M.cen[a]; G.cube_lo[a] = M.cube_lo[a]; }
    G.inv_cell = 1.0 / (double)M.cell; G.cell = M.cell; G.m = m; G.global = M.global; G.k_per_cube = M.k_per_cube;
    k_map_count<<<div_up(m, 256), 256, 0, st>>>(pts_d, m, G, cnt, cell_of, M.cube_count.as<int>());
    MML_LAUNCHED(ctx);
    return MML_OK;
  };

  float cell = cell_hint;
  if (!(cell > 0.f) && M.auto_cell > 0.f && (long long)m * 4 >= (long long)M.auto_m * 3 && (long long)m * 3 <= (long long)M.auto_m * 4)
    cell = M.auto_cell;  // same map, a few frames later: the search result does not depend on the cell edge, only its speed
  if (!(cell > 0.f)) {
    // trial resolution from the volume per point, then rescale so that an occupied cell
    // holds ~3 points (feature maps are 1-D / 2-D manifolds: occupancy ~ cell^2)
    double vol = 1.0;
    for (int a = 0; a < 3; a++) vol *= fmax((double)mx[a] - (double)mn[a], 0.5);
    cell = (float)fmin(fmax(cbrt(vol / (double)m) * 1.5, 0.1), 5.0);
    while (!layout(cell)) cell *= 1.5f;
    MML_CUDA(ctx, ctx->tmp_e.reserve(sizeof(int) * ((size_t)M.ncell + 1) + 64));
    MML_CHECK(count_pass(ctx->tmp_e.as<int>()));
    MML_CUDA(ctx, ctx->counters.reserve(256));
    unsigned long long* nz_d = reinterpret_cast<unsigned long long*>(ctx->counters.p);
    MML_CUDA(ctx, cudaMemsetAsync(nz_d, 0, 8, st));
    k_count_nonzero<<<4 * kNumSMs, 256, 0, st>>>(ctx->tmp_e.as<int>(), M.ncell, nz_d);
    MML_LAUNCHED(ctx);
    unsigned long long nz = 0;
    MML_CUDA(ctx, cudaMemcpyAsync(&nz, nz_d, 8, cudaMemcpyDeviceToHost, st));
    MML_CUDA(ctx, cudaStreamSynchronize(st));
    if (nz == 0) nz = 1;
    const double occ = (double)m / (double)nz;
    // points per non-empty cell the edge is sized for: 3 for scan-sized query sets (one warp per query: fewer candidates
    // per dependent round trip), 6 for map-sized sweeps (thread per query: fewer rows of cells to walk; measured optimum
    // of the S4 sweep, profiles/r2_s4_knn_experiments.txt)
    static const double occ_env = getenv("MML_CELL_OCC") ? atof(getenv("MML_CELL_OCC")) : 0.0;
    const double target_occ = occ_env > 0.0 ? occ_env : (m > 400000 ? 6.0 : 3.0);
    cell = (float)fmin(fmax((double)M.cell * sqrt(target_occ / occ), 0.05), 5.0);
    M.auto_cell = cell;
    M.auto_m = m;
  }
  while (!layout(cell)) cell *= 1.25f;

  MML_CUDA(ctx, M.cell_start.reserve(sizeof(int) * ((size_t)M.ncell + 1)));
  MML_CUDA(ctx, ctx->tmp_e.reserve(sizeof(int) * ((size_t)M.ncell + 1) + 64));
  MML_CUDA(ctx, M.pts.reserve(sizeof(float4) * (size_t)m));
  int* cell_start = M.cell_start.as<int>();
  MML_CHECK(count_pass(cell_start));
  MML_CHECK(exclusive_scan_device(ctx, cell_start, nullptr, (int)(M.ncell + 1), nullptr));
  MML_CUDA(ctx, cudaMemsetAsync(ctx->tmp_e.p, 0, sizeof(int) * (size_t)M.ncell, st));
  k_map_scatter<<<div_up(m, 256), 256, 0, st>>>(pts_d, m, cell_of, cell_start, ctx->tmp_e.as<int>(), M.pts.as<float4>());
  MML_LAUNCHED(ctx);
  // ---- coarse level
  M.coarse = 1;
  const bool aligned = (!M.global || (M.k_per_cube % kCoarseFactor == 0)) && !getenv("MML_NO_COARSE");
  if (aligned && (M.dim[0] > kCoarseFactor || M.dim[1] > kCoarseFactor || M.dim[2] > kCoarseFactor)) {
    M.coarse = kCoarseFactor;
    for (int a = 0; a < 3; a++) M.dim2[a] = (M.dim[a] + kCoarseFactor - 1) / kCoarseFactor;
    const long long ncell2 = (long long)M.dim2[0] * M.dim2[1] * M.dim2[2];
    MML_CUDA(ctx, M.cell_start2.reserve(sizeof(int) * ((size_t)ncell2 + 1)));
    MML_CUDA(ctx, M.pts2.reserve(sizeof(float4) * (size_t)m));
    MML_CUDA(ctx, ctx->tmp_b.reserve(sizeof(int) * (size_t)m));
    int* cell_of2 = ctx->tmp_b.as<int>();
    int* cs2 = M.cell_start2.as<int>();
    MML_CUDA(ctx, cudaMemsetAsync(cs2, 0, sizeof(int) * ((size_t)ncell2 + 1), st));
    k_map_coarse_count<<<div_up(m, 256), 256, 0, st>>>(cell_of, m, M.dim[0], M.dim[1], kCoarseFactor, M.dim2[0], M.dim2[1],
                                                       cell_of2, cs2);
    MML_LAUNCHED(ctx);
    MML_CHECK(exclusive_scan_device(ctx, cs2, nullptr, (int)(ncell2 + 1), nullptr));
    MML_CUDA(ctx, cudaMemsetAsync(ctx->tmp_e.p, 0, sizeof(int) * (size_t)ncell2, st));
    k_map_scatter<<<div_up(m, 256), 256, 0, st>>>(pts_d, m, cell_of2, cs2, ctx->tmp_e.as<int>(), M.pts2.as<float4>());
    MML_LAUNCHED(ctx);
  }
  MML_CUDA(ctx, cudaGetLastError());
  M.valid = true;
  return MML_OK;
}

// Launch association of the frame slot's corner (kind 0) or surf (kind 1) queries.
// T_dev / thres_dev / gate non-null: parameters come from the device-side solver state.
// Device-resident copy of the four maps' descriptors (ctx->grid_table): written on the context's stream whenever a
// map has been (re)built or dropped since the last call, so kernels enqueued afterwards see the new geometry.
int mml_grid_table_sync(mml_ctx* ctx) {
  MML_CUDA(ctx, ctx->grid_table.reserve(sizeof(GridDev) * 4));
  MML_CUDA(ctx, ctx->grid_table_pin.reserve(sizeof(GridDev) * 4 * 8));
  if (!ctx->grid_table_dirty) return MML_OK;
  // a ring of pinned staging copies: an earlier asynchronous copy may still be in flight
  GridDev* h = ctx->grid_table_pin.as<GridDev>() + 4 * (ctx->grid_table_gen++ & 7);
  for (int k = 0; k < 4; k++) {
    h[k] = grid_dev(ctx->maps[k], k);
    for (int a = 0; a < 3; a++) h[k].cen[a] = ctx->maps[k].cen[a];  // the cube rule needs the centre even without a map
  }
  MML_CUDA(ctx, cudaMemcpyAsync(ctx->grid_table.p, h, sizeof(GridDev) * 4, cudaMemcpyHostToDevice, ctx->stream));
  ctx->grid_table_dirty = false;
  return MML_OK;
}

int mml_associate_launch(mml_ctx* ctx, int kind, const double* T16, float thres, const double* T_dev,
                         const float* thres_dev, const int* gate, const int* nq_dev, int cap) {
  const int nq = cap;
  // lanes per query: a whole warp for scan-sized query sets (latency), one thread per query for map-sized sweeps
  const int G = cap <= 32768 ? 32 : 1;
  // group kernels stride over the queries: at most kAssocWave CTAs (4 per SM), however large the query buffer is
  const int kAssocWave = 4 * kNumSMs;
  int grid = G == 1 ? div_up(nq > 0 ? nq : 1, 128) : div_up(nq > 0 ? nq : 1, 128 / G);
  if (G != 1 && grid > kAssocWave) grid = kAssocWave;
  mml::DevBuf& fb = kind == 0 ? ctx->f_line : ctx->f_plane;
  MML_CUDA(ctx, fb.reserve(sizeof(float4) * 3 * (size_t)(nq > 0 ? nq : 1)));
  // assoc_stats layout (doubles): [0..7] line moments/count, [8..15] plane moments/count,
  // then ints: n_line, n_plane, tickets
  MML_CUDA(ctx, ctx->assoc_stats.reserve(512));
  MML_CUDA(ctx, ctx->assoc_part[kind].reserve(sizeof(double) * 8 * (size_t)grid + 64));
  AssocArgs A;
  memset(&A, 0, sizeof(A));
  A.q = (kind == 0 ? ctx->q_corner : ctx->q_surf).as<float4>();
  A.nq = nq;
  A.nq_dev = nq_dev;
  if (T16) for (int k = 0; k < 16; k++) A.T[k] = T16[k];
  A.thres = thres;
  A.G[0] = grid_dev(ctx->maps[kind == 0 ? MML_MAP_CORNER_GLOBAL : MML_MAP_SURF_GLOBAL], kind == 0 ? 0 : 1);
  A.G[1] = grid_dev(ctx->maps[kind == 0 ? MML_MAP_CORNER_LOCAL : MML_MAP_SURF_LOCAL], kind == 0 ? 2 : 3);
  if (!A.G[0].valid) { A.G[0].cen[0] = ctx->maps[kind].cen[0]; A.G[0].cen[1] = ctx->maps[kind].cen[1]; A.G[0].cen[2] = ctx->maps[kind].cen[2]; }
  A.feat = fb.as<float4>();
  double* stats = ctx->assoc_stats.as<double>();
  A.moment_out = stats + 8 * kind;
  int* ints = reinterpret_cast<int*>(stats + 16);
  A.n_feat_out = ints + kind;
  A.ticket = reinterpret_cast<unsigned*>(ints + 4 + kind);
  A.moment_partials = ctx->assoc_part[kind].as<double>();
  A.gate = gate;
  A.T_dev = T_dev;
  A.thres_dev = thres_dev;
  A.tl = ctx->timeline.as<unsigned long long>();
  A.tl_slot = kind == 1 ? 2 : 4;
  A.overflow = ints + 6;
  A.perm = ctx->has_perm[kind] ? ctx->perm[kind].as<unsigned>() : nullptr;
  // map-sized, spatially sorted sets: search kernel, then fit kernel (MML_ASSOC_SPLIT=0: the fused kernel)
  static const int split_env = getenv("MML_ASSOC_SPLIT") ? atoi(getenv("MML_ASSOC_SPLIT")) : 1;
  if (G == 1 && split_env && ctx->has_perm[kind] && nq > 0) {
    mml::DevBuf& pb = ctx->pre_knn[kind];  // [nq][5] positions + [nq] status
    MML_CUDA(ctx, pb.reserve(sizeof(int) * 6 * (size_t)nq + 64));
    A.pre_loc = pb.as<int>();
    A.pre_status = pb.as<int>() + 5 * (size_t)nq;
    A.pre_map = A.G[0].valid ? 0 : 1;
    if (A.G[A.pre_map].valid) {
      // search radii spanning many fine cells (the first outer iterations of a scan-to-map Estimate: thres 25, 10):
      // the two-level walk, which sends displaced queries to the coarse cells after two fine shells
      const GridDev& Gw = A.G[A.pre_map];
      const bool two = Gw.pts2 != nullptr && (int)ceilf(sqrtf(thres) / Gw.cell) + 1 >= MML_TWO_LEVEL_MIN_SHELLS;
      // (a radius that lives on the device - the solve loop's schedule - is only known to the kernels: both are
      // launched and the one that does not apply returns at once)
      if (thres_dev || two) { k_knn_walk<true><<<grid, 128, 0, ctx->stream>>>(A); MML_LAUNCHED(ctx); }
      if (thres_dev || !two) { k_knn_walk<false><<<grid, 128, 0, ctx->stream>>>(A); MML_LAUNCHED(ctx); }
      if (kind == 0) k_associate<0, true><<<grid, 128, 0, ctx->stream>>>(A);
      else k_associate<1, true><<<grid, 128, 0, ctx->stream>>>(A);
      MML_LAUNCHED(ctx);
      MML_CUDA(ctx, cudaGetLastError());
      return MML_OK;
    }
  }
  if (G == 32 && ctx->assoc_table_mode) {
    if (!ctx->grid_table.p) return mml_fail(ctx, MML_ERR_STATE, "map descriptor table not initialised");
    const GridDev* tab = ctx->grid_table.as<GridDev>();
    A.G_tab[0] = tab + (kind == 0 ? MML_MAP_CORNER_GLOBAL : MML_MAP_SURF_GLOBAL);
    A.G_tab[1] = tab + (kind == 0 ? MML_MAP_CORNER_LOCAL : MML_MAP_SURF_LOCAL);
    if (kind == 0) k_associate_g<0, 32, true><<<grid, 128, 0, ctx->stream>>>(A);
    else k_associate_g<1, 32, true><<<grid, 128, 0, ctx->stream>>>(A);
  } else if (G == 32) {
    if (kind == 0) k_associate_g<0, 32><<<grid, 128, 0, ctx->stream>>>(A);
    else k_associate_g<1, 32><<<grid, 128, 0, ctx->stream>>>(A);
  } else {
    if (kind == 0) k_associate<0, false><<<grid, 128, 0, ctx->stream>>>(A);
    else k_associate<1, false><<<grid, 128, 0, ctx->stream>>>(A);
  }
  MML_LAUNCHED(ctx);
  MML_CUDA(ctx, cudaGetLastError());
  return MML_OK;
}

int mml_export_features(mml_ctx* ctx, int kind, int nq, double* out_dev) {
  if (nq <= 0) return MML_OK;
  const float4* f = (kind == 0 ? ctx->f_line : ctx->f_plane).as<float4>();
  if (kind == 0) k_export_features<0><<<div_up(nq, 256), 256, 0, ctx->stream>>>(f, nq, out_dev);
  else k_export_features<1><<<div_up(nq, 256), 256, 0, ctx->stream>>>(f, nq, out_dev);
  MML_LAUNCHED(ctx);
  MML_CUDA(ctx, cudaGetLastError());
  return MML_OK;
}

extern "C" int mml_debug_set_fine_shells(int v) {
  return cudaMemcpyToSymbol(mml::g_fine_shells, &v, sizeof(int)) == cudaSuccess ? 0 : -3;
}
