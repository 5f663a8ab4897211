// ORACLE (test infrastructure only — see oracle.h).
//   orc_associate_line()  <- Estimator::processPointToLine     src/lio/Estimator.cpp:148-365
//   orc_associate_plane() <- Estimator::processPointToPlanVec  src/lio/Estimator.cpp:573-777
//   orc_localizability()  <- Estimator::checkLocalizability    src/lio/Estimator.cpp:536-565
//   orc_map_*             <- the per-cube clouds + kd-trees copied at EST.cpp:1159-1182 and
//                            binned with the cube rule of src/lio/Map_Manager.cpp:159-175
// pcl::KdTreeFLANN::nearestKSearch is restated as an exact k-NN over
// d2 = ((dx*dx)+dy*dy)+dz*dz in float32 (FLANN L2_Simple<float>), results ascending;
// ties, which FLANN leaves unspecified, are broken by the lower point index.
#include "oracle.h"
#include "oracle_math.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <numeric>
#include <thread>
#include <vector>

using namespace orc;

namespace {

struct P4 { float x, y, z, i; };

inline float dist2(const P4& a, float qx, float qy, float qz) {
  float dx = qx - a.x, dy = qy - a.y, dz = qz - a.z;
  return (dx * dx + dy * dy) + dz * dz;
}

struct Knn5 {
  float d[5];
  int id[5];
  int cnt = 0;
  Knn5() {
    for (int k = 0; k < 5; k++) { d[k] = INFINITY; id[k] = -1; }
  }
  inline bool better(float dd, int ii, int slot) const {
    return dd < d[slot] || (dd == d[slot] && ii < id[slot]);
  }
  inline void push(float dd, int ii) {
    if (!(cnt < 5 || better(dd, ii, 4))) return;
    int k = cnt < 5 ? cnt : 4;
    if (cnt < 5) cnt++;
    while (k > 0 && better(dd, ii, k - 1)) {
      d[k] = d[k - 1];
      id[k] = id[k - 1];
      k--;
    }
    d[k] = dd;
    id[k] = ii;
  }
  inline float worst() const { return cnt < 5 ? INFINITY : d[4]; }
};

// Exact kd-tree (median split on the widest axis, leaves of <= 15 points like PCL's default
// FLANN KDTreeSingleIndex leaf_max_size). Pruning uses the float plane distance with
// "<=" so equal-distance candidates with a lower index are never missed.
struct KdTree {
  std::vector<P4> pts;
  std::vector<int> perm;
  struct Node { int lo, hi, axis, left, right; float split; };
  std::vector<Node> nodes;
  void build(const P4* p, int m) {
    pts.assign(p, p + m);
    perm.resize(m);
    std::iota(perm.begin(), perm.end(), 0);
    nodes.clear();
    if (m > 0) rec(0, m);
  }
  int rec(int lo, int hi) {
    int id = (int)nodes.size();
    nodes.push_back({lo, hi, -1, -1, -1, 0.f});
    if (hi - lo <= 15) return id;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = lo; k < hi; k++) {
      const P4& q = pts[perm[k]];
      float v[3] = {q.x, q.y, q.z};
      for (int c = 0; c < 3; c++) { mn[c] = std::min(mn[c], v[c]); mx[c] = std::max(mx[c], v[c]); }
    }
    int ax = 0;
    if (mx[1] - mn[1] > mx[ax] - mn[ax]) ax = 1;
    if (mx[2] - mn[2] > mx[ax] - mn[ax]) ax = 2;
    if (!(mx[ax] > mn[ax])) return id;  // all coincident: keep as leaf
    int mid = (lo + hi) / 2;
    auto key = [&](int i) { return ax == 0 ? pts[i].x : (ax == 1 ? pts[i].y : pts[i].z); };
    std::nth_element(perm.begin() + lo, perm.begin() + mid, perm.begin() + hi,
                     [&](int a, int b) { return key(a) < key(b); });
    float split = key(perm[mid]);
    int l = rec(lo, mid);
    int r = rec(mid, hi);
    nodes[id].axis = ax;
    nodes[id].split = split;
    nodes[id].left = l;
    nodes[id].right = r;
    return id;
  }
  void search(int nid, float qx, float qy, float qz, Knn5& res) const {
    const Node& nd = nodes[nid];
    if (nd.axis < 0) {
      for (int k = nd.lo; k < nd.hi; k++) res.push(dist2(pts[perm[k]], qx, qy, qz), perm[k]);
      return;
    }
    float qv = nd.axis == 0 ? qx : (nd.axis == 1 ? qy : qz);
    float diff = qv - nd.split;
    int near = diff < 0 ? nd.left : nd.right;
    int far = diff < 0 ? nd.right : nd.left;
    search(near, qx, qy, qz, res);
    if (diff * diff <= res.worst()) search(far, qx, qy, qz, res);
  }
  bool knn5(float qx, float qy, float qz, int* idx, float* d2) const {
    Knn5 r;
    if (!nodes.empty()) search(0, qx, qy, qz, r);
    for (int k = 0; k < 5; k++) { idx[k] = r.id[k]; d2[k] = r.d[k]; }
    return r.cnt == 5;
  }
};

const int kNumCubes = 21 * 11 * 21;  // MM.h:117-120

}  // namespace

struct orc_map {
  // global maps: per-cube trees (index = MM.cpp:64-66 ToIndex); local maps: one tree each
  std::vector<std::unique_ptr<KdTree>> cube[2];
  KdTree local[2];
  int cen[3] = {10, 5, 10};  // (CenWidth, CenHeight, CenDepth) MM.h:109-111
  orc_map() {
    cube[0].resize(kNumCubes);
    cube[1].resize(kNumCubes);
  }
};

extern "C" {

orc_map* orc_map_create(void) { return new orc_map(); }
void orc_map_destroy(orc_map* m) { delete m; }

int orc_map_set(orc_map* m, int kind, const float* xyzi, int n, const int* cen3) {
  const P4* p = reinterpret_cast<const P4*>(xyzi);
  if (kind == 2 || kind == 3) {
    m->local[kind - 2].build(p, n);
    return 0;
  }
  if (kind != 0 && kind != 1) return -1;
  if (cen3) { m->cen[0] = cen3[0]; m->cen[1] = cen3[1]; m->cen[2] = cen3[2]; }
  std::vector<std::vector<P4>> bins(kNumCubes);
  for (int i = 0; i < n; i++) {
    // MM.cpp:159-175: same cube rule as FindUsed*Map, out-of-grid points are dropped
    int id = orc_cube_index(&p[i].x, m->cen[0], m->cen[1], m->cen[2]);
    if (id == 5000) continue;
    bins[id].push_back(p[i]);
  }
  for (int c = 0; c < kNumCubes; c++) {
    if (bins[c].empty()) {
      m->cube[kind][c].reset();
    } else {
      m->cube[kind][c].reset(new KdTree());
      m->cube[kind][c]->build(bins[c].data(), (int)bins[c].size());
    }
  }
  return 0;
}

int orc_knn5_brute(const float* cloud, int m, const float* q, int* idx5, float* d2_5) {
  const P4* p = reinterpret_cast<const P4*>(cloud);
  Knn5 r;
  for (int i = 0; i < m; i++) r.push(dist2(p[i], q[0], q[1], q[2]), i);
  for (int k = 0; k < 5; k++) { idx5[k] = r.id[k]; d2_5[k] = r.d[k]; }
  return r.cnt;
}

int orc_knn5_kdtree(const float* cloud, int m, const float* q_xyzi, int nq, int* idx5, float* d2_5) {
  KdTree t;
  t.build(reinterpret_cast<const P4*>(cloud), m);
  for (int i = 0; i < nq; i++)
    t.knn5(q_xyzi[4 * i], q_xyzi[4 * i + 1], q_xyzi[4 * i + 2], idx5 + 5 * i, d2_5 + 5 * i);
  return 0;
}

}  // extern "C"

namespace {

// EST.cpp:204-277: float32 mean/covariance of the 5 neighbours, double eigen-solve,
// accept if lambda2 > 3*lambda1, end points = mean +- 0.1*v2 rounded to float32.
bool fit_line(const KdTree& t, const int* idx, double* p1, double* p2) {
  float cx = 0, cy = 0, cz = 0;
  for (int j = 0; j < 5; j++) { cx += t.pts[idx[j]].x; cy += t.pts[idx[j]].y; cz += t.pts[idx[j]].z; }
  cx /= 5; cy /= 5; cz /= 5;
  float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
  for (int j = 0; j < 5; j++) {
    float ax = t.pts[idx[j]].x - cx, ay = t.pts[idx[j]].y - cy, az = t.pts[idx[j]].z - cz;
    a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
    a22 += ay * ay; a23 += ay * az; a33 += az * az;
  }
  a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
  double A[9] = {a11, a12, a13, a12, a22, a23, a13, a23, a33};
  double ev[3], V[9];
  eig3_sym(A, ev, V);
  if (!(ev[2] > 3 * ev[1])) return false;
  double u[3] = {V[2], V[5], V[8]};
  float x1 = cx + 0.1 * u[0], y1 = cy + 0.1 * u[1], z1 = cz + 0.1 * u[2];
  float x2 = cx - 0.1 * u[0], y2 = cy - 0.1 * u[1], z2 = cz - 0.1 * u[2];
  p1[0] = x1; p1[1] = y1; p1[2] = z1;
  p2[0] = x2; p2[1] = y2; p2[2] = z2;
  return true;
}

// Estimator.h:71-83 FeatureLine::ComputeError
double line_error(const double* po, const double* a, const double* b, const double* T) {
  double P[3];
  for (int r = 0; r < 3; r++) P[r] = ((T[4 * r] * po[0] + T[4 * r + 1] * po[1]) + T[4 * r + 2] * po[2]) + T[4 * r + 3];
  double l12 = std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
  double c0 = (P[0] - a[0]) * (P[1] - b[1]) - (P[0] - b[0]) * (P[1] - a[1]);
  double c1 = (P[0] - a[0]) * (P[2] - b[2]) - (P[0] - b[0]) * (P[2] - a[2]);
  double c2 = (P[1] - a[1]) * (P[2] - b[2]) - (P[1] - b[1]) * (P[2] - a[2]);
  return std::sqrt(c0 * c0 + c1 * c1 + c2 * c2) / l12;
}

// EST.cpp:634-695: plane through the 5 neighbours, float32 normalisation and validity
// test, projection point in double.
bool fit_plane(const KdTree& t, const int* idx, const P4& sel, double* proj, double* nrm) {
  double A[5][3], b[5];
  for (int j = 0; j < 5; j++) {
    A[j][0] = t.pts[idx[j]].x; A[j][1] = t.pts[idx[j]].y; A[j][2] = t.pts[idx[j]].z;
    b[j] = -1.0;
  }
  double X[3];
  qr5x3_solve(A, b, X);
  float pa = X[0], pb = X[1], pc = X[2], pd = 1;
  float ps = std::sqrt(pa * pa + pb * pb + pc * pc);
  pa /= ps; pb /= ps; pc /= ps; pd /= ps;
  for (int j = 0; j < 5; j++) {
    if (std::fabs(pa * t.pts[idx[j]].x + pb * t.pts[idx[j]].y + pc * t.pts[idx[j]].z + pd) > 0.2) return false;
  }
  double dist = pa * sel.x + pb * sel.y + pc * sel.z + pd;
  nrm[0] = pa; nrm[1] = pb; nrm[2] = pc;
  proj[0] = (double)sel.x - dist * nrm[0];
  proj[1] = (double)sel.y - dist * nrm[1];
  proj[2] = (double)sel.z - dist * nrm[2];
  return true;
}

// Estimator.h:118-121 FeaturePlanVec::ComputeError
double plane_error(const double* po, const double* proj, const double* T) {
  double e[3];
  for (int r = 0; r < 3; r++) {
    double P = ((T[4 * r] * po[0] + T[4 * r + 1] * po[1]) + T[4 * r + 2] * po[2]) + T[4 * r + 3];
    e[r] = P - proj[r];
  }
  return std::sqrt((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]);
}

inline void write_feat(double* f, const P4& ori, const double* a, const double* b, double err, int src) {
  f[0] = ori.x; f[1] = ori.y; f[2] = ori.z;
  f[3] = a[0]; f[4] = a[1]; f[5] = a[2];
  f[6] = b[0]; f[7] = b[1]; f[8] = b[2];
  f[9] = err;
  f[10] = (std::fabs(err) > 1e-5) ? 1.0 : 0.0;  // EST.cpp:1313,1385
  f[11] = (double)src;
}

}  // namespace

extern "C" {

static int assoc_line_range(const orc_map* m, const P4* q, int i0, int i1, const double* T, double thres_dist,
                            double* feat) {
  int nf = 0;
  for (int i = i0; i < i1; i++) {
    double* f = feat + 12 * i;
    for (int k = 0; k < 12; k++) f[k] = 0;
    f[10] = -1.0;
    f[11] = (double)i;
    P4 sel = q[i];
    orc_point_to_map(&q[i].x, T, &sel.x);  // EST.cpp:191
    int id = orc_cube_index(&sel.x, m->cen[0], m->cen[1], m->cen[2]);  // EST.cpp:192
    if (id == 5000) continue;
    if (std::isnan(sel.x) || std::isnan(sel.y) || std::isnan(sel.z)) continue;
    int idx[5];
    float d2[5];
    double p1[3], p2[3];
    const KdTree* g = m->cube[0][id].get();
    if (g && g->pts.size() > 100) {  // EST.cpp:198
      g->knn5(sel.x, sel.y, sel.z, idx, d2);
      if (d2[4] < thres_dist && fit_line(*g, idx, p1, p2)) {  // EST.cpp:201,254
        double po[3] = {q[i].x, q[i].y, q[i].z};
        write_feat(f, q[i], p1, p2, line_error(po, p1, p2, T), i);
        nf++;
        continue;
      }
    }
    const KdTree& L = m->local[0];
    if (L.pts.size() > 20) {  // EST.cpp:283
      L.knn5(sel.x, sel.y, sel.z, idx, d2);
      if (d2[4] < thres_dist && fit_line(L, idx, p1, p2)) {
        double po[3] = {q[i].x, q[i].y, q[i].z};
        write_feat(f, q[i], p1, p2, line_error(po, p1, p2, T), i);
        nf++;
      }
    }
  }
  return nf;
}

// threads > 1: queries are split into contiguous ranges (the reference runs line and plane
// association on two threads, EST.cpp:1271-1297; the "all cores" baseline splits the queries)
int orc_associate_line_mt(const orc_map* m, const float* q_xyzi, int nq, const double* T, double thres_dist,
                          double* feat, int* n_feat, int threads) {
  const P4* q = reinterpret_cast<const P4*>(q_xyzi);
  if (threads <= 1 || nq < 64) {
    *n_feat = assoc_line_range(m, q, 0, nq, T, thres_dist, feat);
    return 0;
  }
  std::vector<int> part(threads, 0);
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++)
    th.emplace_back([&, t]() {
      part[t] = assoc_line_range(m, q, (int)((long)nq * t / threads), (int)((long)nq * (t + 1) / threads), T, thres_dist, feat);
    });
  for (auto& t : th) t.join();
  int nf = 0;
  for (int v : part) nf += v;
  *n_feat = nf;
  return 0;
}

int orc_associate_line(const orc_map* m, const float* q_xyzi, int nq, const double* T,
                       double thres_dist, double* feat, int* n_feat) {
  return orc_associate_line_mt(m, q_xyzi, nq, T, thres_dist, feat, n_feat, 1);
}

static int assoc_plane_range(const orc_map* m, const P4* q, int i0, int i1, const double* T, double thres_dist,
                             double* feat, double* M) {
  int nf = 0;
  for (int k = 0; k < 9; k++) M[k] = 0;
  for (int i = i0; i < i1; i++) {
    double* f = feat + 12 * i;
    for (int k = 0; k < 12; k++) f[k] = 0;
    f[10] = -1.0;
    f[11] = (double)i;
    P4 sel = q[i];
    orc_point_to_map(&q[i].x, T, &sel.x);  // EST.cpp:619
    int id = orc_cube_index(&sel.x, m->cen[0], m->cen[1], m->cen[2]);  // EST.cpp:621
    if (id == 5000) continue;
    if (std::isnan(sel.x) || std::isnan(sel.y) || std::isnan(sel.z)) continue;
    int idx[5];
    float d2[5];
    double proj[3], nrm[3];
    bool done = false;
    const KdTree* g = m->cube[1][id].get();
    if (g && g->pts.size() > 50) {  // EST.cpp:627
      g->knn5(sel.x, sel.y, sel.z, idx, d2);
      if (d2[4] < thres_dist && fit_plane(*g, idx, sel, proj, nrm)) done = true;  // EST.cpp:631,667
    }
    if (!done) {
      const KdTree& L = m->local[1];
      if (L.pts.size() > 20) {  // EST.cpp:702
        L.knn5(sel.x, sel.y, sel.z, idx, d2);
        if (d2[4] < thres_dist && fit_plane(L, idx, sel, proj, nrm)) done = true;
      }
    }
    if (!done) continue;
    double po[3] = {q[i].x, q[i].y, q[i].z};
    write_feat(f, q[i], proj, nrm, plane_error(po, proj, T), i);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) M[3 * r + c] += nrm[r] * nrm[c];
    nf++;
  }
  return nf;
}

int orc_associate_plane_mt(const orc_map* m, const float* q_xyzi, int nq, const double* T, double thres_dist,
                           double* feat, int* n_feat, double* M9, int* n_normals, int threads) {
  const P4* q = reinterpret_cast<const P4*>(q_xyzi);
  double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  int nf = 0;
  if (threads <= 1 || nq < 64) {
    nf = assoc_plane_range(m, q, 0, nq, T, thres_dist, feat, M);
  } else {
    std::vector<int> part(threads, 0);
    std::vector<double> Mp(9 * (size_t)threads, 0.0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
      th.emplace_back([&, t]() {
        part[t] = assoc_plane_range(m, q, (int)((long)nq * t / threads), (int)((long)nq * (t + 1) / threads), T, thres_dist,
                                    feat, &Mp[9 * (size_t)t]);
      });
    for (auto& t : th) t.join();
    for (int t = 0; t < threads; t++) {
      nf += part[t];
      for (int k = 0; k < 9; k++) M[k] += Mp[9 * (size_t)t + k];
    }
  }
  *n_feat = nf;
  if (M9) std::memcpy(M9, M, sizeof(M));
  if (n_normals) *n_normals = nf;
  return 0;
}

int orc_associate_plane(const orc_map* m, const float* q_xyzi, int nq, const double* T,
                        double thres_dist, double* feat, int* n_feat, double* M9, int* n_normals) {
  return orc_associate_plane_mt(m, q_xyzi, nq, T, thres_dist, feat, n_feat, M9, n_normals, 1);
}

// EST.cpp:536-565. Singular values of the stacked normals = sqrt(eig(sum n n^T)).
double orc_localizability(const double* M9, int n_normals) {
  if (!(n_normals > 10)) return -1.0;
  double ev[3], V[9];
  eig3_sym(M9, ev, V);
  return std::sqrt(std::max(ev[0], 0.0));
}

}  // extern "C"
