// ORACLE / reference pin (test infrastructure only).
// Stand-ins for the PCL 1.8 types the reference's hot-path text touches. Point layouts carry the
// fields the reference reads; containers behave like std::vector. Third-party ALGORITHMS are
// restated from their published behaviour (same assumptions as oracle/oracle.h):
//   KdTreeFLANN::nearestKSearch  exact k-NN, L2_Simple<float> distance ((dx²)+dy²)+dz², ascending,
//                                ties to the lower index (unspecified in FLANN)
//   VoxelGrid::filter            PCL 1.8 voxel_grid.hpp: voxel = floor(p/leaf) - min, points of a
//                                voxel summed in input order (unspecified after PCL's std::sort),
//                                float32 centroid of x, y, z, intensity; output in voxel order
#ifndef MML_REF_PCL_H
#define MML_REF_PCL_H
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <memory>
#include <vector>

#define pcl_isfinite(x) std::isfinite(x)

namespace pcl {
struct PointXYZI { float x = 0, y = 0, z = 0, intensity = 0; };
struct PointXYZINormal {
  float x = 0, y = 0, z = 0, normal_x = 0, normal_y = 0, normal_z = 0, intensity = 0, curvature = 0;
};
template <class P> struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud<P>>;
  using ConstPtr = std::shared_ptr<const PointCloud<P>>;
  std::vector<P> points;
  uint32_t width = 0, height = 1;
  bool is_dense = true;
  void push_back(const P& p) { points.push_back(p); width = (uint32_t)points.size(); }
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = 0; }
  void reserve(size_t n) { points.reserve(n); }
  P& operator[](size_t i) { return points[i]; }
  const P& operator[](size_t i) const { return points[i]; }
  typename std::vector<P>::iterator begin() { return points.begin(); }
  typename std::vector<P>::iterator end() { return points.end(); }
  typename std::vector<P>::const_iterator begin() const { return points.begin(); }
  typename std::vector<P>::const_iterator end() const { return points.end(); }
  PointCloud& operator+=(const PointCloud& o) { points.insert(points.end(), o.points.begin(), o.points.end()); width = (uint32_t)points.size(); return *this; }
};

template <class P> class KdTreeFLANN {
  struct Node { int lo, hi, left, right, dim; float split; float bmin[3], bmax[3]; };
  std::vector<Node> nodes_;
  std::vector<int> idx_;
  std::shared_ptr<const PointCloud<P>> cloud_;
  static float coord(const P& p, int d) { return d == 0 ? p.x : d == 1 ? p.y : p.z; }
  int build(int lo, int hi) {
    Node nd; nd.lo = lo; nd.hi = hi; nd.left = nd.right = -1; nd.dim = 0; nd.split = 0;
    for (int d = 0; d < 3; d++) { nd.bmin[d] = 1e30f; nd.bmax[d] = -1e30f; }
    for (int i = lo; i < hi; i++) for (int d = 0; d < 3; d++) { float c = coord(cloud_->points[idx_[i]], d); nd.bmin[d] = std::min(nd.bmin[d], c); nd.bmax[d] = std::max(nd.bmax[d], c); }
    int me = (int)nodes_.size(); nodes_.push_back(nd);
    if (hi - lo > 16) {
      int dim = 0; float ext = -1;
      for (int d = 0; d < 3; d++) if (nd.bmax[d] - nd.bmin[d] > ext) { ext = nd.bmax[d] - nd.bmin[d]; dim = d; }
      int mid = (lo + hi) / 2;
      std::nth_element(idx_.begin() + lo, idx_.begin() + mid, idx_.begin() + hi, [&](int a, int b) {
        float ca = coord(cloud_->points[a], dim), cb = coord(cloud_->points[b], dim); return ca < cb || (ca == cb && a < b); });
      int l = build(lo, mid), r = build(mid, hi);
      nodes_[me].left = l; nodes_[me].right = r; nodes_[me].dim = dim;
    }
    return me;
  }
  struct Cand { float d; int i; };
  static bool closer(const Cand& a, const Cand& b) { return a.d < b.d || (a.d == b.d && a.i < b.i); }
  void search(int n, const float q[3], int k, std::vector<Cand>& best) const {
    const Node& nd = nodes_[n];
    if ((int)best.size() == k) {
      double lb = 0;  // conservative lower bound (double, shrunk) so exact float ties are never pruned
      for (int d = 0; d < 3; d++) { double e = 0; if (q[d] < nd.bmin[d]) e = (double)nd.bmin[d] - q[d]; else if (q[d] > nd.bmax[d]) e = (double)q[d] - nd.bmax[d]; lb += e * e; }
      if (lb * (1.0 - 1e-5) > (double)best.back().d) return;
    }
    if (nd.left < 0) {
      for (int i = nd.lo; i < nd.hi; i++) {
        const P& p = cloud_->points[idx_[i]];
        float dx = p.x - q[0], dy = p.y - q[1], dz = p.z - q[2];
        float d = 0; d += dx * dx; d += dy * dy; d += dz * dz;
        Cand c{d, idx_[i]};
        if ((int)best.size() < k) { best.push_back(c); std::sort(best.begin(), best.end(), closer); }
        else if (closer(c, best.back())) { best.back() = c; std::sort(best.begin(), best.end(), closer); }
      }
      return;
    }
    int first = nd.left, second = nd.right;
    const Node& L = nodes_[nd.left];
    if (q[nd.dim] > L.bmax[nd.dim]) std::swap(first, second);
    search(first, q, k, best); search(second, q, k, best);
  }
 public:
  using Ptr = std::shared_ptr<KdTreeFLANN<P>>;
  void setInputCloud(const std::shared_ptr<const PointCloud<P>>& c) {
    cloud_ = c; nodes_.clear(); idx_.resize(c->points.size());
    for (size_t i = 0; i < idx_.size(); i++) idx_[i] = (int)i;
    if (!idx_.empty()) build(0, (int)idx_.size());
  }
  void setInputCloud(const std::shared_ptr<PointCloud<P>>& c) { setInputCloud(std::shared_ptr<const PointCloud<P>>(c)); }
  std::shared_ptr<const PointCloud<P>> getInputCloud() const { return cloud_; }
  int nearestKSearch(const P& p, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
    std::vector<Cand> best; best.reserve(k + 1);
    float q[3] = {p.x, p.y, p.z};
    if (!nodes_.empty()) search(0, q, k, best);
    k_indices.resize(k); k_sqr_distances.resize(k);  // PCL resizes to k before the search
    for (size_t i = 0; i < best.size(); i++) { k_indices[i] = best[i].i; k_sqr_distances[i] = best[i].d; }
    return (int)best.size();
  }
};

template <class P> class VoxelGrid {
  float leaf_[3] = {0, 0, 0};
  std::shared_ptr<const PointCloud<P>> in_;
 public:
  void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
  void setInputCloud(const std::shared_ptr<PointCloud<P>>& c) { in_ = c; }
  void filter(PointCloud<P>& out) {
    const std::vector<P>& pts = in_->points;
    PointCloud<P> res;
    if (pts.empty()) { out = res; return; }
    float inv[3] = {1.0f / leaf_[0], 1.0f / leaf_[1], 1.0f / leaf_[2]};
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (const P& p : pts) {
      if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
      mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
      mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
    }
    int minb[3], maxb[3], divb[3];
    for (int d = 0; d < 3; d++) { minb[d] = (int)std::floor(mn[d] * inv[d]); maxb[d] = (int)std::floor(mx[d] * inv[d]); divb[d] = maxb[d] - minb[d] + 1; }
    int mul[3] = {1, divb[0], divb[0] * divb[1]};
    std::vector<std::pair<unsigned, unsigned>> iv; iv.reserve(pts.size());
    for (unsigned i = 0; i < pts.size(); i++) {
      const P& p = pts[i];
      if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
      int i0 = (int)(std::floor(p.x * inv[0]) - (float)minb[0]);
      int i1 = (int)(std::floor(p.y * inv[1]) - (float)minb[1]);
      int i2 = (int)(std::floor(p.z * inv[2]) - (float)minb[2]);
      iv.emplace_back((unsigned)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), i);
    }
    std::stable_sort(iv.begin(), iv.end(), [](const std::pair<unsigned, unsigned>& a, const std::pair<unsigned, unsigned>& b) { return a.first < b.first; });
    size_t s = 0;
    while (s < iv.size()) {
      size_t e = s + 1;
      while (e < iv.size() && iv[e].first == iv[s].first) e++;
      float sx = 0, sy = 0, sz = 0, si = 0;
      for (size_t k = s; k < e; k++) { const P& p = pts[iv[k].second]; sx += p.x; sy += p.y; sz += p.z; si += p.intensity; }
      float n = (float)(e - s);
      P c; c.x = sx / n; c.y = sy / n; c.z = sz / n; c.intensity = si / n;
      res.push_back(c);
      s = e;
    }
    out = res;
  }
};
}  // namespace pcl
#endif
