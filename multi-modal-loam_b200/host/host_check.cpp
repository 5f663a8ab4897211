// Compile-and-link check of the C++ adapters (build()), and a small GPU run (tests, -m gpu):
//   host_check <n_az>       extracts features of a synthetic ring through
//                           LidarFeatureExtractor::detectFeaturePoint and prints the index lists;
//   host_check <n_az> map   additionally runs Estimator::MapIncrementLocal on the ring;
//   host_check est <dir>    instantiates the whole Estimator adapter on inputs the test wrote to <dir> (float32 /
//                           float64 binaries): setMap, processPointToLine, processPointToPlanVec, EstimateLidarPose,
//                           and prints what they return for the test to compare with the oracle.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "mmloam_shim.hpp"

template <class T>
static std::vector<T> read_bin(const std::string& path) {
  std::vector<T> v;
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open " + path);
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  v.resize((size_t)n / sizeof(T));
  if (n && std::fread(v.data(), 1, (size_t)n, f) != (size_t)n) { std::fclose(f); throw std::runtime_error("short read " + path); }
  std::fclose(f);
  return v;
}
static std::vector<mmloam::PointXYZINormal> cloud_from(const std::vector<float>& a, int stride) {
  std::vector<mmloam::PointXYZINormal> c(a.size() / stride);
  for (size_t i = 0; i < c.size(); i++) {
    c[i] = {};
    c[i].x = a[stride * i]; c[i].y = a[stride * i + 1]; c[i].z = a[stride * i + 2]; c[i].intensity = a[stride * i + 3];
    if (stride >= 7) { c[i].normal_x = a[stride * i + 4]; c[i].normal_y = a[stride * i + 5]; c[i].normal_z = a[stride * i + 6]; }
  }
  return c;
}

static int run_est(const std::string& dir) {
  mmloam::Context ctx(0);
  mmloam::Estimator est(ctx, 0.4f, 0.2f);
  est.setMap(MML_MAP_SURF_LOCAL, cloud_from(read_bin<float>(dir + "/map_surf.bin"), 4));
  est.setMap(MML_MAP_CORNER_LOCAL, cloud_from(read_bin<float>(dir + "/map_corner.bin"), 4));
  const auto corner = cloud_from(read_bin<float>(dir + "/corner.bin"), 4);
  const auto surf = cloud_from(read_bin<float>(dir + "/surf.bin"), 4);
  const auto scan = cloud_from(read_bin<float>(dir + "/scan7.bin"), 7);
  const auto pose = read_bin<double>(dir + "/pose.bin");  // P3 q4 T_wl16 exTlb16
  std::vector<mmloam::Estimator::FeatureLine> lines;
  std::vector<mmloam::Estimator::FeaturePlanVec> planes;
  bool degenerate = false;
  est.thres_dist = 1.0;
  est.processPointToLine(lines, corner, &pose[7]);
  est.processPointToPlanVec(planes, surf, &pose[7], degenerate);
  double sl = 0, sp = 0;
  int vl = 0, vp = 0;
  for (const auto& l : lines) { sl += l.error; vl += l.valid; }
  for (const auto& q : planes) { sp += q.error; vp += q.valid; }
  std::printf("lines %zu %d %.12e\nplanes %zu %d %.12e\ndegenerate %d\n", lines.size(), vl, sl, planes.size(), vp, sp, (int)degenerate);
  mmloam::Estimator::Pose fr;
  for (int k = 0; k < 3; k++) fr.P[k] = pose[k];
  for (int k = 0; k < 4; k++) fr.Q[k] = pose[3 + k];
  est.EstimateLidarPose(fr, scan, &pose[23]);
  std::printf("pose %.15e %.15e %.15e %.15e %.15e %.15e %.15e\nfail %d\n", fr.P[0], fr.P[1], fr.P[2], fr.Q[0], fr.Q[1], fr.Q[2], fr.Q[3],
              (int)est.failureDetected());
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 2 && std::strcmp(argv[1], "est") == 0) {
    try {
      return run_est(argv[2]);
    } catch (const std::exception& e) {
      std::fprintf(stderr, "%s\n", e.what());
      return 2;
    }
  }
  const int n = argc > 1 ? std::atoi(argv[1]) : 1800;
  std::vector<mmloam::PointXYZINormal> line(n);
  for (int i = 0; i < n; i++) {
    // a square room seen from its centre: flat walls and four 90-degree corners
    const double a = 2.0 * M_PI * i / n;
    const double c = std::cos(a), s = std::sin(a);
    const double r = 4.0 / std::fmax(std::fabs(c), std::fabs(s));
    line[i] = {};
    line[i].x = (float)(r * c);
    line[i].y = (float)(r * s);
    line[i].z = 0.3f;
    line[i].intensity = 10.f;
  }
  try {
    mmloam::Context ctx(0);
    mmloam::LidarFeatureExtractor fe(ctx);
    std::vector<int> sharp, flat;
    fe.detectFeaturePoint(line, sharp, flat);
    std::printf("sharp %zu:", sharp.size());
    for (int v : sharp) std::printf(" %d", v);
    std::printf("\nflat %zu:", flat.size());
    for (int v : flat) std::printf(" %d", v);
    std::printf("\n");
    if (argc > 2) {  // host_check <n> map: also the Estimator adapter's map update (the ring doubles as both clouds)
      mmloam::Estimator est(ctx, 0.4f, 0.2f);
      const double I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
      est.MapIncrementLocal(line, line, I16);
      int n_surf_map = 0;
      ctx.check(mml_local_map_get(ctx.get(), 1, nullptr, 0, &n_surf_map));
      std::printf("local surf map %d\n", n_surf_map);
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 2;
  }
  return 0;
}
