// Compile-and-link check of the C++ adapters (build()), and a small GPU run (tests, -m gpu):
//   host_check <n_az>       extracts features of a synthetic ring through
//                           LidarFeatureExtractor::detectFeaturePoint and prints the index lists;
//   host_check <n_az> map   additionally runs Estimator::MapIncrementLocal on the ring.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "mmloam_shim.hpp"

int main(int argc, char** argv) {
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
