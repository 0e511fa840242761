// orc_grid.h — CPU ORACLE (test infrastructure only). Uniform-grid spatial index used to make the
// exact k-NN / radius searches of the oracle finish in seconds. The grid only accelerates: results
// are defined independently of it (exact neighbours, ties broken by the lower point index).
//
// Stands in for pcl::search::KdTree / FLANN KDTreeSingleIndex (pc_manager.cpp:25,
// cluster_segmentation_srv.cpp:57). PINNED CHOICES: squared distance = (dx*dx + dy*dy) + dz*dz in
// float (FLANN L2_Simple accumulates left to right); k-NN ties at equal distance keep the lower
// index (FLANN's order is traversal dependent and unspecified).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <unordered_map>
#include <vector>

namespace orc {

inline float sqdist3(const float* a, const float* b) {
  float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  return (dx * dx + dy * dy) + dz * dz;
}

struct Grid {
  const float* xyz = nullptr;  // n x 4
  int n = 0;
  double h = 1.0;
  double mn[3] = {0, 0, 0};
  std::vector<int> order;  // point indices grouped by cell
  struct Range { int begin, end; };
  std::unordered_map<uint64_t, Range> cells;

  static uint64_t key(int64_t ix, int64_t iy, int64_t iz) {
    return ((uint64_t)(ix & 0x1fffff) << 42) | ((uint64_t)(iy & 0x1fffff) << 21) | (uint64_t)(iz & 0x1fffff);
  }
  void cellOf(const float* p, int64_t c[3]) const {
    for (int a = 0; a < 3; ++a) c[a] = (int64_t)std::floor(((double)p[a] - mn[a]) / h);
  }
  void build(const float* xyz4, int n_, double h_) {
    xyz = xyz4;
    n = n_;
    h = h_;
    for (int a = 0; a < 3; ++a) mn[a] = 0.0;
    bool any = false;
    for (int i = 0; i < n; ++i) {
      const float* p = xyz + 4 * (size_t)i;
      if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
      for (int a = 0; a < 3; ++a) mn[a] = any ? std::min(mn[a], (double)p[a]) : (double)p[a];
      any = true;
    }
    std::vector<std::pair<uint64_t, int>> keyed;
    keyed.reserve(n);
    for (int i = 0; i < n; ++i) {
      const float* p = xyz + 4 * (size_t)i;
      if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
      int64_t c[3];
      cellOf(p, c);
      keyed.emplace_back(key(c[0], c[1], c[2]), i);
    }
    std::sort(keyed.begin(), keyed.end());
    order.resize(keyed.size());
    cells.clear();
    cells.reserve(keyed.size());
    for (size_t i = 0; i < keyed.size();) {
      size_t j = i;
      while (j < keyed.size() && keyed[j].first == keyed[i].first) ++j;
      cells[keyed[i].first] = Range{(int)i, (int)j};
      for (size_t t = i; t < j; ++t) order[t] = keyed[t].second;
      i = j;
    }
  }
  // cell size giving roughly `target` points per occupied cell (surface-like clouds: occupancy ~ h^2)
  static double chooseCell(const float* xyz4, int n, double target) {
    if (n <= 0) return 1.0;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a) {
        double v = xyz4[4 * (size_t)i + a];
        if (!std::isfinite(v)) continue;
        lo[a] = std::min(lo[a], v);
        hi[a] = std::max(hi[a], v);
      }
    double ext = std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2], 1e-6});
    double h = ext / std::max(1.0, std::sqrt((double)n / target));
    for (int it = 0; it < 4; ++it) {
      std::unordered_map<uint64_t, int> occ;
      for (int i = 0; i < n; ++i) {
        const float* p = xyz4 + 4 * (size_t)i;
        if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
        occ[key((int64_t)std::floor((p[0] - lo[0]) / h), (int64_t)std::floor((p[1] - lo[1]) / h),
                (int64_t)std::floor((p[2] - lo[2]) / h))]++;
      }
      double avg = (double)n / std::max<size_t>(1, occ.size());
      double ratio = std::sqrt(target / avg);
      if (ratio > 0.8 && ratio < 1.25) break;
      h *= std::min(4.0, std::max(0.25, ratio));
    }
    return std::max(h, ext * 1e-6);
  }
};

}  // namespace orc
