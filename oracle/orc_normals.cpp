// orc_normals.cpp — CPU ORACLE (test infrastructure only).
// pcl::NormalEstimation<PointXYZ,Normal>::computeFeature with setKSearch(k) as driven by
// PCManager::estimateNormal (reference: src/point_cloud_library/pc_manager.cpp:68-78, k = 50 :18),
// SURVEY.md B.7/B.8. Exact k-NN (query included, ties by lower index, see orc_grid.h), then
// computeMeanAndCovarianceMatrix in float over the neighbours in (distance, index) order,
// solvePlaneParameters (eigen33, curvature) and flipNormalTowardsViewpoint.
#include <queue>

#include "oracle.h"
#include "orc_grid.h"
#include "orc_math.h"

namespace orc {

struct Cand {
  float d;
  int i;
  bool operator<(const Cand& o) const { return d < o.d || (d == o.d && i < o.i); }  // "better" first
};

// exact k-NN of point q among all points, written sorted by (distance, index). returns count.
static int knnQuery(const Grid& g, int q, int k, std::vector<Cand>& heap, int* out_idx, float* out_sq) {
  const float* pq = g.xyz + 4 * (size_t)q;
  heap.clear();
  if (!(std::isfinite(pq[0]) && std::isfinite(pq[1]) && std::isfinite(pq[2]))) return 0;
  int64_t c[3];
  g.cellOf(pq, c);
  const int want = std::min(k, (int)g.order.size());
  auto worse = [](const Cand& a, const Cand& b) { return a < b; };  // max-heap on (d,i)
  // number of rings needed to cover the whole grid is bounded by the cloud extent; stop when no
  // unexplored cell can hold a better candidate
  for (int r = 0;; ++r) {
    bool any_cell = false;
    for (int64_t dz = -r; dz <= r; ++dz)
      for (int64_t dy = -r; dy <= r; ++dy)
        for (int64_t dx = -r; dx <= r; ++dx) {
          if (std::max({std::llabs(dx), std::llabs(dy), std::llabs(dz)}) != r) continue;
          auto it = g.cells.find(Grid::key(c[0] + dx, c[1] + dy, c[2] + dz));
          if (it == g.cells.end()) continue;
          any_cell = true;
          for (int t = it->second.begin; t < it->second.end; ++t) {
            int j = g.order[t];
            Cand cd{sqdist3(pq, g.xyz + 4 * (size_t)j), j};
            if ((int)heap.size() < want) {
              heap.push_back(cd);
              std::push_heap(heap.begin(), heap.end(), worse);
            } else if (cd < heap.front()) {
              std::pop_heap(heap.begin(), heap.end(), worse);
              heap.back() = cd;
              std::push_heap(heap.begin(), heap.end(), worse);
            }
          }
        }
    (void)any_cell;
    if ((int)heap.size() >= want) {
      // everything outside the cube of radius r cells is farther than `bound`
      double bound = 1e300;
      for (int a = 0; a < 3; ++a) {
        double lo = g.mn[a] + (double)(c[a] - r) * g.h, hi = g.mn[a] + (double)(c[a] + r + 1) * g.h;
        bound = std::min({bound, (double)pq[a] - lo, hi - (double)pq[a]});
      }
      if (bound > 0 && (double)heap.front().d < bound * bound * (1.0 - 1e-5)) break;
    }
    if (r > (1 << 21)) break;
  }
  std::sort(heap.begin(), heap.end());
  for (size_t t = 0; t < heap.size(); ++t) {
    if (out_idx) out_idx[t] = heap[t].i;
    if (out_sq) out_sq[t] = heap[t].d;
  }
  return (int)heap.size();
}

}  // namespace orc

using namespace orc;

extern "C" {

int orc_knn(const float* xyz4, int n, int k, int32_t* out_idx, float* out_sqdist) {
  if (n <= 0 || k <= 0) return PITT_OK;
  Grid g;
  g.build(xyz4, n, Grid::chooseCell(xyz4, n, std::max(2.0, k / 4.0)));
  std::vector<Cand> heap;
  std::vector<int> idx(k);
  std::vector<float> sq(k);
  for (int q = 0; q < n; ++q) {
    int m = knnQuery(g, q, k, heap, idx.data(), sq.data());
    for (int t = 0; t < k; ++t) {
      out_idx[(size_t)q * k + t] = t < m ? idx[t] : -1;
      if (out_sqdist) out_sqdist[(size_t)q * k + t] = t < m ? sq[t] : INFINITY;
    }
  }
  return PITT_OK;
}

int orc_estimate_normals(const float* xyz4, int n, int k, const float* vp, float* out4) {
  if (n <= 0) return PITT_OK;
  Grid g;
  g.build(xyz4, n, Grid::chooseCell(xyz4, n, std::max(2.0, k / 4.0)));
  std::vector<Cand> heap;
  std::vector<int> idx(std::max(k, 1));
  const float nanf_ = std::numeric_limits<float>::quiet_NaN();
  for (int q = 0; q < n; ++q) {
    float* o = out4 + 4 * (size_t)q;
    int m = knnQuery(g, q, k, heap, idx.data(), nullptr);
    if (m < 3) {
      o[0] = o[1] = o[2] = o[3] = nanf_;
      continue;
    }
    // computeMeanAndCovarianceMatrix(cloud, indices): float accumulators in index-list order
    float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < m; ++t) {
      const float* p = xyz4 + 4 * (size_t)idx[t];
      accu[0] += p[0] * p[0]; accu[1] += p[0] * p[1]; accu[2] += p[0] * p[2];
      accu[3] += p[1] * p[1]; accu[4] += p[1] * p[2]; accu[5] += p[2] * p[2];
      accu[6] += p[0]; accu[7] += p[1]; accu[8] += p[2];
    }
    float cov[9], cen[4];
    covFromAccu(accu, (float)m, cov, cen);
    float ev, evec[3];
    eigen33(cov, ev, evec);
    float nx = evec[0], ny = evec[1], nz = evec[2];
    float eig_sum = cov[0] + cov[4] + cov[8];
    float curv = (eig_sum != 0.0f) ? fabsf(ev / eig_sum) : 0.0f;
    // flipNormalTowardsViewpoint(point, vp_x, vp_y, vp_z, nx, ny, nz)
    const float* p = xyz4 + 4 * (size_t)q;
    float vx = vp[0] - p[0], vy = vp[1] - p[1], vz = vp[2] - p[2];
    float cos_theta = (vx * nx + vy * ny + vz * nz);
    if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
    o[0] = nx; o[1] = ny; o[2] = nz; o[3] = curv;
  }
  return PITT_OK;
}

}  // extern "C"
