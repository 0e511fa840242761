// orc_cluster.cpp — CPU ORACLE (test infrastructure only).
// pcl::extractEuclideanClusters (segmentation/impl/extract_clusters.hpp, PCL 1.7) as driven by
// clusterize() (reference: src/segmentation_services/cluster_segmentation_srv.cpp:57-69), SURVEY.md B.10:
// BFS over the radius graph (edge iff squared distance < float(tol*tol)... PCL's radiusSearch
// compares FLANN's float squared distance with radius*radius), components kept when
// min <= size <= max, indices ascending, clusters sorted by size descending.
// PINNED CHOICE: equal-size clusters are ordered by their smallest point index (std::sort on the
// reversed range is unstable in PCL, so the reference order is unspecified there).
#include <queue>

#include "oracle.h"
#include "orc_grid.h"

using namespace orc;

namespace orc {
// connected components of the radius graph; comp[i] = component id (by smallest index order) or -1 for non-finite
void radiusComponents(const float* xyz4, int n, double tol, std::vector<int>& comp, std::vector<std::vector<int>>& comps) {
  comp.assign(n, -1);
  comps.clear();
  if (n <= 0) return;
  Grid g;
  g.build(xyz4, n, tol);
  const float r2 = (float)(tol * tol);
  std::vector<char> processed(n, 0);
  std::vector<int> queue;
  for (int i = 0; i < n; ++i) {
    if (processed[i]) continue;
    const float* pi = xyz4 + 4 * (size_t)i;
    if (!(std::isfinite(pi[0]) && std::isfinite(pi[1]) && std::isfinite(pi[2]))) continue;
    queue.clear();
    queue.push_back(i);
    processed[i] = 1;
    for (size_t s = 0; s < queue.size(); ++s) {
      int q = queue[s];
      const float* pq = xyz4 + 4 * (size_t)q;
      int64_t c[3];
      g.cellOf(pq, c);
      for (int64_t dz = -1; dz <= 1; ++dz)
        for (int64_t dy = -1; dy <= 1; ++dy)
          for (int64_t dx = -1; dx <= 1; ++dx) {
            auto it = g.cells.find(Grid::key(c[0] + dx, c[1] + dy, c[2] + dz));
            if (it == g.cells.end()) continue;
            for (int t = it->second.begin; t < it->second.end; ++t) {
              int j = g.order[t];
              if (processed[j]) continue;
              if (sqdist3(pq, xyz4 + 4 * (size_t)j) < r2) {
                processed[j] = 1;
                queue.push_back(j);
              }
            }
          }
    }
    std::sort(queue.begin(), queue.end());
    int id = (int)comps.size();
    for (int j : queue) comp[j] = id;
    comps.push_back(queue);
  }
}
}  // namespace orc

extern "C" int orc_euclidean_clusters(const float* xyz4, int n, double tol, int min_size, int max_size, int32_t* labels,
                                      int* n_clusters) {
  for (int i = 0; i < n; ++i) labels[i] = -1;
  *n_clusters = 0;
  std::vector<int> comp;
  std::vector<std::vector<int>> comps;
  radiusComponents(xyz4, n, tol, comp, comps);
  std::vector<int> keep;
  for (size_t c = 0; c < comps.size(); ++c)
    if ((int)comps[c].size() >= min_size && (int)comps[c].size() <= max_size) keep.push_back((int)c);
  std::stable_sort(keep.begin(), keep.end(), [&](int a, int b) {
    if (comps[a].size() != comps[b].size()) return comps[a].size() > comps[b].size();
    return comps[a][0] < comps[b][0];
  });
  for (size_t r = 0; r < keep.size(); ++r)
    for (int j : comps[keep[r]]) labels[j] = (int)r;
  *n_clusters = (int)keep.size();
  return PITT_OK;
}
