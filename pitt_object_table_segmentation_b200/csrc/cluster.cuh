// cluster.cuh / knn interface used by the service layer
#pragma once
#include <vector>

#include "pitt_common.cuh"

namespace pitt {
// labels (device, n ints): cluster rank in PCL order or -1; sizes_out: cluster sizes in rank order
int euclidean_clusters_impl(pitt_ctx* ctx, const float4* d_xyz, int n, double tolerance, int min_size, int max_size,
                            int* d_labels, std::vector<int>* sizes_out);
// device-resident result of euclidean_clusters_dev (arena memory): d_head = {nc, roots that passed the size filter, sizes[CC_MAXC]
// in PCL order, offsets[CC_MAXC + 1]}; d_idx / d_points = the clusters' ascending index lists / clouds back to back
constexpr int CC_MAXC = 128;
constexpr int CC_HEAD_INTS = 2 + CC_MAXC + CC_MAXC + 1;
struct ClustersOnDevice {
  int* d_head;
  int* d_labels;
  int* d_idx;
  float4* d_points;
};
int euclidean_clusters_dev(pitt_ctx* ctx, const float4* d_xyz, int n, double tolerance, int min_size, int max_size, ClustersOnDevice* out);
constexpr int KNN_BRUTE_MAX = 4096;  // below this size the all-pairs kernel beats grid + search
int estimate_normals_segmented(pitt_ctx* ctx, const float4* d_xyz, int n_total, const int* d_seg_off, const int* d_n_seg, int k,
                               const float vp[3], float4* d_nrm);
int estimate_normals_impl(pitt_ctx* ctx, const float4* d_xyz, int n, int k, const float vp[3], float4* d_nrm);
}  // namespace pitt
