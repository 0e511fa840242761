// cluster.cuh / knn interface used by the service layer
#pragma once
#include <vector>

#include "pitt_common.cuh"

namespace pitt {
// labels (device, n ints): cluster rank in PCL order or -1; sizes_out: cluster sizes in rank order
int euclidean_clusters_impl(pitt_ctx* ctx, const float4* d_xyz, int n, double tolerance, int min_size, int max_size,
                            int* d_labels, std::vector<int>* sizes_out);
int estimate_normals_impl(pitt_ctx* ctx, const float4* d_xyz, int n, int k, const float vp[3], float4* d_nrm);
}  // namespace pitt
