// grid.cuh — uniform grid hash over a staged cloud (device). Shared by the k-NN normal estimator
// (K8) and the Euclidean clusterer (K9): points are counting-sorted by cell so that a cell's points
// are contiguous; a dense cell_start table gives O(1) cell lookup.
//
// Stands in for pcl::search::KdTree (FLANN) of pc_manager.cpp:25 and cluster_segmentation_srv.cpp:57;
// it only accelerates exact searches, it never changes their result.
#pragma once
#include "pitt_common.cuh"

namespace pitt {

struct GridDev {
  float mnx, mny, mnz;  // grid origin (cloud minimum)
  float h, inv_h;
  int dx, dy, dz;       // cells per axis
  int ncells;
  int n;                // points in the grid (non-finite points are dropped)
  const int* cell_start;   // [ncells + 1]
  const float4* sorted;    // [n] {x, y, z, original index as int bits}, grouped by cell
};

__device__ __forceinline__ int grid_coord(float v, float mn, float inv_h, int dim) {
  int c = (int)floorf((v - mn) * inv_h);
  return min(max(c, 0), dim - 1);
}
__device__ __forceinline__ int grid_cell(const GridDev& g, float x, float y, float z) {
  return (grid_coord(z, g.mnz, g.inv_h, g.dz) * g.dy + grid_coord(y, g.mny, g.inv_h, g.dy)) * g.dx +
         grid_coord(x, g.mnx, g.inv_h, g.dx);
}

// Builds the grid on ctx->stream in arena memory (valid until the next outermost API call).
// h <= 0: pick the cell size so that occupied cells hold about `target_per_cell` points.
int grid_build(pitt_ctx* ctx, const float4* d_xyz, int n, float h, float target_per_cell, GridDev* out);

// minimum / maximum over the finite points (one small D2H + stream sync) and their number
int cloud_bbox(pitt_ctx* ctx, const float4* d_xyz, int n, float mn[3], float mx[3], int* n_finite);

// exclusive scan of n ints (device, in place), total written to d_total[0] (may be null)
int device_exclusive_scan(pitt_ctx* ctx, int* d_data, int n, int* d_total);
// exclusive prefix maximum of n floats in place (element 0 becomes -inf)
int device_exclusive_max_scan(pitt_ctx* ctx, float* d_data, int n);

// One-launch variants: the data is scanned in place per chunk of SCAN_CHUNK elements (exclusive, chunk relative) and *d_chunk_off
// (arena) receives the exclusive chunk offsets; element i of the full scan is chunk_off[i >> SCAN_CHUNK_LOG2] (+ or max) data[i].
// d_ticket: a zero-initialised device word (left at zero again).
constexpr int SCAN_CHUNK_LOG2 = 11;
int device_scan_chunks(pitt_ctx* ctx, int* d_data, int n, int** d_chunk_off, unsigned* d_ticket, int* d_total);
int device_max_scan_chunks(pitt_ctx* ctx, float* d_data, int n, float** d_chunk_off, unsigned* d_ticket);

}  // namespace pitt
