// cluster.cu — Euclidean cluster extraction (K9): radius graph on the uniform grid + lock-free
// union-find connected components + size filter + PCL ordering.
//
// Replaces ec.extract() of clusterize() (reference: src/segmentation_services/
// cluster_segmentation_srv.cpp:57-69; pcl::extractEuclideanClusters). The BFS of
// PCL yields exactly the connected components of the graph "squared distance < tol^2", which is
// what the union-find computes; clusters are ordered by size (descending), ties by smallest index.
#include <algorithm>
#include <vector>

#include "cluster.cuh"
#include "grid.cuh"

namespace pitt {

__device__ __forceinline__ int uf_find(int* parent, int x) {
  int p = parent[x];
  while (p != x) {
    int gp = parent[p];
    if (gp != p) parent[x] = gp;  // path halving (benign race: only ever shortcuts towards the root)
    x = p;
    p = gp;
  }
  return x;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }  // link the larger root under the smaller: root = min index
    int old = atomicMin(&parent[a], b);
    if (old == a) return;
    a = old;
  }
}

__global__ void cc_init_kernel(int* parent, int* size, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { parent[i] = i; size[i] = 0; }
}
// One WARP per query point: the 32 lanes stride over the candidates of the 27 surrounding cells. A dense
// object cluster has hundreds of neighbours within the tolerance per point and only a few thousand points
// in total, so a thread per point leaves the GPU almost empty and serialises hundreds of dependent
// find() chains per thread (1.5 ms for 6 k points); a warp per point is 32x wider and each lane keeps the
// query's current root in a register, so an edge inside an already merged component costs one find().
constexpr int CC_WARPS = 4;
__global__ void __launch_bounds__(CC_WARPS * 32) cc_union_kernel(GridDev g, float r2, int* __restrict__ parent) {
  const int t = blockIdx.x * CC_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= g.n) return;
  const float4 q = g.sorted[t];
  const int qi = __float_as_int(q.w);
  const int cx = grid_coord(q.x, g.mnx, g.inv_h, g.dx), cy = grid_coord(q.y, g.mny, g.inv_h, g.dy),
            cz = grid_coord(q.z, g.mnz, g.inv_h, g.dz);
  int rq = uf_find(parent, qi);
  // cells are at least tol wide, but float rounding of the cell assignment can put a neighbour two
  // cells away: search +-1 cells plus the margin handled by building the grid with h slightly > tol
  for (int z = max(cz - 1, 0); z <= min(cz + 1, g.dz - 1); ++z)
    for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dy - 1); ++y) {
      // the cells x-1, x, x+1 of one row are contiguous in the sorted array
      const int xa = max(cx - 1, 0), xb = min(cx + 1, g.dx - 1);
      const int row = (z * g.dy + y) * g.dx;
      const int b = g.cell_start[row + xa], e = g.cell_start[row + xb + 1];
      for (int j = b + lane; j < e; j += 32) {
        const float4 p = g.sorted[j];
        const int pi = __float_as_int(p.w);
        if (pi >= qi) continue;  // each edge once
        const float ddx = q.x - p.x, ddy = q.y - p.y, ddz = q.z - p.z;
        const float d = (ddx * ddx + ddy * ddy) + ddz * ddz;
        if (d < r2) {
          const int rp = uf_find(parent, pi);
          if (rp != rq) {
            uf_union(parent, rq, rp);
            rq = uf_find(parent, qi);
          }
        }
      }
    }
}
__global__ void cc_flatten_kernel(GridDev g, int* __restrict__ parent, int* __restrict__ size) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.n) return;
  const int i = __float_as_int(g.sorted[t].w);
  const int r = uf_find(parent, i);
  parent[i] = r;
  atomicAdd(&size[r], 1);
}
__global__ void cc_roots_kernel(const int* __restrict__ parent, const int* __restrict__ size, int n, int min_size,
                                int max_size, int* __restrict__ n_roots, int2* __restrict__ roots, int cap) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = size[i];
  if (parent[i] == i && s >= min_size && s <= max_size && s > 0) {
    int pos = atomicAdd(n_roots, 1);
    if (pos < cap) roots[pos] = make_int2(i, s);
  }
}
__global__ void cc_rank_kernel(const int2* __restrict__ ranked, int k, int* __restrict__ rank_of) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k) rank_of[ranked[i].x] = i;
}
__global__ void cc_label_kernel(const int* __restrict__ parent, const int* __restrict__ size, const int* __restrict__ rank_of,
                                int n, int* __restrict__ labels) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = parent[i];
  labels[i] = (size[r] > 0) ? rank_of[r] : -1;  // size 0: point never entered the grid (non-finite)
}
__global__ void fill_i32_kernel(int* p, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int euclidean_clusters_impl(pitt_ctx* ctx, const float4* d_xyz, int n, double tolerance, int min_size, int max_size,
                            int* d_labels, std::vector<int>* sizes_out) {
  if (sizes_out) sizes_out->clear();
  if (n <= 0) return PITT_OK;
  const float r2 = (float)(tolerance * tolerance);
  GridDev g;
  // h a hair above tol so that every neighbour within tol lies in the 27 surrounding cells even
  // after float rounding of the cell coordinates
  PITT_TRY(grid_build(ctx, d_xyz, n, (float)tolerance * 1.001f + 1e-7f, 0.f, &g));
  int* d_parent = nullptr;
  int* d_size = nullptr;
  int* d_rank = nullptr;
  int* d_nroots = nullptr;
  int2* d_roots = nullptr;
  const int cap = n;
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_parent));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_size));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_rank));
  PITT_TRY(arena_alloc(ctx, 1, &d_nroots));
  PITT_TRY(arena_alloc(ctx, (size_t)cap, &d_roots));
  cc_init_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_parent, d_size, n);
  fill_i32_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_rank, n, -1);
  PITT_CUDA(ctx, cudaMemsetAsync(d_nroots, 0, sizeof(int), ctx->stream));
  ctx->launches += 2;
  int h_nroots = 0;
  std::vector<int2> roots;
  if (g.n > 0) {
    cc_union_kernel<<<cdiv(g.n, CC_WARPS), CC_WARPS * 32, 0, ctx->stream>>>(g, r2, d_parent);
    cc_flatten_kernel<<<cdiv(g.n, 256), 256, 0, ctx->stream>>>(g, d_parent, d_size);
    cc_roots_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_parent, d_size, n, std::max(min_size, 1), max_size, d_nroots, d_roots, cap);
    ctx->launches += 3;
    PITT_CUDA(ctx, cudaMemcpyAsync(&h_nroots, d_nroots, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
    h_nroots = std::min(h_nroots, cap);
    roots.resize(h_nroots);
    if (h_nroots > 0) {
      PITT_CUDA(ctx, cudaMemcpyAsync(roots.data(), d_roots, (size_t)h_nroots * sizeof(int2), cudaMemcpyDeviceToHost, ctx->stream));
      PITT_CUDA(ctx, pitt::stream_sync(ctx));
      // PCL order: size descending; equal sizes by smallest point index (the root)
      std::sort(roots.begin(), roots.end(), [](const int2& a, const int2& b) { return a.y != b.y ? a.y > b.y : a.x < b.x; });
      PITT_CUDA(ctx, cudaMemcpyAsync(d_roots, roots.data(), (size_t)h_nroots * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
      cc_rank_kernel<<<cdiv(h_nroots, 256), 256, 0, ctx->stream>>>(d_roots, h_nroots, d_rank);
      ctx->launches++;
    }
  }
  cc_label_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_parent, d_size, d_rank, n, d_labels);
  ctx->launches++;
  PITT_CUDA(ctx, cudaGetLastError());
  if (sizes_out)
    for (auto& r : roots) sizes_out->push_back(r.y);
  return PITT_OK;
}

}  // namespace pitt

using namespace pitt;

extern "C" int pitt_euclidean_clusters(pitt_ctx* ctx, const pitt_cloud* c, double tolerance, int min_size, int max_size,
                                       int32_t* labels, int* n_clusters) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!c || !labels || !n_clusters || !(tolerance > 0.0)) return fail(ctx, PITT_ERR_INVALID, "pitt_euclidean_clusters arguments");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  *n_clusters = 0;
  if (c->n > 0) {
    int* d_labels = nullptr;
    PITT_TRY(arena_alloc(ctx, (size_t)c->n, &d_labels));
    std::vector<int> sizes;
    PITT_TRY(euclidean_clusters_impl(ctx, c->d_xyz, c->n, tolerance, min_size, max_size, d_labels, &sizes));
    PITT_CUDA(ctx, cudaMemcpyAsync(labels, d_labels, (size_t)c->n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
    *n_clusters = (int)sizes.size();
  }
  timer.finish();
  return PITT_OK;
}
