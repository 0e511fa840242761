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
#include "mgrid.cuh"

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
// read-only find: used once the unions are complete. (Writing the root back into parent[] there would race with the path halving
// of other threads' finds: a stale grandparent written after the root left parent[i] pointing at a non-root, the point then lost
// its label although it had been counted - about one frame in 3000.)
__device__ __forceinline__ int uf_find_ro(const int* parent, int x) {
  int p = __ldcg(parent + x);
  while (p != x) {
    x = p;
    p = __ldcg(parent + x);
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

__global__ void cc_init_kernel(int* parent, int* size, int* root, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { parent[i] = i; size[i] = 0; root[i] = i; }
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
__global__ void cc_flatten_kernel(GridDev g, const int* __restrict__ parent, int* __restrict__ size, int* __restrict__ root) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.n) return;
  const int i = __float_as_int(g.sorted[t].w);
  const int r = uf_find_ro(parent, i);
  root[i] = r;
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
__global__ void cc_label_kernel(const int* __restrict__ root, const int* __restrict__ size, const int* __restrict__ rank_of,
                                int n, int* __restrict__ labels) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = root[i];
  labels[i] = (size[r] > 0) ? rank_of[r] : -1;  // size 0: point never entered the grid (non-finite)
}
__global__ void fill_i32_kernel(int* p, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------------------------------------
// The same extraction with every data-dependent size kept on the device (the clusterize() path of a frame): the grid is the
// block-Morton grid of knn.cu with fine cells a hair wider than the tolerance, the components come from the same lock-free
// union-find, and the PCL ordering (size descending, ties by smallest index), the size filter, the per-cluster ordered index
// lists, the cluster clouds and the centroid sums are all produced by kernels. The host reads ONE block at the end.
// ---------------------------------------------------------------------------------------------------------------------
// Grid cells of 0.57 tol: the diagonal of a cell is below the tolerance, so ALL points of a cell are mutually connected and one
// union with the cell's first point stands for the whole clique; a point's other neighbours lie within +-2 cells, and for each of
// those 124 cells ONE edge is enough (the cell is a single component): the cell is skipped outright when it already shares the
// point's root, otherwise its points are tried until one is within the tolerance. A warp serves one point, its lanes take the
// cells. A dense object has hundreds of points within the tolerance of each point; this visits a few dozen of them. The
// components are exactly those of the radius graph (every edge used is a radius edge, every radius edge joins two cells that
// end up connected).
constexpr float CC_CELL_FACTOR = 0.57f;  // cell size / tolerance: sqrt(3) * 0.57 = 0.987 < 1, 2 * 0.57 = 1.14 > 1
// Two launches: phase 0 takes the 27 cells at offsets -1..1 (neighbouring cells of a surface are almost always connected and
// the witness is found after a few points), phase 1 the 98 cells of the outer shell, where most pairs are NOT within the
// tolerance: by then the components are nearly final and a single find per cell skips them (run in one launch, every warp
// starts from the unmerged state and scans those cells in full: 86 us instead of 2 x 15 on the objects of a tabletop frame).
// Cells whose box is out of reach of the point are skipped without a look at their points.
__global__ void __launch_bounds__(CC_WARPS * 32)
cc_union_mg_kernel(const MGrid* __restrict__ G, const int* __restrict__ start, const float4* __restrict__ sorted, float r2,
                   int* __restrict__ parent, int phase) {
  const MGrid g = *G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * CC_WARPS + warp;
  if (t >= g.n_finite) return;
  const float4 q = sorted[t];
  const int qi = __float_as_int(q.w);
  const int cx = mg_coord(q.x, g.mnx, g.inv_hf, g.dx), cy = mg_coord(q.y, g.mny, g.inv_hf, g.dy), cz = mg_coord(q.z, g.mnz, g.inv_hf, g.dz);
  int rq = uf_find(parent, qi);
  for (int ci = lane; ci < 125; ci += 32) {
    const int ox = (ci % 5) - 2, oy = ((ci / 5) % 5) - 2, oz = (ci / 25) - 2;
    const bool inner = (ox >= -1 && ox <= 1 && oy >= -1 && oy <= 1 && oz >= -1 && oz <= 1);
    if (inner != (phase == 0)) continue;
    const int x = cx + ox, y = cy + oy, z = cz + oz;
    if (x < 0 || y < 0 || z < 0 || x >= g.dx || y >= g.dy || z >= g.dz) continue;
    const int idx = mg_index(g, x, y, z);
    const int jb = start[idx], je = start[idx + 1];
    if (jb == je) continue;
    const int pf = __float_as_int(sorted[jb].w);
    if (ci == 62) {  // the point's own cell: a clique
      if (pf != qi) {
        uf_union(parent, rq, uf_find(parent, pf));
        rq = uf_find(parent, qi);
      }
      continue;
    }
    if (!inner) {
      // squared distance from the point to the cell's box, shrunk by a relative margin so that rounding can only keep a cell
      const float lx = g.mnx + (float)x * g.hf, ly = g.mny + (float)y * g.hf, lz = g.mnz + (float)z * g.hf;
      const float ex = fmaxf(fmaxf(lx - q.x, q.x - (lx + g.hf)), 0.0f), ey = fmaxf(fmaxf(ly - q.y, q.y - (ly + g.hf)), 0.0f),
                  ez = fmaxf(fmaxf(lz - q.z, q.z - (lz + g.hf)), 0.0f);
      if (((ex * ex + ey * ey) + ez * ez) * 0.98f > r2) continue;
    }
    rq = uf_find(parent, qi);
    if (uf_find(parent, pf) == rq) continue;  // that cell is already in this point's component
    for (int j = jb; j < je; ++j) {
      const float4 p = sorted[j];
      const float ddx = q.x - p.x, ddy = q.y - p.y, ddz = q.z - p.z;
      const float d = (ddx * ddx + ddy * ddy) + ddz * ddz;
      if (d < r2) {
        uf_union(parent, rq, uf_find(parent, __float_as_int(p.w)));
        rq = uf_find(parent, qi);
        break;
      }
    }
  }
}
__global__ void cc_flatten_mg_kernel(const MGrid* __restrict__ G, const float4* __restrict__ sorted, const int* __restrict__ parent,
                                     int* __restrict__ size, int* __restrict__ root) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= G->n_finite) return;
  const int i = __float_as_int(sorted[t].w);
  const int r = uf_find_ro(parent, i);
  root[i] = r;
  atomicAdd(&size[r], 1);
}
// head: [0] clusters kept (nc), [1] roots that passed the size filter (> CC_MAXC: overflow, the caller takes the host path),
// [2 .. 2+CC_MAXC) sizes in PCL order, then offsets [CC_MAXC + 1]
__global__ void __launch_bounds__(CC_MAXC)
cc_rank_dev_kernel(const int* __restrict__ n_roots, const int2* __restrict__ roots, int* __restrict__ rank_of, int* __restrict__ head,
                   const MGrid* __restrict__ G, float h_req) {
  __shared__ int s_size[CC_MAXC], s_root[CC_MAXC], s_sorted[CC_MAXC];
  int R = *n_roots;
  // the cell table did not fit at the requested cell size and the cells were enlarged: the clique argument of cc_union_mg_kernel
  // does not hold, the caller must take the host-ordered path (reported like an overflow of the cluster table)
  if (G->hf > h_req * 1.0001f) R = CC_MAXC + 1;
  if (R > CC_MAXC) {
    if (threadIdx.x == 0) { head[0] = 0; head[1] = R; }
    return;
  }
  const int t = threadIdx.x;
  if (t < R) { s_root[t] = roots[t].x; s_size[t] = roots[t].y; }
  __syncthreads();
  if (t < R) {
    // PCL order: size descending; equal sizes by smallest point index (the root)
    int rank = 0;
    for (int j = 0; j < R; ++j)
      rank += (s_size[j] > s_size[t] || (s_size[j] == s_size[t] && s_root[j] < s_root[t])) ? 1 : 0;
    rank_of[s_root[t]] = rank;
    s_sorted[rank] = s_size[t];
  }
  __syncthreads();
  if (t == 0) {
    head[0] = R;
    head[1] = R;
    int off = 0;
    for (int c = 0; c < R; ++c) {
      head[2 + c] = s_sorted[c];
      head[2 + CC_MAXC + c] = off;
      off += s_sorted[c];
    }
    head[2 + CC_MAXC + R] = off;
  }
}
// one CTA per cluster: ordered (ascending index) compaction of its points -> index list + cluster cloud
__global__ void __launch_bounds__(1024)
cc_compact_dev_kernel(const float4* __restrict__ xyz, const int* __restrict__ labels, int n, const int* __restrict__ head,
                      int* __restrict__ idx_out, float4* __restrict__ pts_out) {
  __shared__ int s_w[32];
  __shared__ int s_base, s_tile;
  const int c = blockIdx.x;
  if (c >= head[0]) return;
  const int off = head[2 + CC_MAXC + c];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int tile = 0; tile < n; tile += 1024) {
    const int i = tile + threadIdx.x;
    const bool in = (i < n) && labels[i] == c;
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
      const int v = s_w[lane];
      int incl = v;
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      s_w[lane] = incl - v;
      if (lane == 31) s_tile = incl;
    }
    __syncthreads();
    const int base = s_base;
    if (in) {
      const int pos = off + base + s_w[warp] + __popc(m & ((1u << lane) - 1u));
      idx_out[pos] = i;
      pts_out[pos] = xyz[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + s_tile;
    __syncthreads();
  }
}

int euclidean_clusters_dev(pitt_ctx* ctx, const float4* d_xyz, int n, double tolerance, int min_size, int max_size, ClustersOnDevice* out) {
  const float r2 = (float)(tolerance * tolerance);
  MGridBuf mg;
  PITT_TRY(mgrid_build(ctx, d_xyz, n, (float)tolerance * CC_CELL_FACTOR, &mg));
  int* d_parent = nullptr;
  int* d_size = nullptr;
  int* d_rank = nullptr;
  int* d_nroots = nullptr;
  int2* d_roots = nullptr;
  int* d_root = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_parent));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_size));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_root));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_rank));
  PITT_TRY(arena_alloc(ctx, 1, &d_nroots));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_roots));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &out->d_labels));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &out->d_idx));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &out->d_points));
  PITT_TRY(arena_alloc(ctx, (size_t)CC_HEAD_INTS, &out->d_head));
  cc_init_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_parent, d_size, d_root, n);
  fill_i32_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_rank, n, -1);
  PITT_CUDA(ctx, cudaMemsetAsync(d_nroots, 0, sizeof(int), ctx->stream));
  PITT_CUDA(ctx, cudaMemsetAsync(out->d_head, 0, CC_HEAD_INTS * sizeof(int), ctx->stream));
  cc_union_mg_kernel<<<cdiv(n, CC_WARPS), CC_WARPS * 32, 0, ctx->stream>>>(mg.d_G, mg.d_start, mg.d_sorted, r2, d_parent, 0);
  cc_union_mg_kernel<<<cdiv(n, CC_WARPS), CC_WARPS * 32, 0, ctx->stream>>>(mg.d_G, mg.d_start, mg.d_sorted, r2, d_parent, 1);
  cc_flatten_mg_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(mg.d_G, mg.d_sorted, d_parent, d_size, d_root);
  cc_roots_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_parent, d_size, n, std::max(min_size, 1), max_size, d_nroots, d_roots, n);
  cc_rank_dev_kernel<<<1, CC_MAXC, 0, ctx->stream>>>(d_nroots, d_roots, d_rank, out->d_head, mg.d_G, (float)tolerance * CC_CELL_FACTOR);
  cc_label_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_root, d_size, d_rank, n, out->d_labels);
  cc_compact_dev_kernel<<<CC_MAXC, 1024, 0, ctx->stream>>>(d_xyz, out->d_labels, n, out->d_head, out->d_idx, out->d_points);
  ctx->launches += 9;
  PITT_CUDA(ctx, cudaGetLastError());
  return PITT_OK;
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
  int* d_root = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_root));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_parent));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_size));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_rank));
  PITT_TRY(arena_alloc(ctx, 1, &d_nroots));
  PITT_TRY(arena_alloc(ctx, (size_t)cap, &d_roots));
  cc_init_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_parent, d_size, d_root, n);
  fill_i32_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_rank, n, -1);
  PITT_CUDA(ctx, cudaMemsetAsync(d_nroots, 0, sizeof(int), ctx->stream));
  ctx->launches += 2;
  int h_nroots = 0;
  std::vector<int2> roots;
  if (g.n > 0) {
    cc_union_kernel<<<cdiv(g.n, CC_WARPS), CC_WARPS * 32, 0, ctx->stream>>>(g, r2, d_parent);
    cc_flatten_kernel<<<cdiv(g.n, 256), 256, 0, ctx->stream>>>(g, d_parent, d_size, d_root);
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
  cc_label_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_root, d_size, d_rank, n, d_labels);
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
