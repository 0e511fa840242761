// services.cu — the service-shaped entry points: findSupports, clusterize, ransac*Detection,
// the primitive selection rule and the per-frame orchestration, all on device-resident clouds.
//
// Reference (paths under /root/reference/src):
//   findSupports        segmentation_services/supports_segmentation_srv.cpp:241-361 (+ :114-238 helpers)
//   clusterize          segmentation_services/cluster_segmentation_srv.cpp:38-108
//   ransac*Detection    segmentation_services/{plane,sphere,cylinder,cone}_segmentation_srv.cpp
//   inlierToVectorMsg   point_cloud_library/pc_manager.cpp:105-111
//   clustersAcquisition ransac_segmentation.cpp:223-343; depthAcquisition obj_segmentation.cpp:251-316
// The O(N*M) linear searches and the O(n^2) axis-extent loop of the reference become flag lookups,
// prefix scans and a tiled all-pairs kernel; observable quirks are kept.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "cluster.cuh"
#include "grid.cuh"
#include "pitt_common.cuh"
#include "sac.cuh"

namespace pitt {

// ------------------------------------------------------------------ small kernels
__global__ void iota_kernel(int* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}
__global__ void set_flags_kernel(const int* __restrict__ idx, const int* __restrict__ n_idx, int* __restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < *n_idx) flag[idx[i]] = 1;
}
__global__ void gather_points_kernel(const float4* __restrict__ src, const int* __restrict__ idx, int m, float4* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) dst[i] = src[idx[i]];
}
// keep[i] = 1 - flag[i] (in place scan input)
__global__ void invert_flags_kernel(const int* __restrict__ flag, int n, int* __restrict__ keep) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = flag[i] ? 0 : 1;
}
// boff (nullable): pos is a chunk-relative scan, boff its chunk offsets (device_scan_chunks)
__global__ void scatter_rest_kernel(const float4* __restrict__ src, const int* __restrict__ flag, const int* __restrict__ pos,
                                    int n, float4* __restrict__ dst, const int* __restrict__ boff = nullptr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !flag[i]) dst[pos[i] + (boff ? boff[i >> SCAN_CHUNK_LOG2] : 0)] = src[i];
}
// createNewIdxMap (supports…:139-157), pass 1: classify. cat: 0 propagate, 1 level, 2 running counter
__global__ void idxmap_classify_kernel(const int* __restrict__ prev, int n0, const int* __restrict__ flag, int n_flag,
                                       int level, int* __restrict__ is_else) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n0) return;
  int v = prev[p];
  int e;
  if (v > level && v < 0) e = 0;
  else if (v >= 0 && v < n_flag && flag[v]) e = 0;
  else e = 1;
  is_else[p] = e;
}
__global__ void idxmap_write_kernel(const int* __restrict__ prev, int n0, const int* __restrict__ flag, int n_flag, int level,
                                    const int* __restrict__ else_pos, int* __restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n0) return;
  int v = prev[p];
  if (v > level && v < 0) out[p] = v;
  else if (v >= 0 && v < n_flag && flag[v]) out[p] = level;
  else out[p] = else_pos[p];
}
// getPointOnPlane (supports…:187-238) bounding box with the order dependent if / else-if scan:
// max = plain maximum; min = minimum over the points that did NOT raise the running maximum.
__global__ void copy_axis_kernel(const float4* __restrict__ pts, int m, int axis, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) out[i] = axis == 0 ? pts[i].x : pts[i].y;
}
// partial[b][0..4] = xMax, xMin(masked), yMax, yMin(masked), zSum of block b's grid-stride share; max / min are order
// independent, the z sum is a fixed tree for a given m (thread-strided partial sums, warp shuffle tree, warps then blocks
// in index order): deterministic. bbox_quirk_final_kernel folds the blocks. (One CTA alone took 0.19 ms on the 290 000
// support points of a full-resolution frame.)
constexpr int BBOX_BLOCKS = 64;
__global__ void __launch_bounds__(1024)
bbox_quirk_kernel(const float4* __restrict__ pts, int m, const float* __restrict__ pmax_x, const float* __restrict__ pmax_y,
                  double* __restrict__ partial /*[gridDim.x][5]*/) {
  __shared__ double s[5][32];
  double xMax = -INFINITY, xMin = INFINITY, yMax = -INFINITY, yMin = INFINITY, zs = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    float4 p = pts[i];
    if (p.x > pmax_x[i]) xMax = fmax(xMax, (double)p.x);  // raised the running maximum
    else if ((double)p.x < xMin) xMin = (double)p.x;
    if (p.y > pmax_y[i]) yMax = fmax(yMax, (double)p.y);
    else if ((double)p.y < yMin) yMin = (double)p.y;
    zs += (double)p.z;
  }
  double v[5] = {xMax, xMin, yMax, yMin, zs};
  for (int o = 16; o > 0; o >>= 1) {
    v[0] = fmax(v[0], __shfl_down_sync(0xffffffffu, v[0], o));
    v[1] = fmin(v[1], __shfl_down_sync(0xffffffffu, v[1], o));
    v[2] = fmax(v[2], __shfl_down_sync(0xffffffffu, v[2], o));
    v[3] = fmin(v[3], __shfl_down_sync(0xffffffffu, v[3], o));
    v[4] += __shfl_down_sync(0xffffffffu, v[4], o);
  }
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 5; ++k) s[k][threadIdx.x >> 5] = v[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    double r[5] = {-INFINITY, INFINITY, -INFINITY, INFINITY, 0.0};
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      r[0] = fmax(r[0], s[0][w]); r[1] = fmin(r[1], s[1][w]);
      r[2] = fmax(r[2], s[2][w]); r[3] = fmin(r[3], s[3][w]);
      r[4] += s[4][w];
    }
    for (int k = 0; k < 5; ++k) partial[blockIdx.x * 5 + k] = r[k];
  }
}
__global__ void bbox_quirk_final_kernel(const double* __restrict__ partial, int blocks, double* __restrict__ out /*xMax,xMin,yMax,yMin,zSum*/) {
  if (threadIdx.x != 0) return;
  double r[5] = {-INFINITY, INFINITY, -INFINITY, INFINITY, 0.0};
  for (int b = 0; b < blocks; ++b) {
    r[0] = fmax(r[0], partial[b * 5 + 0]); r[1] = fmin(r[1], partial[b * 5 + 1]);
    r[2] = fmax(r[2], partial[b * 5 + 2]); r[3] = fmin(r[3], partial[b * 5 + 3]);
    r[4] += partial[b * 5 + 4];
  }
  for (int k = 0; k < 5; ++k) out[k] = r[k];
}
__global__ void on_plane_flag_kernel(const float4* __restrict__ orig, const int* __restrict__ map, int n0, int level,
                                     double xMin, double xMax, double yMin, double yMax, double zMed, int* __restrict__ keep) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n0) return;
  float4 p = orig[i];
  bool k = (map[i] != level) && ((double)p.x > xMin && (double)p.x < xMax && (double)p.z > zMed && (double)p.y > yMin && (double)p.y < yMax);
  keep[i] = k ? 1 : 0;
}
__global__ void compact_points_kernel(const float4* __restrict__ src, const int* __restrict__ keep, const int* __restrict__ pos,
                                      int n, float4* __restrict__ dst, int* __restrict__ dst_idx, const int* __restrict__ boff = nullptr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && keep[i]) {
    const int q = pos[i] + (boff ? boff[i >> SCAN_CHUNK_LOG2] : 0);
    dst[q] = src[i];
    if (dst_idx) dst_idx[q] = i;
  }
}

// per-cluster centroid sums: one block per cluster, fixed-shape double tree (deterministic)
__global__ void __launch_bounds__(256)
cluster_centroid_kernel(const float4* __restrict__ pts, const int* __restrict__ idx, const int* __restrict__ offsets,
                        float* __restrict__ out /*[c][3] sums rounded to float*/) {
  __shared__ double s[3][8];
  const int c = blockIdx.x;
  const int b = offsets[c], e = offsets[c + 1];
  double sx = 0, sy = 0, sz = 0;
  for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
    float4 p = pts[idx[i]];
    sx += (double)p.x; sy += (double)p.y; sz += (double)p.z;
  }
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_down_sync(0xffffffffu, sx, o);
    sy += __shfl_down_sync(0xffffffffu, sy, o);
    sz += __shfl_down_sync(0xffffffffu, sz, o);
  }
  if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = sx; s[1][threadIdx.x >> 5] = sy; s[2][threadIdx.x >> 5] = sz; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[3] = {0, 0, 0};
    for (int w = 0; w < 8; ++w) { t[0] += s[0][w]; t[1] += s[1][w]; t[2] += s[2][w]; }
    out[3 * c] = (float)t[0]; out[3 * c + 1] = (float)t[1]; out[3 * c + 2] = (float)t[2];
  }
}

// ------------------------------------------------------------------ K11: axis extent (cylinder…:143-171, cone…:143-171)
struct AxisFrame {
  float a1x, a1y, a1z, dx, dy, dz, gdiv;  // A1, A1A2, |A1A2|^2 exactly as the reference computes them
};
__global__ void axis_project_kernel(const float4* __restrict__ pts, int n, AxisFrame f, float4* __restrict__ proj) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  float ax = p.x - f.a1x, ay = p.y - f.a1y, az = p.z - f.a1z;
  float G = (ax * f.dx + ay * f.dy + az * f.dz) / f.gdiv;
  proj[i] = make_float4(f.a1x + G * f.dx, f.a1y + G * f.dy, f.a1z + G * f.dz, 0.0f);
}
// all pairs (i > j): maximum of the float distances, first pair in (i, j) lexicographic order on ties
constexpr int AP_TPB = 256;
struct PairBest {
  float d;
  int i, j;
};
__device__ __forceinline__ bool pair_better(float d, int i, int j, const PairBest& b) {
  return d > b.d || (d == b.d && b.i >= 0 && (i < b.i || (i == b.i && j < b.j)));
}
__global__ void __launch_bounds__(AP_TPB)
axis_pairs_kernel(const float4* __restrict__ proj, int n, PairBest* __restrict__ block_best) {
  __shared__ float4 s_j[AP_TPB];
  __shared__ PairBest s_red[AP_TPB / 32];
  const int i = blockIdx.x * AP_TPB + threadIdx.x;
  const int jt = blockIdx.y;  // j tile
  PairBest best{-1.0f, -1, -1};
  if (jt * AP_TPB < (blockIdx.x + 1) * AP_TPB) {  // tile contains some j < max i of this block
    const int j0 = jt * AP_TPB;
    s_j[threadIdx.x] = (j0 + threadIdx.x < n) ? proj[j0 + threadIdx.x] : make_float4(0, 0, 0, 0);
    __syncthreads();
    if (i < n) {
      const float4 pi = proj[i];
      const int jend = min(min(AP_TPB, n - j0), i - j0);  // j < i
      for (int t = 0; t < jend; ++t) {
        const float4 pj = s_j[t];
        const float ddx = pi.x - pj.x, ddy = pi.y - pj.y, ddz = pi.z - pj.z;
        const float d = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
        if (d > best.d) { best.d = d; best.i = i; best.j = j0 + t; }  // ascending j: first max kept
      }
    }
  }
  // block reduction (ties -> lexicographically smallest (i, j))
  for (int o = 16; o > 0; o >>= 1) {
    PairBest ob;
    ob.d = __shfl_down_sync(0xffffffffu, best.d, o);
    ob.i = __shfl_down_sync(0xffffffffu, best.i, o);
    ob.j = __shfl_down_sync(0xffffffffu, best.j, o);
    if (ob.i >= 0 && (best.i < 0 || pair_better(ob.d, ob.i, ob.j, best))) best = ob;
  }
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    PairBest b = s_red[0];
    for (int w = 1; w < AP_TPB / 32; ++w) {
      PairBest ob = s_red[w];
      if (ob.i >= 0 && (b.i < 0 || pair_better(ob.d, ob.i, ob.j, b))) b = ob;
    }
    block_best[blockIdx.y * gridDim.x + blockIdx.x] = b;
  }
}
__global__ void __launch_bounds__(1024) axis_pairs_final_kernel(const PairBest* __restrict__ bb, int nb, const float4* __restrict__ proj,
                                                                float* __restrict__ out /*height, idx1, idx2 (as float bits), p1 xyz, p2 xyz*/) {
  __shared__ PairBest s_red[32];
  PairBest best{-1.0f, -1, -1};
  for (int t = threadIdx.x; t < nb; t += blockDim.x) {
    PairBest ob = bb[t];
    if (ob.i >= 0 && (best.i < 0 || pair_better(ob.d, ob.i, ob.j, best))) best = ob;
  }
  for (int o = 16; o > 0; o >>= 1) {
    PairBest ob;
    ob.d = __shfl_down_sync(0xffffffffu, best.d, o);
    ob.i = __shfl_down_sync(0xffffffffu, best.i, o);
    ob.j = __shfl_down_sync(0xffffffffu, best.j, o);
    if (ob.i >= 0 && (best.i < 0 || pair_better(ob.d, ob.i, ob.j, best))) best = ob;
  }
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    PairBest b = s_red[0];
    for (int w = 1; w < 32; ++w) {
      PairBest ob = s_red[w];
      if (ob.i >= 0 && (b.i < 0 || pair_better(ob.d, ob.i, ob.j, b))) b = ob;
    }
    out[0] = b.d;
    out[1] = __int_as_float(b.i);
    out[2] = __int_as_float(b.j);
    if (b.i >= 0) {
      float4 p1 = proj[b.i], p2 = proj[b.j];
      out[3] = p1.x; out[4] = p1.y; out[5] = p1.z; out[6] = p2.x; out[7] = p2.y; out[8] = p2.z;
    }
  }
}

struct AxisExtent {
  float height;
  int idx1, idx2;
  float p1[3], p2[3];
  float dirn[3];
};
static int axis_extent(pitt_ctx* ctx, const float4* d_pts, int n, const float* co, AxisExtent* out) {
  // getNormalizeAxesDirectionVector / getPointOnAxes / getVectorBetweenPoints in float, as the reference
  float norm = sqrtf(co[3] * co[3] + co[4] * co[4] + co[5] * co[5]);
  out->dirn[0] = co[3] / norm; out->dirn[1] = co[4] / norm; out->dirn[2] = co[5] / norm;
  const float t1 = -1.0f, t2 = +1.0f;
  float A1[3] = {co[0] + out->dirn[0] * t1, co[1] + out->dirn[1] * t1, co[2] + out->dirn[2] * t1};
  float A2[3] = {co[0] + out->dirn[0] * t2, co[1] + out->dirn[1] * t2, co[2] + out->dirn[2] * t2};
  AxisFrame f;
  f.a1x = A1[0]; f.a1y = A1[1]; f.a1z = A1[2];
  f.dx = A2[0] - A1[0]; f.dy = A2[1] - A1[1]; f.dz = A2[2] - A1[2];
  f.gdiv = f.dx * f.dx + f.dy * f.dy + f.dz * f.dz;
  out->height = -1.0f;
  out->idx1 = out->idx2 = -1;
  if (n <= 0) return PITT_OK;
  float4* d_proj = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_proj));
  axis_project_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_pts, n, f, d_proj);
  const int nt = cdiv(n, AP_TPB);
  PairBest* d_bb = nullptr;
  float* d_out = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)nt * nt, &d_bb));
  PITT_TRY(arena_alloc(ctx, 16, &d_out));
  axis_pairs_kernel<<<dim3(nt, nt), AP_TPB, 0, ctx->stream>>>(d_proj, n, d_bb);
  axis_pairs_final_kernel<<<1, 1024, 0, ctx->stream>>>(d_bb, nt * nt, d_proj, d_out);
  ctx->launches += 3;
  float h[9];
  PITT_CUDA(ctx, cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  out->height = h[0];
  memcpy(&out->idx1, &h[1], 4);
  memcpy(&out->idx2, &h[2], 4);
  for (int k = 0; k < 3; ++k) { out->p1[k] = h[3 + k]; out->p2[k] = h[6 + k]; }
  return PITT_OK;
}

// The same without the host in the loop: the axis frame is derived from the refined coefficients on the device (identical
// float operations), gated by the fit's result block (ints[0] best position, ints[3] final inliers): no model or no inliers
// -> height -1, as the reference leaves it.
__device__ __forceinline__ void axis_frame_dev(const float* co, AxisFrame& f, float* dirn) {
  const float norm = sqrtf(co[3] * co[3] + co[4] * co[4] + co[5] * co[5]);
  dirn[0] = co[3] / norm; dirn[1] = co[4] / norm; dirn[2] = co[5] / norm;
  const float t1 = -1.0f, t2 = +1.0f;
  const float A1[3] = {co[0] + dirn[0] * t1, co[1] + dirn[1] * t1, co[2] + dirn[2] * t1};
  const float A2[3] = {co[0] + dirn[0] * t2, co[1] + dirn[1] * t2, co[2] + dirn[2] * t2};
  f.a1x = A1[0]; f.a1y = A1[1]; f.a1z = A1[2];
  f.dx = A2[0] - A1[0]; f.dy = A2[1] - A1[1]; f.dz = A2[2] - A1[2];
  f.gdiv = f.dx * f.dx + f.dy * f.dy + f.dz * f.dz;
}
__device__ __forceinline__ bool axis_gate(const int* ints) { return ints[0] >= 0 && ints[3] > 0; }
__global__ void axis_project_dev_kernel(const float4* __restrict__ pts, int n, const float* __restrict__ co, const int* __restrict__ ints,
                                        float4* __restrict__ proj, const FitDesc* __restrict__ D = nullptr) {
  if (D) {  // batched: problem blockIdx.y
    const FitDesc d = D[blockIdx.y];
    pts = d.xyz; n = d.n; co = d.flt + 8; ints = d.ints; proj = d.proj;
  }
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !axis_gate(ints)) return;
  AxisFrame f;
  float dirn[3];
  axis_frame_dev(co, f, dirn);
  float4 p = pts[i];
  float ax = p.x - f.a1x, ay = p.y - f.a1y, az = p.z - f.a1z;
  float G = (ax * f.dx + ay * f.dy + az * f.dz) / f.gdiv;
  proj[i] = make_float4(f.a1x + G * f.dx, f.a1y + G * f.dy, f.a1z + G * f.dz, 0.0f);
}
__global__ void __launch_bounds__(AP_TPB)
axis_pairs_dev_kernel(const float4* __restrict__ proj, int n, const int* __restrict__ ints, PairBest* __restrict__ block_best,
                      const FitDesc* __restrict__ D = nullptr) {
  if (D) {  // batched: problem blockIdx.z
    const FitDesc d = D[blockIdx.z];
    proj = d.proj; n = d.n; ints = d.ints; block_best = reinterpret_cast<PairBest*>(d.bb);
  }
  if (!axis_gate(ints)) return;
  __shared__ float4 s_j[AP_TPB];
  __shared__ PairBest s_red[AP_TPB / 32];
  const int i = blockIdx.x * AP_TPB + threadIdx.x;
  const int jt = blockIdx.y;
  PairBest best{-1.0f, -1, -1};
  if (jt * AP_TPB < (blockIdx.x + 1) * AP_TPB) {
    const int j0 = jt * AP_TPB;
    s_j[threadIdx.x] = (j0 + threadIdx.x < n) ? proj[j0 + threadIdx.x] : make_float4(0, 0, 0, 0);
    __syncthreads();
    if (i < n) {
      const float4 pi = proj[i];
      const int jend = min(min(AP_TPB, n - j0), i - j0);
      for (int t = 0; t < jend; ++t) {
        const float4 pj = s_j[t];
        const float ddx = pi.x - pj.x, ddy = pi.y - pj.y, ddz = pi.z - pj.z;
        const float d = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
        if (d > best.d) { best.d = d; best.i = i; best.j = j0 + t; }
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    PairBest ob;
    ob.d = __shfl_down_sync(0xffffffffu, best.d, o);
    ob.i = __shfl_down_sync(0xffffffffu, best.i, o);
    ob.j = __shfl_down_sync(0xffffffffu, best.j, o);
    if (ob.i >= 0 && (best.i < 0 || pair_better(ob.d, ob.i, ob.j, best))) best = ob;
  }
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    PairBest b = s_red[0];
    for (int w = 1; w < AP_TPB / 32; ++w) {
      PairBest ob = s_red[w];
      if (ob.i >= 0 && (b.i < 0 || pair_better(ob.d, ob.i, ob.j, b))) b = ob;
    }
    block_best[blockIdx.y * gridDim.x + blockIdx.x] = b;
  }
}
// out: [0] height, [1] idx1, [2] idx2 (int bits), [3..5] p1, [6..8] p2, [9..11] normalised axis direction
__global__ void __launch_bounds__(1024) axis_pairs_final_dev_kernel(const PairBest* __restrict__ bb, int nb, const float4* __restrict__ proj,
                                                                    const float* __restrict__ co, const int* __restrict__ ints,
                                                                    float* __restrict__ out, const FitDesc* __restrict__ D = nullptr) {
  __shared__ PairBest s_red[32];
  if (D) {  // batched: problem blockIdx.x
    const FitDesc d = D[blockIdx.x];
    bb = reinterpret_cast<const PairBest*>(d.bb); proj = d.proj; co = d.flt + 8; ints = d.ints; out = d.flt + 16;
  }
  if (!axis_gate(ints)) {
    if (threadIdx.x == 0) { out[0] = -1.0f; out[1] = __int_as_float(-1); out[2] = __int_as_float(-1); }
    return;
  }
  PairBest best{-1.0f, -1, -1};
  for (int t = threadIdx.x; t < nb; t += blockDim.x) {
    PairBest ob = bb[t];
    if (ob.i >= 0 && (best.i < 0 || pair_better(ob.d, ob.i, ob.j, best))) best = ob;
  }
  for (int o = 16; o > 0; o >>= 1) {
    PairBest ob;
    ob.d = __shfl_down_sync(0xffffffffu, best.d, o);
    ob.i = __shfl_down_sync(0xffffffffu, best.i, o);
    ob.j = __shfl_down_sync(0xffffffffu, best.j, o);
    if (ob.i >= 0 && (best.i < 0 || pair_better(ob.d, ob.i, ob.j, best))) best = ob;
  }
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    PairBest b = s_red[0];
    for (int w = 1; w < 32; ++w) {
      PairBest ob = s_red[w];
      if (ob.i >= 0 && (b.i < 0 || pair_better(ob.d, ob.i, ob.j, b))) b = ob;
    }
    out[0] = b.d;
    out[1] = __int_as_float(b.i);
    out[2] = __int_as_float(b.j);
    if (b.i >= 0) {
      float4 p1 = proj[b.i], p2 = proj[b.j];
      out[3] = p1.x; out[4] = p1.y; out[5] = p1.z; out[6] = p2.x; out[7] = p2.y; out[8] = p2.z;
    }
    AxisFrame f;
    float dirn[3];
    axis_frame_dev(co, f, dirn);
    out[9] = dirn[0]; out[10] = dirn[1]; out[11] = dirn[2];
  }
}
static int axis_extent_async(pitt_ctx* ctx, const float4* d_pts, int n, const float* d_co, const int* d_ints, float* d_out16) {
  if (n <= 0) return PITT_OK;
  float4* d_proj = nullptr;
  PairBest* d_bb = nullptr;
  const int nt = cdiv(n, AP_TPB);
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_proj));
  PITT_TRY(arena_alloc(ctx, (size_t)nt * nt, &d_bb));
  axis_project_dev_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_pts, n, d_co, d_ints, d_proj);
  axis_pairs_dev_kernel<<<dim3(nt, nt), AP_TPB, 0, ctx->stream>>>(d_proj, n, d_ints, d_bb);
  axis_pairs_final_dev_kernel<<<1, 1024, 0, ctx->stream>>>(d_bb, nt * nt, d_proj, d_co, d_ints, d_out16);
  ctx->launches += 3;
  PITT_CUDA(ctx, cudaGetLastError());
  return PITT_OK;
}

// ------------------------------------------------------------------ supports
struct SupportCfg {
  float minCloudPct, minPlanePct, maxVar, minVar, thr, w;
  int maxIter;
  float axis[3], offset[3];
};
static SupportCfg resolve_support(const pitt_support_params& p) {
  SupportCfg c;
  // srvm::getServiceFloatParameter / IntParameter / 3DArrayParameter (srv_manager.h:163-188)
  c.minCloudPct = p.min_iterative_cloud_percentual_size >= 0.0f ? p.min_iterative_cloud_percentual_size : 0.030f;
  c.minPlanePct = p.min_iterative_plane_percentual_size >= 0.0f ? p.min_iterative_plane_percentual_size : 0.030f;
  c.maxVar = p.variance_threshold_for_horizontal >= 0.0f ? p.variance_threshold_for_horizontal : 0.09f;
  c.minVar = -1 * c.maxVar;
  c.thr = p.ransac_distance_point_in_shape_threshold >= 0.0f ? p.ransac_distance_point_in_shape_threshold : 0.02f;
  c.w = p.ransac_model_normal_distance_weigth >= 0.0f ? p.ransac_model_normal_distance_weigth : 0.9f;
  c.maxIter = p.ransac_max_iteration_threshold >= 0 ? p.ransac_max_iteration_threshold : 10;
  const float defAxis[3] = {0.0f, 0.0f, -1.0f};
  const float defOff[3] = {0.02f, 0.02f, 0.005f};
  for (int i = 0; i < 3; ++i) {
    c.axis[i] = p.horizontal_axis_len == 3 ? p.horizontal_axis[i] : defAxis[i];
    c.offset[i] = p.support_edge_remove_offset_len == 3 ? p.support_edge_remove_offset[i] : defOff[i];
  }
  return c;
}
static bool is_horizontal_plane(const float* co, const SupportCfg& c) {  // supports…:161-179
  float div = sqrtf(co[0] * co[0] + co[1] * co[1] + co[2] * co[2]);
  float nx = co[0] / div, ny = co[1] / div, nz = co[2] / div;
  float crossX = ny * c.axis[2] - nz * c.axis[1];
  float crossY = nz * c.axis[0] - nx * c.axis[2];
  float crossZ = nx * c.axis[1] - ny * c.axis[0];
  return ((crossX > c.minVar) && (crossX < c.maxVar)) && ((crossY > c.minVar) && (crossY < c.maxVar)) &&
         ((crossZ > c.minVar) && (crossZ < c.maxVar));
}

struct SupportDev {
  float co[4];
  int n_support, n_on;
  const int* d_map;          // N0 labels
  const float4* d_support;   // support cloud
  const float4* d_on;        // on-support cloud (points of the original cloud, original order)
};

static int stage_reserve(pitt_ctx* ctx, size_t bytes);

// ---- device-side halves of one findSupports trip (no host round trip between them)
__global__ void gather_points_dev_kernel(const float4* __restrict__ src, const int* __restrict__ idx, const int* __restrict__ m,
                                         float4* __restrict__ dst, float* __restrict__ px, float* __restrict__ py) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < *m) {
    const float4 p = src[idx[i]];
    dst[i] = p;
    px[i] = p.x;  // running-maximum scan inputs of getPointOnPlane
    py[i] = p.y;
  }
}
// isHorizontalPlane (supports…:161-179) on the device + the label of this trip: out[0] = horizontal, out[1] = level
__global__ void support_decide_kernel(const float* __restrict__ co, SupportCfg c, int idxMapLayer, int* __restrict__ out) {
  float div = sqrtf(co[0] * co[0] + co[1] * co[1] + co[2] * co[2]);
  float nx = co[0] / div, ny = co[1] / div, nz = co[2] / div;
  float crossX = ny * c.axis[2] - nz * c.axis[1];
  float crossY = nz * c.axis[0] - nx * c.axis[2];
  float crossZ = nx * c.axis[1] - ny * c.axis[0];
  const bool h = ((crossX > c.minVar) && (crossX < c.maxVar)) && ((crossY > c.minVar) && (crossY < c.maxVar)) &&
                 ((crossZ > c.minVar) && (crossZ < c.maxVar));
  out[0] = h ? 1 : 0;
  out[1] = h ? idxMapLayer : -1;
}
__global__ void idxmap_classify_dev_kernel(const int* __restrict__ prev, int n0, const int* __restrict__ flag, int n_flag,
                                           const int* __restrict__ decide, int* __restrict__ is_else) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n0) return;
  const int level = decide[1];
  int v = prev ? prev[p] : p;  // first trip: the identity map
  int e;
  if (v > level && v < 0) e = 0;
  else if (v >= 0 && v < n_flag && flag[v]) e = 0;
  else e = 1;
  is_else[p] = e;
}
__global__ void idxmap_write_dev_kernel(const int* __restrict__ prev, int n0, const int* __restrict__ flag, int n_flag,
                                        const int* __restrict__ decide, const int* __restrict__ else_pos, const int* __restrict__ boff,
                                        int* __restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n0) return;
  const int level = decide[1];
  int v = prev ? prev[p] : p;
  if (v > level && v < 0) out[p] = v;
  else if (v >= 0 && v < n_flag && flag[v]) out[p] = level;
  else out[p] = else_pos[p] + boff[p >> SCAN_CHUNK_LOG2];
}
__global__ void __launch_bounds__(1024)
bbox_quirk_dev_kernel(const float4* __restrict__ pts, const int* __restrict__ m_ptr, const int* __restrict__ decide,
                      const float* __restrict__ pmax_x, const float* __restrict__ pmax_y, const float* __restrict__ boff_x,
                      const float* __restrict__ boff_y, double* __restrict__ partial) {
  __shared__ double s[5][32];
  if (!decide[0]) return;  // not a horizontal plane: no getPointOnPlane
  const int m = *m_ptr;
  double xMax = -INFINITY, xMin = INFINITY, yMax = -INFINITY, yMin = INFINITY, zs = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    float4 p = pts[i];
    const float rmx = fmaxf(pmax_x[i], boff_x[i >> SCAN_CHUNK_LOG2]), rmy = fmaxf(pmax_y[i], boff_y[i >> SCAN_CHUNK_LOG2]);  // running maxima
    if (p.x > rmx) xMax = fmax(xMax, (double)p.x);
    else if ((double)p.x < xMin) xMin = (double)p.x;
    if (p.y > rmy) yMax = fmax(yMax, (double)p.y);
    else if ((double)p.y < yMin) yMin = (double)p.y;
    zs += (double)p.z;
  }
  double v[5] = {xMax, xMin, yMax, yMin, zs};
  for (int o = 16; o > 0; o >>= 1) {
    v[0] = fmax(v[0], __shfl_down_sync(0xffffffffu, v[0], o));
    v[1] = fmin(v[1], __shfl_down_sync(0xffffffffu, v[1], o));
    v[2] = fmax(v[2], __shfl_down_sync(0xffffffffu, v[2], o));
    v[3] = fmin(v[3], __shfl_down_sync(0xffffffffu, v[3], o));
    v[4] += __shfl_down_sync(0xffffffffu, v[4], o);
  }
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 5; ++k) s[k][threadIdx.x >> 5] = v[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    double r[5] = {-INFINITY, INFINITY, -INFINITY, INFINITY, 0.0};
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      r[0] = fmax(r[0], s[0][w]); r[1] = fmin(r[1], s[1][w]);
      r[2] = fmax(r[2], s[2][w]); r[3] = fmin(r[3], s[3][w]);
      r[4] += s[4][w];
    }
    for (int k = 0; k < 5; ++k) partial[blockIdx.x * 5 + k] = r[k];
  }
}
// folds the blocks and applies the offsets (supports…:222-228): out = xMin, xMax, yMin, yMax, zMed
__global__ void bbox_quirk_final_dev_kernel(const double* __restrict__ partial, int blocks, const int* __restrict__ m_ptr,
                                            const int* __restrict__ decide, SupportCfg c, double* __restrict__ out) {
  if (threadIdx.x != 0 || !decide[0]) return;
  double r[5] = {-INFINITY, INFINITY, -INFINITY, INFINITY, 0.0};
  for (int b = 0; b < blocks; ++b) {
    r[0] = fmax(r[0], partial[b * 5 + 0]); r[1] = fmin(r[1], partial[b * 5 + 1]);
    r[2] = fmax(r[2], partial[b * 5 + 2]); r[3] = fmin(r[3], partial[b * 5 + 3]);
    r[4] += partial[b * 5 + 4];
  }
  double xMax = r[0], xMin = r[1], yMax = r[2], yMin = r[3], zMed = r[4];
  xMax -= c.offset[0];
  xMin += c.offset[0];
  yMax -= c.offset[1];
  yMin += c.offset[1];
  zMed = zMed / *m_ptr + c.offset[2];
  out[0] = xMin; out[1] = xMax; out[2] = yMin; out[3] = yMax; out[4] = zMed;
}
__global__ void on_plane_flag_dev_kernel(const float4* __restrict__ orig, const int* __restrict__ map, int n0, const int* __restrict__ decide,
                                         const double* __restrict__ bb, int* __restrict__ flag, int* __restrict__ keep) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n0) return;
  bool k = false;
  if (decide[0]) {
    const int level = decide[1];
    const double xMin = bb[0], xMax = bb[1], yMin = bb[2], yMax = bb[3], zMed = bb[4];
    float4 p = orig[i];
    k = (map[i] != level) && ((double)p.x > xMin && (double)p.x < xMax && (double)p.z > zMed && (double)p.y > yMin && (double)p.y < yMax);
  }
  flag[i] = k ? 1 : 0;
  keep[i] = k ? 1 : 0;
}

struct TripDev {        // device outputs of the second half of a trip
  int* d_decide;        // [0] horizontal, [1] level, [2] points on the support
  int* d_new_map;
  float4* d_support;
  float4* d_rest;
  float4* d_on;
};
// removePlaneInliner + createNewIdxMap + getPointOnPlane of one trip from device-resident inliers (d_inl, *d_n_inl ascending
// indices into the trip's cloud of ni points) and coefficients; every size that depends on the RANSAC result stays on the device
static int supports_trip_second_half(pitt_ctx* ctx, const pitt_cloud* cloud, const float4* d_iter, int ni, const int* d_prev_map,
                                     const int* d_inl, const int* d_n_inl, const float* d_co, const SupportCfg& c, int idxMapLayer,
                                     TripDev* T) {
  const int n0 = cloud->n;
  int* d_flag = nullptr;
  int* d_keep = nullptr;
  float* d_px = nullptr;
  float* d_py = nullptr;
  int* d_else = nullptr;
  double* d_part = nullptr;
  double* d_bb = nullptr;
  int* d_onflag = nullptr;
  int* d_onkeep = nullptr;
  PITT_TRY(arena_alloc(ctx, 4, &T->d_decide));
  PITT_TRY(arena_alloc(ctx, (size_t)ni, &d_flag));
  PITT_TRY(arena_alloc(ctx, (size_t)ni, &d_keep));
  PITT_TRY(arena_alloc(ctx, (size_t)ni, &T->d_support));
  PITT_TRY(arena_alloc(ctx, (size_t)ni, &T->d_rest));
  PITT_TRY(arena_alloc(ctx, (size_t)ni, &d_px));
  PITT_TRY(arena_alloc(ctx, (size_t)ni, &d_py));
  PITT_TRY(arena_alloc(ctx, (size_t)n0, &d_else));
  PITT_TRY(arena_alloc(ctx, (size_t)n0, &T->d_new_map));
  PITT_TRY(arena_alloc(ctx, (size_t)BBOX_BLOCKS * 5, &d_part));
  PITT_TRY(arena_alloc(ctx, 8, &d_bb));
  PITT_TRY(arena_alloc(ctx, (size_t)n0, &d_onflag));
  PITT_TRY(arena_alloc(ctx, (size_t)n0 + 1, &d_onkeep));
  PITT_TRY(arena_alloc(ctx, (size_t)n0, &T->d_on));
  unsigned* d_tickets = nullptr;
  int* boff_keep = nullptr;
  int* boff_else = nullptr;
  int* boff_on = nullptr;
  float* boff_x = nullptr;
  float* boff_y = nullptr;
  PITT_TRY(arena_alloc(ctx, 8, &d_tickets));
  PITT_CUDA(ctx, cudaMemsetAsync(d_tickets, 0, 8 * sizeof(unsigned), ctx->stream));
  PITT_CUDA(ctx, cudaMemsetAsync(d_flag, 0, (size_t)ni * sizeof(int), ctx->stream));
  set_flags_kernel<<<cdiv(ni, 256), 256, 0, ctx->stream>>>(d_inl, d_n_inl, d_flag);
  gather_points_dev_kernel<<<cdiv(ni, 256), 256, 0, ctx->stream>>>(d_iter, d_inl, d_n_inl, T->d_support, d_px, d_py);
  invert_flags_kernel<<<cdiv(ni, 256), 256, 0, ctx->stream>>>(d_flag, ni, d_keep);
  ctx->launches += 3;
  // every prefix scan is one launch (chunk-relative scan + chunk offsets by the last CTA); the consumers add the chunk offset
  PITT_TRY(device_scan_chunks(ctx, d_keep, ni, &boff_keep, d_tickets + 0, nullptr));
  scatter_rest_kernel<<<cdiv(ni, 256), 256, 0, ctx->stream>>>(d_iter, d_flag, d_keep, ni, T->d_rest, boff_keep);
  support_decide_kernel<<<1, 1, 0, ctx->stream>>>(d_co, c, idxMapLayer, T->d_decide);
  idxmap_classify_dev_kernel<<<cdiv(n0, 256), 256, 0, ctx->stream>>>(d_prev_map, n0, d_flag, ni, T->d_decide, d_else);
  ctx->launches += 3;
  PITT_TRY(device_scan_chunks(ctx, d_else, n0, &boff_else, d_tickets + 1, nullptr));
  idxmap_write_dev_kernel<<<cdiv(n0, 256), 256, 0, ctx->stream>>>(d_prev_map, n0, d_flag, ni, T->d_decide, d_else, boff_else, T->d_new_map);
  ctx->launches++;
  // getPointOnPlane (only does anything when the plane is horizontal: the kernels check the flag themselves). The running
  // maxima are scanned over ni entries; the tail beyond the inlier count never reaches an entry that is read.
  PITT_TRY(device_max_scan_chunks(ctx, d_px, ni, &boff_x, d_tickets + 2));
  PITT_TRY(device_max_scan_chunks(ctx, d_py, ni, &boff_y, d_tickets + 3));
  const int bb_blocks = std::max(1, std::min(BBOX_BLOCKS, cdiv(ni, 4096)));
  bbox_quirk_dev_kernel<<<bb_blocks, 1024, 0, ctx->stream>>>(T->d_support, d_n_inl, T->d_decide, d_px, d_py, boff_x, boff_y, d_part);
  bbox_quirk_final_dev_kernel<<<1, 32, 0, ctx->stream>>>(d_part, bb_blocks, d_n_inl, T->d_decide, c, d_bb);
  on_plane_flag_dev_kernel<<<cdiv(n0, 256), 256, 0, ctx->stream>>>(cloud->d_xyz, T->d_new_map, n0, T->d_decide, d_bb, d_onflag, d_onkeep);
  ctx->launches += 3;
  PITT_TRY(device_scan_chunks(ctx, d_onkeep, n0, &boff_on, d_tickets + 4, T->d_decide + 2));
  compact_points_kernel<<<cdiv(n0, 256), 256, 0, ctx->stream>>>(cloud->d_xyz, d_onflag, d_onkeep, n0, T->d_on, nullptr, boff_on);
  ctx->launches++;
  PITT_CUDA(ctx, cudaGetLastError());
  return PITT_OK;
}

// findSupports (supports…:241-361): ONE host synchronisation per trip. The plane RANSAC of a trip (sample table, estimate,
// score, PCL stop rule, refinement, final inliers) and the whole second half of the trip are enqueued back to back; the host then
// reads one small block (inlier count, coefficients, horizontal flag, points on the support, scan flags) and decides whether the
// loop goes on. The second half of the LAST trip is wasted work (the loop always ends with a trip whose plane is rejected), on a
// cloud that has shrunk to the objects by then.
static int find_supports_impl(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_support_params& params, SupportCfg* cfg_out,
                              std::vector<SupportDev>* out, int* loop_trips) {
  const SupportCfg c = resolve_support(params);
  if (cfg_out) *cfg_out = c;
  out->clear();
  *loop_trips = 0;
  const int n0 = cloud->n;
  if (n0 <= 0) return PITT_OK;
  pitt_sac_params sp;
  pitt_default_support_sac_params(&sp);
  sp.distance_threshold = (double)c.thr;
  sp.normal_distance_weight = (double)c.w;
  sp.max_iterations = c.maxIter;

  pitt_cloud iter;  // non-owning view of the shrinking cloud
  iter.n = n0;
  iter.d_xyz = cloud->d_xyz;
  iter.h_valid = false;
  const int* d_prev_map = nullptr;  // nullptr = identity (first trip)
  int idxMapLayer = -2;
  const int k = params.normals_k > 0 ? params.normals_k : 50;
  const size_t stage_words = (size_t)(std::max(c.maxIter, 0) + 1) * 3 + 64;
  for (;;) {
    const int ni = iter.n;
    // ---- first half: seg.segment() of the trip's cloud, enqueued
    PITT_TRY(stage_reserve(ctx, stage_words * 4));
    int* h_stage = reinterpret_cast<int*>(ctx->h_stage);
    SacAsync sa;
    sa.d_ints = nullptr;
    sa.d_flt = nullptr;
    bool issued = false;
    PITT_TRY(sac_segment_async(ctx, &iter, sp, h_stage + 64, &sa, &issued));
    (*loop_trips)++;
    if (!issued) break;  // no solution possible (fewer points than the sample size): n_inl == 0
    // the size test of the reference comes AFTER its RANSAC call and needs nothing from it: no second half for a trip it rejects
    const bool cloud_too_small = (float)ni < (float)n0 * c.minCloudPct;
    TripDev T;
    memset(&T, 0, sizeof(T));
    if (!cloud_too_small)
      PITT_TRY(supports_trip_second_half(ctx, cloud, iter.d_xyz, ni, d_prev_map, sa.d_inl, sa.d_ints + 3, sa.d_flt + 8, c, idxMapLayer, &T));
    PITT_CUDA(ctx, cudaMemcpyAsync(h_stage, sa.d_ints, 16 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, cudaMemcpyAsync(h_stage + 16, sa.d_flt, 16 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (!cloud_too_small) PITT_CUDA(ctx, cudaMemcpyAsync(h_stage + 32, T.d_decide, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
    int n_inl = 0;
    float co[4] = {0.f, 0.f, 0.f, 0.f};
    bool horizontal = false;
    int n_on = 0;
    if (h_stage[8] != 0) {
      // the device-side stop rule raised a flag: repeat the trip's RANSAC on the synchronous path and its second half from that
      SacDeviceResult r;
      PITT_TRY(sac_segment_impl(ctx, &iter, sp, &r));
      n_inl = r.n_inliers;
      if (n_inl > 0 && !cloud_too_small) {
        int* d_n = nullptr;
        float* d_co = nullptr;
        PITT_TRY(arena_alloc(ctx, 1, &d_n));
        PITT_TRY(arena_alloc(ctx, 8, &d_co));
        h_stage[40] = n_inl;
        memcpy(h_stage + 44, r.coeffs, 4 * sizeof(float));
        PITT_CUDA(ctx, cudaMemcpyAsync(d_n, h_stage + 40, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        PITT_CUDA(ctx, cudaMemcpyAsync(d_co, h_stage + 44, 4 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        PITT_TRY(supports_trip_second_half(ctx, cloud, iter.d_xyz, ni, d_prev_map, r.d_inliers, d_n, d_co, c, idxMapLayer, &T));
        PITT_CUDA(ctx, cudaMemcpyAsync(h_stage + 32, T.d_decide, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PITT_CUDA(ctx, pitt::stream_sync(ctx));
      }
      for (int i = 0; i < 4; ++i) co[i] = r.coeffs[i];
    } else if (h_stage[0] >= 0) {
      n_inl = h_stage[3];
      memcpy(co, h_stage + 16 + 8, 4 * sizeof(float));
    }
    if (n_inl == 0) break;
    else if (cloud_too_small) break;
    else if ((float)n_inl < (float)n0 * c.minPlanePct) break;
    horizontal = h_stage[32] != 0;
    n_on = h_stage[34];
    if (params.compute_discarded_normals) {
      // supports…:297,300 — results are never used by the reference; offered for cost fidelity only
      const float vp[3] = {0.f, 0.f, 0.f};
      float4* d_scratch = nullptr;
      PITT_TRY(arena_alloc(ctx, (size_t)std::max(ni, 1), &d_scratch));
      PITT_TRY(estimate_normals_impl(ctx, T.d_rest, ni - n_inl, k, vp, d_scratch));
      PITT_TRY(estimate_normals_impl(ctx, T.d_support, n_inl, k, vp, d_scratch));
    }
    if (horizontal) {
      SupportDev S;
      for (int i = 0; i < 4; ++i) S.co[i] = co[i];
      S.n_support = n_inl;
      S.n_on = n_on;
      S.d_map = T.d_new_map;
      S.d_support = T.d_support;
      S.d_on = T.d_on;
      out->push_back(S);
    }
    d_prev_map = T.d_new_map;
    iter.d_xyz = T.d_rest;
    iter.n = ni - n_inl;
    iter.h_valid = false;
    iter.h_xyz.clear();
    idxMapLayer--;
    if (iter.n <= 0) {
      // the reference would call RANSAC on an empty cloud, get no inliers and stop
      (*loop_trips)++;
      break;
    }
  }
  PITT_CUDA(ctx, cudaGetLastError());
  return PITT_OK;
}

// ------------------------------------------------------------------ clusters
struct ClustersDev {
  std::vector<int> sizes;     // per cluster (PCL order)
  std::vector<int> offsets;   // sizes.size()+1
  std::vector<int> indices;   // host copy, ascending per cluster
  std::vector<float> centroid;// 3 per cluster (sum/(n+1))
  const float4* d_points = nullptr;  // cluster clouds, contiguous in `offsets` order (device)
  const int* d_head = nullptr;       // device copy of {nc, ., sizes, offsets} (device-side clustering only)
  const int* d_labels = nullptr;     // debugging aid
  const int* d_idx = nullptr;
};
static int cluster_service_host_impl(pitt_ctx* ctx, const float4* d_xyz, int n, const pitt_cluster_params& p, ClustersDev* out) {
  out->sizes.clear();
  out->offsets.assign(1, 0);
  out->indices.clear();
  out->centroid.clear();
  out->d_points = nullptr;
  if (!(n >= p.min_input_size) || n <= 0) return PITT_OK;
  const int min_sz = (int)round((double)n * p.min_rate);
  const int max_sz = (int)round((double)n * p.max_rate);
  int* d_labels = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_labels));
  std::vector<int> sizes;
  PITT_TRY(euclidean_clusters_impl(ctx, d_xyz, n, p.tolerance, min_sz, max_sz, d_labels, &sizes));
  const int nc = (int)sizes.size();
  if (nc == 0) return PITT_OK;
  std::vector<int> labels(n);
  PITT_CUDA(ctx, cudaMemcpyAsync(labels.data(), d_labels, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  out->sizes = sizes;
  out->offsets.resize(nc + 1);
  out->offsets[0] = 0;
  for (int c = 0; c < nc; ++c) out->offsets[c + 1] = out->offsets[c] + sizes[c];
  const int total = out->offsets[nc];
  out->indices.resize(total);
  std::vector<int> cur(out->offsets.begin(), out->offsets.end() - 1);
  for (int i = 0; i < n; ++i)
    if (labels[i] >= 0) out->indices[cur[labels[i]]++] = i;  // ascending within each cluster
  int* d_idx = nullptr;
  int* d_off = nullptr;
  float4* d_pts = nullptr;
  float* d_cen = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)total, &d_idx));
  PITT_TRY(arena_alloc(ctx, (size_t)nc + 1, &d_off));
  PITT_TRY(arena_alloc(ctx, (size_t)total, &d_pts));
  PITT_TRY(arena_alloc(ctx, (size_t)nc * 3, &d_cen));
  PITT_CUDA(ctx, cudaMemcpyAsync(d_idx, out->indices.data(), (size_t)total * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  PITT_CUDA(ctx, cudaMemcpyAsync(d_off, out->offsets.data(), (size_t)(nc + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  gather_points_kernel<<<cdiv(total, 256), 256, 0, ctx->stream>>>(d_xyz, d_idx, total, d_pts);
  cluster_centroid_kernel<<<nc, 256, 0, ctx->stream>>>(d_xyz, d_idx, d_off, d_cen);
  ctx->launches += 2;
  std::vector<float> sums((size_t)nc * 3);
  PITT_CUDA(ctx, cudaMemcpyAsync(sums.data(), d_cen, sums.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  out->centroid.resize((size_t)nc * 3);
  for (int c = 0; c < nc; ++c) {
    int cntp1 = 1 + sizes[c];  // `int cnt = 1; ... cnt++` (cluster…:78,91)
    for (int a = 0; a < 3; ++a) out->centroid[3 * c + a] = sums[3 * c + a] / cntp1;
  }
  out->d_points = d_pts;
  return PITT_OK;
}

__global__ void __launch_bounds__(256)
cluster_centroid_dev_kernel(const float4* __restrict__ pts, const int* __restrict__ head, float* __restrict__ out /*[c][3]*/) {
  __shared__ double s[3][8];
  const int c = blockIdx.x;
  if (c >= head[0]) return;
  const int b = head[2 + CC_MAXC + c], e = head[2 + CC_MAXC + c + 1];
  double sx = 0, sy = 0, sz = 0;
  for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
    float4 p = pts[i];  // the cluster clouds lie back to back in index order
    sx += (double)p.x; sy += (double)p.y; sz += (double)p.z;
  }
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_down_sync(0xffffffffu, sx, o);
    sy += __shfl_down_sync(0xffffffffu, sy, o);
    sz += __shfl_down_sync(0xffffffffu, sz, o);
  }
  if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = sx; s[1][threadIdx.x >> 5] = sy; s[2][threadIdx.x >> 5] = sz; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[3] = {0, 0, 0};
    for (int w = 0; w < 8; ++w) { t[0] += s[0][w]; t[1] += s[1][w]; t[2] += s[2][w]; }
    out[3 * c] = (float)t[0]; out[3 * c + 1] = (float)t[1]; out[3 * c + 2] = (float)t[2];
  }
}
// clusterize (cluster…:38-108) with ONE host synchronisation: components, PCL ordering, index lists, cluster clouds and centroid
// sums are made on the device (cluster.cu: euclidean_clusters_dev); the host reads the cluster count, the sizes and the sums.
// want_indices: also fetch the index lists (the service response needs them, the frame path does not). More than CC_MAXC
// clusters (min_rate below 1/128): the host-ordered path.
static int cluster_service_impl(pitt_ctx* ctx, const float4* d_xyz, int n, const pitt_cluster_params& p, ClustersDev* out,
                                bool want_indices = true) {
  out->sizes.clear();
  out->offsets.assign(1, 0);
  out->indices.clear();
  out->centroid.clear();
  out->d_points = nullptr;
  out->d_head = nullptr;
  if (!(n >= p.min_input_size) || n <= 0) return PITT_OK;
  const int min_sz = (int)round((double)n * p.min_rate);
  const int max_sz = (int)round((double)n * p.max_rate);
  {
    static int host_path = -1;  // PITT_DEBUG_CC_HOST=1: the host-ordered clustering (debugging aid)
    if (host_path < 0) { const char* v = getenv("PITT_DEBUG_CC_HOST"); host_path = (v && v[0] == '1') ? 1 : 0; }
    if (host_path) return cluster_service_host_impl(ctx, d_xyz, n, p, out);
  }
  ClustersOnDevice cd;
  PITT_TRY(euclidean_clusters_dev(ctx, d_xyz, n, p.tolerance, min_sz, max_sz, &cd));
  float* d_cen = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)CC_MAXC * 3, &d_cen));
  cluster_centroid_dev_kernel<<<CC_MAXC, 256, 0, ctx->stream>>>(cd.d_points, cd.d_head, d_cen);
  ctx->launches++;
  PITT_TRY(stage_reserve(ctx, (size_t)(CC_HEAD_INTS + CC_MAXC * 3) * 4));
  int* h_head = reinterpret_cast<int*>(ctx->h_stage);
  float* h_cen = reinterpret_cast<float*>(h_head + CC_HEAD_INTS);
  PITT_CUDA(ctx, cudaMemcpyAsync(h_head, cd.d_head, CC_HEAD_INTS * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, cudaMemcpyAsync(h_cen, d_cen, (size_t)CC_MAXC * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  if (h_head[1] > CC_MAXC) return cluster_service_host_impl(ctx, d_xyz, n, p, out);
  const int nc = h_head[0];
  if (nc == 0) return PITT_OK;
  out->sizes.assign(h_head + 2, h_head + 2 + nc);
  out->offsets.assign(h_head + 2 + CC_MAXC, h_head + 2 + CC_MAXC + nc + 1);
  out->centroid.resize((size_t)nc * 3);
  for (int c = 0; c < nc; ++c) {
    const int cntp1 = 1 + out->sizes[c];  // `int cnt = 1; ... cnt++` (cluster…:78,91)
    for (int a = 0; a < 3; ++a) out->centroid[3 * c + a] = h_cen[3 * c + a] / cntp1;
  }
  out->d_points = cd.d_points;
  out->d_head = cd.d_head;
  out->d_labels = cd.d_labels;
  out->d_idx = cd.d_idx;
  if (want_indices) {
    const int total = out->offsets[nc];
    out->indices.resize(total);
    PITT_CUDA(ctx, cudaMemcpyAsync(out->indices.data(), cd.d_idx, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
  }
  return PITT_OK;
}

// ------------------------------------------------------------------ primitives
struct PrimitiveHost {
  SacDeviceResult sac;
  int n_coefficients;
  float coefficients[8];
  float centroid[3];
  int centroid_valid;
};
static int primitive_service_impl(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params& p, PrimitiveHost* out) {
  PITT_TRY(sac_segment_impl(ctx, c, p, &out->sac));
  const SacDeviceResult& r = out->sac;
  out->n_coefficients = r.n_coeffs;
  for (int i = 0; i < 8; ++i) out->coefficients[i] = i < r.n_coeffs ? r.coeffs[i] : 0.0f;
  out->centroid[0] = out->centroid[1] = out->centroid[2] = 0.0f;
  out->centroid_valid = 0;
  if (p.model == PITT_MODEL_SPHERE) {
    if (r.n_coeffs > 0) {
      for (int a = 0; a < 3; ++a) out->centroid[a] = r.coeffs[a];
      out->centroid_valid = 1;
    }
  } else if (p.model == PITT_MODEL_CYLINDER || p.model == PITT_MODEL_CONE) {
    float height = -1.0f;
    if (r.n_inliers > 0) {
      AxisExtent ax;
      PITT_TRY(axis_extent(ctx, c->d_xyz, c->n, r.coeffs, &ax));
      height = ax.height;
      if (p.model == PITT_MODEL_CYLINDER) {
        if (ax.idx1 >= 0) {
          for (int a = 0; a < 3; ++a) out->centroid[a] = (ax.p1[a] + ax.p2[a]) / 2;
          out->centroid_valid = 1;
        }
      } else {
        for (int a = 0; a < 3; ++a) out->centroid[a] = r.coeffs[a] + 3.0f / 4.0f * height * ax.dirn[a];
        out->centroid_valid = 1;
      }
    }
    out->coefficients[r.n_coeffs] = height;  // coefficientVector.push_back(height)
    out->n_coefficients = r.n_coeffs + 1;
  }
  return PITT_OK;
}

static int select_primitive_rule(int64_t planeInl, int64_t sphereInl, int64_t cylinderInl, int64_t coneInl, float prio) {
  if ((!planeInl) && (!sphereInl) && (!cylinderInl) && (!coneInl)) return PITT_TAG_UNKNOWN;
  if ((coneInl >= planeInl) && (coneInl >= sphereInl) && ((float)(uint64_t)coneInl >= (float)(uint64_t)cylinderInl * prio))
    return PITT_TAG_CONE;
  if ((cylinderInl >= planeInl) && (cylinderInl >= coneInl) && (cylinderInl >= sphereInl)) return PITT_TAG_CYLINDER;
  if ((planeInl >= coneInl) && (planeInl >= sphereInl) && (planeInl >= cylinderInl)) return PITT_TAG_PLANE;
  if ((sphereInl >= planeInl) && (sphereInl >= coneInl) && (sphereInl >= cylinderInl)) return PITT_TAG_SPHERE;
  return PITT_TAG_UNKNOWN;
}

// number of inliers after inlierToVectorMsg: index VALUE 0 is dropped (pc_manager.cpp:108). The
// ascending list starts with 0 iff point 0 is an inlier.
static int count_after_zero_drop(pitt_ctx* ctx, const SacDeviceResult& r, int* out) {
  *out = r.n_inliers;
  if (r.n_inliers > 0) {
    int first = -1;
    PITT_CUDA(ctx, cudaMemcpyAsync(&first, r.d_inliers, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
    if (first == 0) *out = r.n_inliers - 1;
  }
  return PITT_OK;
}


}  // namespace pitt

// ------------------------------------------------------------------ frame workers
// A frame has (clusters x 4) independent RANSAC fits whose device work is latency bound (one-CTA
// Levenberg-Marquardt, small scoring grids). Each worker = one helper pitt_ctx (own stream, arena and
// pinned scratch) + one host thread, so the fits overlap on the device and their host-side stop-rule
// scans overlap on the host.
struct pitt_workers {
  std::vector<pitt_ctx*> ctxs;
  std::vector<std::thread> threads;
  std::mutex m;
  std::condition_variable cv_work, cv_done;
  std::deque<std::function<void(pitt_ctx*)>> q;
  int pending = 0;
  bool stop = false;
};

namespace pitt {

static void worker_main(pitt_workers* W, int w) {
  pitt_ctx* ctx = W->ctxs[w];
  cudaSetDevice(ctx->device);
  for (;;) {
    std::function<void(pitt_ctx*)> fn;
    {
      std::unique_lock<std::mutex> lk(W->m);
      W->cv_work.wait(lk, [&] { return W->stop || !W->q.empty(); });
      if (W->stop && W->q.empty()) return;
      fn = std::move(W->q.front());
      W->q.pop_front();
    }
    fn(ctx);
    {
      std::lock_guard<std::mutex> lk(W->m);
      if (--W->pending == 0) W->cv_done.notify_all();
    }
  }
}
static pitt_workers* workers_get(pitt_ctx* ctx) {
  if (ctx->workers || ctx->n_workers <= 0) return ctx->workers;
  pitt_workers* W = new pitt_workers();
  for (int w = 0; w < ctx->n_workers; ++w) {
    pitt_ctx* c = pitt_create(ctx->device, ctx->seed);
    if (!c) break;
    c->n_workers = 0;  // helpers never fan out themselves
    W->ctxs.push_back(c);
  }
  if (W->ctxs.empty()) {
    delete W;
    return nullptr;
  }
  for (size_t w = 0; w < W->ctxs.size(); ++w) W->threads.emplace_back(worker_main, W, (int)w);
  ctx->workers = W;
  return W;
}
void workers_destroy(pitt_ctx* ctx) {
  pitt_workers* W = ctx->workers;
  if (!W) return;
  {
    std::lock_guard<std::mutex> lk(W->m);
    W->stop = true;
  }
  W->cv_work.notify_all();
  for (auto& t : W->threads) t.join();
  for (pitt_ctx* c : W->ctxs) pitt_destroy(c);
  delete W;
  ctx->workers = nullptr;
}
// runs the tasks on the helpers and returns when all are done
static void workers_run(pitt_workers* W, std::vector<std::function<void(pitt_ctx*)>>& tasks) {
  {
    std::lock_guard<std::mutex> lk(W->m);
    for (auto& t : tasks) W->q.push_back(std::move(t));
    W->pending += (int)tasks.size();
  }
  W->cv_work.notify_all();
  std::unique_lock<std::mutex> lk(W->m);
  W->cv_done.wait(lk, [&] { return W->pending == 0; });
}

}  // namespace pitt

namespace pitt {

int g_frame_legacy = 0;  // test hook (pitt_debug_frame_mode): 1 = the round-1 frame path (one synchronous seg.segment() per fit)
int g_frame_no_batch = 0;  // test hook (pitt_debug_frame_mode 2): asynchronous chains per fit instead of per model batch

constexpr int FIT_WORDS = 48;  // result block of one fit: 16 ints (sac_segment_async) + model[8] + refined[8] + axis extent[16]
constexpr int FIT_STREAMS = 4;

static int stage_reserve(pitt_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->h_stage_bytes) return PITT_OK;
  PITT_CUDA(ctx, cudaDeviceSynchronize());  // helper streams may still read the old block
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  ctx->h_stage = nullptr;
  ctx->h_stage_bytes = 0;
  const size_t want = bytes + bytes / 2 + 4096;
  PITT_CUDA(ctx, cudaMallocHost(&ctx->h_stage, want));
  ctx->h_stage_bytes = want;
  return PITT_OK;
}

// one fit's result block (see sac_segment_async) -> the service response fields the frame needs; a block whose scan raised a flag
// is settled by the synchronous path
static int decode_fit(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_sac_params& p, bool issued, const int* I, PrimitiveHost* Pp,
                      int* inl_out) {
  PrimitiveHost& P = *Pp;
  const float* F = reinterpret_cast<const float*>(I + 16);
  const float* A = F + 16;
    if (issued && I[8] != 0) {
      // the device-side scan could not decide exactly like the host would: the synchronous path settles it
      PITT_TRY(primitive_service_impl(ctx, cloud, p, &P));
      PITT_TRY(count_after_zero_drop(ctx, P.sac, inl_out));
      return PITT_OK;
    }
    memset(&P.sac.info, 0, sizeof(P.sac.info));
    P.sac.info.best_hypothesis = -1;
    P.sac.d_inliers = nullptr;
    P.sac.n_inliers = 0;
    P.sac.n_coeffs = 0;
    for (int i = 0; i < 8; ++i) P.sac.coeffs[i] = 0.0f;
    const bool found = issued && I[0] >= 0;
    const int NC = (p.model == PITT_MODEL_PLANE || p.model == PITT_MODEL_SPHERE) ? 4 : 7;
    if (issued) {
      P.sac.info.iterations = I[6];
      P.sac.info.skipped = I[7];
    }
    if (found) {
      P.sac.info.best_hypothesis = I[0];
      P.sac.info.best_count = I[1];
      P.sac.info.n_inliers_model = I[2];
      P.sac.info.lm_info = I[4];
      P.sac.info.lm_nfev = I[5];
      for (int i = 0; i < NC; ++i) { P.sac.info.model_coeffs[i] = F[i]; P.sac.coeffs[i] = F[8 + i]; }
      P.sac.n_coeffs = NC;
      P.sac.n_inliers = I[3];
    }
    const SacDeviceResult& r = P.sac;
    P.n_coefficients = r.n_coeffs;
    for (int i = 0; i < 8; ++i) P.coefficients[i] = i < r.n_coeffs ? r.coeffs[i] : 0.0f;
    P.centroid[0] = P.centroid[1] = P.centroid[2] = 0.0f;
    P.centroid_valid = 0;
    if (p.model == PITT_MODEL_SPHERE) {
      if (r.n_coeffs > 0) {
        for (int a = 0; a < 3; ++a) P.centroid[a] = r.coeffs[a];
        P.centroid_valid = 1;
      }
    } else if (p.model == PITT_MODEL_CYLINDER || p.model == PITT_MODEL_CONE) {
      float height = -1.0f;
      if (r.n_inliers > 0) {
        height = A[0];
        int idx1 = -1;
        memcpy(&idx1, &A[1], 4);
        if (p.model == PITT_MODEL_CYLINDER) {
          if (idx1 >= 0) {
            for (int a = 0; a < 3; ++a) P.centroid[a] = (A[3 + a] + A[6 + a]) / 2;
            P.centroid_valid = 1;
          }
        } else {
          for (int a = 0; a < 3; ++a) P.centroid[a] = r.coeffs[a] + 3.0f / 4.0f * height * A[9 + a];
          P.centroid_valid = 1;
        }
      }
      P.coefficients[r.n_coeffs] = height;  // coefficientVector.push_back(height)
      P.n_coefficients = r.n_coeffs + 1;
    }
    // PCManager::inlierToVectorMsg drops the index VALUE 0 (pc_manager.cpp:108): the ascending list starts with it or not at all
    *inl_out = r.n_inliers - ((r.n_inliers > 0 && I[9] == 0) ? 1 : 0);
  return PITT_OK;
}

// All primitive fits of one support (clusters x {sphere, cylinder, cone, plane}; the reference calls the four services one after
// the other for every cluster, ransac_segmentation.cpp:235-262) enqueued WITHOUT a host round trip: every fit is one chain of
// launches (sample table upload, estimate, score, device-side PCL stop rule, select, refine, select, axis extent) on one of four
// helper streams, all result blocks come back in ONE copy behind ONE synchronisation. A fit whose device-side scan raised a flag
// (batch exhausted, a decision libm could round differently, a collinear speculative plane sample) is repeated on the
// synchronous path; results are identical either way.
static int frame_fits_async(pitt_ctx* ctx, std::vector<pitt_cloud>& cc, const pitt_sac_params* const sp[4], std::vector<PrimitiveHost>& ph,
                            std::vector<int>& inl) {
  const int nc = (int)cc.size(), nfits = nc * 4;
  if (nfits == 0) return PITT_OK;
  size_t words = (size_t)nfits * FIT_WORDS;
  std::vector<size_t> sample_off((size_t)nfits, 0);
  for (int c = 0; c < nc; ++c)
    for (int m = 0; m < 4; ++m) {
      sample_off[c * 4 + m] = words;
      const int S = sp[m]->model == PITT_MODEL_PLANE ? 3 : sp[m]->model == PITT_MODEL_SPHERE ? 4 : sp[m]->model == PITT_MODEL_CYLINDER ? 2 : 3;
      words += (size_t)(std::max(sp[m]->max_iterations, 0) + 1) * S + 4;
    }
  PITT_TRY(stage_reserve(ctx, words * 4));
  int* h_stage = reinterpret_cast<int*>(ctx->h_stage);
  int* d_res = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)nfits * FIT_WORDS, &d_res));
  PITT_CUDA(ctx, cudaMemsetAsync(d_res, 0, (size_t)nfits * FIT_WORDS * sizeof(int), ctx->stream));
  if (!ctx->fit_streams[0]) {
    for (int i = 0; i < FIT_STREAMS; ++i) {
      PITT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->fit_streams[i], cudaStreamNonBlocking));
      PITT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fit_join[i], cudaEventDisableTiming));
    }
    PITT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fit_fork, cudaEventDisableTiming));
  }
  PITT_CUDA(ctx, cudaEventRecord(ctx->ev_fit_fork, ctx->stream));
  for (int i = 0; i < FIT_STREAMS; ++i) PITT_CUDA(ctx, cudaStreamWaitEvent(ctx->fit_streams[i], ctx->ev_fit_fork, 0));
  std::vector<char> issued((size_t)nfits, 0);
  cudaStream_t main_stream = ctx->stream;
  int status = PITT_OK, k = 0;
  const int order[4] = {1, 2, 0, 3};  // longest chains first (cylinder, cone, sphere, plane)
  for (int oi = 0; oi < 4 && status == PITT_OK; ++oi)
    for (int c = 0; c < nc && status == PITT_OK; ++c, ++k) {
      const int m = order[oi], slot = c * 4 + m;
      ctx->stream = ctx->fit_streams[k % FIT_STREAMS];
      SacAsync sa;
      sa.d_ints = d_res + (size_t)slot * FIT_WORDS;
      sa.d_flt = reinterpret_cast<float*>(sa.d_ints + 16);
      bool did = false;
      status = sac_segment_async(ctx, &cc[c], *sp[m], h_stage + sample_off[slot], &sa, &did);
      issued[slot] = did ? 1 : 0;
      if (status == PITT_OK && did && (sp[m]->model == PITT_MODEL_CYLINDER || sp[m]->model == PITT_MODEL_CONE))
        status = axis_extent_async(ctx, cc[c].d_xyz, cc[c].n, sa.d_flt + 8, sa.d_ints, sa.d_flt + 16);
    }
  ctx->stream = main_stream;
  for (int i = 0; i < FIT_STREAMS; ++i) {
    cudaEventRecord(ctx->ev_fit_join[i], ctx->fit_streams[i]);
    cudaStreamWaitEvent(main_stream, ctx->ev_fit_join[i], 0);
  }
  if (status != PITT_OK) {
    pitt::stream_sync(ctx);
    return status;
  }
  PITT_CUDA(ctx, cudaMemcpyAsync(h_stage, d_res, (size_t)nfits * FIT_WORDS * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  for (int slot = 0; slot < nfits; ++slot)
    PITT_TRY(decode_fit(ctx, &cc[slot / 4], *sp[slot % 4], issued[slot] != 0, h_stage + (size_t)slot * FIT_WORDS, &ph[slot], &inl[slot]));
  return PITT_OK;
}

// The same fits with every launch serving ALL clusters of the support (problem = a grid dimension, pointers and sizes from a
// descriptor array on the device): four model chains of about ten launches each instead of about fifteen launches per fit.
// Requires the reference's own configuration (PCL sample stream, adaptive stop, refinement on) and clusters of at most 4096
// points; anything else takes frame_fits_async.
static bool fits_can_batch(const std::vector<pitt_cloud>& cc, const pitt_sac_params* const sp[4]) {
  for (const auto& c : cc)
    if (c.n > 4096 || c.n < 1) return false;
  for (int m = 0; m < 4; ++m) {
    const pitt_sac_params& p = *sp[m];
    if (p.sampler != PITT_SAMPLER_PCL_MT19937 || p.stop != PITT_STOP_PCL_ADAPTIVE || !p.optimize || p.max_iterations < 1 ||
        p.max_iterations > 8000)
      return false;
  }
  return sp[0]->model == PITT_MODEL_SPHERE && sp[1]->model == PITT_MODEL_CYLINDER && sp[2]->model == PITT_MODEL_CONE &&
         sp[3]->model == PITT_MODEL_PLANE;
}
static int frame_fits_batched(pitt_ctx* ctx, std::vector<pitt_cloud>& cc, const pitt_sac_params* const sp[4], std::vector<PrimitiveHost>& ph,
                              std::vector<int>& inl) {
  const int nc = (int)cc.size(), nfits = nc * 4;
  if (nfits == 0) return PITT_OK;
  const int Sm[4] = {4, 2, 3, 3};  // sample sizes of sphere, cylinder, cone, plane
  int n_max = 1;
  for (const auto& c : cc) n_max = std::max(n_max, c.n);
  // pinned staging: result blocks | descriptors (grouped by model) | sample tables
  const size_t res_words = (size_t)nfits * FIT_WORDS;
  const size_t desc_words = ((size_t)nfits * sizeof(FitDesc) + 3) / 4;
  size_t sample_words = 0;
  for (int m = 0; m < 4; ++m) sample_words += (size_t)nc * ((size_t)(sp[m]->max_iterations + 1) * Sm[m]);
  PITT_TRY(stage_reserve(ctx, (res_words + desc_words + sample_words + 64) * 4));
  int* h_stage = reinterpret_cast<int*>(ctx->h_stage);
  FitDesc* h_desc = reinterpret_cast<FitDesc*>(h_stage + ((res_words + 3) & ~(size_t)3));
  int* h_samples = reinterpret_cast<int*>(h_desc + nfits);
  int* d_res = nullptr;
  FitDesc* d_desc = nullptr;
  int* d_samples = nullptr;
  int* d_counts_all = nullptr;
  PITT_TRY(arena_alloc(ctx, res_words, &d_res));
  PITT_TRY(arena_alloc(ctx, (size_t)nfits, &d_desc));
  PITT_TRY(arena_alloc(ctx, sample_words + 4, &d_samples));
  size_t count_words = 0;
  for (int m = 0; m < 4; ++m) count_words += (size_t)nc * (sp[m]->max_iterations + 1);
  PITT_TRY(arena_alloc(ctx, count_words + 4, &d_counts_all));
  const int nt = cdiv(n_max, AP_TPB);
  std::vector<char> issued((size_t)nfits, 0);
  size_t soff = 0, coff = 0;
  for (int m = 0; m < 4; ++m)
    for (int c = 0; c < nc; ++c) {
      const int slot = c * 4 + m;           // result / response order (cluster major)
      FitDesc& d = h_desc[m * nc + c];      // launch order (model major)
      memset(&d, 0, sizeof(d));
      const int n = cc[c].n, S = Sm[m], Hcap = sp[m]->max_iterations + 1;
      d.xyz = cc[c].d_xyz;
      d.nrm = cc[c].d_nrm;
      d.n = n;
      d.ints = d_res + (size_t)slot * FIT_WORDS;
      d.flt = reinterpret_cast<float*>(d.ints + 16);
      d.samples = d_samples + soff;
      d.counts = d_counts_all + coff;
      int have = 0;
      if (n >= S) {
        PclSampleStream stream(n, sp[m]->model, nullptr);  // plane: speculative draws, the scan flags a collinear triple
        for (int h = 0; h < Hcap; ++h) {
          if (!stream.next(h_samples + soff + (size_t)h * S)) break;
          ++have;
        }
      }
      d.H = have;
      issued[slot] = have > 0 ? 1 : 0;
      soff += (size_t)Hcap * S;
      coff += (size_t)Hcap;
      PITT_TRY(arena_alloc(ctx, (size_t)Hcap, &d.recs));
      PITT_TRY(arena_alloc(ctx, (size_t)Hcap * 8, &d.coeffs8));
      PITT_TRY(arena_alloc(ctx, (size_t)Hcap, &d.flags));
      PITT_TRY(arena_alloc(ctx, (size_t)n, &d.inl));
      if (sp[m]->model == PITT_MODEL_PLANE) PITT_TRY(arena_alloc(ctx, (size_t)REF_BLOCKS * 10, &d.partial));
      if (sp[m]->model == PITT_MODEL_CYLINDER || sp[m]->model == PITT_MODEL_CONE) {
        PairBest* bb = nullptr;
        PITT_TRY(arena_alloc(ctx, (size_t)n, &d.proj));
        PITT_TRY(arena_alloc(ctx, (size_t)nt * nt, &bb));
        d.bb = bb;
      }
    }
  PITT_CUDA(ctx, cudaMemsetAsync(d_res, 0, res_words * sizeof(int), ctx->stream));
  PITT_CUDA(ctx, cudaMemsetAsync(d_counts_all, 0, count_words * sizeof(int), ctx->stream));
  PITT_CUDA(ctx, cudaMemcpyAsync(d_desc, h_desc, (size_t)nfits * sizeof(FitDesc), cudaMemcpyHostToDevice, ctx->stream));
  PITT_CUDA(ctx, cudaMemcpyAsync(d_samples, h_samples, sample_words * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  if (!ctx->fit_streams[0]) {
    for (int i = 0; i < FIT_STREAMS; ++i) {
      PITT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->fit_streams[i], cudaStreamNonBlocking));
      PITT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fit_join[i], cudaEventDisableTiming));
    }
    PITT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fit_fork, cudaEventDisableTiming));
  }
  PITT_CUDA(ctx, cudaEventRecord(ctx->ev_fit_fork, ctx->stream));
  for (int i = 0; i < FIT_STREAMS; ++i) PITT_CUDA(ctx, cudaStreamWaitEvent(ctx->fit_streams[i], ctx->ev_fit_fork, 0));
  cudaStream_t main_stream = ctx->stream;
  int status = PITT_OK;
  const int order[4] = {1, 2, 0, 3};  // longest chains first (cylinder, cone, sphere, plane)
  for (int oi = 0; oi < 4 && status == PITT_OK; ++oi) {
    const int m = order[oi];
    ctx->stream = ctx->fit_streams[oi % FIT_STREAMS];
    status = sac_fit_batch_async(ctx, *sp[m], h_desc + m * nc, d_desc + m * nc, nc, sp[m]->model == PITT_MODEL_PLANE);
    if (status == PITT_OK && (sp[m]->model == PITT_MODEL_CYLINDER || sp[m]->model == PITT_MODEL_CONE)) {
      const FitDesc* dd = d_desc + m * nc;
      axis_project_dev_kernel<<<dim3(cdiv(n_max, 256), nc), 256, 0, ctx->stream>>>(nullptr, 0, nullptr, nullptr, nullptr, dd);
      axis_pairs_dev_kernel<<<dim3(nt, nt, nc), AP_TPB, 0, ctx->stream>>>(nullptr, 0, nullptr, nullptr, dd);
      axis_pairs_final_dev_kernel<<<nc, 1024, 0, ctx->stream>>>(nullptr, nt * nt, nullptr, nullptr, nullptr, nullptr, dd);
      ctx->launches += 3;
      if (cudaGetLastError() != cudaSuccess) status = fail(ctx, PITT_ERR_CUDA, "axis extent kernels (batch)");
    }
  }
  ctx->stream = main_stream;
  for (int i = 0; i < FIT_STREAMS; ++i) {
    cudaEventRecord(ctx->ev_fit_join[i], ctx->fit_streams[i]);
    cudaStreamWaitEvent(main_stream, ctx->ev_fit_join[i], 0);
  }
  if (status != PITT_OK) {
    pitt::stream_sync(ctx);
    return status;
  }
  PITT_CUDA(ctx, cudaMemcpyAsync(h_stage, d_res, res_words * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  for (int slot = 0; slot < nfits; ++slot)
    PITT_TRY(decode_fit(ctx, &cc[slot / 4], *sp[slot % 4], issued[slot] != 0, h_stage + (size_t)slot * FIT_WORDS, &ph[slot], &inl[slot]));
  return PITT_OK;
}

}  // namespace pitt

using namespace pitt;

extern "C" {

int pitt_find_supports(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_support_params* params, pitt_support_result* res) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!cloud || !params || !res) return fail(ctx, PITT_ERR_INVALID, "pitt_find_supports arguments");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  std::vector<SupportDev> sup;
  SupportCfg c;
  int trips = 0;
  PITT_TRY(find_supports_impl(ctx, cloud, *params, &c, &sup, &trips));
  res->n_supports = 0;
  res->loop_trips = trips;
  res->maps_used = 0;
  res->points_used = 0;
  res->used_min_iterative_cloud_percentual_size = c.minCloudPct;
  res->used_min_iterative_plane_percentual_size = c.minPlanePct;
  res->used_max_variance_threshold_for_horizontal = c.maxVar;
  res->used_min_variance_threshold_for_horizontal = c.minVar;
  res->used_ransac_max_iteration_threshold = c.maxIter;
  res->used_ransac_distance_point_in_shape_threshold = c.thr;
  res->used_ransac_model_normal_distance_weigth = c.w;
  for (int i = 0; i < 3; ++i) {
    res->used_horizontal_axis[i] = c.axis[i];
    res->used_support_edge_remove_offset[i] = c.offset[i];
  }
  const int n0 = cloud->n;
  int status = PITT_OK;
  for (size_t s = 0; s < sup.size(); ++s) {
    const SupportDev& S = sup[s];
    const int64_t need_pts = (int64_t)S.n_support + S.n_on;
    if ((int)s < res->supports_cap && res->supports && res->maps && res->points && res->maps_used + n0 <= res->maps_cap &&
        res->points_used + need_pts <= res->points_cap) {
      pitt_support& O = res->supports[s];
      O.n_map = n0;
      O.n_support = S.n_support;
      O.n_on_support = S.n_on;
      O.a = S.co[0]; O.b = S.co[1]; O.c = S.co[2]; O.d = S.co[3];
      O.map_offset = res->maps_used;
      O.support_offset = res->points_used;
      O.on_support_offset = res->points_used + S.n_support;
      PITT_CUDA(ctx, cudaMemcpyAsync(res->maps + res->maps_used, S.d_map, (size_t)n0 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      PITT_CUDA(ctx, cudaMemcpyAsync(res->points + 4 * res->points_used, S.d_support, (size_t)S.n_support * 16, cudaMemcpyDeviceToHost, ctx->stream));
      if (S.n_on > 0)
        PITT_CUDA(ctx, cudaMemcpyAsync(res->points + 4 * (res->points_used + S.n_support), S.d_on, (size_t)S.n_on * 16, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      status = PITT_ERR_CAPACITY;
    }
    res->maps_used += n0;
    res->points_used += need_pts;
    res->n_supports++;
  }
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  timer.finish();
  if (status == PITT_ERR_CAPACITY) return fail(ctx, status, "support result buffers too small (see maps_used / points_used)");
  return status;
}

int pitt_cluster_service(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_cluster_params* params, pitt_clusters_result* res) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!cloud || !params || !res) return fail(ctx, PITT_ERR_INVALID, "pitt_cluster_service arguments");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  ClustersDev cd;
  PITT_TRY(cluster_service_impl(ctx, cloud->d_xyz, cloud->n, *params, &cd));
  res->n_clusters = 0;
  res->indices_used = 0;
  int status = PITT_OK;
  for (size_t c = 0; c < cd.sizes.size(); ++c) {
    const int sz = cd.sizes[c];
    if ((int)c < res->clusters_cap && res->clusters && res->indices && res->indices_used + sz <= res->indices_cap) {
      pitt_cluster& C = res->clusters[c];
      C.n = sz;
      C.offset = res->indices_used;
      C.x_centroid = cd.centroid[3 * c]; C.y_centroid = cd.centroid[3 * c + 1]; C.z_centroid = cd.centroid[3 * c + 2];
      memcpy(res->indices + res->indices_used, cd.indices.data() + cd.offsets[c], (size_t)sz * sizeof(int));
    } else {
      status = PITT_ERR_CAPACITY;
    }
    res->indices_used += sz;
    res->n_clusters++;
  }
  timer.finish();
  if (status == PITT_ERR_CAPACITY) return fail(ctx, status, "cluster result buffers too small");
  return status;
}

int pitt_primitive_service(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_sac_params* params, pitt_primitive_result* res) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!cloud || !params || !res) return fail(ctx, PITT_ERR_INVALID, "pitt_primitive_service arguments");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  PrimitiveHost ph;
  PITT_TRY(primitive_service_impl(ctx, cloud, *params, &ph));
  res->n_coefficients = ph.n_coefficients;
  for (int i = 0; i < 8; ++i) res->coefficients[i] = ph.coefficients[i];
  res->x_centroid = ph.centroid[0]; res->y_centroid = ph.centroid[1]; res->z_centroid = ph.centroid[2];
  res->centroid_valid = ph.centroid_valid;
  res->info = ph.sac.info;
  // PCManager::inlierToVectorMsg: drop index value 0
  const int n_inl = ph.sac.n_inliers;
  int status = PITT_OK;
  res->n_inliers = 0;
  if (n_inl > 0) {
    std::vector<int> h(n_inl);
    PITT_CUDA(ctx, cudaMemcpyAsync(h.data(), ph.sac.d_inliers, (size_t)n_inl * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
    int m = 0;
    for (int i = 0; i < n_inl; ++i) {
      if (h[i] == 0) continue;
      if (res->inliers && m < res->inliers_cap) res->inliers[m] = h[i];
      else if (res->inliers) status = PITT_ERR_CAPACITY;
      ++m;
    }
    res->n_inliers = m;
  }
  timer.finish();
  res->info.device_ms = ctx->last_ms;
  if (status == PITT_ERR_CAPACITY) return fail(ctx, status, "primitive inlier buffer too small");
  return status;
}

int pitt_select_primitive(int64_t plane_inl, int64_t sphere_inl, int64_t cylinder_inl, int64_t cone_inl, float prio) {
  return select_primitive_rule(plane_inl, sphere_inl, cylinder_inl, cone_inl, prio);
}

// PITT_DEBUG_HASH=1 (development aid): FNV-1a hashes of the intermediate stages of a frame are stored, as bit patterns, in the
// unused tail of support_coefficients (entries 16..): [16] cloud, [17] on-support cloud, [18] cluster clouds, [19] cluster normals
static bool dbg_hash_on() {
  static int e = -1;
  if (e < 0) { const char* v = getenv("PITT_DEBUG_HASH"); e = (v && v[0] == '1') ? 1 : 0; }
  return e == 1;
}
static unsigned dbg_hash_device(pitt_ctx* ctx, const void* d, size_t bytes) {
  std::vector<unsigned char> h(bytes);
  if (bytes) {
    cudaMemcpyAsync(h.data(), d, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    pitt::stream_sync(ctx);
  }
  unsigned v = 2166136261u;
  for (unsigned char b : h) { v ^= b; v *= 16777619u; }
  return v;
}

int pitt_segment_frame(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_frame_params* fp, pitt_frame_result* res) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!cloud || !fp || !res) return fail(ctx, PITT_ERR_INVALID, "pitt_segment_frame arguments");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  res->n_supports = res->n_clusters = res->n_shapes = 0;
  memset(res->support_coefficients, 0, sizeof(res->support_coefficients));
  memset(res->support_sizes, 0, sizeof(res->support_sizes));
  memset(res->on_support_sizes, 0, sizeof(res->on_support_sizes));
  res->device_ms = 0.0;
  const int n = cloud->n;
  int status = PITT_OK;
  if (n > fp->min_points) {  // obj_segmentation.cpp:251
    // obj_segmentation.cpp:253 estimates the frame normals and ships them with the request; the
    // supports service never uses them (plane RANSAC ignores normals). They are still
    // computed here because they are part of the reference's request (and of its cost).
    float4* d_nrm = nullptr;
    PITT_TRY(arena_alloc(ctx, (size_t)n, &d_nrm));
    // Nothing downstream reads the frame normals, so they run on a second stream of the context BESIDE the supports loop (0.6 of
    // the 3.3 ms a single full-resolution frame takes) and are joined before the clustering stage, which rebuilds the context's
    // grid tables. A context whose stream was supplied by the caller keeps everything on that one stream.
    bool forked = false;
    if (ctx->own_stream && !TraceScope::enabled()) {
      if (!ctx->aux_stream) {
        PITT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
        PITT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_aux_fork, cudaEventDisableTiming));
        PITT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_aux_join, cudaEventDisableTiming));
      }
      PITT_CUDA(ctx, cudaEventRecord(ctx->ev_aux_fork, ctx->stream));
      PITT_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_aux_fork, 0));
      cudaStream_t main_stream = ctx->stream;
      ctx->stream = ctx->aux_stream;
      const int st_n = estimate_normals_impl(ctx, cloud->d_xyz, n, fp->normals_k, fp->viewpoint, d_nrm);
      ctx->stream = main_stream;
      if (st_n != PITT_OK) { cudaStreamSynchronize(ctx->aux_stream); return st_n; }
      PITT_CUDA(ctx, cudaEventRecord(ctx->ev_aux_join, ctx->aux_stream));
      forked = true;
    } else {
      TraceScope ts(ctx, "frame: normals");
      PITT_TRY(estimate_normals_impl(ctx, cloud->d_xyz, n, fp->normals_k, fp->viewpoint, d_nrm));
    }
    std::vector<SupportDev> sup;
    int trips = 0;
    const int st_sup = find_supports_impl(ctx, cloud, fp->support, nullptr, &sup, &trips);
    if (forked) {
      // join: whatever follows on the main stream (grid rebuilds, the final synchronisation, the next call's arena) is ordered
      // after the normals
      if (cudaStreamWaitEvent(ctx->stream, ctx->ev_aux_join, 0) != cudaSuccess) { cudaStreamSynchronize(ctx->aux_stream); return fail(ctx, PITT_ERR_CUDA, "join of the normals stream"); }
    }
    if (st_sup != PITT_OK) { pitt::stream_sync(ctx); return st_sup; }
    res->n_supports = (int)sup.size();
    for (size_t s = 0; s < sup.size(); ++s) {
      const SupportDev& S = sup[s];
      if (s < 8) {
        for (int i = 0; i < 4; ++i) res->support_coefficients[4 * s + i] = S.co[i];
        res->support_sizes[s] = S.n_support;
        res->on_support_sizes[s] = S.n_on;
      }
      ClustersDev cd;
      if (dbg_hash_on() && s == 0) {
        unsigned hv = dbg_hash_device(ctx, cloud->d_xyz, (size_t)n * 16);
        memcpy(&res->support_coefficients[16], &hv, 4);
        hv = dbg_hash_device(ctx, S.d_on, (size_t)S.n_on * 16);
        memcpy(&res->support_coefficients[17], &hv, 4);
      }
      PITT_TRY(cluster_service_impl(ctx, S.d_on, S.n_on, fp->cluster, &cd, false));
      const int nc = (int)cd.sizes.size();
      if (nc == 0) continue;
      // cluster normals (ransac_segmentation.cpp:233) on this ctx's stream, then the 4 fits per cluster
      std::vector<pitt_cloud> cc(nc);  // non-owning views of the cluster clouds
      const int total = cd.offsets[nc];
      float4* d_cn_all = nullptr;
      PITT_TRY(arena_alloc(ctx, (size_t)std::max(total, 1), &d_cn_all));
      bool all_small = cd.d_head != nullptr;
      for (int c = 0; c < nc; ++c) all_small = all_small && cd.sizes[c] <= KNN_BRUTE_MAX;
      // the clusters' normals in ONE launch (every query searches its own cluster), unless a cluster is large enough for the grid
      if (all_small)
        PITT_TRY(estimate_normals_segmented(ctx, cd.d_points, total, cd.d_head + 2 + CC_MAXC, cd.d_head, fp->normals_k, fp->viewpoint, d_cn_all));
      for (int c = 0; c < nc; ++c) {
        cc[c].n = cd.sizes[c];
        cc[c].d_xyz = const_cast<float4*>(cd.d_points) + cd.offsets[c];
        float4* d_cn = d_cn_all + cd.offsets[c];
        if (!all_small) PITT_TRY(estimate_normals_impl(ctx, cc[c].d_xyz, cc[c].n, fp->normals_k, fp->viewpoint, d_cn));
        cc[c].d_nrm = d_cn;
        cc[c].has_normals = true;
      }
      if (dbg_hash_on() && s == 0) {
        unsigned hv = dbg_hash_device(ctx, cd.d_points, (size_t)total * 16);
        memcpy(&res->support_coefficients[18], &hv, 4);
        hv = dbg_hash_device(ctx, d_cn_all, (size_t)total * 16);
        memcpy(&res->support_coefficients[19], &hv, 4);
        if (cd.d_labels) {
          hv = dbg_hash_device(ctx, cd.d_labels, (size_t)S.n_on * 4);
          memcpy(&res->support_coefficients[20], &hv, 4);
          hv = dbg_hash_device(ctx, cd.d_idx, (size_t)total * 4);
          memcpy(&res->support_coefficients[21], &hv, 4);
          hv = dbg_hash_device(ctx, cd.d_head, (size_t)CC_HEAD_INTS * 4);
          memcpy(&res->support_coefficients[22], &hv, 4);
        }
      }
      const pitt_sac_params* sp[4] = {&fp->sphere, &fp->cylinder, &fp->cone, &fp->plane};
      std::vector<PrimitiveHost> ph((size_t)nc * 4);
      std::vector<int> inl((size_t)nc * 4, 0), st((size_t)nc * 4, PITT_OK);
      pitt_workers* W = g_frame_legacy ? workers_get(ctx) : nullptr;
      if (!g_frame_legacy) {
        const int s1 = (fits_can_batch(cc, sp) && !g_frame_no_batch) ? frame_fits_batched(ctx, cc, sp, ph, inl)
                                                                   : frame_fits_async(ctx, cc, sp, ph, inl);
        if (s1 != PITT_OK) {
          for (auto& v : cc) { v.d_xyz = nullptr; v.d_nrm = nullptr; }
          return s1;
        }
      } else if (W) {
        if (!ctx->ev_fan) PITT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fan, cudaEventDisableTiming));
        PITT_CUDA(ctx, cudaEventRecord(ctx->ev_fan, ctx->stream));
        std::vector<std::function<void(pitt_ctx*)>> tasks;
        // longest fits first (cylinder, cone, sphere, plane) so the tail of the schedule is short
        const int order[4] = {1, 2, 0, 3};
        for (int oi = 0; oi < 4; ++oi)
          for (int c = 0; c < nc; ++c) {
            const int m = order[oi];
            const int slot = c * 4 + m;
            tasks.push_back([&, slot, c, m](pitt_ctx* h) {
              arena_reset(h);
              if (cudaStreamWaitEvent(h->stream, ctx->ev_fan, 0) != cudaSuccess) { st[slot] = PITT_ERR_CUDA; return; }
              pitt_cloud view;  // private view: the lazy host mirror of a cloud is per task
              view.n = cc[c].n; view.d_xyz = cc[c].d_xyz; view.d_nrm = cc[c].d_nrm; view.has_normals = true;
              int s1 = primitive_service_impl(h, &view, *sp[m], &ph[slot]);
              if (s1 == PITT_OK) s1 = count_after_zero_drop(h, ph[slot].sac, &inl[slot]);
              if (s1 == PITT_OK && pitt::stream_sync(h) != cudaSuccess) s1 = PITT_ERR_CUDA;
              ph[slot].sac.d_inliers = nullptr;  // helper arena memory: not valid after the task
              st[slot] = s1;
              view.d_xyz = nullptr; view.d_nrm = nullptr;
            });
          }
        workers_run(W, tasks);
        for (pitt_ctx* h : W->ctxs) { ctx->launches += h->launches; h->launches = 0; }
        for (int i = 0; i < nc * 4; ++i)
          if (st[i] != PITT_OK) {
            for (pitt_ctx* h : W->ctxs)
              if (!h->err.empty()) ctx->err = h->err;
            for (auto& v : cc) { v.d_xyz = nullptr; v.d_nrm = nullptr; }
            return st[i];
          }
      } else {
        for (int c = 0; c < nc; ++c)
          for (int m = 0; m < 4; ++m) {
            PITT_TRY(primitive_service_impl(ctx, &cc[c], *sp[m], &ph[c * 4 + m]));
            PITT_TRY(count_after_zero_drop(ctx, ph[c * 4 + m].sac, &inl[c * 4 + m]));
          }
      }
      for (int c = 0; c < nc; ++c) {
        const int64_t sphereInl = inl[c * 4 + 0], cylinderInl = inl[c * 4 + 1], coneInl = inl[c * 4 + 2], planeInl = inl[c * 4 + 3];
        const int tag = select_primitive_rule(planeInl, sphereInl, cylinderInl, coneInl, fp->cone_over_cylinder_priority);
        if (res->shapes && res->n_shapes < res->shapes_cap) {
          pitt_tracked_shape& T = res->shapes[res->n_shapes];
          memset(&T, 0, sizeof(T));
          T.object_id = res->n_clusters;
          T.shape_tag = tag;
          T.x_pc_centroid = cd.centroid[3 * c]; T.y_pc_centroid = cd.centroid[3 * c + 1]; T.z_pc_centroid = cd.centroid[3 * c + 2];
          const PrimitiveHost* sel = tag == PITT_TAG_CONE ? &ph[c * 4 + 2] : tag == PITT_TAG_CYLINDER ? &ph[c * 4 + 1]
                                     : tag == PITT_TAG_PLANE ? &ph[c * 4 + 3] : tag == PITT_TAG_SPHERE ? &ph[c * 4 + 0] : nullptr;
          if (sel) {
            T.x_est_centroid = sel->centroid[0]; T.y_est_centroid = sel->centroid[1]; T.z_est_centroid = sel->centroid[2];
            T.n_coefficients = sel->n_coefficients;
            for (int i = 0; i < 8; ++i) T.coefficients[i] = sel->coefficients[i];
          }
          T.n_points = cc[c].n;
          T.inl_plane = (int)planeInl; T.inl_sphere = (int)sphereInl; T.inl_cylinder = (int)cylinderInl; T.inl_cone = (int)coneInl;
        } else {
          status = PITT_ERR_CAPACITY;
        }
        res->n_shapes++;
        res->n_clusters++;
      }
      for (auto& v : cc) { v.d_xyz = nullptr; v.d_nrm = nullptr; }  // views own nothing
    }
  }
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  timer.finish();
  res->device_ms = ctx->last_ms;
  if (status == PITT_ERR_CAPACITY) return fail(ctx, status, "shapes buffer too small");
  return status;
}

int pitt_segment_raw_frames_batched(pitt_ctx* const* ctxs, int n_ctx, const void* const* frames, const int* n_points,
                                    int stride_bytes, int n_frames, const pitt_prefilter_params* prefilter,
                                    const pitt_frame_params* params, pitt_frame_result* results) {
  if (!ctxs || n_ctx <= 0) return PITT_ERR_CUDA;
  for (int t = 0; t < n_ctx; ++t)
    if (!ctxs[t]) return PITT_ERR_CUDA;
  if (n_frames < 0 || (n_frames > 0 && (!frames || !n_points || !results)) || !params)
    return fail(ctxs[0], PITT_ERR_INVALID, "pitt_segment_frames_batched arguments");
  std::atomic<int> first_error(PITT_OK);
  auto note = [&](int st) {
    if (st != PITT_OK) {
      int expected = PITT_OK;
      first_error.compare_exchange_strong(expected, st);
    }
  };
  // World-frame clouds in the HBM layout (point_step 16, no pre-filter): a context's frames are double-buffered. The copy of
  // its next frame is queued on the context's copy stream before the current frame is segmented, so the 4.9 MB transfer and
  // the wait for it (pitt_stage_cloud synchronises: the caller's buffer is only borrowed for that call; here it is borrowed for
  // the whole batched call) disappear behind the previous frame's kernels.
  auto worker_prefetch = [&](int t) {
    pitt_ctx* ctx = ctxs[t];
    cudaSetDevice(ctx->device);
    if (!ctx->h2d_stream) {
      if (cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&ctx->ev_h2d[0], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&ctx->ev_h2d[1], cudaEventDisableTiming) != cudaSuccess) {
        note(fail(ctx, PITT_ERR_CUDA, "copy stream of the frame stream"));
        return;
      }
    }
    auto issue = [&](int i, int slot, pitt_cloud** out) -> int {
      pitt_cloud* c = new pitt_cloud();
      c->n = n_points[i];
      if (c->n < 0 || (c->n > 0 && !frames[i])) { delete c; return fail(ctx, PITT_ERR_INVALID, "pitt_segment_frames_batched: frame pointer / size"); }
      if (c->n > 0) {
        if (pool_alloc(ctx, (size_t)c->n * sizeof(float4), (void**)&c->d_xyz) != PITT_OK) { delete c; return PITT_ERR_CUDA; }
        cudaError_t e = cudaMemcpyAsync(c->d_xyz, frames[i], (size_t)c->n * 16, cudaMemcpyHostToDevice, ctx->h2d_stream);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_h2d[slot], ctx->h2d_stream);
        if (e != cudaSuccess) { pool_free(ctx, c->d_xyz, (size_t)c->n * sizeof(float4)); delete c; return fail(ctx, PITT_ERR_CUDA, "cudaMemcpyAsync(H2D frame)", e); }
      }
      *out = c;
      return PITT_OK;
    };
    pitt_cloud* cur = nullptr;
    int slot = 0;
    int st = t < n_frames ? issue(t, slot, &cur) : PITT_OK;
    note(st);
    for (int i = t; i < n_frames && st == PITT_OK; i += n_ctx) {
      pitt_cloud* next = nullptr;
      int st_next = PITT_OK;
      if (i + n_ctx < n_frames) st_next = issue(i + n_ctx, slot ^ 1, &next);
      if (cur->n > 0 && cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[slot], 0) != cudaSuccess) st = fail(ctx, PITT_ERR_CUDA, "wait for the frame copy");
      if (st == PITT_OK) st = pitt_segment_frame(ctx, cur, params, &results[i]);
      if (st != PITT_OK) cudaStreamSynchronize(ctx->h2d_stream);
      pitt_release_cloud(ctx, cur);
      cur = next;
      slot ^= 1;
      note(st);
      if (st == PITT_OK) { st = st_next; note(st); }
    }
    if (cur) {
      cudaStreamSynchronize(ctx->h2d_stream);
      pitt_release_cloud(ctx, cur);
    }
  };
  const bool prefetch = !prefilter && stride_bytes == 16;
  auto worker = [&](int t) {
    if (prefetch) { worker_prefetch(t); return; }
    pitt_ctx* ctx = ctxs[t];
    for (int i = t; i < n_frames; i += n_ctx) {
      pitt_cloud* c = nullptr;
      int st = prefilter ? pitt_prefilter_cloud(ctx, frames[i], stride_bytes, n_points[i], prefilter, &c, nullptr)
                         : pitt_stage_cloud(ctx, frames[i], stride_bytes, n_points[i], &c);
      if (st == PITT_OK) st = pitt_segment_frame(ctx, c, params, &results[i]);
      if (c) pitt_release_cloud(ctx, c);
      if (st != PITT_OK) {
        int expected = PITT_OK;
        first_error.compare_exchange_strong(expected, st);
      }
    }
  };
  std::vector<std::thread> threads;
  for (int t = 1; t < n_ctx; ++t) threads.emplace_back(worker, t);
  worker(0);
  for (auto& th : threads) th.join();
  return first_error.load();
}

int pitt_segment_frames_batched(pitt_ctx* const* ctxs, int n_ctx, const void* const* frames, const int* n_points,
                                int stride_bytes, int n_frames, const pitt_frame_params* params, pitt_frame_result* results) {
  return pitt_segment_raw_frames_batched(ctxs, n_ctx, frames, n_points, stride_bytes, n_frames, nullptr, params, results);
}

int pitt_segment_clouds_batched(pitt_ctx* const* ctxs, int n_ctx, const pitt_cloud* const* clouds, int n_frames,
                                const pitt_frame_params* params, pitt_frame_result* results) {
  if (!ctxs || n_ctx <= 0) return PITT_ERR_CUDA;
  for (int t = 0; t < n_ctx; ++t)
    if (!ctxs[t]) return PITT_ERR_CUDA;
  if (n_frames < 0 || (n_frames > 0 && (!clouds || !results)) || !params)
    return fail(ctxs[0], PITT_ERR_INVALID, "pitt_segment_clouds_batched arguments");
  std::atomic<int> first_error(PITT_OK);
  auto worker = [&](int t) {
    for (int i = t; i < n_frames; i += n_ctx) {
      const int st = clouds[i] ? pitt_segment_frame(ctxs[t], clouds[i], params, &results[i]) : PITT_ERR_INVALID;
      if (st != PITT_OK) {
        int expected = PITT_OK;
        first_error.compare_exchange_strong(expected, st);
      }
    }
  };
  std::vector<std::thread> threads;
  for (int t = 1; t < n_ctx; ++t) threads.emplace_back(worker, t);
  worker(0);
  for (auto& th : threads) th.join();
  return first_error.load();
}

}  // extern "C"
