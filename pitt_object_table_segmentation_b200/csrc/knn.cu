// knn.cu — exact k-nearest-neighbour search on the uniform grid and the normal estimator (K8).
//
// Replaces ne.compute() of PCManager::estimateNormal (reference: src/point_cloud_library/
// pc_manager.cpp:68-78; pcl::NormalEstimation::computeFeature + KdTreeFLANN::nearestKSearch,
// SURVEY.md B.7/B.8): k neighbours including the query, ordered by (squared distance, index);
// float covariance accumulated in that order; eigen33; curvature; flip towards the viewpoint.
#include <cmath>

#include "grid.cuh"
#include "pitt_math.cuh"

namespace pitt {

constexpr int KNN_KMAX = 64;
constexpr int KNN_BRUTE_MAX = 4096;  // below this size the all-pairs kernel beats grid + ring search

__device__ __forceinline__ bool cand_less(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }

// max-heap on (d, i). The heap of one thread lives in SHARED memory with the thread index as the
// fastest dimension (entry p of thread t at [p * KNN_TPB + t]): whatever heap positions the 32 lanes of
// a warp touch, they hit 32 different banks, so every access is one conflict-free wavefront. (A
// thread-local array goes through L1 as local memory and divergent positions cost one sector each.)
constexpr int KNN_TPB = 128;
struct Heap {
  float* d;  // [k][KNN_TPB]
  int* i;
  __device__ __forceinline__ float& D(int p) const { return d[p * KNN_TPB]; }
  __device__ __forceinline__ int& I(int p) const { return i[p * KNN_TPB]; }
};
__device__ __forceinline__ void heap_sift_down(const Heap& h, int size, int pos) {
  float d = h.D(pos);
  int i = h.I(pos);
  for (;;) {
    int c = 2 * pos + 1;
    if (c >= size) break;
    float cd = h.D(c);
    int ci = h.I(c);
    if (c + 1 < size) {
      const float cd1 = h.D(c + 1);
      const int ci1 = h.I(c + 1);
      if (cand_less(cd, ci, cd1, ci1)) { ++c; cd = cd1; ci = ci1; }
    }
    if (!cand_less(d, i, cd, ci)) break;
    h.D(pos) = cd;
    h.I(pos) = ci;
    pos = c;
  }
  h.D(pos) = d;
  h.I(pos) = i;
}
__device__ __forceinline__ void heap_sift_up(const Heap& h, int pos) {
  float d = h.D(pos);
  int i = h.I(pos);
  while (pos > 0) {
    int p = (pos - 1) >> 1;
    const float pd = h.D(p);
    const int pi = h.I(p);
    if (!cand_less(pd, pi, d, i)) break;
    h.D(pos) = pd;
    h.I(pos) = pi;
    pos = p;
  }
  h.D(pos) = d;
  h.I(pos) = i;
}

template <int MODE>
__device__ __forceinline__ void knn_finish(const Heap& h, int size, int k, int qi, float4 q, const float4* __restrict__ xyz,
                                           float vpx, float vpy, float vpz, int* __restrict__ out_idx,
                                           float* __restrict__ out_sq, float4* __restrict__ out_nrm);

template <int MODE>  // 0: neighbour lists, 1: normals
__global__ void __launch_bounds__(KNN_TPB)
knn_kernel(GridDev g, const float4* __restrict__ xyz, int k, float vpx, float vpy, float vpz, int* __restrict__ out_idx,
           float* __restrict__ out_sq, float4* __restrict__ out_nrm) {
  extern __shared__ __align__(16) unsigned char knn_smem[];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.n) return;
  const int want = min(k, g.n);
  Heap hp;
  hp.d = reinterpret_cast<float*>(knn_smem) + threadIdx.x;
  hp.i = reinterpret_cast<int*>(knn_smem) + (size_t)k * KNN_TPB + threadIdx.x;
  const float4 q = g.sorted[t];
  const int qi = __float_as_int(q.w);
  int size = 0;
  float top_d = 0.0f;  // register copy of the heap maximum (valid when size == want)
  int top_i = 0;
  const int cx = grid_coord(q.x, g.mnx, g.inv_h, g.dx), cy = grid_coord(q.y, g.mny, g.inv_h, g.dy),
            cz = grid_coord(q.z, g.mnz, g.inv_h, g.dz);
  const int rmax = max(g.dx, max(g.dy, g.dz));
  const float slack = 2e-3f * g.h;
  for (int r = 0; r <= rmax; ++r) {
    const int z0 = max(cz - r, 0), z1 = min(cz + r, g.dz - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, g.dy - 1);
    const int x0 = max(cx - r, 0), x1 = min(cx + r, g.dx - 1);
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) {
        const bool face = (abs(z - cz) == r) || (abs(y - cy) == r);
        const int step = face ? 1 : max(1, x1 - x0);  // interior rows: only the two end cells
        for (int x = x0; x <= x1; x += step) {
          if (!face && abs(x - cx) != r) continue;
          const int cell = (z * g.dy + y) * g.dx + x;
          const int b = g.cell_start[cell], e = g.cell_start[cell + 1];
          if (b == e) continue;
          if (size == want) {
            // prune: no point of this cell can beat the current k-th best if the cell's box is farther away.
            // Walls are pulled in by slack = 2e-3 h (float rounding of the cell assignment), so the bound is safe.
            const float lox = g.mnx + (float)x * g.h, loy = g.mny + (float)y * g.h, loz = g.mnz + (float)z * g.h;
            const float gx = fmaxf(fmaxf(lox - q.x, q.x - (lox + g.h)) - slack, 0.0f);
            const float gy = fmaxf(fmaxf(loy - q.y, q.y - (loy + g.h)) - slack, 0.0f);
            const float gz = fmaxf(fmaxf(loz - q.z, q.z - (loz + g.h)) - slack, 0.0f);
            if ((gx * gx + gy * gy) + gz * gz > top_d) continue;
          }
          for (int j = b; j < e; ++j) {
            const float4 p = g.sorted[j];
            const float ddx = q.x - p.x, ddy = q.y - p.y, ddz = q.z - p.z;
            const float d = (ddx * ddx + ddy * ddy) + ddz * ddz;
            const int pi = __float_as_int(p.w);
            if (size < want) {
              hp.D(size) = d;
              hp.I(size) = pi;
              heap_sift_up(hp, size);
              ++size;
              if (size == want) { top_d = hp.D(0); top_i = hp.I(0); }
            } else if (cand_less(d, pi, top_d, top_i)) {
              hp.D(0) = d;
              hp.I(0) = pi;
              heap_sift_down(hp, size, 0);
              top_d = hp.D(0);
              top_i = hp.I(0);
            }
          }
        }
      }
    if (size >= want) {
      // every point outside the cube of r cells around the query cell is farther than `bound`
      float bound = 3.0e38f;
      if (cx - r > 0) bound = fminf(bound, q.x - (g.mnx + (float)(cx - r) * g.h));
      if (cx + r < g.dx - 1) bound = fminf(bound, (g.mnx + (float)(cx + r + 1) * g.h) - q.x);
      if (cy - r > 0) bound = fminf(bound, q.y - (g.mny + (float)(cy - r) * g.h));
      if (cy + r < g.dy - 1) bound = fminf(bound, (g.mny + (float)(cy + r + 1) * g.h) - q.y);
      if (cz - r > 0) bound = fminf(bound, q.z - (g.mnz + (float)(cz - r) * g.h));
      if (cz + r < g.dz - 1) bound = fminf(bound, (g.mnz + (float)(cz + r + 1) * g.h) - q.z);
      bound -= 2e-3f * g.h;  // float rounding of the cell assignment
      if (bound > 0.0f && top_d < bound * bound) break;
    }
  }
  knn_finish<MODE>(hp, size, k, qi, q, xyz, vpx, vpy, vpz, out_idx, out_sq, out_nrm);
}

// heap -> sorted neighbour list -> outputs (shared by the grid and the brute-force kernels)
template <int MODE>
__device__ __forceinline__ void knn_finish(const Heap& h, int size, int k, int qi, float4 q, const float4* __restrict__ xyz,
                                           float vpx, float vpy, float vpz, int* __restrict__ out_idx,
                                           float* __restrict__ out_sq, float4* __restrict__ out_nrm) {
  // heap sort -> ascending (distance, index)
  for (int s = size - 1; s > 0; --s) {
    float d = h.D(0);
    int i = h.I(0);
    h.D(0) = h.D(s);
    h.I(0) = h.I(s);
    h.D(s) = d;
    h.I(s) = i;
    heap_sift_down(h, s, 0);
  }
  if (MODE == 0) {
    for (int s = 0; s < k; ++s) {
      out_idx[(size_t)qi * k + s] = s < size ? h.I(s) : -1;
      if (out_sq) out_sq[(size_t)qi * k + s] = s < size ? h.D(s) : CUDART_INF_F;
    }
    return;
  }
  if (size < 3) return;  // stays NaN
  // computeMeanAndCovarianceMatrix: float accumulators in neighbour order
  float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < size; ++s) {
    const float4 p = __ldg(xyz + h.I(s));
    accu[0] += p.x * p.x; accu[1] += p.x * p.y; accu[2] += p.x * p.z;
    accu[3] += p.y * p.y; accu[4] += p.y * p.z; accu[5] += p.z * p.z;
    accu[6] += p.x; accu[7] += p.y; accu[8] += p.z;
  }
  float cov[9], cen[3], ev, evec[3];
  cov_from_accu(accu, (float)size, cov, cen);
  eigen33(cov, ev, evec);
  float nx = evec[0], ny = evec[1], nz = evec[2];
  const float eig_sum = cov[0] + cov[4] + cov[8];
  const float curv = (eig_sum != 0.0f) ? fabsf(ev / eig_sum) : 0.0f;
  const float vx = vpx - q.x, vy = vpy - q.y, vz = vpz - q.z;
  const float cos_theta = (vx * nx + vy * ny + vz * nz);
  if (cos_theta < 0.0f) { nx = -nx; ny = -ny; nz = -nz; }
  out_nrm[qi] = make_float4(nx, ny, nz, curv);
}

// ---------------------------------------------------------------------------------------------
// Small clouds (object clusters, a few thousand points): warp-cooperative exact selection.
// A warp serves 32 queries one after the other. For one query the 32 lanes stream all candidate
// points; candidates better than the current k-th best are appended to a 64-entry shared buffer
// (ballot compaction); when the buffer fills, the warp bitonic-sorts {best 64, buffer 64} held four
// per lane in registers. Unlike a per-thread heap this stays fast when the candidates arrive in
// scan order (organised clouds: every candidate beats the current k-th best until the query row).
// The sorted neighbour lists of the 32 queries go to shared memory; then lane l finishes query l
// (sequential float covariance in neighbour order, eigen33) exactly like the per-thread kernels.
// ---------------------------------------------------------------------------------------------
constexpr int WS_WARPS = 8;  // warps (= queries) per CTA
__device__ __forceinline__ void ws_cmpswap(float& da, int& ia, float& db, int& ib, bool up) {
  // after the call (a, b) is ordered ascending when up, descending otherwise
  const bool a_gt_b = cand_less(db, ib, da, ia);
  if (a_gt_b == up) {
    float td = da; da = db; db = td;
    int ti = ia; ia = ib; ib = ti;
  }
}
// full bitonic sort of 128 (d, i) keys: element position p = r*32 + lane, r = 0..3
__device__ __forceinline__ void ws_sort128(float (&kd)[4], int (&ki)[4], int lane) {
#pragma unroll
  for (int size = 2; size <= 128; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= 32) {
        const int rs = stride >> 5;  // partner differs in r
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if ((r & rs) == 0) {
            const int p = r * 32 + lane;
            const bool up = ((p & size) == 0) || size == 128;
            ws_cmpswap(kd[r], ki[r], kd[r | rs], ki[r | rs], up);
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int p = r * 32 + lane;
          const bool up = ((p & size) == 0) || size == 128;
          const float od = __shfl_xor_sync(0xffffffffu, kd[r], stride);
          const int oi = __shfl_xor_sync(0xffffffffu, ki[r], stride);
          const bool lower = (lane & stride) == 0;  // this lane holds the lower position of the pair
          // keep the smaller key at the lower position when ascending
          const bool other_less = cand_less(od, oi, kd[r], ki[r]);
          const bool take = (lower == up) ? other_less : !other_less && !(od == kd[r] && oi == ki[r]);
          if (take) { kd[r] = od; ki[r] = oi; }
        }
      }
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(WS_WARPS * 32)
knn_brute_kernel(const float4* __restrict__ xyz, int n, int k, float vpx, float vpy, float vpz, int* __restrict__ out_idx,
                 float* __restrict__ out_sq, float4* __restrict__ out_nrm) {
  __shared__ float s_bd[WS_WARPS][64];
  __shared__ int s_bi[WS_WARPS][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qi = blockIdx.x * WS_WARPS + warp;  // one warp per query
  if (qi >= n) return;
  const float4 q = __ldg(xyz + qi);
  if (!(isfinite(q.x) && isfinite(q.y) && isfinite(q.z))) return;  // outputs stay NaN / -1
  float kd[4];
  int ki[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { kd[r] = CUDART_INF_F; ki[r] = 0x7fffffff; }
  float thr_d = CUDART_INF_F;
  int thr_i = 0x7fffffff;
  int count = 0;  // entries in the shared buffer (warp uniform)
  auto flush = [&]() {
    kd[2] = (lane < count) ? s_bd[warp][lane] : CUDART_INF_F;
    ki[2] = (lane < count) ? s_bi[warp][lane] : 0x7fffffff;
    kd[3] = (lane + 32 < count) ? s_bd[warp][lane + 32] : CUDART_INF_F;
    ki[3] = (lane + 32 < count) ? s_bi[warp][lane + 32] : 0x7fffffff;
    __syncwarp();
    ws_sort128(kd, ki, lane);
    count = 0;
    // k-th best (position k-1) becomes the admission threshold
    const int r = (k - 1) >> 5, l = (k - 1) & 31;
    const float td = (r == 0) ? kd[0] : kd[1];
    const int ti = (r == 0) ? ki[0] : ki[1];
    thr_d = __shfl_sync(0xffffffffu, td, l);
    thr_i = __shfl_sync(0xffffffffu, ti, l);
  };
  // Candidate chunks of 32 are visited outwards from the query's own index: in organised
  // (scan-ordered) clouds index neighbours are spatial neighbours, so the admission threshold is
  // tight after the first few chunks and almost everything else is rejected by one compare.
  const int n_chunks = (n + 31) >> 5;
  const int c0 = qi >> 5;
  for (int step = 0; step < 2 * n_chunks; ++step) {
    const int off = (step + 1) >> 1;
    const int c = (step & 1) ? c0 - off : c0 + off;  // c0, c0-1, c0+1, c0-2, ...
    if (step == 0 ? false : (c == c0)) continue;
    if (c < 0 || c >= n_chunks) continue;
    const int pi = (c << 5) + lane;
    float d = CUDART_INF_F;
    if (pi < n) {
      const float4 p = __ldg(xyz + pi);
      const float ddx = q.x - p.x, ddy = q.y - p.y, ddz = q.z - p.z;
      d = (ddx * ddx + ddy * ddy) + ddz * ddz;
    }
    // non-finite candidates give inf / NaN distances and never pass
    const bool pass = (d < CUDART_INF_F) && cand_less(d, pi, thr_d, thr_i);
    const unsigned mask = __ballot_sync(0xffffffffu, pass);
    if (mask) {
      bool pass2 = pass;
      unsigned mask2 = mask;
      if (count + __popc(mask) > 64) {
        flush();
        pass2 = pass && cand_less(d, pi, thr_d, thr_i);  // the threshold is tighter after a flush
        mask2 = __ballot_sync(0xffffffffu, pass2);
      }
      if (pass2) {
        const int pos = count + __popc(mask2 & ((1u << lane) - 1u));
        s_bd[warp][pos] = d;
        s_bi[warp][pos] = pi;
      }
      count += __popc(mask2);
      __syncwarp();
    }
  }
  flush();
  // sorted list: position t lives in lane t&31, register t>>5 (t < 64)
  int size = 0;
  {
    const unsigned m0 = __ballot_sync(0xffffffffu, kd[0] < CUDART_INF_F), m1 = __ballot_sync(0xffffffffu, kd[1] < CUDART_INF_F);
    size = min(k, __popc(m0) + __popc(m1));
  }
  if (MODE == 0) {
    for (int t = lane; t < k; t += 32) {
      const float dd_ = (t < 32) ? kd[0] : kd[1];
      const int ii_ = (t < 32) ? ki[0] : ki[1];
      out_idx[(size_t)qi * k + t] = t < size ? ii_ : -1;
      if (out_sq) out_sq[(size_t)qi * k + t] = t < size ? dd_ : CUDART_INF_F;
    }
    return;
  }
  if (size < 3) return;
  // computeMeanAndCovarianceMatrix: sequential float accumulation in neighbour order. Every lane
  // runs the same sequence on broadcast neighbours (identical result), lane 0 stores it.
  float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float4 p0 = (ki[0] < n) ? __ldg(xyz + ki[0]) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 p1 = (ki[1] < n) ? __ldg(xyz + ki[1]) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < size; ++t) {
    const float4 src = (t < 32) ? p0 : p1;
    const float px = __shfl_sync(0xffffffffu, src.x, t & 31), py = __shfl_sync(0xffffffffu, src.y, t & 31),
                pz = __shfl_sync(0xffffffffu, src.z, t & 31);
    accu[0] += px * px; accu[1] += px * py; accu[2] += px * pz;
    accu[3] += py * py; accu[4] += py * pz; accu[5] += pz * pz;
    accu[6] += px; accu[7] += py; accu[8] += pz;
  }
  if (lane != 0) return;
  float cov[9], cen[3], ev, evec[3];
  cov_from_accu(accu, (float)size, cov, cen);
  eigen33(cov, ev, evec);
  float nx = evec[0], ny = evec[1], nz = evec[2];
  const float eig_sum = cov[0] + cov[4] + cov[8];
  const float curv = (eig_sum != 0.0f) ? fabsf(ev / eig_sum) : 0.0f;
  const float vx = vpx - q.x, vy = vpy - q.y, vz = vpz - q.z;
  const float cos_theta = (vx * nx + vy * ny + vz * nz);
  if (cos_theta < 0.0f) { nx = -nx; ny = -ny; nz = -nz; }
  out_nrm[qi] = make_float4(nx, ny, nz, curv);
}

__global__ void fill_f4_kernel(float4* p, int n, float4 v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void fill_knn_kernel(int* idx, float* sq, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    idx[i] = -1;
    if (sq) sq[i] = CUDART_INF_F;
  }
}

// occupied cells hold about k / KNN_CELL_DIV points (tuning knob, PITT_KNN_CELL_DIV overrides for experiments)
static float knn_target_per_cell(int k) {
  static float div = -1.0f;
  if (div < 0.0f) {
    const char* v = getenv("PITT_KNN_CELL_DIV");
    div = v ? (float)atof(v) : 6.0f;  // measured on B200: 4 -> 1.28 ms, 6 -> 1.17 ms, 12 -> 1.12 ms (307 200 points, k = 50)
    if (!(div > 0.0f)) div = 6.0f;
  }
  return fmaxf(2.0f, (float)k / div);
}
static size_t knn_smem_bytes(int k) { return (size_t)k * KNN_TPB * (sizeof(float) + sizeof(int)); }
// heaps of up to KNN_KMAX entries need 64 KB of dynamic shared memory per CTA: opt in once per device
static int knn_smem_opt_in(pitt_ctx* ctx) {
  static bool done[64] = {false};
  const int dev = ctx->device & 63;
  if (done[dev]) return PITT_OK;
  const int bytes = (int)knn_smem_bytes(KNN_KMAX);
  PITT_CUDA(ctx, cudaFuncSetAttribute(knn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  PITT_CUDA(ctx, cudaFuncSetAttribute(knn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done[dev] = true;
  return PITT_OK;
}

int estimate_normals_impl(pitt_ctx* ctx, const float4* d_xyz, int n, int k, const float vp[3], float4* d_nrm) {
  if (n <= 0) return PITT_OK;
  if (k < 1 || k > KNN_KMAX) return fail(ctx, PITT_ERR_INVALID, "k must be in [1, 64]");
  const float nan = nanf("");
  fill_f4_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_nrm, n, make_float4(nan, nan, nan, nan));
  ctx->launches++;
  if (n <= KNN_BRUTE_MAX) {
    knn_brute_kernel<1><<<cdiv(n, WS_WARPS), WS_WARPS * 32, 0, ctx->stream>>>(d_xyz, n, k, vp[0], vp[1], vp[2], nullptr, nullptr, d_nrm);
    ctx->launches++;
    PITT_CUDA(ctx, cudaGetLastError());
    return PITT_OK;
  }
  GridDev g;
  PITT_TRY(grid_build(ctx, d_xyz, n, -1.0f, knn_target_per_cell(k), &g));
  if (g.n <= 0) return PITT_OK;
  PITT_TRY(knn_smem_opt_in(ctx));
  knn_kernel<1><<<cdiv(g.n, KNN_TPB), KNN_TPB, knn_smem_bytes(k), ctx->stream>>>(g, d_xyz, k, vp[0], vp[1], vp[2], nullptr, nullptr, d_nrm);
  ctx->launches++;
  PITT_CUDA(ctx, cudaGetLastError());
  return PITT_OK;
}

}  // namespace pitt

using namespace pitt;

extern "C" {

int pitt_estimate_normals(pitt_ctx* ctx, pitt_cloud* c, int k, const float viewpoint[3]) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!c || !viewpoint) return fail(ctx, PITT_ERR_INVALID, "pitt_estimate_normals arguments");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  if (c->n > 0) {
    if (!c->d_nrm) PITT_TRY(pool_alloc(ctx, (size_t)c->n * sizeof(float4), (void**)&c->d_nrm));
    PITT_TRY(estimate_normals_impl(ctx, c->d_xyz, c->n, k, viewpoint, c->d_nrm));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
  }
  c->has_normals = true;
  timer.finish();
  return PITT_OK;
}

int pitt_knn(pitt_ctx* ctx, const pitt_cloud* c, int k, int32_t* out_idx, float* out_sqdist) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!c || !out_idx || k < 1 || k > KNN_KMAX) return fail(ctx, PITT_ERR_INVALID, "pitt_knn arguments (k in [1,64])");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  const int n = c->n;
  if (n > 0) {
    int* d_idx = nullptr;
    float* d_sq = nullptr;
    PITT_TRY(arena_alloc(ctx, (size_t)n * k, &d_idx));
    PITT_TRY(arena_alloc(ctx, (size_t)n * k, &d_sq));
    size_t tot = (size_t)n * k;
    fill_knn_kernel<<<(unsigned)cdiv64(tot, 256), 256, 0, ctx->stream>>>(d_idx, d_sq, tot);
    ctx->launches++;
    if (n <= KNN_BRUTE_MAX) {
      knn_brute_kernel<0><<<cdiv(n, WS_WARPS), WS_WARPS * 32, 0, ctx->stream>>>(c->d_xyz, n, k, 0.f, 0.f, 0.f, d_idx, d_sq, nullptr);
      ctx->launches++;
    } else {
      GridDev g;
      PITT_TRY(grid_build(ctx, c->d_xyz, n, -1.0f, knn_target_per_cell(k), &g));
      if (g.n > 0) {
        PITT_TRY(knn_smem_opt_in(ctx));
        knn_kernel<0><<<cdiv(g.n, KNN_TPB), KNN_TPB, knn_smem_bytes(k), ctx->stream>>>(g, c->d_xyz, k, 0.f, 0.f, 0.f, d_idx, d_sq, nullptr);
        ctx->launches++;
      }
    }
    PITT_CUDA(ctx, cudaMemcpyAsync(out_idx, d_idx, tot * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_sqdist) PITT_CUDA(ctx, cudaMemcpyAsync(out_sqdist, d_sq, tot * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
  }
  timer.finish();
  return PITT_OK;
}

}  // extern "C"
