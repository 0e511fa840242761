// knn.cu — exact k-nearest-neighbour search (multi-level grid for large clouds, warp-cooperative all-pairs selection for
// object clusters) and the normal estimator (K8).
//
// Replaces ne.compute() of PCManager::estimateNormal (reference: src/point_cloud_library/
// pc_manager.cpp:68-78; pcl::NormalEstimation::computeFeature + KdTreeFLANN::nearestKSearch,
// SURVEY.md B.7/B.8): k neighbours including the query, ordered by (squared distance, index);
// float covariance accumulated in that order; eigen33; curvature; flip towards the viewpoint.
#include <climits>
#include <cmath>
#include <cstdlib>
#include <algorithm>

#include "cluster.cuh"
#include "grid.cuh"
#include "mgrid.cuh"
#include "pitt_math.cuh"

namespace pitt {

constexpr int KNN_KMAX = 64;

__device__ __forceinline__ bool cand_less(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }

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

// SEG: the cloud is a sequence of independent segments (the cluster clouds of a support, back to back): seg_off[0..*n_seg] are
// their offsets on the device, every query searches its own segment only. One launch then serves all clusters of a frame.
template <int MODE, bool SEG>
__global__ void __launch_bounds__(WS_WARPS * 32)
knn_brute_kernel(const float4* __restrict__ xyz, int n, int k, float vpx, float vpy, float vpz, int* __restrict__ out_idx,
                 float* __restrict__ out_sq, float4* __restrict__ out_nrm, const int* __restrict__ seg_off, const int* __restrict__ n_seg) {
  __shared__ float s_bd[WS_WARPS][64];
  __shared__ int s_bi[WS_WARPS][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qi = blockIdx.x * WS_WARPS + warp;  // one warp per query
  if (qi >= n) return;
  int sb = 0, se = n;  // the query's segment [sb, se)
  if (SEG) {
    const int ns = *n_seg;
    if (qi >= seg_off[ns]) return;
    int sidx = 0;
    for (int base = 0; base < ns; base += 32) {
      const bool below = (base + lane < ns) && seg_off[base + lane + 1] <= qi;
      sidx += __popc(__ballot_sync(0xffffffffu, below));
    }
    sb = seg_off[sidx];
    se = seg_off[sidx + 1];
  }
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
  const int n_chunks = ((se - sb) + 31) >> 5;
  const int c0 = (qi - sb) >> 5;
  for (int step = 0; step < 2 * n_chunks; ++step) {
    const int off = (step + 1) >> 1;
    const int c = (step & 1) ? c0 - off : c0 + off;  // c0, c0-1, c0+1, c0-2, ...
    if (step == 0 ? false : (c == c0)) continue;
    if (c < 0 || c >= n_chunks) continue;
    const int pi = sb + (c << 5) + lane;
    float d = CUDART_INF_F;
    if (pi < se) {
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
  const float4 p0 = (ki[0] < se) ? __ldg(xyz + ki[0]) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 p1 = (ki[1] < se) ? __ldg(xyz + ki[1]) : make_float4(0.f, 0.f, 0.f, 0.f);
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


// =====================================================================================================================
// Multi-level grid + exact k-NN for large clouds (the 307 200-point frame of the headline metric).
//
// A Kinect-shaped frame is a surface seen in perspective: its point density varies by more than an order of magnitude
// between the near and the far end of the table, so no single cell size of a uniform grid suits every query. The grid here
// is ONE dense table of fine cells laid out block-Morton: blocks of 8 x 8 x 8 fine cells are ordered linearly, the 512
// cells inside a block in Morton (z-order) code. An aligned cube of 2^l fine cells per side (l = 0..3) is then a contiguous
// range of the table AND of the counting-sorted point array, i.e. the same table serves four cell sizes hf, 2hf, 4hf, 8hf.
// Everything is built on the device from device-resident geometry (no host round trip): bounding box -> geometry -> cell
// histogram -> block totals -> scan of the block totals -> per-block scan -> scatter (the histogram counts back down to zero).
//
// Fast path, two kernels (one thread per query, 32 consecutive queries of the Morton-sorted array per warp: neighbouring
// queries walk the same cells in step, so a candidate load is one broadcast for the warp):
//   knn_collect_kernel  a query picks the finest level whose PARENT cell holds >= `need` points and takes the 3 x 3 x 3 cells
//           around it at that level (their non-empty ranges looked up once into a shared-memory list);
//           pass 1: histogram of the squared distances (32 buckets of 1/8 octave below the distance to the faces of the
//           27-cell block, inside which every point closer than that is guaranteed to lie) -> the bucket edge T at which the
//           cumulative count reaches k; usable when T exists and at most 64 candidates lie below it;
//           pass 2: the <= 64 candidates with d2 <= T go to global memory as 64-bit keys (bits(d2) << 32 | index: unsigned
//           order == (distance, index) order, the tie rule of the oracle), laid out [slot][query]. Few registers (64), 8 CTAs per
//           SM: the latency of the candidate streams is hidden by switching warps.
//   knn_finish_kernel   sorts a query's candidates in REGISTERS with Batcher's odd-even merge network (543 compare-exchanges,
//           compile-time indices) on one 32-bit word per candidate (see knn_sort_body), then the first k in order: neighbour
//           lists, or the sequential float covariance + eigen33 + flip of NormalEstimation::computeFeature,
//           bit-identical to the warp-cooperative kernels. Its first CTAs serve the handed-over queries (below).
// No heap, no data-dependent sift loops, branch-free inner loops (rejected candidates go to a spare histogram row).
// knn_wide_body: the general exact search, one warp per query (ballot compaction + register bitonic top-64 like the
// all-pairs kernel), over the 27 cells at rising levels and then over growing cubes of the coarsest cells, for the queries the
// fast path hands over (96 of the 307 200 of a frame: sparse corners, density steps, > 64 candidates in one bucket,
// duplicates) and for k > KNN_FAST_KMAX. Measured on B200 (307 200-point frame, k = 50, profiles/r02_knn_ncu.md): build
// 0.05 ms, whole normal estimation 0.59-0.61 ms alone on the GPU (134 M + 38 M warp instructions in the two kernels); the
// round-1 heap kernel took 1.01 ms after a 0.15 ms build with two host round trips.
// =====================================================================================================================
constexpr int KNN_FAST_KMAX = 56;      // 64-key buffer: k plus the contents of one 1/8-octave bucket
constexpr int KNN_FAST_TPB = 128;

// scratch ints: [0..2] min, [3..5] max (ordered ints), [6] finite points, [7] fallback queue length
__global__ void mg_init_kernel(int* __restrict__ bb) {
  if (threadIdx.x < 3) bb[threadIdx.x] = INT_MAX;
  else if (threadIdx.x < 6) bb[threadIdx.x] = INT_MIN;
  else if (threadIdx.x < 8) bb[threadIdx.x] = 0;
}
__global__ void __launch_bounds__(256) mg_bbox_kernel(const float4* __restrict__ xyz, int n, int* __restrict__ bb) {
  __shared__ int s_red[8][7];
  int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
  int cnt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = __ldg(xyz + i);
    if (!mg_finite3(p)) continue;
    const int o[3] = {mg_f2ord(p.x), mg_f2ord(p.y), mg_f2ord(p.z)};
#pragma unroll
    for (int a = 0; a < 3; ++a) { mn[a] = min(mn[a], o[a]); mx[a] = max(mx[a], o[a]); }
    ++cnt;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mn[a] = __reduce_min_sync(0xffffffffu, mn[a]);
    mx[a] = __reduce_max_sync(0xffffffffu, mx[a]);
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { s_red[w][a] = mn[a]; s_red[w][3 + a] = mx[a]; }
    s_red[w][6] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 7) {  // one atomic per CTA and quantity (a warp-level atomic per quantity serialised 67 000 of them on the frame)
    int v = s_red[0][threadIdx.x];
    for (int i = 1; i < 8; ++i) {
      const int o = s_red[i][threadIdx.x];
      v = threadIdx.x < 3 ? min(v, o) : threadIdx.x < 6 ? max(v, o) : v + o;
    }
    if (threadIdx.x < 3) atomicMin(&bb[threadIdx.x], v);
    else if (threadIdx.x < 6) atomicMax(&bb[threadIdx.x], v);
    else if (v) atomicAdd(&bb[6], v);
  }
}
// geometry on the device: fine cell size from the mean surface density (about c_avg points per fine cell if the cloud were a
// uniform sheet over the two largest extents), enlarged until the table fits; one thread
__global__ void mg_geom_kernel(const int* __restrict__ bb, float c_avg, float h_fixed, MGrid* __restrict__ G) {
  MGrid g;
  g.n_finite = bb[6];
  if (g.n_finite <= 0) {
    g.mnx = g.mny = g.mnz = 0.0f; g.hf = g.inv_hf = 1.0f;
    g.dx = g.dy = g.dz = 8; g.nbx = g.nby = g.nbz = 1; g.nblocks = 1; g.ncells = 512;
    *G = g;
    return;
  }
  const float mn[3] = {mg_ord2f(bb[0]), mg_ord2f(bb[1]), mg_ord2f(bb[2])};
  const float mx[3] = {mg_ord2f(bb[3]), mg_ord2f(bb[4]), mg_ord2f(bb[5])};
  float e[3] = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
  float emax = fmaxf(e[0], fmaxf(e[1], e[2]));
  float emin = fminf(e[0], fminf(e[1], e[2]));
  float emid = (e[0] + e[1] + e[2]) - emax - emin;
  if (!(emax > 1e-30f)) emax = 1e-30f;
  const float area = fmaxf(emax * emid, emax * emax * 1e-6f);
  float h = sqrtf(c_avg * area / (float)g.n_finite);
  h = fmaxf(h, emax * 1e-5f);
  if (h_fixed > 0.0f) h = h_fixed;
  int d[3], nb[3];
  for (int it = 0; it < 200; ++it) {
    const float inv = 1.0f / h;
    double total = 1.0;
    bool fits = true;
    for (int a = 0; a < 3; ++a) {
      const float c = floorf((mx[a] - mn[a]) * inv);  // the cell of the largest coordinate, same expression as mg_coord
      if (!(c < 2.0e6f)) { fits = false; break; }
      d[a] = (int)c + 1;
      nb[a] = (d[a] + 7) >> 3;
      total *= (double)nb[a];
    }
    if (fits && total * 512.0 <= (double)MG_CAP_CELLS) {
      g.inv_hf = inv;
      break;
    }
    h *= 1.26f;
    g.inv_hf = 1.0f / h;
  }
  g.hf = 1.0f / g.inv_hf;
  g.mnx = mn[0]; g.mny = mn[1]; g.mnz = mn[2];
  g.nbx = nb[0]; g.nby = nb[1]; g.nbz = nb[2];
  g.dx = nb[0] << 3; g.dy = nb[1] << 3; g.dz = nb[2] << 3;
  g.nblocks = nb[0] * nb[1] * nb[2];
  g.ncells = g.nblocks << 9;
  *G = g;
}
__global__ void __launch_bounds__(256) mg_zero_kernel(const MGrid* __restrict__ G, int4* __restrict__ cnt4) {
  const int nc4 = (G->ncells >> 2) + 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc4; i += gridDim.x * blockDim.x) cnt4[i] = make_int4(0, 0, 0, 0);
}
__global__ void __launch_bounds__(256) mg_hist_kernel(const float4* __restrict__ xyz, int n, const MGrid* __restrict__ G,
                                                      int* __restrict__ cnt, int* __restrict__ cellid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const MGrid g = *G;
  const float4 p = __ldg(xyz + i);
  int cell = -1;
  if (mg_finite3(p)) {
    cell = mg_index(g, mg_coord(p.x, g.mnx, g.inv_hf, g.dx), mg_coord(p.y, g.mny, g.inv_hf, g.dy), mg_coord(p.z, g.mnz, g.inv_hf, g.dz));
    atomicAdd(&cnt[cell], 1);
  }
  cellid[i] = cell;
}
// block totals: one warp per block of 512 cells
__global__ void __launch_bounds__(256) mg_blocksum_kernel(const MGrid* __restrict__ G, const int* __restrict__ cnt, int* __restrict__ bsum) {
  const int blk = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (blk >= G->nblocks) return;
  const int4* src = reinterpret_cast<const int4*>(cnt + ((size_t)blk << 9) + lane * 16);
  int sum = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int4 q = src[j];
    sum += (q.x + q.y) + (q.z + q.w);
  }
  sum = __reduce_add_sync(0xffffffffu, sum);
  if (lane == 0) bsum[blk] = sum;
}
// exclusive scan of the block totals (at most MG_CAP_BLOCKS = 8192: 8 per thread), one CTA
__global__ void __launch_bounds__(1024) mg_top_kernel(const MGrid* __restrict__ G, const int* __restrict__ bsum, int* __restrict__ bbase,
                                                      int* __restrict__ start) {
  __shared__ int s_w[32];
  const int nblk = G->nblocks;
  int v[8], sum = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = threadIdx.x * 8 + j;
    v[j] = i < nblk ? bsum[i] : 0;
    sum += v[j];
  }
  int incl = sum;
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += y;
  }
  if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int w = s_w[threadIdx.x];
    int wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, wi, o);
      if ((int)threadIdx.x >= o) wi += y;
    }
    s_w[threadIdx.x] = wi - w;
  }
  __syncthreads();
  int run = s_w[threadIdx.x >> 5] + incl - sum;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = threadIdx.x * 8 + j;
    if (i < nblk) bbase[i] = run;
    run += v[j];
  }
  if (threadIdx.x == 1023) start[G->ncells] = run;  // == number of finite points
}
// one warp per block of 512 cells: lane l scans cells [16 l, 16 l + 16)
__global__ void __launch_bounds__(256) mg_blockscan_kernel(const MGrid* __restrict__ G, const int* __restrict__ cnt,
                                                           const int* __restrict__ bbase, int* __restrict__ start) {
  const int blk = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (blk >= G->nblocks) return;
  const int4* src = reinterpret_cast<const int4*>(cnt + ((size_t)blk << 9) + lane * 16);
  int v[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int4 q = src[j];
    v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
  }
  int sum = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) sum += v[j];
  int incl = sum;
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  int run = bbase[blk] + incl - sum;
  int4* dst = reinterpret_cast<int4*>(start + ((size_t)blk << 9) + lane * 16);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int4 q;
    q.x = run; run += v[4 * j];
    q.y = run; run += v[4 * j + 1];
    q.z = run; run += v[4 * j + 2];
    q.w = run; run += v[4 * j + 3];
    dst[j] = q;
  }
}
// seg_off != nullptr: the cloud is a sequence of segments (seg_off[0..*n_seg], on the device); a point carries its segment in
// the top 8 bits of the index word (n < 2^24), the searches then ignore candidates of other segments
__global__ void __launch_bounds__(256) mg_scatter_kernel(const float4* __restrict__ xyz, int n, const int* __restrict__ cellid,
                                                         const int* __restrict__ start, int* __restrict__ cnt,
                                                         float4* __restrict__ sorted, const int* __restrict__ seg_off,
                                                         const int* __restrict__ n_seg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cell = cellid[i];
  if (cell < 0) return;
  const float4 p = __ldg(xyz + i);
  int w = i;
  if (seg_off) {
    int lo = 0, hi = *n_seg;  // segment s: seg_off[s] <= i < seg_off[s + 1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (seg_off[mid] <= i) lo = mid;
      else hi = mid;
    }
    w = i | (lo << 24);
  }
  const int pos = start[cell] + atomicSub(&cnt[cell], 1) - 1;  // the histogram counts down to zero: no second table
  sorted[pos] = make_float4(p.x, p.y, p.z, __int_as_float(w));
}

// ---- per-query geometry at level l
// squared distance below which every point is guaranteed to lie inside the (2r+1)^3 cells around the query's level-l cell
// (faces on the border of the grid do not count: nothing lies beyond them); INF when the cube covers the whole grid
__device__ __forceinline__ float mg_bound(const MGrid& g, float4 q, int X, int Y, int Z, int l, int r) {
  float b = CUDART_INF_F;
  const int lo[3] = {X - r, Y - r, Z - r}, hi[3] = {X + r + 1, Y + r + 1, Z + r + 1};
  const int dim[3] = {g.dx, g.dy, g.dz};
  const float mn[3] = {g.mnx, g.mny, g.mnz}, qq[3] = {q.x, q.y, q.z};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (lo[a] > 0) b = fminf(b, qq[a] - (mn[a] + (float)(lo[a] << l) * g.hf));
    if ((hi[a] << l) < dim[a]) b = fminf(b, (mn[a] + (float)(hi[a] << l) * g.hf) - qq[a]);
  }
  b -= 2e-3f * g.hf;  // float rounding of the cell assignment
  return b;
}

// Compare-exchange of the sorting network on 32-bit words: one min and one max (two ALU instructions; the first version sorted
// the 64-bit (distance, index) keys with a 64-bit compare and four selects, six instructions and twice the registers).
#define CE(i, j)                               \
  {                                            \
    const unsigned lo_ = min(key[i], key[j]);  \
    const unsigned hi_ = max(key[i], key[j]);  \
    key[i] = lo_;                              \
    key[j] = hi_;                              \
  }

// finest level l such that the query's ancestor cell one level up holds at least `need` points (monotone in l), MG_MAXLVL if
// even the coarsest cell is that sparse. Judging by the PARENT cell keeps points that float next to a dense surface (object
// rims, mixed pixels) at a fine level: their own cells are almost empty at every level, the surface next to them is not.
__device__ __forceinline__ int mg_pick_level(const int* __restrict__ start, int m, int need) {
  int lvl = MG_MAXLVL;
#pragma unroll
  for (int l = MG_MAXLVL; l >= 1; --l) {
    const int base = (m >> (3 * l)) << (3 * l);
    const int c = __ldg(start + base + (1 << (3 * l))) - __ldg(start + base);
    if (c >= need) lvl = l - 1;
  }
  return lvl;
}

// The 27 level-l cells around (X, Y, Z), looked up ONCE per attempt: the range of every NON-EMPTY cell goes to the thread's
// column of a shared-memory list as one word (first point << 10 | number of points; the fast path is limited to clouds of
// <= 2^22 points and to attempts with <= mcap < 1024 candidates). Returns the number of points; ncell = non-empty cells (only
// the first KNN_CELLS_MAX are listed: a query with more of them is handed over).
// The block / Morton terms of z and y are hoisted out of the inner loops by hand (the compiler re-derived the whole index per
// cell: ~45 instructions and two dependent loads per cell in each of the three loops that walked the cells; now the two passes
// over the candidates read one shared-memory word per cell).
constexpr int KNN_CELLS_MAX = 21;  // 33 histogram rows + 21 cell rows of 128 B per warp: eight CTAs per SM
// (Measured and dropped: leaving out the cells that lie entirely beyond a query's histogram window. Which cells those are
// depends on where the query sits in its own cell, so the lanes of a warp - neighbouring queries that otherwise walk the same
// cells in step and share every candidate load as a broadcast - end up with different lists and different trip counts:
// warp instructions 151 M -> 164 M, 310 -> 371 us. The list below only drops cells that are empty, which is the same for
// every query of a cell.)
__device__ __forceinline__ int mg_block27_pack(const MGrid& g, const int* __restrict__ start, int X, int Y, int Z, int l,
                                               unsigned* __restrict__ ccol, int& ncell) {
  const int nx = g.dx >> l, ny = g.dy >> l, nz = g.dz >> l, span = 1 << (3 * l);
  int tot = 0, nc = 0;
  for (int c = -1; c <= 1; ++c) {
    const int z = Z + c, fz = z << l;
    const bool vz = (unsigned)z < (unsigned)nz;
    const int zb = (fz >> 3) * g.nby, zm = mg_spread3(fz & 7) << 2;
    for (int b = -1; b <= 1; ++b) {
      const int y = Y + b, fy = y << l;
      const bool vy = vz && (unsigned)y < (unsigned)ny;
      const int yb = (zb + (fy >> 3)) * g.nbx, ym = zm | (mg_spread3(fy & 7) << 1);
#pragma unroll
      for (int a = -1; a <= 1; ++a) {
        const int x = X + a, fx = x << l;
        if (vy && (unsigned)x < (unsigned)nx) {
          const int idx = ((yb + (fx >> 3)) << 9) | ym | mg_spread3(fx & 7);
          const int jb = __ldg(start + idx), cnt = __ldg(start + idx + span) - jb;
          if (cnt > 0) {
            tot += cnt;
            if (nc < KNN_CELLS_MAX) ccol[nc * 32] = ((unsigned)jb << 10) | (unsigned)min(cnt, 1023);
            ++nc;
          }
        }
      }
    }
  }
  ncell = nc;
  return tot;
}
// visits the points of those cells: f(point). Inside a cell the points are taken four at a time with the next four already
// requested (two register sets used alternately, no index clamping, no per-point validity), so the L1 / L2 latency of a lane's
// private candidate stream overlaps with the arithmetic of the previous four; the last one to three points of a cell follow
// one by one.
template <typename F>
__device__ __forceinline__ void mg_for_cells27(const float4* __restrict__ sorted, const unsigned* __restrict__ ccol, int ncell, F f) {
  for (int ci = 0; ci < ncell; ++ci) {
    const unsigned w = ccol[ci * 32];
    const int cnt = (int)(w & 1023u);
    const float4* p = sorted + (w >> 10);
    int grp = cnt >> 2;
    // the one to three points beyond the last whole group are requested together with the first group and handled first
    // (taken one by one after the groups, each of them exposed a full load latency: 12 % of the long-scoreboard stalls)
    const int rem = cnt & 3;
    const float4* pt = p + (cnt - rem);
    float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0, t2 = t0;
    if (rem > 0) t0 = __ldg(pt);
    if (rem > 1) t1 = __ldg(pt + 1);
    if (rem > 2) t2 = __ldg(pt + 2);
    if (grp) {
      float4 a0 = __ldg(p), a1 = __ldg(p + 1), a2 = __ldg(p + 2), a3 = __ldg(p + 3);
      p += 4;
      --grp;  // groups still to be requested
      if (rem > 0) f(t0);
      if (rem > 1) f(t1);
      if (rem > 2) f(t2);
      while (grp >= 2) {
        const float4 b0 = __ldg(p), b1 = __ldg(p + 1), b2 = __ldg(p + 2), b3 = __ldg(p + 3);
        f(a0); f(a1); f(a2); f(a3);
        a0 = __ldg(p + 4); a1 = __ldg(p + 5); a2 = __ldg(p + 6); a3 = __ldg(p + 7);
        p += 8;
        f(b0); f(b1); f(b2); f(b3);
        grp -= 2;
      }
      if (grp == 1) {
        const float4 b0 = __ldg(p), b1 = __ldg(p + 1), b2 = __ldg(p + 2), b3 = __ldg(p + 3);
        p += 4;
        f(a0); f(a1); f(a2); f(a3);
        f(b0); f(b1); f(b2); f(b3);
      } else {
        f(a0); f(a1); f(a2); f(a3);
      }
    } else {
      if (rem > 0) f(t0);
      if (rem > 1) f(t1);
      if (rem > 2) f(t2);
    }
  }
}


// Squared distance in the oracle's operation order, (dx*dx + dy*dy) + dz*dz with every operation rounded on its own, with
// the x and y lanes on the packed FP32 instructions of sm_100 (FADD2, FMUL2: two issue slots less per candidate). nqxy holds
// (-q.x, -q.y): p + (-q) == -(q - p) bit for bit and the square does not see the sign.
__device__ __forceinline__ unsigned long long knn_pack2(float lo, float hi) {
  return ((unsigned long long)__float_as_uint(hi) << 32) | (unsigned long long)__float_as_uint(lo);
}
__device__ __forceinline__ float knn_d2(unsigned long long nqxy, float qz, const float4 p) {
  unsigned long long d, sq;
  const unsigned long long pxy = knn_pack2(p.x, p.y);
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pxy), "l"(nqxy));
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(sq) : "l"(d));
  const float sx = __uint_as_float((unsigned)(sq & 0xffffffffull)), sy = __uint_as_float((unsigned)(sq >> 32));
  const float ddz = qz - p.z;
  return (sx + sy) + ddz * ddz;
}

// Fast path, first kernel: level choice, pass 1 (histogram -> threshold), pass 2 (the <= 64 candidates below the threshold as
// 64-bit keys). Few registers and 4 KB of shared memory per warp, so an SM holds many warps and the latency of the lane-private
// candidate streams is hidden by switching warps. The keys go to global memory laid out [slot][query] (coalesced: consecutive
// lanes are consecutive queries); ncol[query] = their number, or -1 when the query was handed to the warp-per-query kernel.
template <bool SEG>
__device__ __forceinline__ void
knn_collect_body(const MGrid* __restrict__ G, const int* __restrict__ start, const float4* __restrict__ sorted, int t_base, int t_count,
                 int k, int need, int mcap, unsigned long long* __restrict__ keys, int* __restrict__ ncol, int* __restrict__ fb_count,
                 int* __restrict__ fb_list, unsigned long long* __restrict__ dbg, unsigned (*s_hist)[33][32], unsigned (*s_cell)[KNN_CELLS_MAX][32],
                 unsigned long long* __restrict__ tline, int retry_up) {
  const MGrid g = *G;
  const int tl = blockIdx.x * KNN_FAST_TPB + threadIdx.x;  // query within this chunk
  const int t = t_base + tl;
  if (tl >= t_count) return;
  if (t >= g.n_finite) {  // beyond the finite points of the grid: nothing to sort either
    ncol[tl] = -1;
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned* hcol = &s_hist[warp][1][lane];  // row r of this thread's histogram: hcol[r * 32], r = -1 (outside the window) .. 31
  unsigned* ccol = &s_cell[warp][0][lane];  // cell c of this thread's 27: ccol[c * 32]
  const float4 q = __ldg(sorted + t);
  const unsigned qw = __float_as_uint(q.w);
  const unsigned long long nqxy = knn_pack2(-q.x, -q.y);
  const int cx = mg_coord(q.x, g.mnx, g.inv_hf, g.dx), cy = mg_coord(q.y, g.mny, g.inv_hf, g.dy), cz = mg_coord(q.z, g.mnz, g.inv_hf, g.dz);
  int lvl = mg_pick_level(start, mg_index(g, cx, cy, cz), need);
  // one attempt at this level (two when the 27 cells hold far more points than the level choice expected, i.e. at a density
  // step: finer cells first); whatever does not fit (fewer than k inside the guaranteed radius, more than 64 candidates below
  // the bucket edge) is handed to the warp-per-query kernel with a level hint
  int hand_over = -1;
  unsigned tsel = 0;
  int X, Y, Z, mtot, ncell = 0;
  int dir = 0;  // -1: the level has been lowered (too many points around), +1: raised (fewer than k inside the guaranteed radius)
  for (int attempt = 0;; ++attempt) {
    X = cx >> lvl; Y = cy >> lvl; Z = cz >> lvl;
    const float bound = mg_bound(g, q, X, Y, Z, lvl, 1);
    const float hl = g.hf * (float)(1 << lvl);
    const float top = (bound > 0.0f) ? fminf(bound * bound, 27.5f * hl * hl) : 0.0f;  // cube = whole grid: its diagonal
    const unsigned btop = __float_as_uint(top);
    mtot = mg_block27_pack(g, start, X, Y, Z, lvl, ccol, ncell);
    if (btop < (40u << 20) || attempt == 3) {
      hand_over = lvl;
      break;
    }
    if (mtot > mcap) {
      if (dir <= 0 && lvl > 0) { --lvl; dir = -1; continue; }
      hand_over = lvl;
      break;
    }
    if (ncell > KNN_CELLS_MAX) {  // more non-empty cells than the list holds (a volume rather than a surface)
      hand_over = lvl;
      break;
    }
#pragma unroll
    for (int b = 0; b < 32; ++b) hcol[b * 32] = 0u;
    // bucket qq = (btop - bits(d2)) >> 20 counts DOWN from the top of the window (31 = everything further below); a point
    // outside the window has a negative difference and lands in row -1, which nobody reads (shift, clamp, clamp, address:
    // no compare + select)
    mg_for_cells27(sorted, ccol, ncell, [&](const float4 p) {
      const int dd = (int)btop - __float_as_int(knn_d2(nqxy, q.z, p));  // both are bit patterns of floats >= 0
      int qq = max(min(dd >> 20, 31), -1);
      if (SEG && ((__float_as_uint(p.w) ^ qw) >> 24)) qq = -1;  // a point of another segment does not exist for this query
      atomicAdd(&hcol[qq * 32], 1u);  // the thread's own column: one shared-memory atomic instead of load + add + store
    });
    int cum = 0, bsel = -1, cat = 0;  // bsel in the ascending numbering: bucket b = 31 - qq
#pragma unroll
    for (int b = 0; b < 32; ++b) {
      cum += (int)hcol[(31 - b) * 32];
      if (bsel < 0 && cum >= k) { bsel = b; cat = cum; }
    }
    if (tline) {  // [3] candidates of histogram passes, [4] histogram passes, [5] hand-overs (pitt_debug_knn_timeline)
      atomicAdd(tline + 6 * blockIdx.x + 3, (unsigned long long)mtot);
      atomicAdd(tline + 6 * blockIdx.x + 4, 1ull);
    }
    if (dbg) {  // diagnostics (pitt_debug_knn_stats)
      atomicAdd(&dbg[lvl], 1ull);
      atomicAdd(&dbg[4], (unsigned long long)cum);
      if (bsel >= 0 && cat > 64) atomicAdd(&dbg[5 + (bsel == 0 ? 0 : 1)], 1ull);
      if (bsel < 0) atomicAdd(&dbg[7], 1ull);
      atomicMax(&dbg[8], (unsigned long long)mtot);
    }
    if (bsel < 0) {
      // fewer than k points inside the guaranteed radius: a cube twice as wide
      if (retry_up && dir >= 0 && lvl < MG_MAXLVL) { ++lvl; dir = 1; continue; }
      hand_over = min(lvl + 1, MG_MAXLVL + 1);
    } else if (cat > 64) {
      hand_over = lvl;  // the k-th neighbour lies far below the window, or > 64 candidates in one bucket
    } else {
      tsel = btop - ((unsigned)(31 - bsel) << 20);
    }
    break;
  }
  if (hand_over >= 0) {
    if (dbg && mtot > mcap) atomicAdd(&dbg[9], 1ull);
    if (tline) atomicAdd(tline + 6 * blockIdx.x + 5, 1ull);
    const int pos = atomicAdd(fb_count, 1);
    fb_list[pos] = t | (hand_over << 28);
    ncol[tl] = -1;
    return;
  }
  unsigned long long* kq = keys + tl;  // slot s of this query: kq[s * t_count_padded]
  const size_t stride = (size_t)((t_count + 31) & ~31);
  int c = 0;
  mg_for_cells27(sorted, ccol, ncell, [&](const float4 p) {
    const unsigned db = __float_as_uint(knn_d2(nqxy, q.z, p));
    if (db <= tsel && !(SEG && ((__float_as_uint(p.w) ^ qw) >> 24))) {
      kq[(size_t)c * stride] = ((unsigned long long)db << 32) | (unsigned long long)(unsigned)__float_as_int(p.w);
      ++c;
    }
  });
  ncol[tl] = c;
}

__device__ __forceinline__ unsigned long long knn_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// tline (pitt_debug_knn_timeline, else null): per CTA {first start, last end} in ns of %globaltimer and the SM it ran on
template <bool SEG>
__global__ void __launch_bounds__(KNN_FAST_TPB, 8)
knn_collect_kernel(const MGrid* __restrict__ G, const int* __restrict__ start, const float4* __restrict__ sorted, int t_base, int t_count,
                   int k, int need, int mcap, unsigned long long* __restrict__ keys, int* __restrict__ ncol, int* __restrict__ fb_count,
                   int* __restrict__ fb_list, unsigned long long* __restrict__ dbg, unsigned long long* __restrict__ tline, int retry_up) {
  __shared__ unsigned s_hist[KNN_FAST_TPB / 32][33][32];  // row 0: points outside the histogram window, rows 1..32: the buckets
  __shared__ unsigned s_cell[KNN_FAST_TPB / 32][KNN_CELLS_MAX][32];  // the non-empty cells of every thread's current attempt
  unsigned long long t0 = 0ull;
  if (tline) t0 = knn_globaltimer();
  knn_collect_body<SEG>(G, start, sorted, t_base, t_count, k, need, mcap, keys, ncol, fb_count, fb_list, dbg, s_hist, s_cell, tline, retry_up);
  if (tline) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) atomicMax(tline + 6 * blockIdx.x + 1, knn_globaltimer());
    if (threadIdx.x == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      tline[6 * blockIdx.x] = t0;
      tline[6 * blockIdx.x + 2] = smid;
    }
  }
}

// Fast path, second kernel: the <= 64 candidates of a query sorted in REGISTERS (Batcher's odd-even merge network, compile-time
// indices), then the first k in order: neighbour lists, or the sequential float covariance + eigen33 + flip of computeFeature.
//
// What is sorted is ONE 32-bit word per candidate: the bit pattern of d2 (a non-negative float: unsigned order == numeric order)
// with its low 6 bits replaced by the candidate's slot number. Two candidates whose d2 differ above those 6 bits are ordered
// exactly as the 64-bit (d2, index) keys would order them. Candidates that agree in the upper 26 bits ("ambiguous": closer than
// 64 ulps, about 1 % of the queries of a frame have such a pair) end up ADJACENT after the sort, in slot order; a bubble pass
// over the adjacent pairs compares their true 64-bit keys (fetched from the key array by slot) and swaps where needed, repeated
// until nothing moves, which sorts every ambiguous group exactly. The result is the order of the 64-bit keys, bit for bit
// (tie rule of the oracle: distance, then index), at a third of the compare-exchange instructions and half the registers.
// Empty slots get distinct pad words above every real one (bit 31 set).
template <int MODE, bool SEG, int KC>  // MODE 0: neighbour lists, 1: normals; KC > 0: k is this compile-time constant
__device__ __forceinline__ void
knn_sort_body(const float4* __restrict__ sorted, const float4* __restrict__ xyz, int t_base, int t_count, int k,
              const unsigned long long* __restrict__ keys, const int* __restrict__ ncol, float vpx, float vpy, float vpz,
              int* __restrict__ out_idx, float* __restrict__ out_sq, float4* __restrict__ out_nrm, int vblock,
              unsigned* __restrict__ s_idx /*MODE 1: [64][KNN_FAST_TPB] words of dynamic shared memory*/) {
  const int tl = vblock * KNN_FAST_TPB + threadIdx.x;
  if (tl >= t_count) return;
  const int c = ncol[tl];
  if (c < 0) return;  // handed over
  const size_t stride = (size_t)((t_count + 31) & ~31);
  const unsigned long long* kq = keys + tl;
  unsigned key[64];
  // The keys stream in coalesced (consecutive lanes = consecutive queries). The index words wait in shared memory, one column
  // per thread laid out [slot][thread] (any slot of 32 lanes = 32 banks), and are read back BY SLOT after the sort: fetching
  // them from the key array by slot instead put two dependent gathers (key array in DRAM, then the point) behind every
  // neighbour and left the kernel waiting on memory (long-scoreboard stalls 2.3 -> 6.3 per issue, measured).
  // All 64 slots are requested whatever c is (slots >= c hold stale words of the arena, masked below): loads behind a
  // per-slot predicate were issued a few at a time and every group paid a DRAM round trip (62 % of the kernel's long-scoreboard
  // stalls sat on the first use of a key). The whole column is pulled into L2 first, then read in batches of 16.
  unsigned* icol = s_idx + threadIdx.x;
#pragma unroll
  for (int s = 0; s < 64; ++s) asm volatile("prefetch.global.L2 [%0];" ::"l"(kq + (size_t)s * stride));
#pragma unroll
  for (int b = 0; b < 64; b += 16) {
    unsigned long long kv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) kv[j] = __ldg(kq + (size_t)(b + j) * stride);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int s = b + j;
      key[s] = (s < c) ? (((unsigned)(kv[j] >> 32) & ~63u) | (unsigned)s) : (0x80000000u | ((unsigned)s << 6));
      if (MODE == 1) icol[s * KNN_FAST_TPB] = (unsigned)(kv[j] & 0xffffffffull);
    }
  }
#include "knn_sort64.inc"
  {
    // smallest xor of adjacent words: < 64 <=> some adjacent pair agrees in the upper 26 bits
    unsigned mx = 0xffffffffu;
#pragma unroll
    for (int s = 0; s < 63; ++s) mx = min(mx, key[s] ^ key[s + 1]);
    if (mx < 64u) {  // rare: resolve the ambiguous groups with the true keys
      bool again;
      do {
        again = false;
#pragma unroll
        for (int s = 0; s < 63; ++s) {
          if ((key[s] ^ key[s + 1]) < 64u) {
            const unsigned long long a = __ldg(kq + (size_t)(key[s] & 63u) * stride), b = __ldg(kq + (size_t)(key[s + 1] & 63u) * stride);
            if (b < a) {
              const unsigned t_ = key[s];
              key[s] = key[s + 1];
              key[s + 1] = t_;
              again = true;
            }
          }
        }
      } while (again);
    }
  }
  const float4 q = __ldg(sorted + t_base + tl);
  const unsigned IDX = SEG ? 0x00ffffffu : 0xffffffffu;  // the segment lives in the top 8 bits of the index word
  const int qi = (int)(__float_as_uint(q.w) & IDX);
  const int kk = KC > 0 ? KC : k;
  if (MODE == 0) {
#pragma unroll
    for (int s = 0; s < KNN_FAST_KMAX; ++s) {
      if (s < kk) {
        const unsigned long long tk = __ldg(kq + (size_t)(key[s] & 63u) * stride);
        out_idx[(size_t)qi * kk + s] = (int)((unsigned)(tk & 0xffffffffull) & IDX);
        if (out_sq) out_sq[(size_t)qi * kk + s] = __uint_as_float((unsigned)(tk >> 32));
      }
    }
  } else {
    // computeMeanAndCovarianceMatrix: float accumulators in neighbour order (k >= 3 on this path: the cloud has > 4096 points)
    float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (KC > 0) {
      // k known at compile time (the reference's k = 50): no branch between the gathers, so the loads of a batch are all in
      // flight before the first accumulation waits for one (with a runtime k every neighbour's load sat behind its own branch
      // and the L2 latency of 50 dependent gathers was exposed: 18 % of the kernel's stall samples). Same order of additions.
      constexpr int NB = 10;
#pragma unroll
      for (int b = 0; b < KC; b += NB) {
        unsigned ix[NB];
        float4 p[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j)
          if (b + j < KC) ix[j] = icol[(key[b + j] & 63u) * KNN_FAST_TPB] & IDX;
#pragma unroll
        for (int j = 0; j < NB; ++j)
          if (b + j < KC) p[j] = __ldg(xyz + (int)ix[j]);
#pragma unroll
        for (int j = 0; j < NB; ++j)
          if (b + j < KC) {
            accu[0] += p[j].x * p[j].x; accu[1] += p[j].x * p[j].y; accu[2] += p[j].x * p[j].z;
            accu[3] += p[j].y * p[j].y; accu[4] += p[j].y * p[j].z; accu[5] += p[j].z * p[j].z;
            accu[6] += p[j].x; accu[7] += p[j].y; accu[8] += p[j].z;
          }
      }
    } else {
#pragma unroll
      for (int s = 0; s < KNN_FAST_KMAX; ++s) {
        if (s < kk) {
          const unsigned ix = icol[(key[s] & 63u) * KNN_FAST_TPB] & IDX;
          const float4 p = __ldg(xyz + (int)ix);
          accu[0] += p.x * p.x; accu[1] += p.x * p.y; accu[2] += p.x * p.z;
          accu[3] += p.y * p.y; accu[4] += p.y * p.z; accu[5] += p.z * p.z;
          accu[6] += p.x; accu[7] += p.y; accu[8] += p.z;
        }
      }
    }
    if (kk >= 3) {
      float cov[9], cen[3], ev, evec[3];
      cov_from_accu(accu, (float)kk, cov, cen);
      eigen33(cov, ev, evec);
      float nx = evec[0], ny = evec[1], nz = evec[2];
      const float eig_sum = cov[0] + cov[4] + cov[8];
      const float curv = (eig_sum != 0.0f) ? fabsf(ev / eig_sum) : 0.0f;
      const float vx = vpx - q.x, vy = vpy - q.y, vz = vpz - q.z;
      const float cos_theta = (vx * nx + vy * ny + vz * nz);
      if (cos_theta < 0.0f) { nx = -nx; ny = -ny; nz = -nz; }
      out_nrm[qi] = make_float4(nx, ny, nz, curv);
    }
  }
}
#undef CE

// General exact search, one WARP per query (the selection machinery of knn_brute_kernel: ballot compaction of the candidates
// that beat the current k-th best, register bitonic top-64, ties by index): the candidates are the 27 cells around the query at
// its level; if the k-th best found there is not closer than the guaranteed radius, the level goes up; above MG_MAXLVL the
// whole point array is the candidate set, which always ends the search. Serves the queries the fast kernel hands over (list,
// *n_list on the device) and, with list == nullptr, every point (k > KNN_FAST_KMAX).
struct KnnWideSmem {  // per warp
  float bd[64];
  int bi[64];
  int rb[32], pre[32];
};
template <int MODE, bool SEG, int NW>  // NW warps per CTA; vblock of vgrid CTAs take this role
__device__ __forceinline__ void
knn_wide_body(const MGrid* __restrict__ G, const int* __restrict__ start, const float4* __restrict__ sorted,
              const float4* __restrict__ xyz, int n_xyz, int k, int need, float vpx, float vpy, float vpz, int* __restrict__ out_idx,
              float* __restrict__ out_sq, float4* __restrict__ out_nrm, const int* __restrict__ n_list, const int* __restrict__ list,
              KnnWideSmem* __restrict__ sm_all, int vblock, int vgrid) {
  const MGrid g = *G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* const s_bd_w = sm_all[warp].bd;
  int* const s_bi_w = sm_all[warp].bi;
  int* const s_rb_w = sm_all[warp].rb;
  int* const s_pre_w = sm_all[warp].pre;
  const int nq = list ? *n_list : g.n_finite;
  // a fixed grid of warps strides over the queries (launching a CTA per 8 potential queries, almost all of them empty, costs
  // more than the search itself)
  for (int slot = vblock * NW + warp; slot < nq; slot += vgrid * NW) {
  int t = slot, l = -1;
  if (list) {
    const int e = list[slot];
    t = e & 0x0fffffff;
    l = (e >> 28) & 7;
  }
  const float4 q = __ldg(sorted + t);
  const unsigned IDX = SEG ? 0x00ffffffu : 0xffffffffu;  // the segment lives in the top 8 bits of the index word
  const unsigned qw = __float_as_uint(q.w);
  const int qi = (int)(qw & IDX);
  const int cx = mg_coord(q.x, g.mnx, g.inv_hf, g.dx), cy = mg_coord(q.y, g.mny, g.inv_hf, g.dy), cz = mg_coord(q.z, g.mnz, g.inv_hf, g.dz);
  if (l < 0) l = mg_pick_level(start, mg_index(g, cx, cy, cz), need);
  const int want = min(k, g.n_finite);
  float kd[4];
  int ki[4];
  float thr_d;
  int thr_i, count;
  auto flush = [&]() {
    kd[2] = (lane < count) ? s_bd_w[lane] : CUDART_INF_F;
    ki[2] = (lane < count) ? s_bi_w[lane] : 0x7fffffff;
    kd[3] = (lane + 32 < count) ? s_bd_w[lane + 32] : CUDART_INF_F;
    ki[3] = (lane + 32 < count) ? s_bi_w[lane + 32] : 0x7fffffff;
    __syncwarp();
    ws_sort128(kd, ki, lane);
    count = 0;
    const int r = (k - 1) >> 5, ll = (k - 1) & 31;
    const float td = (r == 0) ? kd[0] : kd[1];
    const int ti = (r == 0) ? ki[0] : ki[1];
    thr_d = __shfl_sync(0xffffffffu, td, ll);
    thr_i = __shfl_sync(0xffffffffu, ti, ll);
  };
  // search sequence: the 27 cells at level l, l + 1, ... MG_MAXLVL, then cubes of 5^3, 7^3, ... cells at MG_MAXLVL; a cube that
  // covers the whole grid has an infinite guaranteed radius and ends the search
  l = min(l, MG_MAXLVL);
  for (int rad = 1;;) {
#pragma unroll
    for (int r = 0; r < 4; ++r) { kd[r] = CUDART_INF_F; ki[r] = 0x7fffffff; }
    thr_d = CUDART_INF_F;
    thr_i = 0x7fffffff;
    count = 0;
    const int X = cx >> l, Y = cy >> l, Z = cz >> l;
    const int side = 2 * rad + 1, ncell = side * side * side;
    const float bound = mg_bound(g, q, X, Y, Z, l, rad);
    for (int cb = 0; cb < ncell; cb += 32) {  // 32 cells at a time: one range per lane
      int jb = 0, len = 0;
      const int ci = cb + lane;
      if (ci < ncell) {
        const int x = X + (ci % side) - rad, y = Y + ((ci / side) % side) - rad, z = Z + (ci / (side * side)) - rad;
        if (x >= 0 && y >= 0 && z >= 0 && x < (g.dx >> l) && y < (g.dy >> l) && z < (g.dz >> l)) {
          const int idx = mg_index(g, x << l, y << l, z << l);
          jb = __ldg(start + idx);
          len = __ldg(start + idx + (1 << (3 * l))) - jb;
        }
      }
      int incl = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      const int mtot = __shfl_sync(0xffffffffu, incl, 31);
      if (mtot == 0) continue;
      __syncwarp();
      s_rb_w[lane] = jb;
      s_pre_w[lane] = incl - len;
      __syncwarp();
      for (int base = 0; base < mtot; base += 32) {
        const int fi = base + lane;
        float d = CUDART_INF_F;
        int pi = 0x7fffffff;
        if (fi < mtot) {
          int r = 0;
#pragma unroll
          for (int s = 16; s > 0; s >>= 1)
            if (s_pre_w[r + s] <= fi) r += s;
          const float4 p = __ldg(sorted + s_rb_w[r] + (fi - s_pre_w[r]));
          const float ddx = q.x - p.x, ddy = q.y - p.y, ddz = q.z - p.z;
          d = (ddx * ddx + ddy * ddy) + ddz * ddz;
          pi = __float_as_int(p.w);
          if (SEG && ((__float_as_uint(p.w) ^ qw) >> 24)) d = CUDART_INF_F;  // another segment: not a candidate
        }
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
            s_bd_w[pos] = d;
            s_bi_w[pos] = pi;
          }
          count += __popc(mask2);
          __syncwarp();
        }
      }
    }
    flush();
    // exact iff the k-th best is closer than everything outside the cube can be (an infinite bound: the cube is the grid)
    if (!(bound < CUDART_INF_F)) break;
    if (thr_d < CUDART_INF_F && bound > 0.0f && thr_d < bound * bound) break;
    if (l < MG_MAXLVL) ++l;
    else ++rad;
  }
  int size = 0;
  {
    const unsigned m0 = __ballot_sync(0xffffffffu, kd[0] < CUDART_INF_F), m1 = __ballot_sync(0xffffffffu, kd[1] < CUDART_INF_F);
    size = min(want, __popc(m0) + __popc(m1));
  }
  if (MODE == 0) {
    for (int tt = lane; tt < k; tt += 32) {
      const float dd_ = (tt < 32) ? kd[0] : kd[1];
      const int ii_ = (tt < 32) ? ki[0] : ki[1];
      out_idx[(size_t)qi * k + tt] = tt < size ? (int)((unsigned)ii_ & IDX) : -1;
      if (out_sq) out_sq[(size_t)qi * k + tt] = tt < size ? dd_ : CUDART_INF_F;
    }
    continue;
  }
  if (size < 3) continue;
  float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int i0 = (ki[0] == 0x7fffffff) ? -1 : (int)((unsigned)ki[0] & IDX), i1 = (ki[1] == 0x7fffffff) ? -1 : (int)((unsigned)ki[1] & IDX);
  const float4 p0 = (i0 >= 0 && i0 < n_xyz) ? __ldg(xyz + i0) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 p1 = (i1 >= 0 && i1 < n_xyz) ? __ldg(xyz + i1) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int tt = 0; tt < size; ++tt) {
    const float4 src = (tt < 32) ? p0 : p1;
    const float px = __shfl_sync(0xffffffffu, src.x, tt & 31), py = __shfl_sync(0xffffffffu, src.y, tt & 31),
                pz = __shfl_sync(0xffffffffu, src.z, tt & 31);
    accu[0] += px * px; accu[1] += px * py; accu[2] += px * pz;
    accu[3] += py * py; accu[4] += py * pz; accu[5] += pz * pz;
    accu[6] += px; accu[7] += py; accu[8] += pz;
  }
  if (lane != 0) continue;
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
  }  // queries of this warp
}


template <int MODE, bool SEG>
__global__ void __launch_bounds__(KNN_FAST_TPB, 6)
knn_sort_kernel(const float4* __restrict__ sorted, const float4* __restrict__ xyz, int t_base, int t_count, int k,
                const unsigned long long* __restrict__ keys, const int* __restrict__ ncol, float vpx, float vpy, float vpz,
                int* __restrict__ out_idx, float* __restrict__ out_sq, float4* __restrict__ out_nrm) {
  extern __shared__ unsigned s_dyn_idx[];
  knn_sort_body<MODE, SEG, 0>(sorted, xyz, t_base, t_count, k, keys, ncol, vpx, vpy, vpz, out_idx, out_sq, out_nrm,
                              (int)gridDim.x - 1 - (int)blockIdx.x, s_dyn_idx);
}
template <int MODE, bool SEG>
__global__ void __launch_bounds__(WS_WARPS * 32)
knn_wide_kernel(const MGrid* __restrict__ G, const int* __restrict__ start, const float4* __restrict__ sorted,
                const float4* __restrict__ xyz, int n_xyz, int k, int need, float vpx, float vpy, float vpz, int* __restrict__ out_idx,
                float* __restrict__ out_sq, float4* __restrict__ out_nrm, const int* __restrict__ n_list, const int* __restrict__ list) {
  __shared__ KnnWideSmem sm[WS_WARPS];
  knn_wide_body<MODE, SEG, WS_WARPS>(G, start, sorted, xyz, n_xyz, k, need, vpx, vpy, vpz, out_idx, out_sq, out_nrm, n_list, list, sm,
                                     blockIdx.x, gridDim.x);
}
// Second kernel of the fast path when the cloud is one chunk: the first `wide_ctas` CTAs serve the handed-over queries a warp each
// (few, long searches: they start first and run beside the sorting CTAs instead of after them), the others sort.
template <int MODE, bool SEG, int KC>
__global__ void __launch_bounds__(KNN_FAST_TPB, 6)
knn_finish_kernel(const MGrid* __restrict__ G, const int* __restrict__ start, const float4* __restrict__ sorted,
                  const float4* __restrict__ xyz, int n_xyz, int t_count, int k, int need, const unsigned long long* __restrict__ keys,
                  const int* __restrict__ ncol, float vpx, float vpy, float vpz, int* __restrict__ out_idx, float* __restrict__ out_sq,
                  float4* __restrict__ out_nrm, const int* __restrict__ n_list, const int* __restrict__ list, int wide_ctas) {
  __shared__ KnnWideSmem sm[KNN_FAST_TPB / 32];
  extern __shared__ unsigned s_dyn_idx[];
  if ((int)blockIdx.x < wide_ctas)
    knn_wide_body<MODE, SEG, KNN_FAST_TPB / 32>(G, start, sorted, xyz, n_xyz, k, need, vpx, vpy, vpz, out_idx, out_sq, out_nrm, n_list,
                                                list, sm, blockIdx.x, wide_ctas);
  else
    knn_sort_body<MODE, SEG, KC>(sorted, xyz, 0, t_count, KC > 0 ? KC : k, keys, ncol, vpx, vpy, vpz, out_idx, out_sq, out_nrm,
                                 // last-written keys first: the collect kernel left the tail of the 157 MB key array in L2 and
                                 // its head in DRAM; walking it backwards the sorting CTAs start on the resident part
                                 (int)gridDim.x - 1 - (int)blockIdx.x, s_dyn_idx);
}

constexpr int KNN_TIMELINE_MAX = 8192;
unsigned long long* g_knn_timeline = nullptr;  // pitt_debug_knn_timeline: device buffer of 3 words per CTA of knn_collect_kernel
int g_knn_timeline_ctas = 0;
int g_knn_stats = 0;  // pitt_debug_knn_stats(ctx, out, enable): collect the diagnostics of knn_fast_kernel
static float knn_env(const char* name, float dflt) {
  const char* v = getenv(name);
  if (!v) return dflt;
  const float f = (float)atof(v);
  return f > 0.0f ? f : dflt;
}
// tuning knobs (measured on B200, 307 200-point frame, k = 50; the environment variables are for experiments only)
static float knn_c_avg() { static float v = knn_env("PITT_KNN_CAVG", 2.5f); return v; }             // mean points per fine cell
// queries with more points than this in their 27 cells go to the warp-per-query kernel
static int knn_mcap() { static float v = knn_env("PITT_KNN_MCAP", 768.0f); return std::min((int)v, 1023); }  // < 1024: the packed cell words  // measured 1536 / 768 / 512: 379 / 318 / 291 us collect, but 512 floods the wide kernel
// a query with fewer than k points inside the guaranteed radius: 1 = its lane repeats the search one level up (the other 31 lanes
// of the warp wait: 0.7 % of the queries make a fifth of the warps twice as long), 0.4 = it joins the hand-over list (measured: the warp-per-query searches cost more than the waiting, 1182 against 1520 frames/s)
static int knn_retry_up() { static float v = knn_env("PITT_KNN_RETRY", 1.0f); return v > 0.5f ? 1 : 0; }
// CTAs per SM that serve the handed-over queries inside the finishing launch
static int knn_wide_mult() { static float v = knn_env("PITT_KNN_WIDE", 2.0f); return std::max(1, (int)v); }
static int knn_seg_grid_min() { static float v = knn_env("PITT_KNN_SEG_MIN", 1500.0f); return (int)v; }
static int knn_need(int k) { static float f = knn_env("PITT_KNN_NEED", 1.2f); return std::max(8, (int)ceilf(f * (float)k)); }  // points in the parent cell

// Builds the multi-level grid of d_xyz[0..n) on ctx->stream; nothing here waits for the device. The two dense tables live
// in the context (allocated once, reused by every call: calls of one context are ordered on its stream), the n-sized arrays
// in the per-call arena.
int mgrid_build(pitt_ctx* ctx, const float4* d_xyz, int n, float h_fixed, MGridBuf* out, const int* d_seg_off, const int* d_n_seg) {
  if (!ctx->mg_tables) {
    const size_t bytes = (size_t)(2 * (MG_CAP_CELLS + 4) + 2 * MG_CAP_BLOCKS) * sizeof(int) + 256;
    PITT_CUDA(ctx, cudaMalloc(&ctx->mg_tables, bytes));
  }
  int* d_cnt = reinterpret_cast<int*>(ctx->mg_tables);
  int* d_start = d_cnt + MG_CAP_CELLS + 4;
  int* d_bsum = d_start + MG_CAP_CELLS + 4;
  int* d_bbase = d_bsum + MG_CAP_BLOCKS;
  MGrid* d_G = reinterpret_cast<MGrid*>(d_bbase + MG_CAP_BLOCKS);
  int* d_scr = nullptr;
  int* d_cellid = nullptr;
  PITT_TRY(arena_alloc(ctx, 8 + 32, &d_scr));
  PITT_CUDA(ctx, cudaMemsetAsync(d_scr + 8, 0, 32 * sizeof(int), ctx->stream));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_cellid));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &out->d_sorted));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &out->d_fb_list));
  const int pb = std::max(1, std::min(cdiv(n, 2048), ctx->sm_count * 2));
  mg_init_kernel<<<1, 32, 0, ctx->stream>>>(d_scr);
  mg_bbox_kernel<<<pb, 256, 0, ctx->stream>>>(d_xyz, n, d_scr);
  mg_geom_kernel<<<1, 1, 0, ctx->stream>>>(d_scr, knn_c_avg(), h_fixed, d_G);
  mg_zero_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_G, reinterpret_cast<int4*>(d_cnt));
  mg_hist_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_xyz, n, d_G, d_cnt, d_cellid);
  mg_blocksum_kernel<<<MG_CAP_BLOCKS / 8, 256, 0, ctx->stream>>>(d_G, d_cnt, d_bsum);
  mg_top_kernel<<<1, 1024, 0, ctx->stream>>>(d_G, d_bsum, d_bbase, d_start);
  mg_blockscan_kernel<<<MG_CAP_BLOCKS / 8, 256, 0, ctx->stream>>>(d_G, d_cnt, d_bbase, d_start);
  mg_scatter_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_xyz, n, d_cellid, d_start, d_cnt, out->d_sorted, d_seg_off, d_n_seg);
  ctx->launches += 9;
  PITT_CUDA(ctx, cudaGetLastError());
  out->d_G = d_G;
  out->d_start = d_start;
  out->d_scr = d_scr;
  return PITT_OK;
}

// dynamic shared memory of the sorting kernels: the index words of MODE 1, [64][KNN_FAST_TPB]
static size_t knn_sort_smem(int mode) { return mode == 1 ? (size_t)64 * KNN_FAST_TPB * sizeof(unsigned) : 0; }

// exact k-NN of every point of a large cloud; MODE 0 writes the neighbour lists, MODE 1 the normals
template <int MODE, bool SEG>
static int knn_large(pitt_ctx* ctx, const float4* d_xyz, int n, int k, const float vp[3], int* d_idx, float* d_sq, float4* d_nrm,
                     const int* d_seg_off = nullptr, const int* d_n_seg = nullptr) {
  MGridBuf mg;
  PITT_TRY(mgrid_build(ctx, d_xyz, n, 0.0f, &mg, d_seg_off, d_n_seg));
  const int need = knn_need(k);
  const int wide_grid = std::min(cdiv(n, WS_WARPS), ctx->sm_count * 8);
  if (k <= KNN_FAST_KMAX && n <= (1 << 22)) {  // the packed cell words hold 22 bits of point offset
    // the collected keys take 512 B per query: clouds of more than a million points go through in chunks of queries
    const int chunk = std::min(n, 1 << 20);
    const size_t stride = (size_t)((chunk + 31) & ~31);
    unsigned long long* d_keys = nullptr;
    int* d_ncol = nullptr;
    PITT_TRY(arena_alloc(ctx, stride * 64, &d_keys));
    PITT_TRY(arena_alloc(ctx, (size_t)chunk, &d_ncol));
    unsigned long long* dbg = g_knn_stats ? reinterpret_cast<unsigned long long*>(mg.d_scr + 8) : nullptr;
    unsigned long long* tline = nullptr;
    if (g_knn_timeline && !SEG && cdiv(chunk, KNN_FAST_TPB) <= KNN_TIMELINE_MAX) {
      tline = g_knn_timeline;
      g_knn_timeline_ctas = cdiv(chunk, KNN_FAST_TPB);
      PITT_CUDA(ctx, cudaMemsetAsync(tline, 0, (size_t)6 * KNN_TIMELINE_MAX * sizeof(unsigned long long), ctx->stream));
    }
    if (n <= chunk) {
      const int wide_ctas = std::min(cdiv(n, KNN_FAST_TPB / 32), ctx->sm_count * knn_wide_mult());
      knn_collect_kernel<SEG><<<cdiv(n, KNN_FAST_TPB), KNN_FAST_TPB, 0, ctx->stream>>>(mg.d_G, mg.d_start, mg.d_sorted, 0, n, k, need,
                                                                                     knn_mcap(), d_keys, d_ncol, mg.d_scr + 7,
                                                                                     mg.d_fb_list, dbg, tline, knn_retry_up());
      auto finish = (MODE == 1 && k == 50) ? knn_finish_kernel<MODE, SEG, (MODE == 1 ? 50 : 0)> : knn_finish_kernel<MODE, SEG, 0>;
      finish<<<wide_ctas + cdiv(n, KNN_FAST_TPB), KNN_FAST_TPB, knn_sort_smem(MODE), ctx->stream>>>(
          mg.d_G, mg.d_start, mg.d_sorted, d_xyz, n, n, k, need, d_keys, d_ncol, vp[0], vp[1], vp[2], d_idx, d_sq, d_nrm, mg.d_scr + 7,
          mg.d_fb_list, wide_ctas);
      ctx->launches += 2;
    } else {
      for (int t0 = 0; t0 < n; t0 += chunk) {
        const int tc = std::min(chunk, n - t0);
        knn_collect_kernel<SEG><<<cdiv(tc, KNN_FAST_TPB), KNN_FAST_TPB, 0, ctx->stream>>>(mg.d_G, mg.d_start, mg.d_sorted, t0, tc, k, need,
                                                                                        knn_mcap(), d_keys, d_ncol, mg.d_scr + 7,
                                                                                        mg.d_fb_list, dbg, tline, knn_retry_up());
        knn_sort_kernel<MODE, SEG><<<cdiv(tc, KNN_FAST_TPB), KNN_FAST_TPB, knn_sort_smem(MODE), ctx->stream>>>(mg.d_sorted, d_xyz, t0, tc, k, d_keys, d_ncol,
                                                                                            vp[0], vp[1], vp[2], d_idx, d_sq, d_nrm);
        ctx->launches += 2;
      }
      // the queries the fast path handed over, a warp each (their number is only known on the device: a fixed grid strides over them)
      knn_wide_kernel<MODE, SEG><<<wide_grid, WS_WARPS * 32, 0, ctx->stream>>>(mg.d_G, mg.d_start, mg.d_sorted, d_xyz, n, k, need, vp[0],
                                                                               vp[1], vp[2], d_idx, d_sq, d_nrm, mg.d_scr + 7, mg.d_fb_list);
      ctx->launches++;
    }
  } else {
    if (SEG) return fail(ctx, PITT_ERR_INVALID, "segmented k-NN: k <= 56 and fewer than 2^24 points");
    knn_wide_kernel<MODE, false><<<wide_grid, WS_WARPS * 32, 0, ctx->stream>>>(mg.d_G, mg.d_start, mg.d_sorted, d_xyz, n, k, need, vp[0], vp[1],
                                                                               vp[2], d_idx, d_sq, d_nrm, nullptr, nullptr);
    ctx->launches++;
  }
  PITT_CUDA(ctx, cudaGetLastError());
  ctx->knn_scr = mg.d_scr;  // diagnostics: pitt_debug_knn_stats
  return PITT_OK;
}

int estimate_normals_segmented(pitt_ctx* ctx, const float4* d_xyz, int n_total, const int* d_seg_off, const int* d_n_seg, int k,
                               const float vp[3], float4* d_nrm) {
  if (n_total <= 0) return PITT_OK;
  if (k < 1 || k > KNN_KMAX) return fail(ctx, PITT_ERR_INVALID, "k must be in [1, 64]");
  const float nan = nanf("");
  fill_f4_kernel<<<cdiv(n_total, 256), 256, 0, ctx->stream>>>(d_nrm, n_total, make_float4(nan, nan, nan, nan));
  ctx->launches++;
  if (n_total >= knn_seg_grid_min() && k <= KNN_FAST_KMAX && n_total <= (1 << 22)) {
    // enough points for the grid to pay: ONE multi-level grid over all segments, every point tagged with its segment, the
    // searches ignore candidates of other segments (6042 cluster points of a frame: 113 us all-pairs -> see profiles/)
    return knn_large<1, true>(ctx, d_xyz, n_total, k, vp, nullptr, nullptr, d_nrm, d_seg_off, d_n_seg);
  }
  knn_brute_kernel<1, true><<<cdiv(n_total, WS_WARPS), WS_WARPS * 32, 0, ctx->stream>>>(d_xyz, n_total, k, vp[0], vp[1], vp[2], nullptr, nullptr,
                                                                                         d_nrm, d_seg_off, d_n_seg);
  ctx->launches++;
  PITT_CUDA(ctx, cudaGetLastError());
  return PITT_OK;
}

int estimate_normals_impl(pitt_ctx* ctx, const float4* d_xyz, int n, int k, const float vp[3], float4* d_nrm) {
  if (n <= 0) return PITT_OK;
  if (k < 1 || k > KNN_KMAX) return fail(ctx, PITT_ERR_INVALID, "k must be in [1, 64]");
  const float nan = nanf("");
  fill_f4_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_nrm, n, make_float4(nan, nan, nan, nan));
  ctx->launches++;
  if (n <= KNN_BRUTE_MAX) {
    knn_brute_kernel<1, false><<<cdiv(n, WS_WARPS), WS_WARPS * 32, 0, ctx->stream>>>(d_xyz, n, k, vp[0], vp[1], vp[2], nullptr, nullptr, d_nrm,
                                                                                          nullptr, nullptr);
    ctx->launches++;
    PITT_CUDA(ctx, cudaGetLastError());
    return PITT_OK;
  }
  return knn_large<1, false>(ctx, d_xyz, n, k, vp, nullptr, nullptr, d_nrm);
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
      knn_brute_kernel<0, false><<<cdiv(n, WS_WARPS), WS_WARPS * 32, 0, ctx->stream>>>(c->d_xyz, n, k, 0.f, 0.f, 0.f, d_idx, d_sq, nullptr, nullptr,
                                                                                          nullptr);
      ctx->launches++;
    } else {
      const float vp[3] = {0.f, 0.f, 0.f};
      PITT_TRY((knn_large<0, false>(ctx, c->d_xyz, n, k, vp, d_idx, d_sq, nullptr)));
    }
    PITT_CUDA(ctx, cudaMemcpyAsync(out_idx, d_idx, tot * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_sqdist) PITT_CUDA(ctx, cudaMemcpyAsync(out_sqdist, d_sq, tot * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
  }
  timer.finish();
  return PITT_OK;
}

/* test / measurement hook (include/pitt_b200_debug.h): enable = collect diagnostics in the following large-cloud k-NN calls;
 * out16 (nullable) = those of the last call of this context (waits for the stream): [0] finite points, [1] queries that took
 * the ring search, [2..5] fast-path attempts at level 0..3, [6] candidates inside the guaranteed radius (sum over attempts),
 * [7] / [8] attempts with more than 64 candidates below the threshold (first bucket / later bucket), [9] attempts with fewer
 * than k candidates inside the guaranteed radius, [10] largest number of candidates of one attempt */
int pitt_debug_knn_stats(pitt_ctx* ctx, int enable, int64_t* out16) {
  g_knn_stats = enable;
  if (!out16) return PITT_OK;
  if (!ctx || !ctx->knn_scr) return PITT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  int h[8 + 32] = {0};
  PITT_CUDA(ctx, cudaMemcpyAsync(h, ctx->knn_scr, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  for (int i = 0; i < 16; ++i) out16[i] = 0;
  out16[0] = h[6];
  out16[1] = h[7];
  const unsigned long long* d = reinterpret_cast<const unsigned long long*>(h + 8);
  for (int i = 0; i < 9; ++i) out16[2 + i] = (int64_t)d[i];
  return PITT_OK;
}

/* measurement hook (include/pitt_b200_debug.h): enable != 0 arms the recording for the following large-cloud k-NN calls of the
 * process (one context at a time); out (nullable, 6 * cap words) receives {start ns, end ns, SM, candidates of the histogram passes, histogram passes, hand-overs} of every CTA of the last
 * knn_collect_kernel<unsegmented> launch. Returns the number of CTAs recorded, or a negative status. */
int pitt_debug_knn_timeline(pitt_ctx* ctx, int enable, uint64_t* out, int cap) {
  if (!ctx) return PITT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (enable && !g_knn_timeline) {
    if (cudaMalloc((void**)&g_knn_timeline, (size_t)6 * KNN_TIMELINE_MAX * sizeof(unsigned long long)) != cudaSuccess) return PITT_ERR_CUDA;
  }
  int n = 0;
  if (out && g_knn_timeline) {
    n = std::min(cap, g_knn_timeline_ctas);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return PITT_ERR_CUDA;
    if (cudaMemcpy(out, g_knn_timeline, (size_t)6 * n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return PITT_ERR_CUDA;
  }
  if (!enable && g_knn_timeline) {
    cudaFree(g_knn_timeline);
    g_knn_timeline = nullptr;
  }
  return n;
}

}  // extern "C"
