// grid.cu — uniform grid construction: bounding box, cell histogram, exclusive scan, scatter.
#include <cfloat>
#include <cmath>

#include <math_constants.h>

#include "grid.cuh"

namespace pitt {

__device__ __forceinline__ int f2ord(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float ord2f_host(int i) {
  int j = i >= 0 ? i : i ^ 0x7fffffff;
  float f;
#ifdef __CUDA_ARCH__
  f = __int_as_float(j);
#else
  memcpy(&f, &j, 4);
#endif
  return f;
}
__device__ __forceinline__ bool finite3(float4 p) { return isfinite(p.x) && isfinite(p.y) && isfinite(p.z); }

__global__ void bbox_init_kernel(int* bb) {
  if (threadIdx.x < 3) bb[threadIdx.x] = INT_MAX;
  else if (threadIdx.x < 6) bb[threadIdx.x] = INT_MIN;
  else if (threadIdx.x == 6) bb[6] = 0;
}
__global__ void __launch_bounds__(256) bbox_kernel(const float4* __restrict__ xyz, int n, int* __restrict__ bb) {
  int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
  int cnt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = __ldg(xyz + i);
    if (!finite3(p)) continue;
    int o[3] = {f2ord(p.x), f2ord(p.y), f2ord(p.z)};
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
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { atomicMin(&bb[a], mn[a]); atomicMax(&bb[3 + a], mx[a]); }
    if (cnt) atomicAdd(&bb[6], cnt);
  }
}

struct GridGeom {
  float mnx, mny, mnz, inv_h;
  int dx, dy, dz;
};
__device__ __forceinline__ int geom_cell(const GridGeom& g, float4 p) {
  int cx = min(max((int)floorf((p.x - g.mnx) * g.inv_h), 0), g.dx - 1);
  int cy = min(max((int)floorf((p.y - g.mny) * g.inv_h), 0), g.dy - 1);
  int cz = min(max((int)floorf((p.z - g.mnz) * g.inv_h), 0), g.dz - 1);
  return (cz * g.dy + cy) * g.dx + cx;
}
__global__ void __launch_bounds__(256) cell_hist_kernel(const float4* __restrict__ xyz, int n, GridGeom g, int* __restrict__ cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = __ldg(xyz + i);
  if (!finite3(p)) return;
  atomicAdd(&cnt[geom_cell(g, p)], 1);
}
__global__ void __launch_bounds__(256) count_nonempty_kernel(const int* __restrict__ cnt, int ncells, int* __restrict__ out) {
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ncells; i += gridDim.x * blockDim.x) c += cnt[i] != 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
__global__ void __launch_bounds__(256) cell_scatter_kernel(const float4* __restrict__ xyz, int n, GridGeom g,
                                                           int* __restrict__ cursor, float4* __restrict__ sorted) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = __ldg(xyz + i);
  if (!finite3(p)) return;
  int pos = atomicAdd(&cursor[geom_cell(g, p)], 1);
  sorted[pos] = make_float4(p.x, p.y, p.z, __int_as_float(i));
}

// ---- multi-block exclusive scan (3 phases), generic in the value type and the associative op
constexpr int SCAN_TPB = 256, SCAN_IPT = 8, SCAN_CHUNK = SCAN_TPB * SCAN_IPT;
struct OpSumI {
  typedef int T;
  __device__ static int identity() { return 0; }
  __device__ static int apply(int a, int b) { return a + b; }
};
struct OpMaxF {
  typedef float T;
  __device__ static float identity() { return -CUDART_INF_F; }
  __device__ static float apply(float a, float b) { return fmaxf(a, b); }
};
template <typename Op>
__device__ __forceinline__ typename Op::T warp_incl_scan(typename Op::T v) {
  for (int o = 1; o < 32; o <<= 1) {
    typename Op::T y = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v = Op::apply(y, v);
  }
  return v;
}
template <typename Op>
__global__ void __launch_bounds__(SCAN_TPB) scan_local_kernel(typename Op::T* __restrict__ data, int n,
                                                              typename Op::T* __restrict__ block_sums) {
  typedef typename Op::T T;
  __shared__ T s_w[SCAN_TPB / 32];
  const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_IPT;
  T v[SCAN_IPT], sum = Op::identity();
#pragma unroll
  for (int j = 0; j < SCAN_IPT; ++j) {
    v[j] = (base + j < n) ? data[base + j] : Op::identity();
    sum = Op::apply(sum, v[j]);
  }
  T incl = warp_incl_scan<Op>(sum);
  if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
  __syncthreads();
  T run = Op::identity();
  for (int w = 0; w < (threadIdx.x >> 5); ++w) run = Op::apply(run, s_w[w]);
  // exclusive prefix of this thread = warps before + lanes before
  T lane_excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if ((threadIdx.x & 31) != 0) run = Op::apply(run, lane_excl);
#pragma unroll
  for (int j = 0; j < SCAN_IPT; ++j) {
    if (base + j < n) data[base + j] = run;
    run = Op::apply(run, v[j]);
  }
  if (threadIdx.x == SCAN_TPB - 1) block_sums[blockIdx.x] = run;
}
template <typename Op>
__global__ void scan_sums_kernel(typename Op::T* __restrict__ v, int nb, typename Op::T* __restrict__ total) {
  // single block, any nb: serial over chunks of blockDim
  typedef typename Op::T T;
  __shared__ T s_carry;
  __shared__ T s_w[32];
  if (threadIdx.x == 0) s_carry = Op::identity();
  __syncthreads();
  for (int base = 0; base < nb; base += blockDim.x) {
    int i = base + threadIdx.x;
    T x = (i < nb) ? v[i] : Op::identity();
    T incl = warp_incl_scan<Op>(x);
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    T excl = s_carry;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) excl = Op::apply(excl, s_w[w]);
    T lane_excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if ((threadIdx.x & 31) != 0) excl = Op::apply(excl, lane_excl);
    if (i < nb) v[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = Op::apply(excl, x);
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = s_carry;
}
template <typename Op>
__global__ void __launch_bounds__(SCAN_TPB) scan_add_kernel(typename Op::T* __restrict__ data, int n,
                                                            const typename Op::T* __restrict__ block_offs) {
  const typename Op::T off = block_offs[blockIdx.x];
  const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_IPT;
#pragma unroll
  for (int j = 0; j < SCAN_IPT; ++j)
    if (base + j < n) data[base + j] = Op::apply(off, data[base + j]);
}

// ---- the same scan in ONE launch for consumers that can add the chunk offset themselves: every CTA scans its chunk of
// SCAN_CHUNK elements in place (exclusive, relative to the chunk) and publishes the chunk total; the last CTA to finish (atomic
// ticket) turns the totals into exclusive chunk offsets. Element i of the full scan = Op(chunk_off[i / SCAN_CHUNK], data[i]).
template <typename Op>
__global__ void __launch_bounds__(SCAN_TPB) scan_chunks_kernel(typename Op::T* __restrict__ data, int n, typename Op::T* __restrict__ chunk_off,
                                                               unsigned* __restrict__ ticket, typename Op::T* __restrict__ total) {
  typedef typename Op::T T;
  __shared__ T s_w[SCAN_TPB / 32];
  __shared__ bool s_last;
  const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_IPT;
  T v[SCAN_IPT], sum = Op::identity();
#pragma unroll
  for (int j = 0; j < SCAN_IPT; ++j) {
    v[j] = (base + j < n) ? data[base + j] : Op::identity();
    sum = Op::apply(sum, v[j]);
  }
  T incl = warp_incl_scan<Op>(sum);
  if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
  __syncthreads();
  T run = Op::identity();
  for (int w = 0; w < (threadIdx.x >> 5); ++w) run = Op::apply(run, s_w[w]);
  T lane_excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if ((threadIdx.x & 31) != 0) run = Op::apply(run, lane_excl);
#pragma unroll
  for (int j = 0; j < SCAN_IPT; ++j) {
    if (base + j < n) data[base + j] = run;
    run = Op::apply(run, v[j]);
  }
  if (threadIdx.x == SCAN_TPB - 1) {
    chunk_off[blockIdx.x] = run;  // chunk total for now
    __threadfence();
    s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // the last CTA: exclusive scan of the gridDim.x chunk totals (serial over strips of SCAN_TPB)
  __shared__ T s_carry;
  if (threadIdx.x == 0) s_carry = Op::identity();
  __syncthreads();
  const int nb = gridDim.x;
  for (int b0 = 0; b0 < nb; b0 += SCAN_TPB) {
    const int i = b0 + threadIdx.x;
    T x = (i < nb) ? __ldcg(chunk_off + i) : Op::identity();
    T in2 = warp_incl_scan<Op>(x);
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = in2;
    __syncthreads();
    T excl = s_carry;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) excl = Op::apply(excl, s_w[w]);
    T le = __shfl_up_sync(0xffffffffu, in2, 1);
    if ((threadIdx.x & 31) != 0) excl = Op::apply(excl, le);
    if (i < nb) chunk_off[i] = excl;
    __syncthreads();
    if (threadIdx.x == SCAN_TPB - 1) s_carry = Op::apply(excl, x);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (total) *total = s_carry;
    *ticket = 0u;  // ready for the next scan that uses this ticket
  }
}
template <typename Op>
static int device_scan_chunks_impl(pitt_ctx* ctx, typename Op::T* d_data, int n, typename Op::T** d_chunk_off, unsigned* d_ticket,
                                   typename Op::T* d_total) {
  typedef typename Op::T T;
  const int nb = std::max(1, cdiv(n, SCAN_CHUNK));
  T* d_off = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)nb + 1, &d_off));
  *d_chunk_off = d_off;
  scan_chunks_kernel<Op><<<nb, SCAN_TPB, 0, ctx->stream>>>(d_data, n, d_off, d_ticket, d_total);
  ctx->launches++;
  PITT_CUDA(ctx, cudaGetLastError());
  return PITT_OK;
}
int device_scan_chunks(pitt_ctx* ctx, int* d_data, int n, int** d_chunk_off, unsigned* d_ticket, int* d_total) {
  return device_scan_chunks_impl<OpSumI>(ctx, d_data, n, d_chunk_off, d_ticket, d_total);
}
int device_max_scan_chunks(pitt_ctx* ctx, float* d_data, int n, float** d_chunk_off, unsigned* d_ticket) {
  return device_scan_chunks_impl<OpMaxF>(ctx, d_data, n, d_chunk_off, d_ticket, nullptr);
}

template <typename Op>
static int device_scan_impl(pitt_ctx* ctx, typename Op::T* d_data, int n, typename Op::T* d_total) {
  typedef typename Op::T T;
  if (n <= 0) return PITT_OK;
  const int nb = cdiv(n, SCAN_CHUNK);
  T* d_sums = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)nb + 1, &d_sums));
  scan_local_kernel<Op><<<nb, SCAN_TPB, 0, ctx->stream>>>(d_data, n, d_sums);
  scan_sums_kernel<Op><<<1, 1024, 0, ctx->stream>>>(d_sums, nb, d_total);
  if (nb > 1) {
    scan_add_kernel<Op><<<nb, SCAN_TPB, 0, ctx->stream>>>(d_data, n, d_sums);
    ctx->launches++;
  }
  ctx->launches += 2;
  PITT_CUDA(ctx, cudaGetLastError());
  return PITT_OK;
}

int device_exclusive_scan(pitt_ctx* ctx, int* d_data, int n, int* d_total) {
  if (n <= 0) {
    if (d_total) PITT_CUDA(ctx, cudaMemsetAsync(d_total, 0, sizeof(int), ctx->stream));
    return PITT_OK;
  }
  return device_scan_impl<OpSumI>(ctx, d_data, n, d_total);
}
// exclusive prefix maximum of n floats in place (first element becomes -inf)
int device_exclusive_max_scan(pitt_ctx* ctx, float* d_data, int n) { return device_scan_impl<OpMaxF>(ctx, d_data, n, nullptr); }

static void make_geom(const float mn[3], const float mx[3], float h, GridGeom* g, int* ncells) {
  // keep the dense table bounded: enlarge h until it fits 2^24 cells
  for (;;) {
    double d[3];
    double total = 1.0;
    for (int a = 0; a < 3; ++a) {
      d[a] = floor(((double)mx[a] - (double)mn[a]) / h) + 1.0;
      total *= d[a];
    }
    if (total <= (double)(1 << 24)) {
      g->dx = (int)d[0]; g->dy = (int)d[1]; g->dz = (int)d[2];
      break;
    }
    h *= 1.26f;
  }
  g->mnx = mn[0]; g->mny = mn[1]; g->mnz = mn[2];
  g->inv_h = 1.0f / h;
  *ncells = g->dx * g->dy * g->dz;
}

int cloud_bbox(pitt_ctx* ctx, const float4* d_xyz, int n, float mn[3], float mx[3], int* n_finite) {
  *n_finite = 0;
  if (n <= 0) return PITT_OK;
  int* d_bb = nullptr;
  PITT_TRY(arena_alloc(ctx, 8, &d_bb));
  bbox_init_kernel<<<1, 32, 0, ctx->stream>>>(d_bb);
  int nb = std::min(cdiv(n, 256), ctx->sm_count * 8);
  bbox_kernel<<<nb, 256, 0, ctx->stream>>>(d_xyz, n, d_bb);
  ctx->launches += 2;
  int h_bb[8];
  PITT_CUDA(ctx, cudaMemcpyAsync(h_bb, d_bb, 7 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  *n_finite = h_bb[6];
  for (int a = 0; a < 3; ++a) { mn[a] = ord2f_host(h_bb[a]); mx[a] = ord2f_host(h_bb[3 + a]); }
  return PITT_OK;
}

int grid_build(pitt_ctx* ctx, const float4* d_xyz, int n, float h, float target_per_cell, GridDev* out) {
  memset(out, 0, sizeof(*out));
  if (n <= 0) return PITT_OK;
  int n_finite = 0;
  float mn[3], mx[3];
  PITT_TRY(cloud_bbox(ctx, d_xyz, n, mn, mx, &n_finite));
  if (n_finite <= 0) return PITT_OK;
  const float ext = std::max(std::max(mx[0] - mn[0], mx[1] - mn[1]), std::max(mx[2] - mn[2], 1e-6f));
  GridGeom g;
  int ncells = 0;
  int* d_cnt = nullptr;
  if (h <= 0.0f) {
    // density probe: occupied cells M0 at a trial size h0 estimate the surface area A ~ M0*h0^2;
    // then n / (A / h^2) = target  =>  h = h0 * sqrt(target * M0 / n)
    float h0 = std::max(ext / 192.0f, ext * 1e-6f);
    make_geom(mn, mx, h0, &g, &ncells);
    h0 = 1.0f / g.inv_h;
    PITT_TRY(arena_alloc(ctx, (size_t)ncells + 2, &d_cnt));
    PITT_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, ((size_t)ncells + 2) * sizeof(int), ctx->stream));
    cell_hist_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_xyz, n, g, d_cnt);
    count_nonempty_kernel<<<std::min(cdiv(ncells, 256), ctx->sm_count * 8), 256, 0, ctx->stream>>>(d_cnt, ncells, d_cnt + ncells + 1);
    ctx->launches += 2;
    int m0 = 0;
    PITT_CUDA(ctx, cudaMemcpyAsync(&m0, d_cnt + ncells + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
    if (m0 < 1) m0 = 1;
    h = h0 * sqrtf(target_per_cell * (float)m0 / (float)n_finite);
    h = std::max(h, ext * 1e-5f);
  }
  make_geom(mn, mx, h, &g, &ncells);
  h = 1.0f / g.inv_h;
  int* d_start = nullptr;
  int* d_cursor = nullptr;
  float4* d_sorted = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)ncells + 1, &d_start));
  PITT_TRY(arena_alloc(ctx, (size_t)ncells + 1, &d_cursor));
  PITT_TRY(arena_alloc(ctx, (size_t)n_finite, &d_sorted));
  PITT_CUDA(ctx, cudaMemsetAsync(d_start, 0, ((size_t)ncells + 1) * sizeof(int), ctx->stream));
  cell_hist_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_xyz, n, g, d_start);
  ctx->launches++;
  PITT_TRY(device_exclusive_scan(ctx, d_start, ncells + 1, nullptr));
  PITT_CUDA(ctx, cudaMemcpyAsync(d_cursor, d_start, ((size_t)ncells + 1) * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
  cell_scatter_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_xyz, n, g, d_cursor, d_sorted);
  ctx->launches++;
  PITT_CUDA(ctx, cudaGetLastError());
  out->mnx = g.mnx; out->mny = g.mny; out->mnz = g.mnz;
  out->h = h; out->inv_h = g.inv_h;
  out->dx = g.dx; out->dy = g.dy; out->dz = g.dz;
  out->ncells = ncells;
  out->n = n_finite;
  out->cell_start = d_start;
  out->sorted = d_sorted;
  return PITT_OK;
}

}  // namespace pitt
