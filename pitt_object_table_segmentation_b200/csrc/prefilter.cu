// prefilter.cu — the pre-path of depthAcquisition on the device:
//   fromROSMsg               point_cloud_library/pc_manager.cpp:85-94    -> one strided H2D + pack
//   PCManager::downSampling  pc_manager.cpp:55-67 (pcl::VoxelGrid, leaf 0.01 :19)
//   deep filter              segmentation_services/deep_filter_srv.cpp:27-58 (threshold 3.0 :21)
//   transformPointCloud      obj_segmentation.cpp:248 (camera -> world Matrix4f)
// VoxelGrid = voxel keys -> stable radix sort of (key, point index) -> segment heads -> one thread per
// voxel accumulates its points in ascending point index (float, like PCL) -> centroids in ascending voxel
// index (PCL's output order). The deep filter and the transform are fused into one ordered compaction.
// The sort is the hand-written stable LSD radix sort below (rs_*: per-tile digit histogram -> one-launch scan -> ranked scatter).
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <limits>

#include "grid.cuh"
#include "pitt_common.cuh"

namespace pitt {

// fromROSMsg: records of `stride` bytes with x,y,z float32 at 0,4,8 -> float4 {x,y,z,1}
static __global__ void pf_pack_kernel(const unsigned char* __restrict__ src, int stride, int n, float4* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = reinterpret_cast<const float*>(src + (size_t)i * stride);
  dst[i] = make_float4(p[0], p[1], p[2], 1.0f);
}

struct VoxGeom {
  float inv0, inv1, inv2;
  int min0, min1, min2;
  int mul1, mul2;
  unsigned sentinel;  // key of a non-finite point: one past the largest voxel index
};
__global__ void __launch_bounds__(256) voxel_key_kernel(const float4* __restrict__ xyz, int n, VoxGeom g, unsigned* __restrict__ key,
                                                        int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(xyz + i);
  unsigned k = g.sentinel;  // non-finite points sort to the end and are cut off
  if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
    // static_cast<int> (floor (x * inverse_leaf_size_[0]) - static_cast<float> (min_b_[0]))
    const int i0 = (int)(floorf(p.x * g.inv0) - (float)g.min0);
    const int i1 = (int)(floorf(p.y * g.inv1) - (float)g.min1);
    const int i2 = (int)(floorf(p.z * g.inv2) - (float)g.min2);
    k = (unsigned)(i0 + i1 * g.mul1 + i2 * g.mul2);
  }
  key[i] = k;
  idx[i] = i;
}
// ---------------------------------------------------------------- stable LSD radix sort of (key, value) pairs
// A pass sorts by `bits` (<= 9) key bits: tiles of 2048 pairs, warp w of a tile owns the contiguous pairs [256 w, 256 w + 256)
// and walks them in 8 rounds of 32, so "warp, round, lane" is ascending input position and equal digits keep their order.
//   rs_hist_kernel     digit histogram of every tile -> hist[digit][tile]
//   device_scan_chunks exclusive scan of hist in that (digit-major) order = where each tile's run of a digit starts
//   rs_scatter_kernel  rank of a pair among the tile's equal digits (match_any inside the warp + running per-warp counters in
//                      shared memory + a prefix over the 8 warps) -> output position
constexpr int RS_TPB = 256, RS_TILE = 2048, RS_MAXB = 512;
__global__ void __launch_bounds__(RS_TPB) rs_hist_kernel(const unsigned* __restrict__ key, int n, int shift, int bits, int ntiles,
                                                         int* __restrict__ hist) {
  __shared__ int cnt[RS_MAXB];
  const int NB = 1 << bits;
  for (int d = threadIdx.x; d < NB; d += RS_TPB) cnt[d] = 0;
  __syncthreads();
  const int base = blockIdx.x * RS_TILE;
#pragma unroll
  for (int r = 0; r < RS_TILE / RS_TPB; ++r) {
    const int i = base + r * RS_TPB + threadIdx.x;
    if (i < n) atomicAdd(&cnt[(__ldg(key + i) >> shift) & (NB - 1)], 1);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < NB; d += RS_TPB) hist[(size_t)d * ntiles + blockIdx.x] = cnt[d];
}
__global__ void __launch_bounds__(RS_TPB) rs_scatter_kernel(const unsigned* __restrict__ key_in, const int* __restrict__ val_in, int n,
                                                            int shift, int bits, int ntiles, const int* __restrict__ off,
                                                            const int* __restrict__ chunk_off, unsigned* __restrict__ key_out,
                                                            int* __restrict__ val_out) {
  __shared__ int cnt[8][RS_MAXB];  // running count of digit d in warp w, then the warp's start inside the tile's run of d
  __shared__ int goff[RS_MAXB];    // start of the tile's run of digit d in the output
  const int NB = 1 << bits;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8 * RS_MAXB; i += RS_TPB) (&cnt[0][0])[i] = 0;
  for (int d = threadIdx.x; d < NB; d += RS_TPB) {
    const size_t j = (size_t)d * ntiles + blockIdx.x;
    goff[d] = off[j] + chunk_off[j >> SCAN_CHUNK_LOG2];
  }
  __syncthreads();
  const int base = blockIdx.x * RS_TILE + w * 256;
  unsigned k[8];
  int v[8], rank[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = base + r * 32 + lane;
    const bool valid = i < n;
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    rank[r] = 0;
    if (valid) {
      k[r] = __ldg(key_in + i);
      v[r] = __ldg(val_in + i);
      const int d = (k[r] >> shift) & (NB - 1);
      const unsigned m = __match_any_sync(act, d);
      const int before = __popc(m & ((1u << lane) - 1u));
      const int prior = cnt[w][d];
      __syncwarp(act);
      if (before == 0) cnt[w][d] = prior + __popc(m);
      __syncwarp(act);
      rank[r] = prior + before;
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < NB; d += RS_TPB) {
    int run = 0;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) {
      const int c = cnt[ww][d];
      cnt[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = base + r * 32 + lane;
    if (i < n) {
      const int d = (k[r] >> shift) & (NB - 1);
      const int pos = goff[d] + cnt[w][d] + rank[r];
      key_out[pos] = k[r];
      val_out[pos] = v[r];
    }
  }
}
// sorts by key bits [0, end_bit); the result is in (*key_a, *val_a) on return (the pointers are swapped per pass)
static int radix_sort_pairs(pitt_ctx* ctx, unsigned** key_a, int** val_a, unsigned** key_b, int** val_b, int n, int end_bit) {
  const int passes = (end_bit + 8) / 9;  // <= 9 bits per pass
  const int ntiles = cdiv(n, RS_TILE);
  int* d_hist = nullptr;
  unsigned* d_ticket = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)RS_MAXB * ntiles, &d_hist));
  PITT_TRY(arena_alloc(ctx, 1, &d_ticket));
  PITT_CUDA(ctx, cudaMemsetAsync(d_ticket, 0, sizeof(unsigned), ctx->stream));
  int shift = 0;
  for (int p = 0; p < passes; ++p) {
    const int bits = (end_bit - shift + (passes - p) - 1) / (passes - p);
    rs_hist_kernel<<<ntiles, RS_TPB, 0, ctx->stream>>>(*key_a, n, shift, bits, ntiles, d_hist);
    int* d_chunk_off = nullptr;
    PITT_TRY(device_scan_chunks(ctx, d_hist, (1 << bits) * ntiles, &d_chunk_off, d_ticket, nullptr));
    rs_scatter_kernel<<<ntiles, RS_TPB, 0, ctx->stream>>>(*key_a, *val_a, n, shift, bits, ntiles, d_hist, d_chunk_off, *key_b, *val_b);
    ctx->launches += 2;
    std::swap(*key_a, *key_b);
    std::swap(*val_a, *val_b);
    shift += bits;
  }
  return PITT_OK;
}

__global__ void __launch_bounds__(256) voxel_heads_kernel(const unsigned* __restrict__ key_sorted, int m, int* __restrict__ head) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  head[i] = (i == 0 || key_sorted[i] != key_sorted[i - 1]) ? 1 : 0;
}
// rank[i] = exclusive scan of head = voxel number of sorted entry i; the thread of a segment head walks its segment
__global__ void __launch_bounds__(128) voxel_centroid_kernel(const float4* __restrict__ xyz, const unsigned* __restrict__ key_sorted,
                                                             const int* __restrict__ idx_sorted, const int* __restrict__ rank, int m,
                                                             float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const unsigned k = key_sorted[i];
  if (i > 0 && key_sorted[i - 1] == k) return;  // not a head
  float4 p = __ldg(xyz + idx_sorted[i]);
  float cx = p.x, cy = p.y, cz = p.z;
  int j = i + 1;
  for (; j < m && key_sorted[j] == k; ++j) {
    p = __ldg(xyz + idx_sorted[j]);
    cx += p.x; cy += p.y; cz += p.z;
  }
  const float cnt = (float)(j - i);
  out[rank[i]] = make_float4(cx / cnt, cy / cnt, cz / cnt, 1.0f);
}
// deep filter: 1 = kept ("closer"), further counted on the side
__global__ void __launch_bounds__(256) deep_flag_kernel(const float4* __restrict__ xyz, int n, float th, int apply, int* __restrict__ keep,
                                                        int* __restrict__ n_further) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int far_ = 0;
  if (i < n) {
    const float z = xyz[i].z;
    int k = 1;
    if (apply) {
      k = 0;
      if (z == z) {
        if (z > th) far_ = 1;
        else k = 1;
      }
    }
    keep[i] = k;
  }
  far_ = __reduce_add_sync(0xffffffffu, far_);
  if ((threadIdx.x & 31) == 0 && far_) atomicAdd(n_further, far_);
}
struct Xform {
  float m[12];
  int apply;
};
__global__ void __launch_bounds__(256) compact_transform_kernel(const float4* __restrict__ src, const int* __restrict__ keep_flag,
                                                                const int* __restrict__ pos, int n, Xform T, float4* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep_flag[i]) return;
  float4 p = src[i];
  if (T.apply && isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
    const float x = p.x, y = p.y, z = p.z;  // -fmad=false: every product and sum rounds once, left to right
    p.x = T.m[0] * x + T.m[1] * y + T.m[2] * z + T.m[3];
    p.y = T.m[4] * x + T.m[5] * y + T.m[6] * z + T.m[7];
    p.z = T.m[8] * x + T.m[9] * y + T.m[10] * z + T.m[11];
  }
  p.w = 1.0f;
  dst[pos[i]] = p;
}

// ---- arm filter: pcl::CropBox (negative) per box, chained (arm_filter_srv.cpp:66-103, 134-141)
struct CropDev {
  float inv[9];  // inverse of the rotation (row major); identity when apply_rot == 0
  float t[3];
  float mn[3], mx[3];
  int apply_rot, apply_t;
};
struct CropSet {
  CropDev b[4];
  int n, dense;
};
// which box removes the point: 0 = none (kept), k + 1 = box k (the first one in the chain that contains it), 5 = dropped as
// non-finite by the first CropBox
__global__ void __launch_bounds__(256) crop_flag_kernel(const float4* __restrict__ xyz, int n, CropSet S, int* __restrict__ keep,
                                                        int* __restrict__ removed /*[5]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int who = 0;
  if (i < n) {
    const float4 p = xyz[i];
    if (!S.dense && !(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) {
      who = 5;
    } else {
      for (int k = 0; k < S.n && who == 0; ++k) {
        const CropDev& B = S.b[k];
        float x = p.x, y = p.y, z = p.z;
        if (B.apply_t) { x -= B.t[0]; y -= B.t[1]; z -= B.t[2]; }
        if (B.apply_rot) {  // Eigen 3x3 * vector, coefficient based: (m0 x + m1 y) + m2 z, unfused
          const float lx = (B.inv[0] * x + B.inv[1] * y) + B.inv[2] * z;
          const float ly = (B.inv[3] * x + B.inv[4] * y) + B.inv[5] * z;
          const float lz = (B.inv[6] * x + B.inv[7] * y) + B.inv[8] * z;
          x = lx; y = ly; z = lz;
        }
        const bool outside = (x < B.mn[0] || y < B.mn[1] || z < B.mn[2]) || (x > B.mx[0] || y > B.mx[1] || z > B.mx[2]);
        if (!outside) who = k + 1;  // NaN compares false everywhere: "inside", removed (negative filter)
      }
    }
    keep[i] = who == 0 ? 1 : 0;
  }
#pragma unroll
  for (int k = 1; k <= 5; ++k) {
    const unsigned m = __ballot_sync(0xffffffffu, who == k);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&removed[k - 1], __popc(m));
  }
}
// pcl::getTransformation(0, 0, 0, roll, pitch, yaw) followed by Eigen's Affine3f::inverse() (3x3 cofactor inverse); float,
// cos/sin = the double function rounded once (the convention of pitt_math.cuh for PCL's float transcendentals)
static void crop_inverse_rotation(const float rpy[3], float inv[9]) {
  const float A = (float)cos((double)rpy[2]), B = (float)sin((double)rpy[2]), C = (float)cos((double)rpy[1]),
              D = (float)sin((double)rpy[1]), E = (float)cos((double)rpy[0]), F = (float)sin((double)rpy[0]);
  const float DE = D * E, DF = D * F;
  float m[3][3];
  m[0][0] = A * C;  m[0][1] = A * DF - B * E;  m[0][2] = B * F + A * DE;
  m[1][0] = B * C;  m[1][1] = A * E + B * DF;  m[1][2] = B * DE - A * F;
  m[2][0] = -D;     m[2][1] = C * F;           m[2][2] = C * E;
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
  };
  const float c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
  const float det = (c0 * m[0][0] + c1 * m[1][0]) + c2 * m[2][0];
  const float invdet = 1.0f / det;
  inv[0] = c0 * invdet; inv[1] = c1 * invdet; inv[2] = c2 * invdet;
  for (int i = 1; i < 3; ++i)
    for (int j = 0; j < 3; ++j) inv[3 * i + j] = cof(j, i) * invdet;
}

// d_in: n float4 on the device. Writes a new pool-allocated cloud.
static int prefilter_impl(pitt_ctx* ctx, const float4* d_in, int n, const pitt_prefilter_params& P, pitt_cloud** out,
                          pitt_prefilter_info* info) {
  pitt_prefilter_info I;
  memset(&I, 0, sizeof(I));
  I.n_input = n;
  I.used_deep_threshold = P.deep_threshold >= 0.0f ? P.deep_threshold : 3.000f;
  const float4* d_cur = d_in;
  int n_cur = n;
  // ---------------- VoxelGrid
  if (n > 0 && P.leaf[0] > 0.0f && P.leaf[1] > 0.0f && P.leaf[2] > 0.0f) {
    float mn[3], mx[3];
    int n_finite = 0;
    PITT_TRY(cloud_bbox(ctx, d_in, n, mn, mx, &n_finite));
    if (n_finite == 0) {
      n_cur = 0;
    } else {
      float inv[3];
      for (int a = 0; a < 3; ++a) inv[a] = 1.0f / P.leaf[a];
      const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv[0]) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv[1]) + 1,
                    dz = (int64_t)((mx[2] - mn[2]) * inv[2]) + 1;
      if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) {
        I.voxel_overflow = 1;  // PCL: "Leaf size is too small for the input dataset": output = input
      } else {
        VoxGeom g;
        int min_b[3], max_b[3], div_b[3];
        for (int a = 0; a < 3; ++a) {
          min_b[a] = (int)floorf(mn[a] * inv[a]);
          max_b[a] = (int)floorf(mx[a] * inv[a]);
          div_b[a] = max_b[a] - min_b[a] + 1;
        }
        g.inv0 = inv[0]; g.inv1 = inv[1]; g.inv2 = inv[2];
        g.min0 = min_b[0]; g.min1 = min_b[1]; g.min2 = min_b[2];
        g.mul1 = div_b[0]; g.mul2 = div_b[0] * div_b[1];
        unsigned *d_key = nullptr, *d_key2 = nullptr;
        int *d_idx = nullptr, *d_idx2 = nullptr, *d_head = nullptr, *d_total = nullptr;
        float4* d_vox = nullptr;
        PITT_TRY(arena_alloc(ctx, (size_t)n, &d_key));
        PITT_TRY(arena_alloc(ctx, (size_t)n, &d_key2));
        PITT_TRY(arena_alloc(ctx, (size_t)n, &d_idx));
        PITT_TRY(arena_alloc(ctx, (size_t)n, &d_idx2));
        PITT_TRY(arena_alloc(ctx, (size_t)n, &d_head));
        PITT_TRY(arena_alloc(ctx, 1, &d_total));
        PITT_TRY(arena_alloc(ctx, (size_t)n_finite, &d_vox));
        // stable sort on the voxel key only: equal keys keep ascending point index; only the bits a key can have are sorted
        const int64_t max_key = (int64_t)div_b[0] * div_b[1] * div_b[2];  // <= INT32_MAX (checked above)
        g.sentinel = (unsigned)max_key;
        int end_bit = 1;
        while (end_bit < 32 && ((int64_t)1 << end_bit) <= max_key) ++end_bit;
        voxel_key_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_in, n, g, d_key, d_idx);
        ctx->launches++;
        PITT_TRY(radix_sort_pairs(ctx, &d_key, &d_idx, &d_key2, &d_idx2, n, end_bit));
        std::swap(d_key, d_key2);  // the code below reads the sorted pairs from d_key2 / d_idx2
        std::swap(d_idx, d_idx2);
        const int m = n_finite;  // the finite points come first
        voxel_heads_kernel<<<cdiv(m, 256), 256, 0, ctx->stream>>>(d_key2, m, d_head);
        ctx->launches++;
        PITT_TRY(device_exclusive_scan(ctx, d_head, m, d_total));
        voxel_centroid_kernel<<<cdiv(m, 128), 128, 0, ctx->stream>>>(d_in, d_key2, d_idx2, d_head, m, d_vox);
        ctx->launches++;
        int n_vox = 0;
        PITT_CUDA(ctx, cudaMemcpyAsync(&n_vox, d_total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        PITT_CUDA(ctx, pitt::stream_sync(ctx));
        d_cur = d_vox;
        n_cur = n_vox;
      }
    }
  }
  I.n_voxel = n_cur;
  // ---------------- deep filter + transform, fused into one ordered compaction
  pitt_cloud* c = new pitt_cloud();
  int n_out = n_cur;
  if (n_cur > 0) {
    int *d_keep = nullptr, *d_pos = nullptr, *d_cnt = nullptr;
    PITT_TRY(arena_alloc(ctx, (size_t)n_cur, &d_keep));
    PITT_TRY(arena_alloc(ctx, (size_t)n_cur, &d_pos));
    PITT_TRY(arena_alloc(ctx, 2, &d_cnt));
    PITT_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int), ctx->stream));
    deep_flag_kernel<<<cdiv(n_cur, 256), 256, 0, ctx->stream>>>(d_cur, n_cur, I.used_deep_threshold, P.apply_deep_filter ? 1 : 0, d_keep, d_cnt + 1);
    ctx->launches++;
    PITT_CUDA(ctx, cudaMemcpyAsync(d_pos, d_keep, (size_t)n_cur * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    PITT_TRY(device_exclusive_scan(ctx, d_pos, n_cur, d_cnt));
    int h_cnt[2] = {0, 0};
    PITT_CUDA(ctx, cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
    n_out = h_cnt[0];
    I.n_further = h_cnt[1];
    if (n_out > 0) {
      int st = pool_alloc(ctx, (size_t)n_out * sizeof(float4), (void**)&c->d_xyz);
      if (st != PITT_OK) { delete c; return st; }
      Xform T;
      for (int k = 0; k < 12; ++k) T.m[k] = P.transform[k];
      T.apply = P.apply_transform ? 1 : 0;
      compact_transform_kernel<<<cdiv(n_cur, 256), 256, 0, ctx->stream>>>(d_cur, d_keep, d_pos, n_cur, T, c->d_xyz);
      ctx->launches++;
    }
  }
  c->n = n_out;
  I.n_closer = n_out;
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { pool_free(ctx, c->d_xyz, (size_t)c->n * sizeof(float4)); delete c; return fail(ctx, PITT_ERR_CUDA, "prefilter kernels", e); }
  *out = c;
  if (info) *info = I;
  return PITT_OK;
}

}  // namespace pitt

using namespace pitt;

extern "C" {

void pitt_default_prefilter_params(pitt_prefilter_params* p) {
  memset(p, 0, sizeof(*p));
  p->leaf[0] = p->leaf[1] = p->leaf[2] = 0.01f;  // PCManager::DEFAULT_DOWSEAMPLIG_RATE, pc_manager.cpp:19
  p->apply_deep_filter = 1;
  p->deep_threshold = -1.0f;                      // -> 3.0, deep_filter_srv.cpp:21
  p->apply_transform = 1;
  p->transform[0] = p->transform[5] = p->transform[10] = p->transform[15] = 1.0f;
}

int pitt_prefilter_cloud(pitt_ctx* ctx, const void* data, int point_step, int n_points, const pitt_prefilter_params* params,
                         pitt_cloud** out, pitt_prefilter_info* info) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!params || !out || n_points < 0 || (n_points > 0 && !data) || point_step < 12 || (point_step & 3))
    return fail(ctx, PITT_ERR_INVALID, "pitt_prefilter_cloud arguments (point_step >= 12, multiple of 4)");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  float4* d_in = nullptr;
  if (n_points > 0) {
    PITT_TRY(arena_alloc(ctx, (size_t)n_points, &d_in));
    if (point_step == 16) {
      PITT_CUDA(ctx, cudaMemcpyAsync(d_in, data, (size_t)n_points * 16, cudaMemcpyHostToDevice, ctx->stream));
    } else {
      unsigned char* d_raw = nullptr;
      PITT_TRY(arena_alloc(ctx, (size_t)n_points * point_step, &d_raw));
      PITT_CUDA(ctx, cudaMemcpyAsync(d_raw, data, (size_t)n_points * point_step, cudaMemcpyHostToDevice, ctx->stream));
      pf_pack_kernel<<<cdiv(n_points, 256), 256, 0, ctx->stream>>>(d_raw, point_step, n_points, d_in);
      ctx->launches++;
    }
  }
  int st = prefilter_impl(ctx, d_in, n_points, *params, out, info);
  timer.finish();
  return st;
}

int pitt_prefilter_staged(pitt_ctx* ctx, const pitt_cloud* in, const pitt_prefilter_params* params, pitt_cloud** out,
                          pitt_prefilter_info* info) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!in || !params || !out) return fail(ctx, PITT_ERR_INVALID, "pitt_prefilter_staged arguments");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  int st = prefilter_impl(ctx, in->d_xyz, in->n, *params, out, info);
  timer.finish();
  return st;
}

void pitt_default_arm_filter_params(pitt_arm_filter_params* p) {
  memset(p, 0, sizeof(*p));
  p->n_boxes = 4;
  const float mn_f[3] = {-0.040f, -0.120f, -0.190f}, mx_f[3] = {0.340f, 0.120f, 0.105f};   // arm_filter_srv.cpp:32-33
  const float mn_e[3] = {-0.090f, -0.135f, -0.160f}, mx_e[3] = {0.440f, 0.135f, 0.110f};   // :34-35
  for (int k = 0; k < 4; ++k)
    for (int a = 0; a < 3; ++a) {
      p->box[k].min_pt[a] = k < 2 ? mn_f[a] : mn_e[a];  // left forearm, right forearm, left elbow, right elbow (:134-141)
      p->box[k].max_pt[a] = k < 2 ? mx_f[a] : mx_e[a];
    }
}

int pitt_arm_filter(pitt_ctx* ctx, const pitt_cloud* in, const pitt_arm_filter_params* P, pitt_cloud** out, int32_t* removed) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!in || !P || !out || P->n_boxes < 0 || P->n_boxes > 4) return fail(ctx, PITT_ERR_INVALID, "pitt_arm_filter arguments");
  cudaSetDevice(ctx->device);
  CallTimer timer(ctx);
  CropSet S;
  memset(&S, 0, sizeof(S));
  S.n = P->n_boxes;
  S.dense = P->input_is_dense ? 1 : 0;
  for (int k = 0; k < S.n; ++k) {
    const pitt_crop_box& b = P->box[k];
    CropDev& d = S.b[k];
    d.apply_rot = (b.rotation_rpy[0] != 0.0f || b.rotation_rpy[1] != 0.0f || b.rotation_rpy[2] != 0.0f) ? 1 : 0;
    d.apply_t = (b.translation[0] != 0.0f || b.translation[1] != 0.0f || b.translation[2] != 0.0f) ? 1 : 0;
    if (d.apply_rot) crop_inverse_rotation(b.rotation_rpy, d.inv);
    for (int a = 0; a < 3; ++a) { d.t[a] = b.translation[a]; d.mn[a] = b.min_pt[a]; d.mx[a] = b.max_pt[a]; }
  }
  pitt_cloud* c = new pitt_cloud();
  const int n = in->n;
  int h_rem[5] = {0, 0, 0, 0, 0};
  int n_out = 0;
  if (n > 0 && S.n == 0) {  // tfError branch: no operation performed
    n_out = n;
    if (pool_alloc(ctx, (size_t)n * sizeof(float4), (void**)&c->d_xyz) != PITT_OK) { delete c; return PITT_ERR_CUDA; }
    PITT_CUDA(ctx, cudaMemcpyAsync(c->d_xyz, in->d_xyz, (size_t)n * 16, cudaMemcpyDeviceToDevice, ctx->stream));
  } else if (n > 0) {
    int *d_keep = nullptr, *d_pos = nullptr, *d_cnt = nullptr;
    PITT_TRY(arena_alloc(ctx, (size_t)n, &d_keep));
    PITT_TRY(arena_alloc(ctx, (size_t)n, &d_pos));
    PITT_TRY(arena_alloc(ctx, 8, &d_cnt));
    PITT_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, 8 * sizeof(int), ctx->stream));
    crop_flag_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(in->d_xyz, n, S, d_keep, d_cnt + 1);
    ctx->launches++;
    PITT_CUDA(ctx, cudaMemcpyAsync(d_pos, d_keep, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    PITT_TRY(device_exclusive_scan(ctx, d_pos, n, d_cnt));
    int h_cnt[6];
    PITT_CUDA(ctx, cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
    n_out = h_cnt[0];
    for (int k = 0; k < 5; ++k) h_rem[k] = h_cnt[1 + k];
    if (n_out > 0) {
      if (pool_alloc(ctx, (size_t)n_out * sizeof(float4), (void**)&c->d_xyz) != PITT_OK) { delete c; return PITT_ERR_CUDA; }
      Xform T;
      memset(&T, 0, sizeof(T));
      compact_transform_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(in->d_xyz, d_keep, d_pos, n, T, c->d_xyz);
      ctx->launches++;
    }
  }
  c->n = n_out;
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { pool_free(ctx, c->d_xyz, (size_t)c->n * sizeof(float4)); delete c; return fail(ctx, PITT_ERR_CUDA, "arm filter kernels", e); }
  timer.finish();
  if (removed) {
    for (int k = 0; k < 4; ++k) removed[k] = h_rem[k];
    removed[0] += h_rem[4];  // non-finite points leave with the first CropBox of the chain
  }
  *out = c;
  return PITT_OK;
}

int pitt_get_points(pitt_ctx* ctx, const pitt_cloud* c, float* out4) {
  if (!ctx) return PITT_ERR_CUDA;
  if (!c || (c->n > 0 && !out4)) return fail(ctx, PITT_ERR_INVALID, "pitt_get_points arguments");
  cudaSetDevice(ctx->device);
  if (c->n > 0) {
    PITT_CUDA(ctx, cudaMemcpyAsync(out4, c->d_xyz, (size_t)c->n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
  }
  return PITT_OK;
}

}  // extern "C"
