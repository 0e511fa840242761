// mgrid.cuh — the multi-level block-Morton grid of knn.cu, shared with the clusterer (cluster.cu): one dense table of fine
// cells, blocks of 8 x 8 x 8 cells ordered linearly and Morton-coded inside, geometry resident on the device (built without a
// host round trip). See knn.cu for the construction.
#pragma once
#include "pitt_common.cuh"

namespace pitt {

constexpr int MG_CAP_CELLS = 1 << 22;  // dense table: at most 4 M fine cells (16 MB of int)
constexpr int MG_CAP_BLOCKS = MG_CAP_CELLS >> 9;
constexpr int MG_MAXLVL = 3;

struct MGrid {
  float mnx, mny, mnz, hf, inv_hf;
  int dx, dy, dz;     // fine cells per axis (multiples of 8)
  int nbx, nby, nbz;  // blocks per axis
  int ncells, nblocks;
  int n_finite;
};

__device__ __forceinline__ int mg_spread3(int v) { return (v & 1) | ((v & 2) << 2) | ((v & 4) << 4); }
__device__ __forceinline__ int mg_index(const MGrid& g, int cx, int cy, int cz) {
  const int blk = ((cz >> 3) * g.nby + (cy >> 3)) * g.nbx + (cx >> 3);
  return (blk << 9) | mg_spread3(cx & 7) | (mg_spread3(cy & 7) << 1) | (mg_spread3(cz & 7) << 2);
}
__device__ __forceinline__ int mg_f2ord(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float mg_ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__device__ __forceinline__ bool mg_finite3(float4 p) { return isfinite(p.x) && isfinite(p.y) && isfinite(p.z); }

__device__ __forceinline__ int mg_coord(float v, float mn, float inv_h, int dim) {
  const int c = (int)floorf((v - mn) * inv_h);
  return min(max(c, 0), dim - 1);
}

struct MGridBuf {
  MGrid* d_G;
  int* d_start;    // [ncells + 1] first point of every fine cell in d_sorted
  float4* d_sorted;  // {x, y, z, original index as int bits}, grouped by cell
  int* d_scr;      // [0..5] bounding box, [6] finite points, [7] ring-search queue length, [8..] 64-bit diagnostics
  int* d_fb_list;
};
// Builds the grid of d_xyz[0..n) on ctx->stream; nothing waits for the device. h_fixed > 0: fine cells of exactly that size
// (enlarged only if the table would not fit) — the clusterer needs cells at least as wide as its tolerance; otherwise the cell
// size follows the mean surface density (c_avg points per fine cell).
// d_seg_off (nullable, with *d_n_seg on the device): segment offsets; points then carry their segment in the top 8 bits of the
// index word (n < 2^24).
int mgrid_build(pitt_ctx* ctx, const float4* d_xyz, int n, float h_fixed, MGridBuf* out, const int* d_seg_off = nullptr,
                const int* d_n_seg = nullptr);

}  // namespace pitt
