// plane_tc.cu — K3t: plane hypothesis scoring with the dot products on the 5th-generation tensor
// cores (tcgen05.mma kind::f16 on BF16 pieces, FP32 accumulators in TMEM) and an exact re-evaluation
// of every evaluation the tensor result cannot decide.
//
// Replaces the point loop of pcl::SampleConsensusModelPlane::countWithinDistance reached from
// seg.segment() at supports_segmentation_srv.cpp:110 and plane_segmentation_srv.cpp:67
// (SURVEY.md B.3): count_h = #{ i : |fl(a x_i + b y_i + c z_i + d)| < thr }.
//
// The exact predicate needs 6 separately rounded FP32 operations per evaluation. Here
//   (1) every float is split without error into three BF16 pieces v = v1 + v2 + v3 (8 + 8 + 8
//       significand bits; each piece has its low 16 bits clear, so it IS a BF16 number),
//   (2) two chained 128 x 256 x 16 MMAs sum, per coordinate, the 8 products a_i x_j with
//       (i, j) != (3, 3), and d1 + d2 + d3 (27 of the 32 K slots); every product is exact (8 x 8 bits)
//       and what is dropped, a3 x3, is below 2^-27 m, m = |a x| + |b y| + |c z| + |d|; the only real
//       error is the tensor core's FP32 accumulation, bounded by TC_ACC_ULPS u m (u = 2^-24;
//       measured, profiles/r02_plane_tc_numerics.md),
//   (3) hypotheses are pre-scaled by a power of two sigma, so the accumulator holds s~ = sigma s and
//       u = sat(C - |s~|), C = fl(sigma thr + 1/2), is exactly 1 for a certain inlier, exactly 0 for a
//       certain outlier and fractional inside the window | |s~| - sigma thr | < 1/2, which sigma makes
//       at least 1/0.36 times wider than the error bound of s~,
//   (4) per (hypothesis, 128-point segment) the epilogue accumulates S1 = sum u and S2 = sum u^2 (one
//       FADD.SAT per evaluation, one packed FADD2 and one packed FFMA2 per two evaluations, no integer
//       work). S1 - S2 <= 0.1 proves that every fractional u is within 0.115 of 0 or 1, i.e. decided
//       with margin, and that rint(S1) is the exact count; otherwise the segment is re-evaluated by
//       the whole warp in the exact operation order.
// Result: counts bit-identical to plane_score_kernel (tests/test_gpu_plane_tc.py).
//
// Thread = hypothesis (TMEM lane), columns = points: no cross-lane reduction anywhere. A 512-point
// chunk is stationary in shared memory (split once per chunk by the epilogue warps), hypothesis
// block images stream through a 4-stage cp.async.bulk pipeline. TMEM holds two 256-column
// accumulators = the two halves ("phases") of the chunk; each of the 16 epilogue warps takes 64
// columns of accumulator 0, hands it back, does the arithmetic, then the same with accumulator 1:
// an accumulator is refilled by the tensor pipe while the FMA pipe works on the other one, so the
// two pipes overlap even though all warps run in lockstep. Counts for up to 2560 hypotheses are kept
// in shared memory and flushed once per super-block.
#include <math_constants.h>

#include "sac.cuh"

namespace pitt {

constexpr int TC_M = 128;                       // hypotheses per block (UMMA M, TMEM lanes)
#ifndef TC_E24
#define TC_E24 0  // 1: 24 epilogue warps (six per sub-partition) on two 192-column accumulators, 32 columns per warp and phase
#endif            // 0: 16 epilogue warps on two 256-column accumulators, 64 columns per warp and phase
constexpr int TC_N = TC_E24 ? 192 : 256;        // points per MMA (UMMA N) = columns of one accumulator
constexpr int TC_PHASES = 2;                    // accumulators in TMEM = tiles per point chunk
constexpr int TC_CHUNK = TC_N * TC_PHASES;      // 512 points stationary in shared memory
#ifndef TC_E8
#define TC_E8 0   // 1: 8 epilogue warps (two per sub-partition), 128 columns per warp and phase in two 64-column steps
#endif
constexpr int TC_RUN = TC_E24 ? 32 : (TC_E8 ? 128 : 64);  // columns one epilogue warp takes from an accumulator
constexpr int TC_MMAS = 2;                      // chained MMAs per tile (K = 16 BF16 each)
constexpr int TC_A_MMA_BYTES = TC_M * 32;       // 4096
constexpr int TC_A_BLOCK_BYTES = TC_MMAS * TC_A_MMA_BYTES;  // 8192 per hypothesis block
constexpr int TC_B_MMA_BYTES = TC_N * 32;       // 8192
constexpr int TC_B_TILE_BYTES = TC_MMAS * TC_B_MMA_BYTES;   // 16384
constexpr int TC_ASTAGES = 4;
constexpr int TC_SB = 40;                       // hypothesis blocks per super-block (counts in smem): 5120 hypotheses
constexpr int TC_EPI_WARPS = 4 * (TC_N / TC_RUN), TC_EPI_THREADS = 32 * TC_EPI_WARPS;
static_assert(TC_CHUNK <= TC_EPI_THREADS || TC_CHUNK % TC_EPI_THREADS == 0, "whole passes of the epilogue threads over a chunk");
constexpr int TC_MMA_WARPS = TC_PHASES;         // one MMA issuing warp per accumulator; warp 0 also streams the images
constexpr int TC_THREADS = TC_EPI_THREADS + 32 * TC_MMA_WARPS;
constexpr float TC_ACC_ULPS = 8.0f;             // bound on the tensor core accumulation error, in u m
constexpr float TC_WINDOW = 0.36f;              // sigma * beta_t must stay below this (see header)

// shared memory carve-up (bytes)
constexpr int TC_OFF_B = 0;
constexpr int TC_OFF_A = TC_OFF_B + TC_PHASES * TC_B_TILE_BYTES;      // 32768
constexpr int TC_OFF_RAW = TC_OFF_A + TC_ASTAGES * TC_A_BLOCK_BYTES;  // 65536
constexpr int TC_OFF_CNT = TC_OFF_RAW + TC_CHUNK * 16;                // 73728
constexpr int TC_OFF_BAR = TC_OFF_CNT + TC_SB * TC_M * 4;             // 94208
constexpr int TC_SMEM_BYTES = TC_OFF_BAR + 16 * 8 + 16;
// barriers
constexpr int TC_BAR_AFULL = 0, TC_BAR_AEMPTY = 4, TC_BAR_FULL = 8, TC_BAR_EMPTY = 10, TC_BAR_B = 12;

struct PlaneTcParams {
  float sigma;   // power of two
  float C;       // fl(sigma * thr_up + 0.5)
  float G;       // bound on m the scale was derived for
  float d_unc;   // scaled d of a hypothesis that must always be re-evaluated (u = 0.5 everywhere)
  float bx, by, bz;
  int use;       // 0: cloud not finite or scale out of range -> exact kernel
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(2000u) /* suspend-time hint (ns): sleep in hardware, wake on completion */
        : "memory");
  } while (!ok);
}
// one lane of the (converged) warp; the same lane every time for a full mask
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, no swizzle: 8-row x 16-byte core matrices (8 BF16 of K each); the two K halves of a row group are
// LBO = 128 B apart, consecutive 8-row groups SBO = 256 B apart (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor: D = F32, A = B = BF16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

#define TC_R32(r) \
  "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), \
  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),   \
  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),  \
  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
#define TC_RW32(r) \
  "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), \
  "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),   \
  "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),  \
  "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
// 32 lanes x 32 consecutive columns: lane i of the warp receives row (lane base + i), columns [col, col + 32)
__device__ __forceinline__ void tc_ld32(uint32_t (&r)[32], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : TC_R32(r)
      : "r"(taddr));
}
// wait for the outstanding tcgen05.ld and make the registers "change" here, so no consumer is scheduled earlier
__device__ __forceinline__ void tc_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : TC_RW32(r)::"memory");
}
__device__ __forceinline__ void tc_ld_wait2(uint32_t (&a)[32], uint32_t (&b)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : TC_RW32(a), TC_RW32(b)::"memory");
}

// exact BF16 pieces: v = v1 + v2 + v3, each with the low 16 bits of its FP32 pattern clear
__device__ __forceinline__ void tc_split3(float v, float& v1, float& v2, float& v3) {
  v1 = __uint_as_float(__float_as_uint(v) & 0xFFFF0000u);
  const float r = v - v1;  // exact
  v2 = __uint_as_float(__float_as_uint(r) & 0xFFFF0000u);
  v3 = r - v2;             // exact, at most 8 significant bits
}
// two BF16-exact floats -> one word (lo in bits 0..15)
__device__ __forceinline__ uint32_t tc_bf2(float lo, float hi) {
  return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
}
__device__ __forceinline__ uint4 tc_bf8(float f0, float f1, float f2, float f3, float f4, float f5, float f6, float f7) {
  return make_uint4(tc_bf2(f0, f1), tc_bf2(f2, f3), tc_bf2(f4, f5), tc_bf2(f6, f7));
}

// ------------------------------------------------------------------------------------------------
// set-up kernels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tc_absmax_kernel(const float4* __restrict__ xyz, int n, unsigned* __restrict__ out3) {
  unsigned mx = 0, my = 0, mz = 0;
  const int stride = gridDim.x * blockDim.x;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  // four independent loads in flight per thread: the kernel is a pure HBM/L2 read of the cloud
  for (; i + 3 * stride < n; i += 4 * stride) {
    const float4 p0 = __ldg(xyz + i), p1 = __ldg(xyz + i + stride), p2 = __ldg(xyz + i + 2 * stride), p3 = __ldg(xyz + i + 3 * stride);
    mx = max(max(mx, __float_as_uint(p0.x) & 0x7fffffffu), max(__float_as_uint(p1.x) & 0x7fffffffu,
             max(__float_as_uint(p2.x) & 0x7fffffffu, __float_as_uint(p3.x) & 0x7fffffffu)));
    my = max(max(my, __float_as_uint(p0.y) & 0x7fffffffu), max(__float_as_uint(p1.y) & 0x7fffffffu,
             max(__float_as_uint(p2.y) & 0x7fffffffu, __float_as_uint(p3.y) & 0x7fffffffu)));
    mz = max(max(mz, __float_as_uint(p0.z) & 0x7fffffffu), max(__float_as_uint(p1.z) & 0x7fffffffu,
             max(__float_as_uint(p2.z) & 0x7fffffffu, __float_as_uint(p3.z) & 0x7fffffffu)));
  }
  for (; i < n; i += stride) {
    float4 p = __ldg(xyz + i);
    mx = max(mx, __float_as_uint(p.x) & 0x7fffffffu);
    my = max(my, __float_as_uint(p.y) & 0x7fffffffu);
    mz = max(mz, __float_as_uint(p.z) & 0x7fffffffu);
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  my = __reduce_max_sync(0xffffffffu, my);
  mz = __reduce_max_sync(0xffffffffu, mz);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out3 + 0, mx);
    atomicMax(out3 + 1, my);
    atomicMax(out3 + 2, mz);
  }
}

__global__ void tc_params_kernel(const unsigned* __restrict__ absmax, float thr_up, float acc_ulps, PlaneTcParams* __restrict__ out) {
  PlaneTcParams P;
  P.use = 0;
  P.sigma = 1.0f; P.C = 0.0f; P.G = 0.0f; P.d_unc = 0.0f; P.bx = P.by = P.bz = 0.0f;
  const unsigned ux = absmax[0], uy = absmax[1], uz = absmax[2];
  if (ux < 0x7f800000u && uy < 0x7f800000u && uz < 0x7f800000u && thr_up > 0.0f && thr_up < 1e18f) {
    const double X = __uint_as_float(ux), Y = __uint_as_float(uy), Z = __uint_as_float(uz);
    // a unit-normal hypothesis through a cloud point has |d| <= |a|X+|b|Y+|c|Z <= R, so m <= 2R
    const double R = sqrt(X * X + Y * Y + Z * Z);
    const double G = 2.0 * R * (1.0 + 1e-6) + 1e-30;
    const double u = 5.9604644775390625e-08;  // 2^-24
    // |s~/sigma - s_exact_order| <= (3.0001 [exact order vs real] + 2^-3 [dropped a3 x3 products] + acc_ulps) u G
    const double beta = (3.0001 + 0.125 + (double)acc_ulps) * u * G + 1e-30;
    const double thr = (double)thr_up;
    // largest power of two with sigma beta + 2^-23 (sigma thr + 1) <= TC_WINDOW
    // (the second term covers the rounding of C and of the saturating subtraction)
    const double per_sigma = beta + thr * 1.1920928955078125e-07;
    double sigma = 0.0;
    if (per_sigma > 0.0) {
      int e = 0;
      frexp(((double)TC_WINDOW - 1.2e-7) / per_sigma, &e);  // = f 2^e, f in [0.5, 1)
      sigma = ldexp(1.0, e - 1);
      while (sigma * per_sigma + 1.2e-7 > (double)TC_WINDOW) sigma *= 0.5;
    }
    if (sigma > 1073741824.0) sigma = 1073741824.0;
    if (sigma >= 1.0 && sigma * G < 1e18 && sigma * thr < 4194304.0) {
      P.sigma = (float)sigma;
      P.C = (float)(sigma * thr + 0.5);
      P.G = (float)G;
      P.d_unc = (float)(sigma * thr);  // exact: power-of-two scaling of a float
      P.bx = (float)X; P.by = (float)Y; P.bz = (float)Z;
      P.use = 1;
    }
  }
  *out = P;
}

// Streaming cloud: the coordinate maxima come from the sample points alone and are doubled; tc_verify_kernel checks after
// the run that no point exceeded them (and that nothing was non-finite or timed out). When the check fails the counts
// are cleared and *ok = 0 sends the job to the exact kernel.
__global__ void tc_widen_kernel(unsigned* __restrict__ absmax) {
  if (threadIdx.x < 3) {
    const float v = __uint_as_float(absmax[threadIdx.x]);
    absmax[threadIdx.x] = __float_as_uint(v * 2.0f);  // inf / NaN stay what they are: the tensor path is then not used at all
  }
}
__global__ void tc_verify_kernel(const unsigned* __restrict__ assumed, const unsigned* __restrict__ seen, const PlaneTcParams* __restrict__ Pp,
                                 int* __restrict__ counts, int H, int* __restrict__ ok_out) {
  const bool ok = Pp->use && seen[0] <= assumed[0] && seen[1] <= assumed[1] && seen[2] <= assumed[2] && seen[3] == 0u;
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (!ok && h < H) counts[h] = 0;
  if (h == 0) *ok_out = ok ? 1 : 0;
}

// K-slot assignment of the two MMAs (hypothesis piece, point piece); the point side mirrors it in tc_point_image.
//   MMA 0 (small terms): (a1,x3) (a2,x2) (a3,x1) (a2,x3) (a3,x2) (b1,y3) (b2,y2) (b3,y1) |
//                        (b2,y3) (b3,y2) (c1,z3) (c2,z2) (c3,z1) (c2,z3) (c3,z2) (d3,1)
//   MMA 1 (large terms): (a1,x1) (a1,x2) (a2,x1) (b1,y1) (b1,y2) (b2,y1) (c1,z1) (c1,z2) |
//                        (c2,z1) (d1,1) (d2,1) 0 0 0 0 0
// A-operand image of one hypothesis block: 2 MMAs x (16 row groups x [8 rows x 16 B | 8 rows x 16 B]).
__global__ void __launch_bounds__(TC_M) tc_hyp_image_kernel(const HypRec* __restrict__ recs, int H, const PlaneTcParams* __restrict__ Pp,
                                                            uint4* __restrict__ image) {
  const PlaneTcParams P = *Pp;
  if (!P.use) return;
  const int h = blockIdx.x * TC_M + threadIdx.x;
  float a = 0.f, b = 0.f, c = 0.f, d = 1e18f;  // padding / invalid: certain outlier everywhere
  if (h < H) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(recs[h].v));
    if (r.x == r.x) {
      const float m = (fabsf(r.x) * P.bx + fabsf(r.y) * P.by) + (fabsf(r.z) * P.bz + fabsf(r.w));
      if (m <= P.G) {
        a = r.x * P.sigma; b = r.y * P.sigma; c = r.z * P.sigma; d = r.w * P.sigma;
      } else {
        d = P.d_unc;  // outside the bound the scale was derived for (or NaN): u = 0.5 everywhere, always re-evaluated
      }
    }
  }
  float a1, a2, a3, b1, b2, b3, c1, c2, c3, d1, d2, d3;
  tc_split3(a, a1, a2, a3);
  tc_split3(b, b1, b2, b3);
  tc_split3(c, c1, c2, c3);
  tc_split3(d, d1, d2, d3);
  const int row = threadIdx.x;
  uint4* blk = image + (size_t)blockIdx.x * (TC_A_BLOCK_BYTES / 16);
  const int o = (row >> 3) * 16 + (row & 7);  // 16-byte index inside one MMA image; the second K half is 8 further
  constexpr int MV = TC_A_MMA_BYTES / 16;
  blk[0 * MV + o] = tc_bf8(a1, a2, a3, a2, a3, b1, b2, b3);
  blk[0 * MV + o + 8] = tc_bf8(b2, b3, c1, c2, c3, c2, c3, d3);
  blk[1 * MV + o] = tc_bf8(a1, a1, a2, b1, b1, b2, c1, c1);
  blk[1 * MV + o + 8] = tc_bf8(c2, d1, d2, 0.f, 0.f, 0.f, 0.f, 0.f);
}
// B-operand rows of one point (pr = row inside its 256-point tile)
__device__ __forceinline__ void tc_point_image(uint4* tile, int pr, float4 p) {
  float x1, x2, x3, y1, y2, y3, z1, z2, z3;
  tc_split3(p.x, x1, x2, x3);
  tc_split3(p.y, y1, y2, y3);
  tc_split3(p.z, z1, z2, z3);
  const int o = (pr >> 3) * 16 + (pr & 7);
  constexpr int MV = TC_B_MMA_BYTES / 16;
  tile[0 * MV + o] = tc_bf8(x3, x2, x1, x3, x2, y3, y2, y1);
  tile[0 * MV + o + 8] = tc_bf8(y3, y2, z3, z2, z1, z3, z2, 1.f);
  tile[1 * MV + o] = tc_bf8(x1, x2, x1, y1, y2, y1, z1, z2);
  tile[1 * MV + o + 8] = tc_bf8(z1, 1.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------------
// the scoring kernel
// ------------------------------------------------------------------------------------------------
// exact count of one hypothesis (a, b, c, d broadcast from its owner lane) over len points of shared
// memory, whole warp (lanes = points)
__device__ __forceinline__ int tc_recount(const float4* pts, int len, float4 r, float thr_up, int lane) {
  int c = 0;
  for (int i = lane; i < len; i += 32) {
    const float4 p = pts[i];
    const float s = (r.x * p.x + r.z * p.z) + (r.y * p.y + r.w);  // -fmad=false: unfused, Eigen order
    c += (fabsf(s) < thr_up) ? 1 : 0;
  }
  return __reduce_add_sync(0xffffffffu, c);
}

// Packed FP32 pairs: FADD2 / FFMA2 retire two lanes per issue slot (tools/epi_probe.cu: 2.06 cycles per evaluation
// for the whole FADD.SAT + FADD2/2 + FFMA2/2 sequence against 3.17 with scalar FADD / FFMA).
__device__ __forceinline__ unsigned long long tc_pack2(float lo, float hi) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ unsigned long long tc_add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long tc_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float tc_sum2(unsigned long long v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
}
__device__ __forceinline__ void tc_accumulate32(const uint32_t (&r)[32], float C, unsigned long long (&S1)[2], unsigned long long (&S2)[2]) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const float nv0 = -fabsf(__uint_as_float(r[i])), nv1 = -fabsf(__uint_as_float(r[i + 1]));
    float u0, u1;
    asm("add.sat.f32 %0, %1, %2;" : "=f"(u0) : "f"(nv0), "f"(C));
    asm("add.sat.f32 %0, %1, %2;" : "=f"(u1) : "f"(nv1), "f"(C));
    const unsigned long long U = tc_pack2(u0, u1);
    S1[(i >> 1) & 1] = tc_add2(S1[(i >> 1) & 1], U);
    S2[(i >> 1) & 1] = tc_fma2(U, U, S2[(i >> 1) & 1]);
  }
}
// ragged run: only the first len columns are points of the cloud
__device__ __forceinline__ void tc_accumulate32_masked(const uint32_t (&r)[32], float C, unsigned long long (&S1)[2],
                                                       unsigned long long (&S2)[2], int len) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const float nv0 = -fabsf(__uint_as_float(r[i])), nv1 = -fabsf(__uint_as_float(r[i + 1]));
    float u0, u1;
    asm("add.sat.f32 %0, %1, %2;" : "=f"(u0) : "f"(nv0), "f"(C));
    asm("add.sat.f32 %0, %1, %2;" : "=f"(u1) : "f"(nv1), "f"(C));
    if (i >= len) u0 = 0.f;
    if (i + 1 >= len) u1 = 0.f;
    const unsigned long long U = tc_pack2(u0, u1);
    S1[(i >> 1) & 1] = tc_add2(S1[(i >> 1) & 1], U);
    S2[(i >> 1) & 1] = tc_fma2(U, U, S2[(i >> 1) & 1]);
  }
}

// Work schedule shared by the three roles of a CTA. An item = (super-block, 512-point chunk, range of the super-block's
// hypothesis blocks). Whole items go round robin over the CTAs (item = CTA + j * grid: the chunks are visited in cloud
// order, which is also the order a streaming cloud arrives in); the items of the last, incomplete round are cut along the
// hypothesis blocks so that every CTA gets the same share of them (the slowest CTA used to run 27 items against an average
// of 26.4: 5 % between the median and the slowest CTA).
struct TcSched {
  int n_chunks, n_hb, item_count, r_full, rem, sbh_t, tail_balanced;
};
__device__ __forceinline__ TcSched tc_sched(int n_chunks, int n_hb, int grid) {
  TcSched S;
  S.n_chunks = n_chunks;
  S.n_hb = n_hb;
  const int n_sb = (n_hb + TC_SB - 1) / TC_SB;
  S.item_count = n_sb * n_chunks;
  S.r_full = S.item_count / grid;
  S.rem = S.item_count - S.r_full * grid;
  const int sb_first = (S.rem > 0) ? (S.r_full * grid) / n_chunks : 0, sb_last = (S.item_count - 1) / n_chunks;
  S.tail_balanced = (S.rem > 0 && sb_first == sb_last) ? 1 : 0;  // the tail lies in one super-block: equal block counts
  S.sbh_t = min(TC_SB, n_hb - sb_last * TC_SB);
  return S;
}
// j-th item of CTA c; false when the CTA has no j-th item
__device__ __forceinline__ bool tc_item(const TcSched& S, int c, int grid, int j, int& sb, int& chunk, int& hb_lo, int& hb_hi) {
  int item;
  if (j < S.r_full || !S.tail_balanced) {
    item = c + j * grid;
    if (j > S.r_full || item >= S.item_count) return false;
    sb = item / S.n_chunks;
    chunk = item - sb * S.n_chunks;
    hb_lo = sb * TC_SB;
    hb_hi = min(S.n_hb, hb_lo + TC_SB);
    return true;
  }
  const int t = j - S.r_full;  // 0 or 1: a CTA's share of the tail touches at most two items
  if (t > 1) return false;
  const long long T = (long long)S.rem * S.sbh_t;
  const long long u0 = T * c / grid, u1 = T * (c + 1) / grid;
  const long long i0 = u0 / S.sbh_t;
  const long long lo = (t == 0) ? u0 : (i0 + 1) * S.sbh_t;
  const long long hi = (t == 0) ? (u1 < (i0 + 1) * S.sbh_t ? u1 : (i0 + 1) * S.sbh_t) : u1;
  if (lo >= hi) return false;
  const int idx = (int)(lo / S.sbh_t);
  item = S.r_full * grid + idx;
  sb = item / S.n_chunks;
  chunk = item - sb * S.n_chunks;
  hb_lo = sb * TC_SB + (int)(lo - (long long)idx * S.sbh_t);
  hb_hi = hb_lo + (int)(hi - lo);
  return true;
}

// Epilogue warp w: TMEM lane quadrant q = w & 3 (hypotheses 32 q .. 32 q + 31 of the block), column run j = w >> 2
// (columns 64 j .. 64 j + 63 of each accumulator, i.e. the points 256 b + 64 j .. + 63 of the chunk in phase b).
// DBG adds the accumulator dump, phase timers and the timing experiments (variant bits: 1 = no accumulation,
// 2 = one MMA per tile, 4 = no TMEM loads).
template <bool DBG>
__global__ void __launch_bounds__(TC_THREADS, 1)
plane_tc_kernel(const float4* __restrict__ xyz, int n, const HypRec* __restrict__ recs, int H, const uint4* __restrict__ image,
                int n_hb /*hypothesis blocks*/, int n_chunks, int n_items, float thr_up, const PlaneTcParams* __restrict__ Pp,
                int* __restrict__ counts, unsigned long long* __restrict__ stats /*nullable: [0] segments, [1] re-evaluated*/,
                float* __restrict__ dbg /*DBG: s~ of hypothesis block 0 x points 0..255 of item 0 (128 x 256)*/, int variant /*DBG*/,
                const int* ready /*nullable: streaming cloud, ready[k] != 0 once points [k ready_pts, (k+1) ready_pts) have landed*/,
                int ready_pts, unsigned* __restrict__ seen /*streaming: [0..2] |x|,|y|,|z| maxima of the points read, [3] time-out*/) {
  extern __shared__ __align__(128) unsigned char smem[];
  const PlaneTcParams P = *Pp;
  if (!P.use) return;  // plane_score_kernel (launched right after) does the work
  const long long t_start = clock64();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar0 = s_base + TC_OFF_BAR;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + TC_OFF_BAR + 16 * 8);
  float4* s_raw = reinterpret_cast<float4*>(smem + TC_OFF_RAW);
  int* s_cnt = reinterpret_cast<int*>(smem + TC_OFF_CNT);

  if (threadIdx.x == 0) {
    static_assert(TC_ASTAGES == 4 && TC_PHASES == 2, "barrier numbering");
    for (int i = 0; i < TC_ASTAGES; ++i) { mbar_init(BAR(TC_BAR_AFULL + i), 1); mbar_init(BAR(TC_BAR_AEMPTY + i), TC_MMA_WARPS); }
    for (int i = 0; i < TC_PHASES; ++i) { mbar_init(BAR(TC_BAR_FULL + i), 1); mbar_init(BAR(TC_BAR_EMPTY + i), TC_EPI_WARPS); }
    mbar_init(BAR(TC_BAR_B), TC_EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < TC_SB * TC_M; i += TC_THREADS) s_cnt[i] = 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const TcSched SCH = tc_sched(n_chunks, n_hb, (int)gridDim.x);
  (void)n_items;

  if (warp >= TC_EPI_WARPS) {
    // ===================== MMA issuers: warp t feeds accumulator t (phase t of every chunk); warp 0 also streams
    // the hypothesis block images global -> shared (bulk async copy), TC_ASTAGES - 1 blocks ahead.
    // The whole warp runs the loop so that every operand stays warp-uniform; one elected lane issues.
    const int t = warp - TC_EPI_WARPS;
    uint32_t ac = 0, cc = 0;
    long long tw = 0, tm = 0, tcm = 0, ta = 0, tb = 0;  // DBG: cycles in empty-wait, MMA issue, commit, A wait, B wait
    // producer cursor: block sequence number, item and hypothesis block of the next image to request
    uint32_t pa = 0;
    int p_j = 0, p_sb = 0, p_chunk = 0, p_hb = 0, p_hb_hi = 0;
    bool p_more = tc_item(SCH, blockIdx.x, gridDim.x, 0, p_sb, p_chunk, p_hb, p_hb_hi);
    auto produce_until = [&](uint32_t limit) {
      while (pa < limit && p_more) {
        const uint32_t st = pa % TC_ASTAGES, ph = (pa / TC_ASTAGES) & 1u;
        mbar_wait(BAR(TC_BAR_AEMPTY + st), ph ^ 1u);  // every MMA that read the previous block of this stage has completed
        if (elect_one()) {
          mbar_expect_tx(BAR(TC_BAR_AFULL + st), TC_A_BLOCK_BYTES);
          const uint4* src = image + (size_t)p_hb * (TC_A_BLOCK_BYTES / 16);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(s_base + TC_OFF_A + st * TC_A_BLOCK_BYTES), "l"(src), "r"(TC_A_BLOCK_BYTES), "r"(BAR(TC_BAR_AFULL + st))
                       : "memory");
        }
        __syncwarp();
        ++pa;
        if (++p_hb >= p_hb_hi) p_more = tc_item(SCH, blockIdx.x, gridDim.x, ++p_j, p_sb, p_chunk, p_hb, p_hb_hi);
      }
    };
    if (t == 0) produce_until(TC_ASTAGES - 1);
    const uint32_t b_addr = s_base + TC_OFF_B + t * TC_B_TILE_BYTES;
    int m_sb, m_chunk, hb0, hb1;
    for (int jj = 0; tc_item(SCH, blockIdx.x, gridDim.x, jj, m_sb, m_chunk, hb0, hb1); ++jj, ++cc) {
      long long c0 = DBG ? clock64() : 0;
      mbar_wait(BAR(TC_BAR_B), cc & 1u);  // the chunk's B image is in shared memory
      tc_fence_after();
      if (DBG) tb += clock64() - c0;
      for (int hb = hb0; hb < hb1; ++hb, ++ac) {
        const uint32_t st = ac % TC_ASTAGES, ph = (ac / TC_ASTAGES) & 1u;
        if (DBG) c0 = clock64();
        mbar_wait(BAR(TC_BAR_AFULL + st), ph);
        tc_fence_after();
        if (DBG) ta += clock64() - c0;
        const uint32_t a_addr = s_base + TC_OFF_A + st * TC_A_BLOCK_BYTES;
        // Everything the issue needs is in registers BEFORE the wait for the drained accumulator: the epilogue warps wait for
        // this warp's MMAs (wake-up + issue + 256 tensor cycles + commit), and this warp shares its issue port with four of them
        uint64_t da[TC_MMAS], db[TC_MMAS];
#pragma unroll
        for (int j = 0; j < TC_MMAS; ++j) {
          da[j] = tc_smem_desc(a_addr + j * TC_A_MMA_BYTES);
          db[j] = tc_smem_desc(b_addr + j * TC_B_MMA_BYTES);
          asm volatile("" : "+l"(da[j]), "+l"(db[j]));
        }
        const uint32_t bar_full = BAR(TC_BAR_FULL + t), bar_aempty = BAR(TC_BAR_AEMPTY + st), d_tmem = tmem + t * TC_N;
        const bool leader = elect_one();
        if (DBG) c0 = clock64();
        mbar_wait(BAR(TC_BAR_EMPTY + t), (ac & 1u) ^ 1u);  // every epilogue warp has the previous contents in registers
        tc_fence_after();
        long long c1 = DBG ? clock64() : 0;
        long long c2 = 0;
        if (leader) {
#pragma unroll
          for (int j = 0; j < TC_MMAS; ++j)
            if (!DBG || j == 0 || !(variant & 2))
              tc_mma_bf16(d_tmem, da[j], db[j], TC_IDESC, j > 0 ? 1u : 0u);
          if (DBG) c2 = clock64();
          tc_commit(bar_full);
          tc_commit(bar_aempty);  // the hypothesis stage is free once the MMAs of both warps have read it
        }
        __syncwarp();
        if (DBG) { tw += c1 - c0; tm += c2 - c1; tcm += clock64() - c2; }
        if (t == 0) produce_until(ac + TC_ASTAGES);  // keep TC_ASTAGES - 1 blocks beyond the one just issued in flight
      }
    }
    if (DBG && stats && blockIdx.x == 0 && lane == 0 && t == 0) {
      stats[162] = tw; stats[163] = tm; stats[164] = tcm; stats[165] = ta; stats[166] = tb; stats[167] = ac;
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps: split the point chunk, drain TMEM, count
    const int q = warp & 3, j = warp >> 2;  // lane quadrant, column run
    const int row = q * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + j * TC_RUN;
    uint32_t uc = 0;  // uses of each accumulator so far
    int cur_sb = -1;
    unsigned n_seg = 0, n_redo = 0;
    long long e_wait = 0, e_ld = 0, e_math = 0, e_tail = 0;  // DBG: cycles per phase
    auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory"); };
    auto flush = [&]() {
      if (cur_sb < 0) return;
      epi_sync();
      for (int i = threadIdx.x; i < TC_SB * TC_M; i += TC_EPI_THREADS) {
        const int h = cur_sb * TC_SB * TC_M + i;
        const int v = s_cnt[i];
        s_cnt[i] = 0;
        if (h < H && v) atomicAdd(&counts[h], v);
      }
      epi_sync();
    };
    int sb, chunk, hb0, hb1;
    for (int jj = 0; tc_item(SCH, blockIdx.x, gridDim.x, jj, sb, chunk, hb0, hb1); ++jj) {
      if (sb != cur_sb) { flush(); cur_sb = sb; }
      const int base = chunk * TC_CHUNK;
      // this warp's two runs of the chunk (phase 0 and 1) and how many of their points exist
      const int run0 = j * TC_RUN, run1 = TC_N + j * TC_RUN;
      const int len0 = max(0, min(TC_RUN, n - (base + run0))), len1 = max(0, min(TC_RUN, n - (base + run1)));
      const bool full_runs = (len0 == TC_RUN) && (len1 == TC_RUN);
      // every MMA that read the previous B image has completed (its accumulators were consumed);
      // nobody may still be re-counting from s_raw
      epi_sync();
      for (int pi = threadIdx.x; pi < max(TC_CHUNK, TC_EPI_THREADS); pi += TC_EPI_THREADS) {  // point of the chunk
        const int gi = base + pi;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pi >= TC_CHUNK) {
          // (24-warp shape: 768 threads, 384 points) nothing to split; warp-uniform, TC_CHUNK is a multiple of 32
        } else if (ready) {
          // The cloud is still arriving from the host (pitt_sac_segment_host): the copy stream raises ready[k] after chunk k.
          // One lane polls (acquire, so the points read below are the copied ones); a time-out flags the run as invalid
          // instead of hanging the GPU (the host then rescoring on the complete cloud).
          if (lane == 0) {
            // flags come up in copy order: the one of the chunk's last point covers a chunk that straddles two regions
            const int* f = ready + (min(base + TC_CHUNK, n) - 1) / ready_pts;
            // The wait is bounded by wall clock (2 s on %globaltimer), and once any warp has timed out (seen[3]) nobody waits
            // again: the run is void, the remaining chunks only have to drain.
            int v;
            unsigned long long t0 = 0ull, t;
            for (;;) {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
              if (v || *reinterpret_cast<volatile unsigned*>(seen + 3)) break;
              asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
              if (t0 == 0ull) t0 = t;
              else if (t - t0 > 2000000000ull) { atomicExch(seen + 3, 1u); break; }
              __nanosleep(200);
            }
          }
          __syncwarp();
          if (gi < n) p = __ldcg(xyz + gi);
          // the scale was derived from the sample points only: record what the cloud really contains
          unsigned mx = __reduce_max_sync(0xffffffffu, __float_as_uint(p.x) & 0x7fffffffu);
          unsigned my = __reduce_max_sync(0xffffffffu, __float_as_uint(p.y) & 0x7fffffffu);
          unsigned mz = __reduce_max_sync(0xffffffffu, __float_as_uint(p.z) & 0x7fffffffu);
          if (lane == 0) { atomicMax(seen + 0, mx); atomicMax(seen + 1, my); atomicMax(seen + 2, mz); }
        } else if (gi < n) {
          p = __ldg(xyz + gi);
        }
        if (pi < TC_CHUNK) {
          s_raw[pi] = p;
          tc_point_image(reinterpret_cast<uint4*>(smem + TC_OFF_B + (pi / TC_N) * TC_B_TILE_BYTES), pi % TC_N, p);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(BAR(TC_BAR_B));
      epi_sync();  // s_raw visible to every epilogue warp

#pragma unroll 1
      for (int hb = hb0; hb < hb1; ++hb, ++uc) {
        const int h = hb * TC_M + row;
        // this lane's hypothesis in the exact form, for re-evaluations (broadcast by shuffle, no memory latency there)
        unsigned long long S1[2] = {0ull, 0ull}, S2[2] = {0ull, 0ull};
        uint32_t ra[32], rb[32];
        // An accumulator is handed back to its MMA warp as soon as this warp's 64 columns are in registers: the
        // tensor pipe refills it (next hypothesis block) while the FMA pipe works through the other accumulator.
#pragma unroll
        for (int b = 0; b < TC_PHASES; ++b) {
          long long e0 = DBG ? clock64() : 0;
          mbar_wait(BAR(TC_BAR_FULL + b), uc & 1u);
          tc_fence_after();
          if (DBG) { e_wait += clock64() - e0; e0 = clock64(); }
          if (DBG && (variant & 5)) {
            if (!(variant & 4)) {
              unsigned x = 0;
              for (int k4 = 0; k4 < TC_RUN / 32; ++k4) {
                tc_ld32(ra, taddr + b * TC_N + 32 * k4);
                tc_ld_wait(ra);
#pragma unroll
                for (int i = 0; i < 32; ++i) x ^= ra[i];
              }
              if (x == 0x12345u) n_seg++;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(TC_BAR_EMPTY + b));
            continue;
          }
          constexpr int STEPS = TC_RUN > 64 ? TC_RUN / 64 : 1;  // 64-column steps of this warp's run (8-warp shape: two)
#pragma unroll
          for (int ks = 0; ks < STEPS; ++ks) {
            tc_ld32(ra, taddr + b * TC_N + 64 * ks);
            if (TC_RUN >= 64) {
              tc_ld32(rb, taddr + b * TC_N + 64 * ks + 32);
              tc_ld_wait2(ra, rb);
            } else {
              tc_ld_wait(ra);
            }
            if (ks == STEPS - 1) {  // the whole run is in registers (or consumed): hand the accumulator back
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(BAR(TC_BAR_EMPTY + b));
            }
            if (DBG && ks == 0) e_ld += clock64() - e0;
            if (DBG && dbg && blockIdx.x == 0 && jj == 0 && hb == 0 && b == 0 && STEPS == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                dbg[row * 256 + j * TC_RUN + i] = __uint_as_float(ra[i]);
                if (TC_RUN == 64) dbg[row * 256 + j * TC_RUN + 32 + i] = __uint_as_float(rb[i]);
              }
            }
            if (DBG) e0 = clock64();
            if (full_runs) {
              tc_accumulate32(ra, P.C, S1, S2);
              if (TC_RUN >= 64) tc_accumulate32(rb, P.C, S1, S2);
            } else {  // ragged last chunk of the cloud (warp-uniform)
              const int len = (b ? len1 : len0) - 64 * ks;
              tc_accumulate32_masked(ra, P.C, S1, S2, len);
              if (TC_RUN >= 64) tc_accumulate32_masked(rb, P.C, S1, S2, len - 32);
            }
          }
          if (DBG) { asm volatile("" : "+l"(S1[0]), "+l"(S1[1]), "+l"(S2[0]), "+l"(S2[1])); e_math += clock64() - e0; }
        }
        if (DBG && (variant & 5)) continue;
        long long p1 = DBG ? clock64() : 0;
        const float s1 = tc_sum2(tc_add2(S1[0], S1[1]));
        const float s2 = tc_sum2(tc_add2(S2[0], S2[1]));
        int c = (int)rintf(s1);
        const bool redo = (s1 - s2) > 0.1f;
        unsigned m = __ballot_sync(0xffffffffu, redo);
        if (DBG && stats && lane == 0) { n_seg += 32; n_redo += __popc(m); }
        while (m) {  // rare
          const int L = __ffs(m) - 1;
          m &= m - 1;
          float4 r;
          // (a segment is only re-evaluated for a real hypothesis: padding rows are certain outliers everywhere)
          r = __ldg(reinterpret_cast<const float4*>(recs[hb * TC_M + q * 32 + L].v));
          const int e = tc_recount(s_raw + run0, len0, r, thr_up, lane) + tc_recount(s_raw + run1, len1, r, thr_up, lane);
          if (lane == L) c = e;
        }
        if (c) atomicAdd(&s_cnt[(hb - sb * TC_SB) * TC_M + row], c);
        if (DBG) e_tail += clock64() - p1;
      }
    }
    flush();
    if (DBG && stats && blockIdx.x == 0 && threadIdx.x == 0) {
      stats[168] = e_wait; stats[169] = 0; stats[170] = 0; stats[171] = uc; stats[172] = e_ld; stats[173] = e_math; stats[174] = e_tail;
    }
    if (DBG && stats && lane == 0) {
      atomicAdd(stats + 0, (unsigned long long)n_seg);
      if (n_redo) atomicAdd(stats + 1, (unsigned long long)n_redo);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
  if (DBG && stats && threadIdx.x == 0) {  // per-CTA cycles and SM id (load-balance diagnostics)
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    stats[2 + blockIdx.x] = ((unsigned long long)smid << 48) | (unsigned long long)(clock64() - t_start);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
unsigned long long g_plane_tc_stats[2 + 160 + 16] = {0, 0};
int g_plane_tc_collect_stats = 0;
int g_plane_tc_dump = 0;
int g_plane_tc_variant = 0;
int g_plane_tc_nwq = 4;
float g_plane_tc_acc_ulps = TC_ACC_ULPS;
std::vector<float> g_plane_tc_dump_host;  // 128 x 256 accumulators + sigma, C

#define TC_LAUNCH_CHECK(ctx, what)                                        \
  do {                                                                    \
    cudaError_t e__ = cudaGetLastError();                                 \
    if (e__ != cudaSuccess) return fail((ctx), PITT_ERR_CUDA, what, e__); \
    (ctx)->launches++;                                                    \
  } while (0)

// Scores H plane hypotheses on the tensor path. *d_use_out points at an int that is 1 when the
// tensor kernel did the work and 0 when the caller must run the exact kernel (non-finite cloud,
// degenerate scale): the decision is made on the device, no host round trip.
// d_extra (nullable): further points every hypothesis may pass through (the sample points when c is only a chunk of the
// cloud): they enter the coordinate bounds the scale is derived from.
int launch_score_plane_tc(pitt_ctx* ctx, const pitt_cloud* c, const HypRec* d_recs, int H, const ScoreParams& sp, int* d_counts,
                          const int** d_use_out, const float4* d_extra, int n_extra, const int* d_ready, int ready_pts) {
  const int n = c->n;
  const int n_hb = cdiv(H, TC_M);
  const int n_chunks = cdiv(n, TC_CHUNK);
  const int n_sb = cdiv(n_hb, TC_SB);
  const long long items = (long long)n_sb * n_chunks;
  if (items > INT_MAX) return fail(ctx, PITT_ERR_INVALID, "plane scoring (tensor path): too many work items");
  unsigned* d_scr = nullptr;
  PlaneTcParams* d_P = nullptr;
  unsigned long long* d_stats = nullptr;
  uint4* d_image = nullptr;
  float* d_dbg = nullptr;
  PITT_TRY(arena_alloc(ctx, 4, &d_scr));
  PITT_TRY(arena_alloc(ctx, 1, &d_P));
  PITT_TRY(arena_alloc(ctx, 2 + 160 + 16, &d_stats));
  PITT_TRY(arena_alloc(ctx, (size_t)n_hb * (TC_A_BLOCK_BYTES / 16), &d_image));
  if (g_plane_tc_dump) PITT_TRY(arena_alloc(ctx, (size_t)TC_M * 256, &d_dbg));
  const bool streaming = d_ready != nullptr;  // one launch over a cloud that is still being copied
  if (streaming && !(d_extra && n_extra > 0)) return fail(ctx, PITT_ERR_INVALID, "streaming tensor path needs the sample points");
  unsigned* d_seen = nullptr;
  int* d_ok = nullptr;
  PITT_TRY(arena_alloc(ctx, 4, &d_seen));
  PITT_TRY(arena_alloc(ctx, 1, &d_ok));
  PITT_CUDA(ctx, cudaMemsetAsync(d_seen, 0, 4 * sizeof(unsigned), ctx->stream));
  PITT_CUDA(ctx, cudaMemsetAsync(d_scr, 0, 4 * sizeof(unsigned), ctx->stream));
  PITT_CUDA(ctx, cudaMemsetAsync(d_stats, 0, (2 + 160 + 16) * sizeof(unsigned long long), ctx->stream));
  if (!streaming) {
    int ab = std::min(cdiv(n, 256 * 4), ctx->sm_count * 8);
    tc_absmax_kernel<<<ab, 256, 0, ctx->stream>>>(c->d_xyz, n, d_scr);
    TC_LAUNCH_CHECK(ctx, "tc_absmax_kernel");
  }
  if (d_extra && n_extra > 0) {
    tc_absmax_kernel<<<std::min(cdiv(n_extra, 256 * 8), ctx->sm_count * 8), 256, 0, ctx->stream>>>(d_extra, n_extra, d_scr);
    TC_LAUNCH_CHECK(ctx, "tc_absmax_kernel");
  }
  if (streaming) {
    tc_widen_kernel<<<1, 32, 0, ctx->stream>>>(d_scr);
    TC_LAUNCH_CHECK(ctx, "tc_widen_kernel");
  }
  tc_params_kernel<<<1, 1, 0, ctx->stream>>>(d_scr, sp.thr_up, g_plane_tc_acc_ulps, d_P);
  TC_LAUNCH_CHECK(ctx, "tc_params_kernel");
  tc_hyp_image_kernel<<<n_hb, TC_M, 0, ctx->stream>>>(d_recs, H, d_P, d_image);
  TC_LAUNCH_CHECK(ctx, "tc_hyp_image_kernel");
  int grid = ctx->sm_count;
  if ((long long)grid > items) grid = (int)items;
  unsigned long long* st = g_plane_tc_collect_stats ? d_stats : nullptr;
  const bool dbgk = g_plane_tc_dump || g_plane_tc_variant || g_plane_tc_collect_stats;  // statistics live in the DBG kernel only
#define TC_LAUNCH(DBGK)                                                                                                         \
  do {                                                                                                                          \
    static bool attr_set[64] = {false}; /* the opt-in is per device */                                                          \
    if (!attr_set[ctx->device & 63]) {                                                                                          \
      PITT_CUDA(ctx, cudaFuncSetAttribute(plane_tc_kernel<DBGK>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));  \
      attr_set[ctx->device & 63] = true;                                                                                        \
    }                                                                                                                           \
    plane_tc_kernel<DBGK><<<grid, TC_THREADS, TC_SMEM_BYTES, ctx->stream>>>(                                                    \
        c->d_xyz, n, d_recs, H, d_image, n_hb, n_chunks, (int)items, sp.thr_up, d_P, d_counts, st, d_dbg, g_plane_tc_variant,  \
        d_ready, ready_pts, d_seen);                                                                                            \
  } while (0)
  if (ctx->time_tc_kernel) {
    if (!ctx->ev_k0) {
      PITT_CUDA(ctx, cudaEventCreate(&ctx->ev_k0));
      PITT_CUDA(ctx, cudaEventCreate(&ctx->ev_k1));
    }
    PITT_CUDA(ctx, cudaEventRecord(ctx->ev_k0, ctx->stream));
  }
  if (dbgk) TC_LAUNCH(true); else TC_LAUNCH(false);
  if (ctx->time_tc_kernel) PITT_CUDA(ctx, cudaEventRecord(ctx->ev_k1, ctx->stream));
#undef TC_LAUNCH
  TC_LAUNCH_CHECK(ctx, "plane_tc_kernel");
  *d_use_out = &d_P->use;
  if (streaming) {
    tc_verify_kernel<<<cdiv(H, 256), 256, 0, ctx->stream>>>(d_scr, d_seen, d_P, d_counts, H, d_ok);
    TC_LAUNCH_CHECK(ctx, "tc_verify_kernel");
    *d_use_out = d_ok;
  }
  if (g_plane_tc_collect_stats || g_plane_tc_dump) {
    if (g_plane_tc_collect_stats)
      PITT_CUDA(ctx, cudaMemcpyAsync(g_plane_tc_stats, d_stats, sizeof(g_plane_tc_stats), cudaMemcpyDeviceToHost, ctx->stream));
    if (g_plane_tc_dump) {
      g_plane_tc_dump_host.assign((size_t)TC_M * 256 + 2, 0.f);
      PITT_CUDA(ctx, cudaMemcpyAsync(g_plane_tc_dump_host.data(), d_dbg, (size_t)TC_M * 256 * sizeof(float), cudaMemcpyDeviceToHost,
                                     ctx->stream));
      PITT_CUDA(ctx, cudaMemcpyAsync(g_plane_tc_dump_host.data() + (size_t)TC_M * 256, d_P, 2 * sizeof(float), cudaMemcpyDeviceToHost,
                                     ctx->stream));
    }
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
  }
  return PITT_OK;
}

}  // namespace pitt
