// sac.cu — RANSAC on the device: hypothesis estimation (K2), inlier scoring (K3), winner
// selection (K4), ordered inlier selection (K5) and plane least-squares refinement (K6).
//
// Replaces pcl::RandomSampleConsensus::computeModel + SACSegmentation::segment reached from
// seg.segment() at supports_segmentation_srv.cpp:110, plane_segmentation_srv.cpp:67,
// sphere…:73, cylinder…:126, cone…:127.
#include <cfloat>
#include <climits>
#include <cmath>
#include <unordered_map>

#include "pitt_common.cuh"
#include <type_traits>

#include "sac.cuh"
#include "sac_device.cuh"

namespace pitt {

// =====================================================================================
// K2: one thread per hypothesis: minimal sample -> coefficients, validity, scoring record
// =====================================================================================
template <int MODEL>
__global__ void estimate_kernel(const float4* __restrict__ xyz, const float4* __restrict__ nrm,
                                const int* __restrict__ samples, int H, Limits L, ScoreParams sp, HypRec* __restrict__ recs,
                                float* __restrict__ coeffs8, uint8_t* __restrict__ flags, const int* __restrict__ skip,
                                const FitDesc* __restrict__ D = nullptr, int hoff = 0) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int SB = (MODEL == PITT_MODEL_PLANE) ? 3 : (MODEL == PITT_MODEL_SPHERE) ? 4 : (MODEL == PITT_MODEL_CYLINDER) ? 2 : 3;
  if (D) {  // batched: problem blockIdx.y, hypotheses [hoff, hoff + H) of its stream
    const FitDesc d = D[blockIdx.y];
    xyz = d.xyz; nrm = d.nrm;
    samples = d.samples + (size_t)hoff * SB;
    H = min(H, d.H - hoff);
    recs = d.recs + hoff;
    coeffs8 = d.coeffs8 + (size_t)hoff * 8;
    flags = d.flags + hoff;
    skip = hoff > 0 ? d.ints + 10 : nullptr;
  }
  if (h >= H || (skip && *skip)) return;  // skip: the stop rule has already ended the loop inside an earlier batch
  constexpr int S = (MODEL == PITT_MODEL_PLANE) ? 3 : (MODEL == PITT_MODEL_SPHERE) ? 4 : (MODEL == PITT_MODEL_CYLINDER) ? 2 : 3;
  int s[4];
#pragma unroll
  for (int i = 0; i < S; ++i) s[i] = samples[(size_t)h * S + i];
  float mc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) mc[i] = 0.0f;
  bool ok;
  if (MODEL == PITT_MODEL_PLANE) ok = estimate_plane(xyz, s, mc);
  else if (MODEL == PITT_MODEL_SPHERE) ok = estimate_sphere(xyz, s, mc);
  else if (MODEL == PITT_MODEL_CYLINDER) ok = estimate_cylinder(xyz, nrm, s, L, mc);
  else ok = estimate_cone(xyz, nrm, s, L, mc);
  bool valid = ok && model_valid<MODEL>(L, mc);
  HypRec r;
  make_rec<MODEL>(mc, valid, sp, r);
  recs[h] = r;
  if (coeffs8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) coeffs8[(size_t)h * 8 + i] = ok ? mc[i] : 0.0f;
  }
  flags[h] = (ok ? 1 : 0) | (valid ? 2 : 0);
}

// coefficients (device, 8 floats) -> one scoring record (isModelValid applied)
template <int MODEL>
__global__ void prep_rec_kernel(const float* __restrict__ coeffs, Limits L, ScoreParams sp, HypRec* __restrict__ rec,
                                const FitDesc* __restrict__ D = nullptr) {
  if (D) {  // batched: problem blockIdx.x, the winner's record goes to recs[0] (the hypotheses are not needed any more)
    coeffs = D[blockIdx.x].flt;
    rec = D[blockIdx.x].recs;
  }
  float mc[8];
  for (int i = 0; i < 8; ++i) mc[i] = coeffs[i];
  bool valid = model_valid<MODEL>(L, mc);
  HypRec r;
  make_rec<MODEL>(mc, valid, sp, r);
  *rec = r;
}

// =====================================================================================
// K3 (generic): lanes = points, hypotheses broadcast from shared memory.
// grid.x = point blocks, grid.y = hypothesis chunks. Each CTA keeps its chunk's counters in
// shared memory over all of its point tiles and flushes them once with one RED per hypothesis.
// =====================================================================================
template <int MODEL, int P, int TPB>
__global__ void __launch_bounds__(TPB)
score_kernel(const float4* __restrict__ xyz, const float4* __restrict__ nrm, int n, const HypRec* __restrict__ recs,
             int H, int hyp_chunk, int pts_per_cta, ScoreParams sp, int* __restrict__ counts, const int* __restrict__ skip,
             const FitDesc* __restrict__ D = nullptr, int hoff = 0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (D) {  // batched: problem blockIdx.z
    const FitDesc d = D[blockIdx.z];
    xyz = d.xyz; nrm = d.nrm; n = d.n;
    recs = d.recs + hoff;
    H = min(H, d.H - hoff);
    counts = d.counts + hoff;
    skip = hoff > 0 ? d.ints + 10 : nullptr;
    if (H <= 0) return;
  }
  if (skip && *skip) return;
  HypRec* s_rec = reinterpret_cast<HypRec*>(smem_raw);
  int* s_cnt = reinterpret_cast<int*>(smem_raw + (size_t)hyp_chunk * sizeof(HypRec));
  constexpr bool NEED_N = (MODEL == PITT_MODEL_CYLINDER || MODEL == PITT_MODEL_CONE);
  const int h0 = blockIdx.y * hyp_chunk;
  const int hc = min(hyp_chunk, H - h0);
  if (hc <= 0) return;
  // stage the hypothesis chunk (16 B per thread per step, coalesced)
  {
    const float4* src = reinterpret_cast<const float4*>(recs + h0);
    float4* dst = reinterpret_cast<float4*>(s_rec);
    for (int i = threadIdx.x; i < hc * 4; i += TPB) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < hc; i += TPB) s_cnt[i] = 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int p_begin = blockIdx.x * pts_per_cta;
  const int p_end = min(n, p_begin + pts_per_cta);
  for (int base = p_begin; base < p_end; base += TPB * P) {
    f3 pt[P], nv[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      int i = base + p * TPB + threadIdx.x;
      if (i < p_end) {
        pt[p] = ld3(xyz, i);
        if (NEED_N) nv[p] = ld3(nrm, i);
        else nv[p] = mk3(0.f, 0.f, 0.f);
      } else {
        pt[p] = mk3(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
        nv[p] = pt[p];
      }
    }
    for (int h = 0; h < hc; ++h) {
      RecRegs<MODEL> r;
      r.load(s_rec + h);
      if (s_rec[h].v[0] != s_rec[h].v[0]) continue;  // invalid hypothesis (uniform branch)
      int c = 0;
#pragma unroll
      for (int p = 0; p < P; ++p) c += r.inlier(pt[p], nv[p], sp) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (lane == 0 && c) atomicAdd(&s_cnt[h], c);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < hc; i += TPB) {
    int c = s_cnt[i];
    if (c) atomicAdd(&counts[h0 + i], c);
  }
}

// =====================================================================================
// K3 (cylinder, cone): two tiers. Tier 1 classifies every evaluation from the axis distance alone
// (Tier1<MODEL>::classify: certain outlier / certain inlier / undecided, ~20-60 instructions, no
// sqrt/acos/division for the cylinder); the undecided ones (thin shells around the model surface, a
// few per cent) are pushed as (hypothesis, point) pairs on a per-warp queue in shared memory and
// evaluated 32 at a time, one pair per lane, with the full predicate RecRegs<MODEL>::inlier
// (FP32 score, exact double sequence inside the band). Without the queue one undecided lane makes
// the whole warp walk the ~200-instruction path: at 2 % undecided that is every second warp.
// Counts are identical to score_kernel (test_two_tier_scoring_equals_generic).
// =====================================================================================
template <int MODEL, int P, int TPB>
__global__ void __launch_bounds__(TPB)
score2_kernel(const float4* __restrict__ xyz, const float4* __restrict__ nrm, int n, const HypRec* __restrict__ recs,
              int H, int hyp_chunk, int pts_per_cta, ScoreParams sp, int* __restrict__ counts, const int* __restrict__ skip,
              const FitDesc* __restrict__ D = nullptr, int hoff = 0) {
  constexpr int QCAP = 32 * P + 32;  // at most 31 left over + 32 * P pushed per hypothesis
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (D) {  // batched: problem blockIdx.z
    const FitDesc d = D[blockIdx.z];
    xyz = d.xyz; nrm = d.nrm; n = d.n;
    recs = d.recs + hoff;
    H = min(H, d.H - hoff);
    counts = d.counts + hoff;
    skip = hoff > 0 ? d.ints + 10 : nullptr;
    if (H <= 0) return;
  }
  if (skip && *skip) return;
  HypRec* s_rec = reinterpret_cast<HypRec*>(smem_raw);
  int* s_cnt = reinterpret_cast<int*>(smem_raw + (size_t)hyp_chunk * sizeof(HypRec));
  int2* s_q = reinterpret_cast<int2*>(smem_raw + (size_t)hyp_chunk * (sizeof(HypRec) + sizeof(int)) + 8 - ((size_t)hyp_chunk * 4) % 8);
  const int h0 = blockIdx.y * hyp_chunk;
  const int hc = min(hyp_chunk, H - h0);
  if (hc <= 0) return;
  {
    const float4* src = reinterpret_cast<const float4*>(recs + h0);
    float4* dst = reinterpret_cast<float4*>(s_rec);
    for (int i = threadIdx.x; i < hc * 4; i += TPB) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < hc; i += TPB) s_cnt[i] = 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int2* q = s_q + (threadIdx.x >> 5) * QCAP;
  const unsigned lt = (1u << lane) - 1u;
  auto full_eval = [&](int2 e) {
    RecRegs<MODEL> rr;
    rr.load(s_rec + e.x);
    const f3 pt = ld3(xyz, e.y), nv = ld3(nrm, e.y);
    if (rr.inlier(pt, nv, sp)) atomicAdd(&s_cnt[e.x], 1);
  };
  const int p_begin = blockIdx.x * pts_per_cta;
  const int p_end = min(n, p_begin + pts_per_cta);
  for (int base = p_begin; base < p_end; base += TPB * P) {
    f3 pt[P];
    unsigned valid = 0, nice = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const int i = base + p * TPB + threadIdx.x;
      pt[p] = mk3(0.f, 0.f, 0.f);
      if (i < p_end) {
        pt[p] = ld3(xyz, i);
        valid |= 1u << p;
        if (nice_point(pt[p], ld3(nrm, i))) nice |= 1u << p;
      }
    }
    int qn = 0;  // warp-uniform
    // Full tiles whose points may all take the certain-inlier shortcut (the normal case) run the loop without the two
    // per-point masks (warp-uniform choice).
    const bool plain = __all_sync(0xffffffffu, valid == (1u << P) - 1u && nice == (1u << P) - 1u);
    auto hyp_loop = [&](auto plain_tag) {
      constexpr bool PLAIN = decltype(plain_tag)::value;
      for (int h = 0; h < hc; ++h) {
        if (s_rec[h].v[0] != s_rec[h].v[0]) continue;  // invalid hypothesis (uniform branch)
        Tier1<MODEL> r;
        r.load(s_rec + h);
        int c = 0;
#pragma unroll
        for (int p = 0; p < P; ++p) {
          bool in, und;
          r.classify(pt[p], PLAIN ? true : (bool)((nice >> p) & 1u), in, und);
          if (!PLAIN) {  // padding lanes classify garbage
            const bool v = (valid >> p) & 1u;
            in &= v;
            und &= v;
          }
          c += in ? 1 : 0;
          const unsigned m = __ballot_sync(0xffffffffu, und);
          if (und) q[qn + __popc(m & lt)] = make_int2(h, base + p * TPB + (int)threadIdx.x);  // no branch around the push:
          qn += __popc(m);                                                                    // it is taken 4 times out of 5
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0 && c) atomicAdd(&s_cnt[h], c);
        if (qn >= 32) {
          __syncwarp();
          do {
            qn -= 32;
            full_eval(q[qn + lane]);
          } while (qn >= 32);
          __syncwarp();
        }
      }
    };
    if (plain) hyp_loop(std::true_type{});
    else hyp_loop(std::false_type{});
    __syncwarp();
    if (lane < qn) full_eval(q[lane]);
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < hc; i += TPB) {
    int c = s_cnt[i];
    if (c) atomicAdd(&counts[h0 + i], c);
  }
}

// =====================================================================================
// K3p: plane scoring, lanes = hypotheses, points broadcast from shared memory, packed f32x2
// arithmetic (FMUL2/FADD2 are IEEE round-to-nearest per element => same bits as scalar).
// Each thread owns KH hypotheses (KH/2 register pairs) and its own counters: no reduction at all
// until the final RED per hypothesis.
// =====================================================================================
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  return ((unsigned long long)__float_as_uint(hi) << 32) | (unsigned long long)__float_as_uint(lo);
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad=false, which would
// round a*x+c*z once instead of three times. The packed add is therefore issued as
// fma(x, ONE, y) with ONE = 1.0f taken from a kernel argument: round(x*1 + y) == round(x + y)
// bit for bit, and an FMA cannot be contracted any further.
__device__ __forceinline__ unsigned long long fadd2_exact(unsigned long long a, unsigned long long b, unsigned long long one) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(one), "l"(b));
  return d;
}
// cnt += (a < b): FSETP + predicated IADD3 (two ALU-pipe instructions; NaN compares false)
__device__ __forceinline__ void count_if_lt(int& cnt, float a, float b) {
  asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt) : "f"(a), "f"(b));
}
__device__ __forceinline__ float lo32(unsigned long long v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
__device__ __forceinline__ float hi32(unsigned long long v) { return __uint_as_float((unsigned)(v >> 32)); }

template <int KH, int TPB, int TILE>
__global__ void __launch_bounds__(TPB)
plane_score_kernel(const float4* __restrict__ xyz, int n, const HypRec* __restrict__ recs, int H, int n_ptiles,
                   int n_items, float thr_up, float one_rt, int* __restrict__ work_counter, int* __restrict__ counts,
                   const int* __restrict__ skip_if /*nullable: *skip_if != 0 => the filter kernel did the work*/) {
  if (skip_if && *skip_if) return;
  // Persistent CTAs pull work items (hypothesis block hb, point tile pt) from an atomic counter:
  // item = hb * n_ptiles + pt. The next item's tile is prefetched with cp.async while the current
  // one is scored, so the FP32 pipe never waits on L2/HBM and there is no wave-quantisation tail.
  __shared__ __align__(16) float4 s_pts[2][TILE];
  __shared__ int s_item[2];
  const unsigned long long ONE = pack2(one_rt, one_rt);
  constexpr int KP = KH / 2;
  unsigned long long A[KP], B[KP], C[KP], D[KP];
  int cnt[KH];
  int cur_hb = -1;
  auto flush = [&]() {
    if (cur_hb < 0) return;
    const int h_base = (cur_hb * TPB + threadIdx.x) * KH;
#pragma unroll
    for (int k = 0; k < KH; ++k)
      if (h_base + k < H && cnt[k]) atomicAdd(&counts[h_base + k], cnt[k]);
  };
  auto load_hyps = [&](int hb) {
    const int h_base = (hb * TPB + threadIdx.x) * KH;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      float4 r0 = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F), r1 = r0;
      if (h_base + 2 * k < H) r0 = __ldg(reinterpret_cast<const float4*>(recs[h_base + 2 * k].v));
      if (h_base + 2 * k + 1 < H) r1 = __ldg(reinterpret_cast<const float4*>(recs[h_base + 2 * k + 1].v));
      A[k] = pack2(r0.x, r1.x); B[k] = pack2(r0.y, r1.y); C[k] = pack2(r0.z, r1.z); D[k] = pack2(r0.w, r1.w);
      cnt[2 * k] = 0; cnt[2 * k + 1] = 0;
    }
    cur_hb = hb;
  };
  auto stage = [&](int item, int buf) {
    if (item < n_items) {
      const int base = (item % n_ptiles) * TILE;
      for (int i = threadIdx.x; i < TILE; i += TPB) {
        int gi = base + i;
        if (gi < n) {
          unsigned saddr = (unsigned)__cvta_generic_to_shared(&s_pts[buf][i]);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(xyz + gi));
        } else {
          s_pts[buf][i] = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
        }
      }
    }
    asm volatile("cp.async.commit_group;");
  };
  if (threadIdx.x == 0) s_item[0] = atomicAdd(work_counter, 1);
  __syncthreads();
  int item = s_item[0];
  stage(item, 0);
  int it = 0;
  while (item < n_items) {
    // claim the next item and start its loads before scoring the current tile
    if (threadIdx.x == 0) s_item[(it + 1) & 1] = atomicAdd(work_counter, 1);
    asm volatile("cp.async.wait_group 0;");
    __syncthreads();  // tile `it` landed for everyone; s_item[(it+1)&1] visible; buffer (it+1)&1 is free
    const int next = s_item[(it + 1) & 1];
    stage(next, (it + 1) & 1);
    const int hb = item / n_ptiles;
    if (hb != cur_hb) {
      flush();
      load_hyps(hb);
    }
    const float4* tile = s_pts[it & 1];
    // two points per step (ILP): 6 packed FP32 instructions + 2x(FSETP, @p IADD3) per evaluation pair
#pragma unroll 2
    for (int i = 0; i < TILE; i += 2) {
      float4 p = tile[i], q = tile[i + 1];  // LDS.128 broadcasts
      unsigned long long PX = pack2(p.x, p.x), PY = pack2(p.y, p.y), PZ = pack2(p.z, p.z);
      unsigned long long QX = pack2(q.x, q.x), QY = pack2(q.y, q.y), QZ = pack2(q.z, q.z);
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        // (a*x + c*z) + (b*y + d), two hypotheses per instruction
        unsigned long long sp = fadd2_exact(fadd2_exact(mul2(A[k], PX), mul2(C[k], PZ), ONE),
                                            fadd2_exact(mul2(B[k], PY), D[k], ONE), ONE);
        unsigned long long sq = fadd2_exact(fadd2_exact(mul2(A[k], QX), mul2(C[k], QZ), ONE),
                                            fadd2_exact(mul2(B[k], QY), D[k], ONE), ONE);
        count_if_lt(cnt[2 * k], fabsf(lo32(sp)), thr_up);
        count_if_lt(cnt[2 * k + 1], fabsf(hi32(sp)), thr_up);
        count_if_lt(cnt[2 * k], fabsf(lo32(sq)), thr_up);
        count_if_lt(cnt[2 * k + 1], fabsf(hi32(sq)), thr_up);
      }
    }
    item = next;
    ++it;
  }
  asm volatile("cp.async.wait_group 0;");
  flush();
}

// =====================================================================================
// K3s: sphere scoring for large jobs, same structure as plane_score_kernel: lanes = hypotheses (KH per
// thread as KH/2 packed pairs), points broadcast from shared memory, persistent CTAs pulling
// (hypothesis block, point tile) items, no reduction until the final RED per hypothesis.
// Per evaluation: d2 = (dx*dx + dy*dy) + dz*dz in PCL's operation order (8 separately rounded FP32
// operations = 4 packed instructions) and the interval test d_lo <= d2 <= d_hi of
// RecRegs<PITT_MODEL_SPHERE>::inlier (two chained FSETP + a predicated IADD3 on the ALU pipe).
// NaN points and invalid (all-NaN) hypotheses compare false. Counts are identical to
// score_kernel<PITT_MODEL_SPHERE> (test_sphere_packed_kernel_equals_generic).
// =====================================================================================
// cnt += (lo <= a && a <= hi)
__device__ __forceinline__ void count_if_within(int& cnt, float a, float lo, float hi) {
  asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\tsetp.le.and.f32 p, %1, %3, p;\n\t@p add.s32 %0, %0, 1;\n\t}"
      : "+r"(cnt) : "f"(a), "f"(lo), "f"(hi));
}

template <int KH, int TPB, int TILE>
__global__ void __launch_bounds__(TPB)
sphere_score_kernel(const float4* __restrict__ xyz, int n, const HypRec* __restrict__ recs, int H, int n_ptiles,
                    int n_items, float one_rt, int* __restrict__ work_counter, int* __restrict__ counts) {
  __shared__ __align__(16) float4 s_pts[2][TILE];
  __shared__ int s_item[2];
  const unsigned long long ONE = pack2(one_rt, one_rt);
  constexpr int KP = KH / 2;
  unsigned long long NX[KP], NY[KP], NZ[KP];  // minus the centre: x - cx == x + (-cx) bit for bit
  float lo[KH], hi[KH];
  int cnt[KH];
  int cur_hb = -1;
  auto flush = [&]() {
    if (cur_hb < 0) return;
    const int h_base = (cur_hb * TPB + threadIdx.x) * KH;
#pragma unroll
    for (int k = 0; k < KH; ++k)
      if (h_base + k < H && cnt[k]) atomicAdd(&counts[h_base + k], cnt[k]);
  };
  auto load_hyps = [&](int hb) {
    const int h_base = (hb * TPB + threadIdx.x) * KH;
    const float nanv = CUDART_NAN_F;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      float4 r0 = make_float4(nanv, nanv, nanv, nanv), r1 = r0;
      float2 i0 = make_float2(nanv, nanv), i1 = i0;
      if (h_base + 2 * k < H) {
        r0 = __ldg(reinterpret_cast<const float4*>(recs[h_base + 2 * k].v));
        i0 = __ldg(reinterpret_cast<const float2*>(recs[h_base + 2 * k].v + 4));
      }
      if (h_base + 2 * k + 1 < H) {
        r1 = __ldg(reinterpret_cast<const float4*>(recs[h_base + 2 * k + 1].v));
        i1 = __ldg(reinterpret_cast<const float2*>(recs[h_base + 2 * k + 1].v + 4));
      }
      NX[k] = pack2(-r0.x, -r1.x); NY[k] = pack2(-r0.y, -r1.y); NZ[k] = pack2(-r0.z, -r1.z);
      lo[2 * k] = i0.x; hi[2 * k] = i0.y; lo[2 * k + 1] = i1.x; hi[2 * k + 1] = i1.y;
      cnt[2 * k] = 0; cnt[2 * k + 1] = 0;
    }
    cur_hb = hb;
  };
  auto stage = [&](int item, int buf) {
    if (item < n_items) {
      const int base = (item % n_ptiles) * TILE;
      for (int i = threadIdx.x; i < TILE; i += TPB) {
        int gi = base + i;
        if (gi < n) {
          unsigned saddr = (unsigned)__cvta_generic_to_shared(&s_pts[buf][i]);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(xyz + gi));
        } else {
          s_pts[buf][i] = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
        }
      }
    }
    asm volatile("cp.async.commit_group;");
  };
  if (threadIdx.x == 0) s_item[0] = atomicAdd(work_counter, 1);
  __syncthreads();
  int item = s_item[0];
  stage(item, 0);
  int it = 0;
  while (item < n_items) {
    if (threadIdx.x == 0) s_item[(it + 1) & 1] = atomicAdd(work_counter, 1);
    asm volatile("cp.async.wait_group 0;");
    __syncthreads();
    const int next = s_item[(it + 1) & 1];
    stage(next, (it + 1) & 1);
    const int hb = item / n_ptiles;
    if (hb != cur_hb) {
      flush();
      load_hyps(hb);
    }
    const float4* tile = s_pts[it & 1];
#pragma unroll 2
    for (int i = 0; i < TILE; ++i) {
      float4 p = tile[i];  // LDS.128 broadcast
      unsigned long long PX = pack2(p.x, p.x), PY = pack2(p.y, p.y), PZ = pack2(p.z, p.z);
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        unsigned long long dx = add2(PX, NX[k]), dy = add2(PY, NY[k]), dz = add2(PZ, NZ[k]);
        // (dx*dx + dy*dy) + dz*dz with every product and sum rounded on its own (see fadd2_exact)
        unsigned long long d2 = fadd2_exact(fadd2_exact(mul2(dx, dx), mul2(dy, dy), ONE), mul2(dz, dz), ONE);
        count_if_within(cnt[2 * k], lo32(d2), lo[2 * k], hi[2 * k]);
        count_if_within(cnt[2 * k + 1], hi32(d2), lo[2 * k + 1], hi[2 * k + 1]);
      }
    }
    item = next;
    ++it;
  }
  asm volatile("cp.async.wait_group 0;");
  flush();
}

// =====================================================================================
// K3f: plane scoring with an FFMA filter and exact re-evaluation of the uncertain evaluations.
//
// The exact predicate |fl(fl(fl(a x)+fl(c z)) + fl(fl(b y)+d))| < thr needs 6 separately rounded
// FP32 operations per evaluation. The filter evaluates s~ = fma(a',x, fma(b',y, fma(c',z, d')))
// with the coefficients scaled by a power of two sigma, then t = fma(s~, s~, -T), T = fl(sigma^2 thr^2):
// 4 FMA-pipe operations. Both s (exact order) and s~/sigma are within 3.0001 u m of the real dot
// product (u = 2^-24, m = |a x|+|b y|+|c z|+|d| <= G), so their difference is below beta = 8 u G.
// sigma is chosen (plane_filter_params_kernel) so that
//     sigma^2 (beta + 2 thr u)(2 thr + beta + 2 thr u)(1 + 4u) < 2,
// which makes |t| >= 2 (bit 30 of t set) imply ||s~|/sigma - thr| > beta + thr u, i.e. the sign of t
// IS the exact predicate. The inner loop therefore only accumulates the sign bits of t (the count)
// and the AND of the t words (bit 30 clear = some evaluation of this (hypothesis, tile) pair is
// uncertain). Uncertain pairs are re-evaluated with the exact operation order by the whole warp
// (lanes = points) and the filter's count for the pair is replaced: the result is bit-identical to
// plane_score_kernel. Non-finite clouds and degenerate scales take plane_score_kernel instead.
// =====================================================================================
struct PlaneFilterParams {
  float sigma;      // power of two
  float neg_T;      // -fl(sigma^2 thr_up^2)
  float G;          // bound on |a x|+|b y|+|c z|+|d| the scale was derived for
  float sd_unc;     // d' of a hypothesis that must always be re-evaluated (t == +0)
  float bx, by, bz; // cloud |x|,|y|,|z| maxima
  int use_filter;   // 0: cloud not finite or scale out of range -> exact kernel
};
__global__ void __launch_bounds__(256) cloud_absmax_kernel(const float4* __restrict__ xyz, int n, unsigned* __restrict__ out3) {
  unsigned mx = 0, my = 0, mz = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = __ldg(xyz + i);
    // |v| as an unsigned word: ordered like the float for finite values, inf/NaN compare above every finite value
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
__global__ void plane_filter_params_kernel(const unsigned* __restrict__ absmax, float thr_up, PlaneFilterParams* __restrict__ out) {
  PlaneFilterParams P;
  P.use_filter = 0;
  P.sigma = 1.0f; P.neg_T = 0.0f; P.G = 0.0f; P.sd_unc = 0.0f; P.bx = P.by = P.bz = 0.0f;
  const unsigned ux = absmax[0], uy = absmax[1], uz = absmax[2];
  if (ux < 0x7f800000u && uy < 0x7f800000u && uz < 0x7f800000u && thr_up > 0.0f && thr_up < 1e18f) {
    const double X = __uint_as_float(ux), Y = __uint_as_float(uy), Z = __uint_as_float(uz);
    // a hypothesis through a cloud point with a unit normal has |d| <= |a|X+|b|Y+|c|Z <= R
    const double R = sqrt(X * X + Y * Y + Z * Z);
    const double G = 2.0 * R * (1.0 + 1e-6) + 1e-30;
    const double u = 5.9604644775390625e-08;  // 2^-24
    const double thr = (double)thr_up;
    const double beta = 8.0 * u * G + 1e-30;
    const double Q0 = (beta + 2.0 * thr * u) * (2.0 * thr + beta + 2.0 * thr * u) * (1.0 + 4.0 * u);
    // largest sigma = 2^k with sigma^2 Q0 < 2
    int e = 0;
    frexp(2.0 / Q0, &e);            // 2/Q0 = f 2^e, f in [0.5,1)  =>  2^(e-1) <= 2/Q0
    int k = (e - 1) / 2;
    if ((e - 1) < 0 && ((e - 1) & 1)) k -= 1;  // floor for negative odd exponents
    double sigma = ldexp(1.0, k);
    while (sigma * sigma * Q0 >= 2.0) sigma *= 0.5;
    if (sigma > 1073741824.0) sigma = 1073741824.0;
    // keep every product finite and normal, and T small enough that fma(sqrt(T), sqrt(T), -T) stays below 2
    if (sigma >= 1.0 && sigma * G < 1e18 && sigma * sigma * thr * thr < 1048576.0) {
      const float T = (float)(sigma * sigma * thr * thr);
      P.sigma = (float)sigma;
      P.neg_T = -T;
      P.G = (float)G;
      P.sd_unc = sqrtf(T);  // fma(sd,sd,-T) is within an ulp of 0: bit 30 clear, always uncertain
      P.bx = (float)X; P.by = (float)Y; P.bz = (float)Z;
      P.use_filter = 1;
    }
  }
  *out = P;
}
// scaled filter coefficients of one hypothesis (shared by the hot loop and the re-evaluation)
__device__ __forceinline__ void plane_filter_coeffs(float4 r, const PlaneFilterParams& P, float& sa, float& sb, float& sc, float& sd) {
  if (!(r.x == r.x)) {  // invalid hypothesis (all NaN): certain outlier everywhere, never re-evaluated
    sa = sb = sc = 0.0f;
    sd = 1e18f;
    return;
  }
  const float m = (fabsf(r.x) * P.bx + fabsf(r.y) * P.by) + (fabsf(r.z) * P.bz + fabsf(r.w));
  if (!(m <= P.G)) {    // outside the bound the scale was derived for (or NaN): always re-evaluated
    sa = sb = sc = 0.0f;
    sd = P.sd_unc;
    return;
  }
  sa = r.x * P.sigma; sb = r.y * P.sigma; sc = r.z * P.sigma; sd = r.w * P.sigma;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// Whole-warp re-evaluation of one (hypothesis, tile) pair: returns exact count - filter count
// (filter_too) or the exact count alone. tile = shared memory, len points.
__device__ __noinline__ int plane_recount(const float4* tile, int len, const HypRec* rec, const PlaneFilterParams* Pp,
                                          float thr_up, int filter_too) {
  const PlaneFilterParams P = *Pp;
  const float4 r = __ldg(reinterpret_cast<const float4*>(rec->v));
  float sa, sb, sc, sd;
  plane_filter_coeffs(r, P, sa, sb, sc, sd);
  int c = 0;
  for (int i = threadIdx.x & 31; i < len; i += 32) {
    const float4 p = tile[i];
    const float s = (r.x * p.x + r.z * p.z) + (r.y * p.y + r.w);  // -fmad=false: unfused, Eigen order
    c += (fabsf(s) < thr_up) ? 1 : 0;
    if (filter_too) {
      const float sf = __fmaf_rn(sa, p.x, __fmaf_rn(sb, p.y, __fmaf_rn(sc, p.z, sd)));
      const float t = __fmaf_rn(sf, sf, P.neg_T);
      c -= (int)(__float_as_uint(t) >> 31);
    }
  }
  return __reduce_add_sync(0xffffffffu, c);
}

template <int KH, int TPB, int TILE>
__global__ void __launch_bounds__(TPB)
plane_filter_kernel(const float4* __restrict__ xyz, int n, const HypRec* __restrict__ recs, int H, int n_ptiles,
                    int n_items, float thr_up, const PlaneFilterParams* __restrict__ Pp, int* __restrict__ work_counter,
                    int* __restrict__ counts, unsigned long long* __restrict__ stats /*[0] pairs, [1] re-evaluated pairs*/) {
  __shared__ __align__(16) float4 s_pts[2][TILE];
  __shared__ int s_item[2];
  const PlaneFilterParams P = *Pp;
  if (!P.use_filter) return;  // plane_score_kernel (launched right after) does the work
  constexpr int KP = KH / 2;
  const unsigned long long NEGT = pack2(P.neg_T, P.neg_T);
  unsigned long long A[KP], B[KP], Cc[KP], D[KP];
  int cnt[KH];
  int cur_hb = -1;
  int n_redo = 0;
  auto flush = [&]() {
    if (cur_hb < 0) return;
    const int h_base = (cur_hb * TPB + threadIdx.x) * KH;
#pragma unroll
    for (int k = 0; k < KH; ++k)
      if (h_base + k < H && cnt[k]) atomicAdd(&counts[h_base + k], cnt[k]);
  };
  auto load_hyps = [&](int hb) {
    const int h_base = (hb * TPB + threadIdx.x) * KH;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      float4 r0 = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F), r1 = r0;
      if (h_base + 2 * k < H) r0 = __ldg(reinterpret_cast<const float4*>(recs[h_base + 2 * k].v));
      if (h_base + 2 * k + 1 < H) r1 = __ldg(reinterpret_cast<const float4*>(recs[h_base + 2 * k + 1].v));
      float a0, b0, c0, d0, a1, b1, c1, d1;
      plane_filter_coeffs(r0, P, a0, b0, c0, d0);
      plane_filter_coeffs(r1, P, a1, b1, c1, d1);
      A[k] = pack2(a0, a1); B[k] = pack2(b0, b1); Cc[k] = pack2(c0, c1); D[k] = pack2(d0, d1);
      cnt[2 * k] = 0; cnt[2 * k + 1] = 0;
    }
    cur_hb = hb;
  };
  auto stage = [&](int item, int buf) {
    if (item < n_items) {
      const int base = (item % n_ptiles) * TILE;
      for (int i = threadIdx.x; i < TILE; i += TPB) {
        int gi = base + i;
        if (gi < n) {
          unsigned saddr = (unsigned)__cvta_generic_to_shared(&s_pts[buf][i]);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(xyz + gi));
        } else {
          s_pts[buf][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    asm volatile("cp.async.commit_group;");
  };
  if (threadIdx.x == 0) s_item[0] = atomicAdd(work_counter, 1);
  __syncthreads();
  int item = s_item[0];
  stage(item, 0);
  int it = 0;
  const int lane = threadIdx.x & 31;
  while (item < n_items) {
    if (threadIdx.x == 0) s_item[(it + 1) & 1] = atomicAdd(work_counter, 1);
    asm volatile("cp.async.wait_group 0;");
    __syncthreads();
    const int next = s_item[(it + 1) & 1];
    stage(next, (it + 1) & 1);
    const int hb = item / n_ptiles;
    if (hb != cur_hb) {
      flush();
      load_hyps(hb);
    }
    const float4* tile = s_pts[it & 1];
    const int len = min(TILE, n - (item % n_ptiles) * TILE);
    const int warp_t0 = threadIdx.x & ~31;
    if (len == TILE) {
      unsigned andw[KH];
#pragma unroll
      for (int k = 0; k < KH; ++k) andw[k] = 0xffffffffu;
      // (A software-pipelined variant that interleaves the ALU work of the previous point pair with the
      // FMA work of the current one was measured slower, 1.20 vs 1.09 ms on C2: the loop is bound by
      // register-file operand reads, not by pipe alternation; tools/plane_variants.cu)
#pragma unroll 2
      for (int i = 0; i < TILE; i += 2) {
        float4 p = tile[i], q = tile[i + 1];  // LDS.128 broadcasts
        unsigned long long PX = pack2(p.x, p.x), PY = pack2(p.y, p.y), PZ = pack2(p.z, p.z);
        unsigned long long QX = pack2(q.x, q.x), QY = pack2(q.y, q.y), QZ = pack2(q.z, q.z);
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          unsigned long long sp = fma2(A[k], PX, fma2(B[k], PY, fma2(Cc[k], PZ, D[k])));
          unsigned long long sq = fma2(A[k], QX, fma2(B[k], QY, fma2(Cc[k], QZ, D[k])));
          unsigned long long tp = fma2(sp, sp, NEGT);
          unsigned long long tq = fma2(sq, sq, NEGT);
          const unsigned tp0 = (unsigned)tp, tp1 = (unsigned)(tp >> 32), tq0 = (unsigned)tq, tq1 = (unsigned)(tq >> 32);
          cnt[2 * k] += (int)(tp0 >> 31);
          cnt[2 * k] += (int)(tq0 >> 31);
          cnt[2 * k + 1] += (int)(tp1 >> 31);
          cnt[2 * k + 1] += (int)(tq1 >> 31);
          andw[2 * k] &= tp0 & tq0;
          andw[2 * k + 1] &= tp1 & tq1;
        }
      }
      unsigned unc = 0;
#pragma unroll
      for (int k = 0; k < KH; ++k) unc |= ((andw[k] >> 30) & 1u) ? 0u : (1u << k);
      if (__any_sync(0xffffffffu, unc != 0)) {
#pragma unroll
        for (int k = 0; k < KH; ++k) {
          unsigned m = __ballot_sync(0xffffffffu, (unc >> k) & 1u);
          while (m) {
            const int L = __ffs(m) - 1;
            m &= m - 1;
            const int h = (cur_hb * TPB + warp_t0 + L) * KH + k;  // < H: padding hypotheses are never uncertain
            const int delta = plane_recount(tile, TILE, recs + h, Pp, thr_up, 1);
            if (lane == L) cnt[k] += delta;
            if (lane == 0) ++n_redo;
          }
        }
      }
    } else {
      // ragged last tile: exact counts for every hypothesis of the block
#pragma unroll
      for (int k = 0; k < KH; ++k) {
#pragma unroll 1
        for (int L = 0; L < 32; ++L) {
          const int h = (cur_hb * TPB + warp_t0 + L) * KH + k;
          if (h >= H) break;  // warp-uniform
          const int c = plane_recount(tile, len, recs + h, Pp, thr_up, 0);
          if (lane == L) cnt[k] += c;
        }
      }
    }
    item = next;
    ++it;
  }
  asm volatile("cp.async.wait_group 0;");
  flush();
  if (stats && lane == 0) {
    atomicAdd(stats + 0, (unsigned long long)it * 32ull * KH);
    if (n_redo) atomicAdd(stats + 1, (unsigned long long)n_redo);
  }
}

// =====================================================================================
// K4: earliest arg-max over the scored hypotheses (ALL_H stop rule)
// =====================================================================================
__global__ void winner_kernel(const int* __restrict__ counts, const uint8_t* __restrict__ flags, int H,
                              const float* __restrict__ coeffs8, int* __restrict__ best /*[2]: idx,count*/,
                              float* __restrict__ best_coeffs /*8*/) {
  __shared__ int s_c[32], s_i[32];
  int bc = INT_MIN, bi = INT_MAX;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    if (!(flags[h] & 1)) continue;  // computeModelCoefficients failed: skipped by PCL
    int c = counts[h];
    if (c > bc || (c == bc && h < bi)) { bc = c; bi = h; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    int oc = __shfl_down_sync(0xffffffffu, bc, o), oi = __shfl_down_sync(0xffffffffu, bi, o);
    if (oc > bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
  }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_c[w] = bc; s_i[w] = bi; }
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    bc = (l < nw) ? s_c[l] : INT_MIN;
    bi = (l < nw) ? s_i[l] : INT_MAX;
    for (int o = 16; o > 0; o >>= 1) {
      int oc = __shfl_down_sync(0xffffffffu, bc, o), oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (oc > bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
    }
    if (l == 0) {
      if (bi == INT_MAX) { best[0] = -1; best[1] = 0; }
      else { best[0] = bi; best[1] = bc; }
    }
    __syncwarp();
    int idx = __shfl_sync(0xffffffffu, (l == 0) ? ((bi == INT_MAX) ? -1 : bi) : 0, 0);
    if (l < 8) best_coeffs[l] = (idx >= 0 && coeffs8) ? coeffs8[(size_t)idx * 8 + l] : 0.0f;
  }
}

// =====================================================================================
// K5: selectWithinDistance — predicate + ordered (ascending index) stream compaction.
// pass 1: per-block inlier counts; pass 2: exclusive scan of the block counts (one block);
// pass 3: predicate again + ballot ranks -> indices. The predicate is evaluated twice instead of
// materialising N flags.
// =====================================================================================
constexpr int SEL_TPB = 256;
constexpr int SEL_PPT = 4;  // consecutive points per thread
template <int MODEL>
__device__ __forceinline__ unsigned select_mask(const float4* xyz, const float4* nrm, int n, const RecRegs<MODEL>& r,
                                                const ScoreParams& sp, int first) {
  constexpr bool NEED_N = (MODEL == PITT_MODEL_CYLINDER || MODEL == PITT_MODEL_CONE);
  unsigned m = 0;
#pragma unroll
  for (int j = 0; j < SEL_PPT; ++j) {
    int i = first + j;
    if (i < n) {
      f3 pt = ld3(xyz, i);
      f3 nv = NEED_N ? ld3(nrm, i) : mk3(0.f, 0.f, 0.f);
      if (r.inlier(pt, nv, sp)) m |= (1u << j);
    }
  }
  return m;
}
template <int MODEL>
__global__ void __launch_bounds__(SEL_TPB)
select_count_kernel(const float4* __restrict__ xyz, const float4* __restrict__ nrm, int n, const HypRec* __restrict__ rec,
                    ScoreParams sp, int* __restrict__ block_counts) {
  __shared__ int s_w[SEL_TPB / 32];
  RecRegs<MODEL> r;
  r.load(rec);
  int first = (blockIdx.x * SEL_TPB + threadIdx.x) * SEL_PPT;
  int c = __popc(select_mask<MODEL>(xyz, nrm, n, r, sp, first));
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < SEL_TPB / 32; ++i) t += s_w[i];
    block_counts[blockIdx.x] = t;
  }
}
// exclusive scan of nb ints in place, total -> *total (single block, any nb)
__global__ void scan_blocks_kernel(int* __restrict__ v, int nb, int* __restrict__ total) {
  __shared__ int s_carry;
  __shared__ int s_w[32];
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += blockDim.x) {
    int i = base + threadIdx.x;
    int x = (i < nb) ? v[i] : 0;
    int incl = x;
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += y;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int nw = blockDim.x >> 5;
      int wv = (threadIdx.x < nw) ? s_w[threadIdx.x] : 0;
      int wi = wv;
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, wi, o);
        if (threadIdx.x >= o) wi += y;
      }
      s_w[threadIdx.x] = wi - wv;  // exclusive warp offsets
    }
    __syncthreads();
    int carry = s_carry;
    int excl = carry + s_w[threadIdx.x >> 5] + incl - x;
    if (i < nb) v[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = excl + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}
template <int MODEL>
__global__ void __launch_bounds__(SEL_TPB)
select_write_kernel(const float4* __restrict__ xyz, const float4* __restrict__ nrm, int n, const HypRec* __restrict__ rec,
                    ScoreParams sp, const int* __restrict__ block_offsets, int* __restrict__ out) {
  __shared__ int s_w[SEL_TPB / 32];
  RecRegs<MODEL> r;
  r.load(rec);
  int first = (blockIdx.x * SEL_TPB + threadIdx.x) * SEL_PPT;
  unsigned m = select_mask<MODEL>(xyz, nrm, n, r, sp, first);
  int c = __popc(m);
  int incl = c;
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += y;
  }
  if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
  __syncthreads();
  int woff = 0;
  for (int i = 0; i < (threadIdx.x >> 5); ++i) woff += s_w[i];
  int pos = block_offsets[blockIdx.x] + woff + incl - c;
#pragma unroll
  for (int j = 0; j < SEL_PPT; ++j)
    if (m & (1u << j)) out[pos++] = first + j;
}

// Small clouds (the object clusters of a frame, <= SEL_SMALL_MAX points): record preparation, predicate and ordered
// compaction in ONE single-CTA launch instead of four (prep_rec, count, scan, write). A frame makes 24 such selections.
constexpr int SEL_SMALL_TPB = 1024;
constexpr int SEL_SMALL_MAX = 16384;
template <int MODEL>
__global__ void __launch_bounds__(SEL_SMALL_TPB)
select_small_kernel(const float4* __restrict__ xyz, const float4* __restrict__ nrm, int n, const float* __restrict__ coeffs, Limits L,
                    ScoreParams sp, int* __restrict__ out, int* __restrict__ total, const FitDesc* __restrict__ D = nullptr,
                    int refined = 0) {
  constexpr bool NEED_N = (MODEL == PITT_MODEL_CYLINDER || MODEL == PITT_MODEL_CONE);
  if (D) {  // batched: problem blockIdx.x; refined = 0: inliers of the winner (count -> ints[2]), 1: of the refined model (-> ints[3])
    const FitDesc d = D[blockIdx.x];
    xyz = d.xyz; nrm = d.nrm; n = d.n;
    coeffs = refined ? d.flt + 8 : d.flt;
    out = d.inl;
    total = d.ints + (refined ? 3 : 2);
  }
  __shared__ HypRec s_rec;
  __shared__ int s_w[SEL_SMALL_TPB / 32];
  __shared__ int s_base, s_tile_total;
  static_assert(SEL_SMALL_TPB == 1024, "warp 0 scans exactly 32 warp counts");
  if (threadIdx.x == 0) {
    float mc[8];
    for (int i = 0; i < 8; ++i) mc[i] = coeffs[i];
    const bool valid = model_valid<MODEL>(L, mc);
    make_rec<MODEL>(mc, valid, sp, s_rec);
    s_base = 0;
  }
  __syncthreads();
  RecRegs<MODEL> r;
  r.load(&s_rec);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int tile = 0; tile < n; tile += SEL_SMALL_TPB) {
    const int i = tile + threadIdx.x;
    bool in = false;
    if (i < n) {
      const f3 pt = ld3(xyz, i);
      const f3 nv = NEED_N ? ld3(nrm, i) : mk3(0.f, 0.f, 0.f);
      in = r.inlier(pt, nv, sp);
    }
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {  // warp 0 scans the 32 warp counts
      const int v = s_w[lane];
      int incl = v;
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      s_w[lane] = incl - v;  // exclusive offsets
      if (lane == 31) s_tile_total = incl;
    }
    __syncthreads();
    const int base = s_base;
    if (in) out[base + s_w[warp] + __popc(m & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + s_tile_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_base;
}

// =====================================================================================
// K6: plane refinement = computeMeanAndCovarianceMatrix over the winner's inliers + eigen33.
// The nine sums are accumulated in double by a fixed-shape tree (deterministic) and rounded once
// to float — the oracle defines them as exact sums rounded once (oracle/orc_math.h DD).
// =====================================================================================
__global__ void __launch_bounds__(REF_TPB)
plane_sums_kernel(const float4* __restrict__ xyz, int n, const HypRec* __restrict__ rec, const int* __restrict__ idx,
                  const int* __restrict__ n_idx, ScoreParams sp, double* __restrict__ partial /*[blocks][10]*/,
                  const FitDesc* __restrict__ D = nullptr) {
  // two modes: idx != nullptr sums the listed points; otherwise the predicate of *rec decides
  __shared__ double s_red[REF_TPB / 32][10];
  if (D) {  // batched: problem blockIdx.y, predicate mode with the record prep_rec_kernel left in recs[0]
    const FitDesc d = D[blockIdx.y];
    xyz = d.xyz; n = d.n; rec = d.recs; idx = nullptr; n_idx = nullptr; partial = d.partial;
  }
  RecRegs<PITT_MODEL_PLANE> r;
  if (rec) r.load(rec);
  double a[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) a[k] = 0.0;
  int total = idx ? *n_idx : n;
  for (int i = blockIdx.x * REF_TPB + threadIdx.x; i < total; i += gridDim.x * REF_TPB) {
    int pi = idx ? idx[i] : i;
    f3 p = ld3(xyz, pi);
    bool in = idx ? true : r.inlier(p, p, sp);
    if (in) {
      double x = p.x, y = p.y, z = p.z;
      a[0] += x * x; a[1] += x * y; a[2] += x * z; a[3] += y * y; a[4] += y * z; a[5] += z * z;
      a[6] += x; a[7] += y; a[8] += z; a[9] += 1.0;
    }
  }
#pragma unroll
  for (int k = 0; k < 10; ++k)
    for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_down_sync(0xffffffffu, a[k], o);
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 10; ++k) s_red[threadIdx.x >> 5][k] = a[k];
  __syncthreads();
  if (threadIdx.x < 10) {
    double t = 0.0;
    for (int w = 0; w < REF_TPB / 32; ++w) t += s_red[w][threadIdx.x];
    partial[(size_t)blockIdx.x * 10 + threadIdx.x] = t;
  }
}
__global__ void plane_refine_final_kernel(const double* __restrict__ partial, int blocks, const float* __restrict__ model,
                                          float* __restrict__ refined, int* __restrict__ n_model_inliers,
                                          const FitDesc* __restrict__ D = nullptr) {
  __shared__ double s[10];
  if (D) {  // batched: problem blockIdx.x
    const FitDesc d = D[blockIdx.x];
    partial = d.partial; model = d.flt; refined = d.flt + 8; n_model_inliers = d.ints + 2;
  }
  __shared__ double s_part[REF_BLOCKS * 10];
  // all threads fetch the partial sums at once (one L2 round trip instead of `blocks` dependent ones); the additions keep
  // their order (block 0, 1, 2, ...), so the sums are bit-identical to the sequential loop
  for (int i = threadIdx.x; i < blocks * 10; i += blockDim.x) s_part[i] = partial[i];
  __syncthreads();
  if (threadIdx.x < 10) {
    double t = 0.0;
    for (int b = 0; b < blocks; ++b) t += s_part[b * 10 + threadIdx.x];
    s[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int cnt = (int)s[9];
    *n_model_inliers = cnt;
    if (cnt < 4) {
      for (int i = 0; i < 8; ++i) refined[i] = model[i];
    } else {
      float accu[9], cov[9], cen[3], ev, evec[3];
      for (int i = 0; i < 9; ++i) accu[i] = (float)s[i];
      cov_from_accu(accu, (float)cnt, cov, cen);
      eigen33(cov, ev, evec);
      f3 o = mk3(evec[0], evec[1], evec[2]);
      refined[0] = o.x; refined[1] = o.y; refined[2] = o.z;
      refined[3] = -dot0(o, mk3(cen[0], cen[1], cen[2]));
      for (int i = 4; i < 8; ++i) refined[i] = 0.0f;
    }
  }
}

// =====================================================================================
// FP32 pipe micro-benchmarks (roofline denominators of K3)
// =====================================================================================
template <int KIND>
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float seed, float one_rt) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + (float)(threadIdx.x + i) * 1e-3f;
  const float m = 1.0000001f, c = 1e-7f;
  if (KIND == 2) {
    unsigned long long v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = pack2(a[2 * i], a[2 * i + 1]);
    unsigned long long M = pack2(m, m), Cc = pack2(c, c), ONE = pack2(one_rt, one_rt);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fadd2_exact(mul2(v[i], M), Cc, ONE);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[2 * i] = lo32(v[i]); a[2 * i + 1] = hi32(v[i]); }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (KIND == 0) a[i] = __fmaf_rn(a[i], m, c);
        else a[i] = __fadd_rn(__fmul_rn(a[i], m), c);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// =====================================================================================
// host side
// =====================================================================================
static inline int sample_size(int model) {
  return model == PITT_MODEL_PLANE ? 3 : model == PITT_MODEL_SPHERE ? 4 : model == PITT_MODEL_CYLINDER ? 2 : 3;
}
static inline int coeff_count(int model) { return (model == PITT_MODEL_PLANE || model == PITT_MODEL_SPHERE) ? 4 : 7; }

Limits limits_for(const pitt_sac_params& p) {
  Limits L;
  L.radius_min = -DBL_MAX; L.radius_max = DBL_MAX; L.min_angle = -DBL_MAX; L.max_angle = DBL_MAX;
  L.eps_angle = 0.0; L.w = 0.0; L.ax = L.ay = L.az = 0.0f;
  const bool fwd_radius = (p.radius_min != -DBL_MAX) && (p.radius_max != DBL_MAX);
  const bool has_axis = (p.axis[0] != 0.0f || p.axis[1] != 0.0f || p.axis[2] != 0.0f);
  if (p.model == PITT_MODEL_SPHERE) {
    if (fwd_radius) { L.radius_min = p.radius_min; L.radius_max = p.radius_max; }
  } else if (p.model == PITT_MODEL_CYLINDER || p.model == PITT_MODEL_CONE) {
    if (p.model == PITT_MODEL_CYLINDER && fwd_radius) { L.radius_min = p.radius_min; L.radius_max = p.radius_max; }
    L.w = p.normal_distance_weight;
    if (has_axis) { L.ax = p.axis[0]; L.ay = p.axis[1]; L.az = p.axis[2]; }
    if (p.eps_angle != 0.0) L.eps_angle = p.eps_angle;
    if (p.model == PITT_MODEL_CONE && p.min_angle != -DBL_MAX && p.max_angle != DBL_MAX) {
      L.min_angle = p.min_angle; L.max_angle = p.max_angle;
    }
  }
  return L;
}

ScoreParams score_params_for(const pitt_sac_params& p, const Limits& L) {
  ScoreParams sp;
  sp.thr = p.distance_threshold;
  float t = (float)p.distance_threshold;
  if ((double)t < p.distance_threshold) t = nextafterf(t, INFINITY);
  sp.thr_up = t;
  sp.w = L.w;
  // |FP32 score - exact score| bound: acosf near |cos|=1 loses up to ~1e-3 rad (times w), the
  // euclidean part a few ulps of the point-axis distance (clouds within ~8 m of the axis).
  sp.band = (float)(2e-3 * fabs(L.w) + 2e-6);
  return sp;
}

#define PITT_LAUNCH_CHECK(ctx, what)                                       \
  do {                                                                     \
    (ctx)->launches++;                                                     \
    cudaError_t e__ = cudaGetLastError();                                  \
    if (e__ != cudaSuccess) return fail((ctx), PITT_ERR_CUDA, what, e__);  \
  } while (0)

template <int MODEL>
static int launch_estimate(pitt_ctx* ctx, const pitt_cloud* c, const int* d_samples, int H, const Limits& L, const ScoreParams& sp,
                           HypRec* d_recs, float* d_coeffs8, uint8_t* d_flags) {
  estimate_kernel<MODEL><<<cdiv(H, 128), 128, 0, ctx->stream>>>(c->d_xyz, c->d_nrm, d_samples, H, L, sp, d_recs, d_coeffs8, d_flags,
                                                                  ctx->skip_flag);
  PITT_LAUNCH_CHECK(ctx, "estimate_kernel");
  return PITT_OK;
}

int sac_estimate(pitt_ctx* ctx, const pitt_cloud* c, int model, const int* d_samples, int H, const Limits& L, const ScoreParams& sp,
                 HypRec* d_recs, float* d_coeffs8, uint8_t* d_flags) {
  if (H <= 0) return PITT_OK;
  switch (model) {
    case PITT_MODEL_PLANE: return launch_estimate<PITT_MODEL_PLANE>(ctx, c, d_samples, H, L, sp, d_recs, d_coeffs8, d_flags);
    case PITT_MODEL_SPHERE: return launch_estimate<PITT_MODEL_SPHERE>(ctx, c, d_samples, H, L, sp, d_recs, d_coeffs8, d_flags);
    case PITT_MODEL_CYLINDER: return launch_estimate<PITT_MODEL_CYLINDER>(ctx, c, d_samples, H, L, sp, d_recs, d_coeffs8, d_flags);
    default: return launch_estimate<PITT_MODEL_CONE>(ctx, c, d_samples, H, L, sp, d_recs, d_coeffs8, d_flags);
  }
}

template <int MODEL, int P, bool TWO_TIER>
static int launch_score_generic(pitt_ctx* ctx, const pitt_cloud* c, const HypRec* d_recs, int H, const ScoreParams& sp,
                                int* d_counts) {
  constexpr int TPB = 256;
  const int n = c->n;
  // Work split: big clouds are cut along the points (each CTA keeps up to 512 hypotheses in shared
  // memory and streams its point tiles); small clouds (object clusters: a few thousand points,
  // 1001 hypotheses) have too few point tiles to fill 148 SMs and are cut along the hypotheses.
  const int tile = TPB * P;
  const int tiles = cdiv(n, tile);
  const int want_ctas = ctx->sm_count * 4;
  int hyp_chunk = H < 512 ? H : 512;
  // Big jobs: many small CTAs (128 hypotheses x one point range), so that the last, partially filled wave of CTAs is a
  // small fraction of the run (600 CTAs on 296 slots ran as 3 waves for 2.03 waves of work).
  if (TWO_TIER && (long long)tiles * cdiv(H, 128) >= 6LL * want_ctas) hyp_chunk = 128;
  if (tiles * cdiv(H, hyp_chunk) < want_ctas) {
    int chunks_wanted = cdiv(want_ctas, tiles);
    hyp_chunk = std::max(16, cdiv(H, chunks_wanted));
    if (hyp_chunk > 512) hyp_chunk = 512;
    if (hyp_chunk > H) hyp_chunk = H;
  }
  int n_chunks = cdiv(H, hyp_chunk);
  int pblocks = (ctx->sm_count * 8) / n_chunks;
  if (pblocks < 1) pblocks = 1;
  if (pblocks > tiles) pblocks = tiles;
  int pts_per_cta = cdiv(tiles, pblocks) * tile;
  pblocks = cdiv(n, pts_per_cta);
  size_t smem = (size_t)hyp_chunk * (sizeof(HypRec) + sizeof(int));
  dim3 grid(pblocks, n_chunks);
  if constexpr (TWO_TIER) {
    smem += 8 + (size_t)(TPB / 32) * (32 * P + 32) * sizeof(int2);  // per-warp queues of undecided evaluations
    score2_kernel<MODEL, P, TPB><<<grid, TPB, smem, ctx->stream>>>(c->d_xyz, c->d_nrm, n, d_recs, H, hyp_chunk, pts_per_cta, sp, d_counts,
                                                                   ctx->skip_flag);
    PITT_LAUNCH_CHECK(ctx, "score2_kernel");
  } else {
    score_kernel<MODEL, P, TPB><<<grid, TPB, smem, ctx->stream>>>(c->d_xyz, c->d_nrm, n, d_recs, H, hyp_chunk, pts_per_cta, sp, d_counts,
                                                                  ctx->skip_flag);
    PITT_LAUNCH_CHECK(ctx, "score_kernel");
  }
  return PITT_OK;
}

int g_force_generic_plane = 0;  // test hook: 1 routes plane scoring through the generic kernel
int g_select_no_fuse = 0;       // test hook: 1 keeps small selections on the four-launch path
int g_score_mode = 0;  // test hook (2 / 3: sphere packed kernel with 512- / 128-point tiles): 0 two-tier kernel for cylinder/cone, 1 generic score_kernel for every model
int g_plane_mode = 0;  // test hook: 0 automatic (tensor path on large jobs), 1 exact packed kernel only,
                       // 2 FFMA filter + exact re-evaluation always, 3 tensor-core path (plane_tc.cu) always
int launch_score_plane_tc(pitt_ctx* ctx, const pitt_cloud* c, const HypRec* d_recs, int H, const ScoreParams& sp, int* d_counts,
                          const int** d_use_out, const float4* d_extra, int n_extra, const int* d_ready, int ready_pts);
unsigned long long g_plane_filter_stats[2] = {0, 0};  // last call with stats enabled: pairs, re-evaluated pairs
int g_plane_filter_collect_stats = 0;

// tensor path on large jobs? (shared with pitt_sac_segment_host, which plans its copy accordingly)
bool plane_job_takes_tensor_path(int n, int H) {
  return H >= 256 && !g_force_generic_plane &&
         ((g_plane_mode == 3) || (g_plane_mode == 0 && (double)n * (double)H >= 134217728.0));
}

// d_ready != NULL: c is the WHOLE cloud, still arriving; the tensor kernel polls the flags, every other kernel here waits
// for ev_all (the last chunk) first.
static int launch_score_plane_packed(pitt_ctx* ctx, const pitt_cloud* c, const HypRec* d_recs, int H, const ScoreParams& sp,
                                     int* d_counts, const float4* d_extra = nullptr, int n_extra = 0, const int* d_ready = nullptr,
                                     int ready_pts = 0, cudaEvent_t ev_all = nullptr) {
  constexpr int KH = 8, TPB = 128, TILE = 512;
  const int n = c->n;
  const int hblocks = cdiv(H, KH * TPB);
  const int n_ptiles = cdiv(n, TILE);
  const long long items = (long long)hblocks * n_ptiles;
  if (items > INT_MAX) return fail(ctx, PITT_ERR_INVALID, "plane scoring: too many work items");
  // scratch: [0..2] |x|,|y|,|z| maxima (as words), [3] exact-kernel work counter, [4] filter work counter,
  // then the filter parameters and the statistics
  unsigned* d_scr = nullptr;
  PlaneFilterParams* d_P = nullptr;
  unsigned long long* d_stats = nullptr;
  PITT_TRY(arena_alloc(ctx, 8, &d_scr));
  PITT_TRY(arena_alloc(ctx, 1, &d_P));
  PITT_TRY(arena_alloc(ctx, 2, &d_stats));
  PITT_CUDA(ctx, cudaMemsetAsync(d_scr, 0, 8 * sizeof(unsigned), ctx->stream));
  PITT_CUDA(ctx, cudaMemsetAsync(d_P, 0, sizeof(PlaneFilterParams), ctx->stream));
  // the filter's two set-up launches only pay off on large jobs; tests force it with mode 2
  // large jobs: dot products on the tensor cores (plane_tc.cu); the FFMA filter (mode 2) is the CUDA-core alternative
  const bool tensor = (g_plane_mode == 3) || (g_plane_mode == 0 && (double)n * (double)H >= 134217728.0);
  const bool filter = (g_plane_mode == 2);
  const int* d_skip = ctx->skip_flag;  // device flag: non-zero = a fast kernel did the work (or the stop rule has ended the loop), the exact kernel returns at once
  if (d_ready && !tensor) return fail(ctx, PITT_ERR_STATE, "streaming scoring needs the tensor path");
  if (tensor) PITT_TRY(launch_score_plane_tc(ctx, c, d_recs, H, sp, d_counts, &d_skip, d_extra, n_extra, d_ready, ready_pts));
  if (d_ready) PITT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_all, 0));  // the exact kernel (fallback) reads the whole cloud
  if (filter) {
    if (d_extra) return fail(ctx, PITT_ERR_STATE, "plane filter mode does not take a chunked cloud");
    d_skip = &d_P->use_filter;
    PITT_CUDA(ctx, cudaMemsetAsync(d_stats, 0, 2 * sizeof(unsigned long long), ctx->stream));
    int ab = std::min(cdiv(n, 256 * 8), ctx->sm_count * 8);
    cloud_absmax_kernel<<<ab, 256, 0, ctx->stream>>>(c->d_xyz, n, d_scr);
    PITT_LAUNCH_CHECK(ctx, "cloud_absmax_kernel");
    plane_filter_params_kernel<<<1, 1, 0, ctx->stream>>>(d_scr, sp.thr_up, d_P);
    PITT_LAUNCH_CHECK(ctx, "plane_filter_params_kernel");
    static int ctas_f = 0;
    if (!ctas_f) {
      PITT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_f, plane_filter_kernel<KH, TPB, TILE>, TPB, 0));
      if (ctas_f < 1) ctas_f = 1;
    }
    int grid = ctx->sm_count * ctas_f;
    if ((long long)grid > items) grid = (int)items;
    plane_filter_kernel<KH, TPB, TILE><<<grid, TPB, 0, ctx->stream>>>(c->d_xyz, n, d_recs, H, n_ptiles, (int)items, sp.thr_up, d_P,
                                                                     (int*)d_scr + 4, d_counts,
                                                                     g_plane_filter_collect_stats ? d_stats : nullptr);
    PITT_LAUNCH_CHECK(ctx, "plane_filter_kernel");
  }
  static int ctas_e = 0;
  if (!ctas_e) {
    PITT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_e, plane_score_kernel<KH, TPB, TILE>, TPB, 0));
    if (ctas_e < 1) ctas_e = 1;
  }
  int grid = ctx->sm_count * ctas_e;
  if ((long long)grid > items) grid = (int)items;
  // exact kernel: the whole job in exact mode; returns at once when the filter kernel was eligible
  plane_score_kernel<KH, TPB, TILE><<<grid, TPB, 0, ctx->stream>>>(c->d_xyz, n, d_recs, H, n_ptiles, (int)items, sp.thr_up, 1.0f,
                                                                  (int*)d_scr + 3, d_counts, d_skip);
  PITT_LAUNCH_CHECK(ctx, "plane_score_kernel");
  if (filter && g_plane_filter_collect_stats) {
    PITT_CUDA(ctx, cudaMemcpyAsync(g_plane_filter_stats, d_stats, sizeof(g_plane_filter_stats), cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
  }
  return PITT_OK;
}


template <int TILE>
static int launch_score_sphere_packed_t(pitt_ctx* ctx, const pitt_cloud* c, const HypRec* d_recs, int H, int* d_counts) {
  constexpr int KH = 8, TPB = 128;
  const int n = c->n;
  const int hblocks = cdiv(H, KH * TPB);
  const int n_ptiles = cdiv(n, TILE);
  const long long items = (long long)hblocks * n_ptiles;
  if (items > INT_MAX) return fail(ctx, PITT_ERR_INVALID, "sphere scoring: too many work items");
  int* d_work = nullptr;
  PITT_TRY(arena_alloc(ctx, 1, &d_work));
  PITT_CUDA(ctx, cudaMemsetAsync(d_work, 0, sizeof(int), ctx->stream));
  static int ctas = 0;
  if (!ctas) {
    PITT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, sphere_score_kernel<KH, TPB, TILE>, TPB, 0));
    if (ctas < 1) ctas = 1;
  }
  int grid = ctx->sm_count * ctas;
  if ((long long)grid > items) grid = (int)items;
  sphere_score_kernel<KH, TPB, TILE><<<grid, TPB, 0, ctx->stream>>>(c->d_xyz, n, d_recs, H, n_ptiles, (int)items, 1.0f, d_work, d_counts);
  PITT_LAUNCH_CHECK(ctx, "sphere_score_kernel");
  return PITT_OK;
}
// Tile size (measured, 10 000 hypotheses): jobs of at most ~4 items per SM run as one partial wave of 512-point items
// (20 000 points: 0.099 ms against 0.116); beyond that 128-point items, several per persistent CTA with the next tile
// prefetched, leave a smaller last round (50 000 points: 0.218 against 0.231; 500 000 points: no difference)
static int launch_score_sphere_packed(pitt_ctx* ctx, const pitt_cloud* c, const HypRec* d_recs, int H, int* d_counts) {
  const long long items512 = (long long)cdiv(H, 1024) * cdiv(c->n, 512);
  const bool small_tiles = g_score_mode == 3 || (g_score_mode != 2 && items512 > 4LL * ctx->sm_count);
  if (small_tiles) return launch_score_sphere_packed_t<128>(ctx, c, d_recs, H, d_counts);
  return launch_score_sphere_packed_t<512>(ctx, c, d_recs, H, d_counts);
}

// a consumer that needs the whole cloud at once
static int stream_complete(pitt_ctx* ctx, const pitt_cloud* c) {
  for (int k = 0; k < c->stream_chunks; ++k) PITT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[k], 0));
  c->stream_chunks = 0;
  return PITT_OK;
}

int sac_score(pitt_ctx* ctx, const pitt_cloud* c, int model, const HypRec* d_recs, int H, const ScoreParams& sp, int* d_counts);

// Plane scoring of a cloud that is still arriving (pitt_sac_segment_host): chunk k is scored as soon as its copy has
// completed, while the copy engine brings in chunk k + 1. Counts add up over the chunks. d_gather = the sample points
// (every hypothesis passes through three of them): with the chunk's own points they bound the tensor path's scale.
int sac_score_plane_streaming(pitt_ctx* ctx, const pitt_cloud* c, const HypRec* d_recs, int H, const ScoreParams& sp, int* d_counts,
                              const float4* d_gather, int n_gather) {
  PITT_CUDA(ctx, cudaMemsetAsync(d_counts, 0, (size_t)H * sizeof(int), ctx->stream));
  if (c->d_ready && !plane_job_takes_tensor_path(c->n, H)) {  // fewer hypotheses than planned: wait for the cloud, score normally
    PITT_TRY(stream_complete(ctx, c));
    return sac_score(ctx, c, PITT_MODEL_PLANE, d_recs, H, sp, d_counts);
  }
  if (c->d_ready) {
    // one launch: the persistent tensor kernel consumes the chunks as their flags come up
    pitt_cloud view;
    view.n = c->n;
    view.d_xyz = c->d_xyz;
    PITT_TRY(launch_score_plane_packed(ctx, &view, d_recs, H, sp, d_counts, d_gather, n_gather, c->d_ready, c->stream_off[1],
                                       ctx->ev_chunk[c->stream_chunks - 1]));
    c->stream_chunks = 0;
    return PITT_OK;
  }
  for (int k = 0; k < c->stream_chunks; ++k) {
    PITT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[k], 0));
    const int off = c->stream_off[k];
    pitt_cloud view;
    view.n = c->stream_off[k + 1] - off;
    view.d_xyz = c->d_xyz + off;
    if (view.n <= 0) continue;
    if (H >= 256 && !g_force_generic_plane) PITT_TRY(launch_score_plane_packed(ctx, &view, d_recs, H, sp, d_counts, d_gather, n_gather));
    else PITT_TRY((launch_score_generic<PITT_MODEL_PLANE, 8, false>(ctx, &view, d_recs, H, sp, d_counts)));
  }
  c->stream_chunks = 0;  // everything this stream does from here on is ordered after the last chunk
  return PITT_OK;
}


int sac_score(pitt_ctx* ctx, const pitt_cloud* c, int model, const HypRec* d_recs, int H, const ScoreParams& sp,
              int* d_counts) {
  if (H <= 0) return PITT_OK;
  PITT_CUDA(ctx, cudaMemsetAsync(d_counts, 0, (size_t)H * sizeof(int), ctx->stream));
  if (c->n <= 0) return PITT_OK;
  switch (model) {
    case PITT_MODEL_PLANE:
      if (H >= 256 && !g_force_generic_plane) return launch_score_plane_packed(ctx, c, d_recs, H, sp, d_counts);
      return launch_score_generic<PITT_MODEL_PLANE, 8, false>(ctx, c, d_recs, H, sp, d_counts);
    case PITT_MODEL_SPHERE:
      // large jobs (C3 clusters x 10 000 hypotheses): lanes = hypotheses, no per-(warp, hypothesis) reduction; below 1e8
      // evaluations the generic kernel is as fast (5 000 x 10 000: 0.048 ms against 0.051)
      if (g_score_mode >= 2 || (g_score_mode == 0 && H >= 512 && (double)c->n * (double)H >= 1.0e8))
        return launch_score_sphere_packed(ctx, c, d_recs, H, d_counts);
      return launch_score_generic<PITT_MODEL_SPHERE, 8, false>(ctx, c, d_recs, H, sp, d_counts);
    case PITT_MODEL_CYLINDER:
      if (g_score_mode == 1) return launch_score_generic<PITT_MODEL_CYLINDER, 2, false>(ctx, c, d_recs, H, sp, d_counts);
      if (c->n < 4096) return launch_score_generic<PITT_MODEL_CYLINDER, 1, true>(ctx, c, d_recs, H, sp, d_counts);
      if (c->n < 32768) return launch_score_generic<PITT_MODEL_CYLINDER, 4, true>(ctx, c, d_recs, H, sp, d_counts);
      return launch_score_generic<PITT_MODEL_CYLINDER, 8, true>(ctx, c, d_recs, H, sp, d_counts);
    default:
      if (g_score_mode == 1) return launch_score_generic<PITT_MODEL_CONE, 2, false>(ctx, c, d_recs, H, sp, d_counts);
      if (c->n < 4096) return launch_score_generic<PITT_MODEL_CONE, 1, true>(ctx, c, d_recs, H, sp, d_counts);
      return launch_score_generic<PITT_MODEL_CONE, 4, true>(ctx, c, d_recs, H, sp, d_counts);
  }
}

template <int MODEL>
static int launch_select(pitt_ctx* ctx, const pitt_cloud* c, const HypRec* d_rec, const ScoreParams& sp, int* d_out,
                         int* d_total) {
  const int n = c->n;
  const int nb = cdiv(n, SEL_TPB * SEL_PPT);
  int* d_bc = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)nb + 1, &d_bc));
  select_count_kernel<MODEL><<<nb, SEL_TPB, 0, ctx->stream>>>(c->d_xyz, c->d_nrm, n, d_rec, sp, d_bc);
  PITT_LAUNCH_CHECK(ctx, "select_count_kernel");
  scan_blocks_kernel<<<1, 1024, 0, ctx->stream>>>(d_bc, nb, d_total);
  PITT_LAUNCH_CHECK(ctx, "scan_blocks_kernel");
  select_write_kernel<MODEL><<<nb, SEL_TPB, 0, ctx->stream>>>(c->d_xyz, c->d_nrm, n, d_rec, sp, d_bc, d_out);
  PITT_LAUNCH_CHECK(ctx, "select_write_kernel");
  return PITT_OK;
}

// d_coeffs: 8 floats on the device. Writes ascending inlier indices to d_out (capacity n) and
// their number to d_total. isModelValid is applied (invalid model => 0 inliers).
int sac_select(pitt_ctx* ctx, const pitt_cloud* c, int model, const float* d_coeffs, const Limits& L,
               const ScoreParams& sp, int* d_out, int* d_total) {
  if (c->n <= 0) {
    PITT_CUDA(ctx, cudaMemsetAsync(d_total, 0, sizeof(int), ctx->stream));
    return PITT_OK;
  }
  if (c->n <= SEL_SMALL_MAX && !g_select_no_fuse) {
    switch (model) {
      case PITT_MODEL_PLANE:
        select_small_kernel<PITT_MODEL_PLANE><<<1, SEL_SMALL_TPB, 0, ctx->stream>>>(c->d_xyz, c->d_nrm, c->n, d_coeffs, L, sp, d_out, d_total);
        break;
      case PITT_MODEL_SPHERE:
        select_small_kernel<PITT_MODEL_SPHERE><<<1, SEL_SMALL_TPB, 0, ctx->stream>>>(c->d_xyz, c->d_nrm, c->n, d_coeffs, L, sp, d_out, d_total);
        break;
      case PITT_MODEL_CYLINDER:
        select_small_kernel<PITT_MODEL_CYLINDER><<<1, SEL_SMALL_TPB, 0, ctx->stream>>>(c->d_xyz, c->d_nrm, c->n, d_coeffs, L, sp, d_out, d_total);
        break;
      default:
        select_small_kernel<PITT_MODEL_CONE><<<1, SEL_SMALL_TPB, 0, ctx->stream>>>(c->d_xyz, c->d_nrm, c->n, d_coeffs, L, sp, d_out, d_total);
        break;
    }
    PITT_LAUNCH_CHECK(ctx, "select_small_kernel");
    return PITT_OK;
  }
  HypRec* d_rec = nullptr;
  PITT_TRY(arena_alloc(ctx, 1, &d_rec));
  switch (model) {
    case PITT_MODEL_PLANE:
      prep_rec_kernel<PITT_MODEL_PLANE><<<1, 1, 0, ctx->stream>>>(d_coeffs, L, sp, d_rec);
      PITT_LAUNCH_CHECK(ctx, "prep_rec_kernel");
      return launch_select<PITT_MODEL_PLANE>(ctx, c, d_rec, sp, d_out, d_total);
    case PITT_MODEL_SPHERE:
      prep_rec_kernel<PITT_MODEL_SPHERE><<<1, 1, 0, ctx->stream>>>(d_coeffs, L, sp, d_rec);
      PITT_LAUNCH_CHECK(ctx, "prep_rec_kernel");
      return launch_select<PITT_MODEL_SPHERE>(ctx, c, d_rec, sp, d_out, d_total);
    case PITT_MODEL_CYLINDER:
      prep_rec_kernel<PITT_MODEL_CYLINDER><<<1, 1, 0, ctx->stream>>>(d_coeffs, L, sp, d_rec);
      PITT_LAUNCH_CHECK(ctx, "prep_rec_kernel");
      return launch_select<PITT_MODEL_CYLINDER>(ctx, c, d_rec, sp, d_out, d_total);
    default:
      prep_rec_kernel<PITT_MODEL_CONE><<<1, 1, 0, ctx->stream>>>(d_coeffs, L, sp, d_rec);
      PITT_LAUNCH_CHECK(ctx, "prep_rec_kernel");
      return launch_select<PITT_MODEL_CONE>(ctx, c, d_rec, sp, d_out, d_total);
  }
}

// Plane optimizeModelCoefficients. Either over an explicit inlier list (d_idx, d_n_idx) or, when
// d_idx == nullptr, over the points satisfying the predicate of d_model (fused select + sums).
int plane_refine(pitt_ctx* ctx, const pitt_cloud* c, const float* d_model, const int* d_idx, const int* d_n_idx,
                 const Limits& L, const ScoreParams& sp, float* d_refined, int* d_n_model_inliers) {
  double* d_partial = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)REF_BLOCKS * 10, &d_partial));
  HypRec* d_rec = nullptr;
  if (!d_idx) {
    PITT_TRY(arena_alloc(ctx, 1, &d_rec));
    prep_rec_kernel<PITT_MODEL_PLANE><<<1, 1, 0, ctx->stream>>>(d_model, L, sp, d_rec);
    PITT_LAUNCH_CHECK(ctx, "prep_rec_kernel");
  }
  plane_sums_kernel<<<REF_BLOCKS, REF_TPB, 0, ctx->stream>>>(c->d_xyz, c->n, d_rec, d_idx, d_n_idx, sp, d_partial);
  PITT_LAUNCH_CHECK(ctx, "plane_sums_kernel");
  plane_refine_final_kernel<<<1, 256, 0, ctx->stream>>>(d_partial, REF_BLOCKS, d_model, d_refined, d_n_model_inliers);
  PITT_LAUNCH_CHECK(ctx, "plane_refine_final_kernel");
  return PITT_OK;
}

int sac_winner(pitt_ctx* ctx, const int* d_counts, const uint8_t* d_flags, int H, const float* d_coeffs8, int* d_best,
               float* d_best_coeffs) {
  winner_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, d_flags, H, d_coeffs8, d_best, d_best_coeffs);
  PITT_LAUNCH_CHECK(ctx, "winner_kernel");
  return PITT_OK;
}

int fp32_peak(pitt_ctx* ctx, int kind, double* tflops) {
  const int blocks = ctx->sm_count * 16, tpb = 256, iters = 4096;
  float* d_out = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)blocks * tpb, &d_out));
  cudaEvent_t e0, e1;
  PITT_CUDA(ctx, cudaEventCreate(&e0));
  PITT_CUDA(ctx, cudaEventCreate(&e1));
  float best_ms = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, ctx->stream);
    if (kind == 0) fp32_peak_kernel<0><<<blocks, tpb, 0, ctx->stream>>>(d_out, iters, 1.0f, 1.0f);
    else if (kind == 1) fp32_peak_kernel<1><<<blocks, tpb, 0, ctx->stream>>>(d_out, iters, 1.0f, 1.0f);
    else fp32_peak_kernel<2><<<blocks, tpb, 0, ctx->stream>>>(d_out, iters, 1.0f, 1.0f);
    ctx->launches++;
    cudaEventRecord(e1, ctx->stream);
    PITT_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_ms) best_ms = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  double flops = (double)blocks * tpb * (double)iters * 16.0 * 2.0;
  *tflops = flops / (best_ms * 1e-3) / 1e12;
  return PITT_OK;
}

// ------------------------------------------------------------------ PCL sample stream (host)
// boost::mt19937(12345) + SampleConsensusModel::drawIndexSample, sparse partial Fisher-Yates.
struct Mt19937 {
  uint32_t s[624];
  int idx;
  explicit Mt19937(uint32_t seed) {
    s[0] = seed;
    for (int i = 1; i < 624; ++i) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
  }
  uint32_t next() {
    if (idx >= 624) {
      for (int i = 0; i < 624; ++i) {
        uint32_t y = (s[i] & 0x80000000u) | (s[(i + 1) % 624] & 0x7fffffffu);
        s[i] = s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = s[idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
};

// The shuffled index array of drawIndexSample is kept sparse: positions 0..3 (touched by every draw) in a small array, the
// other touched positions in an open-addressing table (a frame draws 13 streams of 1001 minimal sets: a node-based map cost
// 2 ms of host time per frame, more than every launch of the frame together).
// every stream starts from boost::mt19937(12345): the first draws of the engine are the same for all of them and are kept in
// a table built once per process (>> 1: uniform_int<>(0, INT_MAX) over a 32-bit engine)
constexpr int MT_TABLE = 16384;
static const uint32_t* mt_table() {
  static const std::vector<uint32_t> t = [] {
    std::vector<uint32_t> v(MT_TABLE);
    Mt19937 mt(12345u);
    for (int i = 0; i < MT_TABLE; ++i) v[i] = mt.next() >> 1;
    return v;
  }();
  return t.data();
}
struct PclSampleStream::Impl {
  const uint32_t* table = mt_table();
  int drawn = 0;
  Mt19937* mt = nullptr;  // only a stream that outruns the table runs its own engine
  ~Impl() { delete mt; }
  uint32_t next_r() {
    if (drawn < MT_TABLE) return table[drawn++];
    if (!mt) {
      mt = new Mt19937(12345u);
      for (int i = 0; i < MT_TABLE; ++i) mt->next();
    }
    ++drawn;
    return mt->next() >> 1;
  }
  int n = 0;
  int front[4] = {0, 1, 2, 3};
  static constexpr int DENSE_MAX = 1 << 16;
  std::vector<int> dense, keys, vals;
  int log2_slots = 13, used = 0;

  size_t probe(const std::vector<int>& k, int key, int lg) const {
    size_t h = (size_t)(((uint32_t)key * 0x9E3779B1u) >> (32 - lg));
    const size_t mask = ((size_t)1 << lg) - 1;
    while (k[h] != -1 && k[h] != key) h = (h + 1) & mask;
    return h;
  }
  void grow_table() {
    const int lg = keys.empty() ? 13 : log2_slots + 1;
    std::vector<int> k2((size_t)1 << lg, -1), v2((size_t)1 << lg, 0);
    for (size_t i = 0; i < keys.size(); ++i)
      if (keys[i] != -1) {
        const size_t h = probe(k2, keys[i], lg);
        k2[h] = keys[i];
        v2[h] = vals[i];
      }
    keys.swap(k2);
    vals.swap(v2);
    log2_slots = lg;
  }
  // the entry of position pos (an untouched position holds its own index)
  int* at(int pos) {
    if (pos < 4) return &front[pos];
    if (n <= DENSE_MAX) {  // a cluster: the plain array
      if (dense.empty()) {
        dense.resize((size_t)n);
        for (int i = 0; i < n; ++i) dense[i] = i;
      }
      return &dense[pos];
    }
    if ((size_t)used * 2 >= keys.size()) grow_table();
    const size_t h = probe(keys, pos, log2_slots);
    if (keys[h] == -1) {
      keys[h] = pos;
      vals[h] = pos;
      ++used;
    }
    return &vals[h];
  }
  void draw(int S, int* out) {
    for (int i = 0; i < S; ++i) {
      const uint32_t r = next_r();
      const int j = i + (int)(r % (uint32_t)(n - i));
      int* pj = at(j);
      std::swap(front[i], *pj);
    }
    for (int i = 0; i < S; ++i) out[i] = front[i];
  }
};
PclSampleStream::PclSampleStream(int n, int model, const float* h_xyz4) : impl_(new Impl), model_(model), h_xyz_(h_xyz4) {
  impl_->n = n;
}
PclSampleStream::~PclSampleStream() { delete impl_; }
static bool plane_sample_good(const float* xyz, const int* s) {
  float q[3];
  for (int k = 0; k < 3; ++k) {
    float a = xyz[4 * (size_t)s[1] + k] - xyz[4 * (size_t)s[0] + k];
    float b = xyz[4 * (size_t)s[2] + k] - xyz[4 * (size_t)s[0] + k];
    q[k] = a / b;
  }
  return (q[0] != q[1]) || (q[2] != q[1]);
}
bool PclSampleStream::next(int* out) {
  const int S = sample_size(model_);
  if (impl_->n < S) return false;
  for (int iter = 0; iter < 1000; ++iter) {
    impl_->draw(S, out);
    // isSampleGood: only the plane model tests anything; without a host mirror the draw is
    // speculative and the device's collinearity flag decides (see sac_segment_impl)
    if (model_ != PITT_MODEL_PLANE || !h_xyz_ || plane_sample_good(h_xyz_, out)) return true;
  }
  return false;
}

// Philox-4x32-10 minimal sets drawn on the device (production sampler)
__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
  uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__global__ void philox_samples_kernel(int* __restrict__ samples, int H, int S, int n, uint64_t seed, uint32_t stream_id) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  uint32_t c0 = (uint32_t)h, c1 = stream_id, c2 = 0x9E3779B9u, c3 = 0xBB67AE85u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  uint32_t rnd[4] = {c0, c1, c2, c3};
  int s[4];
  for (int i = 0; i < S; ++i) {
    // i-th draw without replacement: uniform in [0, n-i), then skip over earlier picks (kept sorted)
    int v = (int)(((uint64_t)rnd[i] * (uint64_t)(n - i)) >> 32);
    int sorted[4];
    for (int j = 0; j < i; ++j) sorted[j] = s[j];
    for (int a = 1; a < i; ++a)
      for (int b = a; b > 0 && sorted[b - 1] > sorted[b]; --b) { int t = sorted[b]; sorted[b] = sorted[b - 1]; sorted[b - 1] = t; }
    for (int j = 0; j < i; ++j)
      if (v >= sorted[j]) ++v;
    s[i] = v;
  }
  for (int i = 0; i < S; ++i) samples[(size_t)h * S + i] = s[i];
}
// raw Philox-4x32-10 block (known-answer tests: the Random123 vectors, tests/test_gpu_plane.py)
__global__ void philox_raw_kernel(const uint32_t* __restrict__ ctr_key /*4 + 2*/, uint32_t* __restrict__ out4) {
  uint32_t c0 = ctr_key[0], c1 = ctr_key[1], c2 = ctr_key[2], c3 = ctr_key[3], k0 = ctr_key[4], k1 = ctr_key[5];
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out4[0] = c0; out4[1] = c1; out4[2] = c2; out4[3] = c3;
}
int sac_philox_raw(pitt_ctx* ctx, const uint32_t* h_ctr_key6, uint32_t* h_out4) {
  uint32_t* d = nullptr;
  PITT_TRY(arena_alloc(ctx, 16, &d));
  PITT_CUDA(ctx, cudaMemcpyAsync(d, h_ctr_key6, 6 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  philox_raw_kernel<<<1, 1, 0, ctx->stream>>>(d, d + 8);
  ctx->launches++;
  PITT_CUDA(ctx, cudaMemcpyAsync(h_out4, d + 8, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  return PITT_OK;
}
int sac_philox_samples(pitt_ctx* ctx, int* d_samples, int H, int S, int n, uint32_t stream_id) {
  philox_samples_kernel<<<cdiv(H, 256), 256, 0, ctx->stream>>>(d_samples, H, S, n, ctx->seed, stream_id);
  PITT_LAUNCH_CHECK(ctx, "philox_samples_kernel");
  return PITT_OK;
}

// ------------------------------------------------------------------ RandomSampleConsensus::computeModel scan (host)
struct RansacScan {
  int iterations = 0, skipped = 0, pos = 0;
  int best = -INT_MAX, best_pos = -1;
  double k = 1.0;
  bool done = false;
};
// Consumes hypotheses [scan.pos, H) of the stream; returns true when the loop has terminated.
static bool ransac_scan(RansacScan& sc, const pitt_sac_params& p, int n, int S, const int* counts, const uint8_t* flags, int H) {
  const double log_probability = log(1.0 - p.probability);
  const double one_over_indices = 1.0 / (double)n;
  const int max_skip = p.max_iterations * 10;
  while ((double)sc.iterations < sc.k && sc.skipped < max_skip) {
    if (sc.pos >= H) return false;  // need more samples
    int h = sc.pos++;
    if (!(flags[h] & 1)) { ++sc.skipped; continue; }
    int cnt = counts[h];
    if (cnt > sc.best) {
      sc.best = cnt;
      sc.best_pos = h;
      double w = (double)cnt * one_over_indices;
      double p_no = 1.0 - pow(w, (double)S);
      p_no = std::max(DBL_EPSILON, p_no);
      p_no = std::min(1.0 - DBL_EPSILON, p_no);
      sc.k = log_probability / log(p_no);
    }
    ++sc.iterations;
    if (sc.iterations > p.max_iterations) break;
  }
  sc.done = true;
  return true;
}

int lm_refine(pitt_ctx* ctx, const pitt_cloud* c, int model, const float* d_model, const int* d_idx, const int* d_n_idx,
              int n_idx_host, float* d_refined, int* d_lm_info /*[2]: status,nfev*/);

// refinement of the winning model + final inlier selection (SACSegmentation::segment after computeModel):
// d_model[8] -> d_refined[8], *d_n_model = inliers of the unrefined model, *d_n_final / d_inl = final set
int sac_finish(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params& p, const Limits& L, const ScoreParams& sp,
               const float* d_model, float* d_refined, int* d_n_model, int* d_n_final, int* d_lm, int* d_inl) {
  const int model = p.model;
  if (model == PITT_MODEL_PLANE) {
    if (p.optimize) {
      PITT_TRY(plane_refine(ctx, c, d_model, nullptr, nullptr, L, sp, d_refined, d_n_model));
      PITT_TRY(sac_select(ctx, c, model, d_refined, L, sp, d_inl, d_n_final));
    } else {
      PITT_TRY(sac_select(ctx, c, model, d_model, L, sp, d_inl, d_n_final));
      PITT_CUDA(ctx, cudaMemcpyAsync(d_refined, d_model, 8 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
      PITT_CUDA(ctx, cudaMemcpyAsync(d_n_model, d_n_final, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    }
  } else {
    PITT_TRY(sac_select(ctx, c, model, d_model, L, sp, d_inl, d_n_model));
    if (p.optimize) {
      PITT_TRY(lm_refine(ctx, c, model, d_model, d_inl, d_n_model, -1, d_refined, d_lm));
      PITT_TRY(sac_select(ctx, c, model, d_refined, L, sp, d_inl, d_n_final));
    } else {
      PITT_CUDA(ctx, cudaMemcpyAsync(d_refined, d_model, 8 * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
      PITT_CUDA(ctx, cudaMemcpyAsync(d_n_final, d_n_model, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    }
  }
  return PITT_OK;
}

// winner of a hypothesis split: the sample of hypothesis best[0] (all ranks hold the whole sample table)
__global__ void pick_sample_kernel(const int* __restrict__ samples, int S, int H, const int* __restrict__ best, int* __restrict__ one) {
  const int h = best[0];
  if (threadIdx.x < S) one[threadIdx.x] = (h >= 0 && h < H) ? samples[(size_t)h * S + threadIdx.x] : 0;
}
int sac_finish_from_winner(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params& p, const int* d_samples_all, int H_all,
                           const int* d_best, SacDeviceResult* out) {
  const int model = p.model;
  const int S = sample_size(model);
  memset(&out->info, 0, sizeof(out->info));
  out->info.best_hypothesis = -1;
  out->n_inliers = 0;
  out->n_coeffs = 0;
  out->d_inliers = nullptr;
  if (c->n < S) return PITT_OK;
  const Limits L = limits_for(p);
  const ScoreParams sp = score_params_for(p, L);
  int* d_ints = nullptr;
  float* d_flt = nullptr;
  int* d_one = nullptr;
  int* d_inl = nullptr;
  HypRec* d_rec = nullptr;
  uint8_t* d_flag = nullptr;
  PITT_TRY(arena_alloc(ctx, 8, &d_ints));
  PITT_TRY(arena_alloc(ctx, 16, &d_flt));
  PITT_TRY(arena_alloc(ctx, 4, &d_one));
  PITT_TRY(arena_alloc(ctx, (size_t)c->n, &d_inl));
  PITT_TRY(arena_alloc(ctx, 1, &d_rec));
  PITT_TRY(arena_alloc(ctx, 4, &d_flag));
  PITT_CUDA(ctx, cudaMemsetAsync(d_ints, 0, 8 * sizeof(int), ctx->stream));
  pick_sample_kernel<<<1, 32, 0, ctx->stream>>>(d_samples_all, S, H_all, d_best, d_one);
  PITT_LAUNCH_CHECK(ctx, "pick_sample_kernel");
  PITT_TRY(sac_estimate(ctx, c, model, d_one, 1, L, sp, d_rec, d_flt, d_flag));  // d_flt[0..8) = model coefficients
  PITT_TRY(sac_finish(ctx, c, p, L, sp, d_flt, d_flt + 8, d_ints + 2, d_ints + 3, d_ints + 4, d_inl));
  PITT_TRY(pinned_reserve(ctx, 256));
  int* h_ints = (int*)ctx->h_pin;
  float* h_flt = (float*)((char*)ctx->h_pin + 64);
  int* h_best = (int*)((char*)ctx->h_pin + 192);
  PITT_CUDA(ctx, cudaMemcpyAsync(h_ints, d_ints, 8 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, cudaMemcpyAsync(h_flt, d_flt, 16 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, cudaMemcpyAsync(h_best, d_best, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  if (h_best[0] < 0 || h_best[0] >= H_all) return PITT_OK;
  const int NC = coeff_count(model);
  out->info.best_hypothesis = h_best[0];
  out->info.best_count = h_best[1];
  out->info.n_inliers_model = h_ints[2];
  out->info.lm_info = h_ints[4];
  out->info.lm_nfev = h_ints[5];
  for (int i = 0; i < 8; ++i) out->info.model_coeffs[i] = i < NC ? h_flt[i] : 0.0f;
  for (int i = 0; i < 8; ++i) out->coeffs[i] = i < NC ? h_flt[8 + i] : 0.0f;
  out->n_coeffs = NC;
  out->n_inliers = h_ints[3];
  out->d_inliers = d_inl;
  return PITT_OK;
}

// ------------------------------------------------------------------ seg.segment()
// sample points of a cloud that is still in (pinned, device-mapped) host memory -> contiguous device copy + identity sample table
__global__ void gather_points_kernel(const float4* __restrict__ host_xyz, const int* __restrict__ samples, int ns,
                                     float4* __restrict__ out, int* __restrict__ iota) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  out[i] = host_xyz[samples[i]];
  iota[i] = i;
}

int sac_segment_impl(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params& p, SacDeviceResult* out) {
  const int model = p.model;
  const int S = sample_size(model);
  const int n = c->n;
  memset(&out->info, 0, sizeof(out->info));
  out->info.best_hypothesis = -1;
  out->n_inliers = 0;
  out->n_coeffs = 0;
  out->d_inliers = nullptr;
  if (model < 0 || model > 3) return fail(ctx, PITT_ERR_INVALID, "bad model");
  if ((model == PITT_MODEL_CYLINDER || model == PITT_MODEL_CONE) && !c->has_normals)
    return fail(ctx, PITT_ERR_STATE, "cylinder/cone segmentation needs normals on the cloud");
  if (n < S || p.max_iterations < 0) return PITT_OK;  // getSamples fails: "No solution found"
  const Limits L = limits_for(p);
  const ScoreParams sp = score_params_for(p, L);
  const bool all_h = (p.stop == PITT_STOP_ALL_H);
  const int H_first = all_h ? p.max_iterations : p.max_iterations + 1;
  if (H_first <= 0) return PITT_OK;

  // result block on the device: best[2], n_model, n_final, lm[2], model coeffs[8], refined[8]
  int* d_ints = nullptr;
  float* d_flt = nullptr;
  PITT_TRY(arena_alloc(ctx, 8, &d_ints));
  PITT_TRY(arena_alloc(ctx, 16, &d_flt));
  PITT_CUDA(ctx, cudaMemsetAsync(d_ints, 0, 8 * sizeof(int), ctx->stream));
  int* d_best = d_ints;         // [0]=idx [1]=count
  int* d_n_model = d_ints + 2;
  int* d_n_final = d_ints + 3;
  int* d_lm = d_ints + 4;       // [4],[5]
  float* d_model = d_flt;
  float* d_refined = d_flt + 8;
  int* d_inl = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)n, &d_inl));

  TraceScope ts_all(ctx, "  sac_segment_impl");
  PclSampleStream stream(n, model, (c->h_valid ? c->h_xyz.data() : c->h_src));
  if (c->stream_chunks > 0 && (model != PITT_MODEL_PLANE || p.sampler == PITT_SAMPLER_PHILOX)) PITT_TRY(stream_complete(ctx, c));
  RansacScan scan;
  std::vector<int> h_samples;
  int H_total = 0;
  int batch = H_first;
  // PCL's adaptive stop usually ends a 10 000-iteration job (C3) after a handful of hypotheses: big jobs are scored in
  // batches of the sample stream (256, then 4096 at a time) and the host scan decides after each batch whether PCL would
  // have stopped inside it. Results do not depend on the batching (the scan runs in stream order); small jobs (the frame
  // path: at most 1001 hypotheses on a few thousand points) stay one batch = one round trip.
  if (!all_h && c->stream_chunks == 0 && (double)n * (double)H_first >= 33554432.0) batch = 256;
  int* d_samples = nullptr;
  HypRec* d_recs = nullptr;
  float* d_coeffs8 = nullptr;
  uint8_t* d_flags = nullptr;
  int* d_counts = nullptr;
  std::vector<int> h_counts;
  std::vector<uint8_t> h_flags;
  int winner = -1, winner_count = 0;

  // Every scanned hypothesis advances PCL's iteration or skip counter and both are bounded (max_iterations, 10 x that), so
  // the scan reports done after finitely many batches; the sample sources run dry on their own (replay_count).
  for (int round = 0;; ++round) {
    const int H = batch;
    PITT_TRY(arena_alloc(ctx, (size_t)H * S, &d_samples));
    PITT_TRY(arena_alloc(ctx, (size_t)H, &d_recs));
    PITT_TRY(arena_alloc(ctx, (size_t)H * 8, &d_coeffs8));
    PITT_TRY(arena_alloc(ctx, (size_t)H, &d_flags));
    PITT_TRY(arena_alloc(ctx, (size_t)H, &d_counts));
    int H_have = H;
    if (p.sampler == PITT_SAMPLER_PHILOX) {
      PITT_TRY(sac_philox_samples(ctx, d_samples, H, S, n, (uint32_t)(round + 1)));
    } else {
      h_samples.resize((size_t)H * S);
      if (p.sampler == PITT_SAMPLER_REPLAY) {
        if (!p.replay_samples) return fail(ctx, PITT_ERR_INVALID, "replay_samples is null");
        H_have = std::min(H, p.replay_count - H_total);
        if (H_have <= 0) break;
        const int* src = p.replay_samples + (size_t)H_total * S;
        for (size_t i = 0; i < (size_t)H_have * S; ++i) {
          // they index the cloud on the device and, in the streaming stage, the caller's host buffer
          if (src[i] < 0 || src[i] >= n) return fail(ctx, PITT_ERR_INVALID, "sample index out of range");
          h_samples[i] = src[i];
        }
      } else {
        H_have = 0;
        for (int h = 0; h < H; ++h) {
          if (!stream.next(h_samples.data() + (size_t)h * S)) break;
          ++H_have;
        }
        if (H_have == 0) break;
      }
      PITT_TRY(pinned_reserve(ctx, (size_t)H * S * sizeof(int) + (size_t)H * 8));
      memcpy(ctx->h_pin, h_samples.data(), (size_t)H_have * S * sizeof(int));
      PITT_CUDA(ctx, cudaMemcpyAsync(d_samples, ctx->h_pin, (size_t)H_have * S * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    trace_mark("  impl: samples ready + upload issued");
    if (c->stream_chunks > 0) {
      // The cloud is still arriving: the sample points are gathered on the host and sent ahead, the hypotheses are
      // estimated from that copy, and every chunk is scored as soon as it is complete.
      const size_t ns = (size_t)H_have * S;
      float4* d_gather = nullptr;
      int* d_iota = nullptr;
      PITT_TRY(arena_alloc(ctx, ns, &d_gather));
      PITT_TRY(arena_alloc(ctx, ns, &d_iota));
      cudaPointerAttributes attr;
      const bool mapped = cudaPointerGetAttributes(&attr, c->h_src) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer;
      if (mapped) {
        // pinned caller buffer: the GPU fetches the 3 H sample points straight from host memory (zero copy, ~20 us); a
        // host-side gather of 15 000 random cache misses would delay the first kernel by far more
        gather_points_kernel<<<cdiv((int)ns, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const float4*>(attr.devicePointer), d_samples,
                                                                           (int)ns, d_gather, d_iota);
        PITT_LAUNCH_CHECK(ctx, "gather_points_kernel");
      } else {
        cudaGetLastError();  // pageable memory: cudaPointerGetAttributes may have set an error on old drivers
        PITT_TRY(pinned2_reserve(ctx, ns * 20));
        float4* hg = reinterpret_cast<float4*>(ctx->h_pin2);
        int* hi = reinterpret_cast<int*>((char*)ctx->h_pin2 + ns * 16);
        const float4* src = reinterpret_cast<const float4*>(c->h_src);
        for (size_t i = 0; i < ns; ++i) { hg[i] = src[h_samples[i]]; hi[i] = (int)i; }
        PITT_CUDA(ctx, cudaMemcpyAsync(d_gather, ctx->h_pin2, ns * 16, cudaMemcpyHostToDevice, ctx->stream));
        PITT_CUDA(ctx, cudaMemcpyAsync(d_iota, (char*)ctx->h_pin2 + ns * 16, ns * 4, cudaMemcpyHostToDevice, ctx->stream));
      }
      pitt_cloud sample_view;
      sample_view.n = (int)ns;
      sample_view.d_xyz = d_gather;
      PITT_TRY(sac_estimate(ctx, &sample_view, model, d_iota, H_have, L, sp, d_recs, d_coeffs8, d_flags));
      PITT_TRY(sac_score_plane_streaming(ctx, c, d_recs, H_have, sp, d_counts, d_gather, (int)ns));
    } else {
      PITT_TRY(sac_estimate(ctx, c, model, d_samples, H_have, L, sp, d_recs, d_coeffs8, d_flags));
      PITT_TRY(sac_score(ctx, c, model, d_recs, H_have, sp, d_counts));
    }
    out->info.hypotheses += H_have;
    trace_mark("  impl: estimate + score issued");
    if (all_h) {
      PITT_TRY(sac_winner(ctx, d_counts, d_flags, H_have, d_coeffs8, d_best, d_model));
      winner = -2;  // decided on the device
      H_total += H_have;
      break;
    }
    // PCL stop rule: scan counts in stream order on the host
    h_counts.resize(H_total + H_have);
    h_flags.resize(H_total + H_have);
    PITT_CUDA(ctx, cudaMemcpyAsync(h_counts.data() + H_total, d_counts, (size_t)H_have * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, cudaMemcpyAsync(h_flags.data() + H_total, d_flags, (size_t)H_have, cudaMemcpyDeviceToHost, ctx->stream));
    PITT_CUDA(ctx, pitt::stream_sync(ctx));
    if (model == PITT_MODEL_PLANE && p.sampler == PITT_SAMPLER_PCL_MT19937 && !c->h_valid) {
      // speculative draws: a collinear triple would have been redrawn by getSamples. Extremely
      // rare; fetch the host mirror and restart with exact isSampleGood on the host.
      bool bad = false;
      for (int h = 0; h < H_have; ++h) bad |= !(h_flags[H_total + h] & 1);
      if (bad) {
        PITT_TRY(ensure_host_mirror(ctx, c));
        return sac_segment_impl(ctx, c, p, out);
      }
    }
    int base = H_total;
    H_total += H_have;
    bool done = ransac_scan(scan, p, n, S, h_counts.data(), h_flags.data(), H_total);
    if (scan.best_pos >= base) {
      // remember the coefficients of a winner found in this batch
      PITT_CUDA(ctx, cudaMemcpyAsync(d_model, d_coeffs8 + (size_t)(scan.best_pos - base) * 8, 8 * sizeof(float),
                                     cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (done || H_have < H) break;
    batch = std::max(64, std::min(p.max_iterations + 1, 4096));
  }
  if (!all_h) {
    out->info.iterations = scan.iterations;
    out->info.skipped = scan.skipped;
    winner = scan.best_pos;
    winner_count = scan.best_pos >= 0 ? scan.best : 0;
    if (winner < 0) return PITT_OK;
  }

  const int NC = coeff_count(model);
  TraceScope ts_fin(ctx, "    refine + select");
  PITT_TRY(sac_finish(ctx, c, p, L, sp, d_model, d_refined, d_n_model, d_n_final, d_lm, d_inl));
  // one small D2H for the scalars
  PITT_TRY(pinned_reserve(ctx, 256));
  int* h_ints = (int*)ctx->h_pin;
  float* h_flt = (float*)((char*)ctx->h_pin + 64);
  PITT_CUDA(ctx, cudaMemcpyAsync(h_ints, d_ints, 8 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PITT_CUDA(ctx, cudaMemcpyAsync(h_flt, d_flt, 16 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  trace_mark("  impl: finish issued");
  PITT_CUDA(ctx, pitt::stream_sync(ctx));
  trace_mark("  impl: scalars on the host");
  if (all_h) {
    winner = h_ints[0];
    winner_count = h_ints[1];
    out->info.iterations = H_total;
    if (winner < 0) return PITT_OK;
  }
  out->info.best_hypothesis = winner;
  out->info.best_count = winner_count;
  out->info.n_inliers_model = h_ints[2];
  out->info.lm_info = h_ints[4];
  out->info.lm_nfev = h_ints[5];
  for (int i = 0; i < 8; ++i) out->info.model_coeffs[i] = i < NC ? h_flt[i] : 0.0f;
  for (int i = 0; i < 8; ++i) out->coeffs[i] = i < NC ? h_flt[8 + i] : 0.0f;
  out->n_coeffs = NC;
  out->n_inliers = h_ints[3];
  out->d_inliers = d_inl;
  return PITT_OK;
}


// ------------------------------------------------------------------ seg.segment() without host round trips
// RandomSampleConsensus::computeModel's stop rule on the device: the same sequential scan as ransac_scan above (best-so-far with
// strict '>', adaptive k from the inlier ratio, skipped samples, iterations_ > max_iterations_ break) over the H scored
// hypotheses of the stream, by one thread on a shared-memory copy of the counts. pow / log are CUDA's double functions; where the
// result could differ from the host's libm in a way that changes a decision (k within 1e-9 of an integer that still matters), or
// the batch ends before the loop does, or a speculative plane sample turns out collinear, a flag is raised and the caller
// repeats the fit on the synchronous path. ints: [0] best position, [1] best count, [6] iterations, [7] skipped, [8] flags.
constexpr int SCAN_FLAG_NEED_MORE = 1, SCAN_FLAG_AMBIGUOUS = 2, SCAN_FLAG_BAD_SAMPLE = 4;
__global__ void __launch_bounds__(256)
ransac_scan_kernel(const int* __restrict__ counts, const uint8_t* __restrict__ flags, int H, int n, int S, int max_iterations,
                   double log_probability, int speculative_plane, const float* __restrict__ coeffs8, int* __restrict__ ints,
                   float* __restrict__ model_out, const int* __restrict__ skip, int final_batch,
                   const FitDesc* __restrict__ D = nullptr) {
  extern __shared__ __align__(16) unsigned char scan_smem[];
  if (D) {  // batched: problem blockIdx.x scans the first min(H, d.H) hypotheses of its stream
    const FitDesc d = D[blockIdx.x];
    counts = d.counts; flags = d.flags; coeffs8 = d.coeffs8; ints = d.ints; model_out = d.flt; n = d.n;
    final_batch = (H >= d.H) ? 1 : 0;
    H = min(H, d.H);
    skip = skip ? d.ints + 10 : nullptr;
    if (H <= 0) return;
  }
  if (skip && *skip) return;  // an earlier batch has already ended the loop
  int* s_cnt = reinterpret_cast<int*>(scan_smem);
  uint8_t* s_flag = reinterpret_cast<uint8_t*>(s_cnt + H);
  for (int i = threadIdx.x; i < H; i += blockDim.x) { s_cnt[i] = counts[i]; s_flag[i] = flags[i]; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const double one_over_indices = 1.0 / (double)n;
  const int max_skip = max_iterations * 10;
  int iterations = 0, skipped = 0, pos = 0, best = -INT_MAX, best_pos = -1, fl = 0;
  double k = 1.0;
  while ((double)iterations < k && skipped < max_skip) {
    if (pos >= H) { fl |= SCAN_FLAG_NEED_MORE; break; }
    const int h = pos++;
    if (!(s_flag[h] & 1)) {
      if (speculative_plane) fl |= SCAN_FLAG_BAD_SAMPLE;  // getSamples would have redrawn this triple
      ++skipped;
      continue;
    }
    const int cnt = s_cnt[h];
    if (cnt > best) {
      best = cnt;
      best_pos = h;
      const double w = (double)cnt * one_over_indices;
      double p_no = 1.0 - pow(w, (double)S);
      p_no = fmax(DBL_EPSILON, p_no);
      p_no = fmin(1.0 - DBL_EPSILON, p_no);
      k = log_probability / log(p_no);
      if (k < (double)max_iterations + 2.0 && fabs(k - rint(k)) <= 1e-9 * fmax(1.0, fabs(k))) fl |= SCAN_FLAG_AMBIGUOUS;
    }
    ++iterations;
    if (iterations > max_iterations) break;
  }
  if ((fl & SCAN_FLAG_NEED_MORE) && !final_batch) return;  // the loop runs on into the next batch: ints[10] stays 0
  ints[0] = best_pos;
  ints[1] = best_pos >= 0 ? best : 0;
  ints[6] = iterations;
  ints[7] = skipped;
  ints[8] = fl;
  ints[10] = 1;  // done: the kernels of the following batch return at once
  for (int i = 0; i < 8; ++i) model_out[i] = best_pos >= 0 ? coeffs8[(size_t)best_pos * 8 + i] : 0.0f;
}
// ALL_H: the winner kernel's {index, count} plus the bookkeeping of the block above
__global__ void all_h_info_kernel(int H, int* __restrict__ ints) {
  ints[6] = H;
  ints[7] = 0;
  ints[8] = 0;
}
__global__ void first_inlier_kernel(const int* __restrict__ inl, const int* __restrict__ n_final, int* __restrict__ out,
                                    const FitDesc* __restrict__ D = nullptr) {
  if (D) {  // batched: problem blockIdx.x
    inl = D[blockIdx.x].inl; n_final = D[blockIdx.x].ints + 3; out = D[blockIdx.x].ints + 9;
  }
  *out = (*n_final > 0) ? inl[0] : -1;
}

// Enqueues one whole seg.segment() on ctx->stream and returns at once. *issued = false: the call is trivially empty (no model,
// nothing enqueued). Results: d_ints[0] best position (-1 none), [1] best count, [2] inliers of the un-refined model, [3] final
// inliers, [4] LM status, [5] LM nfev, [6] iterations, [7] skipped, [8] flags (non-zero: repeat on the synchronous path),
// [9] first final inlier index (-1 none); d_flt[0..8) model, [8..16) refined. h_stage: pinned memory for the sample table,
// valid until the stream has consumed it.
int sac_segment_async(pitt_ctx* ctx, const pitt_cloud* c, const pitt_sac_params& p, int* h_stage, SacAsync* out, bool* issued) {
  const int model = p.model;
  const int S = sample_size(model);
  const int n = c->n;
  *issued = false;
  out->d_inl = nullptr;
  out->H = 0;
  if (model < 0 || model > 3) return fail(ctx, PITT_ERR_INVALID, "bad model");
  if ((model == PITT_MODEL_CYLINDER || model == PITT_MODEL_CONE) && !c->has_normals)
    return fail(ctx, PITT_ERR_STATE, "cylinder/cone segmentation needs normals on the cloud");
  if (n < S || p.max_iterations < 0) return PITT_OK;
  const bool all_h = (p.stop == PITT_STOP_ALL_H);
  int H = all_h ? p.max_iterations : p.max_iterations + 1;
  if (H <= 0) return PITT_OK;
  const Limits L = limits_for(p);
  const ScoreParams sp = score_params_for(p, L);
  int* d_samples = nullptr;
  HypRec* d_recs = nullptr;
  float* d_coeffs8 = nullptr;
  uint8_t* d_flags = nullptr;
  int* d_counts = nullptr;
  if (!out->d_ints) PITT_TRY(arena_alloc(ctx, 16, &out->d_ints));  // the caller may supply the result block
  if (!out->d_flt) PITT_TRY(arena_alloc(ctx, 16, &out->d_flt));
  PITT_TRY(arena_alloc(ctx, (size_t)n, &out->d_inl));
  PITT_CUDA(ctx, cudaMemsetAsync(out->d_ints, 0, 16 * sizeof(int), ctx->stream));
  bool speculative = false;
  if (p.sampler == PITT_SAMPLER_PHILOX) {
    PITT_TRY(arena_alloc(ctx, (size_t)H * S, &d_samples));
    PITT_TRY(sac_philox_samples(ctx, d_samples, H, S, n, 1u));
  } else {
    if (p.sampler == PITT_SAMPLER_REPLAY) {
      if (!p.replay_samples) return fail(ctx, PITT_ERR_INVALID, "replay_samples is null");
      H = std::min(H, p.replay_count);
      if (H <= 0) return PITT_OK;
      for (size_t i = 0; i < (size_t)H * S; ++i) {
        const int v = p.replay_samples[i];
        if (v < 0 || v >= n) return fail(ctx, PITT_ERR_INVALID, "sample index out of range");
        h_stage[i] = v;
      }
    } else {
      const float* h_xyz = c->h_valid ? c->h_xyz.data() : nullptr;
      speculative = (model == PITT_MODEL_PLANE) && !h_xyz;
      PclSampleStream stream(n, model, h_xyz);
      int have = 0;
      for (int h = 0; h < H; ++h) {
        if (!stream.next(h_stage + (size_t)h * S)) break;
        ++have;
      }
      H = have;
      if (H <= 0) return PITT_OK;
    }
    PITT_TRY(arena_alloc(ctx, (size_t)H * S, &d_samples));
    PITT_CUDA(ctx, cudaMemcpyAsync(d_samples, h_stage, (size_t)H * S * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  PITT_TRY(arena_alloc(ctx, (size_t)H, &d_recs));
  PITT_TRY(arena_alloc(ctx, (size_t)H * 8, &d_coeffs8));
  PITT_TRY(arena_alloc(ctx, (size_t)H, &d_flags));
  PITT_TRY(arena_alloc(ctx, (size_t)H, &d_counts));
  int* d_ints = out->d_ints;
  float* d_flt = out->d_flt;
  // PCL's adaptive stop usually ends the loop after a few dozen hypotheses: the first SCAN_BATCH of the stream are estimated,
  // scored and scanned first; the kernels of the remainder read the "done" word the scan leaves behind and return at once
  // (results do not depend on the batching: the scan always runs in stream order from the first hypothesis)
  constexpr int SCAN_BATCH = 256;
  const int H1 = (!all_h && H > SCAN_BATCH + SCAN_BATCH / 2) ? SCAN_BATCH : H;
  PITT_TRY(sac_estimate(ctx, c, model, d_samples, H1, L, sp, d_recs, d_coeffs8, d_flags));
  PITT_TRY(sac_score(ctx, c, model, d_recs, H1, sp, d_counts));
  if (all_h) {
    PITT_TRY(sac_winner(ctx, d_counts, d_flags, H, d_coeffs8, d_ints, d_flt));
    all_h_info_kernel<<<1, 1, 0, ctx->stream>>>(H, d_ints);
    PITT_LAUNCH_CHECK(ctx, "all_h_info_kernel");
  } else {
    if ((size_t)H * 5 + 16 > 200 * 1024) return fail(ctx, PITT_ERR_INVALID, "sac_segment_async: too many hypotheses for the device scan");
    const size_t smem = (size_t)H * 5 + 16;
    static bool opt_in[64] = {false};
    if (smem > 48 * 1024 && !opt_in[ctx->device & 63]) {
      PITT_CUDA(ctx, cudaFuncSetAttribute(ransac_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      opt_in[ctx->device & 63] = true;
    }
    const double log_p = log(1.0 - p.probability);
    ransac_scan_kernel<<<1, 256, (size_t)H1 * 5 + 16, ctx->stream>>>(d_counts, d_flags, H1, n, S, p.max_iterations, log_p, speculative ? 1 : 0,
                                                                    d_coeffs8, d_ints, d_flt, nullptr, H1 == H ? 1 : 0);
    PITT_LAUNCH_CHECK(ctx, "ransac_scan_kernel");
    if (H1 < H) {
      ctx->skip_flag = d_ints + 10;
      int st = sac_estimate(ctx, c, model, d_samples + (size_t)H1 * S, H - H1, L, sp, d_recs + H1, d_coeffs8 + (size_t)H1 * 8, d_flags + H1);
      if (st == PITT_OK) st = sac_score(ctx, c, model, d_recs + H1, H - H1, sp, d_counts + H1);
      ctx->skip_flag = nullptr;
      PITT_TRY(st);
      ransac_scan_kernel<<<1, 256, smem, ctx->stream>>>(d_counts, d_flags, H, n, S, p.max_iterations, log_p, speculative ? 1 : 0, d_coeffs8,
                                                        d_ints, d_flt, d_ints + 10, 1);
      PITT_LAUNCH_CHECK(ctx, "ransac_scan_kernel");
    }
  }
  PITT_TRY(sac_finish(ctx, c, p, L, sp, d_flt, d_flt + 8, d_ints + 2, d_ints + 3, d_ints + 4, out->d_inl));
  first_inlier_kernel<<<1, 1, 0, ctx->stream>>>(out->d_inl, d_ints + 3, d_ints + 9);
  PITT_LAUNCH_CHECK(ctx, "first_inlier_kernel");
  out->H = H;
  *issued = true;
  return PITT_OK;
}


// ------------------------------------------------------------------ a batch of seg.segment() calls in one set of launches
template <int MODEL>
static int fit_batch_estimate_score(pitt_ctx* ctx, const FitDesc* d_desc, int nprob, int n_max, int h0, int hc, const Limits& L,
                                    const ScoreParams& sp) {
  estimate_kernel<MODEL><<<dim3(cdiv(hc, 128), nprob), 128, 0, ctx->stream>>>(nullptr, nullptr, nullptr, hc, L, sp, nullptr, nullptr, nullptr,
                                                                            nullptr, d_desc, h0);
  PITT_LAUNCH_CHECK(ctx, "estimate_kernel (batch)");
  // the configuration of launch_score_generic for the largest problem of the batch; smaller ones leave their surplus CTAs idle
  constexpr int TPB = 256;
  constexpr bool TWO = (MODEL == PITT_MODEL_CYLINDER || MODEL == PITT_MODEL_CONE);
  constexpr int P = TWO ? 1 : 8;
  const int tile = TPB * P;
  const int tiles = cdiv(n_max, tile);
  const int want_ctas = std::max(1, ctx->sm_count * 4 / nprob);
  int hyp_chunk = hc < 512 ? hc : 512;
  if (tiles * cdiv(hc, hyp_chunk) < want_ctas) {
    const int chunks_wanted = cdiv(want_ctas, tiles);
    hyp_chunk = std::max(16, cdiv(hc, chunks_wanted));
    if (hyp_chunk > 512) hyp_chunk = 512;
    if (hyp_chunk > hc) hyp_chunk = hc;
  }
  const int n_chunks = cdiv(hc, hyp_chunk);
  int pblocks = std::max(1, (ctx->sm_count * 8 / nprob) / n_chunks);
  if (pblocks > tiles) pblocks = tiles;
  const int pts_per_cta = cdiv(tiles, pblocks) * tile;
  pblocks = cdiv(n_max, pts_per_cta);
  size_t smem = (size_t)hyp_chunk * (sizeof(HypRec) + sizeof(int));
  const dim3 grid(pblocks, n_chunks, nprob);
  if constexpr (TWO) {
    smem += 8 + (size_t)(TPB / 32) * (32 * P + 32) * sizeof(int2);
    score2_kernel<MODEL, P, TPB><<<grid, TPB, smem, ctx->stream>>>(nullptr, nullptr, 0, nullptr, hc, hyp_chunk, pts_per_cta, sp, nullptr, nullptr,
                                                                   d_desc, h0);
  } else {
    score_kernel<MODEL, P, TPB><<<grid, TPB, smem, ctx->stream>>>(nullptr, nullptr, 0, nullptr, hc, hyp_chunk, pts_per_cta, sp, nullptr, nullptr,
                                                                  d_desc, h0);
  }
  PITT_LAUNCH_CHECK(ctx, "score kernel (batch)");
  return PITT_OK;
}
template <int MODEL>
static int fit_batch_model(pitt_ctx* ctx, const pitt_sac_params& p, const FitDesc* d_desc, int nprob, int n_max, int H_max,
                           bool speculative_plane) {
  const Limits L = limits_for(p);
  const ScoreParams sp = score_params_for(p, L);
  constexpr int S = (MODEL == PITT_MODEL_PLANE) ? 3 : (MODEL == PITT_MODEL_SPHERE) ? 4 : (MODEL == PITT_MODEL_CYLINDER) ? 2 : 3;
  constexpr int SCAN_BATCH = 256;
  const int H1 = H_max > SCAN_BATCH + SCAN_BATCH / 2 ? SCAN_BATCH : H_max;
  const double log_p = log(1.0 - p.probability);
  if ((size_t)H_max * 5 + 16 > 48 * 1024) return fail(ctx, PITT_ERR_INVALID, "sac_fit_batch_async: too many hypotheses");
  PITT_TRY((fit_batch_estimate_score<MODEL>(ctx, d_desc, nprob, n_max, 0, H1, L, sp)));
  ransac_scan_kernel<<<nprob, 256, (size_t)H1 * 5 + 16, ctx->stream>>>(nullptr, nullptr, H1, 0, S, p.max_iterations, log_p, speculative_plane ? 1 : 0,
                                                                      nullptr, nullptr, nullptr, nullptr, 0, d_desc);
  PITT_LAUNCH_CHECK(ctx, "ransac_scan_kernel (batch)");
  if (H1 < H_max) {
    PITT_TRY((fit_batch_estimate_score<MODEL>(ctx, d_desc, nprob, n_max, H1, H_max - H1, L, sp)));
    ransac_scan_kernel<<<nprob, 256, (size_t)H_max * 5 + 16, ctx->stream>>>(nullptr, nullptr, H_max, 0, S, p.max_iterations, log_p,
                                                                           speculative_plane ? 1 : 0, nullptr, nullptr, nullptr,
                                                                           reinterpret_cast<const int*>(d_desc), 1, d_desc);
    PITT_LAUNCH_CHECK(ctx, "ransac_scan_kernel (batch)");
  }
  // refinement of the winners + final inliers (SACSegmentation::segment after computeModel), all problems per launch
  if (MODEL == PITT_MODEL_PLANE) {
    if (p.optimize) {
      prep_rec_kernel<PITT_MODEL_PLANE><<<nprob, 1, 0, ctx->stream>>>(nullptr, L, sp, nullptr, d_desc);
      PITT_LAUNCH_CHECK(ctx, "prep_rec_kernel (batch)");
      plane_sums_kernel<<<dim3(REF_BLOCKS, nprob), REF_TPB, 0, ctx->stream>>>(nullptr, 0, nullptr, nullptr, nullptr, sp, nullptr, d_desc);
      PITT_LAUNCH_CHECK(ctx, "plane_sums_kernel (batch)");
      plane_refine_final_kernel<<<nprob, 256, 0, ctx->stream>>>(nullptr, REF_BLOCKS, nullptr, nullptr, nullptr, d_desc);
      PITT_LAUNCH_CHECK(ctx, "plane_refine_final_kernel (batch)");
      select_small_kernel<MODEL><<<nprob, SEL_SMALL_TPB, 0, ctx->stream>>>(nullptr, nullptr, 0, nullptr, L, sp, nullptr, nullptr, d_desc, 1);
      PITT_LAUNCH_CHECK(ctx, "select_small_kernel (batch)");
    } else {
      return fail(ctx, PITT_ERR_INVALID, "sac_fit_batch_async: optimize = 0 takes the per-fit path");
    }
  } else {
    if (!p.optimize) return fail(ctx, PITT_ERR_INVALID, "sac_fit_batch_async: optimize = 0 takes the per-fit path");
    select_small_kernel<MODEL><<<nprob, SEL_SMALL_TPB, 0, ctx->stream>>>(nullptr, nullptr, 0, nullptr, L, sp, nullptr, nullptr, d_desc, 0);
    PITT_LAUNCH_CHECK(ctx, "select_small_kernel (batch)");
    PITT_TRY(lm_refine_batch(ctx, MODEL, d_desc, nprob, n_max));
    select_small_kernel<MODEL><<<nprob, SEL_SMALL_TPB, 0, ctx->stream>>>(nullptr, nullptr, 0, nullptr, L, sp, nullptr, nullptr, d_desc, 1);
    PITT_LAUNCH_CHECK(ctx, "select_small_kernel (batch)");
  }
  first_inlier_kernel<<<nprob, 1, 0, ctx->stream>>>(nullptr, nullptr, nullptr, d_desc);
  PITT_LAUNCH_CHECK(ctx, "first_inlier_kernel (batch)");
  return PITT_OK;
}
int sac_fit_batch_async(pitt_ctx* ctx, const pitt_sac_params& p, const FitDesc* h_desc, const FitDesc* d_desc, int nprob, bool speculative_plane) {
  if (nprob <= 0) return PITT_OK;
  if (p.stop != PITT_STOP_PCL_ADAPTIVE) return fail(ctx, PITT_ERR_INVALID, "sac_fit_batch_async: PCL adaptive stop only");
  int n_max = 1, H_max = 0;
  for (int i = 0; i < nprob; ++i) {
    n_max = std::max(n_max, h_desc[i].n);
    H_max = std::max(H_max, h_desc[i].H);
  }
  if (n_max > SEL_SMALL_MAX || n_max > 4096) return fail(ctx, PITT_ERR_INVALID, "sac_fit_batch_async: problems of at most 4096 points");
  if (H_max <= 0) return PITT_OK;
  switch (p.model) {
    case PITT_MODEL_PLANE: return fit_batch_model<PITT_MODEL_PLANE>(ctx, p, d_desc, nprob, n_max, H_max, speculative_plane);
    case PITT_MODEL_SPHERE: return fit_batch_model<PITT_MODEL_SPHERE>(ctx, p, d_desc, nprob, n_max, H_max, false);
    case PITT_MODEL_CYLINDER: return fit_batch_model<PITT_MODEL_CYLINDER>(ctx, p, d_desc, nprob, n_max, H_max, false);
    case PITT_MODEL_CONE: return fit_batch_model<PITT_MODEL_CONE>(ctx, p, d_desc, nprob, n_max, H_max, false);
    default: return fail(ctx, PITT_ERR_INVALID, "bad model");
  }
}

}  // namespace pitt
