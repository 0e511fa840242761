// lm.cu — Levenberg-Marquardt refinement of sphere / cylinder / cone coefficients (K7).
//
// Replaces SampleConsensusModel{Sphere,Cylinder,Cone}::optimizeModelCoefficients, i.e.
// Eigen::LevenbergMarquardt<Eigen::NumericalDiff<Functor>, float>::minimize() (Eigen 3.2
// unsupported/NonLinearOptimization) reached from seg.segment() with
// setOptimizeCoefficients(true) at sphere_segmentation_srv.cpp:60, cylinder…:114, cone…:115.
//
// One CTA per problem, or for big problems (m_cap >= LM_CLUSTER_MIN rows: the C3 clusters) one thread-block CLUSTER of 8
// CTAs: rows are owned by (CTA, thread) in a fixed stride, every CTA keeps its own copy of the small state and runs the
// same control flow (all decisions derive from bit-identical reductions), reductions over m exchange one double-double
// partial per CTA through L2, and barrier.cluster (release/acquire) replaces __syncthreads.
// The m residual rows live in global/L2 memory (fvec, the m x n Jacobian);
// every O(m) step is data parallel over the CTA, every reduction over m is accumulated in
// double-double by a fixed-shape tree and rounded once to float (the oracle defines those sums as
// exact sums rounded once), and the n x n algebra (n <= 7: lmpar, qrsolv, Givens) runs on thread 0.
#include <cfloat>

#include "pitt_common.cuh"
#include "sac.cuh"

namespace pitt {

constexpr int LM_TPB = 512;
constexpr int LM_NW = LM_TPB / 32;
constexpr int LM_CLUSTER = 8;           // CTAs per problem on the cluster path (portable maximum)
constexpr int LM_CLUSTER_MIN = 4096;    // rows from which the cluster path pays (tools/lm_crossover.py: 4000 rows 1.74 -> 1.53 ms, 50 000 rows 10.8 -> 2.3 ms)
constexpr int LM_XCH_DOUBLES = 2 * 8 * LM_CLUSTER * 2;  // exchange area: 2 buffers x 8 dots x CTAs x (hi, lo)

template <int CL>
__device__ __forceinline__ int lm_rank() {
  if (CL == 1) return 0;
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return (int)r;
}
// barrier over the problem's threads; on the cluster path it also orders the global-memory writes of the CTAs
template <int CL>
__device__ __forceinline__ void lm_sync() {
  if (CL == 1) {
    __syncthreads();
  } else {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}
// scalar written by another CTA of the cluster: read it from L2, never from this SM's L1
template <int CL>
__device__ __forceinline__ float lm_peek(const float* p) { return CL == 1 ? *p : __ldcg(p); }

struct dd {
  double hi, lo;
};
__device__ __forceinline__ void dd_add(dd& s, double x) {
  double t = s.hi + x;
  double bb = t - s.hi;
  double err = (s.hi - (t - bb)) + (x - bb);
  s.hi = t;
  s.lo += err;
}
__device__ __forceinline__ void dd_merge(dd& s, double ohi, double olo) {
  dd_add(s, ohi);
  s.lo += olo;
}

struct LmShared {
  float x[8], xs[8], diag[8], qtf[8], wa1[8], wa2[8], wa3[8], hcoef[8], colSq[8];
  float r[64];
  int perm[8], transp[8];
  double red_hi[LM_NW], red_lo[LM_NW];
  double md_hi[8][LM_NW], md_lo[8][LM_NW];
  float mdot[8];
  int cidx[8];
  float bcast[4];
  int ibcast[4];
};

// Up to 8 simultaneous dot products sum_{i>=r0[j]} a[j][i]*b[j][i] in ONE pass and ONE reduction tree
// (results in S.mdot[j], visible to all threads after return). The sums are exact to double-double,
// so fusing them does not change a single bit compared with separate block_dot calls.
struct DotSet {
  const float* a[8];
  const float* b[8];
  int r0[8];
  int nv;
};
template <int CL>
__device__ void block_multidot(const DotSet& D, int m, LmShared& S, double* xch, int& xch_parity) {
  dd acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = dd{0.0, 0.0};
  int rmin = m;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < D.nv) rmin = min(rmin, D.r0[j]);
  const int rank = lm_rank<CL>();
  const int tpb = (int)blockDim.x, nw = tpb >> 5;  // small fits are launched with fewer threads: cheaper barriers
  for (int i = rank * tpb + threadIdx.x; i < m; i += tpb * CL) {
    if (i < rmin) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < D.nv && i >= D.r0[j]) dd_add(acc[j], (double)D.a[j][i] * (double)D.b[j][i]);
  }
  __syncthreads();  // previous readers of S.md_* / S.mdot are done
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j < D.nv) {
      dd s = acc[j];
      for (int o = 16; o > 0; o >>= 1) {
        double ohi = __shfl_down_sync(0xffffffffu, s.hi, o), olo = __shfl_down_sync(0xffffffffu, s.lo, o);
        dd_merge(s, ohi, olo);
      }
      if ((threadIdx.x & 31) == 0) { S.md_hi[j][threadIdx.x >> 5] = s.hi; S.md_lo[j][threadIdx.x >> 5] = s.lo; }
    }
  }
  __syncthreads();
  // warp w folds the per-warp partials of the dots w, w + nw, ... with a shuffle tree
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (CL == 1) {
    for (int j = w; j < D.nv; j += nw) {
      dd s{0.0, 0.0};
      if (l < nw) { s.hi = S.md_hi[j][l]; s.lo = S.md_lo[j][l]; }
      for (int o = 16; o > 0; o >>= 1) {
        double ohi = __shfl_down_sync(0xffffffffu, s.hi, o), olo = __shfl_down_sync(0xffffffffu, s.lo, o);
        dd_merge(s, ohi, olo);
      }
      if (l == 0) S.mdot[j] = (float)(s.hi + s.lo);
    }
    __syncthreads();
    return;
  }
  // cluster: every CTA publishes its partial (double buffered: a CTA can be one reduction ahead of the slowest reader),
  // then every CTA folds the CL partials in the same order, so that all of them hold the same bits
  double* buf = xch + (size_t)(xch_parity & 1) * (8 * CL * 2);
  xch_parity ^= 1;
  for (int j = w; j < D.nv; j += nw) {
    dd s{0.0, 0.0};
    if (l < nw) { s.hi = S.md_hi[j][l]; s.lo = S.md_lo[j][l]; }
    for (int o = 16; o > 0; o >>= 1) {
      double ohi = __shfl_down_sync(0xffffffffu, s.hi, o), olo = __shfl_down_sync(0xffffffffu, s.lo, o);
      dd_merge(s, ohi, olo);
    }
    if (l == 0) {
      __stcg(buf + ((size_t)j * CL + rank) * 2, s.hi);
      __stcg(buf + ((size_t)j * CL + rank) * 2 + 1, s.lo);
    }
  }
  lm_sync<CL>();
  for (int j = w; j < D.nv; j += nw) {
    dd s{0.0, 0.0};
    if (l < CL) { s.hi = __ldcg(buf + ((size_t)j * CL + l) * 2); s.lo = __ldcg(buf + ((size_t)j * CL + l) * 2 + 1); }
    for (int o = 16; o > 0; o >>= 1) {
      double ohi = __shfl_down_sync(0xffffffffu, s.hi, o), olo = __shfl_down_sync(0xffffffffu, s.lo, o);
      dd_merge(s, ohi, olo);
    }
    if (l == 0) S.mdot[j] = (float)(s.hi + s.lo);
  }
  __syncthreads();
}

// single dot product, same reduction machinery
template <int CL>
__device__ float block_dot(const float* __restrict__ a, const float* __restrict__ b, int r0, int m, LmShared& S, double* xch,
                           int& xch_parity) {
  DotSet D;
  D.nv = 1;
  D.a[0] = a; D.b[0] = b; D.r0[0] = r0;
  block_multidot<CL>(D, m, S, xch, xch_parity);
  return S.mdot[0];
}

// ---- residual functors (float sequences of the PCL OptimizationFunctor::operator())
template <int MODEL>
__device__ void eval_residuals(const float4* __restrict__ xyz, const int* __restrict__ idx, int m, const float* x,
                               float* __restrict__ fvec, int first, int stride) {
  if (MODEL == PITT_MODEL_SPHERE) {
    const float x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3];
    for (int i = first; i < m; i += stride) {
      float4 p = __ldg(xyz + idx[i]);
      float c0 = p.x - x0, c1 = p.y - x1, c2 = p.z - x2;
      fvec[i] = sqrtf((c0 * c0 + c2 * c2) + c1 * c1) - x3;
    }
  } else if (MODEL == PITT_MODEL_CYLINDER) {
    const f3 lp = mk3(x[0], x[1], x[2]), ld = mk3(x[3], x[4], x[5]);
    const float rr = x[6] * x[6];
    for (int i = first; i < m; i += stride) {
      float4 p = __ldg(xyz + idx[i]);
      fvec[i] = (float)((double)sqr_pt_line(mk3(p.x, p.y, p.z), lp, ld) - (double)rr);
    }
  } else {
    const f3 apex = mk3(x[0], x[1], x[2]), ad = mk3(x[3], x[4], x[5]);
    const float apexdotdir = dot0(apex, ad);
    const float dirdotdir = 1.0f / dot0(ad, ad);
    const float tan_a = tanf_d(x[6]);
    for (int i = first; i < m; i += stride) {
      float4 p4 = __ldg(xyz + idx[i]);
      f3 pt = mk3(p4.x, p4.y, p4.z);
      float k = (dot0(pt, ad) - apexdotdir) * dirdotdir;
      f3 proj = apex + k * ad;
      f3 height = apex - proj;
      float r = tan_a * nrm0(height);
      fvec[i] = (float)((double)sqr_pt_line(pt, apex, ad) - (double)(r * r));
    }
  }
}

__device__ __forceinline__ float norm_n(const float* a, int n) {
  float s = 0.0f;
  for (int i = 0; i < n; ++i) s += a[i] * a[i];
  return sqrtf(s);
}
__device__ void make_givens(float p, float q, float& c, float& s) {
  if (q == 0.0f) { c = p < 0.0f ? -1.0f : 1.0f; s = 0.0f; }
  else if (p == 0.0f) { c = 0.0f; s = q < 0.0f ? 1.0f : -1.0f; }
  else if (fabsf(p) > fabsf(q)) {
    float t = q / p;
    float u = sqrtf(1.0f + t * t);
    if (p < 0.0f) u = -u;
    c = 1.0f / u;
    s = -t * c;
  } else {
    float t = p / q;
    float u = sqrtf(1.0f + t * t);
    if (q < 0.0f) u = -u;
    s = -1.0f / u;
    c = -t * s;
  }
}
// Eigen internal::qrsolv on the N x N triangle; thread 0 only. N is a compile-time constant and every
// loop is fully unrolled so that s, sdiag, wa live in registers (the serial n x n algebra is the
// latency floor of the LM kernel: no local-memory round trips).
template <int N>
__device__ __forceinline__ void qrsolv(float (&s)[N][N], const int* ipvt, const float (&diag)[N], const float (&qtb)[N],
                                       float (&x)[N], float (&sdiag)[N]) {
  float wa[N];
#pragma unroll
  for (int j = 0; j < N; ++j) { x[j] = s[j][j]; wa[j] = qtb[j]; }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (j < i) s[i][j] = s[j][i];
  bool stop = false;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    float dl = 0.0f;
#pragma unroll
    for (int q = 0; q < N; ++q)
      if (ipvt[j] == q) dl = diag[q];
    if (dl == 0.0f) stop = true;
    if (!stop) {
#pragma unroll
      for (int k = 0; k < N; ++k)
        if (k >= j) sdiag[k] = 0.0f;
      sdiag[j] = dl;
      float qtbpj = 0.0f;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        if (k >= j) {
          float gc, gs;
          make_givens(-s[k][k], sdiag[k], gc, gs);
          s[k][k] = gc * s[k][k] + gs * sdiag[k];
          float temp = gc * wa[k] + gs * qtbpj;
          qtbpj = -gs * wa[k] + gc * qtbpj;
          wa[k] = temp;
#pragma unroll
          for (int i = 0; i < N; ++i) {
            if (i > k) {
              temp = gc * s[i][k] + gs * sdiag[i];
              sdiag[i] = -gs * s[i][k] + gc * sdiag[i];
              s[i][k] = temp;
            }
          }
        }
      }
    }
  }
  int nsing = N;
#pragma unroll
  for (int j = N - 1; j >= 0; --j)
    if (sdiag[j] == 0.0f) nsing = j;  // first zero of sdiag
#pragma unroll
  for (int j = 0; j < N; ++j)
    if (j >= nsing) wa[j] = 0.0f;
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    if (i < nsing) {
      float acc = 0.0f;
#pragma unroll
      for (int j = 0; j < N; ++j)
        if (j > i && j < nsing) acc += s[j][i] * wa[j];
      wa[i] = (wa[i] - acc) / s[i][i];
    }
  }
#pragma unroll
  for (int j = 0; j < N; ++j) { sdiag[j] = s[j][j]; s[j][j] = x[j]; }
#pragma unroll
  for (int j = 0; j < N; ++j) {
#pragma unroll
    for (int q = 0; q < N; ++q)
      if (ipvt[j] == q) x[q] = wa[j];
  }
}
template <int N>
__device__ __forceinline__ float norm_t(const float (&a)[N]) {
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < N; ++i) s += a[i] * a[i];
  return sqrtf(s);
}
// gather v[perm[j]] with compile-time indexable registers
template <int N>
__device__ __forceinline__ float pick(const float (&v)[N], int idx) {
  float r = 0.0f;
#pragma unroll
  for (int q = 0; q < N; ++q)
    if (idx == q) r = v[q];
  return r;
}
// Eigen internal::lmpar2; thread 0 only. r8 = top N x N of the QR factor in shared memory (row stride 8).
template <int N>
__device__ void lmpar2(const float* r8, const int* perm, int rank, const float* diag_s, const float* qtb_s, float delta,
                       float& par, float* x_out) {
  const float dwarf = FLT_MIN;
  float r[N][N], diag[N], qtb[N], x[N], wa1[N], wa2[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    diag[i] = diag_s[i];
    qtb[i] = qtb_s[i];
#pragma unroll
    for (int j = 0; j < N; ++j) r[i][j] = r8[i * 8 + j];
  }
#pragma unroll
  for (int j = 0; j < N; ++j) wa1[j] = (j < rank) ? qtb[j] : 0.0f;
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    if (i < rank) {
      wa1[i] /= r[i][i];
#pragma unroll
      for (int q = 0; q < N; ++q)
        if (q < i) wa1[q] -= wa1[i] * r[q][i];
    }
  }
#pragma unroll
  for (int j = 0; j < N; ++j) x[j] = 0.0f;
#pragma unroll
  for (int j = 0; j < N; ++j) {
#pragma unroll
    for (int q = 0; q < N; ++q)
      if (perm[j] == q) x[q] = wa1[j];
  }
  int iter = 0;
#pragma unroll
  for (int j = 0; j < N; ++j) wa2[j] = diag[j] * x[j];
  float dxnorm = norm_t<N>(wa2);
  float fp = dxnorm - delta;
  if (fp <= 0.1f * delta) {
    par = 0.0f;
#pragma unroll
    for (int j = 0; j < N; ++j) x_out[j] = x[j];
    return;
  }
  float parl = 0.0f;
  if (rank == N) {
#pragma unroll
    for (int j = 0; j < N; ++j) wa1[j] = pick<N>(diag, perm[j]) * pick<N>(wa2, perm[j]) / dxnorm;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float acc = 0.0f;
#pragma unroll
      for (int j = 0; j < N; ++j)
        if (j < i) acc += r[j][i] * wa1[j];
      wa1[i] = (wa1[i] - acc) / r[i][i];
    }
    float temp = norm_t<N>(wa1);
    parl = fp / delta / temp / temp;
  }
#pragma unroll
  for (int j = 0; j < N; ++j) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (i <= j) acc += r[i][j] * qtb[i];
    wa1[j] = acc / pick<N>(diag, perm[j]);
  }
  float gnorm = norm_t<N>(wa1);
  float paru = gnorm / delta;
  if (paru == 0.0f) paru = dwarf / fminf(delta, 0.1f);
  par = fmaxf(par, parl);
  par = fminf(par, paru);
  if (par == 0.0f) par = gnorm / dxnorm;
  float s[N][N], sdiag[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    sdiag[i] = 0.0f;
#pragma unroll
    for (int j = 0; j < N; ++j) s[i][j] = r[i][j];
  }
#pragma unroll 1
  for (;;) {
    ++iter;
    if (par == 0.0f) par = fmaxf(dwarf, 0.001f * paru);
    float sp = sqrtf(par);
#pragma unroll
    for (int j = 0; j < N; ++j) wa1[j] = sp * diag[j];
    qrsolv<N>(s, perm, wa1, qtb, x, sdiag);
#pragma unroll
    for (int j = 0; j < N; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = norm_t<N>(wa2);
    float temp = fp;
    fp = dxnorm - delta;
    if (fabsf(fp) <= 0.1f * delta || (parl == 0.0f && fp <= temp && temp < 0.0f) || iter == 10) break;
#pragma unroll
    for (int j = 0; j < N; ++j) wa1[j] = pick<N>(diag, perm[j]) * (pick<N>(wa2, perm[j]) / dxnorm);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      wa1[j] /= sdiag[j];
      temp = wa1[j];
#pragma unroll
      for (int i = 0; i < N; ++i)
        if (i > j) wa1[i] -= s[i][j] * temp;
    }
    temp = norm_t<N>(wa1);
    float parc = fp / delta / temp / temp;
    if (fp > 0.0f) parl = fmaxf(parl, par);
    if (fp < 0.0f) paru = fminf(paru, par);
    par = fmaxf(parl, par + parc);
  }
  if (iter == 0) par = 0.0f;
#pragma unroll
  for (int j = 0; j < N; ++j) x_out[j] = x[j];
}

// work layout (floats): fjac [n*m_cap] | fvec [m_cap] | wa4 [m_cap] | val2 [m_cap] | (cluster) exchange area
// CL = 1: one CTA. CL = LM_CLUSTER: launched with a cluster dimension of CL; row i belongs to thread (i mod (LM_TPB CL)).
// SMEM (CL = 1 only): the work arrays live in dynamic shared memory instead of L2 - the fits of a frame have a few hundred
// rows and are pure latency: ~60 dependent passes per iteration, each an L2 round trip otherwise.
extern __shared__ __align__(16) float lm_dyn_smem[];
template <int MODEL, int CL, bool SMEM>
__global__ void __launch_bounds__(LM_TPB)
lm_kernel(const float4* __restrict__ xyz, const int* __restrict__ idx, const int* __restrict__ n_idx_ptr, int m_cap,
          const float* __restrict__ model_in, float* __restrict__ work, float* __restrict__ refined, int* __restrict__ info_out,
          const FitDesc* __restrict__ D = nullptr) {
  __shared__ LmShared S;
  if (D) {  // batched (one CTA per problem, shared-memory variant only): problem blockIdx.x
    const FitDesc d = D[blockIdx.x];
    xyz = d.xyz; idx = d.inl; n_idx_ptr = d.ints + 2; model_in = d.flt; refined = d.flt + 8; info_out = d.ints + 4;
  }
  constexpr int n = (MODEL == PITT_MODEL_SPHERE) ? 4 : 7;
  const int m = min(*n_idx_ptr, m_cap);
  float* fjac = SMEM ? lm_dyn_smem : work;
  float* fvec = fjac + (size_t)n * m_cap;
  float* wa4 = fvec + m_cap;
  float* val2 = wa4 + m_cap;
  double* xch = reinterpret_cast<double*>(((uintptr_t)(val2 + m_cap) + 15) & ~(uintptr_t)15);  // cluster path only
  int xch_parity = 0;
  const int rank = lm_rank<CL>();
  const int g0 = rank * (int)blockDim.x + threadIdx.x;  // first row of this thread
  const int GS = (int)blockDim.x * CL;                  // row stride
  const bool lead = (threadIdx.x == 0) && (rank == 0);  // the one thread that writes single global elements
  const float eps = 1.1920928955078125e-07f;
  const float ftol = sqrtf(eps), xtol = sqrtf(eps), gtol = 0.0f, factor = 100.0f;
  const int maxfev = 400;
  if (threadIdx.x < 8) S.x[threadIdx.x] = model_in[threadIdx.x];
  lm_sync<CL>();
  int status = -1, nfev = 0;
  bool run = true;
  if (MODEL == PITT_MODEL_SPHERE) { if (m <= 4) { run = false; status = 0; } }
  else if (m == 0) { run = false; status = 0; }
  bool minimized = run;
  if (run && m < n) { run = false; status = 0; }  // ImproperInputParameters: x untouched
  float fnorm = 0.f, par = 0.f, delta = 0.f, xnorm = 0.f;
  int iter = 1;
  if (run) {
    nfev = 1;
    eval_residuals<MODEL>(xyz, idx, m, S.x, fvec, g0, GS);
    lm_sync<CL>();
    fnorm = sqrtf(block_dot<CL>(fvec, fvec, 0, m, S, xch, xch_parity));
  }
  while (run && status == -1) {
    // ---- NumericalDiff (forward): f(x) again, then one perturbed evaluation per parameter
    const float h_eps = sqrtf(eps);
    eval_residuals<MODEL>(xyz, idx, m, S.x, wa4, g0, GS);  // val1
    nfev++;
    lm_sync<CL>();
    for (int j = 0; j < n; ++j) {
      if (threadIdx.x == 0) {
        for (int q = 0; q < n; ++q) S.xs[q] = S.x[q];
        float h = h_eps * fabsf(S.x[j]);
        if (h == 0.0f) h = h_eps;
        S.xs[j] += h;
        S.bcast[1] = h;
      }
      lm_sync<CL>();
      const float h = S.bcast[1];
      eval_residuals<MODEL>(xyz, idx, m, S.xs, val2, g0, GS);
      nfev++;
      lm_sync<CL>();
      float* cj = fjac + (size_t)j * m_cap;
      for (int i = g0; i < m; i += GS) cj[i] = (val2[i] - wa4[i]) / h;
      lm_sync<CL>();
    }
    // ---- column norms, ColPivHouseholderQR (columns are swapped logically through S.cidx),
    //      with Q^T fvec computed on the fly: wa4 rides along as an extra column
    for (int i = g0; i < m; i += GS) wa4[i] = fvec[i];
    if (threadIdx.x < 8) S.cidx[threadIdx.x] = threadIdx.x;
    lm_sync<CL>();
    {
      DotSet D;
      D.nv = n;
      for (int j = 0; j < n; ++j) { D.a[j] = D.b[j] = fjac + (size_t)j * m_cap; D.r0[j] = 0; }
      block_multidot<CL>(D, m, S, xch, xch_parity);
      if (threadIdx.x == 0)
        for (int j = 0; j < n; ++j) { S.wa2[j] = sqrtf(S.mdot[j]); S.colSq[j] = S.mdot[j]; }
      lm_sync<CL>();
    }
    float threshold_helper, maxpivot = 0.0f;
    int nonzero_pivots = n;
    {
      float mx = S.colSq[0];
      for (int k = 1; k < n; ++k) mx = fmaxf(mx, S.colSq[k]);
      threshold_helper = mx * (eps * eps) / (float)m;
    }
    for (int k = 0; k < n; ++k) {
      int big = k;
      for (int j = k + 1; j < n; ++j)
        if (S.colSq[j] > S.colSq[big]) big = j;
      float* cb = fjac + (size_t)S.cidx[big] * m_cap;
      {
        DotSet D;  // squared norm of the pivot column from row k, and of its tail from row k+1
        D.nv = 2;
        D.a[0] = D.b[0] = cb; D.r0[0] = k;
        D.a[1] = D.b[1] = cb; D.r0[1] = k + 1;
        block_multidot<CL>(D, m, S, xch, xch_parity);
      }
      const float bigSq = S.mdot[0];
      const float tailSq = (m - k == 1) ? 0.0f : S.mdot[1];
      lm_sync<CL>();
      if (bigSq < threshold_helper * (float)(m - k)) {
        if (threadIdx.x == 0) S.colSq[big] = bigSq;
        nonzero_pivots = k;
        for (int j = k; j < n; ++j) {
          if (threadIdx.x == 0) { S.hcoef[j] = 0.0f; S.transp[j] = j; }
          float* cj = fjac + (size_t)S.cidx[j] * m_cap;
          for (int i = g0; i < m; i += GS) if (i >= j + 1) cj[i] = 0.0f;
        }
        lm_sync<CL>();
        break;
      }
      if (threadIdx.x == 0) {
        S.colSq[big] = bigSq;
        S.transp[k] = big;
        if (k != big) {
          int t = S.cidx[k]; S.cidx[k] = S.cidx[big]; S.cidx[big] = t;
          float q = S.colSq[k]; S.colSq[k] = S.colSq[big]; S.colSq[big] = q;
        }
      }
      lm_sync<CL>();
      float* ck = fjac + (size_t)S.cidx[k] * m_cap;
      const float c0 = lm_peek<CL>(ck + k);
      float tau, beta, den = 1.0f;
      if (tailSq == 0.0f) {
        tau = 0.0f;
        beta = c0;
      } else {
        beta = sqrtf(c0 * c0 + tailSq);
        if (c0 >= 0.0f) beta = -beta;
        den = c0 - beta;
        tau = (beta - c0) / beta;
      }
      if (fabsf(beta) > maxpivot) maxpivot = fabsf(beta);
      // essential part of the reflector (in place), then v . (remaining columns and wa4) in one pass
      for (int i = g0; i < m; i += GS) if (i >= k + 1) ck[i] = (tailSq == 0.0f) ? 0.0f : ck[i] / den;
      lm_sync<CL>();  // everyone read c0 = ck[k] and the scaled tail is complete
      if (threadIdx.x == 0) S.hcoef[k] = tau;
      if (lead) ck[k] = beta;
      DotSet D;
      D.nv = 0;
      for (int j = k + 1; j < n; ++j) { D.a[D.nv] = ck; D.b[D.nv] = fjac + (size_t)S.cidx[j] * m_cap; D.r0[D.nv] = k + 1; D.nv++; }
      D.a[D.nv] = ck; D.b[D.nv] = wa4; D.r0[D.nv] = k + 1; D.nv++;
      if (m - k > 1) block_multidot<CL>(D, m, S, xch, xch_parity);
      else lm_sync<CL>();
#pragma unroll 1
      for (int t = 0; t < D.nv; ++t) {
        float* cj = const_cast<float*>(D.b[t]);
        if (m - k == 1) {
          if (lead) cj[k] *= (1.0f - tau);
        } else {
          const float tmp = S.mdot[t] + lm_peek<CL>(cj + k);
          for (int i = g0; i < m; i += GS) if (i >= k + 1) cj[i] -= tmp * (tau * ck[i]);
          lm_sync<CL>();  // all threads have read cj[k]
          if (lead) cj[k] -= tau * tmp;
        }
      }
      lm_sync<CL>();
      if (threadIdx.x == 0)
        for (int j = k + 1; j < n; ++j) { float v = lm_peek<CL>(fjac + (size_t)S.cidx[j] * m_cap + k); S.colSq[j] -= v * v; }
      lm_sync<CL>();
    }
    if (threadIdx.x == 0) {
      for (int j = 0; j < n; ++j) S.perm[j] = j;
      for (int k = 0; k < nonzero_pivots; ++k) { int t = S.perm[k]; S.perm[k] = S.perm[S.transp[k]]; S.perm[S.transp[k]] = t; }
      if (iter == 1) {
        for (int j = 0; j < n; ++j) S.diag[j] = (S.wa2[j] == 0.0f) ? 1.0f : S.wa2[j];
        float t[8];
        for (int j = 0; j < n; ++j) t[j] = S.diag[j] * S.x[j];
        S.bcast[2] = norm_n(t, n);
      }
    }
    lm_sync<CL>();
    if (iter == 1) {
      xnorm = S.bcast[2];
      delta = factor * xnorm;
      if (delta == 0.0f) delta = factor;
    }
    // ---- small algebra on thread 0
    if (threadIdx.x == 0) {
      for (int j = 0; j < n; ++j) S.qtf[j] = lm_peek<CL>(wa4 + j);
      for (int i = 0; i < 64; ++i) S.r[i] = 0.0f;
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) S.r[i * 8 + j] = lm_peek<CL>(fjac + (size_t)S.cidx[j] * m_cap + i);
      float gnorm = 0.0f;
      if (fnorm != 0.0f)
        for (int j = 0; j < n; ++j)
          if (S.wa2[S.perm[j]] != 0.0f) {
            float acc = 0.0f;
            for (int i = 0; i <= j; ++i) acc += S.r[i * 8 + j] * (S.qtf[i] / fnorm);
            gnorm = fmaxf(gnorm, fabsf(acc / S.wa2[S.perm[j]]));
          }
      S.bcast[3] = gnorm;
      for (int j = 0; j < n; ++j) S.diag[j] = fmaxf(S.diag[j], S.wa2[j]);
      int rank = 0;
      float thr = fabsf(maxpivot) * (eps * (float)n);
      for (int i = 0; i < nonzero_pivots; ++i) rank += (fabsf(S.r[i * 8 + i]) > thr) ? 1 : 0;
      S.ibcast[0] = rank;
    }
    lm_sync<CL>();
    const float gnorm = S.bcast[3];
    const int rank = S.ibcast[0];
    if (gnorm <= gtol) { status = 4; break; }
    float ratio = 0.0f;
    do {
      if (threadIdx.x == 0) {
        float p = par;
        lmpar2<n>(S.r, S.perm, rank, S.diag, S.qtf, delta, p, S.wa1);
        S.bcast[1] = p;
        for (int j = 0; j < n; ++j) { S.wa1[j] = -S.wa1[j]; S.wa2[j] = S.x[j] + S.wa1[j]; }
        float t[8];
        for (int j = 0; j < n; ++j) t[j] = S.diag[j] * S.wa1[j];
        S.bcast[2] = norm_n(t, n);
        for (int i = 0; i < n; ++i) {
          float acc = 0.0f;
          for (int j = i; j < n; ++j) acc += S.r[i * 8 + j] * S.wa1[S.perm[j]];
          S.wa3[i] = acc;
        }
      }
      lm_sync<CL>();
      par = S.bcast[1];
      const float pnorm = S.bcast[2];
      if (iter == 1) delta = fminf(delta, pnorm);
      eval_residuals<MODEL>(xyz, idx, m, S.wa2, wa4, g0, GS);
      ++nfev;
      lm_sync<CL>();
      const float fnorm1 = sqrtf(block_dot<CL>(wa4, wa4, 0, m, S, xch, xch_parity));
      float actred = -1.0f;
      if (0.1f * fnorm1 < fnorm) { float q = fnorm1 / fnorm; actred = 1.0f - q * q; }
      const float q1 = norm_n(S.wa3, n) / fnorm;
      const float temp1 = q1 * q1;
      const float q2 = sqrtf(par) * pnorm / fnorm;
      const float temp2 = q2 * q2;
      const float prered = temp1 + temp2 / 0.5f;
      const float dirder = -(temp1 + temp2);
      ratio = 0.0f;
      if (prered != 0.0f) ratio = actred / prered;
      if (ratio <= 0.25f) {
        float temp = 0.0f;
        if (actred >= 0.0f) temp = 0.5f;
        if (actred < 0.0f) temp = 0.5f * dirder / (dirder + 0.5f * actred);
        if (0.1f * fnorm1 >= fnorm || temp < 0.1f) temp = 0.1f;
        delta = temp * fminf(delta, pnorm / 0.1f);
        par /= temp;
      } else if (!(par != 0.0f && ratio < 0.75f)) {
        delta = pnorm / 0.5f;
        par = 0.5f * par;
      }
      lm_sync<CL>();  // all threads have read S.wa3 / S.wa2 / bcast before thread 0 rewrites them
      if (ratio >= 1e-4f) {
        if (threadIdx.x == 0) {
          for (int j = 0; j < n; ++j) { S.x[j] = S.wa2[j]; S.wa2[j] = S.diag[j] * S.x[j]; }
          S.bcast[2] = norm_n(S.wa2, n);
        }
        for (int i = g0; i < m; i += GS) fvec[i] = wa4[i];
        lm_sync<CL>();
        xnorm = S.bcast[2];
        fnorm = fnorm1;
        ++iter;
      }
      const bool small_red = fabsf(actred) <= ftol && prered <= ftol && 0.5f * ratio <= 1.0f;
      if (small_red && delta <= xtol * xnorm) { status = 3; break; }
      if (small_red) { status = 1; break; }
      if (delta <= xtol * xnorm) { status = 2; break; }
      if (nfev >= maxfev) { status = 5; break; }
      if (fabsf(actred) <= eps && prered <= eps && 0.5f * ratio <= 1.0f) { status = 6; break; }
      if (delta <= eps * xnorm) { status = 7; break; }
      if (gnorm <= eps) { status = 8; break; }
    } while (ratio < 1e-4f);
  }
  lm_sync<CL>();
  if (lead) {
    float out[8];
    for (int i = 0; i < 8; ++i) out[i] = (i < n) ? S.x[i] : 0.0f;
    if (MODEL != PITT_MODEL_SPHERE && minimized) {
      float nn = sqrtf(out[3] * out[3] + out[4] * out[4] + out[5] * out[5]);
      out[3] /= nn; out[4] /= nn; out[5] /= nn;
    }
    for (int i = 0; i < 8; ++i) refined[i] = out[i];
    info_out[0] = status < 0 ? 0 : status;
    info_out[1] = nfev;
  }
}

constexpr int LM_SMEM_ROWS = 4096;  // (7 + 3) x 4096 floats = 160 KB of dynamic shared memory at most

template <int MODEL>
static cudaError_t lm_launch(pitt_ctx* ctx, bool cluster, const float4* xyz, const int* d_idx, const int* d_n_idx, int m_cap,
                             const float* d_model, float* d_work, float* d_refined, int* d_lm_info) {
  constexpr int n = (MODEL == PITT_MODEL_SPHERE) ? 4 : 7;
  if (!cluster && m_cap <= LM_SMEM_ROWS) {
    static bool attr_set[64] = {false};  // the opt-in is per device (and per template instantiation)
    if (!attr_set[ctx->device & 63]) {
      cudaError_t e = cudaFuncSetAttribute(lm_kernel<MODEL, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (n + 3) * LM_SMEM_ROWS * (int)sizeof(float));
      if (e != cudaSuccess) return e;
      attr_set[ctx->device & 63] = true;
    }
    const size_t smem = (size_t)(n + 3) * m_cap * sizeof(float);
    // (128 / 256 threads for the smallest fits were measured: 0.75 vs 0.77 ms at 300 rows, slower from 1000 rows on)
    lm_kernel<MODEL, 1, true><<<1, LM_TPB, smem, ctx->stream>>>(xyz, d_idx, d_n_idx, m_cap, d_model, d_work, d_refined, d_lm_info);
    return cudaGetLastError();
  }
  if (!cluster) {
    lm_kernel<MODEL, 1, false><<<1, LM_TPB, 0, ctx->stream>>>(xyz, d_idx, d_n_idx, m_cap, d_model, d_work, d_refined, d_lm_info);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(LM_CLUSTER);
  cfg.blockDim = dim3(LM_TPB);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = LM_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, lm_kernel<MODEL, LM_CLUSTER, false>, xyz, d_idx, d_n_idx, m_cap, d_model, d_work, d_refined, d_lm_info,
                            (const FitDesc*)nullptr);
}

// one CTA per problem of a batch (rows <= LM_SMEM_ROWS each: the object clusters of a frame), work arrays in shared memory
template <int MODEL>
static cudaError_t lm_launch_batch(pitt_ctx* ctx, const FitDesc* d_desc, int nprob, int m_cap) {
  constexpr int n = (MODEL == PITT_MODEL_SPHERE) ? 4 : 7;
  static bool attr_set[64] = {false};
  if (!attr_set[ctx->device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(lm_kernel<MODEL, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (n + 3) * LM_SMEM_ROWS * (int)sizeof(float));
    if (e != cudaSuccess) return e;
    attr_set[ctx->device & 63] = true;
  }
  const size_t smem = (size_t)(n + 3) * m_cap * sizeof(float);
  lm_kernel<MODEL, 1, true><<<nprob, LM_TPB, smem, ctx->stream>>>(nullptr, nullptr, nullptr, m_cap, nullptr, nullptr, nullptr, nullptr, d_desc);
  return cudaGetLastError();
}
int lm_refine_batch(pitt_ctx* ctx, int model, const FitDesc* d_desc, int nprob, int m_cap) {
  if (m_cap > LM_SMEM_ROWS) return fail(ctx, PITT_ERR_INVALID, "lm_refine_batch: problems of at most 4096 rows");
  m_cap = std::max(m_cap, 1);
  cudaError_t e;
  switch (model) {
    case PITT_MODEL_SPHERE: e = lm_launch_batch<PITT_MODEL_SPHERE>(ctx, d_desc, nprob, m_cap); break;
    case PITT_MODEL_CYLINDER: e = lm_launch_batch<PITT_MODEL_CYLINDER>(ctx, d_desc, nprob, m_cap); break;
    case PITT_MODEL_CONE: e = lm_launch_batch<PITT_MODEL_CONE>(ctx, d_desc, nprob, m_cap); break;
    default: return fail(ctx, PITT_ERR_INVALID, "lm_refine_batch: model has no LM refinement");
  }
  ctx->launches++;
  if (e != cudaSuccess) return fail(ctx, PITT_ERR_CUDA, "lm_kernel launch (batch)", e);
  return PITT_OK;
}

int g_lm_cluster_min = LM_CLUSTER_MIN;  // test hook: rows from which the cluster path is used

int lm_refine(pitt_ctx* ctx, const pitt_cloud* c, int model, const float* d_model, const int* d_idx, const int* d_n_idx,
              int n_idx_host, float* d_refined, int* d_lm_info) {
  // m is only known on the device (d_n_idx); the workspace is sized for the worst case
  const int m_cap = n_idx_host >= 0 ? std::max(n_idx_host, 1) : std::max(c->n, 1);
  const int n = (model == PITT_MODEL_SPHERE) ? 4 : 7;
  const bool cluster = m_cap >= g_lm_cluster_min;
  float* d_work = nullptr;
  PITT_TRY(arena_alloc(ctx, (size_t)(n + 3) * m_cap + 16 + 2 * LM_XCH_DOUBLES, &d_work));
  cudaError_t e;
  switch (model) {
    case PITT_MODEL_SPHERE:
      e = lm_launch<PITT_MODEL_SPHERE>(ctx, cluster, c->d_xyz, d_idx, d_n_idx, m_cap, d_model, d_work, d_refined, d_lm_info);
      break;
    case PITT_MODEL_CYLINDER:
      e = lm_launch<PITT_MODEL_CYLINDER>(ctx, cluster, c->d_xyz, d_idx, d_n_idx, m_cap, d_model, d_work, d_refined, d_lm_info);
      break;
    case PITT_MODEL_CONE:
      e = lm_launch<PITT_MODEL_CONE>(ctx, cluster, c->d_xyz, d_idx, d_n_idx, m_cap, d_model, d_work, d_refined, d_lm_info);
      break;
    default:
      return fail(ctx, PITT_ERR_INVALID, "lm_refine: model has no LM refinement");
  }
  ctx->launches++;
  if (e != cudaSuccess) return fail(ctx, PITT_ERR_CUDA, "lm_kernel launch", e);
  return PITT_OK;
}

}  // namespace pitt
