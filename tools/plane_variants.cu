// plane_variants.cu — inner-loop variants of the plane filter scoring kernel, timed on synthetic data.
// Diagnostic only (no parity): which instruction mix / packing / occupancy runs fastest on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o plane_variants plane_variants.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { return ((u64)__float_as_uint(hi) << 32) | (u64)__float_as_uint(lo); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

constexpr int TILE = 512;
// MODE bits: 1 = count, 2 = and-accumulate, 4 = software pipelined (packed only)
template <int KH, int TPB, bool PACKED, int MODE>
__global__ void __launch_bounds__(TPB) variant(const float4* __restrict__ xyz, int n_tiles, const float4* __restrict__ hyp, int H,
                                              float negT, int* __restrict__ counts, unsigned* __restrict__ ands) {
  __shared__ __align__(16) float4 s_pts[TILE];
  const int hb = blockIdx.y;
  const int h_base = (hb * TPB + threadIdx.x) * KH;
  int cnt[KH];
  unsigned andw[KH];
#pragma unroll
  for (int k = 0; k < KH; ++k) { cnt[k] = 0; andw[k] = 0xffffffffu; }
  if (PACKED) {
    constexpr int KP = KH / 2;
    u64 A[KP], B[KP], C[KP], D[KP];
    const u64 NEGT = pack2(negT, negT);
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      float4 r0 = hyp[h_base + 2 * k], r1 = hyp[h_base + 2 * k + 1];
      A[k] = pack2(r0.x, r1.x); B[k] = pack2(r0.y, r1.y); C[k] = pack2(r0.z, r1.z); D[k] = pack2(r0.w, r1.w);
    }
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      __syncthreads();
      for (int i = threadIdx.x; i < TILE; i += TPB) s_pts[i] = xyz[t * TILE + i];
      __syncthreads();
      if (MODE & 4) {
        u64 TP[KP], TQ[KP];
        {
          float4 p = s_pts[0], q = s_pts[1];
          u64 PX = pack2(p.x, p.x), PY = pack2(p.y, p.y), PZ = pack2(p.z, p.z), QX = pack2(q.x, q.x), QY = pack2(q.y, q.y), QZ = pack2(q.z, q.z);
#pragma unroll
          for (int k = 0; k < KP; ++k) {
            u64 sp = fma2(A[k], PX, fma2(B[k], PY, fma2(C[k], PZ, D[k])));
            u64 sq = fma2(A[k], QX, fma2(B[k], QY, fma2(C[k], QZ, D[k])));
            TP[k] = fma2(sp, sp, NEGT); TQ[k] = fma2(sq, sq, NEGT);
          }
        }
#pragma unroll 2
        for (int i = 2; i < TILE; i += 2) {
          float4 p = s_pts[i], q = s_pts[i + 1];
          u64 PX = pack2(p.x, p.x), PY = pack2(p.y, p.y), PZ = pack2(p.z, p.z), QX = pack2(q.x, q.x), QY = pack2(q.y, q.y), QZ = pack2(q.z, q.z);
#pragma unroll
          for (int k = 0; k < KP; ++k) {
            u64 sp = fma2(A[k], PX, fma2(B[k], PY, fma2(C[k], PZ, D[k])));
            u64 sq = fma2(A[k], QX, fma2(B[k], QY, fma2(C[k], QZ, D[k])));
            const unsigned tp0 = (unsigned)TP[k], tp1 = (unsigned)(TP[k] >> 32), tq0 = (unsigned)TQ[k], tq1 = (unsigned)(TQ[k] >> 32);
            if (MODE & 1) { cnt[2 * k] += tp0 >> 31; cnt[2 * k] += tq0 >> 31; cnt[2 * k + 1] += tp1 >> 31; cnt[2 * k + 1] += tq1 >> 31; }
            if (MODE & 2) { andw[2 * k] &= tp0 & tq0; andw[2 * k + 1] &= tp1 & tq1; }
            TP[k] = fma2(sp, sp, NEGT); TQ[k] = fma2(sq, sq, NEGT);
          }
        }
#pragma unroll
        for (int k = 0; k < KP; ++k) { andw[2 * k] &= (unsigned)TP[k] & (unsigned)TQ[k]; andw[2 * k + 1] &= (unsigned)(TP[k] >> 32) & (unsigned)(TQ[k] >> 32); }
      } else {
#pragma unroll 2
        for (int i = 0; i < TILE; i += 2) {
          float4 p = s_pts[i], q = s_pts[i + 1];
          u64 PX = pack2(p.x, p.x), PY = pack2(p.y, p.y), PZ = pack2(p.z, p.z), QX = pack2(q.x, q.x), QY = pack2(q.y, q.y), QZ = pack2(q.z, q.z);
#pragma unroll
          for (int k = 0; k < KP; ++k) {
            u64 sp = fma2(A[k], PX, fma2(B[k], PY, fma2(C[k], PZ, D[k])));
            u64 sq = fma2(A[k], QX, fma2(B[k], QY, fma2(C[k], QZ, D[k])));
            u64 tp = fma2(sp, sp, NEGT), tq = fma2(sq, sq, NEGT);
            const unsigned tp0 = (unsigned)tp, tp1 = (unsigned)(tp >> 32), tq0 = (unsigned)tq, tq1 = (unsigned)(tq >> 32);
            if (MODE & 1) { cnt[2 * k] += tp0 >> 31; cnt[2 * k] += tq0 >> 31; cnt[2 * k + 1] += tp1 >> 31; cnt[2 * k + 1] += tq1 >> 31; }
            if (MODE & 2) { andw[2 * k] &= tp0 & tq0; andw[2 * k + 1] &= tp1 & tq1; }
            if (!(MODE & 3)) { andw[2 * k] ^= tp0 ^ tq0 ^ tp1 ^ tq1; }  // keep the FMAs alive with 1 LOP3 per 4... (2 LOP3)
          }
        }
      }
    }
  } else {
    float a[KH], b[KH], c[KH], d[KH];
#pragma unroll
    for (int k = 0; k < KH; ++k) { float4 r = hyp[h_base + k]; a[k] = r.x; b[k] = r.y; c[k] = r.z; d[k] = r.w; }
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      __syncthreads();
      for (int i = threadIdx.x; i < TILE; i += TPB) s_pts[i] = xyz[t * TILE + i];
      __syncthreads();
#pragma unroll 2
      for (int i = 0; i < TILE; i += 2) {
        float4 p = s_pts[i], q = s_pts[i + 1];
#pragma unroll
        for (int k = 0; k < KH; ++k) {
          float sp = __fmaf_rn(a[k], p.x, __fmaf_rn(b[k], p.y, __fmaf_rn(c[k], p.z, d[k])));
          float sq = __fmaf_rn(a[k], q.x, __fmaf_rn(b[k], q.y, __fmaf_rn(c[k], q.z, d[k])));
          unsigned tp = __float_as_uint(__fmaf_rn(sp, sp, negT)), tq = __float_as_uint(__fmaf_rn(sq, sq, negT));
          if (MODE & 1) { cnt[k] += tp >> 31; cnt[k] += tq >> 31; }
          if (MODE & 2) andw[k] &= tp & tq;
          if (!(MODE & 3)) andw[k] ^= tp ^ tq;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KH; ++k) {
    if (cnt[k]) atomicAdd(&counts[h_base + k], cnt[k]);
    atomicAnd(&ands[h_base + k], andw[k]);
  }
}

template <int KH, int TPB, bool PACKED, int MODE>
void run(const char* name, const float4* d_xyz, int n, const float4* d_hyp, int H, int* d_counts, unsigned* d_ands, int sms) {
  int n_tiles = n / TILE;
  int hblocks = H / (KH * TPB);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, variant<KH, TPB, PACKED, MODE>, TPB, 0);
  int gx = (sms * occ) / hblocks;
  if (gx < 1) gx = 1;
  dim3 grid(gx, hblocks);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaMemset(d_counts, 0, H * 4);
    cudaEventRecord(e0);
    variant<KH, TPB, PACKED, MODE><<<grid, TPB>>>(d_xyz, n_tiles, d_hyp, H, -1600.0f, d_counts, d_ands);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, variant<KH, TPB, PACKED, MODE>);
  std::vector<int> hc(H);
  cudaMemcpy(hc.data(), d_counts, H * 4, cudaMemcpyDeviceToHost);
  long long sum = 0; for (int v : hc) sum += v;
  double evals = (double)n_tiles * TILE * H;
  printf("%-44s regs %3d occ %2d grid %4dx%d: %.3f ms  %.2f Tevals/s  %.2f cyc/eval/lane (1.965GHz) sum %lld  err %s\n", name, fa.numRegs, occ,
         gx, hblocks, best, evals / best / 1e9, best * 1e-3 * 1.965e9 * sms * 128 / evals, sum, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int n = 512 * 1954, H = 5120 * 2;  // H multiple of every KH*TPB used below (<= 2048)
  std::vector<float4> hx(n), hh(H);
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.0f - 1.0f; };
  for (auto& p : hx) p = make_float4(rnd(), rnd(), rnd() * 0.5f, 1.0f);
  for (auto& r : hh) { float a = rnd(), b = rnd(), c = rnd(); float s = 2048.0f / sqrtf(a * a + b * b + c * c); r = make_float4(a * s, b * s, c * s, rnd() * s); }
  float4 *d_xyz, *d_hyp; int* d_counts; unsigned* d_ands;
  cudaMalloc(&d_xyz, n * 16); cudaMalloc(&d_hyp, H * 16); cudaMalloc(&d_counts, H * 4); cudaMalloc(&d_ands, H * 4);
  cudaMemcpy(d_xyz, hx.data(), n * 16, cudaMemcpyHostToDevice);
  cudaMemcpy(d_hyp, hh.data(), H * 16, cudaMemcpyHostToDevice);
  cudaMemset(d_ands, 0xff, H * 4);
  run<8, 128, true, 7>("packed KH8 TPB128 count+and pipelined", d_xyz, n, d_hyp, H, d_counts, d_ands, sms);
  run<8, 128, true, 3>("packed KH8 TPB128 count+and", d_xyz, n, d_hyp, H, d_counts, d_ands, sms);
  run<8, 128, true, 1>("packed KH8 TPB128 count only", d_xyz, n, d_hyp, H, d_counts, d_ands, sms);
  run<8, 128, true, 2>("packed KH8 TPB128 and only", d_xyz, n, d_hyp, H, d_counts, d_ands, sms);
  run<8, 128, true, 0>("packed KH8 TPB128 fma only (+xor)", d_xyz, n, d_hyp, H, d_counts, d_ands, sms);
  run<4, 128, true, 7>("packed KH4 TPB128 count+and pipelined", d_xyz, n, d_hyp, H, d_counts, d_ands, sms);
  run<4, 128, true, 3>("packed KH4 TPB128 count+and", d_xyz, n, d_hyp, H, d_counts, d_ands, sms);
  run<4, 256, true, 3>("packed KH4 TPB256 count+and", d_xyz, n, d_hyp, H, d_counts, d_ands, sms);
  run<16, 128, true, 3>("packed KH16 TPB128 count+and", d_xyz, n, d_hyp, H, d_counts, d_ands, sms);
  run<8, 128, false, 3>("scalar KH8 TPB128 count+and", d_xyz, n, d_hyp, H, d_counts, d_ands, sms);
  run<4, 128, false, 3>("scalar KH4 TPB128 count+and", d_xyz, n, d_hyp, H, d_counts, d_ands, sms);
  run<4, 256, false, 3>("scalar KH4 TPB256 count+and", d_xyz, n, d_hyp, H, d_counts, d_ands, sms);
  run<8, 128, false, 1>("scalar KH8 TPB128 count only", d_xyz, n, d_hyp, H, d_counts, d_ands, sms);
  run<8, 128, false, 0>("scalar KH8 TPB128 fma only (+xor)", d_xyz, n, d_hyp, H, d_counts, d_ands, sms);
  run<16, 128, false, 3>("scalar KH16 TPB128 count+and", d_xyz, n, d_hyp, H, d_counts, d_ands, sms);
  return 0;
}
