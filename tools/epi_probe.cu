// epi_probe.cu — stand-alone throughput of the tensor path's epilogue arithmetic (csrc/plane_tc.cu):
// per evaluation one FADD.SAT, half a FADD2 and half an FFMA2, on register operands, 4 warps per SM sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o epi_probe epi_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) { unsigned long long d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE, int NACC>
__global__ void __launch_bounds__(512, 1) probe(const float* in, float* out, int iters, float C) {
  unsigned r[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) r[i] = __float_as_uint(in[(threadIdx.x * 64 + i) & 4095]);
  unsigned long long S1[NACC], S2[NACC];
  float T1[4] = {0, 0, 0, 0}, T2[4] = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < NACC; ++k) S1[k] = S2[k] = 0ull;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 64; ++i) asm volatile("" : "+r"(r[i]));
#pragma unroll
    for (int i = 0; i < 64; i += 2) {
      float u0, u1;
      asm("add.sat.f32 %0, %1, %2;" : "=f"(u0) : "f"(-fabsf(__uint_as_float(r[i]))), "f"(C));
      asm("add.sat.f32 %0, %1, %2;" : "=f"(u1) : "f"(-fabsf(__uint_as_float(r[i + 1]))), "f"(C));
      if (MODE == 0) {  // packed
        unsigned long long U = pack2(u0, u1);
        S1[(i >> 1) % NACC] = add2(S1[(i >> 1) % NACC], U);
        S2[(i >> 1) % NACC] = fma2(U, U, S2[(i >> 1) % NACC]);
      } else if (MODE == 1) {  // scalar
        T1[i & 3] = __fadd_rn(T1[i & 3], u0); T1[(i + 1) & 3] = __fadd_rn(T1[(i + 1) & 3], u1);
        T2[i & 3] = __fmaf_rn(u0, u0, T2[i & 3]); T2[(i + 1) & 3] = __fmaf_rn(u1, u1, T2[(i + 1) & 3]);
      } else {  // sat + packed sum only (no sum of squares)
        unsigned long long U = pack2(u0, u1);
        S1[(i >> 1) % NACC] = add2(S1[(i >> 1) % NACC], U);
      }
    }
  }
  long long t1 = clock64();
  float acc = T1[0] + T1[1] + T1[2] + T1[3] + T2[0] + T2[1] + T2[2] + T2[3];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc += __uint_as_float((unsigned)S1[k]) + __uint_as_float((unsigned)(S2[k] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[148 * 512] = (float)(t1 - t0);
}
template <int MODE, int NACC>
void run(const char* name, float* in, float* out, int threads) {
  int iters = 2000;
  probe<MODE, NACC><<<148, threads>>>(in, out, 10, 3.f);
  cudaDeviceSynchronize();
  probe<MODE, NACC><<<148, threads>>>(in, out, iters, 3.f);
  cudaDeviceSynchronize();
  float cyc;
  cudaMemcpy(&cyc, out + 148 * 512, 4, cudaMemcpyDeviceToHost);
  double evals_per_smsp = (double)iters * 64 * (threads / 32 / 4);
  printf("%-34s threads %3d: %.3f cycles per evaluation per lane (SMSP)\n", name, threads, cyc / evals_per_smsp);
}
int main() {
  float *in, *out;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, (148 * 512 + 4) * 4);
  cudaMemset(in, 0x3f, 4096 * 4);
  for (int threads : {128, 256, 512}) {
    run<0, 2>("packed FADD2/FFMA2, 2 acc pairs", in, out, threads);
    run<0, 4>("packed FADD2/FFMA2, 4 acc pairs", in, out, threads);
    run<1, 2>("scalar FADD/FFMA", in, out, threads);
    run<2, 2>("sat + packed sum only", in, out, threads);
  }
  return 0;
}
