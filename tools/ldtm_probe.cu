// ldtm_probe.cu — how TMEM -> register loads (tcgen05.ld 32x32b.x32) and the tensor path's epilogue arithmetic share an
// SM sub-partition: load-only throughput, arithmetic-only throughput, both from free-running warps, both software
// pipelined inside a warp. No MMA is issued: the loads read whatever the allocation holds.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ldtm_probe ldtm_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define R32(r) \
  "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), \
  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),   \
  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),  \
  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
#define RW32(r) \
  "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), \
  "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),   \
  "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),  \
  "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
__device__ __forceinline__ void ld32(unsigned (&r)[32], unsigned taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : R32(r) : "r"(taddr));
}
__device__ __forceinline__ void ld16(unsigned* r, unsigned taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
__device__ __forceinline__ void ld8(unsigned* r, unsigned taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void ld_wait(unsigned (&r)[32]) { asm volatile("tcgen05.wait::ld.sync.aligned;" : RW32(r)::"memory"); }
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) { unsigned long long d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ void math32(const unsigned (&r)[32], float C, unsigned long long (&S1)[2], unsigned long long (&S2)[2]) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float u0, u1;
    asm("add.sat.f32 %0, %1, %2;" : "=f"(u0) : "f"(-fabsf(__uint_as_float(r[i]))), "f"(C));
    asm("add.sat.f32 %0, %1, %2;" : "=f"(u1) : "f"(-fabsf(__uint_as_float(r[i + 1]))), "f"(C));
    unsigned long long U = pack2(u0, u1);
    S1[(i >> 1) & 1] = add2(S1[(i >> 1) & 1], U);
    S2[(i >> 1) & 1] = fma2(U, U, S2[(i >> 1) & 1]);
  }
}
// variant: S1 = sum u as before (FADD2), certificate sum on the ALU pipe: SI += bits(u0) + bits(u1) (one IADD3 per two evaluations)
__device__ __forceinline__ void math32_int(const unsigned (&r)[32], float C, unsigned long long (&S1)[2], unsigned (&SI)[2]) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float u0, u1;
    asm("add.sat.f32 %0, %1, %2;" : "=f"(u0) : "f"(-fabsf(__uint_as_float(r[i]))), "f"(C));
    asm("add.sat.f32 %0, %1, %2;" : "=f"(u1) : "f"(-fabsf(__uint_as_float(r[i + 1]))), "f"(C));
    unsigned long long U = pack2(u0, u1);
    S1[(i >> 1) & 1] = add2(S1[(i >> 1) & 1], U);
    SI[(i >> 1) & 1] = SI[(i >> 1) & 1] + __float_as_uint(u0) + __float_as_uint(u1);
  }
}
// MODE 0: loads only   1: arithmetic only (registers)   2: load 64 columns, wait, arithmetic on them (free-running warps)
// MODE 3: software pipeline inside the warp (load of the next 32 columns in flight under the arithmetic of the current 32)
// MODE 4: as 2, warps of a sub-partition start staggered by a quarter of the iteration
// MODE 5: as 2 with x16 loads (4 per 64 columns)
template <int MODE>
__global__ void __launch_bounds__(MODE == 11 ? 1024 : 512, 1) probe(float* out, int iters, float C, int stagger) {
  __shared__ unsigned s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = s_tmem;
  const int q = warp & 3, j = warp >> 2;
  const unsigned taddr = tmem + ((unsigned)(q * 32) << 16) + (j & 3) * 64;
  unsigned ra[32], rb[32];
  unsigned long long S1[2] = {0, 0}, S2[2] = {0, 0};
  unsigned x = 0;
  unsigned SI[2] = {0, 0};
#pragma unroll
  for (int i = 0; i < 32; ++i) { ra[i] = 0x3f000000u + i; rb[i] = 0x3e000000u + i; }
  if (MODE == 4) { long long t = clock64(); while (clock64() - t < (long long)j * stagger) {} }
  long long t0 = clock64();
  if (MODE == 3) { ld32(ra, taddr); ld_wait(ra); }
  for (int it = 0; it < iters; ++it) {
    const unsigned col = (it & 1) * 256;
    if (MODE == 0) {
      ld32(ra, taddr + col); ld32(rb, taddr + col + 32);
      ld_wait(ra); ld_wait(rb);
      x ^= ra[0] ^ rb[31];
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(ra[i]), "+r"(rb[i]));
      math32(ra, C, S1, S2); math32(rb, C, S1, S2);
    } else if (MODE == 2 || MODE == 4) {
      ld32(ra, taddr + col); ld32(rb, taddr + col + 32);
      ld_wait(ra); ld_wait(rb);
      math32(ra, C, S1, S2); math32(rb, C, S1, S2);
    } else if (MODE == 5) {
      ld16(ra, taddr + col); ld16(ra + 16, taddr + col + 16); ld16(rb, taddr + col + 32); ld16(rb + 16, taddr + col + 48);
      ld_wait(ra); ld_wait(rb);
      math32(ra, C, S1, S2); math32(rb, C, S1, S2);
    } else if (MODE == 6) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { ld8(ra + 8 * k, taddr + col + 8 * k); ld8(rb + 8 * k, taddr + col + 32 + 8 * k); }
      ld_wait(ra); ld_wait(rb);
      math32(ra, C, S1, S2); math32(rb, C, S1, S2);
    } else if (MODE == 7) {  // even column-run warps only load, odd ones only compute: do the two overlap across warps at all?
      if (j & 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(ra[i]), "+r"(rb[i]));
        math32(ra, C, S1, S2); math32(rb, C, S1, S2);
      } else {
        ld32(ra, taddr + col); ld32(rb, taddr + col + 32);
        ld_wait(ra); ld_wait(rb);
        x ^= ra[0] ^ rb[31];
      }
    } else if (MODE == 8) {  // load 32, wait, arithmetic 32 (finer interleave)
      ld32(ra, taddr + col); ld_wait(ra); math32(ra, C, S1, S2);
      ld32(rb, taddr + col + 32); ld_wait(rb); math32(rb, C, S1, S2);
    } else if (MODE == 9) {   // arithmetic only, integer certificate
#pragma unroll
      for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(ra[i]), "+r"(rb[i]));
      math32_int(ra, C, S1, SI); math32_int(rb, C, S1, SI);
    } else if (MODE == 10) {  // load 64, wait, arithmetic with the integer certificate
      ld32(ra, taddr + col); ld32(rb, taddr + col + 32);
      ld_wait(ra); ld_wait(rb);
      math32_int(ra, C, S1, SI); math32_int(rb, C, S1, SI);
    } else if (MODE == 11) {  // 32 columns per step (one register block): runs with up to 8 warps per sub-partition
      ld32(ra, taddr + col); ld_wait(ra); math32(ra, C, S1, S2);
      ld32(ra, taddr + col + 32); ld_wait(ra); math32(ra, C, S1, S2);
    } else if (MODE == 3) {
      ld32(rb, taddr + col + 32);
      math32(ra, C, S1, S2);
      ld_wait(rb);
      ld32(ra, taddr + (col ^ 256));
      math32(rb, C, S1, S2);
      ld_wait(ra);
    }
  }
  long long t1 = clock64();
  float acc = __uint_as_float(x ^ SI[0] ^ SI[1]);
#pragma unroll
  for (int k = 0; k < 2; ++k) acc += __uint_as_float((unsigned)S1[k]) + __uint_as_float((unsigned)(S2[k] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  // the slowest warp of CTA 0 defines the time
  __shared__ unsigned long long s_max;
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) atomicMax(&s_max, (unsigned long long)(t1 - t0));
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[148 * 512] = (float)s_max;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
template <int MODE>
void run(const char* name, float* out, int threads, int stagger = 0) {
  int iters = 4000;
  probe<MODE><<<148, threads>>>(out, 10, 3.f, stagger);
  cudaDeviceSynchronize();
  probe<MODE><<<148, threads>>>(out, iters, 3.f, stagger);
  cudaError_t e = cudaDeviceSynchronize();
  float cyc;
  cudaMemcpy(&cyc, out + 148 * 512, 4, cudaMemcpyDeviceToHost);
  const int wps = threads / 128;  // warps per sub-partition
  // one iteration of one warp = 64 columns x 32 lanes (8 KB of TMEM, 64 evaluations per lane)
  printf("%-44s %d warps/SMSP: %7.1f cycles per warp-iteration, %6.1f cycles per 64 columns per SMSP, %6.1f B/clk/SMSP, %.3f cycles/eval/SMSP  %s\n",
         name, wps, cyc / iters, cyc / iters / wps, 8192.0 * wps * iters / cyc, cyc / ((double)iters * 64 * wps), e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
  float* out;
  cudaMalloc(&out, (148 * 512 + 4) * 4);
  for (int threads : {128, 256, 512}) {
    run<0>("loads only (2 x32 + wait)", out, threads);
    run<1>("arithmetic only", out, threads);
    run<2>("load 64, wait, arithmetic (free running)", out, threads);
    run<3>("software pipelined inside the warp", out, threads);
  }
  run<5>("x16 loads (4 per 64 columns)", out, 512);
  run<6>("x8 loads (8 per 64 columns)", out, 512);
  run<7>("2 warps load only + 2 warps arithmetic only", out, 512);
  run<8>("load 32, wait, arithmetic 32", out, 512);
  run<11>("load 32, wait, arithmetic 32, one register block", out, 512);
  run<11>("load 32, wait, arithmetic 32, one register block", out, 768);
  run<11>("load 32, wait, arithmetic 32, one register block", out, 1024);
  run<9>("arithmetic only, IADD3 certificate", out, 512);
  run<10>("load 64, wait, arithmetic, IADD3 certificate", out, 512);
  run<4>("free running, staggered start 100", out, 512, 100);
  run<4>("free running, staggered start 300", out, 512, 300);
  return 0;
}
