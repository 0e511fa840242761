// pipe_probe.cu — issue/pipe micro-benchmark for sm_100a: do the FMA pipe (FFMA, FFMA2) and the ALU pipe
// (LOP3, LEA.HI / SHF / IADD3) overlap, and what are their rates? Prints warp-instructions per cycle per SMSP.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

#define U 8
template <int MODE>
__global__ void __launch_bounds__(256) probe(unsigned* out, int iters, unsigned seed, float fs) {
  unsigned long long f[U];
  unsigned a[U], c[U];
  float g[U];
#pragma unroll
  for (int i = 0; i < U; ++i) {
    f[i] = ((unsigned long long)__float_as_uint(fs + i) << 32) | __float_as_uint(fs * 0.5f + i);
    a[i] = seed * (i + 1) + threadIdx.x;
    c[i] = 0;
    g[i] = fs + i;
  }
  const unsigned long long M = ((unsigned long long)__float_as_uint(1.0000001f) << 32) | __float_as_uint(0.9999999f);
  const unsigned long long A = ((unsigned long long)__float_as_uint(1e-7f) << 32) | __float_as_uint(-1e-7f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (MODE == 0 || MODE == 3 || MODE == 5 || MODE == 7) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(f[i]) : "l"(M), "l"(A));
      if (MODE == 1 || MODE == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(seed), "r"(c[(i + 1) % U]));
      if (MODE == 2 || MODE == 5) c[i] += a[i] >> 31;  // LEA.HI (or SHF+IADD)
      if (MODE == 4 || MODE == 6) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(g[i]) : "f"(1.0000001f), "f"(1e-7f));
      if (MODE == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(seed), "r"(c[(i + 1) % U]));
      if (MODE == 7) {  // 2 ALU per FFMA2
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(seed), "r"(c[(i + 1) % U]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(c[i]) : "r"(seed), "r"(a[(i + 1) % U]));
      }
      if (MODE == 8) {  // FSETP + predicated IADD (old counting)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(c[i]) : "f"(g[i]), "f"(fs));
      }
    }
  }
  unsigned r = 0;
#pragma unroll
  for (int i = 0; i < U; ++i) r ^= (unsigned)f[i] ^ (unsigned)(f[i] >> 32) ^ a[i] ^ c[i] ^ __float_as_uint(g[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, int instr_per_iter_per_chain, int sms, int warps_per_smsp) {
  int tpb = 256;
  int blocks = sms * (warps_per_smsp * 4 * 32 / tpb);
  unsigned* d;
  cudaMalloc(&d, (size_t)blocks * tpb * 4);
  int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    probe<MODE><<<blocks, tpb>>>(d, iters, 12345u, 1.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cycles = best * 1e-3 * clk * 1e3;
  double winstr = (double)iters * U * instr_per_iter_per_chain * warps_per_smsp;  // per SMSP
  printf("%-34s warps/SMSP %d: %.3f ms, %.3f warp-instr/cycle/SMSP (assuming %d kHz)\n", name, warps_per_smsp, best, winstr / cycles, clk);
  cudaFree(d);
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int w : {4, 8}) {
    run<0>("FFMA2 only", 1, sms, w);
    run<4>("FFMA only", 1, sms, w);
    run<1>("LOP3 only", 1, sms, w);
    run<2>("x>>31 accumulate only", 1, sms, w);
    run<3>("FFMA2 + LOP3 (1:1)", 2, sms, w);
    run<7>("FFMA2 + 2 LOP3 (1:2)", 3, sms, w);
    run<5>("FFMA2 + x>>31 acc (1:1)", 2, sms, w);
    run<6>("FFMA + LOP3 (1:1)", 2, sms, w);
    run<8>("FSETP + @p IADD", 2, sms, w);
  }
  return 0;
}
