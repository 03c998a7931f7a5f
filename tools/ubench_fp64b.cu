// ubench_fp64b.cu -- issue cost of FP64 instructions on sm_100a as a function of how many
// distinct register operands they read.  Aggregate cycles per instruction per SM
// sub-partition (event time at 1.965 GHz, 8 warps per sub-partition).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void __launch_bounds__(128) k(double *out, int iters, double cc)
{
  double a[16], b[16], g[8], d[8];
#pragma unroll
  for (int q = 0; q < 16; q++) { a[q] = threadIdx.x * 1e-3 + q; b[q] = 1.0 + 1e-7 * (q + threadIdx.x); }
#pragma unroll
  for (int q = 0; q < 8; q++) { g[q] = 1.0 + 1e-9 * q * threadIdx.x; d[q] = 1e-9 * (q + 1) * threadIdx.x; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int q = 0; q < 16; q++) {
      if (MODE == 0) a[q] = fma(a[q], cc, cc);                  // 1 register + 2 uniform/const
      if (MODE == 1) a[q] = a[q] * b[q];                        // DMUL, 2 distinct regs
      if (MODE == 2) a[q] = a[q] + b[q];                        // DADD, 2 distinct regs
      if (MODE == 3) a[q] = fma(b[q], b[q], a[q]);              // DFMA, 2 distinct regs
      if (MODE == 4) a[q] = fma(a[q], b[q], b[(q + 5) & 15]);   // DFMA, 3 distinct regs
      if (MODE == 5) a[q] = fma(g[q >> 1], d[q >> 1], a[q]);    // pairs of DFMA sharing 2 sources
      if (MODE == 6) a[q] = fma(a[q], b[q], cc);                // DFMA 2 regs + const
      if (MODE == 7) a[q] = fma(g[q & 7], d[(q * 3) & 7], a[q]);  // 3 regs, sources rarely repeat
    }
  }
  double s = 0;
#pragma unroll
  for (int q = 0; q < 16; q++) s += a[q] + b[q];
#pragma unroll
  for (int q = 0; q < 8; q++) s += g[q] + d[q];
  if (s == 1234.5) out[0] = s;
}

template <int MODE> void run(const char *name)
{
  double *out; cudaMalloc(&out, 8);
  int iters = 20000, w = 8, blocks = 148 * w;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 128>>>(out, 100, 1e-9);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, 128>>>(out, iters, 1e-9);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-52s cycles per instr per SMSP = %.3f\n", name, ms * 1e-3 * 1.965e9 / ((double)iters * 16 * w));
  cudaFree(out);
}

int main()
{
  run<0>("DFMA reg, const, const");
  run<1>("DMUL 2 distinct regs");
  run<2>("DADD 2 distinct regs");
  run<3>("DFMA b,b,a (2 distinct regs)");
  run<4>("DFMA 3 distinct regs");
  run<5>("DFMA g,d,a pairs sharing g,d");
  run<6>("DFMA 2 regs + const");
  run<7>("DFMA g,d,a sources from small sets");
  return 0;
}
