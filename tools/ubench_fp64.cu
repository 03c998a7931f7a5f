// ubench_fp64.cu -- FP64 pipe microbenchmarks for sm_100a: how does DFMA throughput react to
// (a) distinct source registers, (b) interleaved integer-pipe instructions, (c) MUFU.RCP64H,
// (d) LDS / SHFL in the stream.  Prints cycles per warp-level DFMA per SM sub-partition.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench_fp64 ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NALU, int NMUFU, int NSHFL, bool DISTINCT>
__global__ void __launch_bounds__(128) k(double *out, int iters, long long *cyc)
{
  double a[16];
  double b[16];
  int ia[16];
#pragma unroll
  for (int q = 0; q < 16; q++) {
    a[q] = threadIdx.x * 1e-3 + q;
    b[q] = 1.0 + 1e-7 * (q + threadIdx.x);
    ia[q] = threadIdx.x + q;
  }
  double c = 1e-9 * threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int q = 0; q < 16; q++) {
      if (DISTINCT) a[q] = fma(a[q], b[q], b[(q + 5) & 15]);
      else a[q] = fma(a[q], c, c);
      if (q < NALU) ia[q] = (ia[q] + 0x3fee) ^ (ia[q] >> 3);
      if (q < NMUFU) {
        double y;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b[q]));
        b[q] = y;
      }
      if (q < NSHFL) ia[(q + 8) & 15] = __shfl_sync(0xffffffffu, ia[(q + 8) & 15], (threadIdx.x + 1) & 31);
    }
  }
  long long t1 = clock64();
  double s = 0;
  int si = 0;
#pragma unroll
  for (int q = 0; q < 16; q++) { s += a[q] + b[q]; si += ia[q]; }
  if (s == 1234.5 || si == 77) out[0] = s + si;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NALU, int NMUFU, int NSHFL, bool DISTINCT> void run(const char *name, int warps_per_smsp)
{
  double *out; long long *cyc;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  int iters = 20000;
  int blocks = 148 * warps_per_smsp;  // 128-thread blocks: one warp per SMSP each
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NALU, NMUFU, NSHFL, DISTINCT><<<blocks, 128>>>(out, 100, cyc);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<NALU, NMUFU, NSHFL, DISTINCT><<<blocks, 128>>>(out, iters, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  // aggregate: warp-level DFMAs issued per SMSP per cycle, from the event time at the
  // clock implied by block 0's own cycle counter
  double clk_ghz = (double)c / (ms * 1e6);
  double dfma_per_smsp = (double)iters * 16 * warps_per_smsp;
  printf("%-44s w/SMSP=%d  block0 cycles/iter %.1f | kernel %.3f ms (%.2f GHz if block0 spans it) "
         "-> cycles per DFMA per SMSP (event time @1.965GHz) = %.3f\n", name, warps_per_smsp,
         (double)c / iters, ms, clk_ghz, ms * 1e-3 * 1.965e9 / dfma_per_smsp);
  cudaFree(out); cudaFree(cyc);
}

int main()
{
  for (int w : {1, 4, 8, 12}) {
    run<0, 0, 0, false>("DFMA reuse operands", w);
    run<0, 0, 0, true>("DFMA distinct operands", w);
    run<4, 0, 0, true>("DFMA distinct + 4 ALU-pairs /16", w);
    run<8, 0, 0, true>("DFMA distinct + 8 ALU-pairs /16", w);
    run<16, 0, 0, true>("DFMA distinct + 16 ALU-pairs /16", w);
    run<0, 2, 0, true>("DFMA distinct + 2 MUFU.RCP64H /16", w);
    run<0, 0, 2, true>("DFMA distinct + 2 SHFL /16", w);
    run<8, 1, 1, true>("DFMA distinct + 8 ALUp + 1 MUFU + 1 SHFL", w);
  }
  return 0;
}
