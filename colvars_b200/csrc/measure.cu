// measure.cu -- measurement helpers of the C ABI: sustained DFMA throughput (the FP64 roofline
// denominator) and the accuracy self-test of the in-kernel reciprocal.
#include "engine.h"

using namespace b200cv;

extern "C" {

int b200cv_fp64_peak(int device, double target_ms, double *tflops, double *elapsed_ms)
{
  if (!tflops) return B200CV_BUG_ERROR;
#define PK(call)                                                                 \
  do {                                                                           \
    cudaError_t e_ = (cudaError_t)(call);                                        \
    if (e_ != cudaSuccess) return fail(nullptr, B200CV_ERROR, cudaGetErrorString(e_)); \
  } while (0)
  PK(cudaSetDevice(device));
  cudaDeviceProp prop;
  PK(cudaGetDeviceProperties(&prop, device));
  int const blocks = prop.multiProcessorCount * 8;
  double *d_out = nullptr;
  PK(cudaMalloc(&d_out, 64));
  cudaEvent_t e0, e1;
  PK(cudaEventCreate(&e0));
  PK(cudaEventCreate(&e1));
  auto run = [&](int iters, float &ms) -> int {
    PK(cudaEventRecord(e0, 0));
    PK(launch_fp64_peak(d_out, iters, blocks, 0));
    PK(cudaEventRecord(e1, 0));
    PK(cudaEventSynchronize(e1));
    PK(cudaEventElapsedTime(&ms, e0, e1));
    return 0;
  };
  float ms = 0.f;
  int iters = 2000;
  if (run(iters, ms)) return B200CV_ERROR;  // warm-up + calibration
  if (run(iters, ms)) return B200CV_ERROR;
  double const per_iter = ms / iters;
  iters = (int)std::max(1000.0, std::min(2.0e7, target_ms / std::max(per_iter, 1e-9)));
  if (run(iters, ms)) return B200CV_ERROR;
  double const flops = 2.0 * 256.0 * (double)iters * 256.0 * (double)blocks;
  *tflops = flops / (ms * 1e-3) / 1e12;
  if (elapsed_ms) *elapsed_ms = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return B200CV_OK;
}

int b200cv_selftest_rcp(int device, int n, double lo, double hi, double *max_ulp)
{
  if (!max_ulp || n < 1) return B200CV_BUG_ERROR;
  PK(cudaSetDevice(device));
  std::vector<double> x(n);
  // log-uniform deterministic sample
  double const llo = std::log(lo), lhi = std::log(hi);
  unsigned long long s = 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < n; i++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    double const u = (double)(s >> 11) / 9007199254740992.0;
    x[i] = std::exp(llo + u * (lhi - llo));
  }
  double *d_x = nullptr, *d_m = nullptr;
  PK(cudaMalloc(&d_x, sizeof(double) * n));
  PK(cudaMalloc(&d_m, sizeof(double)));
  PK(cudaMemcpy(d_x, x.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
  PK(cudaMemset(d_m, 0, sizeof(double)));
  PK(launch_rcp_selftest(d_x, n, d_m, 0));
  PK(cudaMemcpy(max_ulp, d_m, sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(d_x);
  cudaFree(d_m);
  return B200CV_OK;
#undef PK
}

}  // extern "C"
