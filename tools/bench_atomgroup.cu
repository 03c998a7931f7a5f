// bench_atomgroup.cu -- the O(N) atom-group kernels of this repository next to the reference's
// own CUDA kernels (src/cuda/colvaratoms_kernel.cu, unmodified, linked from
// oracle/_ref/libcolvars_ref_cuda.a) on the same device and the same buffers:
//   gather      atoms_pos_from_proxy_kernel (src/cuda/colvaratoms_kernel.cu:18-34)
//   COM + COG   atoms_calc_cog_com_kernel   (:134-243)
//   scatter     apply_colvar_force_to_proxy_kernel (:434-474)
// Each launch is timed alone with CUDA events after a 256 MiB write that evicts L2; reported
// against the algorithmic HBM bytes: gather 4 + 24 + 24 B/atom, COM+COG 32 B/atom,
// scatter 4 + 24 B read + 3 FP64 atomics (24 B read + 24 B written) per atom.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <numeric>
#include <random>
#include <vector>

#include "colvarmodule.h"
#include "cuda/colvaratoms_kernel.h"

#include "../colvars_b200/csrc/kernels.cuh"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (cudaError_t)(x);                                             \
    if (e_ != cudaSuccess) {                                                       \
      std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

template <typename F> static float time_launches(F launch, int reps, char *flush, size_t flush_bytes)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  std::vector<float> ms;
  for (int r = 0; r < reps + 3; r++) {
    cudaMemsetAsync(flush, r, flush_bytes, 0);
    cudaEventRecord(e0, 0);
    launch();
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float t;
    cudaEventElapsedTime(&t, e0, e1);
    if (r >= 3) ms.push_back(t);
  }
  std::sort(ms.begin(), ms.end());
  return ms[ms.size() / 2];
}

static cudaGraphExec_t make_exec(cudaGraph_t g)
{
  cudaGraphExec_t ex;
  cudaGraphInstantiate(&ex, g, 0);
  return ex;
}

int main(int argc, char **argv)
{
  int const n = argc > 1 ? std::atoi(argv[1]) : 1000000;
  int const P = n + n / 10;
  bool const shuffled = argc > 2 && std::atoi(argv[2]) != 0;
  size_t const stride = (size_t)(n + 31) / 32 * 32 + 32;
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), P - n);
  if (shuffled) {
    std::mt19937 rng(7);
    std::shuffle(idx.begin(), idx.end(), rng);
  }
  std::vector<double> h_proxy(3 * (size_t)P), h_mass(n, 1.0);
  for (size_t k = 0; k < h_proxy.size(); k++) h_proxy[k] = 1e-3 * (double)(k % 100003);
  double *d_proxy, *d_force, *d_pos, *d_pos_ref, *d_grad, *d_mass, *d_small, *d_scratch;
  int *d_idx;
  unsigned int *d_ticket, *d_tb;
  char *flush;
  size_t const flush_bytes = 256u << 20;
  CK(cudaMalloc(&d_proxy, sizeof(double) * 3 * P));
  CK(cudaMalloc(&d_force, sizeof(double) * 3 * P));
  CK(cudaMalloc(&d_pos, sizeof(double) * 3 * stride));
  CK(cudaMalloc(&d_pos_ref, sizeof(double) * 3 * n));
  CK(cudaMalloc(&d_grad, sizeof(double) * 3 * stride));
  CK(cudaMalloc(&d_mass, sizeof(double) * n));
  CK(cudaMalloc(&d_small, sizeof(double) * 64));
  CK(cudaMalloc(&d_scratch, sizeof(double) * b200cv::COM_SCRATCH_DOUBLES));
  CK(cudaMalloc(&d_idx, sizeof(int) * n));
  CK(cudaMalloc(&d_ticket, 4));
  CK(cudaMalloc(&d_tb, 4));
  CK(cudaMalloc(&flush, flush_bytes));
  CK(cudaMemset(d_ticket, 0, 4));
  CK(cudaMemset(d_tb, 0, 4));
  CK(cudaMemset(d_small, 0, sizeof(double) * 64));
  CK(cudaMemset(d_force, 0, sizeof(double) * 3 * P));
  CK(cudaMemcpy(d_proxy, h_proxy.data(), sizeof(double) * 3 * P, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_mass, h_mass.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_idx, idx.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_grad, 0, sizeof(double) * 3 * stride));
  double *h_force_scalar;  // the reference kernel reads the scalar force from pinned host memory
  CK(cudaHostAlloc(&h_force_scalar, sizeof(double), cudaHostAllocMapped));
  *h_force_scalar = -0.37;
  cvm::rvector *h_cog, *h_com;
  CK(cudaHostAlloc(&h_cog, sizeof(cvm::rvector), cudaHostAllocMapped));
  CK(cudaHostAlloc(&h_com, sizeof(cvm::rvector), cudaHostAllocMapped));
  cvm::rvector *d_rv;  // cog_tmp, cog_out, cog_orig, com_tmp, com_out
  CK(cudaMalloc(&d_rv, sizeof(cvm::rvector) * 8));
  CK(cudaMemset(d_rv, 0, sizeof(cvm::rvector) * 8));

  int const reps = 21;
  std::printf("atom-group kernels, %d atoms of a %d-atom proxy, %s indices\n", n, P,
              shuffled ? "shuffled" : "contiguous");
  std::printf("%-34s %10s %10s   %s\n", "kernel", "ours us", "ref us", "ours GB/s (algorithmic bytes)");

  // ---- gather -------------------------------------------------------------------------
  {
    cudaGraph_t g;
    cudaGraphCreate(&g, 0);
    cudaGraphNode_t node;
    colvars_gpu::atoms_pos_from_proxy(d_idx, d_proxy, d_pos_ref, n, P, node, g, {});
    cudaGraphExec_t ex = make_exec(g);
    float const t_ref = time_launches([&]() { cudaGraphLaunch(ex, 0); }, reps, flush, flush_bytes);
    float const t_our = time_launches(
        [&]() { b200cv::launch_gather(d_proxy, P, 0, d_idx, n, d_pos, stride, 0); }, reps, flush,
        flush_bytes);
    std::printf("%-34s %10.2f %10.2f   %.0f\n", "gather (read_positions)", 1e3 * t_our, 1e3 * t_ref,
                52.0 * n / (t_our * 1e-3) / 1e9);
  }
  // ---- COM + COG -----------------------------------------------------------------------
  {
    cudaGraph_t g;
    cudaGraphCreate(&g, 0);
    cudaGraphNode_t node;
    colvars_gpu::atoms_calc_cog_com(d_pos_ref, d_mass, n, d_rv + 0, d_rv + 1, d_rv + 2, d_rv + 3,
                                    d_rv + 4, h_cog, h_com, (double)n, d_tb, node, g, {});
    cudaGraphExec_t ex = make_exec(g);
    float const t_ref = time_launches([&]() { cudaGraphLaunch(ex, 0); }, reps, flush, flush_bytes);
    float const t_our = time_launches(
        [&]() {
          b200cv::launch_com_cog(d_pos, stride, d_mass, (double)n, n, d_small, d_small + 3, d_scratch,
                                 d_ticket, 0);
        },
        reps, flush, flush_bytes);
    double ours[6];
    cudaMemcpy(ours, d_small, sizeof(ours), cudaMemcpyDeviceToHost);
    std::printf("%-34s %10.2f %10.2f   %.0f      (cog.x ours %.12g ref %.12g)\n",
                "COM + COG", 1e3 * t_our, 1e3 * t_ref, 32.0 * n / (t_our * 1e-3) / 1e9, ours[3],
                h_cog->x);
  }
  // ---- force scatter ------------------------------------------------------------------
  {
    cudaGraph_t g;
    cudaGraphCreate(&g, 0);
    cudaGraphNode_t node;
    colvars_gpu::apply_main_colvar_force_to_proxy(d_idx, d_force, d_pos_ref, h_force_scalar, false,
                                                  nullptr, n, P, node, g, {});
    cudaGraphExec_t ex = make_exec(g);
    float const t_ref = time_launches([&]() { cudaGraphLaunch(ex, 0); }, reps, flush, flush_bytes);
    float const t_our = time_launches(
        [&]() {
          b200cv::launch_scatter_force(d_pos, stride, d_idx, n, -0.37, nullptr, d_force, P, 0, 0);
        },
        reps, flush, flush_bytes);
    std::printf("%-34s %10.2f %10.2f   %.0f\n", "force scatter (apply_colvar_force)", 1e3 * t_our,
                1e3 * t_ref, 76.0 * n / (t_our * 1e-3) / 1e9);
  }
  return 0;
}
