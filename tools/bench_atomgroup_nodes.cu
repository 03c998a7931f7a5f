// bench_atomgroup_nodes.cu -- times the atom-group kernel boundary of the reference
// (src/cuda/colvaratoms_kernel.h: one graph kernel node per call) for whichever implementation
// is linked: built twice by tools/Makefile,
//   _build/bench_agnodes_ref   reference archive only (member from src/cuda/colvaratoms_kernel.cu)
//   _build/bench_agnodes_b200  colvars_b200/host/atomgroup_b200.cu linked in front of it.
// Same buffers, same calls; each graph launch is timed alone with CUDA events after a 256 MiB
// write that evicts L2.  A checksum of every output is printed so the two runs can be compared.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <numeric>
#include <random>
#include <vector>

#include "colvarmodule.h"
#include "colvartypes.h"
#include "cuda/colvaratoms_kernel.h"
#include "cuda/colvartypes_kernel.h"

extern "C" int b200cv_atomgroup_kernels_linked(void) __attribute__((weak));

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (cudaError_t)(x);                                                     \
    if (e_ != cudaSuccess) {                                                               \
      std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

static char *g_flush;
static size_t const g_flush_bytes = 256u << 20;

static float time_graph(cudaGraph_t g, int reps)
{
  cudaGraphExec_t ex;
  cudaGraphInstantiate(&ex, g, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  std::vector<float> ms;
  for (int r = 0; r < reps + 3; r++) {
    cudaMemsetAsync(g_flush, r, g_flush_bytes, 0);
    cudaEventRecord(e0, 0);
    cudaGraphLaunch(ex, 0);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float t;
    cudaEventElapsedTime(&t, e0, e1);
    if (r >= 3) ms.push_back(t);
  }
  cudaGraphExecDestroy(ex);
  std::sort(ms.begin(), ms.end());
  return 1e3f * ms[ms.size() / 2];
}

static double checksum(const double *d, size_t n)
{
  std::vector<double> h(n);
  cudaMemcpy(h.data(), d, sizeof(double) * n, cudaMemcpyDeviceToHost);
  long double s = 0;
  for (size_t k = 0; k < n; k++) s += (long double)h[k] * (long double)(1 + k % 7);
  return (double)s;
}

int main(int argc, char **argv)
{
  unsigned int const n = argc > 1 ? std::atoi(argv[1]) : 1000000;
  unsigned int const P = n + n / 10;
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), (int)(P - n));
  std::mt19937 rng(7);
  if (argc > 2 && std::atoi(argv[2]) != 0) std::shuffle(idx.begin(), idx.end(), rng);
  std::vector<double> h_proxy(3 * (size_t)P), h_mass(n);
  for (size_t k = 0; k < h_proxy.size(); k++) h_proxy[k] = 1e-3 * (double)(k % 100003);
  for (unsigned int k = 0; k < n; k++) h_mass[k] = 1.0 + (k % 5);
  double const total_mass = std::accumulate(h_mass.begin(), h_mass.end(), 0.0);

  double *d_proxy, *d_force, *d_pos, *d_grad, *d_mass, *d_scalar;
  int *d_idx;
  unsigned int *d_tb;
  cvm::rvector *d_rv, *h_cog, *h_com;
  cvm::quaternion *d_q;
  CK(cudaMalloc(&d_proxy, sizeof(double) * 3 * P));
  CK(cudaMalloc(&d_force, sizeof(double) * 3 * P));
  CK(cudaMalloc(&d_pos, sizeof(double) * 3 * n));
  CK(cudaMalloc(&d_grad, sizeof(double) * 3 * n));
  CK(cudaMalloc(&d_mass, sizeof(double) * n));
  CK(cudaMalloc(&d_scalar, sizeof(double)));
  CK(cudaMalloc(&d_idx, sizeof(int) * n));
  CK(cudaMalloc(&d_tb, 4));
  CK(cudaMalloc(&d_rv, sizeof(cvm::rvector) * 8));
  CK(cudaMalloc(&d_q, sizeof(cvm::quaternion)));
  CK(cudaMalloc(&g_flush, g_flush_bytes));
  CK(cudaHostAlloc(&h_cog, sizeof(cvm::rvector), cudaHostAllocMapped));
  CK(cudaHostAlloc(&h_com, sizeof(cvm::rvector), cudaHostAllocMapped));
  CK(cudaMemset(d_tb, 0, 4));
  CK(cudaMemset(d_rv, 0, sizeof(cvm::rvector) * 8));
  CK(cudaMemset(d_force, 0, sizeof(double) * 3 * P));
  CK(cudaMemcpy(d_proxy, h_proxy.data(), sizeof(double) * 3 * P, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_mass, h_mass.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_idx, idx.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
  double const fscalar = -0.37;
  CK(cudaMemcpy(d_scalar, &fscalar, sizeof(double), cudaMemcpyHostToDevice));
  cvm::quaternion hq(std::cos(0.3), std::sin(0.3) * 0.6, std::sin(0.3) * 0.0, std::sin(0.3) * 0.8);
  CK(cudaMemcpy(d_q, &hq, sizeof(hq), cudaMemcpyHostToDevice));

  int const reps = 21;
  std::printf("atom-group kernel boundary: %s implementation, %u atoms of a %u-atom proxy\n",
              b200cv_atomgroup_kernels_linked ? "b200" : "reference", n, P);
  std::printf("%-44s %10s   %s\n", "call", "us", "checksum of the output");

  auto one = [&](const char *name, auto add, auto sum) {
    cudaGraph_t g;
    cudaGraphCreate(&g, 0);
    cudaGraphNode_t node;
    int const rc = add(node, g);
    float const us = rc == 0 ? time_graph(g, reps) : -1.f;
    cudaGraphDestroy(g);
    std::printf("%-44s %10.2f   %.15e\n", name, us, sum());
  };

  one("atoms_pos_from_proxy (gather)",
      [&](cudaGraphNode_t &nd, cudaGraph_t &g) {
        return colvars_gpu::atoms_pos_from_proxy(d_idx, d_proxy, d_pos, n, P, nd, g, {});
      },
      [&]() { return checksum(d_pos, 3 * (size_t)n); });
  one("atoms_calc_cog_com",
      [&](cudaGraphNode_t &nd, cudaGraph_t &g) {
        return colvars_gpu::atoms_calc_cog_com(d_pos, d_mass, n, d_rv + 0, d_rv + 1, d_rv + 2,
                                               d_rv + 3, d_rv + 4, h_cog, h_com, total_mass, d_tb,
                                               nd, g, {});
      },
      [&]() { return h_cog->x + 2 * h_cog->y + 3 * h_cog->z + 5 * h_com->x + 7 * h_com->y + 11 * h_com->z; });
  one("atoms_calc_cog",
      [&](cudaGraphNode_t &nd, cudaGraph_t &g) {
        return colvars_gpu::atoms_calc_cog(d_pos, n, d_rv + 0, d_rv + 1, nullptr, h_cog, d_tb, nd,
                                           g, {});
      },
      [&]() { return h_cog->x + 2 * h_cog->y + 3 * h_cog->z; });
  // gradients := positions (any data will do), then the scatters
  CK(cudaMemcpy(d_grad, d_pos, sizeof(double) * 3 * n, cudaMemcpyDeviceToDevice));
  one("apply_main_colvar_force_to_proxy",
      [&](cudaGraphNode_t &nd, cudaGraph_t &g) {
        return colvars_gpu::apply_main_colvar_force_to_proxy(d_idx, d_force, d_grad, d_scalar,
                                                             false, nullptr, n, P, nd, g, {});
      },
      [&]() { return checksum(d_force, 3 * (size_t)P) / (reps + 3); });
  CK(cudaMemset(d_force, 0, sizeof(double) * 3 * P));
  one("apply_main_colvar_force_to_proxy (rotated)",
      [&](cudaGraphNode_t &nd, cudaGraph_t &g) {
        return colvars_gpu::apply_main_colvar_force_to_proxy(d_idx, d_force, d_grad, d_scalar,
                                                             true, d_q, n, P, nd, g, {});
      },
      [&]() { return checksum(d_force, 3 * (size_t)P) / (reps + 3); });
  CK(cudaMemset(d_force, 0, sizeof(double) * 3 * P));
  one("apply_force",
      [&](cudaGraphNode_t &nd, cudaGraph_t &g) {
        return colvars_gpu::apply_force(d_grad, d_idx, d_force, n, P, nd, g, {});
      },
      [&]() { return checksum(d_force, 3 * (size_t)P) / (reps + 3); });
  // optimal rotation: correlation matrix of (positions, gradients-as-reference) + 4x4 eigenproblem
  {
    double *d_S, *d_Sv, *d_L;
    cvm::rmatrix *h_C;
    cvm::quaternion *d_qq, *d_qold;
    int *h_flags;
    CK(cudaMalloc(&d_S, 16 * sizeof(double)));
    CK(cudaMalloc(&d_Sv, 16 * sizeof(double)));
    CK(cudaMalloc(&d_L, 4 * sizeof(double)));
    CK(cudaMalloc(&d_qq, sizeof(cvm::quaternion)));
    CK(cudaMalloc(&d_qold, sizeof(cvm::quaternion)));
    CK(cudaMemset(d_qold, 0, sizeof(cvm::quaternion)));
    CK(cudaHostAlloc(&h_C, sizeof(cvm::rmatrix), cudaHostAllocMapped));
    CK(cudaHostAlloc(&h_flags, 2 * sizeof(int), cudaHostAllocMapped));
    h_flags[0] = h_flags[1] = 0;
    auto lead = [&]() {
      double v[4], l[4];
      cudaMemcpy(v, d_Sv, sizeof(v), cudaMemcpyDeviceToHost);
      cudaMemcpy(l, d_L, sizeof(l), cudaMemcpyDeviceToHost);
      double const sg = v[0] < 0 ? -1.0 : 1.0;  // the sign of an eigenvector is arbitrary
      return l[0] + sg * (v[0] + 2 * v[1] + 3 * v[2] + 5 * v[3]);
    };
    one("memset S + build_overlapping_matrix",
        [&](cudaGraphNode_t &nd, cudaGraph_t &g) {
          cudaGraphNode_t ms;
          cudaMemsetParams mp = {0};
          mp.dst = d_S; mp.value = 0; mp.elementSize = 4; mp.width = 32; mp.height = 1; mp.pitch = 128;
          if (cudaGraphAddMemsetNode(&ms, g, nullptr, 0, &mp) != cudaSuccess) return 1;
          return colvars_gpu::build_overlapping_matrix(d_pos, d_pos + n, d_pos + 2 * n, d_grad,
                                                       d_grad + n, d_grad + 2 * n, d_S, d_Sv, h_C,
                                                       d_tb, n, nd, g, {ms});
        },
        [&]() { return checksum(d_S, 16); });
    one("jacobi_4x4",
        [&](cudaGraphNode_t &nd, cudaGraph_t &g) {
          // the eigen-solver overwrites its input: restore it from S first
          cudaGraphNode_t cp;
          if (cudaGraphAddMemcpyNode1D(&cp, g, nullptr, 0, d_Sv, d_S, 16 * sizeof(double),
                                       cudaMemcpyDeviceToDevice) != cudaSuccess) return 1;
          return colvars_gpu::jacobi_4x4(d_Sv, d_L, h_flags, d_qq, false, 1e-2, d_qold,
                                         h_flags + 1, nd, g, {cp});
        },
        lead);
  }
  // in-place transforms last: they change d_pos on every launch
  one("apply_translation",
      [&](cudaGraphNode_t &nd, cudaGraph_t &g) {
        return colvars_gpu::apply_translation(d_pos, -1.0, d_rv + 1, n, nd, g, {});
      },
      [&]() { return checksum(d_pos, 3 * (size_t)n); });
  one("rotate_with_quaternion",
      [&](cudaGraphNode_t &nd, cudaGraph_t &g) {
        return colvars_gpu::rotate_with_quaternion(d_pos, d_q, n, nd, g, {});
      },
      [&]() { return checksum(d_pos, 3 * (size_t)n); });
  return 0;
}
