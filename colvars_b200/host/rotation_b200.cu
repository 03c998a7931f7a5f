// rotation_b200.cu -- B200 kernels behind the reference's optimal-rotation kernel boundary.
//
// Companion of atomgroup_b200.cu: the three remaining free functions that the `smp gpu` graphs
// of a fitted atom group are assembled from,
//   colvars_gpu::build_overlapping_matrix, colvars_gpu::jacobi_4x4
//       (src/cuda/colvartypes_kernel.h:11-38; callers src/colvartypes.cpp:549-611),
//   colvars_gpu::prepare_derivative
//       (src/cuda/colvar_rotation_derivative_kernel.h:12-20; caller
//        src/colvar_rotation_derivative.cpp:48-60).
// Linked in front of the reference archive they supersede the members built from
// src/cuda/colvartypes_kernel.cu and src/cuda/colvar_rotation_derivative_kernel.cu
// (correlation matrix, 4x4 eigenproblem, derivative tables).
//
// Contract kept (what the unmodified host and rotation_derivative_gpu's device methods read):
//  * S (16 doubles) is zero on entry (the caller's memset node) and holds the symmetric overlap
//    matrix on exit; S_eigvec starts as a copy of it; h_C (mapped host) receives C; the ticket is
//    zero again on exit;
//  * jacobi_4x4 replaces S_eigvec by the eigenvectors as ROWS, S_eigval by the eigenvalues in the
//    same order, the largest first; q = that leading eigenvector; max_reached / crossing flags
//    as in the reference;
//  * prepare_derivative fills tmp_Q0Q0 (4x4) and tmp_Q0Q0_L (4x4x4) from them.
// Different inside: the correlation sums run on every SM (296 CTAs x 256 threads, shuffle tree,
// one atomic per CTA and sum; the reference caps them at 64 CTAs of 128 threads), and the
// eigenproblem is a cyclic Jacobi iteration in one thread with a RELATIVE convergence test
// (off-diagonal norm <= 1e-15 of the matrix norm; the reference stops at an absolute 1e-8).
#include "colvar_gpu_support.h"
#include "colvarmodule.h"
#include "colvartypes.h"
#include "colvar_rotation_derivative.h"
#include "cuda/colvartypes_kernel.h"
#include "cuda/colvar_rotation_derivative_kernel.h"

#include <algorithm>
#include <vector>

#if defined(COLVARS_CUDA)

extern "C" int b200cv_rotation_kernels_linked(void) { return 1; }

namespace {

constexpr unsigned int kThreads = 256;
constexpr unsigned int kReduceCtas = 296;

int add_kernel(void *func, unsigned int ctas, unsigned int threads, void **args,
               cudaGraphNode_t &node, cudaGraph_t &graph, const std::vector<cudaGraphNode_t> &deps)
{
  cudaKernelNodeParams p = {0};
  p.func = func;
  p.gridDim = dim3(ctas, 1, 1);
  p.blockDim = dim3(threads, 1, 1);
  p.sharedMemBytes = 0;
  p.kernelParams = args;
  p.extra = NULL;
  return checkGPUError(cudaGraphAddKernelNode(&node, graph, deps.data(), deps.size(), &p));
}

// S from C (src/colvartypes.cpp:231-255, compute_overlap_matrix)
__device__ void overlap_from_correlation(const double C[9], double *S)
{
  double const xx = C[0], xy = C[1], xz = C[2], yx = C[3], yy = C[4], yz = C[5], zx = C[6],
               zy = C[7], zz = C[8];
  S[0] = xx + yy + zz;
  S[1] = S[4] = yz - zy;
  S[2] = S[8] = -xz + zx;
  S[3] = S[12] = xy - yx;
  S[5] = xx - yy - zz;
  S[6] = S[9] = xy + yx;
  S[7] = S[13] = xz + zx;
  S[10] = -xx + yy - zz;
  S[11] = S[14] = yz + zy;
  S[15] = -xx - yy + zz;
}

__global__ void __launch_bounds__(kThreads)
correlation_kernel(const double *__restrict__ ax, const double *__restrict__ ay,
                   const double *__restrict__ az, const double *__restrict__ bx,
                   const double *__restrict__ by, const double *__restrict__ bz, double *S,
                   double *S_eigvec, cvm::rmatrix *h_C, unsigned int *ticket, unsigned int n)
{
  __shared__ double smem[kThreads / 32];
  __shared__ bool last;
  double c[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (unsigned int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    double const x1 = ax[i], y1 = ay[i], z1 = az[i];
    double const x2 = bx[i], y2 = by[i], z2 = bz[i];
    c[0] = fma(x1, x2, c[0]);
    c[1] = fma(x1, y2, c[1]);
    c[2] = fma(x1, z2, c[2]);
    c[3] = fma(y1, x2, c[3]);
    c[4] = fma(y1, y2, c[4]);
    c[5] = fma(y1, z2, c[5]);
    c[6] = fma(z1, x2, c[6]);
    c[7] = fma(z1, y2, c[7]);
    c[8] = fma(z1, z2, c[8]);
  }
#pragma unroll
  for (int k = 0; k < 9; k++) {
    double v = c[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
#pragma unroll
      for (unsigned int w = 0; w < kThreads / 32; w++) s += smem[w];
      atomicAdd(S + k, s);  // S[0..8] doubles as the accumulator, as in the reference
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    double C[9];
#pragma unroll
    for (int k = 0; k < 9; k++) C[k] = atomicAdd(S + k, 0.0);
    overlap_from_correlation(C, S);
#pragma unroll
    for (int k = 0; k < 16; k++) S_eigvec[k] = S[k];
    cvm::rmatrix M;
    M.xx = C[0]; M.xy = C[1]; M.xz = C[2];
    M.yx = C[3]; M.yy = C[4]; M.yz = C[5];
    M.zx = C[6]; M.zy = C[7]; M.zz = C[8];
    *h_C = M;
    *ticket = 0;
    __threadfence();
  }
}

// Cyclic Jacobi for a symmetric 4x4 matrix, one thread.  A -> diagonal, V accumulates the
// rotations (columns = eigenvectors).  Returns false when the sweep limit was hit.
__device__ bool jacobi_sweeps(double (&A)[4][4], double (&V)[4][4])
{
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 4; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
  }
  double norm2 = 0.0;
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 4; j++) norm2 += A[i][j] * A[i][j];
  }
  if (!(norm2 > 0.0)) return true;
  double const tiny = 1.0e-30 * norm2;  // (1e-15 relative)^2
  for (int sweep = 0; sweep < 50; sweep++) {
    double off2 = 0.0;
    for (int p = 0; p < 3; p++) {
      for (int q = p + 1; q < 4; q++) off2 += A[p][q] * A[p][q];
    }
    if (off2 <= tiny) return true;
#pragma unroll
    for (int p = 0; p < 3; p++) {
#pragma unroll
      for (int q = p + 1; q < 4; q++) {
        double const apq = A[p][q];
        if (apq == 0.0) continue;
        // rotation angle that annihilates A[p][q] (Rutishauser's stable form)
        double const theta = 0.5 * (A[q][q] - A[p][p]) / apq;
        double const t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double const c = 1.0 / sqrt(t * t + 1.0);
        double const s = t * c;
        double const tau = s / (1.0 + c);
        A[p][p] -= t * apq;
        A[q][q] += t * apq;
        A[p][q] = A[q][p] = 0.0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
          if (r != p && r != q) {
            double const arp = A[r][p], arq = A[r][q];
            A[r][p] = A[p][r] = arp - s * (arq + tau * arp);
            A[r][q] = A[q][r] = arq + s * (arp - tau * arq);
          }
          double const vrp = V[r][p], vrq = V[r][q];
          V[r][p] = vrp - s * (vrq + tau * vrp);
          V[r][q] = vrq + s * (vrp - tau * vrq);
        }
      }
    }
  }
  return false;
}

__global__ void eigen4_kernel(double *S_eigvec, double *S_eigval, int *max_reached,
                              cvm::quaternion *q, bool monitor_crossings,
                              double crossing_threshold, cvm::quaternion *q_old,
                              int *discontinuous_rotation)
{
  if (threadIdx.x != 0) return;
  double A[4][4], V[4][4];
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 4; j++) A[i][j] = S_eigvec[4 * i + j];
  }
  bool const converged = jacobi_sweeps(A, V);
  if (max_reached) *max_reached = converged ? 0 : 1;
  // largest eigenvalue first; the other three keep their places
  int lead = 0;
  for (int k = 1; k < 4; k++) {
    if (A[k][k] > A[lead][lead]) lead = k;
  }
  int order[4] = {lead, 0, 0, 0};
  for (int k = 0, o = 1; k < 4; k++) {
    if (k != lead) order[o++] = k;
  }
  for (int r = 0; r < 4; r++) {
    int const k = order[r];
    S_eigval[r] = A[k][k];
    for (int j = 0; j < 4; j++) S_eigvec[4 * r + j] = V[j][k];  // eigenvectors as rows
  }
  q->q0 = V[0][lead];
  q->q1 = V[1][lead];
  q->q2 = V[2][lead];
  q->q3 = V[3][lead];
  if (monitor_crossings && q_old->norm2() > 0) {
    q->match(*q_old);
    if (q_old->inner(*q) < (1.0 - crossing_threshold)) atomicAdd(discontinuous_rotation, 1);
  }
}

// tmp_Q0Q0[i][j] = Q0_i Q0_j ; tmp_Q0Q0_L[i][j][k] = sum_{m=1..3} Q_m[j] Q0[k] / (L0 - L_m) Q_m[i]
// (the tables read by rotation_derivative_gpu::calc_derivative_impl,
//  src/colvar_rotation_derivative.h:687-786)
__global__ void derivative_tables_kernel(rotation_derivative_dldq need,
                                         const double *__restrict__ L,
                                         const double *__restrict__ Q, double *tmp_Q0Q0,
                                         double *tmp_Q0Q0_L)
{
  int const idx = threadIdx.x;
  if ((need & rotation_derivative_dldq::use_dl) && idx < 16) {
    tmp_Q0Q0[idx] = Q[idx >> 2] * Q[idx & 3];
  }
  if ((need & rotation_derivative_dldq::use_dq) && idx < 64) {
    int const i = idx >> 4, j = (idx >> 2) & 3, k = idx & 3;
    double const q0k = Q[k];
    double acc = 0.0;
#pragma unroll
    for (int m = 1; m < 4; m++) acc += (Q[4 * m + j] * q0k) / (L[0] - L[m]) * Q[4 * m + i];
    tmp_Q0Q0_L[idx] = acc;
  }
}

}  // namespace

namespace colvars_gpu {

// src/cuda/colvartypes_kernel.h:11-25
int build_overlapping_matrix(const cvm::real *pos1_x, const cvm::real *pos1_y,
                             const cvm::real *pos1_z, const cvm::real *pos2_x,
                             const cvm::real *pos2_y, const cvm::real *pos2_z, cvm::real *S,
                             cvm::real *S_eigvec, cvm::rmatrix *h_C, unsigned int *tbcount,
                             unsigned int num_atoms, cudaGraphNode_t &node, cudaGraph_t &graph,
                             const std::vector<cudaGraphNode_t> &dependencies)
{
  unsigned int const ctas =
      std::max(1u, std::min(kReduceCtas, (num_atoms + kThreads - 1) / kThreads));
  void *args[] = {&pos1_x, &pos1_y, &pos1_z, &pos2_x, &pos2_y, &pos2_z,
                  &S,      &S_eigvec, &h_C,  &tbcount, &num_atoms};
  return add_kernel((void *)correlation_kernel, ctas, kThreads, args, node, graph, dependencies);
}

// src/cuda/colvartypes_kernel.h:27-38
int jacobi_4x4(double *S_eigvec, double *S_eigval, int *max_reached, cvm::quaternion *q,
               bool monitor_crossings, cvm::real crossing_threshold, cvm::quaternion *q_old,
               int *discontinuous_rotation, cudaGraphNode_t &node, cudaGraph_t &graph,
               const std::vector<cudaGraphNode_t> &dependencies)
{
  void *args[] = {&S_eigvec, &S_eigval, &max_reached, &q, &monitor_crossings,
                  &crossing_threshold, &q_old, &discontinuous_rotation};
  return add_kernel((void *)eigen4_kernel, 1, 32, args, node, graph, dependencies);
}

// src/cuda/colvar_rotation_derivative_kernel.h:12-20
int prepare_derivative(rotation_derivative_dldq dldq, const cvm::real *S_eigval,
                       const cvm::real *S_eigvec, cvm::real *tmp_Q0Q0, cvm::real *tmp_Q0Q0_L,
                       cudaGraphNode_t &node, cudaGraph_t &graph,
                       const std::vector<cudaGraphNode_t> &dependencies)
{
  void *args[] = {&dldq, &S_eigval, &S_eigvec, &tmp_Q0Q0, &tmp_Q0Q0_L};
  return add_kernel((void *)derivative_tables_kernel, 1, 64, args, node, graph, dependencies);
}

}  // namespace colvars_gpu

#endif  // COLVARS_CUDA
