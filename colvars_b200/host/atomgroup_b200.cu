// atomgroup_b200.cu -- B200 kernels behind the reference's atom-group kernel boundary.
//
// Under `smp gpu` every cvm::atom_group owns a colvaratoms_gpu whose graphs are assembled from
// the free functions declared in the reference's src/cuda/colvaratoms_kernel.h:13-202: each adds
// ONE kernel node (or launches on a stream) and returns a Colvars error code.  This file defines
// all 17 of them, so linking it in front of the reference archive replaces the archive member
// built from src/cuda/colvaratoms_kernel.cu without touching the host: gather from the proxy
// (SURVEY.md 8a a11), COM / COG (a12), translate / rotate (a12), the two fit-gradient loops
// (a15), force scatter (a16) and the small helpers.
//
// What is different from the member it replaces (measured: profiles/README.md):
//  * reductions use every SM (the reference caps them at 64 CTAs of 128 threads): 296 CTAs x 256
//    threads, two independent accumulation chains per thread, shuffle tree, one atomic per CTA
//    and sum, the last CTA (ticket) publishes and re-arms the scratch;
//  * element-wise kernels are grid-stride with 256-thread CTAs sized to the machine;
//  * scatter kernels issue fire-and-forget reductions (no returned value, RED.ADD.F64).
// Semantics (what is read, written, zeroed again, and in which layout) follow the reference
// wrappers one to one; the rotation-derivative projections are the reference's own device
// methods (src/colvar_rotation_derivative.h:843-935), part of the unmodified host framework.
#include "colvar_gpu_support.h"
#include "colvarmodule.h"
#include "colvartypes.h"
#include "cuda/colvaratoms_kernel.h"

#include <algorithm>
#include <tuple>
#include <vector>

#if defined(COLVARS_CUDA)

extern "C" int b200cv_atomgroup_kernels_linked(void) { return 1; }

namespace {

constexpr unsigned int kThreads = 256;
constexpr unsigned int kReduceCtas = 296;  // 2 CTAs x 148 SMs

inline unsigned int ctas_for(unsigned int n, unsigned int per_thread = 1)
{
  unsigned int const b = (n + kThreads * per_thread - 1) / (kThreads * per_thread);
  return std::max(1u, b);
}

// one kernel node from a typed argument pack
template <typename... P, typename... A>
int add_node(void (*kernel)(P...), unsigned int ctas, cudaGraphNode_t &node, cudaGraph_t &graph,
             const std::vector<cudaGraphNode_t> &deps, A... args)
{
  static_assert(sizeof...(P) == sizeof...(A), "argument count");
  std::tuple<P...> typed(args...);
  void *ptrs[sizeof...(P)];
  std::apply([&](auto &...v) { size_t k = 0; ((ptrs[k++] = (void *)&v), ...); }, typed);
  cudaKernelNodeParams p = {0};
  p.func = (void *)kernel;
  p.gridDim = dim3(ctas, 1, 1);
  p.blockDim = dim3(kThreads, 1, 1);
  p.sharedMemBytes = 0;
  p.kernelParams = ptrs;
  p.extra = NULL;
  return checkGPUError(cudaGraphAddKernelNode(&node, graph, deps.data(), deps.size(), &p));
}

struct quat4 {
  double v[4];
  __device__ double operator[](int k) const { return v[k]; }
};

// sum over the CTA; result valid in thread 0
__device__ __forceinline__ double cta_sum(double v, double *smem)
{
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();  // smem reuse between consecutive calls
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (unsigned int w = 0; w < kThreads / 32; w++) s += smem[w];
  }
  return s;
}

// returns true in every thread of the CTA that arrived last
__device__ __forceinline__ bool last_cta(unsigned int *ticket)
{
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  return last;
}

__device__ __forceinline__ double read_fresh(double *p) { return atomicAdd(p, 0.0); }

// ---- gather --------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) gather_kernel(const int *__restrict__ index,
                                                          const double *__restrict__ proxy,
                                                          unsigned int proxy_stride,
                                                          double *__restrict__ pos,
                                                          unsigned int n)
{
  for (unsigned int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    unsigned int const p = (unsigned int)index[i];
    pos[i] = __ldg(proxy + p);
    pos[n + i] = __ldg(proxy + proxy_stride + p);
    pos[2 * n + i] = __ldg(proxy + 2 * (size_t)proxy_stride + p);
  }
}

__global__ void nudge_kernel(double *pos, size_t element, double delta) { pos[element] += delta; }

// ---- centres -------------------------------------------------------------------------------
// tmp vectors are zero on entry and zero again on exit; so is the ticket
template <bool COM>
__global__ void __launch_bounds__(kThreads)
centers_kernel(const double *__restrict__ pos, const double *__restrict__ mass, unsigned int n,
               cvm::rvector *cog_tmp, cvm::rvector *cog_out, cvm::rvector *cog_saved,
               cvm::rvector *com_tmp, cvm::rvector *com_out, cvm::rvector *h_cog,
               cvm::rvector *h_com, double total_mass, unsigned int *ticket)
{
  __shared__ double smem[kThreads / 32];
  double g[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
  double m[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
  unsigned int const stride = gridDim.x * kThreads;
  unsigned int i = blockIdx.x * kThreads + threadIdx.x;
  for (; i + stride < n; i += 2 * stride) {
#pragma unroll
    for (int u = 0; u < 2; u++) {
      unsigned int const a = i + u * stride;
      double const x = pos[a], y = pos[n + a], z = pos[2 * n + a];
      g[u][0] += x;
      g[u][1] += y;
      g[u][2] += z;
      if (COM) {
        double const w = mass[a];
        m[u][0] = fma(w, x, m[u][0]);
        m[u][1] = fma(w, y, m[u][1]);
        m[u][2] = fma(w, z, m[u][2]);
      }
    }
  }
  if (i < n) {
    double const x = pos[i], y = pos[n + i], z = pos[2 * n + i];
    g[0][0] += x;
    g[0][1] += y;
    g[0][2] += z;
    if (COM) {
      double const w = mass[i];
      m[0][0] = fma(w, x, m[0][0]);
      m[0][1] = fma(w, y, m[0][1]);
      m[0][2] = fma(w, z, m[0][2]);
    }
  }
  double *gt = &cog_tmp->x;
  double *mt = COM ? &com_tmp->x : nullptr;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    double const s = cta_sum(g[0][d] + g[1][d], smem);
    if (threadIdx.x == 0) atomicAdd(gt + d, s);
    if (COM) {
      double const t = cta_sum(m[0][d] + m[1][d], smem);
      if (threadIdx.x == 0) atomicAdd(mt + d, t);
    }
  }
  if (last_cta(ticket) && threadIdx.x == 0) {
    cvm::rvector c;
    c.x = read_fresh(gt) / n;
    c.y = read_fresh(gt + 1) / n;
    c.z = read_fresh(gt + 2) / n;
    *cog_out = c;
    if (h_cog) *h_cog = c;
    if (cog_saved) *cog_saved = c;
    gt[0] = gt[1] = gt[2] = 0.0;
    if (COM) {
      cvm::rvector w;
      w.x = read_fresh(mt) / total_mass;
      w.y = read_fresh(mt + 1) / total_mass;
      w.z = read_fresh(mt + 2) / total_mass;
      *com_out = w;
      if (h_com) *h_com = w;
      mt[0] = mt[1] = mt[2] = 0.0;
    }
    *ticket = 0;
    __threadfence();
  }
}

// ---- per-atom transforms -------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) translate_kernel(double *__restrict__ pos,
                                                             double factor,
                                                             const cvm::rvector *__restrict__ t,
                                                             unsigned int n)
{
  double const tx = factor * t->x, ty = factor * t->y, tz = factor * t->z;
  for (unsigned int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    pos[i] += tx;
    pos[n + i] += ty;
    pos[2 * n + i] += tz;
  }
}

__global__ void __launch_bounds__(kThreads) rotate_kernel(double *__restrict__ pos,
                                                          const cvm::quaternion *__restrict__ q,
                                                          unsigned int n)
{
  cvm::rmatrix const R = q->rotation_matrix();
  for (unsigned int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    double const x = pos[i], y = pos[n + i], z = pos[2 * n + i];
    pos[i] = R.xx * x + R.xy * y + R.xz * z;
    pos[n + i] = R.yx * x + R.yy * y + R.yz * z;
    pos[2 * n + i] = R.zx * x + R.zy * y + R.zz * z;
  }
}

// total forces: proxy -> group, optionally into the rotated frame
template <bool ROTATE>
__global__ void __launch_bounds__(kThreads)
total_force_kernel(const int *__restrict__ index, const double *__restrict__ proxy,
                   unsigned int proxy_stride, double *__restrict__ out,
                   const cvm::quaternion *__restrict__ q, unsigned int n)
{
  cvm::rmatrix R;
  if (ROTATE) R = q->rotation_matrix();
  for (unsigned int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    unsigned int const p = (unsigned int)index[i];
    double const fx = __ldg(proxy + p);
    double const fy = __ldg(proxy + proxy_stride + p);
    double const fz = __ldg(proxy + 2 * (size_t)proxy_stride + p);
    if (ROTATE) {
      out[i] = R.xx * fx + R.xy * fy + R.xz * fz;
      out[n + i] = R.yx * fx + R.yy * fy + R.yz * fz;
      out[2 * n + i] = R.zx * fx + R.zy * fy + R.zz * fz;
    } else {
      out[i] = fx;
      out[n + i] = fy;
      out[2 * n + i] = fz;
    }
  }
}

// proxy[index[i]] += scale * (R^-1) v_i ; scale read from the device when scale_ptr != null
template <bool ROTATE>
__global__ void __launch_bounds__(kThreads)
scatter_kernel(const int *__restrict__ index, double *__restrict__ proxy,
               unsigned int proxy_stride, const double *__restrict__ v,
               const double *__restrict__ scale_ptr, const cvm::quaternion *__restrict__ q,
               unsigned int n)
{
  double const s = scale_ptr ? *scale_ptr : 1.0;
  cvm::rmatrix Ri;
  if (ROTATE) Ri = q->conjugate().rotation_matrix();
  for (unsigned int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    unsigned int const p = (unsigned int)index[i];
    double x = v[i], y = v[n + i], z = v[2 * n + i];
    if (ROTATE) {
      double const rx = Ri.xx * x + Ri.xy * y + Ri.xz * z;
      double const ry = Ri.yx * x + Ri.yy * y + Ri.yz * z;
      double const rz = Ri.zx * x + Ri.zy * y + Ri.zz * z;
      x = rx;
      y = ry;
      z = rz;
    }
    atomicAdd(proxy + p, s * x);
    atomicAdd(proxy + proxy_stride + p, s * y);
    atomicAdd(proxy + 2 * (size_t)proxy_stride + p, s * z);
  }
}

// forces computed by a CPU-side component (mapped host array) added to the group's device array
__global__ void __launch_bounds__(kThreads) add_host_force_kernel(const double *__restrict__ h,
                                                                  double *__restrict__ d,
                                                                  unsigned int n3)
{
  for (unsigned int i = blockIdx.x * kThreads + threadIdx.x; i < n3; i += gridDim.x * kThreads) {
    atomicAdd(d + i, h[i]);  // a group may be shared by several components
  }
}

// ---- fit gradients / fit forces (src/colvaratoms.cpp:1471-1527 on the device) ---------------
// loop 1 over the main group: S = sum g, C = sum g (x) x_unrotated; dq = d(R x)/dq : C;
// the last CTA turns the accumulated dq into dxdC and S into -R^-1 S / Nfit.
template <bool CENTER, bool ROTATE>
__global__ void __launch_bounds__(kThreads)
fit_loop1_kernel(const double *__restrict__ g, const double *__restrict__ xu,
                 const colvars_gpu::rotation_derivative_gpu *__restrict__ rot_deriv,
                 const cvm::quaternion *__restrict__ q, unsigned int n, unsigned int nfit,
                 double3 *atom_grad, double *sum_dxdq, cvm::rmatrix *dxdC, unsigned int *ticket)
{
  __shared__ double smem[kThreads / 32];
  double S[3] = {0.0, 0.0, 0.0};
  double C[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
  for (unsigned int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    double const gx = g[i], gy = g[n + i], gz = g[2 * n + i];
    S[0] += gx;
    S[1] += gy;
    S[2] += gz;
    if (ROTATE) {
      double const x = xu[i], y = xu[n + i], z = xu[2 * n + i];
      C[0][0] = fma(gx, x, C[0][0]);
      C[0][1] = fma(gx, y, C[0][1]);
      C[0][2] = fma(gx, z, C[0][2]);
      C[1][0] = fma(gy, x, C[1][0]);
      C[1][1] = fma(gy, y, C[1][1]);
      C[1][2] = fma(gy, z, C[1][2]);
      C[2][0] = fma(gz, x, C[2][0]);
      C[2][1] = fma(gz, y, C[2][1]);
      C[2][2] = fma(gz, z, C[2][2]);
    }
  }
  double *ag = &atom_grad->x;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    double const s = cta_sum(S[d], smem);
    if (threadIdx.x == 0) atomicAdd(ag + d, s);
  }
  if (ROTATE) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
      for (int b = 0; b < 3; b++) C[a][b] = cta_sum(C[a][b], smem);
    }
    if (threadIdx.x == 0) {
      // linear in C: project this CTA's share onto the quaternion components
      quat4 const dq = q->derivative_element_wise_product_sum<quat4>(C);
#pragma unroll
      for (int k = 0; k < 4; k++) atomicAdd(sum_dxdq + k, dq[k]);
    }
  }
  if (last_cta(ticket) && threadIdx.x == 0) {
    if (ROTATE) {
      quat4 total;
#pragma unroll
      for (int k = 0; k < 4; k++) total.v[k] = read_fresh(sum_dxdq + k);
      *dxdC = rot_deriv->project_force_to_C_from_dxdq(total);
#pragma unroll
      for (int k = 0; k < 4; k++) sum_dxdq[k] = 0.0;
    }
    if (CENTER) {
      double sx = read_fresh(ag), sy = read_fresh(ag + 1), sz = read_fresh(ag + 2);
      if (ROTATE) {
        cvm::rmatrix const Ri = q->conjugate().rotation_matrix();
        double const rx = sx * Ri.xx + sy * Ri.xy + sz * Ri.xz;
        double const ry = sx * Ri.yx + sy * Ri.yy + sz * Ri.yz;
        double const rz = sx * Ri.zx + sy * Ri.zy + sz * Ri.zz;
        sx = rx;
        sy = ry;
        sz = rz;
      }
      double const c = -1.0 / nfit;
      ag[0] = sx * c;
      ag[1] = sy * c;
      ag[2] = sz * c;
    }
    *ticket = 0;
    __threadfence();
  }
}

// loop 2 over the fitting group: centre term + rotation term; written to fit_gradients or
// (fit forces) added to the proxy's force array
template <bool CENTER, bool ROTATE, bool TO_PROXY>
__global__ void __launch_bounds__(kThreads)
fit_loop2_kernel(double *__restrict__ fit_grad,
                 const colvars_gpu::rotation_derivative_gpu *__restrict__ rot_deriv,
                 const double3 *__restrict__ atom_grad, const cvm::rmatrix *__restrict__ dxdC,
                 const int *__restrict__ index, double *__restrict__ proxy,
                 unsigned int proxy_stride, unsigned int nfit)
{
  cvm::rmatrix M;
  if (ROTATE) M = *dxdC;
  double cx = 0.0, cy = 0.0, cz = 0.0;
  if (CENTER) {
    cx = atom_grad->x;
    cy = atom_grad->y;
    cz = atom_grad->z;
  }
  for (unsigned int i = blockIdx.x * kThreads + threadIdx.x; i < nfit; i += gridDim.x * kThreads) {
    double fx = cx, fy = cy, fz = cz;
    if (ROTATE) {
      cvm::rvector const r = rot_deriv->project_force_to_group1(i, M);
      fx += r.x;
      fy += r.y;
      fz += r.z;
    }
    if (TO_PROXY) {
      unsigned int const p = (unsigned int)index[i];
      atomicAdd(proxy + p, fx);
      atomicAdd(proxy + proxy_stride + p, fy);
      atomicAdd(proxy + 2 * (size_t)proxy_stride + p, fz);
    } else {
      fit_grad[i] = fx;
      fit_grad[nfit + i] = fy;
      fit_grad[2 * nfit + i] = fz;
    }
  }
}

}  // namespace

// ============================ the boundary (src/cuda/colvaratoms_kernel.h) ====================
namespace colvars_gpu {

// :13-21
int atoms_pos_from_proxy(const int *atoms_proxy_index, const cvm::real *atoms_pos_proxy,
                         cvm::real *atoms_pos_ag, unsigned int num_atoms,
                         unsigned int proxy_stride, cudaGraphNode_t &node, cudaGraph_t &graph,
                         const std::vector<cudaGraphNode_t> &dependencies)
{
  return add_node(gather_kernel, ctas_for(num_atoms), node, graph, dependencies,
                  atoms_proxy_index, atoms_pos_proxy, proxy_stride, atoms_pos_ag, num_atoms);
}

// :24-30 (debug gradients)
int atoms_pos_from_proxy(const int *atoms_proxy_index, const cvm::real *atoms_pos_proxy,
                         cvm::real *atoms_pos_ag, unsigned int num_atoms,
                         unsigned int proxy_stride, cudaStream_t stream)
{
  gather_kernel<<<ctas_for(num_atoms), kThreads, 0, stream>>>(atoms_proxy_index, atoms_pos_proxy,
                                                             proxy_stride, atoms_pos_ag,
                                                             num_atoms);
  return checkGPUError(cudaGetLastError());
}

// :32-37
int change_one_coordinate(cvm::real *atoms_pos_ag, size_t atom_id_in_group, int xyz,
                          cvm::real step_size, unsigned int num_atoms, cudaStream_t stream)
{
  if (xyz < 0 || xyz > 2) return COLVARS_OK;
  nudge_kernel<<<1, 1, 0, stream>>>(atoms_pos_ag, (size_t)num_atoms * xyz + atom_id_in_group,
                                    step_size);
  return checkGPUError(cudaGetLastError());
}

// :39-54
int atoms_calc_cog_com(const cvm::real *atoms_pos_ag, const cvm::real *atoms_mass,
                       unsigned int num_atoms, cvm::rvector *d_cog_tmp, cvm::rvector *d_cog_out,
                       cvm::rvector *d_cog_origin, cvm::rvector *d_com_tmp,
                       cvm::rvector *d_com_out, cvm::rvector *h_cog_out,
                       cvm::rvector *h_com_out, cvm::real total_mass, unsigned int *tbcount,
                       cudaGraphNode_t &node, cudaGraph_t &graph,
                       const std::vector<cudaGraphNode_t> &dependencies)
{
  unsigned int const ctas = std::min(kReduceCtas, ctas_for(num_atoms, 2));
  return add_node(centers_kernel<true>, ctas, node, graph, dependencies, atoms_pos_ag,
                  atoms_mass, num_atoms, d_cog_tmp, d_cog_out, d_cog_origin, d_com_tmp,
                  d_com_out, h_cog_out, h_com_out, total_mass, tbcount);
}

// :56-66
int atoms_calc_cog(const cvm::real *atoms_pos_ag, unsigned int num_atoms,
                   cvm::rvector *d_cog_tmp, cvm::rvector *d_cog_out, cvm::rvector *d_cog_origin,
                   cvm::rvector *h_cog_out, unsigned int *tbcount, cudaGraphNode_t &node,
                   cudaGraph_t &graph, const std::vector<cudaGraphNode_t> &dependencies)
{
  unsigned int const ctas = std::min(kReduceCtas, ctas_for(num_atoms, 2));
  return add_node(centers_kernel<false>, ctas, node, graph, dependencies, atoms_pos_ag,
                  (const double *)nullptr, num_atoms, d_cog_tmp, d_cog_out, d_cog_origin,
                  (cvm::rvector *)nullptr, (cvm::rvector *)nullptr, h_cog_out,
                  (cvm::rvector *)nullptr, 0.0, tbcount);
}

// :68-76
int atoms_total_force_from_proxy(const int *atoms_proxy_index,
                                 const cvm::real *atoms_total_force_proxy,
                                 cvm::real *atoms_total_force_ag, bool rotate,
                                 const cvm::quaternion *q, unsigned int num_atoms,
                                 unsigned int proxy_stride, cudaStream_t stream)
{
  if (num_atoms == 0) return COLVARS_OK;
  unsigned int const ctas = ctas_for(num_atoms);
  if (rotate) {
    total_force_kernel<true><<<ctas, kThreads, 0, stream>>>(
        atoms_proxy_index, atoms_total_force_proxy, proxy_stride, atoms_total_force_ag, q,
        num_atoms);
  } else {
    total_force_kernel<false><<<ctas, kThreads, 0, stream>>>(
        atoms_proxy_index, atoms_total_force_proxy, proxy_stride, atoms_total_force_ag, q,
        num_atoms);
  }
  return checkGPUError(cudaGetLastError());
}

// :78-89
int apply_main_colvar_force_to_proxy(const int *atoms_proxy_index,
                                     cvm::real *atoms_applied_force_proxy,
                                     const cvm::real *atoms_grad_ag, cvm::real *colvar_force,
                                     bool rotate, const cvm::quaternion *q,
                                     unsigned int num_atoms, unsigned int proxy_stride,
                                     cudaGraphNode_t &node, cudaGraph_t &graph,
                                     const std::vector<cudaGraphNode_t> &dependencies)
{
  return add_node(rotate ? scatter_kernel<true> : scatter_kernel<false>, ctas_for(num_atoms),
                  node, graph, dependencies, atoms_proxy_index, atoms_applied_force_proxy,
                  proxy_stride, atoms_grad_ag, (const double *)colvar_force, q, num_atoms);
}

// :91-100
int apply_fitting_colvar_force_to_proxy(const int *atoms_proxy_index,
                                        cvm::real *atoms_applied_force_proxy,
                                        const cvm::real *atoms_grad_ag, cvm::real *colvar_force,
                                        unsigned int num_atoms, unsigned int proxy_stride,
                                        cudaGraphNode_t &node, cudaGraph_t &graph,
                                        const std::vector<cudaGraphNode_t> &dependencies)
{
  return add_node(scatter_kernel<false>, ctas_for(num_atoms), node, graph, dependencies,
                  atoms_proxy_index, atoms_applied_force_proxy, proxy_stride, atoms_grad_ag,
                  (const double *)colvar_force, (const cvm::quaternion *)nullptr, num_atoms);
}

// :102-108
int accumulate_cpu_force(const cvm::real *h_atoms_force, cvm::real *d_atoms_force,
                         unsigned int num_atoms, cudaGraphNode_t &node, cudaGraph_t &graph,
                         const std::vector<cudaGraphNode_t> &dependencies)
{
  return add_node(add_host_force_kernel, ctas_for(3 * num_atoms), node, graph, dependencies,
                  h_atoms_force, d_atoms_force, 3 * num_atoms);
}

namespace {

int add_fit_loop1(const cvm::real *pos_unrotated, const cvm::real *main_grad,
                  const rotation_derivative_gpu *rot_deriv, const cvm::quaternion *q,
                  unsigned int n, unsigned int nfit, double3 *atom_grad, double *sum_dxdq,
                  cvm::rmatrix *dxdC, unsigned int *tbcount, bool center, bool rotate,
                  cudaGraphNode_t &node, cudaGraph_t &graph,
                  const std::vector<cudaGraphNode_t> &deps)
{
  if (!center && !rotate) return COLVARS_OK;
  auto kernel = center ? (rotate ? fit_loop1_kernel<true, true> : fit_loop1_kernel<true, false>)
                       : fit_loop1_kernel<false, true>;
  unsigned int const ctas = std::min(kReduceCtas, ctas_for(n));
  return add_node(kernel, ctas, node, graph, deps, main_grad, pos_unrotated, rot_deriv, q, n,
                  nfit, atom_grad, sum_dxdq, dxdC, tbcount);
}

template <bool TO_PROXY>
int add_fit_loop2(cvm::real *fit_grad, const rotation_derivative_gpu *rot_deriv,
                  const double3 *atom_grad, const cvm::rmatrix *dxdC, const int *index,
                  cvm::real *proxy, unsigned int proxy_stride, unsigned int nfit, bool center,
                  bool rotate, cudaGraphNode_t &node, cudaGraph_t &graph,
                  const std::vector<cudaGraphNode_t> &deps)
{
  if (!center && !rotate) return COLVARS_OK;
  auto kernel = center ? (rotate ? fit_loop2_kernel<true, true, TO_PROXY>
                                 : fit_loop2_kernel<true, false, TO_PROXY>)
                       : fit_loop2_kernel<false, true, TO_PROXY>;
  return add_node(kernel, ctas_for(nfit), node, graph, deps, fit_grad, rot_deriv, atom_grad,
                  dxdC, index, proxy, proxy_stride, nfit);
}

}  // namespace

// :110-124
int calc_fit_gradients_impl_loop1(const cvm::real *pos_unrotated, cvm::real *main_grad,
                                  const rotation_derivative_gpu *rot_deriv,
                                  const cvm::quaternion *q, unsigned int num_atoms_main,
                                  unsigned int num_atoms_fitting, double3 *atom_grad,
                                  double *sum_dxdq, cvm::rmatrix *dxdC, unsigned int *tbcount,
                                  bool ag_center, bool ag_rotate, cudaGraphNode_t &node,
                                  cudaGraph_t &graph,
                                  const std::vector<cudaGraphNode_t> &dependencies)
{
  return add_fit_loop1(pos_unrotated, main_grad, rot_deriv, q, num_atoms_main,
                       num_atoms_fitting, atom_grad, sum_dxdq, dxdC, tbcount, ag_center,
                       ag_rotate, node, graph, dependencies);
}

// :126-135
int calc_fit_gradients_impl_loop2(cvm::real *fit_grad, const rotation_derivative_gpu *rot_deriv,
                                  const double3 *atom_grad, const cvm::rmatrix *dxdC,
                                  unsigned int group_for_fit_size, bool ag_center,
                                  bool ag_rotate, cudaGraphNode_t &node, cudaGraph_t &graph,
                                  const std::vector<cudaGraphNode_t> &dependencies)
{
  return add_fit_loop2<false>(fit_grad, rot_deriv, atom_grad, dxdC, (const int *)nullptr,
                              (cvm::real *)nullptr, 0u, group_for_fit_size, ag_center,
                              ag_rotate, node, graph, dependencies);
}

// :137-144
int apply_translation(cvm::real *atoms_pos_ag, cvm::real translation_vector_factor,
                      const cvm::rvector *translation_vector, unsigned int num_atoms,
                      cudaGraphNode_t &node, cudaGraph_t &graph,
                      const std::vector<cudaGraphNode_t> &dependencies)
{
  return add_node(translate_kernel, ctas_for(num_atoms), node, graph, dependencies, atoms_pos_ag,
                  translation_vector_factor, translation_vector, num_atoms);
}

// :146-152
int rotate_with_quaternion(cvm::real *atoms_pos_ag, const cvm::quaternion *q,
                           unsigned int num_atoms, cudaGraphNode_t &node, cudaGraph_t &graph,
                           const std::vector<cudaGraphNode_t> &dependencies)
{
  return add_node(rotate_kernel, ctas_for(num_atoms), node, graph, dependencies, atoms_pos_ag, q,
                  num_atoms);
}

// :154-163
int apply_force_with_inverse_rotation(const cvm::real *atoms_force, const cvm::quaternion *q,
                                      const int *atoms_proxy_index, cvm::real *proxy_new_force,
                                      unsigned int num_atoms, unsigned int proxy_stride,
                                      cudaGraphNode_t &node, cudaGraph_t &graph,
                                      const std::vector<cudaGraphNode_t> &dependencies)
{
  return add_node(scatter_kernel<true>, ctas_for(num_atoms), node, graph, dependencies,
                  atoms_proxy_index, proxy_new_force, proxy_stride, atoms_force,
                  (const double *)nullptr, q, num_atoms);
}

// :165-173
int apply_force(const cvm::real *atoms_force, const int *atoms_proxy_index,
                cvm::real *proxy_new_force, unsigned int num_atoms, unsigned int proxy_stride,
                cudaGraphNode_t &node, cudaGraph_t &graph,
                const std::vector<cudaGraphNode_t> &dependencies)
{
  return add_node(scatter_kernel<false>, ctas_for(num_atoms), node, graph, dependencies,
                  atoms_proxy_index, proxy_new_force, proxy_stride, atoms_force,
                  (const double *)nullptr, (const cvm::quaternion *)nullptr, num_atoms);
}

// :175-189
int calc_fit_forces_impl_loop1(const cvm::real *pos_unrotated, cvm::real *main_force,
                               const rotation_derivative_gpu *rot_deriv,
                               const cvm::quaternion *q, unsigned int num_atoms_main,
                               unsigned int num_atoms_fitting, double3 *atom_grad,
                               double *sum_dxdq, cvm::rmatrix *dxdC, unsigned int *tbcount,
                               bool ag_center, bool ag_rotate, cudaGraphNode_t &node,
                               cudaGraph_t &graph,
                               const std::vector<cudaGraphNode_t> &dependencies)
{
  return add_fit_loop1(pos_unrotated, main_force, rot_deriv, q, num_atoms_main,
                       num_atoms_fitting, atom_grad, sum_dxdq, dxdC, tbcount, ag_center,
                       ag_rotate, node, graph, dependencies);
}

// :191-201
int calc_fit_forces_impl_loop2(const rotation_derivative_gpu *rot_deriv,
                               const double3 *atom_grad, const cvm::rmatrix *dxdC,
                               const int *atoms_proxy_index, cvm::real *proxy_new_force,
                               unsigned int group_for_fit_size, unsigned int proxy_stride,
                               bool ag_center, bool ag_rotate, cudaGraphNode_t &node,
                               cudaGraph_t &graph,
                               const std::vector<cudaGraphNode_t> &dependencies)
{
  return add_fit_loop2<true>((cvm::real *)nullptr, rot_deriv, atom_grad, dxdC,
                             atoms_proxy_index, proxy_new_force, proxy_stride,
                             group_for_fit_size, ag_center, ag_rotate, node, graph,
                             dependencies);
}

}  // namespace colvars_gpu

#endif  // COLVARS_CUDA
