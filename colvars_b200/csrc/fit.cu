// fit.cu -- fitting kernels (see fit.cuh).
//
// d_small layout (doubles):
//   [0..3]   q          rotation quaternion (leading eigenvector of the overlap matrix)
//   [4..6]   cog        centre of geometry of the fitted-on group (lab frame)
//   [7..9]   ref_cog    centre of geometry of the reference positions
//   [10..18] C          correlation matrix  sum_j x_j (x) ref_j       (src/colvartypes.cpp:211-229)
//   [19..22] eigval     eigenvalues of the overlap matrix, descending
//   [23..38] eigvec     eigenvectors, rows
//   [39..41] sum_g      sum of the main group's gradients
//   [42..50] Cg         sum_i g_i (x) x_i^unrotated                   (src/colvaratoms.cpp:1483-1497)
//   [51..59] dxdC       gradient w.r.t. the correlation matrix
//   [60..62] atom_grad  centring term  -R^T sum_g / Nfit
#include "fit.cuh"

#include <cmath>

#include "../../include/b200cv.h"
#include "kernels.cuh"

namespace b200cv {

namespace {

constexpr int SM_Q = 0, SM_COG = 4, SM_REFCOG = 7, SM_C = 10, SM_VAL = 19, SM_VEC = 23,
              SM_SUMG = 39, SM_CG = 42, SM_DXDC = 51, SM_AGRAD = 60, SM_COUNT = 64;

template <int NV> __device__ __forceinline__ void block_reduce_store(double (&v)[NV], double *out)
{
  __shared__ double red[NV][32];
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < NV; c++) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], d);
    if (lane == 0) red[c][warp] = v[c];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < NV; c++) {
      double x = lane < (int)(blockDim.x >> 5) ? red[c][lane] : 0.0;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
      if (lane == 0) out[c] = x;
    }
  }
}

// centre of geometry (calc_center_of_geometry, src/colvaratoms.cpp:1354-1372)
__global__ void __launch_bounds__(1024) cog_kernel(const double *pos, size_t stride, int n,
                                                   double *cog)
{
  double s[3] = {0, 0, 0};
  for (int i = threadIdx.x; i < n; i += 1024) {
    s[0] += pos[i];
    s[1] += pos[stride + i];
    s[2] += pos[2 * stride + i];
  }
  __shared__ double tmp[3];
  block_reduce_store<3>(s, tmp);
  __syncthreads();
  if (threadIdx.x < 3) cog[threadIdx.x] = tmp[threadIdx.x] / (double)n;
}

// correlation matrix C = sum_j pos_j (x) ref_j
__global__ void __launch_bounds__(1024) corr_kernel(const double *pos, size_t stride,
                                                    const double *ref, size_t rstride, int n,
                                                    double *C)
{
  double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += 1024) {
    double const p[3] = {pos[i], pos[stride + i], pos[2 * stride + i]};
    double const r[3] = {ref[i], ref[rstride + i], ref[2 * rstride + i]};
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
      for (int b = 0; b < 3; b++) s[a * 3 + b] = fma(p[a], r[b], s[a * 3 + b]);
    }
  }
  block_reduce_store<9>(s, C);
}

// overlap matrix (compute_overlap_matrix, src/colvartypes.cpp:231-255) + symmetric 4x4
// eigen-decomposition by cyclic Jacobi in one thread; leading eigenvector -> q
__global__ void eigen_kernel(double *sm)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double *C = sm + SM_C;
  double A[4][4], V[4][4];
  A[0][0] = C[0] + C[4] + C[8];
  A[1][0] = A[0][1] = C[5] - C[7];
  A[2][0] = A[0][2] = -C[2] + C[6];
  A[3][0] = A[0][3] = C[1] - C[3];
  A[1][1] = C[0] - C[4] - C[8];
  A[2][1] = A[1][2] = C[1] + C[3];
  A[3][1] = A[1][3] = C[2] + C[6];
  A[2][2] = -C[0] + C[4] - C[8];
  A[3][2] = A[2][3] = C[5] + C[7];
  A[3][3] = -C[0] - C[4] + C[8];
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 4; j++) V[i][j] = i == j ? 1.0 : 0.0;
  }
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = 0.0, diag = 0.0;
    for (int p = 0; p < 4; p++) {
      diag += A[p][p] * A[p][p];
      for (int q = p + 1; q < 4; q++) off += A[p][q] * A[p][q];
    }
    if (off <= 1e-34 * diag || off == 0.0) break;
    for (int p = 0; p < 3; p++) {
      for (int q = p + 1; q < 4; q++) {
        double const apq = A[p][q];
        if (apq == 0.0) continue;
        double const theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        double const c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; k++) {
          double const akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 4; k++) {
          double const apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 4; k++) {
          double const vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
    }
  }
  int order[4] = {0, 1, 2, 3};
  for (int i = 0; i < 4; i++) {
    for (int j = i + 1; j < 4; j++) {
      if (A[order[j]][order[j]] > A[order[i]][order[i]]) {
        int const t = order[i];
        order[i] = order[j];
        order[j] = t;
      }
    }
  }
  for (int e = 0; e < 4; e++) {
    int const col = order[e];
    sm[SM_VAL + e] = A[col][col];
    double n2 = 0.0;
    for (int k = 0; k < 4; k++) n2 += V[k][col] * V[k][col];
    double const inv = 1.0 / sqrt(n2);
    for (int k = 0; k < 4; k++) sm[SM_VEC + 4 * e + k] = V[k][col] * inv;
  }
  for (int k = 0; k < 4; k++) sm[SM_Q + k] = sm[SM_VEC + k];
}

__device__ __forceinline__ void quat_matrix(const double *q, double R[3][3])
{
  double const q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
  R[0][0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3;
  R[1][1] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3;
  R[2][2] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
  R[0][1] = 2.0 * (q1 * q2 - q0 * q3);
  R[0][2] = 2.0 * (q0 * q2 + q1 * q3);
  R[1][0] = 2.0 * (q0 * q3 + q1 * q2);
  R[1][2] = 2.0 * (q2 * q3 - q0 * q1);
  R[2][0] = 2.0 * (q1 * q3 - q0 * q2);
  R[2][1] = 2.0 * (q0 * q1 + q2 * q3);
}

// loop 1 of calc_fit_forces_impl (src/colvaratoms.cpp:1483-1512) + the projection of the
// summed gradient onto q, then onto the correlation matrix
__global__ void __launch_bounds__(1024) fitgrad_reduce_kernel(const double *grad, size_t gstride,
                                                              const double *unrot, size_t ustride,
                                                              int n, double *sm, int center,
                                                              int rotate, int nfit)
{
  double s[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += 1024) {
    double const g[3] = {grad[i], grad[gstride + i], grad[2 * gstride + i]};
    s[0] += g[0];
    s[1] += g[1];
    s[2] += g[2];
    if (rotate) {
      double const x[3] = {unrot[i], unrot[ustride + i], unrot[2 * ustride + i]};
#pragma unroll
      for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int b = 0; b < 3; b++) s[3 + a * 3 + b] = fma(g[a], x[b], s[3 + a * 3 + b]);
      }
    }
  }
  block_reduce_store<12>(s, sm + SM_SUMG);  // SUMG[3] and CG[9] are contiguous
  __syncthreads();
  if (threadIdx.x != 0) return;
  const double *sumg = sm + SM_SUMG;
  const double *Cg = sm + SM_CG;
  double R[3][3];
  if (rotate) quat_matrix(sm + SM_Q, R);
  for (int a = 0; a < 3; a++) {
    double v = 0.0;
    if (center) {
      v = rotate ? (R[0][a] * sumg[0] + R[1][a] * sumg[1] + R[2][a] * sumg[2]) : sumg[a];
      v *= (-1.0) / (double)nfit;
    }
    sm[SM_AGRAD + a] = v;
  }
  double dxdC[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (rotate) {
    double const q0 = sm[SM_Q], q1 = sm[SM_Q + 1], q2 = sm[SM_Q + 2], q3 = sm[SM_Q + 3];
    // dR/dq_p (src/colvartypes.h:1249-1264)
    double const dR[4][9] = {{q0, -q3, q2, q3, q0, -q1, -q2, q1, q0},
                             {q1, q2, q3, q2, -q1, -q0, q3, q0, -q1},
                             {-q2, q1, q0, q1, q2, q3, -q0, q3, -q2},
                             {-q3, -q0, q1, q0, -q3, q2, q1, q2, q3}};
    const double *L = sm + SM_VAL;
    const double *Q = sm + SM_VEC;
    for (int p = 0; p < 4; p++) {
      double f = 0.0;
      for (int e = 0; e < 9; e++) f += dR[p][e] * Cg[e];
      f *= 2.0;
      // T[i][j] = dq_p/dS_ij = sum_k Qk[p] Qk[i] Q0[j] / (L0 - Lk)
      // (src/colvar_rotation_derivative.h:125-329)
      double T[4][4];
      for (int i = 0; i < 4; i++) {
        for (int j = 0; j < 4; j++) {
          double t = 0.0;
          for (int k = 1; k < 4; k++) {
            t += (Q[4 * k + i] * Q[j]) / (L[0] - L[k]) * Q[4 * k + p];
          }
          T[i][j] = t;
        }
      }
      // dS/dC chain (src/colvar_rotation_derivative.h:535-548)
      dxdC[0] += f * (T[0][0] + T[1][1] - T[2][2] - T[3][3]);
      dxdC[1] += f * (T[0][3] + T[1][2] + T[2][1] + T[3][0]);
      dxdC[2] += f * (-T[0][2] + T[1][3] - T[2][0] + T[3][1]);
      dxdC[3] += f * (-T[0][3] + T[1][2] + T[2][1] - T[3][0]);
      dxdC[4] += f * (T[0][0] - T[1][1] + T[2][2] - T[3][3]);
      dxdC[5] += f * (T[0][1] + T[1][0] + T[2][3] + T[3][2]);
      dxdC[6] += f * (T[0][2] + T[1][3] + T[2][0] + T[3][1]);
      dxdC[7] += f * (-T[0][1] - T[1][0] + T[2][3] + T[3][2]);
      dxdC[8] += f * (T[0][0] - T[1][1] - T[2][2] + T[3][3]);
    }
  }
  for (int e = 0; e < 9; e++) sm[SM_DXDC + e] = dxdC[e];
}

// loop 2 (src/colvaratoms.cpp:1513-1526): fit_grad_j = atom_grad + dxdC ref_j  (assignment)
__global__ void fitgrad_apply_kernel(const double *ref, size_t rstride, int nfit,
                                     const double *sm, int center, int rotate, double *fg,
                                     size_t fstride)
{
  int const j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nfit) return;
  double g[3] = {0, 0, 0};
  if (center) {
    g[0] = sm[SM_AGRAD];
    g[1] = sm[SM_AGRAD + 1];
    g[2] = sm[SM_AGRAD + 2];
  }
  if (rotate) {
    double const rx = ref[j], ry = ref[rstride + j], rz = ref[2 * rstride + j];
    const double *D = sm + SM_DXDC;
    g[0] += D[0] * rx + D[1] * ry + D[2] * rz;
    g[1] += D[3] * rx + D[4] * ry + D[5] * rz;
    g[2] += D[6] * rx + D[7] * ry + D[8] * rz;
  }
  fg[j] = g[0];
  fg[fstride + j] = g[1];
  fg[2 * fstride + j] = g[2];
}

#define FIT_OK(call)                                                          \
  do {                                                                        \
    cudaError_t e_ = (cudaError_t)(call);                                     \
    if (e_ != cudaSuccess) {                                                  \
      err = std::string(#call) + ": " + cudaGetErrorString(e_);               \
      return e_ == cudaErrorMemoryAllocation ? B200CV_MEMORY_ERROR : B200CV_ERROR; \
    }                                                                         \
  } while (0)

}  // namespace

void fit_free(FitState &F)
{
  auto fr = [](auto *&p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  fr(F.d_fit_index);
  fr(F.d_fit_pos);
  fr(F.d_ref);
  fr(F.d_main_unrot);
  fr(F.d_fit_grad);
  fr(F.d_small);
  F.d_q = nullptr;
  F.d_index = nullptr;
  F.active = false;
}

int fit_setup(FitState &F, int center, int rotate, const int *fit_proxy_index, int nfit,
              int nmain, size_t mstride, const int *d_main_index,
              const std::vector<int> &h_main_index, const double *ref_pos_soa,
              int enable_fit_gradients, std::string &err)
{
  fit_free(F);
  F.center = center != 0;
  F.rotate = rotate != 0;
  if (!F.center && !F.rotate) return B200CV_OK;
  if (!ref_pos_soa) {
    err = "Error: no reference positions provided.";
    return B200CV_INPUT_ERROR;
  }
  F.separate = fit_proxy_index != nullptr && nfit > 0;
  F.nfit = F.separate ? nfit : nmain;
  F.nmain = nmain;
  F.mstride = mstride;
  F.fstride = F.separate ? ((size_t)(F.nfit + 31) / 32 * 32 + 32) : mstride;
  F.fit_gradients = enable_fit_gradients != 0;
  FIT_OK(cudaMalloc(&F.d_small, sizeof(double) * SM_COUNT));
  FIT_OK(cudaMemset(F.d_small, 0, sizeof(double) * SM_COUNT));
  F.d_q = F.d_small;
  // center_ref_pos (src/colvaratoms.cpp:1186-1211): sequential sum, divide, subtract
  int const nref = F.nfit;
  double cog[3] = {0, 0, 0};
  for (int i = 0; i < nref; i++) {
    cog[0] += ref_pos_soa[i];
    cog[1] += ref_pos_soa[(size_t)nref + i];
    cog[2] += ref_pos_soa[2 * (size_t)nref + i];
  }
  for (double &c : cog) c /= (double)nref;
  std::vector<double> ref(3 * F.fstride, 0.0);
  for (int d = 0; d < 3; d++) {
    for (int i = 0; i < nref; i++) ref[d * F.fstride + i] = ref_pos_soa[(size_t)d * nref + i] - cog[d];
    F.ref_pos_cog[d] = cog[d];
  }
  FIT_OK(cudaMalloc(&F.d_ref, sizeof(double) * 3 * F.fstride));
  FIT_OK(cudaMemcpy(F.d_ref, ref.data(), sizeof(double) * 3 * F.fstride, cudaMemcpyHostToDevice));
  FIT_OK(cudaMemcpy(F.d_small + SM_REFCOG, cog, sizeof(cog), cudaMemcpyHostToDevice));
  double const qid[4] = {1.0, 0.0, 0.0, 0.0};
  FIT_OK(cudaMemcpy(F.d_small + SM_Q, qid, sizeof(qid), cudaMemcpyHostToDevice));
  if (F.separate) {
    F.h_index.assign(fit_proxy_index, fit_proxy_index + nfit);
    FIT_OK(cudaMalloc(&F.d_fit_index, sizeof(int) * nfit));
    FIT_OK(cudaMemcpy(F.d_fit_index, fit_proxy_index, sizeof(int) * nfit, cudaMemcpyHostToDevice));
    F.d_index = F.d_fit_index;
    FIT_OK(cudaMalloc(&F.d_fit_pos, sizeof(double) * 3 * F.fstride));
    FIT_OK(cudaMemset(F.d_fit_pos, 0, sizeof(double) * 3 * F.fstride));
  } else {
    F.h_index = h_main_index;
    F.d_index = d_main_index;
  }
  if (F.fit_gradients) {
    FIT_OK(cudaMalloc(&F.d_main_unrot, sizeof(double) * 3 * mstride));
    FIT_OK(cudaMalloc(&F.d_fit_grad, sizeof(double) * 3 * F.fstride));
    FIT_OK(cudaMemset(F.d_fit_grad, 0, sizeof(double) * 3 * F.fstride));
  }
  F.active = true;
  return B200CV_OK;
}

int fit_read_and_apply(FitState &F, const double *proxy_pos, size_t pstride, int aos,
                       double *d_pos, size_t stride, int n, cudaStream_t s, std::string &err)
{
  if (!F.active) return B200CV_OK;
  double *fitp = F.separate ? F.d_fit_pos : d_pos;
  size_t const fs = F.separate ? F.fstride : stride;
  if (F.separate) {
    FIT_OK(launch_gather(proxy_pos, pstride, aos, F.d_fit_index, F.nfit, F.d_fit_pos, F.fstride, s));
  }
  if (F.center) {
    cog_kernel<<<1, 1024, 0, s>>>(fitp, fs, F.nfit, F.d_small + SM_COG);
  note_launch();
    FIT_OK(cudaGetLastError());
    FIT_OK(launch_translate(d_pos, stride, n, F.d_small + SM_COG, -1.0, s));
    if (F.separate) FIT_OK(launch_translate(fitp, fs, F.nfit, F.d_small + SM_COG, -1.0, s));
  }
  if (F.fit_gradients) {
    FIT_OK(cudaMemcpyAsync(F.d_main_unrot, d_pos, sizeof(double) * 3 * stride,
                           cudaMemcpyDeviceToDevice, s));
  }
  if (F.rotate) {
    corr_kernel<<<1, 1024, 0, s>>>(fitp, fs, F.d_ref, F.fstride, F.nfit, F.d_small + SM_C);
  note_launch();
    FIT_OK(cudaGetLastError());
    eigen_kernel<<<1, 32, 0, s>>>(F.d_small);
  note_launch();
    FIT_OK(cudaGetLastError());
    FIT_OK(launch_rotate(d_pos, stride, n, F.d_q, s));
    if (F.separate) FIT_OK(launch_rotate(fitp, fs, F.nfit, F.d_q, s));
  }
  if (F.center) {
    FIT_OK(launch_translate(d_pos, stride, n, F.d_small + SM_REFCOG, 1.0, s));
    if (F.separate) FIT_OK(launch_translate(fitp, fs, F.nfit, F.d_small + SM_REFCOG, 1.0, s));
  }
  return B200CV_OK;
}

int fit_calc_gradients(FitState &F, const double *d_grad, size_t stride, int n, cudaStream_t s,
                       std::string &err)
{
  if (!F.active || !F.fit_gradients) return B200CV_OK;
  fitgrad_reduce_kernel<<<1, 1024, 0, s>>>(d_grad, stride, F.d_main_unrot, F.mstride, n, F.d_small,
                                           F.center ? 1 : 0, F.rotate ? 1 : 0, F.nfit);
  note_launch();
  FIT_OK(cudaGetLastError());
  fitgrad_apply_kernel<<<(F.nfit + 255) / 256, 256, 0, s>>>(F.d_ref, F.fstride, F.nfit, F.d_small,
                                                            F.center ? 1 : 0, F.rotate ? 1 : 0,
                                                            F.d_fit_grad, F.fstride);
  note_launch();
  FIT_OK(cudaGetLastError());
  return B200CV_OK;
}

int fit_scatter_force(FitState &F, double force, double *proxy_force, size_t pstride, int aos,
                      cudaStream_t s, std::string &err)
{
  if (!F.active || !F.fit_gradients) return B200CV_OK;
  // fit gradients are already in the laboratory frame (src/colvaratoms.cpp:1691-1703)
  FIT_OK(launch_scatter_force(F.d_fit_grad, F.fstride, F.d_index, F.nfit, force, nullptr,
                              proxy_force, pstride, aos, s));
  return B200CV_OK;
}

int fit_scatter_force_host(FitState &F, double force, double *h_proxy_force, size_t n_proxy,
                           int rank, int world, cudaStream_t s, std::string &err)
{
  if (!F.active || !F.fit_gradients) return B200CV_OK;
  // this rank's share of the fitting atoms (all of them on a single rank)
  size_t const j0 = (size_t)F.nfit * (size_t)rank / (size_t)world;
  size_t const j1 = (size_t)F.nfit * ((size_t)rank + 1) / (size_t)world;
  size_t const m = j1 - j0;
  if (m == 0) return B200CV_OK;
  F.h_stage.resize(3 * m);
  FIT_OK(cudaMemcpy2DAsync(F.h_stage.data(), sizeof(double) * m, F.d_fit_grad + j0,
                           sizeof(double) * F.fstride, sizeof(double) * m, 3,
                           cudaMemcpyDeviceToHost, s));
  FIT_OK(cudaStreamSynchronize(s));
  for (size_t j = 0; j < m; j++) {
    size_t const p = (size_t)F.h_index[j0 + j];
    if (p >= n_proxy) {
      err = "proxy index of a fitting atom out of range";
      return B200CV_BUG_ERROR;
    }
    h_proxy_force[3 * p] += force * F.h_stage[j];
    h_proxy_force[3 * p + 1] += force * F.h_stage[m + j];
    h_proxy_force[3 * p + 2] += force * F.h_stage[2 * m + j];
  }
  return B200CV_OK;
}

int fit_host_inverse_matrix(FitState &F, double Rinv[3][3], std::string &err)
{
  double q[4] = {1, 0, 0, 0};
  if (F.active && F.rotate) FIT_OK(cudaMemcpy(q, F.d_q, sizeof(q), cudaMemcpyDeviceToHost));
  double const q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
  double R[3][3];
  R[0][0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3;
  R[1][1] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3;
  R[2][2] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
  R[0][1] = 2.0 * (q1 * q2 - q0 * q3);
  R[0][2] = 2.0 * (q0 * q2 + q1 * q3);
  R[1][0] = 2.0 * (q0 * q3 + q1 * q2);
  R[1][2] = 2.0 * (q2 * q3 - q0 * q1);
  R[2][0] = 2.0 * (q1 * q3 - q0 * q2);
  R[2][1] = 2.0 * (q0 * q1 + q2 * q3);
  for (int a = 0; a < 3; a++) {
    for (int b = 0; b < 3; b++) Rinv[a][b] = R[b][a];
  }
  return B200CV_OK;
}

int fit_get_gradients(FitState &F, double *out_soa, int nfit, std::string &err)
{
  if (!F.active || !F.fit_gradients) {
    err = "group has no fit gradients";
    return B200CV_INPUT_ERROR;
  }
  if (nfit != F.nfit) {
    err = "fit-gradient count mismatch";
    return B200CV_INPUT_ERROR;
  }
  FIT_OK(cudaMemcpy2D(out_soa, sizeof(double) * nfit, F.d_fit_grad, sizeof(double) * F.fstride,
                      sizeof(double) * nfit, 3, cudaMemcpyDeviceToHost));
  return B200CV_OK;
}

}  // namespace b200cv
