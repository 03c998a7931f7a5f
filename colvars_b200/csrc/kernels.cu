// kernels.cu -- kernels around the tile sweep: exact (reference-order) pair queue,
// pair-list walk, dense fallback, centre-of-mass variants, finalize, atom-group
// gather / COM / rotate / scatter, and the measurement helpers.
#include "kernels.cuh"

#include <algorithm>
#include <atomic>

namespace b200cv {

static std::atomic<unsigned long long> g_launches{0ull};
void note_launch() { g_launches.fetch_add(1ull, std::memory_order_relaxed); }
unsigned long long launches_so_far() { return g_launches.load(std::memory_order_relaxed); }

// ============================ block reduction helper =======================================

template <int THREADS> __device__ __forceinline__ double block_sum(double v, double *red)
{
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
    for (int w = 0; w < THREADS / 32; w++) s += red[w];
  }
  __syncthreads();
  return s;  // valid on thread 0
}

// ---- the final sum (finalize_kernel, or the last CTA of exact_kernel) -------------------------
// fixed summation order: thread t takes partials t, t+T, ...; then a binary tree.  volatile
// loads: when the last CTA of a kernel runs this, the partials were written by other CTAs of
// the same launch (made visible by their fence + ticket).
template <int T>
__device__ __forceinline__ void finalize_body(FinalizeArgs const &A, double *red)
{
  double s = 0.0;
  for (int k = threadIdx.x; k < A.n; k += T) s += ((const volatile double *)A.partials)[k];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int d = T / 2; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) red[threadIdx.x] += red[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *A.d_value = red[0];
    if (A.h_counters) {
      A.h_counters[0] =
          A.rare_count ? (unsigned long long)*(const volatile unsigned int *)A.rare_count : 0ull;
      A.h_counters[1] = A.pl_count ? *(const volatile unsigned long long *)A.pl_count : 0ull;
    }
    bool rebuilt = A.rebuilt != 0;
    if (A.gate) {
      int const mode = *A.gate;
      if (A.h_counters) A.h_counters[2] = (unsigned long long)mode;
      rebuilt = mode == 1;
    }
    if (A.row_counts && A.h_counters) {
      A.h_counters[3] = (unsigned long long)*(const volatile unsigned *)A.row_counts;
      A.h_counters[4] = *(const volatile unsigned long long *)(A.row_counts + 4);
      A.h_counters[5] = (unsigned long long)*(const volatile unsigned *)(A.row_counts + 3);
    }
    if (A.overflow) {
      bool over = A.rare_count && *(const volatile unsigned int *)A.rare_count > A.rare_cap;
      if (rebuilt) {
        over = over || (A.pl_count && *(const volatile unsigned long long *)A.pl_count > A.pl_cap);
        over = over || (A.row_counts && (*(const volatile unsigned *)A.row_counts > A.row_cap_chunks ||
                                         *(const volatile unsigned *)(A.row_counts + 3) > A.row_cap_idx));
      }
      *A.overflow = over ? 1.0 : 0.0;
    }
    if (rebuilt) {
      if (A.pl_persist) {
        unsigned long long const n = *(const volatile unsigned long long *)A.pl_count;
        *A.pl_persist = n < A.pl_cap ? n : A.pl_cap;
      }
      if (A.row_counts) {
        unsigned const u = *(const volatile unsigned *)A.row_counts;
        A.row_counts[6] = u < A.row_cap_chunks ? u : A.row_cap_chunks;
      }
    }
    // mapped pinned copies: the host reads them after synchronising with the stream, and
    // kernel completion makes every write visible -- no fence needed
    if (A.h_value) *A.h_value = red[0];
  }
}

// ============================ exact pair queue / pair-list walk ================================

// GENERAL: the walk's fast form in mixed / triclinic cells as well (image search of the sweep,
// general_image; ambiguous decisions fall through to the reference-order pair).  A variant of
// its own because the search costs registers the other walks should not pay for.
template <bool GENERAL>
__global__ void __launch_bounds__(EXACT_THREADS) exact_kernel(const __grid_constant__ ExactArgs A)
{
  __shared__ double red[EXACT_THREADS / 32];
  __shared__ double tritab[GENERAL ? 42 : 1];
  if (A.gate != nullptr && *A.gate != A.gate_on) {
    if (threadIdx.x == 0) A.partials[blockIdx.x] = 0.0;
    return;
  }
  if constexpr (GENERAL) {
    if (threadIdx.x < 42) {
      tritab[threadIdx.x] = threadIdx.x < 39 ? A.P.tri_shift[threadIdx.x / 3][threadIdx.x % 3] : 0.0;
    }
    __syncthreads();
  }
  unsigned long long n1 = A.count32 ? (unsigned long long)*A.count32 : *A.count64;
  if (n1 > A.cap) n1 = A.cap;
  unsigned long long n = n1;
  if (A.list2 != nullptr) {
    unsigned long long n2 = *A.count2;
    if (n2 > A.cap2) n2 = A.cap2;
    n += n2;
  }
  bool const grad = A.flags & EF_GRADIENTS;
  bool const rebuild = (A.flags & EF_REBUILD_PAIRLIST) && A.pl;
  if (A.gs != nullptr && grad) {
    // the partner shares of the neighbour rows (rows.cuh): cell order -> the group's gradient
    double *const tx = A.gs_group == 1 ? A.g1x : A.g2x;
    double *const ty = A.gs_group == 1 ? A.g1y : A.g2y;
    double *const tz = A.gs_group == 1 ? A.g1z : A.g2z;
    for (int p = blockIdx.x * EXACT_THREADS + threadIdx.x; p < A.gs_n; p += gridDim.x * EXACT_THREADS) {
      double const sx = A.gs[p], sy = A.gs[A.gs_stride + p], sz = A.gs[2 * A.gs_stride + p];
      if (sx != 0.0 || sy != 0.0 || sz != 0.0) {
        int const atom = A.gs_sorted[p];
        atomicAdd(&tx[atom], A.P.c_grad[0] * sx);
        atomicAdd(&ty[atom], A.P.c_grad[1] * sy);
        atomicAdd(&tz[atom], A.P.c_grad[2] * sz);
      }
    }
  }
  double fsum = 0.0;
  // warp-uniform trip count: the group1 gradients of a warp's 32 entries are summed per run of
  // equal rows before they go to memory (a list rebuilt through the cell list is made of such
  // runs; 32 lanes adding to one address would serialise, and 3 of the 6 atomics per pair go away)
  int const lane = threadIdx.x & 31;
  unsigned long long const first = ((unsigned long long)blockIdx.x * EXACT_THREADS + threadIdx.x) -
                                   (unsigned long long)lane;
  for (unsigned long long e0 = first; e0 < n; e0 += (unsigned long long)gridDim.x * EXACT_THREADS) {
    unsigned long long const e = e0 + (unsigned long long)lane;
    bool const valid = e < n;
    int i = -1 - lane, j = 0;
    double Gx = 0.0, Gy = 0.0, Gz = 0.0, F = 0.0;
    unsigned long long ent = 0ull;
    if (valid) {
      ent = e < n1 ? A.list[e] : A.list2[e - n1];
      i = (int)(ent >> 32);
      j = (int)(ent & 0xffffffffu);
      double const x1 = A.x1[i], y1 = A.y1[i], z1 = A.z1[i];
      double const x2 = A.x2[j], y2 = A.y2[j], z2 = A.z2[j];
      bool done = false;
      if (A.fast_exp == 3) {
        // 6/12, non-periodic or orthorhombic: the sweep's fast form for every pair that is not
        // next to a decision threshold (same windows as the sweep), reference order otherwise
        double dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
        bool ambiguous = false;
        if constexpr (GENERAL) {
          ambiguous = general_image(A.P, tritab, dx, dy, dz);
        } else if (A.P.bc_type == 2) {
          double const shx = floor(__dadd_rn(__dmul_rn(A.P.recip[0][0], dx), 0.5));
          double const shy = floor(__dadd_rn(__dmul_rn(A.P.recip[1][1], dy), 0.5));
          double const shz = floor(__dadd_rn(__dmul_rn(A.P.recip[2][2], dz), 0.5));
          dx = fma(-shx, A.P.cell[0][0], dx);
          dy = fma(-shy, A.P.cell[1][1], dy);
          dz = fma(-shz, A.P.cell[2][2], dz);
        }
        double const sx = dx * A.P.inv_r0[0], sy = dy * A.P.inv_r0[1], sz = dz * A.P.inv_r0[2];
        double const u = fma(sz, sz, fma(sy, sy, sx * sx));
        if (!ambiguous && !pair_is_rare<true>(u, A.P)) {
          double gq;
          switching_fast<3, true>(u, false, A.P, F, gq);
          Gx = A.P.c_grad[0] * gq * dx;
          Gy = A.P.c_grad[1] * gq * dy;
          Gz = A.P.c_grad[2] * gq * dz;
          done = true;
        }
      }
      if (!done) F = pair_exact(A.P, x1, y1, z1, x2, y2, z2, A.flags, Gx, Gy, Gz);
    }
    fsum += F;
    bool const counts = valid && F > 0.0;
    if (rebuild && counts) {
      unsigned long long const slot = atomicAdd(A.pl_count, 1ull);
      if (slot < A.pl_cap) A.pl[slot] = ent;
    }
    if (grad) {
      if (counts) {
        atomicAdd(&A.g2x[j], Gx);
        atomicAdd(&A.g2y[j], Gy);
        atomicAdd(&A.g2z[j], Gz);
      } else {
        Gx = Gy = Gz = 0.0;
      }
      // runs of equal i -> segment numbers -> segmented inclusive scan -> the last lane of a
      // run holds its sum
      int const prev = __shfl_up_sync(0xffffffffu, i, 1);
      unsigned const heads = __ballot_sync(0xffffffffu, lane == 0 || prev != i);
      int const seg = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        double const px = __shfl_up_sync(0xffffffffu, Gx, d);
        double const py = __shfl_up_sync(0xffffffffu, Gy, d);
        double const pz = __shfl_up_sync(0xffffffffu, Gz, d);
        int const ps = __shfl_up_sync(0xffffffffu, seg, d);
        if (lane >= d && ps == seg) {
          Gx += px;
          Gy += py;
          Gz += pz;
        }
      }
      bool const tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
      if (tail && valid && (Gx != 0.0 || Gy != 0.0 || Gz != 0.0)) {
        atomicAdd(&A.g1x[i], -Gx);
        atomicAdd(&A.g1y[i], -Gy);
        atomicAdd(&A.g1z[i], -Gz);
      }
    }
  }
  double const s = block_sum<EXACT_THREADS>(fsum, red);
  if (threadIdx.x == 0) A.partials[blockIdx.x] = s;
  if (A.fin_ticket != nullptr) {
    // fused finalize: the CTA that arrives last adds up all partial sums (one launch less)
    __shared__ bool last;
    __shared__ double fin_red[EXACT_THREADS];
    if (threadIdx.x == 0) {
      __threadfence();
      last = (atomicAdd(A.fin_ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
      __threadfence();
      finalize_body<EXACT_THREADS>(A.fin, fin_red);
      if (threadIdx.x == 0) *A.fin_ticket = 0u;
    }
  }
}

int launch_exact(const ExactArgs &args, cudaStream_t stream, int blocks)
{
  bool const general = args.fast_exp == 3 && (args.P.bc_type == 1 || args.P.bc_type == 3);
  if (general) exact_kernel<true><<<blocks, EXACT_THREADS, 0, stream>>>(args);
  else exact_kernel<false><<<blocks, EXACT_THREADS, 0, stream>>>(args);
  note_launch();
  return (int)cudaGetLastError();
}

// entries (k, k), k < n, and their count: a batch of independent single pairs for exact_kernel
__global__ void fill_diagonal_kernel(unsigned long long *list, unsigned int *count, int n)
{
  int const k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) list[k] = ((unsigned long long)(unsigned)k << 32) | (unsigned long long)(unsigned)k;
  if (k == 0) *count = (unsigned)n;
}

int launch_fill_diagonal(unsigned long long *list, unsigned int *count, int n, cudaStream_t stream)
{
  fill_diagonal_kernel<<<(n + 255) / 256, 256, 0, stream>>>(list, count, n);
  note_launch();
  return (int)cudaGetLastError();
}

// ============================ dense fallback ==================================================

__global__ void __launch_bounds__(EXACT_THREADS) dense_kernel(const __grid_constant__ DenseArgs A)
{
  __shared__ double red[EXACT_THREADS / 32];
  if (A.gate != nullptr && *A.gate != A.gate_on) {
    if (threadIdx.x == 0) A.partials[blockIdx.x] = 0.0;
    return;
  }
  bool const grad = A.flags & EF_GRADIENTS;
  bool const rebuild = (A.flags & EF_REBUILD_PAIRLIST) && A.pl;
  double fsum = 0.0;
  unsigned long long const rows = (unsigned long long)(A.i_end - A.i_begin);
  unsigned long long const total = rows * (unsigned long long)A.n2;
  for (unsigned long long e = (unsigned long long)blockIdx.x * EXACT_THREADS + threadIdx.x;
       e < total; e += (unsigned long long)gridDim.x * EXACT_THREADS) {
    int const i = A.i_begin + (int)(e / (unsigned long long)A.n2);
    int const j = (int)(e % (unsigned long long)A.n2);
    if (A.self && j <= i) continue;
    double Gx, Gy, Gz;
    double const F = pair_exact(A.P, A.x1[i], A.y1[i], A.z1[i], A.x2[j], A.y2[j], A.z2[j],
                                A.flags, Gx, Gy, Gz);
    fsum += F;
    if (grad && F > 0.0) {
      atomicAdd(&A.g1x[i], -Gx);
      atomicAdd(&A.g1y[i], -Gy);
      atomicAdd(&A.g1z[i], -Gz);
      atomicAdd(&A.g2x[j], Gx);
      atomicAdd(&A.g2y[j], Gy);
      atomicAdd(&A.g2z[j], Gz);
    }
    if (rebuild && F > 0.0) {
      unsigned long long const slot = atomicAdd(A.pl_count, 1ull);
      if (slot < A.pl_cap) A.pl[slot] = ((unsigned long long)(unsigned)i << 32) | (unsigned)j;
    }
  }
  double const s = block_sum<EXACT_THREADS>(fsum, red);
  if (threadIdx.x == 0) A.partials[blockIdx.x] = s;
}

int launch_dense(const DenseArgs &args, cudaStream_t stream)
{
  dense_kernel<<<EXACT_BLOCKS, EXACT_THREADS, 0, stream>>>(args);
  note_launch();
  return (int)cudaGetLastError();
}

// ============================ distancePairs ====================================================
// colvar::distance_pairs (src/colvarcomp_distances.cpp:486-548): no reduction at all -- the
// value IS the N1 x N2 matrix of distances, the force of pair (i, j) its force component times
// the unit vector.  One thread per pair, reference operation order (position_distance,
// rvector::norm / unit, src/colvartypes.h:837-841); HBM-bound: 8 B written per pair.
__device__ __forceinline__ double pair_distance_exact(PairParams const &P, double x1, double y1,
                                                      double z1, double x2, double y2, double z2,
                                                      double &dx, double &dy, double &dz)
{
  position_distance_exact(P, x1, y1, z1, x2, y2, z2, dx, dy, dz);
  return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

__global__ void __launch_bounds__(256) distance_pairs_kernel(const __grid_constant__ DistPairsArgs A)
{
  size_t const total = (size_t)A.n1 * (size_t)A.n2;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    int const i = (int)(e / (size_t)A.n2), j = (int)(e % (size_t)A.n2);
    double dx, dy, dz;
    double const d = pair_distance_exact(A.P, A.x1[i], A.y1[i], A.z1[i], A.x2[j], A.y2[j], A.z2[j],
                                         dx, dy, dz);
    if (A.force == nullptr) {
      A.dist[e] = d;
    } else {
      // f2 = force_ij * dv.unit(); group1 atom gets -f2, group2 atom +f2
      double const f = A.force[e];
      double ux = 1.0, uy = 0.0, uz = 0.0;
      if (d > 0.0) {
        ux = __ddiv_rn(dx, d);
        uy = __ddiv_rn(dy, d);
        uz = __ddiv_rn(dz, d);
      }
      double const fx = __dmul_rn(f, ux), fy = __dmul_rn(f, uy), fz = __dmul_rn(f, uz);
      atomicAdd(&A.f1x[i], -fx);
      atomicAdd(&A.f1y[i], -fy);
      atomicAdd(&A.f1z[i], -fz);
      atomicAdd(&A.f2x[j], fx);
      atomicAdd(&A.f2y[j], fy);
      atomicAdd(&A.f2z[j], fz);
    }
  }
}

int launch_distance_pairs(const DistPairsArgs &args, cudaStream_t stream)
{
  size_t const total = (size_t)args.n1 * (size_t)args.n2;
  int const blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  distance_pairs_kernel<<<std::max(blocks, 1), 256, 0, stream>>>(args);
  note_launch();
  return (int)cudaGetLastError();
}

// ============================ centre-of-mass variants ===========================================

__global__ void __launch_bounds__(EXACT_THREADS) center_kernel(const __grid_constant__ CenterArgs A)
{
  __shared__ double red[EXACT_THREADS / 32];
  bool const grad = A.flags & EF_GRADIENTS;
  bool const use_pl = A.flags & EF_USE_PAIRLIST;
  if (A.gate != nullptr && *A.gate == 2) {  // skipped launch of a captured evaluation
    if (threadIdx.x == 0) A.partials[blockIdx.x] = 0.0;
    return;
  }
  bool const rebuild = A.gate ? (*A.gate == 1) : (bool)(A.flags & EF_REBUILD_PAIRLIST);
  double const cx = A.center[0], cy = A.center[1], cz = A.center[2];
  double fsum = 0.0, sgx = 0.0, sgy = 0.0, sgz = 0.0;
  if (A.n == 0) {
    // both sides are centres: one pair (groupCoord; coordNum with both groups centre-only, which
    // may carry a one-entry pair list: src/colvarcomp_coordnums.cpp:219-232)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      bool const within = !use_pl || rebuild || A.pairlist[0];
      double Gx = 0.0, Gy = 0.0, Gz = 0.0, F = 0.0;
      if (within) {
        F = pair_exact(A.P, cx, cy, cz, A.center2[0], A.center2[1], A.center2[2], A.flags, Gx,
                       Gy, Gz);
      }
      if (use_pl && rebuild) A.pairlist[0] = F > 0.0 ? 1 : 0;
      fsum = F;
      if (grad && F > 0.0) {
        A.center_grad[0] += -Gx;
        A.center_grad[1] += -Gy;
        A.center_grad[2] += -Gz;
        A.center2_grad[0] += Gx;
        A.center2_grad[1] += Gy;
        A.center2_grad[2] += Gz;
      }
    }
  } else {
    for (int a = blockIdx.x * EXACT_THREADS + threadIdx.x; a < A.n;
         a += gridDim.x * EXACT_THREADS) {
      bool const within = !use_pl || rebuild || A.pairlist[a];
      double F = 0.0, Gx = 0.0, Gy = 0.0, Gz = 0.0;
      if (within) {
        if (A.center_is_group1) {
          F = pair_exact(A.P, cx, cy, cz, A.x[a], A.y[a], A.z[a], A.flags, Gx, Gy, Gz);
        } else {
          F = pair_exact(A.P, A.x[a], A.y[a], A.z[a], cx, cy, cz, A.flags, Gx, Gy, Gz);
        }
      }
      if (use_pl && rebuild) A.pairlist[a] = F > 0.0 ? 1 : 0;
      fsum += F;
      if (grad && F > 0.0) {
        // G is the gradient w.r.t. the second position of the pair
        double const s = A.center_is_group1 ? 1.0 : -1.0;
        atomicAdd(&A.gx[a], s * Gx);
        atomicAdd(&A.gy[a], s * Gy);
        atomicAdd(&A.gz[a], s * Gz);
        sgx -= s * Gx;
        sgy -= s * Gy;
        sgz -= s * Gz;
      }
    }
  }
  double const s = block_sum<EXACT_THREADS>(fsum, red);
  if (threadIdx.x == 0) A.partials[blockIdx.x] = s;
  if (A.n != 0 && grad) {
    double const tx = block_sum<EXACT_THREADS>(sgx, red);
    double const ty = block_sum<EXACT_THREADS>(sgy, red);
    double const tz = block_sum<EXACT_THREADS>(sgz, red);
    if (threadIdx.x == 0) {
      atomicAdd(&A.center_grad[0], tx);
      atomicAdd(&A.center_grad[1], ty);
      atomicAdd(&A.center_grad[2], tz);
    }
  }
}

int launch_center(const CenterArgs &args, cudaStream_t stream)
{
  int blocks = args.n == 0 ? 1 : (args.n + EXACT_THREADS - 1) / EXACT_THREADS;
  if (blocks > EXACT_BLOCKS) blocks = EXACT_BLOCKS;
  center_kernel<<<blocks, EXACT_THREADS, 0, stream>>>(args);
  note_launch();
  return (int)cudaGetLastError();
}

// ============================ finalize ===========================================================

__global__ void __launch_bounds__(256) finalize_kernel(const __grid_constant__ FinalizeArgs A)
{
  __shared__ double red[256];
  finalize_body<256>(A, red);
}

int launch_finalize(const FinalizeArgs &args, cudaStream_t stream)
{
  finalize_kernel<<<1, 256, 0, stream>>>(args);
  note_launch();
  return (int)cudaGetLastError();
}

// mode word of a captured evaluation: h_mode[0] = 1 rebuild / 0 walk (or plain evaluation),
// h_mode[1] != 0: skip this launch altogether (mode 2: no kernel of the evaluation does any work)
__global__ void fetch_gate_kernel(const int *h_mode_dev, int *d_gate, int with_list)
{
  *d_gate = h_mode_dev[1] != 0 ? 2 : (with_list ? h_mode_dev[0] : 0);
}

int launch_fetch_gate(const int *h_mode_dev, int *d_gate, int with_list, cudaStream_t stream)
{
  fetch_gate_kernel<<<1, 1, 0, stream>>>(h_mode_dev, d_gate, with_list);
  note_launch();
  return (int)cudaGetLastError();
}

// distanceInv epilogue: every thread recomputes the two scalars from the reduced sum (cheaper
// than a second launch); a cooperative-groups-free two-phase scheme -- the raw sum is left
// untouched until all blocks have read it: block 0 publishes x through a separate kernel.
__global__ void __launch_bounds__(256) distinv_scale_kernel(const __grid_constant__ DistInvPostArgs A)
{
  double const sum = *A.d_value;
  double const x = pow(sum * (1.0 / A.npairs), -1.0 / (double)A.exponent);
  double const dxdsum = (-1.0 / (double)A.exponent) * ipow_rt(x, A.exponent + 1) / A.npairs;
  for (size_t k = (size_t)blockIdx.x * 256 + threadIdx.x; k < A.count; k += (size_t)gridDim.x * 256) {
    A.grad[k] *= dxdsum;
  }
}

__global__ void distinv_value_kernel(const __grid_constant__ DistInvPostArgs A)
{
  double const sum = *A.d_value;
  double const x = pow(sum * (1.0 / A.npairs), -1.0 / (double)A.exponent);
  *A.d_value = x;
  if (A.h_value) *A.h_value = x;  // visible to the host once the kernel has completed
}

int launch_distinv_post(const DistInvPostArgs &args, cudaStream_t stream)
{
  if (args.grad && args.count > 0) {
    int blocks = (int)std::min<size_t>((args.count + 255) / 256, 296 * 4);
    distinv_scale_kernel<<<blocks, 256, 0, stream>>>(args);
  note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  distinv_value_kernel<<<1, 1, 0, stream>>>(args);
  note_launch();
  return (int)cudaGetLastError();
}

// ============================ atom-group kernels ================================================

__global__ void __launch_bounds__(256) gather_kernel(const double *__restrict__ proxy,
                                                     size_t pstride, int aos, GroupView g);

// Four atoms per thread: one 128-bit index load, twelve coordinate loads in flight, and per plane
// one 32-byte run of stores (the planes are padded to multiples of 32 atoms and 256-byte aligned,
// alloc_group).  HBM-bound: 4 B index + 24 B in + 24 B out per atom.
__device__ __forceinline__ void gather_four(const double *__restrict__ proxy, size_t pstride,
                                            int aos, GroupView const &g, int q)
{
  int const i0 = 4 * q;
  if (i0 >= g.n) return;
  if (i0 + 4 <= g.n) {
    int4 const p4 = __ldg(reinterpret_cast<const int4 *>(g.idx) + q);
    size_t const p[4] = {(size_t)p4.x, (size_t)p4.y, (size_t)p4.z, (size_t)p4.w};
    double v[3][4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
#pragma unroll
      for (int d = 0; d < 3; d++) v[d][a] = aos ? proxy[3 * p[a] + d] : proxy[d * pstride + p[a]];
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
      double2 *out = reinterpret_cast<double2 *>(g.data + d * g.stride + i0);
      out[0] = make_double2(v[d][0], v[d][1]);
      out[1] = make_double2(v[d][2], v[d][3]);
    }
  } else {
    for (int i = i0; i < g.n; i++) {
      size_t const p = (size_t)g.idx[i];
#pragma unroll
      for (int d = 0; d < 3; d++) {
        g.data[d * g.stride + i] = aos ? proxy[3 * p + d] : proxy[d * pstride + p];
      }
    }
  }
}

__global__ void __launch_bounds__(256) gather_kernel(const double *__restrict__ proxy,
                                                     size_t pstride, int aos, GroupView g)
{
  gather_four(proxy, pstride, aos, g, (int)(blockIdx.x * blockDim.x + threadIdx.x));
}

int launch_gather(const double *proxy, size_t pstride, int aos, const int *idx, int n,
                  double *pos, size_t stride, cudaStream_t stream)
{
  if (n <= 0) return 0;
  GroupView const g{idx, n, pos, stride, nullptr};
  gather_kernel<<<(n + 1023) / 1024, 256, 0, stream>>>(proxy, pstride, aos, g);
  note_launch();
  return (int)cudaGetLastError();
}

__global__ void __launch_bounds__(256) gather2_kernel(const double *__restrict__ proxy,
                                                      size_t pstride, int aos, GroupView a,
                                                      GroupView b, int blocks_a)
{
  bool const first = (int)blockIdx.x < blocks_a;
  int const q = ((int)blockIdx.x - (first ? 0 : blocks_a)) * (int)blockDim.x + (int)threadIdx.x;
  gather_four(proxy, pstride, aos, first ? a : b, q);
}

int launch_gather2(const double *proxy, size_t pstride, int aos, const GroupView &a,
                   const GroupView &b, cudaStream_t stream)
{
  int const ba = (a.n + 1023) / 1024, bb = (b.n + 1023) / 1024;
  if (ba + bb <= 0) return 0;
  gather2_kernel<<<ba + bb, 256, 0, stream>>>(proxy, pstride, aos, a, b, ba);
  note_launch();
  return (int)cudaGetLastError();
}

// calc_center_of_mass / calc_center_of_geometry (src/colvaratoms.cpp:1354-1395; GPU twin
// atoms_calc_cog_com_kernel, src/cuda/colvaratoms_kernel.cu:134-243).  Up to COM_BLOCKS CTAs
// reduce a grid-strided share each into scratch[block][6]; the last CTA to finish (ticket
// counter) sums the partials in block order, so the result does not depend on scheduling.
constexpr int COM_BLOCKS = 296;
constexpr int COM_THREADS = 256;

__global__ void __launch_bounds__(COM_THREADS) com_cog_kernel(const double *__restrict__ pos,
                                                              size_t stride,
                                                              const double *__restrict__ mass,
                                                              double total_mass, int n,
                                                              double *com, double *cog,
                                                              double *scratch,
                                                              unsigned int *ticket)
{
  __shared__ double red[6][COM_THREADS / 32];
  __shared__ bool last;
  double s[6] = {0, 0, 0, 0, 0, 0};
  int const step = gridDim.x * COM_THREADS;
  int i = blockIdx.x * COM_THREADS + threadIdx.x;
  // four atoms per trip: 16 independent loads in flight per thread
  for (; i + 3 * step < n; i += 4 * step) {
    double m[4], x[4], y[4], z[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      int const k = i + u * step;
      m[u] = mass[k];
      x[u] = pos[k];
      y[u] = pos[stride + k];
      z[u] = pos[2 * stride + k];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      s[0] = fma(m[u], x[u], s[0]);
      s[1] = fma(m[u], y[u], s[1]);
      s[2] = fma(m[u], z[u], s[2]);
      s[3] += x[u];
      s[4] += y[u];
      s[5] += z[u];
    }
  }
  for (; i < n; i += step) {
    double const m = mass[i];
    double const x = pos[i], y = pos[stride + i], z = pos[2 * stride + i];
    s[0] = fma(m, x, s[0]);
    s[1] = fma(m, y, s[1]);
    s[2] = fma(m, z, s[2]);
    s[3] += x;
    s[4] += y;
    s[5] += z;
  }
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 6; c++) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], d);
    if (lane == 0) red[c][warp] = s[c];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0.0;
    for (int w = 0; w < COM_THREADS / 32; w++) v += red[threadIdx.x][w];
    scratch[blockIdx.x * 6 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  // last CTA: partials of block b go to thread b % 256 (at most two per thread, loaded
  // together), then the same fixed shuffle tree -- independent of which CTA came last
  double t[6];
#pragma unroll
  for (int c = 0; c < 6; c++) {
    double v = 0.0;
    for (unsigned bb = threadIdx.x; bb < gridDim.x; bb += COM_THREADS) {
      v += __ldcg(&scratch[bb * 6 + c]);
    }
    t[c] = v;
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 6; c++) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t[c] += __shfl_xor_sync(0xffffffffu, t[c], d);
    if (lane == 0) red[c][warp] = t[c];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0.0;
    for (int w = 0; w < COM_THREADS / 32; w++) v += red[threadIdx.x][w];
    if (threadIdx.x < 3) com[threadIdx.x] = v / total_mass;
    else cog[threadIdx.x - 3] = v / (double)n;
  }
  if (threadIdx.x == 0) *ticket = 0u;  // ready for the next launch
}

int launch_com_cog(const double *pos, size_t stride, const double *mass, double total_mass,
                   int n, double *com, double *cog, double *scratch, unsigned int *ticket,
                   cudaStream_t stream)
{
  int blocks = (n + COM_THREADS * 8 - 1) / (COM_THREADS * 8);
  blocks = blocks < 1 ? 1 : (blocks > COM_BLOCKS ? COM_BLOCKS : blocks);
  com_cog_kernel<<<blocks, COM_THREADS, 0, stream>>>(pos, stride, mass, total_mass, n, com, cog,
                                                     scratch, ticket);
  note_launch();
  return (int)cudaGetLastError();
}

__global__ void translate_kernel(double *pos, size_t stride, int n, const double *t, double s)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pos[i] += s * t[0];
  pos[stride + i] += s * t[1];
  pos[2 * stride + i] += s * t[2];
}

int launch_translate(double *pos, size_t stride, int n, const double *t_dev, double s,
                     cudaStream_t stream)
{
  if (n <= 0) return 0;
  translate_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pos, stride, n, t_dev, s);
  note_launch();
  return (int)cudaGetLastError();
}

// cvm::quaternion::rotation_matrix (src/colvartypes.h)
__device__ __forceinline__ void quat_to_matrix(const double *q, double R[3][3])
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

__global__ void rotate_kernel(double *pos, size_t stride, int n, const double *q)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double R[3][3];
  quat_to_matrix(q, R);
  double const x = pos[i], y = pos[stride + i], z = pos[2 * stride + i];
  pos[i] = R[0][0] * x + R[0][1] * y + R[0][2] * z;
  pos[stride + i] = R[1][0] * x + R[1][1] * y + R[1][2] * z;
  pos[2 * stride + i] = R[2][0] * x + R[2][1] * y + R[2][2] * z;
}

int launch_rotate(double *pos, size_t stride, int n, const double *q_dev, cudaStream_t stream)
{
  if (n <= 0) return 0;
  rotate_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pos, stride, n, q_dev);
  note_launch();
  return (int)cudaGetLastError();
}

__global__ void weighted_gradient_kernel(const double *__restrict__ w, const double *g, int n,
                                         double *grad, size_t stride)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  grad[i] = w[i] * g[0];
  grad[stride + i] = w[i] * g[1];
  grad[2 * stride + i] = w[i] * g[2];
}

int launch_weighted_gradient(const double *w, const double *g_dev, int n, double *grad,
                             size_t stride, cudaStream_t stream)
{
  if (n <= 0) return 0;
  weighted_gradient_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w, g_dev, n, grad, stride);
  note_launch();
  return (int)cudaGetLastError();
}

// atom_group::apply_colvar_force (src/colvaratoms.cpp:1637-1689): atomics because several
// groups / components may push on the same proxy atom.
__global__ void scatter_force_kernel(const double *__restrict__ grad, size_t stride,
                                     const int *__restrict__ idx, int n, double f,
                                     const double *q, double *proxy_force, size_t pstride,
                                     int aos)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double gx = grad[i], gy = grad[stride + i], gz = grad[2 * stride + i];
  if (q) {
    // rot.inverse().matrix() = transpose of R(q)
    double R[3][3];
    quat_to_matrix(q, R);
    double const tx = R[0][0] * gx + R[1][0] * gy + R[2][0] * gz;
    double const ty = R[0][1] * gx + R[1][1] * gy + R[2][1] * gz;
    double const tz = R[0][2] * gx + R[1][2] * gy + R[2][2] * gz;
    gx = tx;
    gy = ty;
    gz = tz;
  }
  size_t const p = (size_t)idx[i];
  if (aos) {
    atomicAdd(&proxy_force[3 * p], f * gx);
    atomicAdd(&proxy_force[3 * p + 1], f * gy);
    atomicAdd(&proxy_force[3 * p + 2], f * gz);
  } else {
    atomicAdd(&proxy_force[p], f * gx);
    atomicAdd(&proxy_force[pstride + p], f * gy);
    atomicAdd(&proxy_force[2 * pstride + p], f * gz);
  }
}

int launch_scatter_force(const double *grad, size_t stride, const int *idx, int n, double f,
                         const double *q_dev, double *proxy_force, size_t pstride, int aos,
                         cudaStream_t stream)
{
  if (n <= 0) return 0;
  scatter_force_kernel<<<(n + 255) / 256, 256, 0, stream>>>(grad, stride, idx, n, f, q_dev,
                                                            proxy_force, pstride, aos);
  note_launch();
  return (int)cudaGetLastError();
}

__global__ void scatter_force2_kernel(GroupView a, GroupView b, int blocks_a, double f,
                                      double *proxy_force, size_t pstride, int aos)
{
  bool const first = (int)blockIdx.x < blocks_a;
  GroupView const &g = first ? a : b;
  int const i = ((int)blockIdx.x - (first ? 0 : blocks_a)) * (int)blockDim.x + (int)threadIdx.x;
  if (i >= g.n) return;
  double gx = g.data[i], gy = g.data[g.stride + i], gz = g.data[2 * g.stride + i];
  if (g.q) {
    double R[3][3];
    quat_to_matrix(g.q, R);
    double const tx = R[0][0] * gx + R[1][0] * gy + R[2][0] * gz;
    double const ty = R[0][1] * gx + R[1][1] * gy + R[2][1] * gz;
    double const tz = R[0][2] * gx + R[1][2] * gy + R[2][2] * gz;
    gx = tx;
    gy = ty;
    gz = tz;
  }
  size_t const p = (size_t)g.idx[i];
  if (aos) {
    atomicAdd(&proxy_force[3 * p], f * gx);
    atomicAdd(&proxy_force[3 * p + 1], f * gy);
    atomicAdd(&proxy_force[3 * p + 2], f * gz);
  } else {
    atomicAdd(&proxy_force[p], f * gx);
    atomicAdd(&proxy_force[pstride + p], f * gy);
    atomicAdd(&proxy_force[2 * pstride + p], f * gz);
  }
}

int launch_scatter_force2(const GroupView &a, const GroupView &b, double f, double *proxy_force,
                          size_t pstride, int aos, cudaStream_t stream)
{
  int const ba = (a.n + 255) / 256, bb = (b.n + 255) / 256;
  if (ba + bb <= 0) return 0;
  scatter_force2_kernel<<<ba + bb, 256, 0, stream>>>(a, b, ba, f, proxy_force, pstride, aos);
  note_launch();
  return (int)cudaGetLastError();
}

__global__ void repack_kernel(const double *__restrict__ src, size_t sstride, double *dst,
                              size_t dstride, int n, int accumulate)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    double const v = src[d * sstride + i];
    // accumulate: the destination may be shared with other components running concurrently
    // in the same graph (src/cuda/colvarcomp_distance_kernel.cu:137 does the same)
    if (accumulate) atomicAdd(&dst[d * dstride + i], v);
    else dst[d * dstride + i] = v;
  }
}

int launch_repack(const double *src, size_t sstride, double *dst, size_t dstride, int n,
                  int accumulate, cudaStream_t stream)
{
  if (n <= 0) return 0;
  repack_kernel<<<(n + 255) / 256, 256, 0, stream>>>(src, sstride, dst, dstride, n, accumulate);
  note_launch();
  return (int)cudaGetLastError();
}

// ============================ measurement helpers =================================================

// FP64 roofline denominator: 16 independent DFMA chains per thread, 256 DFMA per loop trip.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters)
{
  double a[16];
  double const b = 1.0000001, c = 1e-9;
#pragma unroll
  for (int k = 0; k < 16; k++) a[k] = (double)(threadIdx.x + k) * 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 16; rep++) {
#pragma unroll
      for (int k = 0; k < 16; k++) a[k] = fma(a[k], b, c);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; k++) s += a[k];
  if (s == 12345.678) out[0] = s;  // keep the chains alive
}

int launch_fp64_peak(double *d_out, int iters, int blocks, cudaStream_t stream)
{
  fp64_peak_kernel<<<blocks, 256, 0, stream>>>(d_out, iters);
  note_launch();
  return (int)cudaGetLastError();
}

__global__ void rcp_selftest_kernel(const double *x, int n, double *maxulp)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  double err = 0.0;
  if (i < n) {
    double const q = x[i];
    double const y = fast_rcp(q);
    double const yref = __ddiv_rn(1.0, q);
    // distance in units of the last place of the correctly rounded result
    long long const a = __double_as_longlong(y), b = __double_as_longlong(yref);
    err = (double)(a > b ? a - b : b - a);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) err = fmax(err, __shfl_xor_sync(0xffffffffu, err, d));
  if ((threadIdx.x & 31) == 0) {
    // non-negative doubles order like their bit patterns
    atomicMax((unsigned long long *)maxulp, (unsigned long long)__double_as_longlong(err));
  }
}

int launch_rcp_selftest(const double *d_x, int n, double *d_maxulp, cudaStream_t stream)
{
  rcp_selftest_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_x, n, d_maxulp);
  note_launch();
  return (int)cudaGetLastError();
}

}  // namespace b200cv
