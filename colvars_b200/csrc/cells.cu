// cells.cu -- see cells.cuh.  Grid construction (bounding box, cell ids, cell order (keysort.cuh), cell
// offsets) and the neighbour-cell pair kernel of the pair-list rebuild.
#include "cells.cuh"

#include <algorithm>
#include <cmath>


#include "kernels.cuh"

namespace b200cv {

namespace {

constexpr int GRID_THREADS = 1024;
#ifndef B200CV_CELL_MINB
#define B200CV_CELL_MINB 2
#endif
constexpr int CELL_MINB = B200CV_CELL_MINB;  // CTAs per SM the register allocation targets

// ---- bounding box of group2 -> grid parameters (one CTA; a rebuild happens every
// pairListFrequency steps, this is not on the per-step path) -------------------------------
__global__ void __launch_bounds__(GRID_THREADS)
grid_kernel(const double *__restrict__ x, const double *__restrict__ y,
            const double *__restrict__ z, int n, double cutx, double cuty, double cutz,
            CellGrid *grid, const int *gate)
{
  __shared__ double red[6][GRID_THREADS / 32];
  if (gate != nullptr && *gate != 1) return;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i = threadIdx.x; i < n; i += GRID_THREADS) {
    double const p[3] = {x[i], y[i], z[i]};
#pragma unroll
    for (int d = 0; d < 3; d++) {
      lo[d] = fmin(lo[d], p[d]);
      hi[d] = fmax(hi[d], p[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], s));
      hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], s));
    }
    if ((threadIdx.x & 31) == 0) {
      red[d][threadIdx.x >> 5] = lo[d];
      red[3 + d][threadIdx.x >> 5] = hi[d];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double const cut[3] = {cutx, cuty, cutz};
    int ncells = 1;
    for (int d = 0; d < 3; d++) {
      double a = red[d][0], b = red[3 + d][0];
      for (int w = 1; w < GRID_THREADS / 32; w++) {
        a = fmin(a, red[d][w]);
        b = fmax(b, red[3 + d][w]);
      }
      double const ext = fmax(b - a, 0.0);
      // at least one cut-off radius wide (with a margin far above the rounding of the cell
      // index), and at most CELL_MAX_DIM cells per axis
      double cell = fmax(cut[d] * (1.0 + 1.0e-9), ext / (double)CELL_MAX_DIM * (1.0 + 1.0e-9));
      if (!(cell > 0.0)) cell = 1.0;
      int dim = (int)floor(ext / cell) + 1;
      dim = max(1, min(dim, CELL_MAX_DIM));
      grid->origin[d] = a;
      grid->inv_cell[d] = 1.0 / cell;
      grid->dim[d] = dim;
      ncells *= dim;
    }
    grid->ncells = ncells;
    grid->periodic = 0;
    for (int a = 0; a < 3; a++) {
      for (int d = 0; d < 3; d++) grid->frac[a][d] = 0.0;
    }
  }
}

__global__ void set_grid_kernel(CellGrid value, CellGrid *grid, const int *gate)
{
  if (gate != nullptr && *gate != 1) return;
  *grid = value;
}

__device__ __forceinline__ int cell_coord(double p, double origin, double inv_cell, int dim)
{
  // clamped in floating point first: a query atom may be arbitrarily far outside the grid
  double const c = floor((p - origin) * inv_cell);
  return (int)fmin(fmax(c, -2.0), (double)dim + 1.0);
}

// periodic: cell of the atom's image inside the unit cell, along fractional axis a
__device__ __forceinline__ int cell_coord_periodic(CellGrid const &g, int a, double x, double y,
                                                   double z)
{
  double const t = fma(z, g.frac[a][2], fma(y, g.frac[a][1], x * g.frac[a][0]));
  double const f = t - floor(t);  // [0, 1]; 1.0 only by rounding for tiny negative t
  return min((int)(f * (double)g.dim[a]), g.dim[a] - 1);
}

__global__ void cell_key_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                const double *__restrict__ z, int n,
                                const CellGrid *__restrict__ grid, int *key, int *cnt,
                                const int *gate)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (gate != nullptr && *gate != 1) return;
  CellGrid const g = *grid;
  int cx, cy, cz;
  if (g.periodic) {
    cx = cell_coord_periodic(g, 0, x[i], y[i], z[i]);
    cy = cell_coord_periodic(g, 1, x[i], y[i], z[i]);
    cz = cell_coord_periodic(g, 2, x[i], y[i], z[i]);
  } else {
    cx = min(max(cell_coord(x[i], g.origin[0], g.inv_cell[0], g.dim[0]), 0), g.dim[0] - 1);
    cy = min(max(cell_coord(y[i], g.origin[1], g.inv_cell[1], g.dim[1]), 0), g.dim[1] - 1);
    cz = min(max(cell_coord(z[i], g.origin[2], g.inv_cell[2], g.dim[2]), 0), g.dim[2] - 1);
  }
  key[i] = (cz * g.dim[1] + cy) * g.dim[0] + cx;
  keysort_count(cnt, key[i]);
}


// coordinates in cell order: the pair kernel's candidate loads become contiguous
__global__ void cell_gather_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                   const double *__restrict__ z, const int *__restrict__ sorted,
                                   int n, double *__restrict__ xs, double *__restrict__ ys,
                                   double *__restrict__ zs, const int *gate, int gate_on)
{
  int const p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (gate != nullptr && *gate != gate_on) return;
  int const j = sorted[p];
  xs[p] = x[j];
  ys[p] = y[j];
  zs[p] = z[j];
}

// gradients accumulated in cell order go back to the group's planes (one thread per atom, plain
// read-modify-write: every atom appears once in the order)
__global__ void cell_scatter_add_kernel(const double *__restrict__ gs, const int *__restrict__ sorted,
                                        int n, double *gx, double *gy, double *gz,
                                        const int *gate, int gate_on)
{
  int const p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (gate != nullptr && *gate != gate_on) return;
  int const j = sorted[p];
  gx[j] += gs[p];
  gy[j] += gs[n + p];
  gz[j] += gs[2 * n + p];
}

// ---- candidate pairs: one warp per group1 row, lanes over the atoms of the 3 x 3 x-spans of
// neighbouring cells (cells adjacent in x are contiguous in key order) -------------------------
// GENERAL: triclinic cell, candidates through the sweep's image search (general_image) with the
// reference-order pair as the fallback for ambiguous decisions and threshold neighbours
template <bool GENERAL>
__global__ void __launch_bounds__(CELL_THREADS, CELL_MINB) cell_pairs_kernel(const __grid_constant__ CellPairArgs A)
{
  __shared__ double red[CELL_THREADS / 32];
  __shared__ double tritab[GENERAL ? 42 : 1];
  if constexpr (GENERAL) {
    if (threadIdx.x < 42) {
      tritab[threadIdx.x] = threadIdx.x < 39 ? A.P.tri_shift[threadIdx.x / 3][threadIdx.x % 3] : 0.0;
    }
    __syncthreads();
  }
  // listed neighbours of the warp's current row: written to the list as ONE contiguous run per
  // row (one counter atomic per run, coalesced stores); the list walk sums the row's gradient
  // over such a run before it touches memory
  constexpr int ROW_BUF = 480;
  __shared__ int jbuf[CELL_THREADS / 32][ROW_BUF];
  if (A.gate != nullptr && *A.gate != 1) {
    if (threadIdx.x == 0) A.partials[blockIdx.x] = 0.0;
    return;
  }
  CellGrid const g = *A.grid;
  int const lane = threadIdx.x & 31;
  int *const mybuf = jbuf[threadIdx.x >> 5];
  int nbuf = 0;  // warp-uniform
  auto flush_row = [&](int row_i) {
    if (nbuf > 0 && A.pl != nullptr) {
      __syncwarp();
      unsigned long long base = 0ull;
      if (lane == 0) base = atomicAdd(A.pl_count, (unsigned long long)nbuf);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (int k = lane; k < nbuf; k += 32) {
        unsigned long long const slot = base + (unsigned long long)k;
        if (slot < A.pl_cap) {
          A.pl[slot] = ((unsigned long long)(unsigned)row_i << 32) | (unsigned)mybuf[k];
        }
      }
      __syncwarp();
    }
    nbuf = 0;
  };
  int const warps_per_block = CELL_THREADS / 32;
  int const gwarp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  int const nwarps = gridDim.x * warps_per_block;
  bool const grad = A.flags & EF_GRADIENTS;
  bool const self_sorted = A.self && A.sorted_entries;
  double fsum = 0.0;
  // selfCoordNum: rows in cell order as well (row q of the sorted atoms; any disjoint split of
  // the rows over warps and ranks lists every pair once), so that the warps of a CTA work on
  // neighbouring cells and share their candidates' cache lines
  for (int q = A.i_begin + gwarp; q < A.i_end; q += nwarps) {
    int const i = A.self ? A.sorted2[q] : q;
    int const row_id = A.sorted_entries ? q : i;  // q == i unless selfCoordNum
    double const xi = A.x1[i], yi = A.y1[i], zi = A.z1[i];
    bool const per = g.periodic != 0;
    int const cx = per ? cell_coord_periodic(g, 0, xi, yi, zi)
                       : cell_coord(xi, g.origin[0], g.inv_cell[0], g.dim[0]);
    int const cy = per ? cell_coord_periodic(g, 1, xi, yi, zi)
                       : cell_coord(yi, g.origin[1], g.inv_cell[1], g.dim[1]);
    int const cz = per ? cell_coord_periodic(g, 2, xi, yi, zi)
                       : cell_coord(zi, g.origin[2], g.inv_cell[2], g.dim[2]);
    double gx = 0.0, gy = 0.0, gz = 0.0;
    // neighbouring cells: non-periodic -- clipped to the grid, the three x-cells as one span;
    // periodic -- wrapped around (dim >= 3, so the three are distinct), one cell at a time
    for (int dzc = -1; dzc <= 1; dzc++) {
      int nz = cz + dzc;
      if (per) nz = (nz + g.dim[2]) % g.dim[2];
      else if (nz < 0 || nz >= g.dim[2]) continue;
      for (int dyc = -1; dyc <= 1; dyc++) {
        int ny = cy + dyc;
        if (per) ny = (ny + g.dim[1]) % g.dim[1];
        else if (ny < 0 || ny >= g.dim[1]) continue;
        int const row = (nz * g.dim[1] + ny) * g.dim[0];
        int const nspans = per ? 3 : 1;
        for (int sxc = 0; sxc < nspans; sxc++) {
          int xa, xb;
          if (per) {
            xa = xb = (cx + sxc - 1 + g.dim[0]) % g.dim[0];
          } else {
            xa = max(cx - 1, 0);
            xb = min(cx + 1, g.dim[0] - 1);
            if (xa > xb) continue;
          }
          int begin = A.cell_start[row + xa];
          int const end = A.cell_start[row + xb + 1];
          // selfCoordNum in cell order: a pair belongs to the row that comes first in that order
          // (any strict order lists every pair once; b200cv_pairlist puts the smaller atom index
          // first again), so everything up to the row's own position is skipped -- half of the
          // neighbourhood goes away without being looked at
          if (self_sorted) begin = max(begin, q + 1);
          for (int p0 = begin; p0 < end; p0 += 32) {
            int const p = p0 + lane;
            double F = 0.0, Gx = 0.0, Gy = 0.0, Gz = 0.0;
            int j = -1;
            if (p < end) {
              j = A.sorted2[p];
              if (!A.self || self_sorted || j > i) {
                double const xj = A.xs[p], yj = A.ys[p], zj = A.zs[p];
                bool done = false;
                if (A.fast_exp == 3) {
                  // 6/12: the sweep's fast form for every pair that is not next to a decision
                  // threshold (same windows as the sweep), reference order otherwise
                  double dx = xj - xi, dy = yj - yi, dz = zj - zi;
                  bool ambiguous = false;
                  if constexpr (GENERAL) {
                    ambiguous = general_image(A.P, tritab, dx, dy, dz);
                  } else if (A.P.bc_type == 2) {
                    // orthorhombic minimum image, image index in the reference's operation
                    // order (src/colvars_system.h:176-184), as in the tile sweep
                    double const shx = floor(__dadd_rn(__dmul_rn(A.P.recip[0][0], dx), 0.5));
                    double const shy = floor(__dadd_rn(__dmul_rn(A.P.recip[1][1], dy), 0.5));
                    double const shz = floor(__dadd_rn(__dmul_rn(A.P.recip[2][2], dz), 0.5));
                    dx = fma(-shx, A.P.cell[0][0], dx);
                    dy = fma(-shy, A.P.cell[1][1], dy);
                    dz = fma(-shz, A.P.cell[2][2], dz);
                  }
                  double const sx = dx * A.P.inv_r0[0], sy = dy * A.P.inv_r0[1],
                               sz = dz * A.P.inv_r0[2];
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
                if (!done) F = pair_exact(A.P, xi, yi, zi, xj, yj, zj, A.flags, Gx, Gy, Gz);
              }
            }
            bool const listed = F > 0.0;
            fsum += F;
            if (grad && listed) {
              gx += Gx;
              gy += Gy;
              gz += Gz;
              atomicAdd(&A.g2x[j], Gx);
              atomicAdd(&A.g2y[j], Gy);
              atomicAdd(&A.g2z[j], Gz);
            }
            unsigned const m = __ballot_sync(0xffffffffu, listed);
            if (m != 0u) {
              if (listed) mybuf[nbuf + __popc(m & ((1u << lane) - 1u))] = A.sorted_entries ? p : j;
              nbuf += __popc(m);
              if (nbuf > ROW_BUF - 32) flush_row(row_id);
            }
          }
        }
      }
    }
    flush_row(row_id);
    if (grad) {
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        gx += __shfl_xor_sync(0xffffffffu, gx, s);
        gy += __shfl_xor_sync(0xffffffffu, gy, s);
        gz += __shfl_xor_sync(0xffffffffu, gz, s);
      }
      if (lane == 0 && (gx != 0.0 || gy != 0.0 || gz != 0.0)) {
        atomicAdd(&A.g1x[i], -gx);
        atomicAdd(&A.g1y[i], -gy);
        atomicAdd(&A.g1z[i], -gz);
      }
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, s);
  if (lane == 0) red[threadIdx.x >> 5] = fsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < warps_per_block; w++) s += red[w];
    A.partials[blockIdx.x] = s;
  }
}

}  // namespace

int cells_setup(CellBuffers &B, int n2)
{
  if (B.d_grid && B.n2 == n2) return 0;
  cells_free(B);
  B.n2 = n2;
  cudaError_t e;
  if ((e = cudaMalloc(&B.d_grid, sizeof(CellGrid))) != cudaSuccess) return (int)e;
  size_t const nb = sizeof(int) * (size_t)std::max(n2, 1);
  if ((e = cudaMalloc(&B.d_key, nb)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&B.d_sorted, nb)) != cudaSuccess) return (int)e;
  if (int const rc = keysort_setup(B.sort, n2, CELL_MAX)) return rc;
  size_t const nbd = sizeof(double) * (size_t)std::max(n2, 1);
  if ((e = cudaMalloc(&B.d_xs, nbd)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&B.d_ys, nbd)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&B.d_zs, nbd)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&B.d_gs, 3 * nbd)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&B.d_cell_start, sizeof(int) * ((size_t)CELL_MAX + 2))) != cudaSuccess) {
    return (int)e;
  }
  return 0;
}

void cells_free(CellBuffers &B)
{
  cudaFree(B.d_grid);
  cudaFree(B.d_key);
  cudaFree(B.d_sorted);
  keysort_free(B.sort);
  cudaFree(B.d_xs);
  cudaFree(B.d_ys);
  cudaFree(B.d_zs);
  cudaFree(B.d_gs);
  cudaFree(B.d_cell_start);
  B = CellBuffers();
}

bool cells_periodic_dims(const double cut[3], const double frac[9], int dim[3])
{
  bool const diagonal = frac[1] == 0.0 && frac[2] == 0.0 && frac[3] == 0.0 && frac[5] == 0.0 &&
                        frac[6] == 0.0 && frac[7] == 0.0;
  double const cmax = std::max(cut[0], std::max(cut[1], cut[2]));
  for (int a = 0; a < 3; a++) {
    double const len = std::sqrt(frac[3 * a] * frac[3 * a] + frac[3 * a + 1] * frac[3 * a + 1] +
                                 frac[3 * a + 2] * frac[3 * a + 2]);
    double const c = diagonal ? cut[a] : cmax;
    if (!(len > 0.0) || !(c > 0.0)) return false;
    double const width = 1.0 / len;  // of the unit cell, perpendicular to the faces of axis a
    double const n = std::floor(width / (c * (1.0 + 1.0e-9)));
    if (n < 3.0) return false;  // the three neighbours of a cell must be distinct
    dim[a] = (int)std::min(n, (double)CELL_MAX_DIM);
  }
  return true;
}

int cells_build(CellBuffers &B, const double *x2, const double *y2, const double *z2, int n2,
                const double cut[3], const double *periodic_frac, const int *gate,
                cudaStream_t stream)
{
  if (periodic_frac != nullptr) {
    CellGrid g{};
    int dim[3];
    if (!cells_periodic_dims(cut, periodic_frac, dim)) return (int)cudaErrorInvalidValue;
    g.ncells = 1;
    for (int d = 0; d < 3; d++) {
      g.origin[d] = 0.0;
      g.dim[d] = dim[d];
      g.inv_cell[d] = 0.0;  // unused: cells are counted in fractional coordinates
      for (int k = 0; k < 3; k++) g.frac[d][k] = periodic_frac[3 * d + k];
      g.ncells *= dim[d];
    }
    g.periodic = 1;
    set_grid_kernel<<<1, 1, 0, stream>>>(g, B.d_grid, gate);
  } else {
    grid_kernel<<<1, GRID_THREADS, 0, stream>>>(x2, y2, z2, n2, cut[0], cut[1], cut[2], B.d_grid,
                                                gate);
  }
  note_launch();
  cell_key_kernel<<<(n2 + 255) / 256, 256, 0, stream>>>(x2, y2, z2, n2, B.d_grid, B.d_key,
                                                        B.sort.d_cnt, gate);
  note_launch();
  if (int const rc = keysort_run(B.sort, B.d_key, 0, B.d_sorted, B.d_cell_start, gate, stream)) return rc;
  return launch_cell_gather(B, x2, y2, z2, n2, gate, 1, stream);
}

int launch_cell_gather(CellBuffers &B, const double *x2, const double *y2, const double *z2,
                       int n2, const int *gate, int gate_on, cudaStream_t stream)
{
  cell_gather_kernel<<<(n2 + 255) / 256, 256, 0, stream>>>(x2, y2, z2, B.d_sorted, n2, B.d_xs,
                                                           B.d_ys, B.d_zs, gate, gate_on);
  note_launch();
  return (int)cudaGetLastError();
}

int launch_cell_scatter_add(CellBuffers &B, double *g2x, double *g2y, double *g2z, int n2,
                            const int *gate, int gate_on, cudaStream_t stream)
{
  cell_scatter_add_kernel<<<(n2 + 255) / 256, 256, 0, stream>>>(B.d_gs, B.d_sorted, n2, g2x, g2y,
                                                                g2z, gate, gate_on);
  note_launch();
  return (int)cudaGetLastError();
}

int launch_cell_pairs(const CellPairArgs &args, cudaStream_t stream)
{
  bool const general = args.fast_exp == 3 && (args.P.bc_type == 1 || args.P.bc_type == 3);
  if (general) cell_pairs_kernel<true><<<CELL_BLOCKS, CELL_THREADS, 0, stream>>>(args);
  else cell_pairs_kernel<false><<<CELL_BLOCKS, CELL_THREADS, 0, stream>>>(args);
  note_launch();
  return (int)cudaGetLastError();
}

}  // namespace b200cv
