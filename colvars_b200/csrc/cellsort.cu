// cellsort.cu -- see cellsort.cuh.  Grid of the group2 atoms and sort of both groups.
#include "cellsort.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "kernels.cuh"

namespace b200cv {

namespace {

#ifndef B200CV_SORT_XCELL
#define B200CV_SORT_XCELL 0.5  // cell edge along x in units of the cube that holds 32 atoms
#endif


// ---- bounding box of group2 -> grid with cells about as large as the volume of 32 atoms ------
// doubles as unsigned integers that order the same way (for atomicMin / atomicMax)
__device__ __forceinline__ unsigned long long ordered_bits(double v)
{
  unsigned long long const u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double from_ordered_bits(unsigned long long u)
{
  u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
  return __longlong_as_double((long long)u);
}

constexpr int BBOX_THREADS = 256;

__device__ void sort_grid(unsigned long long *box, int n, int cell_bits, SortGrid *grid);

// box[0..2] = min, box[3..5] = max (ordered bits); armed by sort_grid for the next time
__global__ void __launch_bounds__(BBOX_THREADS)
sort_bbox_kernel(const double *__restrict__ x, const double *__restrict__ y,
                 const double *__restrict__ z, int n, unsigned long long *box, int cell_bits,
                 SortGrid *grid, const int *gate)
{
  __shared__ double red[6][BBOX_THREADS / 32];
  if (gate != nullptr && *gate != 1) return;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i = blockIdx.x * BBOX_THREADS + threadIdx.x; i < n; i += gridDim.x * BBOX_THREADS) {
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
  if (threadIdx.x < 6) {
    int const d = threadIdx.x;
    double v = red[d][0];
    for (int w = 1; w < BBOX_THREADS / 32; w++) v = d < 3 ? fmin(v, red[d][w]) : fmax(v, red[d][w]);
    if (d < 3) atomicMin(&box[d], ordered_bits(v));
    else atomicMax(&box[d], ordered_bits(v));
    __threadfence();
  }
  // the CTA that arrives last turns the box into the grid (box[6]: ticket, re-armed)
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned *const ticket = reinterpret_cast<unsigned *>(box + 6);
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      __threadfence();
      *ticket = 0u;
      sort_grid(box, n, cell_bits, grid);
    }
  }
}

// the grid from the finished box (one thread: the last CTA of sort_bbox_kernel)
__device__ void sort_grid(unsigned long long *box, int n, int cell_bits, SortGrid *grid)
{
  double a[3], ext[3], emax = 0.0;
  for (int d = 0; d < 3; d++) {
    double const l = from_ordered_bits(((volatile unsigned long long *)box)[d]);
    double const h = from_ordered_bits(((volatile unsigned long long *)box)[3 + d]);
    a[d] = l;
    ext[d] = fmax(h - l, 0.0);
    emax = fmax(emax, ext[d]);
    box[d] = ordered_bits(1.0e300);       // armed for the next rebuild
    box[3 + d] = ordered_bits(-1.0e300);
  }
  // cell edge: the cube that holds 32 atoms at the mean density of the occupied box (flat or
  // linear systems: only the axes that have an extent count); half of it along x, where the
  // range a row search covers is rounded to whole cells
  int k = 0;
  double vol = 1.0;
  for (int d = 0; d < 3; d++) {
    if (ext[d] > 1.0e-9 * emax && ext[d] > 0.0) {
      vol *= ext[d];
      k++;
    }
  }
  double s = 1.0;
  if (k > 0) s = pow(32.0 * vol / (double)max(n, 1), 1.0 / (double)k);
  int dim[3];
  double cell[3];
  for (int attempt = 0; attempt < 64; attempt++) {
    long long ncells = 1;
    for (int d = 0; d < 3; d++) {
      cell[d] = fmax(d == 0 ? B200CV_SORT_XCELL * s : s, ext[d] / (double)SORT_MAX_DIM * (1.0 + 1.0e-9));
      if (!(cell[d] > 0.0)) cell[d] = 1.0;
      dim[d] = max(1, min((int)floor(ext[d] / cell[d]) + 1, SORT_MAX_DIM));
      ncells *= dim[d];
    }
    if (ncells <= (1ll << cell_bits)) break;
    s *= 1.26;  // coarser cells until the sort key fits
  }
  for (int d = 0; d < 3; d++) {
    grid->origin[d] = a[d];
    grid->inv_cell[d] = 1.0 / cell[d];
    grid->dim[d] = dim[d];
  }
  grid->ncells = dim[0] * dim[1] * dim[2];
  grid->nrows = dim[1] * dim[2];
}

__global__ void sort_key_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                const double *__restrict__ z, int n,
                                const SortGrid *__restrict__ grid, int *key, int *cnt,
                                const int *gate)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (gate != nullptr && *gate != 1) return;
  SortGrid const g = *grid;
  int const cx = sort_cell(x[i], g.origin[0], g.inv_cell[0], g.dim[0]);
  int const cy = sort_cell(y[i], g.origin[1], g.inv_cell[1], g.dim[1]);
  int const cz = sort_cell(z[i], g.origin[2], g.inv_cell[2], g.dim[2]);
  // position inside the cell along x, SORT_FINE_BITS bits: consecutive atoms of a row of cells
  // are neighbours along x as well
  double const fx = (x[i] - g.origin[0]) * g.inv_cell[0] - (double)cx;
  int const fine = (int)fmin(fmax(fx * (double)(1 << SORT_FINE_BITS), 0.0),
                             (double)((1 << SORT_FINE_BITS) - 1));
  int const cell = ((cz * g.dim[1] + cy) * g.dim[0]) + cx;
  key[i] = (cell << SORT_FINE_BITS) | fine;
  keysort_count(cnt, cell);
}

// coordinates into sorted order; both sides in one launch
__global__ void sort_gather_kernel(const double *__restrict__ x1, const double *__restrict__ y1,
                                   const double *__restrict__ z1, const int *__restrict__ s1,
                                   int n1, double *xs1, double *ys1, double *zs1,
                                   const double *__restrict__ x2, const double *__restrict__ y2,
                                   const double *__restrict__ z2, const int *__restrict__ s2,
                                   int n2, double *xs2, double *ys2, double *zs2, double *gs2,
                                   size_t gs_stride2, int blocks1, const int *gate, int gate_on)
{
  if (gate != nullptr && *gate != gate_on) return;
  if ((int)blockIdx.x < blocks1) {
    int const p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) xs1[n1] = ys1[n1] = zs1[n1] = SORT_FAR;
    if (p >= n1) return;
    int const i = s1[p];
    xs1[p] = x1[i];
    ys1[p] = y1[i];
    zs1[p] = z1[i];
  } else {
    int const p = ((int)blockIdx.x - blocks1) * blockDim.x + threadIdx.x;
    if (p == 0) {
      xs2[n2] = ys2[n2] = zs2[n2] = SORT_FAR;
      gs2[n2] = gs2[gs_stride2 + n2] = gs2[2 * gs_stride2 + n2] = 0.0;
    }
    if (p >= n2) return;
    gs2[p] = gs2[gs_stride2 + p] = gs2[2 * gs_stride2 + p] = 0.0;
    int const j = s2[p];
    xs2[p] = x2[j];
    ys2[p] = y2[j];
    zs2[p] = z2[j];
  }
}

void side_free(SortedGroup &S)
{
  cudaFree(S.d_key);
  cudaFree(S.d_sorted);
  keysort_free(S.sort);
  cudaFree(S.d_cell_start);
  cudaFree(S.d_xs);
  cudaFree(S.d_ys);
  cudaFree(S.d_zs);
  cudaFree(S.d_gs);
  S = SortedGroup();
}

int side_setup(SortedGroup &S, int n, int cell_bits)
{
  S.n = n;
  cudaError_t e;
  size_t const nb = sizeof(int) * (size_t)std::max(n, 1);
  if ((e = cudaMalloc(&S.d_key, nb)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&S.d_sorted, nb)) != cudaSuccess) return (int)e;
  if (int const rc = keysort_setup(S.sort, n, 1 << cell_bits)) return rc;
  size_t const nbd = sizeof(double) * ((size_t)std::max(n, 1) + 1);  // + the far-away point
  if ((e = cudaMalloc(&S.d_xs, nbd)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&S.d_ys, nbd)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&S.d_zs, nbd)) != cudaSuccess) return (int)e;
  S.gs_stride = ((size_t)std::max(n, 1) + 2) & ~(size_t)1;
  if ((e = cudaMalloc(&S.d_gs, 3 * sizeof(double) * S.gs_stride)) != cudaSuccess) return (int)e;
  if ((e = cudaMemset(S.d_gs, 0, 3 * sizeof(double) * S.gs_stride)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&S.d_cell_start, sizeof(int) * (((size_t)1 << cell_bits) + 2))) != cudaSuccess) {
    return (int)e;
  }
  return 0;
}

int sort_side(CellSort &B, SortedGroup &S, const double *x, const double *y, const double *z,
              const int *gate, cudaStream_t stream)
{
  sort_key_kernel<<<(S.n + 255) / 256, 256, 0, stream>>>(x, y, z, S.n, B.d_grid, S.d_key,
                                                         S.sort.d_cnt, gate);
  note_launch();
  return keysort_run(S.sort, S.d_key, SORT_FINE_BITS, S.d_sorted, S.d_cell_start, gate, stream);
}

}  // namespace

int cellsort_setup(CellSort &B, int n1, int n2)
{
  bool const self = n1 <= 0;
  if (B.d_grid && B.side[1].n == n2 && B.side[0].n == (self ? 0 : n1)) return 0;
  cellsort_free(B);
  B.self = self;
  // about n2 / 16 cells (32 atoms per cell, half cells along x); the grid kernel coarsens the
  // cells if rounding up per axis takes it beyond 2^cell_bits
  int bits = 8;
  while (bits < SORT_MAX_CELL_BITS && ((long long)1 << bits) < (long long)std::max(n2, 1) / 2 + 64) bits++;
  B.cell_bits = bits;
  cudaError_t e;
  if ((e = cudaMalloc(&B.d_grid, sizeof(SortGrid))) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&B.d_box, 7 * sizeof(unsigned long long))) != cudaSuccess) return (int)e;
  if ((e = cudaMemset(B.d_box, 0, 7 * sizeof(unsigned long long))) != cudaSuccess) return (int)e;
  {
    // min slots at +1e300, max slots at -1e300 (ordered bits, see sort_bbox_kernel)
    auto ordered = [](double v) {
      unsigned long long u;
      std::memcpy(&u, &v, 8);
      return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    };
    unsigned long long const init[6] = {ordered(1e300), ordered(1e300), ordered(1e300),
                                        ordered(-1e300), ordered(-1e300), ordered(-1e300)};
    if ((e = cudaMemcpy(B.d_box, init, sizeof(init), cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
  }
  if ((e = cudaMalloc(&B.d_counts, 8 * sizeof(unsigned))) != cudaSuccess) return (int)e;
  if ((e = cudaMemset(B.d_counts, 0, 8 * sizeof(unsigned))) != cudaSuccess) return (int)e;
  int rc = side_setup(B.side[1], n2, bits);
  if (rc) return rc;
  if (!self) {
    rc = side_setup(B.side[0], n1, bits);
    if (rc) return rc;
  }
  return 0;
}

void cellsort_free(CellSort &B)
{
  cudaFree(B.d_grid);
  cudaFree(B.d_box);
  cudaFree(B.d_counts);
  side_free(B.side[0]);
  side_free(B.side[1]);
  B = CellSort();
}

int cellsort_build(CellSort &B, const double *x1, const double *y1, const double *z1,
                   const double *x2, const double *y2, const double *z2, const int *gate,
                   cudaStream_t stream)
{
  SortedGroup &J = B.side[1];
  int const bb = std::max(1, std::min(296, (J.n + BBOX_THREADS * 4 - 1) / (BBOX_THREADS * 4)));
  sort_bbox_kernel<<<bb, BBOX_THREADS, 0, stream>>>(x2, y2, z2, J.n, B.d_box, B.cell_bits, B.d_grid,
                                                    gate);
  note_launch();
  int rc = sort_side(B, J, x2, y2, z2, gate, stream);
  if (rc) return rc;
  if (!B.self) {
    rc = sort_side(B, B.side[0], x1, y1, z1, gate, stream);
    if (rc) return rc;
  }
  return (int)cudaGetLastError();
}

int cellsort_gather(CellSort &B, const double *x1, const double *y1, const double *z1,
                    const double *x2, const double *y2, const double *z2, const int *gate,
                    int gate_on, cudaStream_t stream)
{
  SortedGroup &I = B.side[0], &J = B.side[1];
  int const b1 = B.self ? 0 : (I.n + 255) / 256, b2 = (J.n + 255) / 256;
  sort_gather_kernel<<<b1 + b2, 256, 0, stream>>>(x1, y1, z1, I.d_sorted, B.self ? 0 : I.n, I.d_xs,
                                                  I.d_ys, I.d_zs, x2, y2, z2, J.d_sorted, J.n,
                                                  J.d_xs, J.d_ys, J.d_zs, J.d_gs, J.gs_stride, b1, gate,
                                                  gate_on);
  note_launch();
  return (int)cudaGetLastError();
}

}  // namespace b200cv
