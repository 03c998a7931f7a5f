// cellsort.cuh -- both atom groups sorted along a cell grid (open systems): the order in which the
// sparse pair list (rows.cuh) keeps and walks its partners.
//
// Grid: the bounding box of group2, found on the device, cut into cells about as large as the
// volume 32 atoms take up (half that along x); sort key = cell index (x fastest) and, as the minor
// key, a fine x-coordinate inside the cell, so that atoms that follow each other in the order are
// neighbours in space.  group1 is sorted on the same grid (atoms outside the box end up in the
// outermost cells).  One sort per group and rebuild (keysort.cuh: the order of a stable sort by
// key); the coordinates are copied into sorted order again on every list step (the order stays
// that of the last rebuild).
#pragma once
#include <cuda_runtime.h>

#include "keysort.cuh"

namespace b200cv {

constexpr int SORT_MAX_DIM = 128;    // cells per axis at most
constexpr int SORT_FINE_BITS = 8;    // fine x-coordinate inside a cell: minor sort key
// Behind the n sorted coordinates of a group sits one more point, far away on every axis: the
// padding entries of a neighbour row (rows.cuh) point at it.  A pair with it is beyond any
// tolerance_l2_max (value masked to +0) and its gradient term is 2^-400 * 1e30: nothing next to
// a real contribution.  (Coordinates and cut-offs of 1e29 length units are out of scope.)
constexpr double SORT_FAR = 1.0e30;
constexpr int SORT_MAX_CELL_BITS = 21;

struct SortGrid {
  double origin[3];
  double inv_cell[3];
  int dim[3];
  int ncells;
  int nrows;  // dim[1] * dim[2]: rows of cells along x
};

// cell index along one axis, clamped to the grid: monotonic in p, so "every atom with
// coordinate >= a sits in a cell >= cell(a)" holds for atoms inside and outside the grid alike
__device__ __forceinline__ int sort_cell(double p, double origin, double inv_cell, int dim)
{
  double const c = floor((p - origin) * inv_cell);
  return (int)fmin(fmax(c, 0.0), (double)(dim - 1));
}

// one sorted group
struct SortedGroup {
  int n = 0;
  int *d_key = nullptr;
  int *d_sorted = nullptr;      // d_sorted[p] = atom index at sorted position p
  int *d_cell_start = nullptr;  // [2^cell_bits + 2]
  KeySort sort;
  double *d_xs = nullptr, *d_ys = nullptr, *d_zs = nullptr;  // coordinates in sorted order
  // gradient contributions the rows of the OTHER atoms make to this group's atoms (rows.cuh):
  // 3 planes of gs_stride in sorted order, unscaled (sum of gq * d), zeroed by the per-step gather
  double *d_gs = nullptr;
  size_t gs_stride = 0;
};

struct CellSort {
  SortGrid *d_grid = nullptr;
  unsigned long long *d_box = nullptr;  // bounding box of group2 being reduced (ordered bits)
  SortedGroup side[2];  // [0] = group1 (unused for selfCoordNum), [1] = group2 / the self group
  int cell_bits = SORT_MAX_CELL_BITS;  // the grid is kept below 2^cell_bits cells (sort key width)
  // counters of the row list (device, 8 words).  Zeroed before every evaluation: [0] overflow
  // chunks written by this build, [3] row entries of this build, [4..5] listed pairs in rows of
  // this build (u64).  Kept: [6] overflow chunks of the current list (written by the final sum
  // of a rebuild).
  unsigned *d_counts = nullptr;
  bool self = false;
};

// allocate for n1 group1 atoms (0: selfCoordNum) and n2 group2 atoms; not capturable
int cellsort_setup(CellSort &B, int n1, int n2);
void cellsort_free(CellSort &B);
// grid of the group2 atoms, sort of both groups, cell offsets.  x1.. may be NULL (selfCoordNum).
// gate (may be NULL): captured pair-list graphs, work only when *gate == 1.  Launches are counted
// by note_launch().
int cellsort_build(CellSort &B, const double *x1, const double *y1, const double *z1,
                   const double *x2, const double *y2, const double *z2, const int *gate,
                   cudaStream_t stream);
// coordinates of both groups into sorted order (every step)
int cellsort_gather(CellSort &B, const double *x1, const double *y1, const double *z1,
                    const double *x2, const double *y2, const double *z2, const int *gate,
                    int gate_on, cudaStream_t stream);

}  // namespace b200cv
