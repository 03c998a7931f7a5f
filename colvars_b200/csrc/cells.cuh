// cells.cuh -- cell-list rebuild of the pair list (open systems, orthorhombic and triclinic cells).
//
// On a rebuild step the reference visits every pair (src/colvarcomp_coordnums.cpp:219-232), but a
// pair with l2 > tolerance_l2_max returns 0 before anything is evaluated and is not listed
// (src/colvarcomp_coordnums.h:256-261).  With cells at least one cut-off radius wide per axis
// (|d_axis| > cell_axis  =>  (d_axis / r0_axis)^2 > l2_max  =>  l2 > l2_max) every pair outside
// the 27 neighbouring cells is such a pair, so it can be skipped without changing any number:
// value, gradients and list are those of the full sweep.  Candidates next to a decision threshold
// (and ambiguous images in a triclinic cell) are evaluated in the reference's operation order
// (pair_exact), so every listed / not-listed decision is the reference's own; the others use the
// sweep's fast form.
#pragma once
#include <cuda_runtime.h>

#include "pair_math.cuh"
#include "keysort.cuh"

namespace b200cv {

constexpr int CELL_MAX_DIM = 128;  // cells per axis at most (2^21 cells)
constexpr int CELL_MAX = CELL_MAX_DIM * CELL_MAX_DIM * CELL_MAX_DIM;
constexpr int CELL_BLOCKS = 592;
constexpr int CELL_THREADS = 256;

// grid of the group2 atoms, computed on the device (no host round trip)
struct CellGrid {
  double origin[3];
  double inv_cell[3];
  int dim[3];
  int ncells;
  // periodic cell (orthorhombic or triclinic): atoms are binned by their fractional coordinates
  // frac[a] . p wrapped into [0, 1) (frac = reciprocal vectors; diag(1 / edge) when orthorhombic),
  // neighbours wrap around; dim >= 3 per axis so that the three neighbours are distinct
  int periodic;
  double frac[3][3];
};

struct CellBuffers {
  CellGrid *d_grid = nullptr;
  int *d_key = nullptr;         // cell id per group2 atom
  int *d_sorted = nullptr;      // atom index, sorted by cell (keysort.cuh)
  int *d_cell_start = nullptr;  // [CELL_MAX + 2]
  KeySort sort;
  double *d_xs = nullptr, *d_ys = nullptr, *d_zs = nullptr;  // group2 coordinates in cell order
  double *d_gs = nullptr;  // 3 planes of n2: group2 gradients in cell order (list walk)
  int n2 = 0;
};

struct CellPairArgs {
  const double *x1, *y1, *z1;
  const double *x2, *y2, *z2;
  double *g1x, *g1y, *g1z;
  double *g2x, *g2y, *g2z;
  int n2, self;
  int i_begin, i_end;  // group1 rows of this rank
  const int *sorted2;
  const double *xs, *ys, *zs;  // group2 coordinates in cell order: xs[p] = x2[sorted2[p]]
  const int *cell_start;
  const CellGrid *grid;
  unsigned long long *pl;
  unsigned long long pl_cap;
  unsigned long long *pl_count;
  double *partials;  // [CELL_BLOCKS]
  int flags;
  int fast_exp;  // 3: candidates may use the sweep's fast 6/12 form (see ExactArgs::fast_exp)
  // list entries in cell order: (row << 32 | p) with p the sorted position of the group2 atom
  // (and row the sorted position of the group1 atom for selfCoordNum), see launch_cell_gather
  int sorted_entries;
  const int *gate;  // captured pair-list graphs: work only when *gate == 1 (see SweepArgs::gate)
  PairParams P;
};

// allocate for n2 group2 atoms (not capturable: call before any stream capture)
int cells_setup(CellBuffers &B, int n2);
void cells_free(CellBuffers &B);
// grid + sort of the group2 atoms; cut[d] = cut-off radius along axis d
// gate (may be NULL): see CellPairArgs::gate.
// periodic_frac: NULL for a non-periodic system (grid = bounding box of group2, found on the
// device), else the reciprocal vectors of the cell, row by row (diag(1 / edge) for an
// orthorhombic cell); the grid is then fixed by the host.  cells_periodic_dims tells whether the
// cell is large enough (>= 3 cells per axis): a fractional cell must be at least one cut-off wide
// measured perpendicular to its faces, width_a = 1 / |frac_a| -- per-axis cut-offs for an
// orthorhombic cell, the largest of the three otherwise (|d| <= max cut for every listed pair,
// so |frac_a . d| <= 1 / dim_a: the two atoms sit in the same or in adjacent cells whatever
// image of the pair the reference ends up using).
int cells_build(CellBuffers &B, const double *x2, const double *y2, const double *z2, int n2,
                const double cut[3], const double *periodic_frac, const int *gate,
                cudaStream_t stream);
bool cells_periodic_dims(const double cut[3], const double frac[9], int dim[3]);
// A list with sorted_entries is walked in cell order: the neighbours of a row are a few
// contiguous spans there, so the walk's coordinate loads and gradient atomics coalesce instead
// of scattering over the group.  Per walk step: refresh B.d_xs/ys/zs from the current positions
// (the order is that of the last rebuild), zero B.d_gs, walk, add B.d_gs back to the group's
// gradient planes.  gate / gate_on as in SweepArgs (gate NULL: always).
int launch_cell_gather(CellBuffers &B, const double *x2, const double *y2, const double *z2,
                       int n2, const int *gate, int gate_on, cudaStream_t stream);
int launch_cell_scatter_add(CellBuffers &B, double *g2x, double *g2y, double *g2z, int n2,
                            const int *gate, int gate_on, cudaStream_t stream);
int launch_cell_pairs(const CellPairArgs &args, cudaStream_t stream);

}  // namespace b200cv
