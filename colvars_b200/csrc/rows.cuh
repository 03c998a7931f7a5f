// rows.cuh -- the pair list as neighbour ROWS (open systems; the sparse-list form).
//
// The reference keeps one bool per pair (src/colvarcomp_coordnums.cpp:147-151) and, between
// rebuilds, evaluates exactly the pairs whose bool is set (:219-228).  Here the listed pairs are
// kept as rows: the atoms of group1 (selfCoordNum: of the group) each own the sorted positions
// (cell order, cellsort.cuh) of their listed partners in group2 (selfCoordNum: of the partners
// BEHIND them in cell order), so every listed pair is one entry of one row.  A warp works on one
// row, 32 partners at a time, with coalesced index loads, partner coordinates that sit next to
// each other in cell order and no mask (a row is padded to whole steps with the index of a
// far-away point, cellsort.cuh).
//
//   value      F of every entry.
//   gradient   own atom: -c_grad * sum gq * (x_partner - x_own), a register sum reduced once per
//              row; partner: + the same term, added (unscaled) to a gradient buffer in cell order
//              (SortedGroup::d_gs, three reductions per entry that fall into a few sectors),
//              which exact_kernel adds to the group's gradient afterwards
//              (the reference's -G for group1 and +G for group2, src/colvarcomp_coordnums.h:267-277).
//              [The first form of the rows kept every pair in the rows of BOTH its atoms -- no
//              reductions in the pair loop, twice the entries, group2 rows as well; measured:
//              the three reductions per pair cost less than the second evaluation, and the
//              rebuild searches half the cells.]
//   rare pairs next to a decision threshold (pair_is_rare, pair_math.cuh) contribute nothing in a
//              row; they are queued (atom indices) for exact_kernel, which evaluates them in the
//              reference's operation order and updates both atoms; they are never row entries.
//              |l2 - 1| < 2^-4 pairs are listed for certain (F ~ 1/2), pairs inside the window
//              around tolerance_l2_max are listed by exact_kernel's own decision.  What
//              exact_kernel lists on a rebuild becomes the EXTRAS (packed (i << 32 | j) entries,
//              atom indices), evaluated by every list step next to the rows.
//   rebuild    row_build_kernel, one warp per atom: partners from the neighbouring rows of cells
//              (x-range per row of cells from the distance of the atom to that row;
//              selfCoordNum: only what lies behind the atom in cell order), first a cheap pass
//              l2 <= tolerance_l2_max (everything else returns 0 in the reference before
//              anything is evaluated and is not listed, src/colvarcomp_coordnums.h:256-261), then
//              the survivors 32 at a time at full lane occupancy: value, gradients, row.
//   list step  row_walk_kernel: one wave of CTAs, a warp per block of consecutive rows (rows.cu).
// The listed set is: row entries + extras; b200cv_pairlist_export maps it back to canonical
// indices.
#pragma once
#include <cuda_runtime.h>

#include "pair_math.cuh"
#include "cellsort.cuh"

namespace b200cv {

constexpr int ROW_THREADS = 128;   // 4 warps per CTA
constexpr int ROW_BLOCKS = 148 * 8;
#ifndef B200CV_ROW_EXACT_BLOCKS
#define B200CV_ROW_EXACT_BLOCKS 592
#endif
// CTAs of the exact kernel behind the walk: the extras (the pairs next to a threshold when the
// list was built, most of them still there) take the reference-order path, a long dependent
// chain per pair -- one pair per thread
constexpr int ROW_EXACT_BLOCKS = B200CV_ROW_EXACT_BLOCKS;
constexpr int ROW_BUF = 512;       // row entries buffered per warp before they are written out

struct RowList {
  int *d_idx = nullptr;        // partner positions (cell order of the partner's group)
  unsigned cap_idx = 0;
  // {own position, start, length, 0}: one slot per row and behind them the overflow chunks of
  // rows with more than ROW_BUF entries
  int4 *d_chunks = nullptr;
  unsigned cap_chunks = 0;
  // counters live in CellSort::d_counts: [0] overflow chunks of this build, [3] entries of this
  // build, [4..5] listed pairs in rows of this build, [6] overflow chunks
  // of the current list
};

struct RowSideView {
  const double *xs, *ys, *zs;  // coordinates in cell order
  const int *sorted;           // atom index at a sorted position
  const int *cell_start;
  int n;
  double *grad;                // the group's gradient planes (atom order)
  size_t gstride;
  double *gs;                  // partner shares in cell order (SortedGroup::d_gs), 3 planes
  size_t gs_stride;
};

struct RowBuildArgs {
  RowSideView own, nbr;
  const SortGrid *grid;
  int self;                    // own == nbr: partners behind the own position only
  int row_begin, row_end;      // rows (sorted positions of the own group) of this rank
  int *idx;
  unsigned cap_idx;
  int4 *chunks;
  unsigned cap_chunks;
  unsigned chunk_base;         // slot of row 0
  unsigned chunk_extra;        // first overflow slot (= number of rows)
  unsigned *counts;            // CellSort::d_counts
  unsigned long long *rare;
  unsigned int rare_cap;
  unsigned int *rare_count;
  double *partials;            // [ROW_BLOCKS]
  int want_grad;
  double cut[3];               // cut-off radius per axis (with its safety margin)
  const int *gate;             // captured pair-list graphs: work only when *gate == 1
  PairParams P;
};

struct RowWalkArgs {
  RowSideView own, nbr;        // group1 and group2 (selfCoordNum: the group, twice)
  int self;
  const int *idx;
  const int4 *chunks;
  unsigned cap_chunks;
  int row_begin, row_end;      // this rank's rows
  unsigned chunk_extra;        // first overflow slot
  unsigned *counts;
  unsigned long long *rare;
  unsigned int rare_cap;
  unsigned int *rare_count;
  double *partials;            // [ROW_BLOCKS]
  int want_grad;
  const int *gate;             // work only when *gate == 0
  PairParams P;
};

int rows_reserve(RowList &R, unsigned entries, unsigned chunks);
void rows_free(RowList &R);
int launch_row_build(const RowBuildArgs &args, int exp_special, bool aniso, cudaStream_t stream);
int launch_row_walk(const RowWalkArgs &args, int exp_special, bool aniso, cudaStream_t stream);

}  // namespace b200cv
