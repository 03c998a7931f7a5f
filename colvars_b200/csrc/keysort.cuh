// keysort.cuh -- atoms into cell order without a library sort: a counting sort by cell and a rank
// sort inside the cell, every kernel behind the pair-list gate.
//
// Both cell lists (cells.cuh: periodic cells; cellsort.cuh: the grid of an open system) order
// their atoms by an integer key = cell << fine_bits | fine.  A library radix sort does that in
// five kernels that a captured graph cannot switch off: the list steps of a captured evaluation
// (gate word != 1) would sort again on every launch.  Here
//   count     cnt[cell]++ per atom (by the caller's key kernel: keysort_count)
//   scan      cell_start[c] = atoms in cells < c, c in [0, ncells_max + 1]: tiles of 4096 cells, then
//             the tile offsets
//   scatter   atoms into the slots of their cell, in arrival order (cnt counts down to zero again)
//   rank      inside the cell: position = number of atoms of the cell with a smaller (key, index)
// The result is the order of a stable sort by key -- ascending key, ties by atom index -- so it
// does not depend on the arrival order of the scatter.  Cells hold a few dozen atoms by
// construction of the grids; the rank sort is quadratic in the population of a cell.
#pragma once
#include <cuda_runtime.h>

namespace b200cv {

struct KeySort {
  int *d_cnt = nullptr;       // [ncells_max + 2], zero between sorts
  int *d_tile = nullptr;      // totals of the scan tiles
  int *d_tmp_key = nullptr;   // [n] keys in arrival order inside the cells
  int *d_tmp_val = nullptr;   // [n] atom indices in arrival order inside the cells
  int n = 0, ncells_max = 0;
};

int keysort_setup(KeySort &S, int n, int ncells_max);
void keysort_free(KeySort &S);
// the count step, for the kernel that computes the keys
__device__ __forceinline__ void keysort_count(int *cnt, int cell) { atomicAdd(&cnt[cell], 1); }

// key[i] of atom i (cell = key >> fine_bits < ncells_max), counted into S.d_cnt by the caller's
// key kernel -> sorted[p] = atom at position p,
// cell_start[c] for c in [0, ncells_max + 1] (= n from the first empty cell behind the last atom
// on).  gate (may be NULL): work only when *gate == 1.  Three kernels (four beyond 4096 cells).
// kernels keysort_run launches
int keysort_launches(const KeySort &S);
int keysort_run(KeySort &S, const int *key, int fine_bits, int *sorted, int *cell_start,
                const int *gate, cudaStream_t stream);

}  // namespace b200cv
