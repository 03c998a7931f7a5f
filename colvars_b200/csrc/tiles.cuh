// tiles.cuh -- the pair list as masked cluster tiles (open systems; the sparse-list form).
//
// The reference keeps one bool per pair (src/colvarcomp_coordnums.cpp:147-151) and, between
// rebuilds, evaluates exactly the pairs whose bool is set (:219-228).  Here the same set is kept
// as 32 x 32 TILES: the atoms of both groups are sorted along a cell grid (cells about as large
// as the volume 32 atoms take up, a fine x-coordinate inside the cell as the minor key), 32
// consecutive atoms of group1 in that order form an i-CLUSTER (clusters never straddle a row of
// cells), and a tile pairs one i-cluster with 32 group2 atoms that lie within one cut-off radius
// of the cluster's bounding box.  A tile stores the 32 sorted positions of its group2 atoms and
// one 32-bit mask per i-atom: bit t of lane l's mask is the reference's bool of the pair
// (i-atom l, tile slot (l + t) mod 32).  Tiles of one cluster are grouped into UNITS of up to
// TILE_UNIT tiles, the work item of the walk.
//
// rebuild (tile_build_kernel, one warp per i-cluster): candidates from the neighbouring rows of
//   cells are filtered by their distance to the cluster's bounding box -- a pair beyond it has
//   l2 > tolerance_l2_max, for which the reference returns 0 before evaluating anything and
//   lists nothing (src/colvarcomp_coordnums.h:256-261) -- compacted 32 at a time, and each tile
//   is evaluated with the sweep's rotation scheme (lane l meets slot (l + t) mod 32 at step t,
//   its group2 gradient accumulators travel with the slot): value, gradients and the masks in
//   one pass.
// walk (tile_walk_kernel, one warp per unit): the same tile body under the stored masks:
//   coalesced loads, register / shuffle accumulation, three RED.ADD.F64 per lane and tile.
// Pairs next to a decision threshold (pair_is_rare, pair_math.cuh) never get a mask bit: they go
// to the exact-pair queue in sorted positions and exact_kernel evaluates them in the reference's
// operation order; on a rebuild it appends those it lists to the EXTRAS list (the packed
// (i << 32 | j) list of the other forms), which every walk step evaluates next to the tiles.
// The listed set is: mask bits + extras; b200cv_pairlist_export maps it back to canonical indices.
#pragma once
#include <cuda_runtime.h>

#include "pair_math.cuh"

namespace b200cv {

#ifndef B200CV_TILE_UNIT
#define B200CV_TILE_UNIT 4
#endif
#ifndef B200CV_TILE_MINB
#define B200CV_TILE_MINB 6          // CTAs per SM the walk's register allocation targets
#endif
#ifndef B200CV_TILE_UNROLL
#define B200CV_TILE_UNROLL 2        // rotation steps interleaved per loop trip
#endif
constexpr int TILE_UNIT = B200CV_TILE_UNIT;  // tiles per unit
constexpr int TILE_MAX_DIM = 128;   // cells per axis at most
constexpr int TILE_MAX_CELLS = TILE_MAX_DIM * TILE_MAX_DIM * TILE_MAX_DIM;
constexpr int TILE_THREADS = 128;   // 4 warps per CTA
constexpr int TILE_BLOCKS = 148 * 8;
constexpr int TILE_FINE_BITS = 8;   // fine x-coordinate inside a cell: minor sort key
constexpr int TILE_KEY_BITS = 21 + TILE_FINE_BITS;

struct TileGrid {
  double origin[3];
  double inv_cell[3];
  int dim[3];
  int ncells;
  int nrows;  // dim[1] * dim[2]: rows of cells along x
};

// cell index along one axis, clamped to the grid: monotonic in p, so "every atom with
// coordinate >= a sits in a cell >= cell(a)" holds for atoms inside and outside the grid alike
__device__ __forceinline__ int tile_cell(double p, double origin, double inv_cell, int dim)
{
  double const c = floor((p - origin) * inv_cell);
  return (int)fmin(fmax(c, 0.0), (double)(dim - 1));
}

// one sorted group
struct TileSide {
  int n = 0;
  int *d_key = nullptr, *d_key_sorted = nullptr;
  int *d_val = nullptr, *d_sorted = nullptr;  // d_sorted[p] = atom index at sorted position p
  int *d_cell_start = nullptr;                // [ncells + 1]
  double *d_xs = nullptr, *d_ys = nullptr, *d_zs = nullptr;  // coordinates in sorted order
  double *d_gs = nullptr;                     // gradients in sorted order, 3 planes of n
  int4 *d_clusters = nullptr;                 // {p0, count, row of cells, 0}
  int *d_ncl = nullptr;                       // number of clusters (device)
  int max_clusters = 0;
};

struct TileBuffers {
  TileGrid *d_grid = nullptr;
  TileSide side[2];  // [0] = group1 (unused for selfCoordNum), [1] = group2 / the self group
  void *d_temp = nullptr;
  size_t temp_bytes = 0;
  // the list
  int4 *d_unit_hdr = nullptr;     // {p0, count, ntiles, 0} per unit
  int *d_tile_j = nullptr;        // [cap_units * TILE_UNIT * 32] sorted positions (-1: empty slot)
  unsigned *d_tile_m = nullptr;   // [cap_units * TILE_UNIT * 32] masks
  unsigned cap_units = 0;
  // counters (device, 8 words).  Zeroed before every evaluation: [0] units written by this
  // build, [1] cluster ticket (build), [2] unit ticket (walk), [4..5] listed pairs of this build
  // (u64).  Kept: [6] units of the current list (written by the final sum of a rebuild).
  unsigned *d_counts = nullptr;
  bool self = false;
};

struct TileBuildArgs {
  const double *xi, *yi, *zi;  // i side, sorted order
  double *gi;                  // 3 planes of ni
  int ni;
  const int4 *clusters;
  const int *ncl;
  int rank, world;             // this rank builds clusters [ncl*rank/world, ncl*(rank+1)/world)
  const double *xj, *yj, *zj;  // j side, sorted order
  double *gj;
  int nj;
  const int *cell_start;
  const TileGrid *grid;
  int self;
  int4 *unit_hdr;
  int *tile_j;
  unsigned *tile_m;
  unsigned cap_units;
  unsigned *counts;            // TileBuffers::d_counts
  unsigned long long *rare;    // exact-pair queue, entries in sorted positions
  unsigned int rare_cap;
  unsigned int *rare_count;
  double *partials;            // [TILE_BLOCKS]
  int want_grad;
  double cut[3];               // cut-off radius per axis: beyond it l2 > tolerance_l2_max
  const int *gate;             // captured pair-list graphs: work only when *gate == 1
  PairParams P;
};

struct TileWalkArgs {
  const double *xi, *yi, *zi;
  double *gi;
  int ni;
  const double *xj, *yj, *zj;
  double *gj;
  int nj;
  const int4 *unit_hdr;
  const int *tile_j;
  const unsigned *tile_m;
  unsigned cap_units;
  unsigned *counts;
  unsigned long long *rare;
  unsigned int rare_cap;
  unsigned int *rare_count;
  double *partials;            // [TILE_BLOCKS]
  int want_grad;
  const int *gate;             // work only when *gate == 0
  PairParams P;
};

// allocate for n1 group1 atoms (0: selfCoordNum) and n2 group2 atoms; not capturable
int tiles_setup(TileBuffers &B, int n1, int n2);
void tiles_free(TileBuffers &B);
// (re)allocate the tile arrays for `units` units; not capturable
int tiles_reserve(TileBuffers &B, unsigned units);
// grid of the group2 atoms (bounding box, found on the device), sort of both groups, cluster
// table of the i side.  x1.. may be NULL (selfCoordNum).  gate: see TileBuildArgs::gate (the
// radix sorts themselves always run).  Launches are counted by note_launch().
int tiles_sort(TileBuffers &B, const double *x1, const double *y1, const double *z1,
               const double *x2, const double *y2, const double *z2, const int *gate,
               cudaStream_t stream);
// coordinates of both groups into sorted order, sorted gradient planes zeroed (every step)
int tiles_gather(TileBuffers &B, const double *x1, const double *y1, const double *z1,
                 const double *x2, const double *y2, const double *z2, const int *gate,
                 int gate_on, cudaStream_t stream);
// sorted gradient planes added back to the groups' planes (plane stride s1 / s2)
int tiles_scatter_add(TileBuffers &B, double *g1, size_t s1, double *g2, size_t s2,
                      const int *gate, int gate_on, cudaStream_t stream);
int launch_tile_build(const TileBuildArgs &args, int exp_special, bool aniso, cudaStream_t stream);
int launch_tile_walk(const TileWalkArgs &args, int exp_special, bool aniso, cudaStream_t stream);

}  // namespace b200cv
