// calc.cu -- one evaluation: every kernel of coordnum::calc_value
// (src/colvarcomp_coordnums.cpp:251-313) enqueued on a stream, and the pair-list state machine
// around it (which form of rebuild / walk runs, queue capacities, commit and overflow redo after
// the stream has drained).
#include "engine.h"

using namespace b200cv;

namespace b200cv {

bool groups_ready(b200cv_cvc *h)
{
  return h->g[0].d_index && (h->self() || h->g[1].d_index);
}

// calc_required_properties (src/colvaratoms.cpp:1325-1352): COM and COG of a group as read
// (and fitted).  The pair sweep never reads them, so they are computed with the step only
// for a group that enters as its centre (group{1,2}CenterOnly, groupCoord) and otherwise on
// demand (b200cv_get_centers): two launch-bound reductions less per step.
bool centers_in_step(const b200cv_cvc *h, const Group &G)
{
  return (&G == &h->g[0]) ? h->center1() : h->center2();
}

int centers_now(b200cv_cvc *h, Group &G, cudaStream_t s)
{
  CUDA_OK(h, launch_com_cog(G.d_pos, G.stride, G.d_mass, G.total_mass, G.n, G.com(), G.cog(),
                            G.d_com_scratch, G.d_com_ticket, s));
  return B200CV_OK;
}

int centers_after_read(b200cv_cvc *h, Group &G, cudaStream_t s)
{
  if (!centers_in_step(h, G)) return B200CV_OK;
  return centers_now(h, G, s);
}

// lists built by the cell kernel are kept, and walked, in cell order (cells.cuh)
bool sorted_walk_enabled() { return std::getenv("B200CV_NO_SORTED_WALK") == nullptr; }

// the last rebuild listed fewer than 2 % of this rank's pairs
bool sparse_list(const b200cv_cvc *h)
{
  unsigned long long const n1 = h->g[0].n, n2 = h->self() ? n1 : h->g[1].n;
  unsigned long long const pairs = (h->self() ? n1 * (n1 - 1) / 2 : n1 * n2) / (unsigned)h->world;
  return h->pl_n * 50ull < pairs;
}

// cut-off radius per axis: beyond it l2 > tolerance_l2_max whatever the other two components are
void cell_cutoffs(const b200cv_cvc *h, double cut[3])
{
  for (int d = 0; d < 3; d++) cut[d] = std::sqrt(h->P.l2_max) / h->P.inv_r0[d];
}

// reciprocal vectors for the cell list's fractional binning: diag(1 / |edge|) for an orthorhombic
// cell, the reference's reciprocal cell vectors for a triclinic one; false for any other kind
bool cell_list_frac(const b200cv_cvc *h, double frac[9])
{
  for (int k = 0; k < 9; k++) frac[k] = 0.0;
  if (h->P.bc_type == 2) {
    for (int d = 0; d < 3; d++) frac[4 * d] = 1.0 / std::fabs(h->P.cell[d][d]);
    return true;
  }
  if (h->P.bc_type == 3) {
    for (int a = 0; a < 3; a++) {
      for (int d = 0; d < 3; d++) frac[3 * a + d] = h->P.recip[a][d];
    }
    return true;
  }
  return false;
}

// the cell list (cells.cuh) is set up and applies: a list has been built before, and the system
// is non-periodic, orthorhombic or triclinic with at least three cells per axis (mixed cells keep
// the all-pairs sweep)
bool cells_usable(const b200cv_cvc *h)
{
  if (!(h->cells_ready && h->d_pl && h->pl_n > 0) || std::getenv("B200CV_NO_CELLS")) {
    return false;
  }
  if (h->P.bc_type == 0) return true;
  double cut[3], frac[9];
  int dim[3];
  if (!cell_list_frac(h, frac)) return false;
  cell_cutoffs(h, cut);
  return cells_periodic_dims(cut, frac, dim);
}

// rebuild through the cell list instead of the all-pairs sweep?
bool rebuild_with_cells(const b200cv_cvc *h, int flags)
{
  return (flags & B200CV_EF_REBUILD_PAIRLIST) && (flags & B200CV_EF_USE_PAIRLIST) &&
         cells_usable(h) && h->pl_valid && !(h->gated) && sparse_list(h);
}

// the row form of the list (rows.cuh) applies: an open system whose list a previous rebuild
// has shown to be sparse (B200CV_NO_ROWS=1 keeps the cell-ordered entries of cells.cuh there)
bool rows_usable(const b200cv_cvc *h)
{
  return h->sort_ready && h->d_pl && h->pl_n > 0 && h->P.bc_type == 0 &&
         !(h->center1() || h->center2()) && std::getenv("B200CV_NO_ROWS") == nullptr &&
         std::getenv("B200CV_NO_CELLS") == nullptr;
}

bool rebuild_with_rows(const b200cv_cvc *h, int flags)
{
  return (flags & B200CV_EF_REBUILD_PAIRLIST) && (flags & B200CV_EF_USE_PAIRLIST) &&
         rows_usable(h) && h->pl_valid && !(h->gated) && sparse_list(h);
}

// chunk slots taken by the rows themselves (one per atom of group1); overflow chunks follow
unsigned row_slots(const b200cv_cvc *h)
{
  return (unsigned)h->g[0].n;
}

// entries / chunks the rows of the next rebuild may need: one entry per listed pair, one chunk
// per atom of group1 plus one per ROW_BUF entries; a guess that turns out too small costs one redo
// of the step
static void row_estimate(const b200cv_cvc *h, unsigned &entries, unsigned &chunks)
{
  unsigned long long const atoms = (unsigned long long)h->g[0].n;
  unsigned long long const e = h->pl_n + h->pl_n / 4ull + 32ull * atoms + 65536ull;  // chunks are padded to 32
  unsigned long long const c = atoms + e / (unsigned long long)ROW_BUF + 4096ull;
  entries = (unsigned)std::min<unsigned long long>(e, 0xfffffff0ull);
  chunks = (unsigned)std::min<unsigned long long>(c, 0xfffffff0ull);
}

int prepare_calc(b200cv_cvc *h, int flags)
{
  if (!(flags & B200CV_EF_USE_PAIRLIST)) return B200CV_OK;
  bool const rows_next = h->gated ? (rows_usable(h) && sparse_list(h)) : rebuild_with_rows(h, flags);
  if (rows_next && h->rows.cap_idx == 0) {
    unsigned entries, chunks;
    row_estimate(h, entries, chunks);
    h->epoch++;
    CUDA_OK(h, rows_reserve(h->rows, entries, chunks));
  }
  return B200CV_OK;
}

int ensure_pairlist_capacity(b200cv_cvc *h, unsigned long long need)
{
  if (need <= h->pl_cap && h->d_pl) return B200CV_OK;
  free_dev(h->d_pl);
  h->epoch++;
  unsigned long long cap = std::max<unsigned long long>(need + need / 2, 1ull << 20);
  CUDA_OK(h, cudaMalloc(&h->d_pl, sizeof(unsigned long long) * cap));
  h->pl_cap = cap;
  return B200CV_OK;
}

int ensure_rare_capacity(b200cv_cvc *h, unsigned long long need)
{
  if (need <= h->rare_cap && h->d_rare) return B200CV_OK;
  free_dev(h->d_rare);
  h->epoch++;
  unsigned long long cap = std::max<unsigned long long>(need + need / 2, 1ull << 20);
  if (cap > 0xffffffffull) return fail(h, B200CV_MEMORY_ERROR, "exact-pair queue too large");
  CUDA_OK(h, cudaMalloc(&h->d_rare, sizeof(unsigned long long) * cap));
  h->rare_cap = (unsigned int)cap;
  return B200CV_OK;
}

// ---- one evaluation (all kernels of coordnum::calc_value) on `stream` ----------------------
int enqueue_calc(b200cv_cvc *h, int flags, cudaStream_t stream)
{
  Group &G1 = h->g[0];
  Group &G2 = h->grp2();
  bool const grad = flags & B200CV_EF_GRADIENTS;
  bool const use_pl = flags & B200CV_EF_USE_PAIRLIST;
  bool const rebuild = flags & B200CV_EF_REBUILD_PAIRLIST;
  h->launches = 0;
  int npart = 0;
  bool const gated = h->gated && use_pl;
  // captured evaluation without a pair list: one form only, but the host may still have to skip
  // a launch (mode 2, b200cv_async_set_skip); its kernels run in mode 0
  bool const skip_gated = h->gated && !use_pl;

  CUDA_OK(h, cudaMemsetAsync(h->d_counters, 0, 16, stream));
  // this evaluation builds or walks the row form of the list (rows.cuh)
  bool const rows_gated = gated && rows_usable(h) && sparse_list(h) && h->rows.cap_idx > 0;
  bool const rows_eval =
      use_pl && !(h->center1() || h->center2()) &&
      (rows_gated || (!gated && ((!rebuild && h->pl_valid && h->pl_rows) ||
                                 (rebuild_with_rows(h, flags) && h->rows.cap_idx > 0))));
  if (rows_eval) CUDA_OK(h, cudaMemsetAsync(h->sort.d_counts, 0, 6 * sizeof(unsigned), stream));
  if (gated || skip_gated) {
    CUDA_OK(h, launch_fetch_gate(h->h_mode_dev, h->gate(), gated ? 1 : 0, stream));
    h->launches++;
  }
  // the final sum: a launch of its own, or the last CTA of the exact kernel (`finalized`)
  bool finalized = false;
  auto finalize_args = [&](int nparts) {
    FinalizeArgs F{};
    F.partials = h->d_partials;
    F.n = nparts;
    F.d_value = h->d_value;
    F.h_value = (h->world > 1 || h->dinv()) ? nullptr : h->h_value_dev;
    F.rare_count = h->rare_count();
    F.pl_count = h->pl_count();
    F.h_counters = h->h_counters_dev;
    F.gate = (gated || skip_gated) ? h->gate() : nullptr;
    F.rebuilt = (!gated && use_pl && rebuild && !(h->center1() || h->center2())) ? 1 : 0;
    F.pl_persist = h->pl_count() + 1;
    F.pl_cap = h->pl_cap;
    F.row_counts = rows_eval ? h->sort.d_counts : nullptr;
    F.row_cap_chunks = h->rows.cap_chunks > row_slots(h) ? h->rows.cap_chunks - row_slots(h) : 0u;
    F.row_cap_idx = h->rows.cap_idx;
    F.rare_cap = h->rare_cap;
    F.overflow = (h->world > 1) ? h->d_reduce + 1 : nullptr;
    return F;
  };

  if (h->center1() || h->center2()) {
    // COM variants (src/colvarcomp_coordnums.cpp:190-247): O(N) work, reference-order pairs
    int rc = ensure_partials(h, PARTIALS_TAIL);
    if (rc) return rc;
    CenterArgs A{};
    A.P = h->P;
    A.flags = flags;
    A.partials = h->d_partials;
    A.pairlist = use_pl ? h->d_pl_center : nullptr;
    A.gate = (gated || skip_gated) ? h->gate() : nullptr;
    CUDA_OK(h, cudaMemsetAsync(G1.com_grad(), 0, 3 * sizeof(double), stream));
    CUDA_OK(h, cudaMemsetAsync(G2.com_grad(), 0, 3 * sizeof(double), stream));
    if (h->center1() && h->center2()) {
      A.center = G1.com();
      A.center2 = G2.com();
      A.center_grad = G1.com_grad();
      A.center2_grad = G2.com_grad();
      A.n = 0;
      A.center_is_group1 = 1;
      npart = 1;
    } else if (h->center1()) {
      A.center = G1.com();
      A.center_grad = G1.com_grad();
      A.x = G2.d_pos;
      A.y = G2.d_pos + G2.stride;
      A.z = G2.d_pos + 2 * G2.stride;
      A.gx = G2.d_grad;
      A.gy = G2.d_grad + G2.stride;
      A.gz = G2.d_grad + 2 * G2.stride;
      A.n = G2.n;
      A.center_is_group1 = 1;
      npart = std::min((G2.n + EXACT_THREADS - 1) / EXACT_THREADS, EXACT_BLOCKS);
    } else {
      A.center = G2.com();
      A.center_grad = G2.com_grad();
      A.x = G1.d_pos;
      A.y = G1.d_pos + G1.stride;
      A.z = G1.d_pos + 2 * G1.stride;
      A.gx = G1.d_grad;
      A.gy = G1.d_grad + G1.stride;
      A.gz = G1.d_grad + 2 * G1.stride;
      A.n = G1.n;
      A.center_is_group1 = 0;
      npart = std::min((G1.n + EXACT_THREADS - 1) / EXACT_THREADS, EXACT_BLOCKS);
    }
    // a sharded run evaluates the O(N) centre variants on rank 0 only
    if (h->rank == 0) {
      CUDA_OK(h, launch_center(A, stream));
      h->launches++;
      if (grad) {
        // set_weighted_gradient (src/colvaratoms.cpp:1415-1431): assignment
        if (h->center1()) {
          CUDA_OK(h, launch_weighted_gradient(G1.d_weight, G1.com_grad(), G1.n, G1.d_grad,
                                              G1.stride, stream));
          h->launches++;
        }
        if (h->center2()) {
          CUDA_OK(h, launch_weighted_gradient(G2.d_weight, G2.com_grad(), G2.n, G2.d_grad,
                                              G2.stride, stream));
          h->launches++;
        }
      }
    } else {
      npart = 0;
    }
  } else {
    if (h->items_dirty) {
      int rc = build_items(h);
      if (rc) return rc;
    }
    double *x1 = G1.d_pos, *y1 = G1.d_pos + G1.stride, *z1 = G1.d_pos + 2 * G1.stride;
    double *x2 = G2.d_pos, *y2 = G2.d_pos + G2.stride, *z2 = G2.d_pos + 2 * G2.stride;
    double *g1x = G1.d_grad, *g1y = G1.d_grad + G1.stride, *g1z = G1.d_grad + 2 * G1.stride;
    double *g2x = G2.d_grad, *g2y = G2.d_grad + G2.stride, *g2z = G2.d_grad + 2 * G2.stride;
    // mixed / triclinic cells: the tile sweep's general image search; B200CV_DENSE=1 keeps the
    // reference-order dense kernel (one pair per thread) for comparison
    bool const general_cell = (h->P.bc_type == 1 || h->P.bc_type == 3);
    bool const general_bc = general_cell && std::getenv("B200CV_DENSE") != nullptr;
    bool const walk_list = use_pl && !rebuild && h->pl_valid;

    // non-rebuild step: only listed pairs (src/colvarcomp_coordnums.cpp:219-228)
    auto enqueue_walk = [&](double *partials, const int *gate) -> int {
      bool const fuse = (gate == nullptr);
      ExactArgs E{};
      E.x1 = x1; E.y1 = y1; E.z1 = z1; E.x2 = x2; E.y2 = y2; E.z2 = z2;
      E.g1x = g1x; E.g1y = g1y; E.g1z = g1z; E.g2x = g2x; E.g2y = g2y; E.g2z = g2z;
      if (h->pl_sorted) {
        // list in cell order: walk the cell-ordered copy of group2 (positions refreshed in the
        // order of the last rebuild), gradients go back through the same order afterwards
        CellBuffers &B = h->cells;
        CUDA_OK(h, launch_cell_gather(B, x2, y2, z2, G2.n, gate, 0, stream));
        CUDA_OK(h, cudaMemsetAsync(B.d_gs, 0, sizeof(double) * 3 * (size_t)G2.n, stream));
        h->launches++;
        E.x2 = B.d_xs; E.y2 = B.d_ys; E.z2 = B.d_zs;
        E.g2x = B.d_gs; E.g2y = B.d_gs + G2.n; E.g2z = B.d_gs + 2 * (size_t)G2.n;
        if (h->self()) {
          E.x1 = E.x2; E.y1 = E.y2; E.z1 = E.z2;
          E.g1x = E.g2x; E.g1y = E.g2y; E.g1z = E.g2z;
        }
      }
      E.list = h->d_pl;
      E.count64 = h->pl_count() + 1;  // persistent copy of the list length
      E.cap = h->pl_cap;
      E.partials = partials;
      E.flags = flags & ~B200CV_EF_REBUILD_PAIRLIST;
      E.fast_exp = (h->exp_special == 3) ? 3 : 0;  // every kind of cell has a fast form
      E.gate = gate;
      E.gate_on = 0;
      if (fuse) {
        E.fin_ticket = h->fin_ticket();
        E.fin = finalize_args(WALK_BLOCKS);
        finalized = true;
      }
      E.P = h->P;
      CUDA_OK(h, launch_exact(E, stream, WALK_BLOCKS));
      h->launches++;
      if (h->pl_sorted && grad) {
        CUDA_OK(h, launch_cell_scatter_add(h->cells, g2x, g2y, g2z, G2.n, gate, 0, stream));
        h->launches++;
      }
      return B200CV_OK;
    };
    // every pair: tile sweep (+ exact-pair queue) or, for mixed / triclinic cells, the dense
    // kernel; records the pair list when `record`.  Returns the number of partial sums.
    auto enqueue_all_pairs = [&](int sweep_flags, bool record, const int *gate,
                                 int &nparts, int gate_on = 1) -> int {
      if (record) h->pl_sorted = h->pl_rows = false;  // the sweep and the dense kernel list (i, j) as they are
      if (general_bc) {
        DenseArgs D{};
        D.x1 = x1; D.y1 = y1; D.z1 = z1; D.x2 = x2; D.y2 = y2; D.z2 = z2;
        D.g1x = g1x; D.g1y = g1y; D.g1z = g1z; D.g2x = g2x; D.g2y = g2y; D.g2z = g2z;
        D.n1 = G1.n;
        D.n2 = G2.n;
        D.self = h->self() ? 1 : 0;
        D.i_begin = (int)((long long)G1.n * h->rank / h->world);
        D.i_end = (int)((long long)G1.n * (h->rank + 1) / h->world);
        D.partials = h->d_partials;
        D.pl = record ? h->d_pl : nullptr;
        D.pl_cap = h->pl_cap;
        D.pl_count = h->pl_count();
        D.flags = sweep_flags;
        D.gate = gate;
        D.gate_on = gate_on;
        D.P = h->P;
        CUDA_OK(h, launch_dense(D, stream));
        h->launches++;
        nparts = EXACT_BLOCKS;
        return B200CV_OK;
      }
      SweepArgs S{};
      S.x1 = x1; S.y1 = y1; S.z1 = z1; S.x2 = x2; S.y2 = y2; S.z2 = z2;
      S.g1x = g1x; S.g1y = g1y; S.g1z = g1z; S.g2x = g2x; S.g2y = g2y; S.g2z = g2z;
      S.items = h->d_items;
      S.partials = h->d_partials;
      S.rare = h->d_rare;
      S.rare_cap = h->rare_cap;
      S.rare_count = h->rare_count();
      S.pl = h->d_pl;
      S.pl_cap = record ? h->pl_cap : 0;  // all-true list before the first rebuild: no recording
      S.pl_count = h->pl_count();
      S.i_end = G1.n;
      S.want_grad = grad ? 1 : 0;
      S.gate = gate;
      S.gate_on = gate_on;
      S.P = h->P;
      int const pbc = h->P.bc_type == 2 ? BC_ORTHO : (general_cell ? BC_GENERAL : BC_NONE);
      CUDA_OK(h, launch_sweep(S, h->nitems, h->exp_special, h->aniso, pbc, use_pl, stream));
      if (h->nitems > 0) h->launches++;
      // pairs next to a decision threshold: reference-order arithmetic (distanceInv has none:
      // the launch then only zeroes its partial sums)
      ExactArgs E{};
      E.x1 = x1; E.y1 = y1; E.z1 = z1; E.x2 = x2; E.y2 = y2; E.z2 = z2;
      E.g1x = g1x; E.g1y = g1y; E.g1z = g1z; E.g2x = g2x; E.g2y = g2y; E.g2z = g2z;
      E.list = h->d_rare;
      E.count32 = h->rare_count();
      E.cap = h->rare_cap;
      E.partials = h->d_partials + std::max(h->nitems, 0);
      E.pl = record ? h->d_pl : nullptr;
      E.pl_cap = h->pl_cap;
      E.pl_count = h->pl_count();
      E.flags = sweep_flags;
      E.gate = gate;
      E.gate_on = gate_on;
      nparts = h->nitems + EXACT_BLOCKS;
      if (gate == nullptr) {
        E.fin_ticket = h->fin_ticket();
        E.fin = finalize_args(nparts);
        finalized = true;
      }
      E.P = h->P;
      CUDA_OK(h, launch_exact(E, stream));
      h->launches++;
      return B200CV_OK;
    };

    // rebuild of a list that a previous rebuild showed to be sparse: only pairs in neighbouring
    // cells can be within tolerance_l2_max, everything else is exactly zero (cells.cuh)
    auto enqueue_cells = [&](int cell_flags, const int *gate) -> int {
      static_assert(CELL_BLOCKS + WALK_BLOCKS <= PARTIALS_TAIL, "partials are sized for PARTIALS_TAIL");
      double cut[3];
      cell_cutoffs(h, cut);
      double frac[9];
      bool const periodic = cell_list_frac(h, frac);
      CUDA_OK(h, cells_build(h->cells, x2, y2, z2, G2.n, cut, periodic ? frac : nullptr, gate,
                             stream));
      h->launches += 3 + keysort_launches(h->cells.sort);  // grid, keys, sort, gather
      CellPairArgs C{};
      C.x1 = x1; C.y1 = y1; C.z1 = z1; C.x2 = x2; C.y2 = y2; C.z2 = z2;
      C.g1x = g1x; C.g1y = g1y; C.g1z = g1z; C.g2x = g2x; C.g2y = g2y; C.g2z = g2z;
      C.n2 = G2.n;
      C.self = h->self() ? 1 : 0;
      C.i_begin = (int)((long long)G1.n * h->rank / h->world);
      C.i_end = (int)((long long)G1.n * (h->rank + 1) / h->world);
      C.sorted2 = h->cells.d_sorted;
      C.xs = h->cells.d_xs;
      C.ys = h->cells.d_ys;
      C.zs = h->cells.d_zs;
      C.cell_start = h->cells.d_cell_start;
      C.grid = h->cells.d_grid;
      C.pl = h->d_pl;
      C.pl_cap = h->pl_cap;
      C.pl_count = h->pl_count();
      C.partials = h->d_partials;
      C.flags = cell_flags;
      C.fast_exp = (h->exp_special == 3) ? 3 : 0;  // every kind of cell has a fast form
      h->pl_sorted = sorted_walk_enabled();
      h->pl_rows = false;
      C.sorted_entries = h->pl_sorted ? 1 : 0;
      C.gate = gate;
      C.P = h->P;
      CUDA_OK(h, launch_cell_pairs(C, stream));
      h->launches++;
      return B200CV_OK;
    };

    // ---- the list as neighbour rows (rows.cuh): open systems, sparse lists --------------------
    auto row_sides = [&](RowSideView &s1, RowSideView &s2) {
      SortedGroup &J = h->sort.side[1];
      SortedGroup &I = h->self() ? J : h->sort.side[0];
      s1 = RowSideView{I.d_xs, I.d_ys, I.d_zs, I.d_sorted, I.d_cell_start, I.n, G1.d_grad, G1.stride,
                       I.d_gs, I.gs_stride};
      s2 = RowSideView{J.d_xs, J.d_ys, J.d_zs, J.d_sorted, J.d_cell_start, J.n, G2.d_grad, G2.stride,
                       J.d_gs, J.gs_stride};
    };
    // exact-pair queue (and, on list steps, the extras) on the groups' own arrays: the rows queue
    // atom indices.  Partial sums of the exact kernel at `partials`.
    auto enqueue_row_exact = [&](int exact_flags, bool walk, double *partials, int nparts,
                                 const int *gate, int gate_on, int blocks = EXACT_BLOCKS) -> int {
      ExactArgs E{};
      E.x1 = x1; E.y1 = y1; E.z1 = z1; E.x2 = x2; E.y2 = y2; E.z2 = z2;
      E.g1x = g1x; E.g1y = g1y; E.g1z = g1z; E.g2x = g2x; E.g2y = g2y; E.g2z = g2z;
      E.list = h->d_rare;
      E.count32 = h->rare_count();
      E.cap = h->rare_cap;
      {
        // what the rows added to their partners: group2 (selfCoordNum: the group) in cell order
        SortedGroup const &J = h->sort.side[1];
        E.gs = J.d_gs;
        E.gs_stride = J.gs_stride;
        E.gs_sorted = J.d_sorted;
        E.gs_n = J.n;
        E.gs_group = h->self() ? 1 : 2;
      }
      if (walk) {
        E.list2 = h->d_pl;
        E.count2 = h->pl_count() + 1;  // persistent length of the extras
        E.cap2 = h->pl_cap;
        E.fast_exp = (h->exp_special == 3) ? 3 : 0;
      } else {
        E.pl = h->d_pl;
        E.pl_cap = h->pl_cap;
        E.pl_count = h->pl_count();
      }
      E.partials = partials;
      E.flags = exact_flags;
      E.gate = gate;
      E.gate_on = gate_on;
      if (gate == nullptr) {
        E.fin_ticket = h->fin_ticket();
        E.fin = finalize_args(nparts);
        finalized = true;
      }
      E.P = h->P;
      CUDA_OK(h, launch_exact(E, stream, blocks));
      h->launches++;
      return B200CV_OK;
    };
    // rebuild.  Partial sums: [ROW_BLOCKS | EXACT_BLOCKS]
    auto enqueue_rows_build = [&](int build_flags, double *partials, const int *gate,
                                  int &nparts) -> int {
      CellSort &B = h->sort;
      CUDA_OK(h, cellsort_build(B, x1, y1, z1, x2, y2, z2, gate, stream));
      // bounding box + grid; keys and sort per group
      h->launches += 1 + (h->self() ? 1 : 2) * (1 + keysort_launches(B.side[1].sort));
      CUDA_OK(h, cellsort_gather(B, x1, y1, z1, x2, y2, z2, gate, 1, stream));
      h->launches++;
      RowSideView s1, s2;
      row_sides(s1, s2);
      {
        RowBuildArgs T{};
        T.own = s1;
        T.nbr = h->self() ? s1 : s2;
        T.grid = B.d_grid;
        T.self = h->self() ? 1 : 0;
        T.row_begin = (int)((long long)T.own.n * h->rank / h->world);
        T.row_end = (int)((long long)T.own.n * (h->rank + 1) / h->world);
        T.idx = h->rows.d_idx;
        T.cap_idx = h->rows.cap_idx;
        T.chunks = h->rows.d_chunks;
        T.cap_chunks = h->rows.cap_chunks;
        T.chunk_base = 0u;
        T.chunk_extra = row_slots(h);
        T.counts = B.d_counts;
        T.rare = h->d_rare;
        T.rare_cap = h->rare_cap;
        T.rare_count = h->rare_count();
        T.partials = partials;
        T.want_grad = grad ? 1 : 0;
        cell_cutoffs(h, T.cut);
        for (int d = 0; d < 3; d++) T.cut[d] *= (1.0 + 1.0e-9);
        T.gate = gate;
        T.P = h->P;
        CUDA_OK(h, launch_row_build(T, h->exp_special, h->aniso, stream));
        h->launches++;
      }
      h->pl_rows = true;
      h->pl_sorted = false;
      nparts = ROW_BLOCKS + EXACT_BLOCKS;
      return enqueue_row_exact(build_flags, false, partials + ROW_BLOCKS, nparts, gate, 1);
    };
    // list step.  Partial sums: [ROW_BLOCKS | ROW_EXACT_BLOCKS]
    auto enqueue_rows_walk = [&](double *partials, const int *gate) -> int {
      CellSort &B = h->sort;
      CUDA_OK(h, cellsort_gather(B, x1, y1, z1, x2, y2, z2, gate, 0, stream));
      h->launches++;
      RowWalkArgs T{};
      RowSideView s1, s2;
      row_sides(s1, s2);
      T.own = s1;
      T.nbr = h->self() ? s1 : s2;
      T.self = h->self() ? 1 : 0;
      T.idx = h->rows.d_idx;
      T.chunks = h->rows.d_chunks;
      T.cap_chunks = h->rows.cap_chunks;
      T.row_begin = (int)((long long)T.own.n * h->rank / h->world);
      T.row_end = (int)((long long)T.own.n * (h->rank + 1) / h->world);
      T.chunk_extra = row_slots(h);
      T.counts = B.d_counts;
      T.rare = h->d_rare;
      T.rare_cap = h->rare_cap;
      T.rare_count = h->rare_count();
      T.partials = partials;
      T.want_grad = grad ? 1 : 0;
      T.gate = gate;
      T.P = h->P;
      CUDA_OK(h, launch_row_walk(T, h->exp_special, h->aniso, stream));
      h->launches++;
      return enqueue_row_exact(flags & ~B200CV_EF_REBUILD_PAIRLIST, true, partials + ROW_BLOCKS,
                               ROW_BLOCKS + ROW_EXACT_BLOCKS, gate, 0, ROW_EXACT_BLOCKS);
    };

    if (gated) {
      // captured evaluation with a pair list: the graph holds the rebuild sweep AND the list
      // walk; the mode word decides on the device which of the two does any work
      int nparts = 0;
      int rc;
      if (rows_gated) {
        static_assert(2 * ROW_BLOCKS + EXACT_BLOCKS + ROW_EXACT_BLOCKS <= PARTIALS_TAIL, "partials are sized for PARTIALS_TAIL");
        int nb = 0;
        rc = enqueue_rows_build(flags | B200CV_EF_REBUILD_PAIRLIST, h->d_partials, h->gate(), nb);
        if (rc) return rc;
        rc = enqueue_rows_walk(h->d_partials + nb, h->gate());
        if (rc) return rc;
        npart = nb + ROW_BLOCKS + ROW_EXACT_BLOCKS;
      } else {
        if (cells_usable(h) && sparse_list(h)) {
          // sized and found sparse by b200cv_pairlist_reserve before the capture
          rc = enqueue_cells(flags | B200CV_EF_REBUILD_PAIRLIST, h->gate());
          nparts = CELL_BLOCKS;
        } else {
          rc = enqueue_all_pairs(flags | B200CV_EF_REBUILD_PAIRLIST, true, h->gate(), nparts);
        }
        if (rc) return rc;
        rc = enqueue_walk(h->d_partials + nparts, h->gate());
        if (rc) return rc;
        npart = nparts + WALK_BLOCKS;
      }
    } else if (walk_list) {
      int rc = h->pl_rows ? enqueue_rows_walk(h->d_partials, nullptr)
                          : enqueue_walk(h->d_partials, nullptr);
      if (rc) return rc;
      npart = h->pl_rows ? ROW_BLOCKS + ROW_EXACT_BLOCKS : WALK_BLOCKS;
    } else if (rows_eval) {
      int rc = enqueue_rows_build(flags, h->d_partials, nullptr, npart);
      if (rc) return rc;
    } else if (rebuild_with_cells(h, flags)) {
      int rc = enqueue_cells(flags, nullptr);
      if (rc) return rc;
      npart = CELL_BLOCKS;
    } else {
      int rc = enqueue_all_pairs(flags, rebuild, skip_gated ? h->gate() : nullptr, npart, 0);
      if (rc) return rc;
    }
  }

  if (!finalized) {
    FinalizeArgs const F = finalize_args(npart);
    CUDA_OK(h, launch_finalize(F, stream));
    h->launches++;
  }

  if (h->world > 1) {
    int const rc = comm_allreduce(h, stream);
    if (rc) return rc;
  }
  if (h->dinv()) {
    DistInvPostArgs D{};
    D.d_value = h->d_value;
    D.h_value = h->world > 1 ? nullptr : h->h_value_dev;
    D.grad = grad ? h->d_reduce + 32 : nullptr;
    D.count = h->reduce_count - 32;
    D.exponent = h->cfg.exp_numer;
    D.npairs = (double)h->g[0].n * (double)h->g[1].n;
    CUDA_OK(h, launch_distinv_post(D, stream));
    h->launches += grad ? 2 : 1;
  }
  if (h->world > 1) {
    // value + the ranks' overflow count (d_reduce[1], see finalize_body)
    CUDA_OK(h, cudaMemcpyAsync(h->h_value, h->d_value, 2 * sizeof(double), cudaMemcpyDeviceToHost,
                               stream));
  }
  return B200CV_OK;
}

int reset_gradients(b200cv_cvc *h, cudaStream_t stream)
{
  CUDA_OK(h, cudaMemsetAsync(h->d_reduce, 0, sizeof(double) * h->reduce_count, stream));
  return B200CV_OK;
}

int calc_flags(b200cv_cvc *h, long long step, int gradients)
{
  // compute_coordnum (src/colvarcomp_coordnums.cpp:251-271)
  int flags = gradients ? B200CV_EF_GRADIENTS : 0;
  bool const use_pl = h->cfg.tolerance > 0 && h->cfg.kind != B200CV_GROUPCOORD;
  if (use_pl) {
    flags |= B200CV_EF_USE_PAIRLIST;
    if (step % h->cfg.pairlist_freq == 0) flags |= B200CV_EF_REBUILD_PAIRLIST;
  }
  return flags;
}

// After the stream has drained: queue overflow -> grow and redo the evaluation.
// Returns 1 when a redo was enqueued.
int post_sync(b200cv_cvc *h, bool &redo)
{
  redo = false;
  unsigned long long const rare_n = h->h_counters[0], pl_n = h->h_counters[1];
  unsigned long long const units = h->h_counters[3], row_pairs = h->h_counters[4];
  // gated (captured) evaluations: finalize_kernel reports the mode the finished launch ran in
  bool const rebuild = h->gated ? (h->h_counters[2] == 1)
                                : (bool)(h->last_flags & B200CV_EF_REBUILD_PAIRLIST);
  bool const rows_built = rebuild && h->pl_rows;
  unsigned long long const row_entries = h->h_counters[5];
  h->exact_pairs = (size_t)rare_n;
  if (rare_n > h->rare_cap) {
    if (h->async_eval) {
      // a captured graph holds the queue's address: it cannot be re-allocated here
      return fail(h, B200CV_MEMORY_ERROR,
                  "exact-pair queue overflow in a captured evaluation (" +
                      std::to_string(rare_n) + " pairs with |l2-1| < 1/16, capacity " +
                      std::to_string(h->rare_cap) + "): this step's result is incomplete");
    }
    int rc = ensure_rare_capacity(h, rare_n);
    if (rc) return rc;
    redo = true;
  }
  if (rebuild && !(h->center1() || h->center2())) {
    bool const over_pl = pl_n > h->pl_cap;
    bool const over_rows = rows_built && (units + row_slots(h) > h->rows.cap_chunks ||
                                          row_entries > h->rows.cap_idx);
    if ((over_pl || over_rows) && h->async_eval) {
      return fail(h, B200CV_MEMORY_ERROR,
                  "pair-list overflow in a captured evaluation (" + std::to_string(pl_n) +
                      " listed pairs, capacity " + std::to_string(h->pl_cap) + "; " +
                      std::to_string(row_entries) + " row entries, capacity " +
                      std::to_string(h->rows.cap_idx) +
                      "): reserve more with b200cv_pairlist_reserve before capturing");
    }
    if (over_pl) {
      int rc = ensure_pairlist_capacity(h, pl_n);
      if (rc) return rc;
    }
    if (over_rows) {
      h->epoch++;
      CUDA_OK(h, rows_reserve(h->rows,
                              (unsigned)std::min<unsigned long long>(row_entries + row_entries / 4 + 1024, 0xfffffff0ull),
                              (unsigned)std::min<unsigned long long>(
                                  row_slots(h) + units + units / 4 + 1024, 0xfffffff0ull)));
    }
    if (over_pl || over_rows) {
      redo = true;
    } else {
      h->pl_entries = pl_n;
      h->row_entries = rows_built ? row_entries : 0ull;
      h->pl_n = pl_n + (rows_built ? row_pairs : 0ull);
      h->pl_valid = true;
    }
  }
  return B200CV_OK;
}

// After the stream has drained: commit what the evaluation left behind (list lengths; the
// device keeps its own copies, written by the final sum) and repair queue overflows by growing
// the queue and running the evaluation again.  Sharded runs agree on the redo through the
// overflow word that travels with the all-reduce (d_reduce[1]), so that every rank re-enters
// the collective together.
int sync_and_check(b200cv_cvc *h)
{
  for (int attempt = 0; attempt < 6; attempt++) {
    CUDA_OK(h, cudaStreamSynchronize(h->last_stream));
    bool redo = false;
    int rc = post_sync(h, redo);
    if (rc) {
      h->unchecked = false;  // reported; the caller reserves again or gives up
      return rc;
    }
    if (h->world > 1) {
      if (h->comm) {
        redo = redo || (h->h_value[1] > 0.5);
      } else if (redo) {
        // b200cv_set_shard without a communicator: the caller reduces, nobody to agree with
        return fail(h, B200CV_MEMORY_ERROR,
                    "pair queue overflow in a sharded run without a communicator");
      }
    }
    if (!redo) {
      h->unchecked = false;
      return B200CV_OK;
    }
    rc = reset_gradients(h, h->last_stream);
    if (rc) return rc;
    rc = prepare_calc(h, h->last_flags);
    if (rc) return rc;
    rc = enqueue_calc(h, h->last_flags, h->last_stream);
    if (rc) return rc;
    if (h->fit_after_calc) {
      // the fit gradients were derived from the incomplete first pass
      rc = b200cv_calc_fit_gradients(h, h->last_stream);
      if (rc) return rc;
      h->fit_after_calc = true;
    }
  }
  return fail(h, B200CV_BUG_ERROR, "queue capacity did not converge");
}

}  // namespace b200cv
