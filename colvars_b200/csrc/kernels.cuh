// kernels.cuh -- launch interface between the C-ABI layer (api.cu) and the kernels.
#pragma once
#include "sweep.cuh"

namespace b200cv {

// every kernel this library launches is counted (bench.py reports the number it launched
// inside its timed region): one relaxed atomic increment on the host per launch
void note_launch();
unsigned long long launches_so_far();
#if defined(B200CV_TRACE)
int trace_dump(unsigned long long *out, int n);  // sweep_inst.cu, debug builds only
#endif

// sweep geometry (compile-time): R register atoms per lane, W warps per CTA
#ifndef B200CV_SWEEP_R
#define B200CV_SWEEP_R 8
#endif
#ifndef B200CV_SWEEP_W
#define B200CV_SWEEP_W 4
#endif
#ifndef B200CV_SWEEP_MINB
#define B200CV_SWEEP_MINB 1
#endif
constexpr int SWEEP_R = B200CV_SWEEP_R;
constexpr int SWEEP_W = B200CV_SWEEP_W;
constexpr int SWEEP_MINB = B200CV_SWEEP_MINB;  // CTAs per SM the register allocation targets
constexpr int SWEEP_IB = SWEEP_W * 32 * SWEEP_R;  // group1 atoms per work item (1024)

// Launch the tile sweep over `nitems` work items.  exp_special = n' when ed == 2*en and a
// specialised instantiation exists (currently n' = 3, i.e. 6/12), else 0 (runtime exponents).
// pbc: BC_NONE or BC_ORTHO.  Returns a cudaError_t.
int launch_sweep(const SweepArgs &args, int nitems, int exp_special, bool aniso, int pbc,
                 bool pairlist, cudaStream_t stream);
bool sweep_has_special(int en, int ed);

// value = sum(partials[0..n)) in a fixed order -> device scalar + mapped pinned host copy;
// also publishes the queue counters to the host.
struct FinalizeArgs {
  const double *partials;
  int n;
  double *d_value;
  double *h_value;  // mapped pinned
  const unsigned int *rare_count;
  const unsigned long long *pl_count;
  // mapped pinned [8]: rare_count, pl_count, gate mode, overflow chunks of the rows, listed
  // pairs in rows, row entries
  unsigned long long *h_counters;
  // after a rebuild -- `rebuilt` != 0, or *gate == 1 in a gated (captured) evaluation -- the
  // lengths of the new list are kept on the device for the list steps that follow (clamped to
  // the capacities; the host sees the unclamped counts and redoes the step on overflow)
  const int *gate;
  int rebuilt;
  unsigned long long *pl_persist;
  unsigned long long pl_cap;
  unsigned *row_counts;     // CellSort::d_counts (counters of the row list, rows.cuh) or null
  unsigned row_cap_chunks;  // capacity in overflow chunks (counts[0])
  unsigned row_cap_idx;     // capacity in entries (counts[3])
  // sharded runs: 1.0 when a queue of this rank overflowed, else 0.0, into the word that the
  // all-reduce sums next to the value (null: single rank)
  unsigned int rare_cap;
  double *overflow;
};
// Pairs queued by the sweep (or every listed pair on non-rebuild steps, or every pair for
// the general-PBC fallback) evaluated in the reference's operation order.
struct ExactArgs {
  const double *x1, *y1, *z1;
  const double *x2, *y2, *z2;
  double *g1x, *g1y, *g1z;
  double *g2x, *g2y, *g2z;
  const unsigned long long *list;      // packed (i << 32 | j)
  const unsigned int *count32;         // number of entries (device), or NULL
  const unsigned long long *count64;   // ... 64-bit variant (pair list), or NULL
  unsigned long long cap;              // entries beyond cap were never written
  // a second list evaluated behind the first in the same launch (tile walk: exact-pair queue +
  // the extras of the tile list), or NULL
  const unsigned long long *list2;
  const unsigned long long *count2;
  unsigned long long cap2;
  // row form of the list (rows.cuh): what the rows added to their partners, in cell order and
  // unscaled -- added to the gradient of group `gs_group` (1 or 2) before anything else, or NULL
  const double *gs;
  size_t gs_stride;
  const int *gs_sorted;                // atom at a sorted position
  int gs_n;
  int gs_group;
  double *partials;                    // one per CTA
  unsigned long long *pl;              // pair list to append to on rebuild (or NULL)
  unsigned long long pl_cap;
  unsigned long long *pl_count;
  int flags;                           // EF_*
  int fast_exp;                        // 3: pair-list walk may use the sweep's fast 6/12 form
  const int *gate;                     // see SweepArgs::gate
  int gate_on;
  // fused finalize (fin_ticket != NULL, never together with a gate): the last CTA to finish
  // sums fin.partials and publishes the value; the ticket re-arms itself
  unsigned int *fin_ticket;
  FinalizeArgs fin;
  PairParams P;
};
constexpr int EXACT_BLOCKS = 296;
constexpr int EXACT_THREADS = 128;
// the pair-list walk is a latency-bound gather (6 coordinate loads per entry): it wants every
// warp slot the register file allows (58 registers -> 8 CTAs per SM)
constexpr int WALK_BLOCKS = 1184;
// partial sums behind the sweep items: [rare queue | walk], [cell kernel (592) | walk] or
// [row build (| rows of group2) | its exact queue | row walk | its exact lists]
constexpr int PARTIALS_TAIL = 6144;
int launch_exact(const ExactArgs &args, cudaStream_t stream, int blocks = EXACT_BLOCKS);
// list[k] = (k, k) for k < n, *count = n: a batch of independent single pairs for exact_kernel
int launch_fill_diagonal(unsigned long long *list, unsigned int *count, int n, cudaStream_t stream);

// Dense fallback for boundary conditions the tile sweep does not specialise (mixed /
// triclinic): every pair through pair_exact, gradients by atomics.  self: i < j only.
struct DenseArgs {
  const double *x1, *y1, *z1;
  const double *x2, *y2, *z2;
  double *g1x, *g1y, *g1z;
  double *g2x, *g2y, *g2z;
  int n1, n2, self;
  int i_begin, i_end;  // rows of this shard
  double *partials;    // [EXACT_BLOCKS]
  unsigned long long *pl;
  unsigned long long pl_cap;
  unsigned long long *pl_count;
  int flags;
  const int *gate;  // see SweepArgs::gate
  int gate_on;
  PairParams P;
};
int launch_dense(const DenseArgs &args, cudaStream_t stream);

// One pseudo-atom (a centre of mass held in device memory) against a group of n atoms:
// coordNum with group{1,2}CenterOnly and groupCoord (src/colvarcomp_coordnums.cpp:190-247).
struct CenterArgs {
  const double *center;     // device double[3]: the COM pseudo-atom
  const double *x, *y, *z;  // the other side: n atoms (or another centre when n == 0)
  const double *center2;    // used when both sides are centres
  double *gx, *gy, *gz;     // gradients of the n atoms (accumulated)
  double *center_grad;      // device double[3] += gradient of `center`
  double *center2_grad;     // when both are centres
  int n;
  int center_is_group1;     // diff = p2 - p1: which side the centre is on
  unsigned char *pairlist;  // dense bool per pair (n entries) or NULL
  double *partials;         // [EXACT_BLOCKS]
  int flags;
  const int *gate;          // non-null: *gate (1 rebuild / 0 walk) replaces EF_REBUILD_PAIRLIST
  PairParams P;
};
int launch_center(const CenterArgs &args, cudaStream_t stream);

// colvar::distance_pairs: dist[i * n2 + j] = |position_distance(p1_i, p2_j)| (force == NULL), or
// the forces f1_i -= force_ij * unit, f2_j += force_ij * unit accumulated into the f planes
struct DistPairsArgs {
  const double *x1, *y1, *z1;
  const double *x2, *y2, *z2;
  int n1, n2;
  double *dist;         // [n1 * n2]
  const double *force;  // [n1 * n2] or NULL
  double *f1x, *f1y, *f1z;
  double *f2x, *f2y, *f2z;
  PairParams P;
};
int launch_distance_pairs(const DistPairsArgs &args, cudaStream_t stream);

int launch_finalize(const FinalizeArgs &args, cudaStream_t stream);

// copies the host-written mode word (mapped pinned) into device memory once per evaluation:
// 2 when the host asked to skip the launch, else rebuild (1) / walk (0) for pair-list
// evaluations (with_list) and 0 for the others
int launch_fetch_gate(const int *h_mode_dev, int *d_gate, int with_list, cudaStream_t stream);

// distanceInv epilogue (src/colvarcomp_distances.cpp:442-455): *d_value holds sum_ij d^-n
// (already reduced over ranks); x = (sum / npairs)^(-1/n) replaces it and every gradient entry
// in grad[0..count) is scaled by dx/dsum = (-1/n) x^(n+1) / npairs.
struct DistInvPostArgs {
  double *d_value;
  double *h_value;  // mapped pinned or null
  double *grad;     // contiguous gradient block (group1 | group2), or null
  size_t count;
  int exponent;
  double npairs;
};
int launch_distinv_post(const DistInvPostArgs &args, cudaStream_t stream);

// ---- atom-group kernels (src/colvaratoms.cpp, src/cuda/colvaratoms_kernel.cu) -----------
// gather: pos[d*stride + i] = proxy[d*pstride + idx[i]]  (aos != 0: proxy[3*idx[i] + d])
int launch_gather(const double *proxy, size_t pstride, int aos, const int *idx, int n,
                  double *pos, size_t stride, cudaStream_t stream);
// both groups of a component in one launch (steps of a few hundred microseconds are bound by
// the number of launches, not by these kernels' work)
struct GroupView {
  const int *idx;
  int n;
  double *data;     // positions (gather) or gradients (scatter), SoA planes
  size_t stride;
  const double *q;  // scatter: device quaternion of a rotated group, or NULL
};
int launch_gather2(const double *proxy, size_t pstride, int aos, const GroupView &a,
                   const GroupView &b, cudaStream_t stream);
int launch_scatter_force2(const GroupView &a, const GroupView &b, double f, double *proxy_force,
                          size_t pstride, int aos, cudaStream_t stream);
// COM = sum m x / M and COG = sum x / n in one pass (<= 296 CTAs, fixed summation order).
// scratch: device double[296 * 6]; ticket: device counter, zero before the first launch.
constexpr int COM_SCRATCH_DOUBLES = 296 * 6;
int launch_com_cog(const double *pos, size_t stride, const double *mass, double total_mass,
                   int n, double *com, double *cog, double *scratch, unsigned int *ticket,
                   cudaStream_t stream);
// pos += s * t (t in device memory, plus optional constant shift c)
int launch_translate(double *pos, size_t stride, int n, const double *t_dev, double s,
                     cudaStream_t stream);
// pos = R pos, R = rotation matrix of device quaternion q (inverse != 0: R^T)
int launch_rotate(double *pos, size_t stride, int n, const double *q_dev, cudaStream_t stream);
// grad[i] = w[i] * g (assignment; set_weighted_gradient)
int launch_weighted_gradient(const double *w, const double *g_dev, int n, double *grad,
                             size_t stride, cudaStream_t stream);
// F_proxy[idx[i]] += f * (Rinv grad_i); q_dev == NULL: no rotation
int launch_scatter_force(const double *grad, size_t stride, const int *idx, int n, double f,
                         const double *q_dev, double *proxy_force, size_t pstride, int aos,
                         cudaStream_t stream);
// copy n atoms between SoA layouts with different strides (dst += src when accumulate)
int launch_repack(const double *src, size_t sstride, double *dst, size_t dstride, int n,
                  int accumulate, cudaStream_t stream);

// ---- measurement helpers ------------------------------------------------------------------
int launch_fp64_peak(double *d_out, int iters, int blocks, cudaStream_t stream);
int launch_rcp_selftest(const double *d_x, int n, double *d_maxulp, cudaStream_t stream);

}  // namespace b200cv
