// engine.h -- host-side state of one coordination-number component and the internal interfaces
// between the translation units that implement the C ABI of include/b200cv.h:
//   api.cu        create / destroy / groups / cell / read_data / calc_value / value / apply_force,
//                 bare-array entries, inspection
//   planner.cu    work items of the tile sweep (per rank), b200cv_plan_items
//   calc.cu       one evaluation = the kernels of coordnum::calc_value; pair-list state machine,
//                 queue capacities, overflow redo
//   graph_step.cu b200cv_step: cache of CUDA graphs of the fused device-resident step
//   comm.cu       NCCL communicator (dlopen), sharding
//   host_path.cu  host-buffer entries (CPU-proxy arrays): eval_host, calc_host, apply_force_host
//   measure.cu    FP64 peak and reciprocal self-test helpers
// No CPU compute path exists here: every evaluation is a kernel launch, and creating a handle
// without a CUDA device fails.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200cv.h"
#include "cells.cuh"
#include "fit.cuh"
#include "kernels.cuh"
#include "rows.cuh"

namespace b200cv {

// NVTX ranges around the phases of a step, named like the reference's GPU scheduler phases
// (src/colvar_gpu_calc.cpp:711-722).  Header-only NVTX3: a no-op unless a profiler is attached.
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

inline size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

struct Group {
  int n = 0;
  size_t stride = 0;  // plane stride of d_pos / d_grad (multiple of 32, >= n + 32)
  int *d_index = nullptr;
  double *d_mass = nullptr;
  double *d_weight = nullptr;
  double total_mass = 0.0;
  double *d_pos = nullptr;
  double *d_grad = nullptr;     // points into the reduce buffer
  double *d_centers = nullptr;  // com[3] cog[3] com_grad[3] (+3 pad)
  double *d_com_scratch = nullptr;       // per-CTA partials of com_cog_kernel
  unsigned int *d_com_ticket = nullptr;  // its last-block-done counter
  std::vector<int> h_index;
  int max_index = -1;
  std::vector<double> h_mass;
  FitState fit;
  double *com() const { return d_centers; }
  double *cog() const { return d_centers + 3; }
  double *com_grad() const { return d_centers + 6; }
};

}  // namespace b200cv

struct b200cv_cvc {
  b200cv_config cfg;
  b200cv::PairParams P;
  double l2_max = 1.0e20;
  bool aniso = false;
  int exp_special = 0;
  b200cv::Group g[2];

  // value / counters
  double *d_reduce = nullptr;  // [value | pad | grad1 (3*stride1) | grad2 (3*stride2)]
  size_t reduce_count = 0;
  double *d_value = nullptr;
  double *h_value = nullptr;                // mapped pinned
  double *h_value_dev = nullptr;            // device alias of h_value
  unsigned long long *h_counters = nullptr; // mapped pinned [8]: rare, listed (entries), gate mode,
                                            // overflow chunks / listed pairs / entries of the rows
  unsigned long long *h_counters_dev = nullptr;
  unsigned char *d_counters = nullptr;      // rare_count (u32 @0), pl_count (u64 @8),
                                            // persistent list length (u64 @16), gate (i32 @24),
                                            // fused-finalize ticket (u32 @28)
  // captured pair-list evaluations: 1 = the next launch rebuilds the list, 0 = it walks it
  int *h_mode = nullptr;                    // mapped pinned, written by the host between launches
  int *h_mode_dev = nullptr;
  bool gated = false;                       // last evaluation was enqueued in gated form
  double *d_partials = nullptr;
  int partial_cap = 0;

  // queues
  unsigned long long *d_rare = nullptr;
  unsigned int rare_cap = 0;
  unsigned long long *d_pl = nullptr;
  unsigned long long pl_cap = 0;
  unsigned long long pl_n = 0;  // pairs of the current list (all forms)
  unsigned long long pl_entries = 0;  // ... of them as packed entries in d_pl (tile form: the extras)
  bool pl_valid = false;        // a rebuild has happened
  // entries of the current list are in cell order (built by the cell kernel with
  // sorted_entries): the walk runs on the cell-ordered copies (cells.cuh, launch_cell_gather)
  bool pl_sorted = false;
  unsigned char *d_pl_center = nullptr;  // dense flags for the centre-only variants

  // cell-list rebuild of the pair list (cells.cuh); used once a rebuild has shown the list to
  // be sparse
  b200cv::CellBuffers cells;
  bool cells_ready = false;
  // the sparse list of an open system as neighbour rows (rows.cuh): both groups sorted along a
  // cell grid (cellsort.cuh), every atom keeps the sorted positions of all its listed partners;
  // d_pl then holds the extras (pairs listed by the reference-order path), atom indices
  b200cv::CellSort sort;
  bool sort_ready = false;
  b200cv::RowList rows;
  bool pl_rows = false;    // the current list is in row form
  unsigned long long row_entries = 0;  // entries of all rows of this rank (padding included)
  // an evaluation has been enqueued whose queue counters / new list lengths the host has not
  // looked at yet (sync_and_check): the next call that builds on its result checks first
  bool unchecked = false;
  // calc_fit_gradients has been enqueued behind the last evaluation (an overflow redo repeats it)
  bool fit_after_calc = false;

  // sweep work items
  b200cv::SweepItem *d_items = nullptr;
  int nitems = 0;
  int chunk = b200cv::JCMAX;
  bool items_dirty = true;

  // sharding
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;

  // distancePairs: the n1 x n2 matrix of distances (or of force components) on the device
  double *d_pairs = nullptr;
  size_t pairs_cap = 0;

  // host staging (calc_host / apply_force_host)
  double *d_stage_pos = nullptr;
  size_t stage_n = 0;
  double *h_stage_grad = nullptr;  // pinned
  size_t h_stage_count = 0;
  cudaStream_t own_stream = nullptr;

  // last calc (for b200cv_value and the overflow redo)
  cudaStream_t last_stream = nullptr;
  long long last_step = 0;
  int last_flags = 0;
  bool have_calc = false;
  bool data_read = false;
  bool async_eval = false;  // last evaluation cannot be re-run (graph capture)
  int launches = 0;
  size_t exact_pairs = 0;

  // CUDA graphs of the fused device-resident step (b200cv_step), keyed by what they froze
  struct StepGraph {
    int flags;
    const double *pos;
    size_t stride;
    unsigned long long epoch;
    int launches;
    cudaGraphExec_t exec;
  };
  std::vector<StepGraph> step_graphs;
  unsigned long long epoch = 0;  // bumped whenever a buffer, the cell or the work items change

  std::string err;

  unsigned int *rare_count() const { return reinterpret_cast<unsigned int *>(d_counters); }
  unsigned long long *pl_count() const
  {
    return reinterpret_cast<unsigned long long *>(d_counters + 8);
  }
  int *gate() const { return reinterpret_cast<int *>(d_counters + 24); }
  unsigned int *fin_ticket() const { return reinterpret_cast<unsigned int *>(d_counters + 28); }
  bool self() const { return cfg.kind == B200CV_SELFCOORDNUM; }
  bool dinv() const { return cfg.kind == B200CV_DISTANCEINV; }
  bool center1() const { return cfg.group1_center_only != 0; }
  bool center2() const { return cfg.group2_center_only != 0; }
  b200cv::Group &grp2() { return self() ? g[0] : g[1]; }
};

namespace b200cv {

// message of the last failed call without a handle (b200cv_create, comm_unique_id), per thread
std::string &create_error();

inline int fail(b200cv_cvc *h, int code, const std::string &msg)
{
  if (h) h->err = msg;
  else create_error() = msg;
  return code;
}

#define CUDA_OK(h, call)                                                                     \
  do {                                                                                       \
    cudaError_t e_ = (cudaError_t)(call);                                                    \
    if (e_ != cudaSuccess) {                                                                 \
      return fail(h, e_ == cudaErrorMemoryAllocation ? B200CV_MEMORY_ERROR : B200CV_ERROR,   \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                       \
    }                                                                                        \
  } while (0)

template <typename T> void free_dev(T *&p)
{
  if (p) cudaFree(p);
  p = nullptr;
}

#define CUDA_OK_RC(expr)          \
  do {                            \
    int rc_ = (expr);             \
    if (rc_ != B200CV_OK) return rc_; \
  } while (0)

// ---- planner.cu -------------------------------------------------------------------------------
void plan_items(bool self, int n1, int n2_in, int rank, int world, std::vector<SweepItem> &mine,
                int &chunk_out);
int build_items(b200cv_cvc *h);
int ensure_partials(b200cv_cvc *h, int need);

// ---- calc.cu ----------------------------------------------------------------------------------
bool groups_ready(b200cv_cvc *h);
bool centers_in_step(const b200cv_cvc *h, const Group &G);
int centers_now(b200cv_cvc *h, Group &G, cudaStream_t s);
int centers_after_read(b200cv_cvc *h, Group &G, cudaStream_t s);
bool sorted_walk_enabled();
bool sparse_list(const b200cv_cvc *h);
bool cells_usable(const b200cv_cvc *h);
bool rebuild_with_cells(const b200cv_cvc *h, int flags);
bool rows_usable(const b200cv_cvc *h);
bool rebuild_with_rows(const b200cv_cvc *h, int flags);
unsigned row_slots(const b200cv_cvc *h);
// allocations an evaluation with these flags may need (the row list): before any stream capture
int prepare_calc(b200cv_cvc *h, int flags);
int ensure_pairlist_capacity(b200cv_cvc *h, unsigned long long need);
int ensure_rare_capacity(b200cv_cvc *h, unsigned long long need);
// all kernels of one coordnum::calc_value on `stream`
int enqueue_calc(b200cv_cvc *h, int flags, cudaStream_t stream);
int reset_gradients(b200cv_cvc *h, cudaStream_t stream);
// compute_coordnum's flags for a step (src/colvarcomp_coordnums.cpp:251-271)
int calc_flags(b200cv_cvc *h, long long step, int gradients);
// waits for the last evaluation, commits pair-list state, grows queues and re-runs on overflow
int sync_and_check(b200cv_cvc *h);

// ---- comm.cu ----------------------------------------------------------------------------------
// the one collective of the path: [value | grad1 | grad2] summed over ranks, in place
int comm_allreduce(b200cv_cvc *h, cudaStream_t stream);
// buf holds world * count_per_rank doubles; rank r has filled its own block: gather the others
int comm_allgather(b200cv_cvc *h, double *buf, size_t count_per_rank, cudaStream_t stream);
void comm_destroy(b200cv_cvc *h);

}  // namespace b200cv
