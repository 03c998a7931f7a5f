// api.cu -- implementation of the C ABI declared in include/b200cv.h.
//
// Host-side state of one coordination-number component: the two atom groups (device SoA
// buffers, padded so that every 32-atom tile and every TMA bulk copy is in bounds), the
// pair parameters, the work-item table of the tile sweep, the exact-pair queue, the
// compacted pair list and the NCCL communicator.  No CPU compute path exists here: every
// evaluation is a kernel launch, and creating a handle without a CUDA device fails.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/b200cv.h"
#include "cells.cuh"
#include "fit.cuh"
#include "kernels.cuh"

using namespace b200cv;

namespace {

// message of the last failed b200cv_create on this thread (no handle exists to carry it)
thread_local std::string g_create_error;

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

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::mutex mu;
  bool load(std::string &err)
  {
    std::lock_guard<std::mutex> lock(mu);
    if (lib && AllReduce) return true;
    // picks up the copy already mapped into the process (e.g. torch's) or the system one
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      err = std::string("cannot load NCCL: ") + dlerror();
      return false;
    }
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce) {
      err = "NCCL library lacks required symbols";
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;

}  // namespace

struct b200cv_cvc {
  b200cv_config cfg;
  PairParams P;
  double l2_max = 1.0e20;
  bool aniso = false;
  int exp_special = 0;
  Group g[2];

  // value / counters
  double *d_reduce = nullptr;  // [value | pad | grad1 (3*stride1) | grad2 (3*stride2)]
  size_t reduce_count = 0;
  double *d_value = nullptr;
  double *h_value = nullptr;                // mapped pinned
  double *h_value_dev = nullptr;            // device alias of h_value
  unsigned long long *h_counters = nullptr; // mapped pinned [4]: rare, listed, gate mode, -
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
  unsigned long long pl_n = 0;  // entries of the current list
  bool pl_valid = false;        // a rebuild has happened
  // entries of the current list are in cell order (built by the cell kernel with
  // sorted_entries): the walk runs on the cell-ordered copies (cells.cuh, launch_cell_gather)
  bool pl_sorted = false;
  unsigned char *d_pl_center = nullptr;  // dense flags for the centre-only variants

  // cell-list rebuild of the pair list (cells.cuh); used once a rebuild has shown the list to
  // be sparse
  CellBuffers cells;
  bool cells_ready = false;

  // sweep work items
  SweepItem *d_items = nullptr;
  int nitems = 0;
  int chunk = JCMAX;
  bool items_dirty = true;

  // sharding
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;

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
  Group &grp2() { return self() ? g[0] : g[1]; }
};

namespace {

int fail(b200cv_cvc *h, int code, const std::string &msg)
{
  if (h) h->err = msg;
  else g_create_error = msg;
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

// coordnum::compute_tolerance_l2_max (src/colvarcomp_coordnums.cpp:162-184): Newton on the
// pair-list-shifted switching function, host side, reference operation order.
double integer_power_host(double x, int n)
{
  if (x == 0.0) return 0.0;
  double yy = 1.0, ww = x;
  for (int nn = n; nn != 0; nn >>= 1, ww *= ww) {
    if (nn & 1) yy *= ww;
  }
  return yy;
}

double switching_host(double l2, double &dFdl2, int en, int ed, double tol)
{
  int const en2 = en / 2, ed2 = ed / 2;
  double const xn = integer_power_host(l2, en2), xd = integer_power_host(l2, ed2);
  double const h = l2 - 1.0, en2_r = en2, ed2_r = ed2;
  double f0;
  if (std::fabs(h) < 1.0e-7) {
    double const c0 = en2_r / ed2_r;
    double const c1 = (en2_r * (en2_r - ed2_r)) / (2.0 * ed2_r);
    double const c2 = (en2_r * (en2_r - ed2_r) * (2.0 * en2_r - ed2_r - 3.0)) / (12.0 * ed2_r);
    f0 = c0 + h * (c1 + h * c2);
  } else {
    f0 = (1.0 - xn) / (1.0 - xd);
  }
  double const inv = 1 / (1.0 - tol);
  double const func = (f0 - tol) * inv;
  if (func < 0) return 0;
  double ld;
  if (std::fabs(h) < 1.0e-7) {
    ld = 0.5 * (en2_r - ed2_r) + h * (((en2_r - ed2_r) * (en2_r + ed2_r - 6.0)) / 12.0);
  } else {
    ld = (ed2_r * xd / ((1.0 - xd) * l2)) - (en2_r * xn / ((1.0 - xn) * l2));
  }
  dFdl2 = f0 * inv * ld;
  return func;
}

double compute_tolerance_l2_max(int en, int ed, double tol)
{
  double l2 = 1.001, F = 0.0, dF = 0.0;
  for (size_t i = 0; i < 1000000; i++) {
    F = switching_host(l2, dF, en, ed, tol);
    if ((std::fabs(F) < 1.0e-6) || (std::fabs(dF) < 1.0e-9)) break;
    l2 -= F / dF;
  }
  return l2;
}

int hi_word(double v)
{
  unsigned long long b;
  std::memcpy(&b, &v, 8);
  return (int)(b >> 32);
}

void fill_pair_params(b200cv_cvc *h)
{
  PairParams &P = h->P;
  b200cv_config const &c = h->cfg;
  // update_cutoffs (src/colvarcomp_coordnums.cpp:28-43)
  for (int d = 0; d < 3; d++) {
    P.inv_r0[d] = 1.0 / c.cutoff3[d];
    P.inv_r0sq[d] = P.inv_r0[d] * P.inv_r0[d];
  }
  h->aniso = !(c.cutoff3[0] == c.cutoff3[1] && c.cutoff3[0] == c.cutoff3[2]);
  P.en2 = c.exp_numer / 2;
  P.ed2 = c.exp_denom / 2;
  P.tol = c.tolerance > 0 ? c.tolerance : 0.0;
  P.inv1mt = 1 / (1.0 - P.tol);
  P.ntol = -P.tol * P.inv1mt;
  P.l2_max = h->l2_max;
  // decision window of the pair list: l2_max (Newton, reference order) and the exact root of
  // f0(l2) = tol (bisection; the function is monotonic on [1, 1e12] whenever a root exists)
  {
    int lo = hi_word(P.l2_max), hi = lo;
    if (P.tol > 0) {
      auto f0 = [&](double u) {
        double d = 0.0;
        return switching_host(u, d, c.exp_numer, c.exp_denom, 0.0) - P.tol;
      };
      double a = 1.0 + 1e-6, bb = 1e12;
      if (f0(a) * f0(bb) < 0) {
        for (int it = 0; it < 200; it++) {
          double const mid = 0.5 * (a + bb);
          if (f0(a) * f0(mid) <= 0) bb = mid; else a = mid;
        }
        int const r = hi_word(0.5 * (a + bb));
        lo = std::min(lo, r);
        hi = std::max(hi, r);
      }
    }
    P.thr_lo_hi = lo - 1;
    P.thr_width = (unsigned)(hi - lo + 2);
  }
  h->exp_special = sweep_has_special(c.exp_numer, c.exp_denom) ? P.en2 : 0;
  for (int d = 0; d < 3; d++) {
    P.c_grad[d] = h->exp_special ? -2.0 * (double)P.en2 * P.inv1mt * P.inv_r0sq[d]
                                 : 2.0 * P.inv_r0sq[d];
  }
  P.func = 0;
  if (h->dinv()) {
    // pair term d^-n on the unscaled squared distance; G = -n F / d^2 * diff
    // (src/colvarcomp_distances.cpp:415-427)
    P.func = 1;
    h->aniso = false;
    h->exp_special = (P.en2 == 3) ? -3 : -1;
    for (int d = 0; d < 3; d++) {
      P.inv_r0[d] = P.inv_r0sq[d] = 1.0;
      P.c_grad[d] = -(double)c.exp_numer;
    }
  }
}

void reset_boundaries(PairParams &P)
{
  P.bc_type = 0;
  std::memset(P.periodic, 0, sizeof(P.periodic));
  std::memset(P.cell, 0, sizeof(P.cell));
  std::memset(P.recip, 0, sizeof(P.recip));
  std::memset(P.cell2, 0, sizeof(P.cell2));
  std::memset(P.tri_n, 0, sizeof(P.tri_n));
  std::memset(P.tri_shift, 0, sizeof(P.tri_shift));
  P.tri_margin = 0.0;
}

// tables of the tile sweep's image search in mixed / triclinic cells (general_image,
// pair_math.cuh): the 13 shift vectors in the operation order of get_triclinic_shift
// (src/colvars_system.h:207), their squared lengths, and the decision margin
void fill_general_cell_tables(PairParams &P)
{
  double max_n = 0.0;
  for (int k = 0; k < 13; k++) {
    int const ix = tri_dir(k, 0), iy = tri_dir(k, 1), iz = tri_dir(k, 2);
    bool const allowed = (ix == 0 || P.periodic[0]) && (iy == 0 || P.periodic[1]) &&
                         (iz == 0 || P.periodic[2]);
    double n2 = 0.0;
    for (int d = 0; d < 3; d++) {
      double const sd = ((double)ix * P.cell[0][d] + (double)iy * P.cell[1][d]) +
                        (double)iz * P.cell[2][d];
      P.tri_shift[k][d] = allowed ? sd : 0.0;
      n2 += sd * sd;
    }
    P.tri_n[k] = (allowed && n2 > 0.0) ? n2 : 1.0e300;
    if (allowed) max_n = std::max(max_n, n2);
  }
  for (int a = 0; a < 3; a++) {
    for (int d = 0; d < 3; d++) P.cell2[a][d] = 2.0 * P.cell[a][d];
  }
  P.tri_margin = std::ldexp(max_n, -26);
}

int alloc_group(b200cv_cvc *h, Group &G, int n)
{
  G.n = n;
  G.stride = round_up((size_t)n, 32) + 32;
  CUDA_OK(h, cudaMalloc(&G.d_index, sizeof(int) * std::max(n, 1)));
  CUDA_OK(h, cudaMalloc(&G.d_mass, sizeof(double) * std::max(n, 1)));
  CUDA_OK(h, cudaMalloc(&G.d_weight, sizeof(double) * std::max(n, 1)));
  CUDA_OK(h, cudaMalloc(&G.d_pos, sizeof(double) * 3 * G.stride));
  {
    // padding atoms [n, stride) sit far away (-1e30; rows beyond a group use +1e30 in the
    // sweep): finite everywhere, and invisible to the specialised path without masking
    std::vector<double> init(3 * G.stride, 0.0);
    for (int d = 0; d < 3; d++) {
      for (size_t i = (size_t)n; i < G.stride; i++) init[d * G.stride + i] = -1.0e30;
    }
    CUDA_OK(h, cudaMemcpy(G.d_pos, init.data(), sizeof(double) * 3 * G.stride,
                          cudaMemcpyHostToDevice));
  }
  CUDA_OK(h, cudaMalloc(&G.d_centers, sizeof(double) * 12));
  CUDA_OK(h, cudaMemset(G.d_centers, 0, sizeof(double) * 12));
  CUDA_OK(h, cudaMalloc(&G.d_com_scratch, sizeof(double) * COM_SCRATCH_DOUBLES));
  CUDA_OK(h, cudaMalloc(&G.d_com_ticket, sizeof(unsigned int)));
  CUDA_OK(h, cudaMemset(G.d_com_ticket, 0, sizeof(unsigned int)));
  return B200CV_OK;
}

int ensure_partials(b200cv_cvc *h, int need)
{
  if (need > h->partial_cap) {
    free_dev(h->d_partials);
    CUDA_OK(h, cudaMalloc(&h->d_partials, sizeof(double) * need));
    h->partial_cap = need;
  }
  return B200CV_OK;
}

// Work items of the tile sweep for one rank: (group1 block) x (group2 chunk).  Pure host
// logic (no CUDA): also exported as b200cv_plan_items for the CPU-side sharding tests.
void plan_items(bool self, int n1, int n2_in, int rank, int world, std::vector<SweepItem> &mine,
                int &chunk_out)
{
  int const n2 = self ? n1 : n2_in;
  int const nb1 = (n1 + SWEEP_IB - 1) / SWEEP_IB;
  // chunk length: the longest that still fills the machine.  CTAs run two per SM (register
  // bound), i.e. in waves of 296 per GPU; a chunk costs (tiles + ~1.5 tiles of fixed
  // work), the last wave is as long as a full one.  Pick the candidate with the least
  // estimated time; big problems end up at JCMAX, a 1k x 100k problem at exactly one wave.
  long long const slots = 148LL * 2;
  long long chunk = JCMAX;
  double best = 1e300;
  for (int cand = 32 * SWEEP_W; cand <= JCMAX; cand += 32) {
    long long const nch = (n2 + cand - 1) / cand;
    long long items = (long long)nb1 * nch;
    if (self) items = items / 2 + nb1;  // upper triangle
    long long const per_rank = (items + world - 1) / world;
    long long const waves = (per_rank + slots - 1) / slots;
    double const t = (double)waves * ((double)cand / 32.0 + 1.5);
    if (t < best * (1.0 - 1e-9) || (t <= best * (1.0 + 1e-9) && cand > chunk)) {
      best = t;
      chunk = cand;
    }
  }
  if (const char *forced = std::getenv("B200CV_CHUNK")) {
    // tuning experiments only: group2 atoms per work item (multiple of 32, <= JCMAX)
    long long const c = std::atoll(forced);
    if (c >= 32 && c <= JCMAX && c % 32 == 0) chunk = c;
  }
  chunk_out = (int)chunk;
  int const nch = (n2 + (int)chunk - 1) / (int)chunk;

  std::vector<SweepItem> all;
  all.reserve((size_t)nb1 * nch);
  for (int b = 0; b < nb1; b++) {
    int const i0 = b * SWEEP_IB;
    for (int c = 0; c < nch; c++) {
      int const j0 = c * (int)chunk;
      int const jn = std::min((int)chunk, n2 - j0);
      int diag = 0;
      if (self) {
        if (j0 + jn <= i0 + 1) continue;       // no j > i in this chunk
        diag = (j0 < i0 + SWEEP_IB) ? 1 : 0;   // overlaps the block's own rows
      }
      all.push_back(SweepItem{i0, j0, jn, diag});
    }
  }
  mine.clear();
  if (world == 1) {
    mine.swap(all);
  } else if (!self) {
    // equal contiguous ranges of the block-major item list: a rank owns whole group1 blocks
    // except at the two ends of its range, and the load is balanced to one item
    size_t const lo = all.size() * (size_t)rank / (size_t)world;
    size_t const hi = all.size() * (size_t)(rank + 1) / (size_t)world;
    mine.assign(all.begin() + lo, all.begin() + hi);
  } else {
    // cyclic over the flat list (balances the selfCoordNum triangle)
    for (size_t k = 0; k < all.size(); k++) {
      if ((int)(k % (size_t)world) == rank) mine.push_back(all[k]);
    }
  }
  // The kernel ends when its last CTA does: with several waves of equal items the SMs drain
  // one by one over the duration of a whole item (0.4 ms for 32 tiles -- 2 % of a step at 8
  // GPUs).  Cut the items of the last two waves into quarters so the drain is a quarter as long.
  if ((long long)mine.size() > 2 * slots && !std::getenv("B200CV_CHUNK")) {
    size_t const first_small = mine.size() - (size_t)(2 * slots);
    std::vector<SweepItem> tail;
    tail.reserve((size_t)(8 * slots));
    for (size_t k = first_small; k < mine.size(); k++) {
      SweepItem const it = mine[k];
      int const tiles = (it.jn + 31) / 32;
      int const per = std::max(4, (tiles + 3) / 4);  // tiles per piece
      for (int t0 = 0; t0 < tiles; t0 += per) {
        int const j0 = it.j0 + t0 * 32;
        int const jn = std::min(per * 32, it.j0 + it.jn - j0);
        int diag = 0;
        if (self) {
          if (j0 + jn <= it.i0 + 1) continue;     // no j > i in this piece
          diag = (j0 < it.i0 + SWEEP_IB) ? 1 : 0;
        }
        tail.push_back(SweepItem{it.i0, j0, jn, diag});
      }
    }
    mine.resize(first_small);
    mine.insert(mine.end(), tail.begin(), tail.end());
  }
}

int build_items(b200cv_cvc *h)
{
  h->epoch++;
  std::vector<SweepItem> mine;
  plan_items(h->self(), h->g[0].n, h->g[1].n, h->rank, h->world, mine, h->chunk);
  free_dev(h->d_items);
  h->nitems = (int)mine.size();
  if (h->nitems > 0) {
    CUDA_OK(h, cudaMalloc(&h->d_items, sizeof(SweepItem) * mine.size()));
    CUDA_OK(h, cudaMemcpy(h->d_items, mine.data(), sizeof(SweepItem) * mine.size(),
                          cudaMemcpyHostToDevice));
  }
  int rc = ensure_partials(h, std::max(h->nitems, 1) + PARTIALS_TAIL);
  if (rc) return rc;
  h->items_dirty = false;
  return B200CV_OK;
}

int ensure_reduce_buffer(b200cv_cvc *h)
{
  if (h->d_reduce) return B200CV_OK;
  size_t const s1 = h->g[0].stride, s2 = h->self() ? 0 : h->g[1].stride;
  h->reduce_count = 32 + 3 * s1 + 3 * s2;
  CUDA_OK(h, cudaMalloc(&h->d_reduce, sizeof(double) * h->reduce_count));
  CUDA_OK(h, cudaMemset(h->d_reduce, 0, sizeof(double) * h->reduce_count));
  h->d_value = h->d_reduce;
  h->g[0].d_grad = h->d_reduce + 32;
  if (!h->self()) h->g[1].d_grad = h->d_reduce + 32 + 3 * s1;
  return B200CV_OK;
}

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

#define CUDA_OK_RC(expr)          \
  do {                            \
    int rc_ = (expr);             \
    if (rc_ != B200CV_OK) return rc_; \
  } while (0)

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

  CUDA_OK(h, cudaMemsetAsync(h->d_counters, 0, 16, stream));
  if (gated) {
    CUDA_OK(h, launch_fetch_gate(h->h_mode_dev, h->gate(), stream));
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
    F.gate = gated ? h->gate() : nullptr;
    F.pl_persist = h->pl_count() + 1;
    F.pl_cap = h->pl_cap;
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
    A.gate = gated ? h->gate() : nullptr;
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
                                 int &nparts) -> int {
      if (record) h->pl_sorted = false;  // the sweep and the dense kernel list (i, j) as they are
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
        D.gate_on = 1;
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
      S.gate_on = 1;
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
      E.gate_on = 1;
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
      h->launches += 5;
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
      C.sorted_entries = h->pl_sorted ? 1 : 0;
      C.gate = gate;
      C.P = h->P;
      CUDA_OK(h, launch_cell_pairs(C, stream));
      h->launches++;
      return B200CV_OK;
    };

    if (gated) {
      // captured evaluation with a pair list: the graph holds the rebuild sweep AND the list
      // walk; the mode word decides on the device which of the two does any work
      int nparts = 0;
      int rc;
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
    } else if (walk_list) {
      int rc = enqueue_walk(h->d_partials, nullptr);
      if (rc) return rc;
      npart = WALK_BLOCKS;
    } else if (rebuild_with_cells(h, flags)) {
      int rc = enqueue_cells(flags, nullptr);
      if (rc) return rc;
      npart = CELL_BLOCKS;
    } else {
      int rc = enqueue_all_pairs(flags, rebuild, nullptr, npart);
      if (rc) return rc;
    }
  }

  if (!finalized) {
    FinalizeArgs const F = finalize_args(npart);
    CUDA_OK(h, launch_finalize(F, stream));
    h->launches++;
  }

  if (h->world > 1) {
    if (h->comm) {
      ncclResult_t r = g_nccl.AllReduce(h->d_reduce, h->d_reduce, h->reduce_count, ncclDouble,
                                        ncclSum, h->comm, stream);
      if (r != ncclSuccess) {
        return fail(h, B200CV_ERROR,
                    std::string("ncclAllReduce: ") +
                        (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
      }
    }
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
    CUDA_OK(h, cudaMemcpyAsync(h->h_value, h->d_value, sizeof(double), cudaMemcpyDeviceToHost,
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
  // gated (captured) evaluations: finalize_kernel reports the mode the finished launch ran in
  bool const rebuild = h->gated ? (h->h_counters[2] == 1)
                                : (bool)(h->last_flags & B200CV_EF_REBUILD_PAIRLIST);
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
    if (pl_n > h->pl_cap && h->async_eval) {
      return fail(h, B200CV_MEMORY_ERROR,
                  "pair-list overflow in a captured evaluation (" + std::to_string(pl_n) +
                      " listed pairs, capacity " + std::to_string(h->pl_cap) +
                      "): reserve more with b200cv_pairlist_reserve before capturing");
    }
    if (pl_n > h->pl_cap) {
      int rc = ensure_pairlist_capacity(h, pl_n);
      if (rc) return rc;
      redo = true;
    } else {
      h->pl_n = pl_n;
      h->pl_valid = true;
    }
  }
  return B200CV_OK;
}

int sync_and_check(b200cv_cvc *h)
{
  for (int attempt = 0; attempt < 4; attempt++) {
    CUDA_OK(h, cudaStreamSynchronize(h->last_stream));
    bool redo = false;
    int rc = post_sync(h, redo);
    if (rc) return rc;
    if (redo && h->world > 1) {
      // ranks cannot re-run independently: their all-reduces would no longer pair up
      return fail(h, B200CV_MEMORY_ERROR,
                  "pair queue overflow in a sharded run; raise the pair-list capacity");
    }
    if (!redo) {
      bool const rebuild = !h->gated && (h->last_flags & B200CV_EF_REBUILD_PAIRLIST);
      if (rebuild && !(h->center1() || h->center2())) {
        // keep the list length on the device for the list-walk steps (gated evaluations
        // do that in finalize_kernel)
        CUDA_OK(h, cudaMemcpyAsync(h->pl_count() + 1, &h->h_counters[1], 8,
                                   cudaMemcpyHostToDevice, h->last_stream));
        CUDA_OK(h, cudaStreamSynchronize(h->last_stream));
      }
      return B200CV_OK;
    }
    rc = reset_gradients(h, h->last_stream);
    if (rc) return rc;
    rc = enqueue_calc(h, h->last_flags, h->last_stream);
    if (rc) return rc;
  }
  return fail(h, B200CV_BUG_ERROR, "queue capacity did not converge");
}

}  // namespace

// ======================================== C ABI =================================================

extern "C" {

void b200cv_config_defaults(b200cv_config *cfg)
{
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->kind = B200CV_COORDNUM;
  cfg->cutoff3[0] = cfg->cutoff3[1] = cfg->cutoff3[2] = 4.0;
  cfg->exp_numer = 6;
  cfg->exp_denom = 12;
  cfg->tolerance = 0.0;
  cfg->pairlist_freq = 100;
}

const char *b200cv_last_error(const b200cv_cvc *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int b200cv_create(const b200cv_config *cfg, b200cv_cvc **out)
{
  if (!cfg || !out) return fail(nullptr, B200CV_BUG_ERROR, "null argument");
  *out = nullptr;
  b200cv_config c = *cfg;
  if (c.kind < 0 || c.kind > 3) return fail(nullptr, B200CV_INPUT_ERROR, "unknown component kind");
  if (c.kind == B200CV_DISTANCEINV) {
    // distance_inv::init (src/colvarcomp_distances.cpp:384-391)
    if (c.exp_numer % 2) {
      return fail(nullptr, B200CV_INPUT_ERROR,
                  "Error: odd exponent provided, can only use even ones.");
    }
    if (c.exp_numer <= 0) {
      return fail(nullptr, B200CV_INPUT_ERROR, "Error: negative or zero exponent provided.");
    }
    c.exp_denom = 2 * c.exp_numer;
    c.cutoff3[0] = c.cutoff3[1] = c.cutoff3[2] = 1.0;
    c.tolerance = 0.0;
    c.group1_center_only = c.group2_center_only = 0;
  }
  if (c.n1 < 1) return fail(nullptr, B200CV_INPUT_ERROR, "Error: group1 is empty");
  if (c.kind != B200CV_SELFCOORDNUM && c.n2 < 1) {
    return fail(nullptr, B200CV_INPUT_ERROR, "Error: group2 is empty");
  }
  if ((c.exp_numer % 2) || (c.exp_denom % 2)) {
    return fail(nullptr, B200CV_INPUT_ERROR,
                "Error: odd exponent(s) provided, can only use even ones.");
  }
  if ((c.exp_numer <= 0) || (c.exp_denom <= 0)) {
    return fail(nullptr, B200CV_INPUT_ERROR, "Error: negative exponent(s) provided.");
  }
  for (int d = 0; d < 3; d++) {
    if (c.cutoff3[d] < 0.0) c.cutoff3[d] = -c.cutoff3[d];  // "remove meaningless negative signs"
    if (!(c.cutoff3[d] > 0.0)) return fail(nullptr, B200CV_INPUT_ERROR, "Error: zero cutoff.");
  }
  if (c.kind == B200CV_GROUPCOORD) {
    c.group1_center_only = c.group2_center_only = 1;
    c.tolerance = 0.0;  // groupCoord never uses a pair list (src/colvarcomp_coordnums.cpp:135)
  }
  if (c.kind == B200CV_SELFCOORDNUM) {
    c.group1_center_only = c.group2_center_only = 0;
    c.n2 = 0;
  }
  if (c.tolerance > 0 && !(c.pairlist_freq > 0)) {
    return fail(nullptr, B200CV_INPUT_ERROR, "Error: non-positive pairlistfrequency provided.");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    return fail(nullptr, B200CV_ERROR,
                "no CUDA device: the B200 coordination-number engine has no CPU fallback");
  }
  if (c.device < 0 || c.device >= ndev) {
    return fail(nullptr, B200CV_INPUT_ERROR, "invalid CUDA device ordinal");
  }
  b200cv_cvc *h = new b200cv_cvc();
  h->cfg = c;
  std::memset(&h->P, 0, sizeof(h->P));
  reset_boundaries(h->P);
  if (c.tolerance > 0) h->l2_max = compute_tolerance_l2_max(c.exp_numer, c.exp_denom, c.tolerance);
  fill_pair_params(h);

  auto bail = [&](int rc) {
    g_create_error = h->err;
    b200cv_destroy(h);
    return rc;
  };
#define CREATE_OK(call)                                                                       \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
      return bail(e_ == cudaErrorMemoryAllocation ? B200CV_MEMORY_ERROR : B200CV_ERROR);      \
    }                                                                                         \
  } while (0)
  CREATE_OK(cudaSetDevice(c.device));
  CREATE_OK(cudaHostAlloc(&h->h_value, sizeof(double), cudaHostAllocMapped));
  CREATE_OK(cudaHostGetDevicePointer(&h->h_value_dev, h->h_value, 0));
  CREATE_OK(cudaHostAlloc(&h->h_counters, 4 * sizeof(unsigned long long), cudaHostAllocMapped));
  CREATE_OK(cudaHostGetDevicePointer(&h->h_counters_dev, h->h_counters, 0));
  h->h_counters[0] = h->h_counters[1] = h->h_counters[2] = h->h_counters[3] = 0;
  *h->h_value = 0.0;
  CREATE_OK(cudaHostAlloc(&h->h_mode, sizeof(int), cudaHostAllocMapped));
  CREATE_OK(cudaHostGetDevicePointer(&h->h_mode_dev, h->h_mode, 0));
  *h->h_mode = 1;
  CREATE_OK(cudaMalloc(&h->d_counters, 32));
  CREATE_OK(cudaMemset(h->d_counters, 0, 32));
  CREATE_OK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
#undef CREATE_OK
  int rc = ensure_rare_capacity(h, 1);
  if (rc) return bail(rc);
  *out = h;
  return B200CV_OK;
}

int b200cv_destroy(b200cv_cvc *h)
{
  if (!h) return B200CV_OK;
  cudaSetDevice(h->cfg.device);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  for (auto &g : h->step_graphs) cudaGraphExecDestroy(g.exec);
  for (auto &G : h->g) {
    free_dev(G.d_index);
    free_dev(G.d_mass);
    free_dev(G.d_weight);
    free_dev(G.d_pos);
    free_dev(G.d_centers);
    free_dev(G.d_com_scratch);
    free_dev(G.d_com_ticket);
    fit_free(G.fit);
  }
  free_dev(h->d_reduce);
  free_dev(h->d_counters);
  free_dev(h->d_partials);
  free_dev(h->d_rare);
  free_dev(h->d_pl);
  free_dev(h->d_pl_center);
  free_dev(h->d_items);
  cells_free(h->cells);
  free_dev(h->d_stage_pos);
  if (h->h_value) cudaFreeHost(h->h_value);
  if (h->h_counters) cudaFreeHost(h->h_counters);
  if (h->h_mode) cudaFreeHost(h->h_mode);
  if (h->h_stage_grad) cudaFreeHost(h->h_stage_grad);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return B200CV_OK;
}

int b200cv_set_group(b200cv_cvc *h, int which, const int *proxy_index, const double *mass, int n)
{
  if (!h) return B200CV_BUG_ERROR;
  if (which != 1 && which != 2) return fail(h, B200CV_INPUT_ERROR, "group must be 1 or 2");
  if (which == 2 && h->self()) {
    return fail(h, B200CV_INPUT_ERROR, "selfCoordNum has no group2");
  }
  int const expect = which == 1 ? h->cfg.n1 : h->cfg.n2;
  if (n != expect) return fail(h, B200CV_INPUT_ERROR, "group size differs from the configuration");
  if (!proxy_index) return fail(h, B200CV_BUG_ERROR, "null index array");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  Group &G = h->g[which - 1];
  if (G.d_index) return fail(h, B200CV_INPUT_ERROR, "group already defined");
  G.h_index.assign(proxy_index, proxy_index + n);
  G.h_mass.resize(n);
  double total = 0.0;
  for (int i = 0; i < n; i++) {
    G.h_mass[i] = mass ? mass[i] : 1.0;
    total += G.h_mass[i];
    if (proxy_index[i] < 0) return fail(h, B200CV_INPUT_ERROR, "negative proxy index");
    G.max_index = std::max(G.max_index, proxy_index[i]);
  }
  {
    // an atom can be in a group only once (atom_group::add_atom_id; the host scatter and the
    // device gather rely on it)
    std::unordered_set<int> own;
    own.reserve((size_t)n * 2);
    for (int i = 0; i < n; i++) {
      if (!own.insert(proxy_index[i]).second) {
        G.h_index.clear();
        return fail(h, B200CV_INPUT_ERROR,
                    "Error: the same atom appears twice in a group (proxy index " +
                        std::to_string(proxy_index[i]) + ")");
      }
    }
  }
  // atom_group::overlap (src/colvaratoms.cpp:1096-1105), O(N) here
  Group &O = h->g[2 - which];
  if (!h->self() && !O.h_index.empty()) {
    std::unordered_set<int> seen(O.h_index.begin(), O.h_index.end());
    for (int i = 0; i < n; i++) {
      if (seen.count(proxy_index[i])) {
        G.h_index.clear();
        if (h->dinv()) {
          // distance_inv::init (src/colvarcomp_distances.cpp:393-397)
          return fail(h, B200CV_INPUT_ERROR,
                      "Error: group1 and group2 have some atoms in common: this is not "
                      "allowed for distanceInv.");
        }
        return fail(h, B200CV_INPUT_ERROR,
                    "Error: group1 and group2 share a common atom (number: " +
                        std::to_string(proxy_index[i] + 1) + ")");
      }
    }
  }
  G.total_mass = total;
  int rc = alloc_group(h, G, n);
  if (rc) return rc;
  std::vector<double> w(n);
  for (int i = 0; i < n; i++) w[i] = G.h_mass[i] / total;
  CUDA_OK(h, cudaMemcpy(G.d_index, proxy_index, sizeof(int) * n, cudaMemcpyHostToDevice));
  CUDA_OK(h, cudaMemcpy(G.d_mass, G.h_mass.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
  CUDA_OK(h, cudaMemcpy(G.d_weight, w.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
  h->items_dirty = true;
  if (groups_ready(h)) {
    rc = ensure_reduce_buffer(h);
    if (rc) return rc;
    if (h->cfg.tolerance > 0) {
      if (h->center1() || h->center2()) {
        int const m = h->center1() ? h->g[1].n : h->g[0].n;
        CUDA_OK(h, cudaMalloc(&h->d_pl_center, std::max(m, 1)));
        CUDA_OK(h, cudaMemset(h->d_pl_center, 1, std::max(m, 1)));  // list starts all-true
      } else {
        unsigned long long const n1 = h->g[0].n, n2 = h->self() ? n1 : h->g[1].n;
        unsigned long long const pairs = h->self() ? n1 * (n1 - 1) / 2 : n1 * n2;
        CUDA_OK(h, cells_setup(h->cells, (int)n2));
        h->cells_ready = true;
        rc = ensure_pairlist_capacity(h, std::min<unsigned long long>(pairs, 64ull * (n1 + n2)));
        if (rc) return rc;
      }
    }
  }
  return B200CV_OK;
}

int b200cv_set_boundaries(b200cv_cvc *h, int px, int py, int pz, const double A[3],
                          const double B[3], const double C[3])
{
  if (!h) return B200CV_BUG_ERROR;
  // system_boundary_conditions::set_boundaries (src/colvars_system.h:80-158)
  PairParams &P = h->P;
  PairParams const before = P;
  struct Bump {  // kernel parameters are frozen inside captured graphs
    b200cv_cvc *h;
    PairParams const &old;
    ~Bump() { if (std::memcmp(&old, &h->P, sizeof(PairParams)) != 0) h->epoch++; }
  } bump{h, before};
  reset_boundaries(P);
  P.periodic[0] = px ? 1 : 0;
  P.periodic[1] = py ? 1 : 0;
  P.periodic[2] = pz ? 1 : 0;
  if ((P.periodic[0] == P.periodic[1]) && (P.periodic[0] == P.periodic[2])) {
    if (P.periodic[0]) {
      P.bc_type = 2;
    } else {
      reset_boundaries(P);
      return B200CV_OK;
    }
  } else {
    P.bc_type = 1;
  }
  if (!A || !B || !C) return fail(h, B200CV_BUG_ERROR, "null cell vector");
  const double *V[3] = {A, B, C};
  double const diagonal_tol2 = 1.0e-10;
  bool off_diagonal = false;
  for (int d = 0; d < 3; d++) {
    if (!P.periodic[d]) continue;
    int const o1 = (d + 1) % 3, o2 = (d + 2) % 3;
    for (int k = 0; k < 3; k++) P.cell[d][k] = V[d][k];
    if ((V[d][o1] * V[d][o1]) > diagonal_tol2 || (V[d][o2] * V[d][o2]) > diagonal_tol2) {
      off_diagonal = true;
    } else {
      double const n2 = V[d][0] * V[d][0] + V[d][1] * V[d][1] + V[d][2] * V[d][2];
      for (int k = 0; k < 3; k++) P.recip[d][k] = V[d][k] / n2;
    }
  }
  if (P.bc_type == 2 && off_diagonal) P.bc_type = 3;
  if (P.bc_type == 3) {
    auto cross = [](const double a[3], const double b[3], double c[3]) {
      c[0] = a[1] * b[2] - b[1] * a[2];
      c[1] = -a[0] * b[2] + b[0] * a[2];
      c[2] = a[0] * b[1] - b[0] * a[1];
    };
    auto dot = [](const double a[3], const double b[3]) {
      return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    };
    double v[3];
    cross(P.cell[1], P.cell[2], v);
    double s = dot(v, P.cell[0]);
    for (int k = 0; k < 3; k++) P.recip[0][k] = v[k] / s;
    cross(P.cell[2], P.cell[0], v);
    s = dot(v, P.cell[1]);
    for (int k = 0; k < 3; k++) P.recip[1][k] = v[k] / s;
    cross(P.cell[0], P.cell[1], v);
    s = dot(v, P.cell[2]);
    for (int k = 0; k < 3; k++) P.recip[2][k] = v[k] / s;
  }
  if (P.bc_type == 2) {
    // the tile sweep's orthogonal path reads only the diagonal; the reference also adds the
    // (<= 1e-5) off-diagonal terms it tolerates -- route those cells to the general path
    bool exact_diag = true;
    for (int d = 0; d < 3; d++) {
      for (int k = 0; k < 3; k++) {
        if (k != d && P.cell[d][k] != 0.0) exact_diag = false;
      }
    }
    if (!exact_diag) P.bc_type = 3;
  }
  if (P.bc_type == 1 || P.bc_type == 3) fill_general_cell_tables(P);
  return B200CV_OK;
}

int b200cv_set_fitting(b200cv_cvc *h, int which, int center, int rotate,
                       const int *fit_proxy_index, int nfit, const double *ref_pos_soa,
                       int enable_fit_gradients)
{
  if (!h) return B200CV_BUG_ERROR;
  if (which != 1 && which != 2) return fail(h, B200CV_INPUT_ERROR, "group must be 1 or 2");
  Group &G = h->g[which - 1];
  if (!G.d_index) return fail(h, B200CV_INPUT_ERROR, "define the group before its fitting options");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  std::string err;
  h->epoch++;
  int rc = fit_setup(G.fit, center, rotate, fit_proxy_index, nfit, G.n, G.stride, G.d_index,
                     G.h_index, ref_pos_soa, enable_fit_gradients, err);
  if (rc) return fail(h, rc, err);
  return B200CV_OK;
}

int b200cv_read_data(b200cv_cvc *h, const double *d_proxy_pos, size_t proxy_stride, void *stream)
{
  NvtxRange nvtx_range("b200cv::atom_group_read_data");
  if (!h) return B200CV_BUG_ERROR;
  if (!groups_ready(h)) return fail(h, B200CV_INPUT_ERROR, "atom groups are not defined");
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  // reset_atoms_data (src/colvaratoms.cpp:1731-1736): gradients (and the value) to zero
  int rc = reset_gradients(h, s);
  if (rc) return rc;
  int const ng = h->self() ? 1 : 2;
  bool const one_launch = ng == 2 && !h->g[0].fit.active && !h->g[1].fit.active;
  if (one_launch) {
    GroupView const a{h->g[0].d_index, h->g[0].n, h->g[0].d_pos, h->g[0].stride, nullptr};
    GroupView const b{h->g[1].d_index, h->g[1].n, h->g[1].d_pos, h->g[1].stride, nullptr};
    CUDA_OK(h, launch_gather2(d_proxy_pos, proxy_stride, 0, a, b, s));
  }
  for (int k = 0; k < ng; k++) {
    Group &G = h->g[k];
    if (!one_launch) {
      CUDA_OK(h, launch_gather(d_proxy_pos, proxy_stride, 0, G.d_index, G.n, G.d_pos, G.stride,
                               s));
    }
    std::string err;
    rc = fit_read_and_apply(G.fit, d_proxy_pos, proxy_stride, 0, G.d_pos, G.stride, G.n, s, err);
    if (rc) return fail(h, rc, err);
    // calc_required_properties (src/colvaratoms.cpp:1325-1352)
    CUDA_OK_RC(centers_after_read(h, G, s));
  }
  h->data_read = true;
  return B200CV_OK;
}

int b200cv_calc_value(b200cv_cvc *h, long long step_relative, int gradients, void *stream)
{
  NvtxRange nvtx_range("b200cv::cvc_calc_value");
  if (!h) return B200CV_BUG_ERROR;
  if (!h->data_read) return fail(h, B200CV_BUG_ERROR, "calc_value before read_data");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  int const flags = calc_flags(h, step_relative, gradients);
  h->async_eval = false;
  h->gated = false;
  h->last_stream = (cudaStream_t)stream;
  h->last_step = step_relative;
  h->last_flags = flags;
  h->have_calc = true;
  return enqueue_calc(h, flags, h->last_stream);
}

int b200cv_step(b200cv_cvc *h, const double *d_proxy_pos, size_t proxy_stride,
                long long step_relative, int gradients, void *stream)
{
  if (!h) return B200CV_BUG_ERROR;
  if (!groups_ready(h)) return fail(h, B200CV_INPUT_ERROR, "atom groups are not defined");
  NvtxRange nvtx_range("b200cv::step");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t user = (cudaStream_t)stream;
  int const flags = calc_flags(h, step_relative, gradients);
  h->gated = false;  // this entry keeps one graph per flag set instead of gating
  bool const fitted = h->g[0].fit.active || h->g[1].fit.active;
  auto plain = [&]() -> int {
    int rc = b200cv_read_data(h, d_proxy_pos, proxy_stride, stream);
    if (rc == B200CV_OK) rc = b200cv_calc_value(h, step_relative, gradients, stream);
    if (rc == B200CV_OK && fitted && gradients) rc = b200cv_calc_fit_gradients(h, stream);
    return rc;
  };
  // not captured: sharded runs (the collective), a pair list that has not been built yet
  bool const use_pl = flags & B200CV_EF_USE_PAIRLIST;
  bool const rebuild = flags & B200CV_EF_REBUILD_PAIRLIST;
  if (h->world > 1 || (use_pl && !rebuild && !h->pl_valid)) return plain();

  // everything that allocates or copies synchronously has to happen before the capture
  if (!(h->center1() || h->center2())) {
    if (h->items_dirty) {
      int rc = build_items(h);
      if (rc) return rc;
    }
  } else {
    int rc = ensure_partials(h, PARTIALS_TAIL);
    if (rc) return rc;
  }
  // a rebuild is captured in two forms: all-pairs sweep (first rebuild) and cell list
  // (the host does not run enqueue_calc on a replay: what it would have noted -- the index
  // space of the list a rebuild leaves behind -- is noted below), a list walk in two as well:
  // entries as atom indices or in cell order
  bool const cells_rebuild = rebuild_with_cells(h, flags);
  int graph_key = flags | (cells_rebuild ? (1 << 30) : 0);
  if (use_pl && !rebuild && h->pl_sorted) graph_key |= (1 << 29);
  b200cv_cvc::StepGraph *hit = nullptr;
  for (auto &g : h->step_graphs) {
    if (g.flags == graph_key && g.pos == d_proxy_pos && g.stride == proxy_stride &&
        g.epoch == h->epoch) {
      hit = &g;
    }
  }
  if (!hit) {
    // drop graphs that can never match again
    for (size_t k = 0; k < h->step_graphs.size();) {
      if (h->step_graphs[k].epoch != h->epoch) {
        cudaGraphExecDestroy(h->step_graphs[k].exec);
        h->step_graphs.erase(h->step_graphs.begin() + k);
      } else {
        k++;
      }
    }
    cudaStream_t cap = h->own_stream;  // the caller's stream may be the legacy default stream
    CUDA_OK(h, cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed));
    int rc = b200cv_read_data(h, d_proxy_pos, proxy_stride, cap);
    if (rc == B200CV_OK) rc = enqueue_calc(h, flags, cap);
    if (rc == B200CV_OK && fitted && gradients) rc = b200cv_calc_fit_gradients(h, cap);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(cap, &graph);
    if (rc != B200CV_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    CUDA_OK(h, e);
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    CUDA_OK(h, e);
    size_t nodes = 0;
    h->step_graphs.push_back({graph_key, d_proxy_pos, proxy_stride, h->epoch, h->launches, exec});
    hit = &h->step_graphs.back();
    (void)nodes;
  }
  CUDA_OK(h, cudaGraphLaunch(hit->exec, user));
  if (use_pl && rebuild && !(h->center1() || h->center2())) {
    h->pl_sorted = cells_rebuild && sorted_walk_enabled();
  }
  h->launches = hit->launches;
  h->async_eval = false;
  h->gated = false;
  h->last_stream = user;
  h->last_step = step_relative;
  h->last_flags = flags;
  h->have_calc = true;
  h->data_read = true;
  return B200CV_OK;
}

int b200cv_calc_fit_gradients(b200cv_cvc *h, void *stream)
{
  NvtxRange nvtx_range("b200cv::atom_group_calc_fit_gradients");
  if (!h) return B200CV_BUG_ERROR;
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  int const ng = h->self() ? 1 : 2;
  for (int k = 0; k < ng; k++) {
    Group &G = h->g[k];
    std::string err;
    int rc = fit_calc_gradients(G.fit, G.d_grad, G.stride, G.n, (cudaStream_t)stream, err);
    if (rc) return fail(h, rc, err);
  }
  return B200CV_OK;
}

int b200cv_value(b200cv_cvc *h, double *value)
{
  NvtxRange nvtx_range("b200cv::cvc_calc_value_after_gpu");
  if (!h || !value) return B200CV_BUG_ERROR;
  if (!h->have_calc) return fail(h, B200CV_BUG_ERROR, "value requested before calc_value");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  int rc = sync_and_check(h);
  if (rc) return rc;
  *value = *h->h_value;
  return B200CV_OK;
}

int b200cv_apply_force(b200cv_cvc *h, double force, double *d_proxy_force, size_t proxy_stride,
                       void *stream)
{
  NvtxRange nvtx_range("b200cv::apply_forces");
  if (!h) return B200CV_BUG_ERROR;
  if (!h->have_calc) return fail(h, B200CV_BUG_ERROR, "apply_force before calc_value");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  int const ng = h->self() ? 1 : 2;
  bool const one_launch = ng == 2;
  if (one_launch) {
    Group &A = h->g[0], &B = h->g[1];
    GroupView const a{A.d_index, A.n, A.d_grad, A.stride, A.fit.rotate ? A.fit.d_q : nullptr};
    GroupView const b{B.d_index, B.n, B.d_grad, B.stride, B.fit.rotate ? B.fit.d_q : nullptr};
    CUDA_OK(h, launch_scatter_force2(a, b, force, d_proxy_force, proxy_stride, 0, s));
  }
  for (int k = 0; k < ng; k++) {
    Group &G = h->g[k];
    if (!one_launch) {
      CUDA_OK(h, launch_scatter_force(G.d_grad, G.stride, G.d_index, G.n, force,
                                      G.fit.rotate ? G.fit.d_q : nullptr, d_proxy_force,
                                      proxy_stride, 0, s));
    }
    std::string err;
    int rc = fit_scatter_force(G.fit, force, d_proxy_force, proxy_stride, 0, s, err);
    if (rc) return fail(h, rc, err);
  }
  return B200CV_OK;
}

static int eval_device(b200cv_cvc *h, const double *d_pos1, const double *d_pos2, int flags,
                       double *d_grad1, double *d_grad2, void *stream, bool sync)
{
  if (!h) return B200CV_BUG_ERROR;
  if (!groups_ready(h)) return fail(h, B200CV_INPUT_ERROR, "atom groups are not defined");
  if (!d_pos1 || (!h->self() && !d_pos2)) return fail(h, B200CV_BUG_ERROR, "null positions");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  int rc = reset_gradients(h, s);
  if (rc) return rc;
  Group &G1 = h->g[0];
  CUDA_OK(h, launch_repack(d_pos1, (size_t)G1.n, G1.d_pos, G1.stride, G1.n, 0, s));
  CUDA_OK_RC(centers_after_read(h, G1, s));
  if (!h->self()) {
    Group &G2 = h->g[1];
    CUDA_OK(h, launch_repack(d_pos2, (size_t)G2.n, G2.d_pos, G2.stride, G2.n, 0, s));
    CUDA_OK_RC(centers_after_read(h, G2, s));
  }
  if (!(h->cfg.tolerance > 0) || h->cfg.kind == B200CV_GROUPCOORD) {
    flags &= ~(B200CV_EF_USE_PAIRLIST | B200CV_EF_REBUILD_PAIRLIST);
  }
  h->last_stream = s;
  h->last_flags = flags;
  h->have_calc = true;
  h->data_read = true;
  h->async_eval = !sync;
  h->gated = !sync && (flags & B200CV_EF_USE_PAIRLIST);
  rc = enqueue_calc(h, flags, s);
  if (rc) return rc;
  if (sync) {
    // queue overflow has to be resolved before the gradients leave our buffers
    rc = sync_and_check(h);
    if (rc) return rc;
  }
  if (flags & B200CV_EF_GRADIENTS) {
    if (d_grad1) CUDA_OK(h, launch_repack(G1.d_grad, G1.stride, d_grad1, (size_t)G1.n, G1.n, 1, s));
    if (!h->self() && d_grad2) {
      Group &G2 = h->g[1];
      CUDA_OK(h, launch_repack(G2.d_grad, G2.stride, d_grad2, (size_t)G2.n, G2.n, 1, s));
    }
  }
  return B200CV_OK;
}

int b200cv_eval(b200cv_cvc *h, const double *d_pos1, const double *d_pos2, int flags,
                double *d_grad1, double *d_grad2, void *stream)
{
  return eval_device(h, d_pos1, d_pos2, flags, d_grad1, d_grad2, stream, true);
}

int b200cv_eval_async(b200cv_cvc *h, const double *d_pos1, const double *d_pos2, int flags,
                      double *d_grad1, double *d_grad2, void *stream)
{
  if (h && (flags & B200CV_EF_USE_PAIRLIST) && h->cfg.tolerance > 0 &&
      h->cfg.kind != B200CV_GROUPCOORD) {
    // gated form: rebuild sweep and list walk are both enqueued, b200cv_pairlist_set_rebuild
    // picks one per launch.  The list (allocated by b200cv_set_group) cannot grow under a
    // captured graph: size it with b200cv_pairlist_reserve first.
    if (!groups_ready(h)) return fail(h, B200CV_INPUT_ERROR, "atom groups are not defined");
  }
  if (h && h->items_dirty) {
    // allocations are not capturable: build the work-item table before any stream capture
    CUDA_OK(h, cudaSetDevice(h->cfg.device));
    if (!(h->center1() || h->center2())) {
      int rc = build_items(h);
      if (rc) return rc;
    } else {
      int rc = ensure_partials(h, PARTIALS_TAIL);
      if (rc) return rc;
    }
  }
  return eval_device(h, d_pos1, d_pos2, flags, d_grad1, d_grad2, stream, false);
}

int b200cv_pairlist_set_rebuild(b200cv_cvc *h, int rebuild)
{
  if (!h) return B200CV_BUG_ERROR;
  *h->h_mode = rebuild ? 1 : 0;  // mapped pinned: read by fetch_gate_kernel of the next launch
  return B200CV_OK;
}

int b200cv_pairlist_reserve(b200cv_cvc *h, const double *d_pos1, const double *d_pos2,
                            double headroom, void *stream)
{
  if (!h) return B200CV_BUG_ERROR;
  if (!(h->cfg.tolerance > 0) || h->cfg.kind == B200CV_GROUPCOORD || h->center1() ||
      h->center2()) {
    // no compacted list to size (dense flags of fixed length, or none); still make sure
    // nothing is left to allocate inside a capture
    CUDA_OK(h, cudaSetDevice(h->cfg.device));
    if (h->center1() || h->center2()) return ensure_partials(h, PARTIALS_TAIL);
    if (h->items_dirty) return build_items(h);
    return B200CV_OK;
  }
  // one synchronous rebuild sweep on the current positions (value only): grows the list until
  // it fits, then adds the headroom
  int rc = eval_device(h, d_pos1, d_pos2,
                       B200CV_EF_USE_PAIRLIST | B200CV_EF_REBUILD_PAIRLIST, nullptr, nullptr,
                       stream, true);
  if (rc) return rc;
  if (!(headroom >= 1.0)) headroom = 1.0;
  unsigned long long const n1 = h->g[0].n, n2 = h->self() ? n1 : h->g[1].n;
  unsigned long long const pairs = h->self() ? n1 * (n1 - 1) / 2 : n1 * n2;
  unsigned long long want = (unsigned long long)std::ceil(headroom * (double)h->pl_n) + 1024;
  want = std::min(want, pairs);
  if (want > h->pl_cap) {
    // ensure_pairlist_capacity over-allocates by 1.5x: ask for exactly `want`
    free_dev(h->d_pl);
    h->epoch++;
    CUDA_OK(h, cudaMalloc(&h->d_pl, sizeof(unsigned long long) * want));
    h->pl_cap = want;
  }
  h->pl_valid = false;  // the first captured launch rebuilds (mode word starts at 1);
                        // pl_n stays: it tells the capture whether the list is sparse
  *h->h_mode = 1;
  return B200CV_OK;
}

int b200cv_accumulate_gradients(b200cv_cvc *h, double *d_grad1, double *d_grad2, void *stream)
{
  if (!h) return B200CV_BUG_ERROR;
  if (!h->have_calc) return fail(h, B200CV_BUG_ERROR, "accumulate_gradients before an evaluation");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  Group &G1 = h->g[0];
  if (d_grad1) CUDA_OK(h, launch_repack(G1.d_grad, G1.stride, d_grad1, (size_t)G1.n, G1.n, 1, s));
  if (!h->self() && d_grad2) {
    Group &G2 = h->g[1];
    CUDA_OK(h, launch_repack(G2.d_grad, G2.stride, d_grad2, (size_t)G2.n, G2.n, 1, s));
  }
  return B200CV_OK;
}

int b200cv_eval_host(b200cv_cvc *h, const double *h_pos1, const double *h_pos2, int flags,
                     double *h_grad1, double *h_grad2, double *value)
{
  if (!h || !value) return B200CV_BUG_ERROR;
  if (!groups_ready(h)) return fail(h, B200CV_INPUT_ERROR, "atom groups are not defined");
  if (!h_pos1 || (!h->self() && !h_pos2)) return fail(h, B200CV_BUG_ERROR, "null positions");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = h->own_stream;
  int const ng = h->self() ? 1 : 2;
  const double *hp[2] = {h_pos1, h_pos2};
  double *hg[2] = {h_grad1, h_grad2};
  int rc = reset_gradients(h, s);
  if (rc) return rc;
  for (int k = 0; k < ng; k++) {
    Group &G = h->g[k];
    // SoA planes of n atoms -> padded planes of the device buffers
    CUDA_OK(h, cudaMemcpy2DAsync(G.d_pos, sizeof(double) * G.stride, hp[k],
                                 sizeof(double) * G.n, sizeof(double) * G.n, 3,
                                 cudaMemcpyHostToDevice, s));
    CUDA_OK_RC(centers_after_read(h, G, s));
  }
  if (!(h->cfg.tolerance > 0) || h->cfg.kind == B200CV_GROUPCOORD) {
    flags &= ~(B200CV_EF_USE_PAIRLIST | B200CV_EF_REBUILD_PAIRLIST);
  }
  h->async_eval = false;
  h->gated = false;
  h->last_stream = s;
  h->last_flags = flags;
  h->have_calc = true;
  h->data_read = true;
  rc = enqueue_calc(h, flags, s);
  if (rc) return rc;
  rc = sync_and_check(h);
  if (rc) return rc;
  *value = *h->h_value;
  if (flags & B200CV_EF_GRADIENTS) {
    if (h->h_stage_count < h->reduce_count) {
      if (h->h_stage_grad) cudaFreeHost(h->h_stage_grad);
      h->h_stage_grad = nullptr;
      CUDA_OK(h, cudaHostAlloc(&h->h_stage_grad, sizeof(double) * h->reduce_count,
                               cudaHostAllocDefault));
      h->h_stage_count = h->reduce_count;
    }
    CUDA_OK(h, cudaMemcpyAsync(h->h_stage_grad, h->d_reduce, sizeof(double) * h->reduce_count,
                               cudaMemcpyDeviceToHost, s));
    CUDA_OK(h, cudaStreamSynchronize(s));
    for (int k = 0; k < ng; k++) {
      if (!hg[k]) continue;
      Group &G = h->g[k];
      const double *g = h->h_stage_grad + (G.d_grad - h->d_reduce);
      for (int d = 0; d < 3; d++) {
        for (int i = 0; i < G.n; i++) hg[k][(size_t)d * G.n + i] += g[(size_t)d * G.stride + i];
      }
    }
  }
  return B200CV_OK;
}

// ---- host-buffer step --------------------------------------------------------------------------

int b200cv_calc_host(b200cv_cvc *h, const double *h_proxy_pos, size_t n_proxy,
                     long long step_relative, int gradients, double *value)
{
  if (!h || !h_proxy_pos || !value) return B200CV_BUG_ERROR;
  if (!groups_ready(h)) return fail(h, B200CV_INPUT_ERROR, "atom groups are not defined");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = h->own_stream;
  if (h->stage_n < n_proxy) {
    free_dev(h->d_stage_pos);
    CUDA_OK(h, cudaMalloc(&h->d_stage_pos, sizeof(double) * 3 * n_proxy));
    h->stage_n = n_proxy;
  }
  // the CPU proxy's AoS rvector array (src/colvarproxy.h:275) goes up in one copy; the
  // gather by atoms_index runs on the device
  CUDA_OK(h, cudaMemcpyAsync(h->d_stage_pos, h_proxy_pos, sizeof(double) * 3 * n_proxy,
                             cudaMemcpyHostToDevice, s));
  int rc = reset_gradients(h, s);
  if (rc) return rc;
  int const ng = h->self() ? 1 : 2;
  for (int k = 0; k < ng; k++) {
    Group &G = h->g[k];
    CUDA_OK(h, launch_gather(h->d_stage_pos, n_proxy, 1, G.d_index, G.n, G.d_pos, G.stride, s));
    std::string err;
    rc = fit_read_and_apply(G.fit, h->d_stage_pos, n_proxy, 1, G.d_pos, G.stride, G.n, s, err);
    if (rc) return fail(h, rc, err);
    CUDA_OK_RC(centers_after_read(h, G, s));
  }
  h->data_read = true;
  rc = b200cv_calc_value(h, step_relative, gradients, s);
  if (rc) return rc;
  if (gradients) {
    rc = b200cv_calc_fit_gradients(h, s);
    if (rc) return rc;
  }
  return b200cv_value(h, value);
}

int b200cv_apply_force_host(b200cv_cvc *h, double force, double *h_proxy_force, size_t n_proxy)
{
  if (!h || !h_proxy_force) return B200CV_BUG_ERROR;
  if (!h->have_calc) return fail(h, B200CV_BUG_ERROR, "apply_force before calc_value");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = h->last_stream;
  // one D2H copy of [grad1 | grad2]; the scatter into the AoS force array is the same
  // host loop as atom_group::apply_colvar_force (src/colvaratoms.cpp:1637-1704)
  if (h->h_stage_count < h->reduce_count) {
    if (h->h_stage_grad) cudaFreeHost(h->h_stage_grad);
    h->h_stage_grad = nullptr;
    CUDA_OK(h, cudaHostAlloc(&h->h_stage_grad, sizeof(double) * h->reduce_count,
                             cudaHostAllocDefault));
    h->h_stage_count = h->reduce_count;
  }
  CUDA_OK(h, cudaMemcpyAsync(h->h_stage_grad, h->d_reduce, sizeof(double) * h->reduce_count,
                             cudaMemcpyDeviceToHost, s));
  CUDA_OK(h, cudaStreamSynchronize(s));
  int const ng = h->self() ? 1 : 2;
  for (int k = 0; k < ng; k++) {
    Group &G = h->g[k];
    const double *g = h->h_stage_grad + (G.d_grad - h->d_reduce);
    double Rinv[3][3];
    bool const rot = G.fit.rotate;
    if (rot) {
      std::string err;
      int rc = fit_host_inverse_matrix(G.fit, Rinv, err);
      if (rc) return fail(h, rc, err);
    }
    if (G.max_index >= 0 && (size_t)G.max_index >= n_proxy) {
      return fail(h, B200CV_BUG_ERROR, "proxy index out of range");
    }
    const int *idx = G.h_index.data();
    size_t const stride = G.stride;
    int const n = G.n;
    // distinct atoms of one group: the rows are independent, large groups are split over the
    // host threads (the reference's loop is sequential, src/colvaratoms.cpp:1637-1704)
#pragma omp parallel for schedule(static) if (n > 16384)
    for (int i = 0; i < n; i++) {
      size_t const p = (size_t)idx[i];
      double gx = g[i], gy = g[stride + i], gz = g[2 * stride + i];
      if (rot) {
        double const tx = Rinv[0][0] * gx + Rinv[0][1] * gy + Rinv[0][2] * gz;
        double const ty = Rinv[1][0] * gx + Rinv[1][1] * gy + Rinv[1][2] * gz;
        double const tz = Rinv[2][0] * gx + Rinv[2][1] * gy + Rinv[2][2] * gz;
        gx = tx; gy = ty; gz = tz;
      }
      h_proxy_force[3 * p] += force * gx;
      h_proxy_force[3 * p + 1] += force * gy;
      h_proxy_force[3 * p + 2] += force * gz;
    }
    std::string err;
    int rc = fit_scatter_force_host(G.fit, force, h_proxy_force, n_proxy, s, err);
    if (rc) return fail(h, rc, err);
  }
  return B200CV_OK;
}

// ---- inspection ----------------------------------------------------------------------------------

static int copy_planes(b200cv_cvc *h, const double *d, size_t stride, int n, double *out)
{
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  if (h->last_stream) CUDA_OK(h, cudaStreamSynchronize(h->last_stream));
  CUDA_OK(h, cudaDeviceSynchronize());
  CUDA_OK(h, cudaMemcpy2D(out, sizeof(double) * n, d, sizeof(double) * stride, sizeof(double) * n,
                          3, cudaMemcpyDeviceToHost));
  return B200CV_OK;
}

int b200cv_get_positions(b200cv_cvc *h, int which, double *out)
{
  if (!h || !out || which < 1 || which > 2) return B200CV_BUG_ERROR;
  Group &G = h->g[which - 1];
  if (!G.d_pos) return fail(h, B200CV_INPUT_ERROR, "group not defined");
  return copy_planes(h, G.d_pos, G.stride, G.n, out);
}

int b200cv_get_gradients(b200cv_cvc *h, int which, double *out)
{
  if (!h || !out || which < 1 || which > 2) return B200CV_BUG_ERROR;
  Group &G = h->g[which - 1];
  if (!G.d_grad) return fail(h, B200CV_INPUT_ERROR, "group not defined");
  if (h->have_calc) {
    int rc = sync_and_check(h);
    if (rc) return rc;
  }
  return copy_planes(h, G.d_grad, G.stride, G.n, out);
}

int b200cv_get_fit_gradients(b200cv_cvc *h, int which, double *out, int nfit)
{
  if (!h || !out || which < 1 || which > 2) return B200CV_BUG_ERROR;
  Group &G = h->g[which - 1];
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  CUDA_OK(h, cudaDeviceSynchronize());
  std::string err;
  int rc = fit_get_gradients(G.fit, out, nfit, err);
  if (rc) return fail(h, rc, err);
  return B200CV_OK;
}

int b200cv_get_centers(b200cv_cvc *h, int which, double com[3], double cog[3])
{
  if (!h || which < 1 || which > 2) return B200CV_BUG_ERROR;
  Group &G = h->g[which - 1];
  if (!G.d_centers) return fail(h, B200CV_INPUT_ERROR, "group not defined");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  CUDA_OK(h, cudaDeviceSynchronize());
  if (!centers_in_step(h, G)) {
    // not part of the step for this group: reduce the positions it currently holds
    if (!h->data_read) return fail(h, B200CV_BUG_ERROR, "centers requested before read_data");
    int rc = centers_now(h, G, h->own_stream);
    if (rc) return rc;
    CUDA_OK(h, cudaStreamSynchronize(h->own_stream));
  }
  double c[6];
  CUDA_OK(h, cudaMemcpy(c, G.d_centers, sizeof(c), cudaMemcpyDeviceToHost));
  if (com) std::memcpy(com, c, 3 * sizeof(double));
  if (cog) std::memcpy(cog, c + 3, 3 * sizeof(double));
  return B200CV_OK;
}

int b200cv_get_rotation(b200cv_cvc *h, int which, double q[4])
{
  if (!h || !q || which < 1 || which > 2) return B200CV_BUG_ERROR;
  Group &G = h->g[which - 1];
  if (!G.fit.rotate) return fail(h, B200CV_INPUT_ERROR, "group is not rotated");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  CUDA_OK(h, cudaDeviceSynchronize());
  CUDA_OK(h, cudaMemcpy(q, G.fit.d_q, 4 * sizeof(double), cudaMemcpyDeviceToHost));
  return B200CV_OK;
}

int b200cv_get_tolerance_l2_max(const b200cv_cvc *h, double *l2_max)
{
  if (!h || !l2_max) return B200CV_BUG_ERROR;
  *l2_max = h->l2_max;
  return B200CV_OK;
}

int b200cv_pairlist_count(b200cv_cvc *h, size_t *count)
{
  if (!h || !count) return B200CV_BUG_ERROR;
  if (!(h->cfg.tolerance > 0)) return fail(h, B200CV_INPUT_ERROR, "no pair list (tolerance = 0)");
  if (h->have_calc) {
    int rc = sync_and_check(h);
    if (rc) return rc;
  }
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  if (h->center1() || h->center2()) {
    int const m = h->center1() ? h->g[1].n : h->g[0].n;
    std::vector<unsigned char> f(m);
    CUDA_OK(h, cudaMemcpy(f.data(), h->d_pl_center, m, cudaMemcpyDeviceToHost));
    size_t c = 0;
    for (unsigned char v : f) c += v ? 1 : 0;
    *count = c;
    return B200CV_OK;
  }
  if (!h->pl_valid) {
    // before the first rebuild the reference's list is all-true
    unsigned long long const n1 = h->g[0].n, n2 = h->self() ? n1 : h->g[1].n;
    *count = (size_t)(h->self() ? n1 * (n1 - 1) / 2 : n1 * n2);
    return B200CV_OK;
  }
  *count = (size_t)h->pl_n;
  return B200CV_OK;
}

int b200cv_pairlist_export(b200cv_cvc *h, uint64_t *out, size_t capacity)
{
  if (!h || !out) return B200CV_BUG_ERROR;
  size_t count = 0;
  int rc = b200cv_pairlist_count(h, &count);
  if (rc) return rc;
  if (capacity < count) return fail(h, B200CV_MEMORY_ERROR, "pair-list export buffer too small");
  unsigned long long const n1 = h->g[0].n, n2 = h->self() ? n1 : h->g[1].n;
  if (h->center1() || h->center2()) {
    int const m = h->center1() ? h->g[1].n : h->g[0].n;
    std::vector<unsigned char> f(m);
    CUDA_OK(h, cudaMemcpy(f.data(), h->d_pl_center, m, cudaMemcpyDeviceToHost));
    size_t k = 0;
    for (int a = 0; a < m; a++) {
      if (f[a]) out[k++] = (uint64_t)a;
    }
    return B200CV_OK;
  }
  if (!h->pl_valid) {
    for (size_t k = 0; k < count; k++) out[k] = k;
    return B200CV_OK;
  }
  std::vector<unsigned long long> ent(count);
  CUDA_OK(h, cudaMemcpy(ent.data(), h->d_pl, sizeof(unsigned long long) * count,
                        cudaMemcpyDeviceToHost));
  std::vector<int> order;
  if (h->pl_sorted) {
    // entries in cell order: back to atom indices through the order of the last rebuild
    order.resize((size_t)n2);
    CUDA_OK(h, cudaMemcpy(order.data(), h->cells.d_sorted, sizeof(int) * (size_t)n2,
                          cudaMemcpyDeviceToHost));
  }
  for (size_t k = 0; k < count; k++) {
    unsigned long long i = ent[k] >> 32, j = ent[k] & 0xffffffffull;
    if (h->pl_sorted) {
      j = (unsigned long long)order[(size_t)j];
      if (h->self()) {
        i = (unsigned long long)order[(size_t)i];
        if (i > j) std::swap(i, j);  // listed by the row that comes first in cell order
      }
    }
    // canonical index of the reference's dense bool array
    // (src/colvarcomp_coordnums.cpp:197-238 / :481-520)
    out[k] = h->self() ? i * (2 * n1 - i - 1) / 2 + (j - i - 1) : i * n2 + j;
  }
  std::sort(out, out + count);
  return B200CV_OK;
}

int b200cv_last_stats(b200cv_cvc *h, int *kernel_launches, size_t *exact_pairs)
{
  if (!h) return B200CV_BUG_ERROR;
  if (kernel_launches) *kernel_launches = h->launches;
  if (exact_pairs) *exact_pairs = h->exact_pairs;
  return B200CV_OK;
}

unsigned long long b200cv_kernel_launches(void) { return launches_so_far(); }

#if defined(B200CV_TRACE)
// debug builds only: {start, tiles begin, tiles end, end (ns), smid} per CTA of the last sweep
int b200cv_trace_dump(unsigned long long *out, int n) { return b200cv::trace_dump(out, n); }
#endif

// ---- multi-GPU -----------------------------------------------------------------------------------

int b200cv_comm_unique_id(unsigned char id[128])
{
  std::string err;
  if (!g_nccl.load(err)) return fail(nullptr, B200CV_ERROR, err);
  ncclUniqueId uid;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclResult_t r = g_nccl.GetUniqueId(&uid);
  if (r != ncclSuccess) return fail(nullptr, B200CV_ERROR, "ncclGetUniqueId failed");
  std::memcpy(id, &uid, 128);
  return B200CV_OK;
}

int b200cv_set_shard(b200cv_cvc *h, int rank, int world)
{
  if (!h) return B200CV_BUG_ERROR;
  if (world < 1 || rank < 0 || rank >= world) return fail(h, B200CV_INPUT_ERROR, "bad rank/world");
  h->rank = rank;
  h->world = world;
  h->items_dirty = true;
  h->pl_valid = false;
  return B200CV_OK;
}

int b200cv_comm_init(b200cv_cvc *h, const unsigned char id[128], int rank, int world)
{
  if (!h || !id) return B200CV_BUG_ERROR;
  int rc = b200cv_set_shard(h, rank, world);
  if (rc) return rc;
  std::string err;
  if (!g_nccl.load(err)) return fail(h, B200CV_ERROR, err);
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  ncclUniqueId uid;
  std::memcpy(&uid, id, 128);
  ncclResult_t r = g_nccl.CommInitRank(&h->comm, world, uid, rank);
  if (r != ncclSuccess) {
    return fail(h, B200CV_ERROR,
                std::string("ncclCommInitRank: ") +
                    (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
  }
  return B200CV_OK;
}

int b200cv_plan_items(int kind, int n1, int n2, int rank, int world, int *items, size_t capacity,
                      size_t *count, int *chunk)
{
  if (!count || world < 1 || rank < 0 || rank >= world || n1 < 1) return B200CV_INPUT_ERROR;
  std::vector<SweepItem> mine;
  int ch = 0;
  plan_items(kind == B200CV_SELFCOORDNUM, n1, n2, rank, world, mine, ch);
  *count = mine.size();
  if (chunk) *chunk = ch;
  if (items) {
    if (capacity < mine.size()) return B200CV_MEMORY_ERROR;
    std::memcpy(items, mine.data(), sizeof(SweepItem) * mine.size());
  }
  return B200CV_OK;
}

// ---- measurement helpers ---------------------------------------------------------------------------

int b200cv_fp64_peak(int device, double target_ms, double *tflops, double *elapsed_ms)
{
  if (!tflops) return B200CV_BUG_ERROR;
#define PK(call)                                                                 \
  do {                                                                           \
    cudaError_t e_ = (cudaError_t)(call);                                        \
    if (e_ != cudaSuccess) return fail(nullptr, B200CV_ERROR, cudaGetErrorString(e_)); \
  } while (0)
  PK(cudaSetDevice(device));
  cudaDeviceProp prop;
  PK(cudaGetDeviceProperties(&prop, device));
  int const blocks = prop.multiProcessorCount * 8;
  double *d_out = nullptr;
  PK(cudaMalloc(&d_out, 64));
  cudaEvent_t e0, e1;
  PK(cudaEventCreate(&e0));
  PK(cudaEventCreate(&e1));
  auto run = [&](int iters, float &ms) -> int {
    PK(cudaEventRecord(e0, 0));
    PK(launch_fp64_peak(d_out, iters, blocks, 0));
    PK(cudaEventRecord(e1, 0));
    PK(cudaEventSynchronize(e1));
    PK(cudaEventElapsedTime(&ms, e0, e1));
    return 0;
  };
  float ms = 0.f;
  int iters = 2000;
  if (run(iters, ms)) return B200CV_ERROR;  // warm-up + calibration
  if (run(iters, ms)) return B200CV_ERROR;
  double const per_iter = ms / iters;
  iters = (int)std::max(1000.0, std::min(2.0e7, target_ms / std::max(per_iter, 1e-9)));
  if (run(iters, ms)) return B200CV_ERROR;
  double const flops = 2.0 * 256.0 * (double)iters * 256.0 * (double)blocks;
  *tflops = flops / (ms * 1e-3) / 1e12;
  if (elapsed_ms) *elapsed_ms = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return B200CV_OK;
}

int b200cv_selftest_rcp(int device, int n, double lo, double hi, double *max_ulp)
{
  if (!max_ulp || n < 1) return B200CV_BUG_ERROR;
  PK(cudaSetDevice(device));
  std::vector<double> x(n);
  // log-uniform deterministic sample
  double const llo = std::log(lo), lhi = std::log(hi);
  unsigned long long s = 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < n; i++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    double const u = (double)(s >> 11) / 9007199254740992.0;
    x[i] = std::exp(llo + u * (lhi - llo));
  }
  double *d_x = nullptr, *d_m = nullptr;
  PK(cudaMalloc(&d_x, sizeof(double) * n));
  PK(cudaMalloc(&d_m, sizeof(double)));
  PK(cudaMemcpy(d_x, x.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
  PK(cudaMemset(d_m, 0, sizeof(double)));
  PK(launch_rcp_selftest(d_x, n, d_m, 0));
  PK(cudaMemcpy(max_ulp, d_m, sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(d_x);
  cudaFree(d_m);
  return B200CV_OK;
#undef PK
}

}  // extern "C"
