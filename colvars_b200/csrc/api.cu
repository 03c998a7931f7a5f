// api.cu -- the core of the C ABI declared in include/b200cv.h: component life cycle, atom
// groups, periodic cell, and the device-resident step (read_data / calc_value / value /
// apply_force), the bare-array entries and the inspection calls.  See engine.h for the map of
// the other translation units.
#include <unordered_set>

#include "engine.h"

using namespace b200cv;

namespace b200cv {
std::string &create_error()
{
  thread_local std::string msg;
  return msg;
}
}  // namespace b200cv

namespace {

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
  P.tri_search = 1;
  std::memset(P.recipf, 0, sizeof(P.recipf));
  std::memset(P.cell2f, 0, sizeof(P.cell2f));
  std::memset(P.tri_nf, 0, sizeof(P.tri_nf));
  P.tri_marginf = 0.0f;
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
  // single-precision copies (decisions of general_image)
  for (int a = 0; a < 3; a++) {
    for (int d = 0; d < 3; d++) {
      P.recipf[a][d] = (float)P.recip[a][d];
      P.cell2f[a][d] = (float)P.cell2[a][d];
    }
  }
  for (int k = 0; k < 13; k++) P.tri_nf[k] = (float)std::min(P.tri_n[k], 3.0e38);
  P.tri_marginf = (float)std::ldexp(max_n, -15);
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

const char *b200cv_last_error(const b200cv_cvc *h) { return h ? h->err.c_str() : create_error().c_str(); }

int b200cv_create(const b200cv_config *cfg, b200cv_cvc **out)
{
  if (!cfg || !out) return fail(nullptr, B200CV_BUG_ERROR, "null argument");
  *out = nullptr;
  b200cv_config c = *cfg;
  if (c.kind < 0 || c.kind > 4) return fail(nullptr, B200CV_INPUT_ERROR, "unknown component kind");
  if (c.kind == B200CV_DISTANCEPAIRS) {
    // colvar::distance_pairs has no keywords beyond its groups
    c.tolerance = 0.0;
    c.group1_center_only = c.group2_center_only = 0;
    c.exp_numer = 6;
    c.exp_denom = 12;
    c.cutoff3[0] = c.cutoff3[1] = c.cutoff3[2] = 1.0;
  }
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
    create_error() = h->err;
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
  CREATE_OK(cudaHostAlloc(&h->h_value, 4 * sizeof(double), cudaHostAllocMapped));
  CREATE_OK(cudaHostGetDevicePointer(&h->h_value_dev, h->h_value, 0));
  CREATE_OK(cudaHostAlloc(&h->h_counters, 8 * sizeof(unsigned long long), cudaHostAllocMapped));
  CREATE_OK(cudaHostGetDevicePointer(&h->h_counters_dev, h->h_counters, 0));
  for (int k = 0; k < 8; k++) h->h_counters[k] = 0;
  for (int k = 0; k < 4; k++) h->h_value[k] = 0.0;
  CREATE_OK(cudaHostAlloc(&h->h_mode, 2 * sizeof(int), cudaHostAllocMapped));
  CREATE_OK(cudaHostGetDevicePointer(&h->h_mode_dev, h->h_mode, 0));
  h->h_mode[0] = 1;
  h->h_mode[1] = 0;
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
  comm_destroy(h);
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
  cellsort_free(h->sort);
  rows_free(h->rows);
  free_dev(h->d_stage_pos);
  free_dev(h->d_pairs);
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
  // (colvar::distance_pairs does not mind atoms common to both groups)
  if (!h->self() && h->cfg.kind != B200CV_DISTANCEPAIRS && !O.h_index.empty()) {
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
        int const m = (h->center1() && h->center2()) ? 1 : (h->center1() ? h->g[1].n : h->g[0].n);
        CUDA_OK(h, cudaMalloc(&h->d_pl_center, std::max(m, 1)));
        CUDA_OK(h, cudaMemset(h->d_pl_center, 1, std::max(m, 1)));  // list starts all-true
      } else {
        unsigned long long const n1 = h->g[0].n, n2 = h->self() ? n1 : h->g[1].n;
        unsigned long long const pairs = h->self() ? n1 * (n1 - 1) / 2 : n1 * n2;
        CUDA_OK(h, cells_setup(h->cells, (int)n2));
        h->cells_ready = true;
        CUDA_OK(h, cellsort_setup(h->sort, h->self() ? 0 : (int)n1, (int)n2));
        h->sort_ready = true;
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
    // (<= 1e-5) off-diagonal terms it tolerates, without searching the neighbouring images --
    // route those cells to the general path with the search switched off
    bool exact_diag = true;
    for (int d = 0; d < 3; d++) {
      for (int k = 0; k < 3; k++) {
        if (k != d && P.cell[d][k] != 0.0) exact_diag = false;
      }
    }
    if (!exact_diag) {
      P.bc_type = 3;
      P.tri_search = 0;
    }
  }
  if (P.bc_type == 1 || P.bc_type == 3) fill_general_cell_tables(P);
  if (!P.tri_search) {
    // no image search: no direction can ever win in general_image
    for (int k = 0; k < 13; k++) {
      P.tri_n[k] = 1.0e300;
      P.tri_nf[k] = 3.0e38f;
    }
  }
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
  // the previous evaluation may have left a new list (or an overflow) nobody has looked at
  if (h->unchecked) CUDA_OK_RC(sync_and_check(h));
  int const flags = calc_flags(h, step_relative, gradients);
  h->async_eval = false;
  h->gated = false;
  CUDA_OK_RC(prepare_calc(h, flags));
  h->last_stream = (cudaStream_t)stream;
  h->last_step = step_relative;
  h->last_flags = flags;
  h->have_calc = true;
  h->fit_after_calc = false;
  h->unchecked = true;
  return enqueue_calc(h, flags, h->last_stream);
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
  h->fit_after_calc = true;
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
  // gradients of an evaluation whose queues overflowed are incomplete: never scatter them
  // unchecked (a no-op after b200cv_value; captured evaluations are checked there)
  if (h->unchecked && !h->async_eval) CUDA_OK_RC(sync_and_check(h));
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
  if (h->unchecked && sync) CUDA_OK_RC(sync_and_check(h));
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
  h->gated = !sync;  // every captured evaluation carries the mode word (rebuild / walk / skip)
  h->fit_after_calc = false;
  h->unchecked = true;
  if (sync) CUDA_OK_RC(prepare_calc(h, flags));
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

int b200cv_async_set_skip(b200cv_cvc *h, int skip)
{
  if (!h) return B200CV_BUG_ERROR;
  h->h_mode[1] = skip ? 1 : 0;  // mapped pinned: read by fetch_gate_kernel of the next launch
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
  if (rows_usable(h) && sparse_list(h)) {
    // the row form: one entry per listed pair, one chunk per atom of group1 and per ROW_BUF
    // entries (every chunk padded to a multiple of 32 entries)
    double const ent = headroom * (double)h->pl_n + 32.0 * (double)row_slots(h) + 65536.0;
    double const chk = (double)row_slots(h) + headroom * ent / ROW_BUF + 4096.0;
    if ((unsigned)std::min(4.0e9, ent) > h->rows.cap_idx ||
        (unsigned)std::min(4.0e9, chk) > h->rows.cap_chunks) {
      h->epoch++;
      CUDA_OK(h, rows_reserve(h->rows, (unsigned)std::min(4.0e9, ent),
                              (unsigned)std::min(4.0e9, chk)));
    }
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
    int const m = (h->center1() && h->center2()) ? 1 : (h->center1() ? h->g[1].n : h->g[0].n);
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
    int const m = (h->center1() && h->center2()) ? 1 : (h->center1() ? h->g[1].n : h->g[0].n);
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
  // canonical index of the reference's dense bool array
  // (src/colvarcomp_coordnums.cpp:197-238 / :481-520)
  auto canonical = [&](unsigned long long i, unsigned long long j) {
    if (h->self() && i > j) std::swap(i, j);
    return h->self() ? i * (2 * n1 - i - 1) / 2 + (j - i - 1) : i * n2 + j;
  };
  if (h->pl_rows) {
    // row form (rows.cuh): the entries of every row chunk + the extras (atom indices)
    CellSort &B = h->sort;
    SortedGroup &J = B.side[1];
    SortedGroup &I = h->self() ? J : B.side[0];
    unsigned counts[8];
    CUDA_OK(h, cudaMemcpy(counts, B.d_counts, sizeof(counts), cudaMemcpyDeviceToHost));
    // valid slots: this rank's rows, then the overflow chunks
    std::vector<int4> all((size_t)h->rows.cap_chunks), chunks;
    CUDA_OK(h, cudaMemcpy(all.data(), h->rows.d_chunks, sizeof(int4) * all.size(),
                          cudaMemcpyDeviceToHost));
    {
      size_t const slots = row_slots(h);
      size_t const r0 = (size_t)I.n * (size_t)h->rank / (size_t)h->world;
      size_t const r1 = (size_t)I.n * ((size_t)h->rank + 1) / (size_t)h->world;
      for (size_t r = r0; r < r1; r++) chunks.push_back(all[r]);
      size_t const extra = std::min<size_t>(counts[6], all.size() > slots ? all.size() - slots : 0);
      for (size_t x = 0; x < extra; x++) chunks.push_back(all[slots + x]);
    }
    std::vector<int> oi(I.n), oj(J.n);
    std::vector<unsigned long long> ent((size_t)h->pl_entries);
    CUDA_OK(h, cudaMemcpy(oi.data(), I.d_sorted, sizeof(int) * oi.size(), cudaMemcpyDeviceToHost));
    CUDA_OK(h, cudaMemcpy(oj.data(), J.d_sorted, sizeof(int) * oj.size(), cudaMemcpyDeviceToHost));
    if (!ent.empty()) {
      CUDA_OK(h, cudaMemcpy(ent.data(), h->d_pl, sizeof(unsigned long long) * ent.size(),
                            cudaMemcpyDeviceToHost));
    }
    size_t k = 0, used = 0;
    for (auto const &c : chunks) used = std::max(used, (size_t)c.y + (size_t)std::max(c.z, 0));
    used = std::min(used, (size_t)h->rows.cap_idx);
    std::vector<int> idx(used);
    if (used) {
      CUDA_OK(h, cudaMemcpy(idx.data(), h->rows.d_idx, sizeof(int) * used, cudaMemcpyDeviceToHost));
    }
    for (auto const &c : chunks) {
      int const p = c.x, start = c.y, len = c.z;
      if (len <= 0 || (size_t)start + (size_t)len > used) continue;
      for (int e = start; e < start + len; e++) {
        int const q = idx[(size_t)e];
        if (q >= J.n) continue;  // padding (the far-away point, cellsort.cuh)
        if (k >= count) return fail(h, B200CV_BUG_ERROR, "inconsistent row list");
        out[k++] = canonical((unsigned long long)oi[(size_t)p], (unsigned long long)oj[(size_t)q]);
      }
    }
    for (unsigned long long e : ent) {
      if (k >= count) return fail(h, B200CV_BUG_ERROR, "inconsistent row list");
      out[k++] = canonical(e >> 32, e & 0xffffffffull);
    }
    if (k != count) return fail(h, B200CV_BUG_ERROR, "row list does not add up to its pair count");
    std::sort(out, out + count);
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
    out[k] = canonical(i, j);
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

int b200cv_pairlist_stats(b200cv_cvc *h, int *form, size_t *pairs, size_t *entries,
                          size_t *row_entries, size_t *row_chunks)
{
  if (!h) return B200CV_BUG_ERROR;
  if (!(h->cfg.tolerance > 0)) return fail(h, B200CV_INPUT_ERROR, "no pair list (tolerance = 0)");
  if (h->have_calc) CUDA_OK_RC(sync_and_check(h));
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  size_t chunks = 0, rentries = 0;
  if (h->pl_valid && h->pl_rows) {
    unsigned counts[8];
    CUDA_OK(h, cudaMemcpy(counts, h->sort.d_counts, sizeof(counts), cudaMemcpyDeviceToHost));
    chunks = (size_t)row_slots(h) / (size_t)h->world + counts[6];
    rentries = (size_t)h->row_entries;
  }
  if (form) *form = !h->pl_valid ? -1 : (h->pl_rows ? 2 : (h->pl_sorted ? 1 : 0));
  if (pairs) *pairs = (size_t)h->pl_n;
  if (entries) *entries = (size_t)h->pl_entries;
  if (row_entries) *row_entries = rentries;
  if (row_chunks) *row_chunks = chunks;
  return B200CV_OK;
}

unsigned long long b200cv_kernel_launches(void) { return launches_so_far(); }

int b200cv_current_device(int *device)
{
  if (!device) return B200CV_BUG_ERROR;
  cudaError_t const e = cudaGetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, B200CV_ERROR, cudaGetErrorString(e));
  return B200CV_OK;
}

#if defined(B200CV_TRACE)
// debug builds only: {start, tiles begin, tiles end, end (ns), smid} per CTA of the last sweep
int b200cv_trace_dump(unsigned long long *out, int n) { return b200cv::trace_dump(out, n); }
#endif

}  // extern "C"
