// host_path.cu -- entries for group / proxy arrays that live in HOST memory (a CPU proxy):
// b200cv_eval_host (atom_group arrays after the reference's own read_data) and the whole step
// b200cv_calc_host / b200cv_apply_force_host on AoS proxy arrays.  Host<->device copies happen
// here; the arithmetic is the same kernels as the device-resident step.
#include "engine.h"

using namespace b200cv;

extern "C" {

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

// A batch of independent single pairs: pair k = (atom k of group1, atom k of group2), k < n1 = n2.
// What colvar::alpha_angles asks of its hydrogen-bond terms (src/colvarcomp_protein.cpp:153-166,
// :213-231): every term is one compute_pair_coordnum, and the component needs their sum (all
// terms carry the same weight) and, per atom, the gradient.  One launch of exact_kernel -- the
// reference's operation order for every pair, there are only a handful -- over the entries (k, k).
int b200cv_eval_pairs_host(b200cv_cvc *h, const double *h_pos1, const double *h_pos2, int flags,
                           double *h_grad1, double *h_grad2, double *value_sum)
{
  if (!h || !value_sum || !h_pos1 || !h_pos2) return B200CV_BUG_ERROR;
  if (!groups_ready(h)) return fail(h, B200CV_INPUT_ERROR, "atom groups are not defined");
  if (h->cfg.kind != B200CV_COORDNUM || h->g[0].n != h->g[1].n || h->center1() || h->center2() ||
      h->cfg.tolerance > 0 || h->world > 1) {
    return fail(h, B200CV_INPUT_ERROR,
                "a pair batch needs a coordNum handle with n1 == n2, no centre-only group, no "
                "pair list and no sharding");
  }
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = h->own_stream;
  if (h->unchecked) CUDA_OK_RC(sync_and_check(h));
  int const n = h->g[0].n;
  CUDA_OK_RC(ensure_rare_capacity(h, (unsigned long long)n));
  CUDA_OK_RC(ensure_partials(h, PARTIALS_TAIL));
  int rc = reset_gradients(h, s);
  if (rc) return rc;
  const double *hp[2] = {h_pos1, h_pos2};
  for (int k = 0; k < 2; k++) {
    Group &G = h->g[k];
    CUDA_OK(h, cudaMemcpy2DAsync(G.d_pos, sizeof(double) * G.stride, hp[k], sizeof(double) * G.n,
                                 sizeof(double) * G.n, 3, cudaMemcpyHostToDevice, s));
  }
  flags &= B200CV_EF_GRADIENTS;
  CUDA_OK(h, cudaMemsetAsync(h->d_counters, 0, 16, s));
  CUDA_OK(h, launch_fill_diagonal(h->d_rare, h->rare_count(), n, s));
  Group &G1 = h->g[0], &G2 = h->g[1];
  ExactArgs E{};
  E.x1 = G1.d_pos; E.y1 = G1.d_pos + G1.stride; E.z1 = G1.d_pos + 2 * G1.stride;
  E.x2 = G2.d_pos; E.y2 = G2.d_pos + G2.stride; E.z2 = G2.d_pos + 2 * G2.stride;
  E.g1x = G1.d_grad; E.g1y = G1.d_grad + G1.stride; E.g1z = G1.d_grad + 2 * G1.stride;
  E.g2x = G2.d_grad; E.g2y = G2.d_grad + G2.stride; E.g2z = G2.d_grad + 2 * G2.stride;
  E.list = h->d_rare;
  E.count32 = h->rare_count();
  E.cap = h->rare_cap;
  E.partials = h->d_partials;
  E.flags = flags;
  E.fin_ticket = h->fin_ticket();
  E.fin = FinalizeArgs{};
  E.fin.partials = h->d_partials;
  E.fin.n = EXACT_BLOCKS;
  E.fin.d_value = h->d_value;
  E.fin.h_value = h->h_value_dev;
  E.P = h->P;
  CUDA_OK(h, launch_exact(E, s));
  CUDA_OK(h, cudaStreamSynchronize(s));
  h->launches = 2;
  h->last_stream = s;
  h->last_flags = flags;
  h->have_calc = true;
  h->data_read = true;
  h->async_eval = h->gated = h->unchecked = false;
  *value_sum = *h->h_value;
  if (flags & B200CV_EF_GRADIENTS) {
    double *hg[2] = {h_grad1, h_grad2};
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
    for (int k = 0; k < 2; k++) {
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

// ---- distancePairs (src/colvarcomp_distances.cpp:486-548) ---------------------------------------
// The value is the whole matrix of distances, so the staging buffer d_pairs holds n1 * n2
// doubles (and the force matrix on the way back).

static int distance_pairs_common(b200cv_cvc *h, const double *h_pos1, const double *h_pos2)
{
  if (!groups_ready(h)) return fail(h, B200CV_INPUT_ERROR, "atom groups are not defined");
  if (h->cfg.kind != B200CV_DISTANCEPAIRS) {
    return fail(h, B200CV_INPUT_ERROR, "not a distancePairs handle");
  }
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  size_t const total = (size_t)h->g[0].n * (size_t)h->g[1].n;
  if (h->pairs_cap < total) {
    free_dev(h->d_pairs);
    CUDA_OK(h, cudaMalloc(&h->d_pairs, sizeof(double) * total));
    h->pairs_cap = total;
  }
  if (h_pos1 && h_pos2) {
    const double *hp[2] = {h_pos1, h_pos2};
    for (int k = 0; k < 2; k++) {
      Group &G = h->g[k];
      CUDA_OK(h, cudaMemcpy2DAsync(G.d_pos, sizeof(double) * G.stride, hp[k], sizeof(double) * G.n,
                                   sizeof(double) * G.n, 3, cudaMemcpyHostToDevice, h->own_stream));
    }
    h->data_read = true;
  }
  return B200CV_OK;
}

static DistPairsArgs distance_pairs_args(b200cv_cvc *h)
{
  Group &G1 = h->g[0], &G2 = h->g[1];
  DistPairsArgs A{};
  A.x1 = G1.d_pos; A.y1 = G1.d_pos + G1.stride; A.z1 = G1.d_pos + 2 * G1.stride;
  A.x2 = G2.d_pos; A.y2 = G2.d_pos + G2.stride; A.z2 = G2.d_pos + 2 * G2.stride;
  A.n1 = G1.n;
  A.n2 = G2.n;
  A.f1x = G1.d_grad; A.f1y = G1.d_grad + G1.stride; A.f1z = G1.d_grad + 2 * G1.stride;
  A.f2x = G2.d_grad; A.f2y = G2.d_grad + G2.stride; A.f2z = G2.d_grad + 2 * G2.stride;
  A.P = h->P;
  return A;
}

int b200cv_distance_pairs_host(b200cv_cvc *h, const double *h_pos1, const double *h_pos2,
                               double *h_dist)
{
  if (!h || !h_pos1 || !h_pos2 || !h_dist) return B200CV_BUG_ERROR;
  CUDA_OK_RC(distance_pairs_common(h, h_pos1, h_pos2));
  cudaStream_t s = h->own_stream;
  DistPairsArgs A = distance_pairs_args(h);
  A.dist = h->d_pairs;
  A.force = nullptr;
  CUDA_OK(h, launch_distance_pairs(A, s));
  CUDA_OK(h, cudaMemcpyAsync(h_dist, h->d_pairs, sizeof(double) * (size_t)A.n1 * (size_t)A.n2,
                             cudaMemcpyDeviceToHost, s));
  CUDA_OK(h, cudaStreamSynchronize(s));
  h->launches = 1;
  return B200CV_OK;
}

int b200cv_distance_pairs_force_host(b200cv_cvc *h, const double *h_force, double *h_f1,
                                     double *h_f2)
{
  if (!h || !h_force || !h_f1 || !h_f2) return B200CV_BUG_ERROR;
  if (!h->data_read) return fail(h, B200CV_BUG_ERROR, "distance_pairs force before its value");
  CUDA_OK_RC(distance_pairs_common(h, nullptr, nullptr));
  cudaStream_t s = h->own_stream;
  DistPairsArgs A = distance_pairs_args(h);
  size_t const total = (size_t)A.n1 * (size_t)A.n2;
  CUDA_OK_RC(reset_gradients(h, s));
  CUDA_OK(h, cudaMemcpyAsync(h->d_pairs, h_force, sizeof(double) * total, cudaMemcpyHostToDevice, s));
  A.dist = nullptr;
  A.force = h->d_pairs;
  CUDA_OK(h, launch_distance_pairs(A, s));
  double *out[2] = {h_f1, h_f2};
  for (int k = 0; k < 2; k++) {
    Group &G = h->g[k];
    CUDA_OK(h, cudaMemcpy2DAsync(out[k], sizeof(double) * G.n, G.d_grad, sizeof(double) * G.stride,
                                 sizeof(double) * G.n, 3, cudaMemcpyDeviceToHost, s));
  }
  CUDA_OK(h, cudaStreamSynchronize(s));
  h->launches = 1;
  return B200CV_OK;
}

// ---- host-buffer step --------------------------------------------------------------------------
// Sharded runs (b200cv_comm_init): every rank passes the same proxy array, as the ranks of a
// replicated-data engine hold it.  Each rank uploads 1/world of it over its own PCIe link and the
// blocks are exchanged GPU to GPU (ncclAllGather over NVLink); on the way back rank r scatters
// the forces of its share of every group's atoms -- atoms [n r / world, n (r+1) / world) of a
// group of n -- so each rank moves 1/world of the gradients and the ranks' force arrays add up
// to the full force.

int b200cv_calc_host(b200cv_cvc *h, const double *h_proxy_pos, size_t n_proxy,
                     long long step_relative, int gradients, double *value)
{
  if (!h || !h_proxy_pos || !value) return B200CV_BUG_ERROR;
  if (!groups_ready(h)) return fail(h, B200CV_INPUT_ERROR, "atom groups are not defined");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t s = h->own_stream;
  bool const exchange = h->world > 1 && h->comm != nullptr;
  size_t const per = exchange ? (n_proxy + (size_t)h->world - 1) / (size_t)h->world : n_proxy;
  size_t const need = exchange ? per * (size_t)h->world : n_proxy;
  if (h->stage_n < need) {
    free_dev(h->d_stage_pos);
    CUDA_OK(h, cudaMalloc(&h->d_stage_pos, sizeof(double) * 3 * need));
    h->stage_n = need;
  }
  // the CPU proxy's AoS rvector array (src/colvarproxy.h:275) goes up as it is; the gather by
  // atoms_index runs on the device
  if (exchange) {
    size_t const a0 = std::min(per * (size_t)h->rank, n_proxy);
    size_t const a1 = std::min(a0 + per, n_proxy);
    if (a1 > a0) {
      CUDA_OK(h, cudaMemcpyAsync(h->d_stage_pos + 3 * a0, h_proxy_pos + 3 * a0,
                                 sizeof(double) * 3 * (a1 - a0), cudaMemcpyHostToDevice, s));
    }
    CUDA_OK_RC(comm_allgather(h, h->d_stage_pos, 3 * per, s));
  } else {
    CUDA_OK(h, cudaMemcpyAsync(h->d_stage_pos, h_proxy_pos, sizeof(double) * 3 * n_proxy,
                               cudaMemcpyHostToDevice, s));
  }
  int rc = reset_gradients(h, s);
  if (rc) return rc;
  int const ng = h->self() ? 1 : 2;
  for (int k = 0; k < ng; k++) {
    Group &G = h->g[k];
    if (G.max_index >= 0 && (size_t)G.max_index >= n_proxy) {
      return fail(h, B200CV_BUG_ERROR, "proxy index out of range");
    }
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
  if (h->unchecked) CUDA_OK_RC(sync_and_check(h));
  cudaStream_t s = h->last_stream;
  int const ng = h->self() ? 1 : 2;
  // this rank's share of each group (all of it on a single rank)
  // (b200cv_set_shard without a communicator leaves the reduction to the caller: every rank then
  // holds partial gradients of all atoms and scatters all of them)
  bool const shares = h->world > 1 && h->comm != nullptr;
  size_t const srank = shares ? (size_t)h->rank : 0, sworld = shares ? (size_t)h->world : 1;
  size_t b0[2] = {0, 0}, b1[2] = {0, 0}, off[2] = {0, 0}, total = 0;
  for (int k = 0; k < ng; k++) {
    size_t const n = (size_t)h->g[k].n;
    b0[k] = n * srank / sworld;
    b1[k] = n * (srank + 1) / sworld;
    off[k] = total;
    total += 3 * (b1[k] - b0[k]);
  }
  if (h->h_stage_count < total) {
    if (h->h_stage_grad) cudaFreeHost(h->h_stage_grad);
    h->h_stage_grad = nullptr;
    CUDA_OK(h, cudaHostAlloc(&h->h_stage_grad, sizeof(double) * std::max<size_t>(total, 1),
                             cudaHostAllocDefault));
    h->h_stage_count = total;
  }
  // one strided D2H copy per group (three plane segments); the scatter into the AoS force array
  // is the same host loop as atom_group::apply_colvar_force (src/colvaratoms.cpp:1637-1704)
  for (int k = 0; k < ng; k++) {
    Group &G = h->g[k];
    size_t const m = b1[k] - b0[k];
    if (m == 0) continue;
    CUDA_OK(h, cudaMemcpy2DAsync(h->h_stage_grad + off[k], sizeof(double) * m, G.d_grad + b0[k],
                                 sizeof(double) * G.stride, sizeof(double) * m, 3,
                                 cudaMemcpyDeviceToHost, s));
  }
  CUDA_OK(h, cudaStreamSynchronize(s));
  for (int k = 0; k < ng; k++) {
    Group &G = h->g[k];
    size_t const m = b1[k] - b0[k];
    const double *g = h->h_stage_grad + off[k];
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
    const int *idx = G.h_index.data() + b0[k];
    // distinct atoms of one group: the rows are independent, large groups are split over the
    // host threads (the reference's loop is sequential, src/colvaratoms.cpp:1637-1704)
#pragma omp parallel for schedule(static) if (m > 16384)
    for (long long ii = 0; ii < (long long)m; ii++) {
      size_t const i = (size_t)ii;
      size_t const p = (size_t)idx[i];
      double gx = g[i], gy = g[m + i], gz = g[2 * m + i];
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
    int rc = fit_scatter_force_host(G.fit, force, h_proxy_force, n_proxy, (int)srank, (int)sworld, s, err);
    if (rc) return fail(h, rc, err);
  }
  return B200CV_OK;
}

}  // extern "C"
