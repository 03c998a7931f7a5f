// graph_step.cu -- b200cv_step: read_data + calc_value (+ fit gradients) as one CUDA-graph
// launch, captured once per (proxy buffer, stride, step flags, list form) and replayed.
#include "engine.h"

using namespace b200cv;

extern "C" {

int b200cv_step(b200cv_cvc *h, const double *d_proxy_pos, size_t proxy_stride,
                long long step_relative, int gradients, void *stream)
{
  if (!h) return B200CV_BUG_ERROR;
  if (!groups_ready(h)) return fail(h, B200CV_INPUT_ERROR, "atom groups are not defined");
  NvtxRange nvtx_range("b200cv::step");
  CUDA_OK(h, cudaSetDevice(h->cfg.device));
  cudaStream_t user = (cudaStream_t)stream;
  // which graph applies depends on the list the previous step left behind: commit it first
  if (h->unchecked) CUDA_OK_RC(sync_and_check(h));
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
  CUDA_OK_RC(prepare_calc(h, flags));
  // a rebuild is captured in three forms: all-pairs sweep (first rebuild), cell list, neighbour
  // rows (the host does not run enqueue_calc on a replay: what it would have noted -- the form
  // of the list a rebuild leaves behind -- is noted below), a list walk in three as well:
  // entries as atom indices, entries in cell order, rows
  bool const rows_rebuild = rebuild_with_rows(h, flags) && h->rows.cap_idx > 0;
  bool const cells_rebuild = !rows_rebuild && rebuild_with_cells(h, flags);
  int graph_key = flags | (cells_rebuild ? (1 << 30) : 0) | (rows_rebuild ? (1 << 28) : 0);
  if (use_pl && !rebuild && h->pl_sorted) graph_key |= (1 << 29);
  if (use_pl && !rebuild && h->pl_rows) graph_key |= (1 << 27);
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
    h->pl_rows = rows_rebuild;
  }
  h->launches = hit->launches;
  h->async_eval = false;
  h->gated = false;
  h->last_stream = user;
  h->last_step = step_relative;
  h->last_flags = flags;
  h->have_calc = true;
  h->data_read = true;
  h->unchecked = true;
  h->fit_after_calc = fitted && gradients;
  return B200CV_OK;
}

}  // extern "C"
