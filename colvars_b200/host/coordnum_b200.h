// coordnum_b200.h -- `colvar::cvc` components that run the coordination-number hot path on a
// B200 through the C ABI of include/b200cv.h.
//
// Drop-in for colvar::coordnum / selfcoordnum / groupcoordnum
// (src/colvarcomp_coordnums.h:23-141 of the Colvars tree): same configuration keywords, same
// error messages, same `cvc` virtuals (init / calc_value / calc_gradients; read_data,
// calc_fit_gradients, debug_gradients and apply_force are the base-class ones and keep
// working because group positions are read from, and gradients written to, the reference's own
// cvm::atom_group arrays).  Compiles only against the unmodified Colvars headers.
#ifndef COORDNUM_B200_H
#define COORDNUM_B200_H

#include <atomic>
#include <string>

#include "colvarmodule.h"
#include "colvar.h"
#include "colvarcomp.h"

#include "b200cv.h"

class coordnum_b200 : public colvar::cvc {
public:
  coordnum_b200();
  ~coordnum_b200() override;
  int init(std::string const &conf) override;
  void calc_value() override;
  void calc_gradients() override;

#if defined(COLVARS_CUDA)
  // `smp gpu` (src/colvar_gpu_calc.cpp:268-437): this component contributes graph nodes
  // that work on the atom groups' device buffers (colvaratoms_gpu_buffer_t).  With a pair
  // list the one captured graph holds the rebuild sweep and the list walk; a mode word picks.
  bool has_gpu_implementation() const override;
  int add_calc_value_node(cudaGraph_t &graph,
                          std::unordered_map<std::string, cudaGraphNode_t> &nodes_map) override;
  int calc_value_after_gpu() override;
  int add_calc_gradients_node(cudaGraph_t &graph,
                              std::unordered_map<std::string, cudaGraphNode_t> &nodes_map) override;
#endif

  /// Replace the factories of "coordNum", "selfCoordNum" and "groupCoord" in
  /// colvar::global_cvc_map (src/colvar.cpp:822-830, :908-910).  Call once after the
  /// colvarmodule exists and before the first configuration is parsed.
  static void install(colvarmodule *cvmodule);

protected:
  explicit coordnum_b200(char const *function_type_in);
  int b200_error(int code);
  /// b200cv_create + group upload (+ device-buffer route under `smp gpu`)
  int create_engine(b200cv_config const &cfg);

  cvm::atom_group *group1 = nullptr;
  cvm::atom_group *group2 = nullptr;
  cvm::rvector r0_vec;
  int en = 6;
  int ed = 12;
  bool b_group1_center_only = false;
  bool b_group2_center_only = false;
  cvm::real tolerance = 0.0;
  int pairlist_freq = 100;
  size_t num_pairs = 0;
  /// group2 is a dummy atom (`dummyAtom (x, y, z)`): a fixed point that receives no force
  bool dummy2 = false;
  cvm::atom_pos dummy2_pos;  // as parsed from `dummyAtom`
  b200cv_cvc *engine = nullptr;
  bool sharded = false;  // one rank of several (join_shards): group1 rows split, results all-reduced
#if defined(COLVARS_CUDA)
  cudaStream_t capture_stream = nullptr;
  // A captured evaluation freezes the periodic cell (and the addresses of the engine's queues).
  // cvc::read_data copies the cell from the proxy every step (src/colvarcomp.cpp:457-465), so
  // the graph carries a host node in front of the kernels (cell_check, run by the graph itself,
  // i.e. after wait_for_extra_info_ready has delivered this step's lattice,
  // src/colvar_gpu_calc.cpp:753-757) that compares the proxy's cell with the frozen one and,
  // if they differ, makes the kernels of this launch do nothing; calc_value_after_gpu then
  // evaluates the step directly (evaluate_now) with the proxy's cell of now, on this and every
  // later step (a cell that has moved once belongs to a barostat), until the scheduler builds
  // its graphs again (colvarmodule_gpu_calc::init) and add_calc_value_node captures afresh.
  // The same switch serves a queue overflow of a captured evaluation: evaluate_now grows the
  // queue, the captured kernels -- they hold the old address -- stay off.
  double frozen_cell[12] = {0};     // periodic flags + unit cell the captured graph was built with
  std::atomic<int> launch_skipped{0};
  bool dynamic_cell = false;        // the cell has moved: the graph's nodes stay switched off
  bool graph_stale = false;         // a queue was re-allocated: likewise
  int fake_overflow_step = -1;      // B200CV_SHIM_FAKE_OVERFLOW_STEP: exercise the recovery path
  void current_cell(double out[12]);
  static void CUDART_CB cell_check(void *self);
  int evaluate_now();
#endif
};

class selfcoordnum_b200 : public coordnum_b200 {
public:
  selfcoordnum_b200() : coordnum_b200("selfCoordNum") {}
};

class groupcoordnum_b200 : public coordnum_b200 {
public:
  groupcoordnum_b200() : coordnum_b200("groupCoord") {}
};

/// Widening: colvar::distance_inv, "distanceInv"
/// (src/colvarcomp_distances.cpp:365-470) on the same pair sweep; keywords group1, group2,
/// exponent.
class distanceinv_b200 : public coordnum_b200 {
public:
  distanceinv_b200();
  int init(std::string const &conf) override;

protected:
  int exponent = 6;
};

/// colvar::h_bond, "hBond" (src/colvarcomp_coordnums.h:147-166, .cpp:323-469): ONE
/// acceptor-donor pair through the same pair function (cutoff 3.3 A, exponents 6 / 8 by
/// default), both atoms in one atom group; value in calc_value(), gradients in
/// calc_gradients(), as the reference splits them.  Keywords acceptor, donor, cutoff,
/// expNumer, expDenom.  A single pair has nothing to put on a graph: under `smp gpu` it takes
/// the reference's CPU-buffer route.
class hbond_b200 : public coordnum_b200 {
public:
  hbond_b200();
  int init(std::string const &conf) override;
  void calc_value() override;
  void calc_gradients() override;
#if defined(COLVARS_CUDA)
  bool has_gpu_implementation() const override { return false; }
#endif

private:
  void evaluate(int flags);
};

/// colvar::alpha_angles, "alpha" (src/colvarcomp.h:807-846, src/colvarcomp_protein.cpp:18-330):
/// the Calpha-Calpha-Calpha angle terms stay the reference's own objects; the hydrogen-bond
/// terms -- a vector of colvar::h_bond, i.e. one compute_pair_coordnum each -- are evaluated as
/// ONE batch of single pairs behind the C ABI (b200cv_eval_pairs_host).  Keywords, index groups,
/// collect_gradients and apply_force are the base class's: the batch writes its gradients into
/// the h_bond terms' own atom groups, where the base class looks for them.
class alpha_b200 : public colvar::alpha_angles {
public:
  alpha_b200() = default;
  ~alpha_b200() override;
  int init(std::string const &conf) override;
  void calc_value();
  void calc_gradients();

private:
  int evaluate_hbonds(int flags, double &sum);
  b200cv_cvc *engine = nullptr;
};

/// Widening: colvar::distance_pairs, "distancePairs"
/// (src/colvarcomp.h:561-584, src/colvarcomp_distances.cpp:466-548): the vector of all
/// group1 x group2 distances, forces per pair.  Everything but calc_value and apply_force is the
/// base class's.
class distancepairs_b200 : public colvar::distance_pairs {
public:
  distancepairs_b200() = default;
  ~distancepairs_b200() override;
  int init(std::string const &conf) override;
  void calc_value() override;
  void apply_force(colvarvalue const &force) override;

private:
  b200cv_cvc *engine = nullptr;
};

#endif
