/*
 * b200cv.h -- C ABI of the B200-native coordination-number engine.
 *
 * This is the drop-in boundary for the Colvars coordination-number hot path
 * (coordNum / selfCoordNum / groupCoord): everything a `colvar::cvc` subclass
 * needs to implement init / read_data / calc_value / calc_gradients /
 * apply_force on top of hand-written sm_100a kernels.  Plain C types only; no
 * torch, no C++ in the signatures.  Device pointers are raw `double*` in the
 * layouts the reference already uses; streams are passed as `void*`
 * (a `cudaStream_t`).
 *
 * Reference interfaces replaced (paths relative to the Colvars tree):
 *   b200cv_create / destroy        colvar::coordnum::coordnum + init        src/colvarcomp_coordnums.cpp:18-156
 *                                  selfcoordnum / groupcoordnum ctors        src/colvarcomp_coordnums.cpp:472-475,564
 *   b200cv_set_group               cvc::parse_group + atom_group arrays      src/colvarcomp.cpp:178-240, src/colvaratoms.h:770-815
 *   b200cv_set_fitting             atom_group::parse_fitting_options         src/colvaratoms.cpp:815-934
 *   b200cv_set_boundaries          system_boundary_conditions::set_boundaries src/colvars_system.h:80-158 (copied in cvc::read_data, src/colvarcomp.cpp:461)
 *   b200cv_read_data               cvc::read_data                            src/colvarcomp.cpp:457-476
 *                                  (reset_atoms_data + read_positions +      src/colvaratoms.cpp:1731,1107-1125,1325-1352;
 *                                  calc_required_properties; GPU twin:       src/colvaratoms_gpu.cpp:176-465,
 *                                  atoms_pos_from_proxy / cog_com kernels)   src/cuda/colvaratoms_kernel.cu:18-243
 *   b200cv_calc_value              coordnum/selfcoordnum/groupcoordnum::calc_value  src/colvarcomp_coordnums.cpp:274-313,547-555,567-577
 *                                  (= cvc::add_calc_value_node in smp gpu)   src/colvarcomp.h:170-173
 *   b200cv_value                   cvc::calc_value_after_gpu                 src/colvarcomp.h:176
 *   b200cv_step                    the three calls above fused into one CUDA-graph launch
 *   b200cv_calc_fit_gradients      cvc::calc_fit_gradients                   src/colvarcomp.cpp:558-565, src/colvaratoms.cpp:1433-1527
 *   b200cv_apply_force             cvc::apply_force -> apply_colvar_force    src/colvarcomp.cpp:568-577, src/colvaratoms.cpp:1637-1704;
 *                                  apply_colvar_force_to_proxy_kernel        src/cuda/colvaratoms_kernel.cu:434-570
 *   b200cv_eval                    coordnum::main_loop on bare arrays        src/colvarcomp_coordnums.cpp:187-248, :478-523
 *   b200cv_eval_host               the same, atom_group host arrays (CPU proxy)     src/colvaratoms.h:556-584
 *   b200cv_calc_host / b200cv_apply_force_host
 *                                  the same step for a CPU proxy
 *                                  (AoS atoms_positions / atoms_new_colvar_forces)  src/colvarproxy.h:139-142,275-300
 *
 * Error convention: every call returns the Colvars error bitmask
 * (src/colvarmodule.h:986-995); b200cv_last_error() gives the message.
 * No exceptions cross this boundary.  Calls on one handle must be serialised by
 * the caller; different handles may be used from different host threads
 * (`smp cvcs`, src/colvarproxy.cpp:299-315).
 *
 * There is no CPU fallback: creating a handle without a usable CUDA device
 * fails with B200CV_ERROR.
 */
#ifndef B200CV_H
#define B200CV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Colvars error bitmask (src/colvarmodule.h:986-995) */
#define B200CV_OK 0
#define B200CV_ERROR 1
#define B200CV_NOT_IMPLEMENTED (1 << 1)
#define B200CV_INPUT_ERROR (1 << 2)
#define B200CV_BUG_ERROR (1 << 3)
#define B200CV_FILE_ERROR (1 << 4)
#define B200CV_MEMORY_ERROR (1 << 5)

/* component kinds = function_type() strings "coordNum" / "selfCoordNum" / "groupCoord" */
#define B200CV_COORDNUM 0
#define B200CV_SELFCOORDNUM 1
#define B200CV_GROUPCOORD 2
/* widening: colvar::distance_inv, "distanceInv" (src/colvarcomp_distances.cpp:365-470) -- the
 * same group1 x group2 pair sweep with the pair term d^-n and the scalar epilogue
 * (sum / (N1 N2))^(-1/n).  exp_numer carries the keyword "exponent" (even, > 0, default 6);
 * cutoff3, exp_denom, tolerance and the *_center_only flags are ignored. */
#define B200CV_DISTANCEINV 3
/* widening: colvar::distance_pairs, "distancePairs" (src/colvarcomp_distances.cpp:486-548): the
 * value is the vector of all N1 x N2 distances; only n1, n2 and the device are read from the
 * config, and only b200cv_distance_pairs_host / _force_host evaluate it. */
#define B200CV_DISTANCEPAIRS 4

/* evaluation flags, same values as colvar::coordnum::ef_* (src/colvarcomp_coordnums.h:32-37) */
#define B200CV_EF_NULL 0
#define B200CV_EF_GRADIENTS 1
#define B200CV_EF_USE_PAIRLIST (1 << 9)
#define B200CV_EF_REBUILD_PAIRLIST (1 << 10)

typedef struct b200cv_cvc b200cv_cvc;

/* Keywords of a coordNum / selfCoordNum / groupCoord block
 * (src/colvarcomp_coordnums.cpp:73-153; doc/colvars-refman-main.tex:1913-2016). */
typedef struct {
  int kind;               /* B200CV_COORDNUM | _SELFCOORDNUM | _GROUPCOORD | _DISTANCEINV */
  int n1;                 /* atoms in group1 */
  int n2;                 /* atoms in group2 (ignored for selfCoordNum) */
  double cutoff3[3];      /* r0 per axis ("cutoff" = all three equal); default 4 A */
  int exp_numer;          /* expNumer, even > 0; default 6 (distanceInv: "exponent") */
  int exp_denom;          /* expDenom, even > 0; default 12 */
  int group1_center_only; /* group1CenterOnly (forced on for groupCoord) */
  int group2_center_only; /* group2CenterOnly (forced on for groupCoord) */
  double tolerance;       /* > 0 enables the pair list */
  int pairlist_freq;      /* pairListFrequency; default 100 */
  int device;             /* CUDA device ordinal */
} b200cv_config;

/* Fill a config with the reference defaults (r0 = 4, n = 6, m = 12, tolerance 0, freq 100). */
void b200cv_config_defaults(b200cv_config *cfg);

/* coordnum::init.  Validates like the reference: odd / non-positive exponents, n < 1,
 * non-positive pairListFrequency -> B200CV_INPUT_ERROR.  With tolerance > 0 runs the
 * reference's Newton solve for tolerance_l2_max on the host
 * (src/colvarcomp_coordnums.cpp:162-184). */
int b200cv_create(const b200cv_config *cfg, b200cv_cvc **out);
int b200cv_destroy(b200cv_cvc *h);
/* Message of the last failing call on this handle (or of the last failed create when
 * h == NULL).  Never NULL. */
const char *b200cv_last_error(const b200cv_cvc *h);

/* Atom group `which` (1 or 2): proxy slot of every atom (atoms_index), masses (NULL = 1.0).
 * Overlapping group1 / group2 -> B200CV_INPUT_ERROR (src/colvarcomp_coordnums.cpp:67-71). */
int b200cv_set_group(b200cv_cvc *h, int which, const int *proxy_index, const double *mass,
                     int n);

/* Fitting options of group `which` (centerToReference / rotateToReference / fittingGroup /
 * refPositions / enableFitGradients).  fit_proxy_index == NULL: the group is fitted on
 * itself.  ref_pos is SoA x..x y..y z..z of nref = (nfit or n) atoms, not yet centred. */
int b200cv_set_fitting(b200cv_cvc *h, int which, int center, int rotate,
                       const int *fit_proxy_index, int nfit, const double *ref_pos_soa,
                       int enable_fit_gradients);

/* Periodic flags and Bravais vectors; all-zero flags = non-periodic.  The four kinds of
 * system_boundary_conditions::position_distance (src/colvars_system.h:161-219: non-periodic,
 * orthogonal, mixed, triclinic with its search over the 27 neighbouring images) all run inside
 * the pair sweep. */
int b200cv_set_boundaries(b200cv_cvc *h, int periodic_x, int periodic_y, int periodic_z,
                          const double A[3], const double B[3], const double C[3]);

/* ---- device-resident step (colvarproxy_gpu layout) ---------------------------------
 * d_proxy_pos / d_proxy_force: device double[3 * proxy_stride], SoA planes x|y|z with
 * stride = number of proxy atoms (src/colvaratoms_gpu.cpp:210-214,
 * src/cuda/colvaratoms_kernel.cu:47-52).  All work is enqueued on `stream`
 * (proxy->get_default_stream()); b200cv_value() is where the host normally waits: it returns
 * the scalar, commits what a pair-list rebuild left behind and repairs a queue overflow (grow,
 * run the evaluation again).  A caller that skips it loses nothing: the list lengths of a
 * rebuild are kept on the device by the evaluation itself, and the next b200cv_calc_value /
 * b200cv_step / b200cv_apply_force on the handle first does the same check (one extra
 * synchronisation in that case), so stale list lengths or gradients of an overflowed evaluation
 * are never used. */
int b200cv_read_data(b200cv_cvc *h, const double *d_proxy_pos, size_t proxy_stride,
                     void *stream);
int b200cv_calc_value(b200cv_cvc *h, long long step_relative, int gradients, void *stream);
int b200cv_calc_fit_gradients(b200cv_cvc *h, void *stream);
/* read_data + calc_value (+ calc_fit_gradients) as ONE CUDA-graph launch: the sequence is
 * captured the first time it is seen for (proxy buffer, stride, step flags) and replayed
 * afterwards; it is re-captured when anything it froze changes (cell, fitting options, queue
 * capacities, work items).  For launch-bound small systems.  Falls back to the three plain
 * calls in sharded runs and before the first pair-list rebuild. */
int b200cv_step(b200cv_cvc *h, const double *d_proxy_pos, size_t proxy_stride,
                long long step_relative, int gradients, void *stream);
/* Waits for the stream used by the last calc_value and returns x.real_value. */
int b200cv_value(b200cv_cvc *h, double *value);
/* F_proxy[idx_i] += force * (R^-1 grad_i) for both groups (+ fit gradients). */
int b200cv_apply_force(b200cv_cvc *h, double force, double *d_proxy_force, size_t proxy_stride,
                       void *stream);

/* ---- bare-array entry (what add_calc_value_node wraps) ------------------------------
 * d_pos1/d_pos2: device SoA double[3*n] exactly as colvaratoms_gpu_buffer_t::d_atoms_pos
 * (src/colvaratoms_gpu.h:19-64); d_grad1/d_grad2: d_atoms_grad, ACCUMULATED into (the
 * reference zeroes them every step, src/colvaratoms_gpu.cpp:192).  For selfCoordNum pass
 * d_pos2 = d_grad2 = NULL.  `flags` = B200CV_EF_*; rebuild is the caller's decision here.
 * The value is retrieved with b200cv_value(). */
int b200cv_eval(b200cv_cvc *h, const double *d_pos1, const double *d_pos2, int flags,
                double *d_grad1, double *d_grad2, void *stream);

/* b200cv_eval without its internal synchronisation: only enqueues work on `stream`, so the
 * call can sit inside a CUDA stream capture and become a child graph of the reference's
 * calc_value graph (add_calc_value_node).  Not available with a pair list.  A queue
 * overflow (more than 1 M pairs with |l2 - 1| < 1/16 in one evaluation) cannot be repaired
 * in flight: b200cv_value() then returns B200CV_MEMORY_ERROR. */
int b200cv_eval_async(b200cv_cvc *h, const double *d_pos1, const double *d_pos2, int flags,
                      double *d_grad1, double *d_grad2, void *stream);

/* Pair list under a captured graph (`smp gpu` route).  b200cv_eval_async with
 * B200CV_EF_USE_PAIRLIST enqueues the rebuild sweep AND the list walk; a mode word written by
 * the host between launches selects on the device which of the two does any work, so the same
 * instantiated graph serves rebuild and list steps (compute_coordnum's
 * `step_relative() % pairListFrequency == 0`, src/colvarcomp_coordnums.cpp:255, evaluated by the
 * caller).  The mode starts at "rebuild".  Buffers cannot grow while a graph holds their
 * addresses: b200cv_pairlist_reserve runs one synchronous rebuild on the current positions
 * and sizes the list to headroom x the number of listed pairs (>= 1); a later overflow is
 * reported by b200cv_value as B200CV_MEMORY_ERROR. */
int b200cv_pairlist_reserve(b200cv_cvc *h, const double *d_pos1, const double *d_pos2,
                            double headroom, void *stream);
int b200cv_pairlist_set_rebuild(b200cv_cvc *h, int rebuild);
/* Captured evaluations freeze what they were enqueued with: the periodic cell and the
 * addresses of the queues.  skip != 0 makes the NEXT launch of a captured b200cv_eval_async do
 * nothing (every kernel of it returns at once) -- for the step on which the caller has found
 * the frozen state out of date (the cell changed: cvc::read_data copies it every step,
 * src/colvarcomp.cpp:457-465) and evaluates with b200cv_eval instead, before it captures again.
 * May be called from a host node of the same graph (cudaLaunchHostFunc), i.e. after the engine
 * has delivered this step's cell and before the kernels read the word. */
int b200cv_async_set_skip(b200cv_cvc *h, int skip);

/* d_grad1 / d_grad2 (d_atoms_grad layout) += the gradients of the last evaluation, with
 * atomics (groups may be shared between components).  Only enqueues work: this is the node
 * add_calc_gradients_node contributes when add_calc_value_node evaluated with
 * d_grad1 = d_grad2 = NULL. */
int b200cv_accumulate_gradients(b200cv_cvc *h, double *d_grad1, double *d_grad2, void *stream);

/* Same evaluation for group arrays that live in HOST memory: atom_group::atoms_pos /
 * atoms_grad of a CPU-proxy run (SoA double[3*n], src/colvaratoms.h:770-815), i.e. after the
 * reference's own read_data() has gathered (and fitted) the groups.  Positions go up,
 * gradients are ACCUMULATED into h_grad1/h_grad2 (may be NULL), the value is returned.
 * This is the entry the `colvar::cvc` shim of colvars_b200/host calls from calc_value(). */
int b200cv_eval_host(b200cv_cvc *h, const double *h_pos1_soa, const double *h_pos2_soa,
                     int flags, double *h_grad1_soa, double *h_grad2_soa, double *value);

/* A batch of independent single pairs on a coordNum handle with n1 == n2: pair k is (atom k of
 * group1, atom k of group2).  *value_sum = sum over k of the switching function, gradients
 * accumulate into h_grad1 / h_grad2 (SoA double[3*n], may be NULL).  This is the hydrogen-bond
 * part of colvar::alpha_angles -- a vector of colvar::h_bond terms, each one
 * compute_pair_coordnum, all with the same weight (src/colvarcomp_protein.cpp:153-166,
 * :213-231) -- and any other batch of hBond-like terms.  Every pair is evaluated in the
 * reference's operation order. */
int b200cv_eval_pairs_host(b200cv_cvc *h, const double *h_pos1_soa, const double *h_pos2_soa,
                           int flags, double *h_grad1_soa, double *h_grad2_soa,
                           double *value_sum);

/* colvar::distance_pairs::calc_value (src/colvarcomp_distances.cpp:486-517): group positions as
 * SoA host arrays (atom_group::atoms_pos), h_dist[i1 * n2 + i2] = |position_distance(p1, p2)|.
 * ... ::apply_force (:526-548): h_force[i1 * n2 + i2] is the force on that distance; the forces
 * on the atoms of the two groups are WRITTEN to h_f1 / h_f2 (SoA double[3*n]), computed on the
 * positions of the last b200cv_distance_pairs_host call. */
int b200cv_distance_pairs_host(b200cv_cvc *h, const double *h_pos1_soa, const double *h_pos2_soa,
                               double *h_dist);
int b200cv_distance_pairs_force_host(b200cv_cvc *h, const double *h_force, double *h_f1_soa,
                                     double *h_f2_soa);

/* ---- host-buffer step (CPU proxy arrays; host<->device copies included) ---------------
 * h_proxy_pos: AoS rvector array double[3*n_proxy] (src/colvarproxy.h:275).
 * b200cv_calc_host = read_data + calc_value (+ fit gradients) and returns the value;
 * b200cv_apply_force_host adds force * gradients into h_proxy_force (AoS).
 * Sharded runs (after b200cv_comm_init): every rank passes the same position array (a
 * replicated-data engine); each rank uploads 1/world of it and the blocks are exchanged GPU to
 * GPU (ncclAllGather).  b200cv_apply_force_host on rank r adds the forces of atoms
 * [n r / world, n (r+1) / world) of every group of n atoms (and of the fitting atoms), so each
 * rank moves 1/world of the gradients and the ranks' force arrays SUM to the full force. */
int b200cv_calc_host(b200cv_cvc *h, const double *h_proxy_pos, size_t n_proxy,
                     long long step_relative, int gradients, double *value);
int b200cv_apply_force_host(b200cv_cvc *h, double force, double *h_proxy_force, size_t n_proxy);

/* ---- inspection (tests, debug_gradients, host shim) ----------------------------------- */
/* Copy a group's current positions / gradients / fit gradients to host SoA arrays. */
int b200cv_get_positions(b200cv_cvc *h, int which, double *h_pos_soa);
int b200cv_get_gradients(b200cv_cvc *h, int which, double *h_grad_soa);
int b200cv_get_fit_gradients(b200cv_cvc *h, int which, double *h_fit_grad_soa, int nfit);
int b200cv_get_centers(b200cv_cvc *h, int which, double com[3], double cog[3]);
/* Rotation quaternion (q0..q3) of a fitted group after read_data. */
int b200cv_get_rotation(b200cv_cvc *h, int which, double q[4]);
/* tolerance_l2_max found by the Newton solve at create time (1e20 without pair list). */
int b200cv_get_tolerance_l2_max(const b200cv_cvc *h, double *l2_max);
/* Number of listed pairs after the last rebuild, and their canonical indices
 * (ip = i*N2 + j, or the row-major upper triangle for selfCoordNum; sorted ascending). */
int b200cv_pairlist_count(b200cv_cvc *h, size_t *count);
int b200cv_pairlist_export(b200cv_cvc *h, uint64_t *h_indices, size_t capacity);
/* How the current list is kept: form 0 = packed (i, j) entries as the all-pairs sweep compacts
 * them, 1 = entries in cell order (cell-list rebuild, periodic cells), 2 = neighbour rows + a
 * few entries (open systems, sparse lists: every group1 atom keeps its listed partners; for
 * selfCoordNum those behind it in cell order), -1 = no rebuild yet; listed pairs, of them as packed entries, entries of all rows
 * (every row padded to a multiple of 32 entries), row chunks. */
int b200cv_pairlist_stats(b200cv_cvc *h, int *form, size_t *pairs, size_t *entries,
                          size_t *row_entries, size_t *row_chunks);
/* Counters of the last calc_value: kernels launched, pairs sent to the exact
 * (reference-order) path. */
int b200cv_last_stats(b200cv_cvc *h, int *kernel_launches, size_t *exact_pairs);
/* Kernels this library has launched in this process so far (all handles, every stream).
 * bench.py reports the difference over its timed region as "gpu_launches". */
unsigned long long b200cv_kernel_launches(void);

/* The CUDA device current on the calling thread: the one the engine's proxy works on (its
 * stream, src/colvarproxy_gpu.h:25, belongs to that device).  The `colvar::cvc` shim creates its
 * components there. */
int b200cv_current_device(int *device);

/* ---- multi-GPU (one process per GPU; group1 rows sharded, group2 replicated) ---------- */
/* NCCL unique id (128 bytes) created on rank 0 and handed to every rank by the caller. */
int b200cv_comm_unique_id(unsigned char id[128]);
/* Join the communicator; afterwards calc_value evaluates only this rank's rows of
 * group1 (cyclic tiles for selfCoordNum) and all-reduces [value | grad1 | grad2] with
 * ncclAllReduce(ncclDouble, ncclSum) on the compute stream.  A queue overflow on any rank
 * travels with that all-reduce, so all ranks grow and repeat the evaluation together. */
int b200cv_comm_init(b200cv_cvc *h, const unsigned char id[128], int rank, int world);
/* Shard without a communicator (the caller reduces): for tests and external plumbing. */
int b200cv_set_shard(b200cv_cvc *h, int rank, int world);

/* The work-item table rank `rank` of `world` sweeps: `count` records of 4 ints
 * {i0, j0, jn, diag} = 1024-row block of group1 x chunk [j0, j0+jn) of group2 (diag: the
 * selfCoordNum chunk overlaps the block's own rows, so j > i is applied).  Pure host logic,
 * needs no GPU; items == NULL only counts. */
int b200cv_plan_items(int kind, int n1, int n2, int rank, int world, int *items,
                      size_t capacity, size_t *count, int *chunk);

/* ---- measurement helpers ---------------------------------------------------------------- */
/* Sustained DFMA throughput of the device (the FP64 roofline denominator):
 * runs a pure dependent-chain DFMA kernel for about `target_ms` and reports TFLOP/s. */
int b200cv_fp64_peak(int device, double target_ms, double *tflops, double *elapsed_ms);
/* Max relative error (in ulps) of the in-kernel reciprocal over n samples in [lo, hi]. */
int b200cv_selftest_rcp(int device, int n, double lo, double hi, double *max_ulp);

#ifdef __cplusplus
}
#endif
#endif /* B200CV_H */
