// fit.cuh -- centerToReference / rotateToReference of an atom group on the device
// (atom_group::calc_apply_roto_translation src/colvaratoms.cpp:1127-1173, calc_fit_gradients
// :1433-1527; optimal rotation src/colvartypes.cpp:211-489; rotation derivative
// src/colvar_rotation_derivative.h).  No host round trip: correlation matrix, 4x4
// eigen-solve, rotation and the fit-gradient chain all run as kernels on the caller's stream.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

namespace b200cv {

struct FitState {
  bool active = false;
  bool center = false;
  bool rotate = false;
  bool fit_gradients = false;
  bool separate = false;  // fittingGroup given
  int nfit = 0;           // atoms of the group the fit is computed on
  int nmain = 0;
  size_t mstride = 0;     // plane stride of the main group's buffers
  size_t fstride = 0;     // plane stride of d_fit_pos / d_ref / d_fit_grad
  std::vector<int> h_index;         // proxy slots of the fitted-on atoms (host copy)
  int *d_fit_index = nullptr;       // separate fitting group only (owned)
  const int *d_index = nullptr;     // slots of the fitted-on atoms on the device (owned or main's)
  double *d_fit_pos = nullptr;      // separate fitting group positions (owned)
  double *d_ref = nullptr;          // centred reference positions
  double *d_main_unrot = nullptr;   // atoms_pos_unrotated of the main group
  double *d_fit_grad = nullptr;     // fit_gradients
  double *d_small = nullptr;        // see fit.cu for the layout
  double *d_q = nullptr;            // = d_small (quaternion q0..q3)
  double ref_pos_cog[3] = {0, 0, 0};
  std::vector<double> h_stage;
};

void fit_free(FitState &F);
int fit_setup(FitState &F, int center, int rotate, const int *fit_proxy_index, int nfit,
              int nmain, size_t mstride, const int *d_main_index,
              const std::vector<int> &h_main_index, const double *ref_pos_soa,
              int enable_fit_gradients, std::string &err);
// gather the fitting group (if separate), centre, rotate, translate: modifies d_pos
int fit_read_and_apply(FitState &F, const double *proxy_pos, size_t pstride, int aos,
                       double *d_pos, size_t stride, int n, cudaStream_t s, std::string &err);
int fit_calc_gradients(FitState &F, const double *d_grad, size_t stride, int n, cudaStream_t s,
                       std::string &err);
int fit_scatter_force(FitState &F, double force, double *proxy_force, size_t pstride, int aos,
                      cudaStream_t s, std::string &err);
// host-buffer route: forces of this rank's share [nfit rank / world, nfit (rank+1) / world) of the
// fitting atoms
int fit_scatter_force_host(FitState &F, double force, double *h_proxy_force, size_t n_proxy,
                           int rank, int world, cudaStream_t s, std::string &err);
int fit_host_inverse_matrix(FitState &F, double Rinv[3][3], std::string &err);
int fit_get_gradients(FitState &F, double *out_soa, int nfit, std::string &err);

}  // namespace b200cv
