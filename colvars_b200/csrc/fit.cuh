// fit.cuh -- centerToReference / rotateToReference support of an atom group
// (atom_group::calc_apply_roto_translation, calc_fit_gradients; src/colvaratoms.cpp:1127-1173,
// 1433-1527; optimal rotation src/colvartypes.cpp:231-489; rotation derivative
// src/colvar_rotation_derivative.h).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

namespace b200cv {

struct FitState {
  bool center = false;
  bool rotate = false;
  bool fit_gradients = false;
  bool separate_group = false;  // fittingGroup given
  int nfit = 0;                 // atoms in the group used for the fit
  size_t fstride = 0;
  int nmain = 0;
  std::vector<int> h_fit_index;
  int *d_fit_index = nullptr;
  double *d_fit_pos = nullptr;       // fitting group positions (centred, rotated at the end)
  double *d_fit_pos_unrot = nullptr; // centred, unrotated (input of the optimal rotation)
  double *d_ref_pos = nullptr;       // centred reference positions
  double *d_main_unrot = nullptr;    // atoms_pos_unrotated of the main group
  double *d_fit_grad = nullptr;      // fit_gradients
  double *d_small = nullptr;         // q[4], cog_fit[3], ref_cog[3], S/eig workspace ...
  double *d_q = nullptr;             // alias into d_small
  double ref_pos_cog[3] = {0, 0, 0};
  double *h_pin = nullptr;           // pinned scratch for host-side application
};

void fit_free(FitState &F);
int fit_setup(FitState &F, int center, int rotate, const int *fit_proxy_index, int nfit,
              int nmain, const double *ref_pos_soa, int enable_fit_gradients, std::string &err);
// gather the fitting group (if separate), centre, rotate, translate: modifies d_pos
int fit_read_and_apply(FitState &F, const double *proxy_pos, size_t pstride, int aos,
                       double *d_pos, size_t stride, int n, cudaStream_t s, std::string &err);
int fit_calc_gradients(FitState &F, const double *d_grad, size_t stride, int n, cudaStream_t s,
                       std::string &err);
int fit_scatter_force(FitState &F, double force, double *proxy_force, size_t pstride, int aos,
                      cudaStream_t s, std::string &err);
int fit_scatter_force_host(FitState &F, double force, double *h_proxy_force, size_t n_proxy,
                           cudaStream_t s, std::string &err);
int fit_host_inverse_matrix(FitState &F, double Rinv[3][3], std::string &err);
int fit_get_gradients(FitState &F, double *out_soa, int nfit, std::string &err);

}  // namespace b200cv
