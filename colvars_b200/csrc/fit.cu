// fit.cu -- fitting (centerToReference / rotateToReference) kernels.  Placeholder stage:
// groups without fitting options pass through untouched.
#include "fit.cuh"

#include "../../include/b200cv.h"

namespace b200cv {

void fit_free(FitState &F)
{
  auto fr = [](auto *&p) { if (p) cudaFree(p); p = nullptr; };
  fr(F.d_fit_index); fr(F.d_fit_pos); fr(F.d_fit_pos_unrot); fr(F.d_ref_pos);
  fr(F.d_main_unrot); fr(F.d_fit_grad); fr(F.d_small);
  if (F.h_pin) cudaFreeHost(F.h_pin);
  F.h_pin = nullptr;
}

int fit_setup(FitState &F, int center, int rotate, const int *, int, int, const double *, int,
              std::string &err)
{
  if (center || rotate) {
    err = "fitting options are not implemented yet";
    return B200CV_NOT_IMPLEMENTED;
  }
  F.center = F.rotate = false;
  return B200CV_OK;
}

int fit_read_and_apply(FitState &, const double *, size_t, int, double *, size_t, int,
                       cudaStream_t, std::string &) { return B200CV_OK; }
int fit_calc_gradients(FitState &, const double *, size_t, int, cudaStream_t, std::string &)
{ return B200CV_OK; }
int fit_scatter_force(FitState &, double, double *, size_t, int, cudaStream_t, std::string &)
{ return B200CV_OK; }
int fit_scatter_force_host(FitState &, double, double *, size_t, cudaStream_t, std::string &)
{ return B200CV_OK; }
int fit_host_inverse_matrix(FitState &, double Rinv[3][3], std::string &)
{
  for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Rinv[a][b] = a == b ? 1.0 : 0.0;
  return B200CV_OK;
}
int fit_get_gradients(FitState &, double *, int, std::string &err)
{
  err = "group has no fit gradients";
  return B200CV_INPUT_ERROR;
}

}  // namespace b200cv
