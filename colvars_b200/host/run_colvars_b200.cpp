// run_colvars_b200.cpp -- functional-test runner: the unmodified Colvars library + stub proxy
// with the three coordination-number components replaced by the B200 ones.  Same command
// line and same output files as the reference's tests/functional/run_colvars_test
// (-c config -t trajectory.xyz -o prefix [--force]), so the reference's golden files apply.
#include <fstream>
#include <iostream>
#include <string>

#include "colvarmodule.h"
#include "colvarproxy.h"
#include "colvarproxy_stub.h"

#include "coordnum_b200.h"

namespace {

// the stub is non-periodic (misc_interfaces/stubs/colvarproxy_stub.cpp:55-56); this one can be
// given an orthorhombic cell so that minimum-image paths run through the real host as well
class cell_proxy : public colvarproxy_stub {
public:
  void set_cell(double ax, double by, double cz)
  {
    boundaries_.set_boundaries(true, true, true, cvm::rvector(ax, 0, 0), cvm::rvector(0, by, 0),
                               cvm::rvector(0, 0, cz));
  }
};

}  // namespace

int main(int argc, char **argv)
{
  std::string config, traj, prefix;
  bool forces = false, reference_components = false;
  double cell[3] = {0, 0, 0}, drift = 0.0;
  for (int a = 1; a < argc; a++) {
    std::string const s = argv[a];
    if ((s == "-c" || s == "--configuration_file") && a + 1 < argc) config = argv[++a];
    else if ((s == "-t" || s == "--trajectory_file") && a + 1 < argc) traj = argv[++a];
    else if ((s == "-o" || s == "--output_prefix") && a + 1 < argc) prefix = argv[++a];
    else if (s == "--force") forces = true;
    // extensions of this runner: the reference's own components (for goldens made on the spot)
    // and a periodic orthorhombic cell
    else if (s == "--reference-components") reference_components = true;
    else if (s == "--cell" && a + 3 < argc) {
      for (int d = 0; d < 3; d++) cell[d] = std::stod(argv[++a]);
    }
    // ... that grows by this fraction every frame (a barostat: cvc::read_data has to pick the
    // cell up each step, src/colvarcomp.cpp:457-465)
    else if (s == "--cell-drift" && a + 1 < argc) drift = std::stod(argv[++a]);
  }
  if (config.empty() || traj.empty()) {
    std::cerr << "usage: run_colvars_b200 -c test.in -t trajectory.xyz [-o prefix [--force]]\n";
    return 2;
  }
  cell_proxy *proxy = new cell_proxy();
  colvarmodule *cvmodule = proxy->cvmodule;
  int err = proxy->set_unit_system("real", false);
  if (!prefix.empty()) err |= proxy->set_output_prefix(prefix);
  err |= cvmodule->setup_input();
  err |= cvmodule->setup_output();
  if (!reference_components) {
    coordnum_b200::install(cvmodule);  // <-- the only difference from the reference runner
  }
  err |= cvmodule->read_config_file(config.c_str());
  if (cell[0] > 0) proxy->set_cell(cell[0], cell[1], cell[2]);

  std::ifstream ifs(traj);
  int natoms = 0;
  ifs >> natoms;
  ifs.close();
  for (int ai = 0; ai < natoms; ai++) proxy->init_atom(ai + 1);
  int io_err = 0;
  for (int frame = 0; !io_err; frame++) {
    if (cell[0] > 0 && drift != 0.0) {
      double const f = 1.0 + drift * frame;
      proxy->set_cell(f * cell[0], f * cell[1], f * cell[2]);
    }
    io_err = proxy->read_frame_xyz(traj.c_str(), forces && !prefix.empty());
  }
  proxy->post_run();
  double const max_gradient_error = cvmodule->get_max_gradient_error();
  if (max_gradient_error > 0.) {
    std::cout << "Max gradient error (debugGradients): " << max_gradient_error << "\n";
    if (max_gradient_error > 1e-3) err = 1;
  }
  if (cvmodule->get_error()) err |= 1;
  delete proxy;
  return err;
}
