// run_colvars_b200_gpu.cpp -- functional-test runner for the GPU-resident route (`smp gpu`):
// the unmodified Colvars library built with COLVARS_CUDA (its graph scheduler, its atom-group
// gather / COM / force-scatter kernels, its colvarproxy_gpu hooks) + the B200
// coordination-number components contributing graph nodes (coordnum_b200, COLVARS_CUDA part).
// A minimal device-resident proxy plays the MD engine: SoA position / force planes in device
// memory, one stream.  Same command line and output files as the reference's runners.
#include <cuda_runtime.h>

#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

#include "colvarmodule.h"
#include "colvarproxy.h"
#include "colvars_version.h"
#include "colvar_gpu_calc.h"

#include "coordnum_b200.h"

namespace {

class device_proxy : public colvarproxy {
public:
  device_proxy()
  {
    smp_mode = smp_mode_t::none;
    support_gpu = true;
    version_int = get_version_from_string(COLVARS_VERSION);
    b_simulation_running = false;
    updated_masses_ = updated_charges_ = true;
    cudaStreamCreate(&stream);
    cvmodule = new colvarmodule(this);
    cvmodule->cv_traj_freq = 0;
    cvmodule->restart_out_freq = 0;
    cvm::rotation::monitor_crossings = false;
    cvmodule->setup_input();
    cvmodule->setup_output();
    setup();
  }
  ~device_proxy() override
  {
    cudaStreamSynchronize(stream);
    release();
    cudaStreamDestroy(stream);
  }
  int setup() override
  {
    boundaries_.reset();
    if (cvmodule) {
      cvmodule->it = cvmodule->it_restart = 0;
      return cvmodule->update_engine_parameters();
    }
    return COLVARS_OK;
  }
  void set_cell(double ax, double by, double cz)
  {
    boundaries_.set_boundaries(true, true, true, cvm::rvector(ax, 0, 0), cvm::rvector(0, by, 0),
                               cvm::rvector(0, 0, cz));
  }
  // the lattice of the coming step the way NAMD's CudaGlobalMaster delivers it: not before
  // calc(), but when the scheduler asks for the "extra info" between the atom groups' read_data
  // graph and the components' calc_value graph (src/colvar_gpu_calc.cpp:753-757,
  // namd/cudaglobalmaster/colvarproxy_cudaglobalmaster.C:731-763)
  void set_cell_late(double ax, double by, double cz)
  {
    late_cell[0] = ax;
    late_cell[1] = by;
    late_cell[2] = cz;
    late_pending = true;
  }
  int wait_for_extra_info_ready() override
  {
    if (late_pending) {
      set_cell(late_cell[0], late_cell[1], late_cell[2]);
      late_pending = false;
    }
    return COLVARS_OK;
  }
  void request_total_force(bool yesno) override { total_force_requested = yesno; }
  bool total_forces_enabled() const override { return total_force_requested; }
  bool total_forces_same_step() const override { return total_force_requested; }
  void log(std::string const &message) override { std::cout << "colvars: " << message; }
  void error(std::string const &message) override
  {
    add_error_msg(message);
    std::cerr << "colvars: " << message;
  }
  int set_unit_system(std::string const &units_in, bool) override
  {
    if (units_in != "real") return COLVARS_ERROR;
    angstrom_value_ = 1.;
    kcal_mol_value_ = 1.;
    units = units_in;
    return COLVARS_OK;
  }
  int check_atom_id(int atom_number) override { return atom_number - 1; }
  int init_atom(int atom_number) override
  {
    int const aid = atom_number - 1;
    for (size_t i = 0; i < atoms_ids.size(); i++) {
      if (atoms_ids[i] == aid) {
        atoms_refcount[i] += 1;
        return (int)i;
      }
    }
    if (aid < 0) return COLVARS_INPUT_ERROR;
    changed = true;
    return add_atom_slot(aid);
  }
  void clear_atom(int index) override
  {
    colvarproxy::clear_atom(index);
    changed = true;
  }
  cvm::real *proxy_atoms_positions_gpu() override { return d_pos; }
  cvm::real *proxy_atoms_total_forces_gpu() override { return d_total; }
  cvm::real *proxy_atoms_new_colvar_forces_gpu() override { return d_force; }
  cudaStream_t get_default_stream() override { return stream; }
  smp_mode_t get_preferred_smp_mode() const override { return smp_mode_t::gpu; }
  std::vector<smp_mode_t> get_available_smp_modes() const override { return {smp_mode_t::gpu}; }
  int set_smp_mode(smp_mode_t mode) override
  {
    if (mode == smp_mode_t::gpu) {
      smp_mode = mode;
      support_gpu = true;
      return COLVARS_OK;
    }
    return COLVARS_NOT_IMPLEMENTED;
  }

  // one frame: positions up (SoA planes, stride = #slots), calc(), forces down
  int step_xyz(const char *filename, bool write_forces)
  {
    size_t const n = atoms_ids.size();
    std::vector<cvm::rvector> positions(n);
    int err = cvmodule->load_coords_xyz(filename, &positions, nullptr, true);
    if (err) return err;
    if (changed) {
      release();
      cudaMalloc(&d_pos, sizeof(double) * 3 * n);
      cudaMalloc(&d_force, sizeof(double) * 3 * n);
      cudaMalloc(&d_total, sizeof(double) * 3 * n);
      if (cvmodule->gpu_calc) cvmodule->gpu_calc->init();  // graphs hold the old addresses
      changed = false;
    }
    std::vector<double> soa(3 * n);
    for (size_t i = 0; i < n; i++) {
      soa[i] = positions[i].x;
      soa[n + i] = positions[i].y;
      soa[2 * n + i] = positions[i].z;
    }
    cudaMemcpy(d_pos, soa.data(), sizeof(double) * 3 * n, cudaMemcpyHostToDevice);
    cudaMemset(d_force, 0, sizeof(double) * 3 * n);
    cudaMemset(d_total, 0, sizeof(double) * 3 * n);
    std::string const fname =
        cvmodule->output_prefix() + "_forces_" + cvm::to_str(cvmodule->it) + ".dat";
    cvmodule->calc();
    cvmodule->it++;
    cudaStreamSynchronize(stream);
    cudaMemcpy(soa.data(), d_force, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost);
    if (write_forces) {
      std::ofstream ofs(fname);
      for (size_t i = 0; i < n; i++) {
        for (int d = 0; d < 3; d++) {
          ofs << std::scientific << std::setprecision(12) << std::setw(20) << soa[d * n + i]
              << std::setw(0) << " ";
        }
        ofs << std::endl;
      }
    }
    return COLVARS_OK;
  }

private:
  void release()
  {
    if (d_pos) cudaFree(d_pos);
    if (d_force) cudaFree(d_force);
    if (d_total) cudaFree(d_total);
    d_pos = d_force = d_total = nullptr;
  }
  cudaStream_t stream = nullptr;
  double *d_pos = nullptr, *d_force = nullptr, *d_total = nullptr;
  bool changed = true;
  double late_cell[3] = {0, 0, 0};
  bool late_pending = false;
};

}  // namespace

// defined by atomgroup_b200.cu when it is linked in front of the reference archive
extern "C" int b200cv_atomgroup_kernels_linked(void) __attribute__((weak));
extern "C" int b200cv_rotation_kernels_linked(void) __attribute__((weak));

int main(int argc, char **argv)
{
  std::string config, traj, prefix;
  bool forces = false;
  double cell[3] = {0, 0, 0}, drift = 0.0;
  bool late_lattice = false;
  for (int a = 1; a < argc; a++) {
    std::string const s = argv[a];
    if ((s == "-c" || s == "--configuration_file") && a + 1 < argc) config = argv[++a];
    else if ((s == "-t" || s == "--trajectory_file") && a + 1 < argc) traj = argv[++a];
    else if ((s == "-o" || s == "--output_prefix") && a + 1 < argc) prefix = argv[++a];
    else if (s == "--force") forces = true;
    else if (s == "--cell" && a + 3 < argc) {
      for (int d = 0; d < 3; d++) cell[d] = std::stod(argv[++a]);
    }
    // the cell grows by this fraction every frame; --late-lattice delivers each frame's cell
    // through wait_for_extra_info_ready() instead of before calc()
    else if (s == "--cell-drift" && a + 1 < argc) drift = std::stod(argv[++a]);
    else if (s == "--late-lattice") late_lattice = true;
  }
  if (config.empty() || traj.empty()) {
    std::cerr << "usage: run_colvars_b200_gpu -c test.in -t trajectory.xyz [-o prefix [--force]]\n";
    return 2;
  }
  // one process per GPU: the device of this rank (torchrun-style LOCAL_RANK, or B200CV_DEVICE);
  // the components pick it up from the context (coordnum_b200.cpp, read_shard_env)
  if (char const *dev = std::getenv("B200CV_DEVICE") ? std::getenv("B200CV_DEVICE")
                                                     : std::getenv("LOCAL_RANK")) {
    if (cudaSetDevice(std::atoi(dev)) != cudaSuccess) {
      std::cerr << "cannot select CUDA device " << dev << "\n";
      return 2;
    }
  }
  device_proxy *proxy = new device_proxy();
  colvarmodule *cvmodule = proxy->cvmodule;
  int err = proxy->set_unit_system("real", false);
  if (!prefix.empty()) err |= proxy->set_output_prefix(prefix);
  err |= cvmodule->setup_input();
  err |= cvmodule->setup_output();
  coordnum_b200::install(cvmodule);
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
      if (late_lattice && frame > 0) proxy->set_cell_late(f * cell[0], f * cell[1], f * cell[2]);
      else proxy->set_cell(f * cell[0], f * cell[1], f * cell[2]);
    }
    io_err = proxy->step_xyz(traj.c_str(), forces && !prefix.empty());
  }
  proxy->post_run();
  bool gpu_route = false;
  for (colvar *cv : *cvmodule->variables()) {
    for (auto &c : cv->get_cvcs()) gpu_route = gpu_route || c->has_gpu_implementation();
  }
  std::cout << "B200 components on the graph route: " << (gpu_route ? "yes" : "no") << "\n";
  // which implementation of src/cuda/colvaratoms_kernel.h the atom groups' graphs were built on
  std::cout << "atom-group kernels: "
            << (b200cv_atomgroup_kernels_linked ? "b200 (atomgroup_b200.cu)" : "reference")
            << "\n";
  std::cout << "optimal-rotation kernels: "
            << (b200cv_rotation_kernels_linked ? "b200 (rotation_b200.cu)" : "reference") << "\n";
  double const max_gradient_error = cvmodule->get_max_gradient_error();
  if (max_gradient_error > 0.) {
    std::cout << "Max gradient error (debugGradients): " << max_gradient_error << "\n";
    if (max_gradient_error > 1e-3) err = 1;
  }
  if (cvmodule->get_error()) err |= 1;
  delete proxy;
  return err;
}
