// ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Our own driver around the UNMODIFIED reference library (oracle/_ref/libcolvars_ref.a,
// built by oracle/Makefile from the sources where they lie): evaluates coordNum /
// selfCoordNum / groupCoord on raw synthetic arrays through the reference's own public
// path -- config text -> colvarmodule::read_config_string -> colvarmodule::calc() -> harmonic
// bias -> atom_group::apply_colvar_force -- and dumps value, per-atom gradients, applied
// forces and the dense pair list, or times calc().  It is how the C restatement
// (coordnum_oracle.c) and the CUDA path are validated against the reference itself on
// inputs the reference's regression tests do not cover, and it is
// the "reference" CPU baseline of bench.py.
//
// Proxy: the reference's stub (misc_interfaces/stubs/colvarproxy_stub.*) with an O(1)
// init_atom (the stub's linear scan is O(N^2) at set-up); slots are
// pre-registered in atom-number order so slot k = atom k+1 = input row k.
//
// Input file (little endian):
//   int32[12] kind (0 coordNum, 1 selfCoordNum, 2 groupCoord), n1, n2, en, ed,
//             pairListFrequency, flags, nframes (0 = 1), nfit, reserved x3
//               flags: 1 group1CenterOnly, 2 group2CenterOnly, 4 boundaries follow,
//                      8 masses follow, 16 group1 is fitted (32 centerToReference,
//                      64 rotateToReference); nfit > 0: separate fittingGroup = the nfit
//                      atoms after group2, nfit = 0: group1 is fitted on itself
//   double[4] cutoff3 x, y, z, tolerance
//   [flags&4]  int32[4] periodic x y z + pad ; double[9] A, B, C
//   [flags&8]  double[P] masses            (P = n1 + n2 + nfit)
//   [flags&16] double[nref*3] refPositions, AoS (nref = nfit or n1)
//   nframes x double[P*3]   positions, AoS: group1 rows, group2 rows, fitting-group rows
// Output file (--out):
//   per frame: double value, double fa, double[n1*3] grad1 (AoS), double[n2*3] grad2,
//              double[P*3] applied forces, int64 npl, uint64[npl] listed canonical
//              indices (npl = -1 when there is no pair list or it is too large to dump),
//              [flags&16] double[4] rotation quaternion, double[nref*3] fit gradients (AoS),
//                         double[n1*3] fitted group1 positions (AoS)
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "colvarmodule.h"
#include "colvar.h"
#include "colvaratoms.h"
#include "colvarcomp.h"
#include "colvarcomp_coordnums.h"
#include "colvarproxy.h"
#include "colvarproxy_stub.h"

namespace {

class driver_proxy : public colvarproxy_stub {
public:
  bool quiet = true;
  void log(std::string const &message) override
  {
    if (!quiet) std::cout << "colvars: " << message;
  }
  // identity map, O(1)
  int init_atom(int atom_number) override
  {
    int const aid = atom_number - 1;
    if (aid < 0 || aid >= (int)atoms_ids.size()) return COLVARS_INPUT_ERROR;
    atoms_refcount[aid] += 1;
    return aid;
  }
  void preregister(size_t n, const double *masses)
  {
    for (size_t k = 0; k < n; k++) {
      add_atom_slot((int)k);
      atoms_refcount[k] = 0;
      if (masses) atoms_masses[k] = masses[k];
    }
  }
  void set_cell(bool px, bool py, bool pz, const double *c)
  {
    boundaries_.set_boundaries(px, py, pz, cvm::rvector(c[0], c[1], c[2]),
                               cvm::rvector(c[3], c[4], c[5]), cvm::rvector(c[6], c[7], c[8]));
  }
  void set_positions(const double *aos, size_t n)
  {
    auto &pos = *modify_atom_positions();
    for (size_t k = 0; k < n; k++) pos[k] = cvm::rvector(aos[3 * k], aos[3 * k + 1], aos[3 * k + 2]);
    auto &f = *modify_atom_applied_forces();
    std::fill(f.begin(), f.end(), cvm::rvector{0, 0, 0});
  }
};

// read access to the protected pair list of the unmodified class
struct coordnum_peek : public colvar::coordnum {
  static const bool *list(colvar::coordnum *c) { return static_cast<coordnum_peek *>(c)->pairlist.get(); }
};

std::string range(size_t a, size_t b)  // 1-based inclusive
{
  return std::to_string(a) + "-" + std::to_string(b);
}

}  // namespace

int main(int argc, char **argv)
{
  std::string in_path, out_path;
  int threads = 1, steps = 1;
  bool timing = false, verbose = false;
  for (int a = 1; a < argc; a++) {
    std::string const s = argv[a];
    if (s == "--in" && a + 1 < argc) in_path = argv[++a];
    else if (s == "--out" && a + 1 < argc) out_path = argv[++a];
    else if (s == "--threads" && a + 1 < argc) threads = std::atoi(argv[++a]);
    else if (s == "--steps" && a + 1 < argc) steps = std::atoi(argv[++a]);
    else if (s == "--time") timing = true;
    else if (s == "--verbose") verbose = true;
  }
  if (in_path.empty()) {
    std::cerr << "usage: ref_driver --in case.bin [--out res.bin] [--threads T] [--steps S --time]\n";
    return 2;
  }
  std::ifstream in(in_path, std::ios::binary);
  if (!in) { std::cerr << "cannot open " << in_path << "\n"; return 2; }
  int32_t hdr[12];
  double dbl[4];
  in.read((char *)hdr, sizeof(hdr));
  in.read((char *)dbl, sizeof(dbl));
  int const kind = hdr[0];
  size_t const n1 = hdr[1], n2 = hdr[2];
  int const en = hdr[3], ed = hdr[4], freq = hdr[5], flags = hdr[6];
  int const nframes = hdr[7] > 0 ? hdr[7] : 1;
  size_t const nfit = hdr[8];
  bool const fitted = flags & 16;
  size_t const nref = fitted ? (nfit ? nfit : n1) : 0;
  size_t const P = n1 + n2 + nfit;
  int32_t per[4] = {0, 0, 0, 0};
  double cell[9] = {0};
  if (flags & 4) {
    in.read((char *)per, sizeof(per));
    in.read((char *)cell, sizeof(cell));
  }
  std::vector<double> masses;
  if (flags & 8) {
    masses.resize(P);
    in.read((char *)masses.data(), sizeof(double) * P);
  }
  std::vector<double> refpos(3 * nref);
  if (fitted) in.read((char *)refpos.data(), sizeof(double) * 3 * nref);
  std::vector<std::vector<double>> frames(nframes, std::vector<double>(3 * P));
  for (auto &f : frames) in.read((char *)f.data(), sizeof(double) * 3 * P);
  if (!in) { std::cerr << "short input file\n"; return 2; }

  driver_proxy *proxy = new driver_proxy();
  proxy->quiet = !verbose;
  colvarmodule *cvmodule = proxy->cvmodule;
  int err = proxy->set_unit_system("real", false);
  proxy->preregister(P, masses.empty() ? nullptr : masses.data());
  if (flags & 4) proxy->set_cell(per[0], per[1], per[2], cell);

  bool const self = kind == 1;
  const char *key = kind == 0 ? "coordNum" : (kind == 1 ? "selfCoordNum" : "groupCoord");
  int ncomp = 1;
  if (kind == 0 && !(flags & 3) && threads > 1) ncomp = (int)std::min<size_t>(threads, n1);
  std::ostringstream cfg;
  cfg.precision(17);
  cfg << "colvarsTrajFrequency 0\ncolvarsRestartFrequency 0\n";
  cfg << "smp " << (ncomp > 1 ? "cvcs" : "off") << "\n";
  cfg << "colvar {\n  name one\n  width 0.5\n";
  for (int c = 0; c < ncomp; c++) {
    size_t const a = n1 * c / ncomp, b = n1 * (c + 1) / ncomp;
    cfg << "  " << key << " {\n";
    if (ncomp > 1) cfg << "    name c" << c << "\n";
    cfg << "    group1 {\n      atomNumbersRange " << range(a + 1, b) << "\n";
    if (fitted) {
      cfg << "      centerToReference " << ((flags & 32) ? "yes" : "no") << "\n";
      cfg << "      rotateToReference " << ((flags & 64) ? "yes" : "no") << "\n";
      if (nfit) {
        cfg << "      fittingGroup { atomNumbersRange " << range(n1 + n2 + 1, n1 + n2 + nfit)
            << " }\n";
      }
      cfg << "      refPositions";
      for (size_t r = 0; r < nref; r++) {
        cfg << " (" << refpos[3 * r] << ", " << refpos[3 * r + 1] << ", " << refpos[3 * r + 2]
            << ")";
      }
      cfg << "\n";
    }
    cfg << "    }\n";
    if (!self) cfg << "    group2 { atomNumbersRange " << range(n1 + 1, n1 + n2) << " }\n";
    if (dbl[0] == dbl[1] && dbl[0] == dbl[2]) cfg << "    cutoff " << dbl[0] << "\n";
    else cfg << "    cutoff3 (" << dbl[0] << ", " << dbl[1] << ", " << dbl[2] << ")\n";
    cfg << "    expNumer " << en << "\n    expDenom " << ed << "\n";
    if (kind == 0) {
      if (flags & 1) cfg << "    group1CenterOnly yes\n";
      if (flags & 2) cfg << "    group2CenterOnly yes\n";
    }
    if (dbl[3] > 0 && kind != 2) {
      cfg << "    tolerance " << dbl[3] << "\n    pairListFrequency " << freq << "\n";
    }
    cfg << "  }\n";
  }
  cfg << "}\nharmonic {\n  colvars one\n  centers 0.1\n  forceConstant 0.001\n}\n";
  if (threads > 0) {
    // the stub proxy runs OpenMP with the environment's thread count
  }
  err |= cvmodule->read_config_string(cfg.str());
  if (err || cvmodule->get_error()) {
    std::cerr << "reference rejected the configuration\n" << cfg.str();
    return 1;
  }
  colvar *cv = (*cvmodule->variables())[0];

  if (timing) {
    proxy->set_positions(frames[0].data(), P);
    cvmodule->calc();  // warm-up (also the pair-list rebuild when enabled)
    cvmodule->it++;
    auto const t0 = std::chrono::steady_clock::now();
    for (int s = 0; s < steps; s++) {
      proxy->set_positions(frames[s % nframes].data(), P);
      cvmodule->calc();
      cvmodule->it++;
    }
    double const sec =
        std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / steps;
    std::printf("{\"seconds_per_step\": %.9g, \"value\": %.17g, \"components\": %d, "
                "\"threads\": %d}\n", sec, cv->value().real_value, ncomp, threads);
    delete proxy;
    return 0;
  }

  std::ofstream out;
  if (!out_path.empty()) out.open(out_path, std::ios::binary);
  for (int f = 0; f < nframes; f++) {
    proxy->set_positions(frames[f].data(), P);
    cvmodule->calc();
    double const value = cv->value().real_value;
    double const fa = cv->applied_force().real_value;
    if (out.is_open()) {
      out.write((const char *)&value, 8);
      out.write((const char *)&fa, 8);
      // gradients: concatenate the components' group1 slices; group2 summed over components
      std::vector<double> g1(3 * n1, 0.0), g2(3 * n2, 0.0);
      size_t row = 0;
      for (auto &c : cv->get_cvcs()) {
        cvm::atom_group *ag1 = c->atom_groups[0];
        for (size_t i = 0; i < ag1->size(); i++, row++) {
          g1[3 * row] = ag1->grad_x(i);
          g1[3 * row + 1] = ag1->grad_y(i);
          g1[3 * row + 2] = ag1->grad_z(i);
        }
        if (!self) {
          cvm::atom_group *ag2 = c->atom_groups[1];
          for (size_t j = 0; j < ag2->size(); j++) {
            g2[3 * j] += ag2->grad_x(j);
            g2[3 * j + 1] += ag2->grad_y(j);
            g2[3 * j + 2] += ag2->grad_z(j);
          }
        }
      }
      out.write((const char *)g1.data(), 8 * g1.size());
      out.write((const char *)g2.data(), 8 * g2.size());
      auto const &forces = *proxy->modify_atom_applied_forces();
      for (size_t k = 0; k < P; k++) {
        double const v[3] = {forces[k].x, forces[k].y, forces[k].z};
        out.write((const char *)v, 24);
      }
      int64_t npl = -1;
      std::vector<uint64_t> idx;
      if (dbl[3] > 0 && kind != 2 && ncomp == 1) {
        auto *c = dynamic_cast<colvar::coordnum *>(cv->get_cvcs()[0].get());
        const bool *pl = c ? coordnum_peek::list(c) : nullptr;
        size_t const n1c = (flags & 1) ? 1 : n1, n2c = (flags & 2) ? 1 : n2;
        size_t const npairs = self ? n1 * (n1 - 1) / 2 : n1c * n2c;
        if (pl && npairs <= (size_t)400000000) {
          for (size_t k = 0; k < npairs; k++) {
            if (pl[k]) idx.push_back(k);
          }
          npl = (int64_t)idx.size();
        }
      }
      out.write((const char *)&npl, 8);
      if (npl > 0) out.write((const char *)idx.data(), 8 * idx.size());
      if (fitted) {
        cvm::atom_group *ag1 = cv->get_cvcs()[0]->atom_groups[0];
        cvm::quaternion const qq = ag1->rot.q;
        double const qv[4] = {qq.q0, qq.q1, qq.q2, qq.q3};
        out.write((const char *)qv, 32);
        cvm::atom_group *fg = ag1->fitting_group ? ag1->fitting_group : ag1;
        for (size_t j = 0; j < fg->size(); j++) {
          double const v[3] = {fg->fit_gradients_x(j), fg->fit_gradients_y(j),
                               fg->fit_gradients_z(j)};
          out.write((const char *)v, 24);
        }
        for (size_t i = 0; i < ag1->size(); i++) {
          double const v[3] = {ag1->pos_x(i), ag1->pos_y(i), ag1->pos_z(i)};
          out.write((const char *)v, 24);
        }
      }
    } else {
      std::printf("frame %d value %.17g fa %.17g\n", f, value, fa);
    }
    cvmodule->it++;
  }
  delete proxy;
  return 0;
}
