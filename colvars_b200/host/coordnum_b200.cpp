// coordnum_b200.cpp -- see coordnum_b200.h.  Host side of the drop-in: keyword parsing with
// the reference's own parser, then one C-ABI call per calc_value().
#include "coordnum_b200.h"

#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <thread>
#include <vector>

#include "colvaratoms.h"
#include "colvarcomp_coordnums.h"
#include "colvarproxy.h"
#include "colvars_system.h"

namespace {

// read access to the cell the base class copied into cvc::boundary_conditions
// (src/colvarcomp.cpp:461); the fields are protected in the reference
struct boundaries_view : public cvm::system_boundary_conditions {
  void numbers(double out[12]) const
  {
    out[0] = periodic_x;
    out[1] = periodic_y;
    out[2] = periodic_z;
    cvm::rvector const v[3] = {unit_cell_x, unit_cell_y, unit_cell_z};
    for (int a = 0; a < 3; a++) {
      out[3 + 3 * a] = v[a].x;
      out[4 + 3 * a] = v[a].y;
      out[5 + 3 * a] = v[a].z;
    }
  }
  void push(b200cv_cvc *h) const
  {
    double const A[3] = {unit_cell_x.x, unit_cell_x.y, unit_cell_x.z};
    double const B[3] = {unit_cell_y.x, unit_cell_y.y, unit_cell_y.z};
    double const C[3] = {unit_cell_z.x, unit_cell_z.y, unit_cell_z.z};
    b200cv_set_boundaries(h, periodic_x, periodic_y, periodic_z, A, B, C);
  }
};

// the component map is a protected static of colvar
struct component_registry : public colvar {
  explicit component_registry(colvarmodule *m) : colvar(m) {}
  void replace()
  {
    if (global_cvc_map.empty()) define_component_types();
    global_cvc_map["coordNum"] = []() -> colvar::cvc * { return new coordnum_b200(); };
    global_cvc_map["selfCoordNum"] = []() -> colvar::cvc * { return new selfcoordnum_b200(); };
    global_cvc_map["groupCoord"] = []() -> colvar::cvc * { return new groupcoordnum_b200(); };
    global_cvc_map["distanceInv"] = []() -> colvar::cvc * { return new distanceinv_b200(); };
    global_cvc_map["hBond"] = []() -> colvar::cvc * { return new hbond_b200(); };
    global_cvc_map["alpha"] = []() -> colvar::cvc * { return new alpha_b200(); };
    global_cvc_map["distancePairs"] = []() -> colvar::cvc * { return new distancepairs_b200(); };
  }
};

// One process per GPU: the engine -- or the launcher, torchrun style --
// tells every process its place through the environment.  B200CV_RANK / B200CV_WORLD (else
// RANK / WORLD_SIZE), B200CV_DEVICE (else LOCAL_RANK when sharded, else the device current on
// the calling thread, i.e. the one the proxy's stream lives on, src/colvarproxy_gpu.h:25),
// B200CV_NCCL_ID_FILE: a path unique to the run through which rank 0 hands the NCCL id of every
// component to the other ranks (plain files: the host has no MPI of its own).
struct shard_env {
  int rank = 0, world = 1, device = 0;
  std::string id_file;
};

int env_int(char const *a, char const *b, int def)
{
  char const *v = std::getenv(a);
  if (!v && b) v = std::getenv(b);
  return v ? std::atoi(v) : def;
}

shard_env read_shard_env()
{
  shard_env e;
  if (char const *f = std::getenv("B200CV_NCCL_ID_FILE")) e.id_file = f;
  if (!e.id_file.empty()) {
    e.world = env_int("B200CV_WORLD", "WORLD_SIZE", 1);
    e.rank = env_int("B200CV_RANK", "RANK", 0);
  }
  int current = 0;
  if (b200cv_current_device(&current) != B200CV_OK) current = 0;
  e.device = env_int("B200CV_DEVICE", e.world > 1 ? "LOCAL_RANK" : nullptr, current);
  return e;
}

// components are created in the same order on every rank: the serial number names the id file
int g_component_serial = 0;

int join_shards(b200cv_cvc *h, shard_env const &e, std::string &err)
{
  std::string const path = e.id_file + "." + std::to_string(g_component_serial++);
  unsigned char id[128];
  if (e.rank == 0) {
    if (b200cv_comm_unique_id(id) != B200CV_OK) {
      err = b200cv_last_error(nullptr);
      return COLVARS_ERROR;
    }
    std::string const tmp = path + ".tmp";
    {
      std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
      f.write(reinterpret_cast<char const *>(id), sizeof(id));
    }
    if (std::rename(tmp.c_str(), path.c_str()) != 0) {
      err = "cannot write " + path;
      return COLVARS_FILE_ERROR;
    }
  } else {
    bool ok = false;
    for (int attempt = 0; attempt < 1200 && !ok; attempt++) {  // up to two minutes
      std::ifstream f(path, std::ios::binary);
      if (f && f.read(reinterpret_cast<char *>(id), sizeof(id)) && f.gcount() == sizeof(id)) {
        ok = true;
      } else {
        std::this_thread::sleep_for(std::chrono::milliseconds(100));
      }
    }
    if (!ok) {
      err = "timed out waiting for " + path + " (rank 0 writes it)";
      return COLVARS_FILE_ERROR;
    }
  }
  if (b200cv_comm_init(h, id, e.rank, e.world) != B200CV_OK) {
    err = b200cv_last_error(h);
    return COLVARS_ERROR;
  }
  return COLVARS_OK;
}

int set_engine_group(b200cv_cvc *h, int which, cvm::atom_group *g)
{
  if (g->b_dummy) {
    // one pseudo-atom that is not part of the system: an id no real atom can have
    int const id = 0x7ffffff0;
    double const m = 1.0;
    return b200cv_set_group(h, which, &id, &m, 1);
  }
  size_t const n = g->size();
  std::vector<int> ids(n);
  std::vector<double> m(n);
  for (size_t i = 0; i < n; i++) {
    ids[i] = g->id(i);
    m[i] = g->mass(i);
  }
  return b200cv_set_group(h, which, ids.data(), m.data(), (int)n);
}

}  // namespace

void coordnum_b200::install(colvarmodule *cvmodule)
{
  component_registry reg(cvmodule);
  reg.replace();
}

coordnum_b200::coordnum_b200() : coordnum_b200("coordNum") {}

coordnum_b200::coordnum_b200(char const *function_type_in)
{
  set_function_type(function_type_in);
  x.type(colvarvalue::type_scalar);
  cvm::real const r0 = cvmodule->proxy->angstrom_to_internal(4.0);
  r0_vec = cvm::rvector(r0, r0, r0);
}

coordnum_b200::~coordnum_b200()
{
  if (engine) b200cv_destroy(engine);
#if defined(COLVARS_CUDA)
  if (capture_stream) cudaStreamDestroy(capture_stream);
#endif
}

int coordnum_b200::b200_error(int code)
{
  return cvmodule->error(std::string(b200cv_last_error(engine)) + "\n", code);
}

int coordnum_b200::init(std::string const &conf)
{
  int error_code = cvc::init(conf);
  std::string const ftype = function_type();
  bool const is_self = (ftype == "selfCoordNum");
  bool const is_group = (ftype == "groupCoord");

  group1 = parse_group(conf, "group1");
  if (!group1) return error_code | COLVARS_INPUT_ERROR;
  if (group1->b_dummy) {
    error_code |= cvmodule->error("Error: group1 may not be a dummy atom\n", COLVARS_INPUT_ERROR);
  }
  if (!is_self) {
    group2 = parse_group(conf, "group2");
    if (!group2) return error_code | COLVARS_INPUT_ERROR;
    if (!is_group) {
      get_keyval(conf, "group1CenterOnly", b_group1_center_only, group2->b_dummy);
      get_keyval(conf, "group2CenterOnly", b_group2_center_only, group2->b_dummy);
    } else {
      b_group1_center_only = b_group2_center_only = true;
    }
    size_t const c1 = b_group1_center_only ? 1 : group1->size();
    size_t const c2 = b_group2_center_only ? 1 : group2->size();
    num_pairs = c1 * c2;
  } else {
    num_pairs = (group1->size() * (group1->size() - 1)) / 2;
  }
  init_scalar_boundaries(0.0, num_pairs);

  cvm::real r0 = r0_vec[0];
  bool const have_cutoff = get_keyval(conf, "cutoff", r0, r0);
  if (get_keyval(conf, "cutoff3", r0_vec, r0_vec)) {
    if (have_cutoff) {
      error_code |= cvmodule->error(
          "Error: cannot specify \"cutoff\" and \"cutoff3\" at the same time.\n",
          COLVARS_INPUT_ERROR);
    }
  } else if (have_cutoff) {
    r0_vec = cvm::rvector(r0, r0, r0);
  }
  get_keyval(conf, "expNumer", en, en);
  get_keyval(conf, "expDenom", ed, ed);
  if (!is_group) {
    get_keyval(conf, "tolerance", tolerance, tolerance);
    if (tolerance > 0) {
      cvmodule->cite_feature("coordNum pairlist");
      get_keyval(conf, "pairListFrequency", pairlist_freq, pairlist_freq);
    }
  }
  if (error_code != COLVARS_OK) return error_code;

  // everything else (exponent / frequency validation, the Newton solve for
  // tolerance_l2_max, overlap check) happens behind the C ABI with the reference's messages
  b200cv_config cfg;
  b200cv_config_defaults(&cfg);
  cfg.kind = is_self ? B200CV_SELFCOORDNUM : (is_group ? B200CV_GROUPCOORD : B200CV_COORDNUM);
  cfg.n1 = (int)group1->size();
  dummy2 = !is_self && group2->b_dummy;
  if (dummy2) {
    // the position is private to the atom group; read the keyword once more -- by hand, so
    // that this component's keyword registry (positions inside `conf`, used by
    // check_keywords) is not disturbed
    std::string low(conf);
    for (char &ch : low) ch = (char)std::tolower((unsigned char)ch);
    size_t const g2 = low.find("group2");
    size_t const at = (g2 == std::string::npos) ? g2 : low.find("dummyatom", g2);
    double v[3] = {0.0, 0.0, 0.0};
    if (at != std::string::npos) {
      std::string rest = conf.substr(at + 9);
      for (char &ch : rest) {
        if (ch == '(' || ch == ')' || ch == ',') ch = ' ';
      }
      std::sscanf(rest.c_str(), "%lf %lf %lf", &v[0], &v[1], &v[2]);
    }
    dummy2_pos = cvm::atom_pos(v[0], v[1], v[2]);
  }
  cfg.n2 = is_self ? 0 : (dummy2 ? 1 : (int)group2->size());
  cfg.cutoff3[0] = r0_vec.x;
  cfg.cutoff3[1] = r0_vec.y;
  cfg.cutoff3[2] = r0_vec.z;
  cfg.exp_numer = en;
  cfg.exp_denom = ed;
  cfg.group1_center_only = b_group1_center_only;
  cfg.group2_center_only = b_group2_center_only;
  cfg.tolerance = tolerance;
  cfg.pairlist_freq = pairlist_freq;
  cfg.device = 0;
  return error_code | create_engine(cfg);
}

int coordnum_b200::create_engine(b200cv_config const &cfg_in)
{
  // the device of the proxy's context; one rank of a sharded run when the environment says so
  shard_env const env = read_shard_env();
  b200cv_config cfg = cfg_in;
  cfg.device = env.device;
  int rc = b200cv_create(&cfg, &engine);
  if (rc != B200CV_OK) {
    return cvmodule->error(std::string(b200cv_last_error(nullptr)) + "\n", rc);
  }
  rc = set_engine_group(engine, 1, group1);
  if (rc == B200CV_OK && group2) rc = set_engine_group(engine, 2, group2);
  if (rc != B200CV_OK) return b200_error(rc);
  if (env.world > 1) {
    // group1 rows over the ranks, one ncclAllReduce of [value | grad1 | grad2] per evaluation:
    // every rank ends up with the full value and gradients, the host above notices nothing
    std::string err;
    int const jrc = join_shards(engine, env, err);
    if (jrc != COLVARS_OK) return cvmodule->error("Error: " + err + "\n", jrc);
    sharded = true;
    cvmodule->log("b200cv: component \"" + name + "\" sharded, rank " + cvm::to_str(env.rank) +
                  " of " + cvm::to_str(env.world) + " on device " + cvm::to_str(env.device) + "\n");
  }
#if defined(COLVARS_CUDA)
  if (has_gpu_implementation()) {
    // positions and gradients stay on the device (as colvar::rmsd does,
    // src/colvarcomp_distances.cpp:930)
    disable(f_cvc_require_cpu_buffers);
  }
#endif
  return COLVARS_OK;
}

// ---- distanceInv (src/colvarcomp_distances.cpp:365-470) ------------------------------------

distanceinv_b200::distanceinv_b200() : coordnum_b200("distanceInv") { init_as_distance(); }

int distanceinv_b200::init(std::string const &conf)
{
  int error_code = cvc::init(conf);
  group1 = parse_group(conf, "group1");
  group2 = parse_group(conf, "group2");
  if (!group1 || !group2) return error_code | COLVARS_INPUT_ERROR;
  get_keyval(conf, "exponent", exponent, exponent);
  if (cvm::atom_group::overlap(*group1, *group2) != 0) {
    error_code |= cvmodule->error("Error: group1 and group2 have some atoms in common: this is not "
                                  "allowed for distanceInv.\n",
                                  COLVARS_INPUT_ERROR);
  }
  if (is_enabled(f_cvc_debug_gradient)) {
    cvmodule->log("Warning: debugGradients will not give correct results "
                  "for distanceInv, because its value and gradients are computed "
                  "simultaneously.\n");
  }
  if (error_code != COLVARS_OK) return error_code;
  // parity of the exponent and its sign are checked behind the C ABI (reference messages)
  b200cv_config cfg;
  b200cv_config_defaults(&cfg);
  cfg.kind = B200CV_DISTANCEINV;
  cfg.n1 = (int)group1->size();
  cfg.n2 = (int)group2->size();
  cfg.exp_numer = exponent;
  cfg.device = 0;
  return error_code | create_engine(cfg);
}

// ---- hBond (src/colvarcomp_coordnums.cpp:323-469) -------------------------------------------

hbond_b200::hbond_b200() : coordnum_b200("hBond")
{
  cvm::real const r0 = cvmodule->proxy->angstrom_to_internal(3.3);
  r0_vec = cvm::rvector(r0, r0, r0);
  en = 6;
  ed = 8;
  init_scalar_boundaries(0.0, 1.0);
}

int hbond_b200::init(std::string const &conf)
{
  int error_code = cvc::init(conf);
  set_function_type("hBond");
  x.type(colvarvalue::type_scalar);
  init_scalar_boundaries(0.0, 1.0);

  int a_num = -1, d_num = -1;
  get_keyval(conf, "acceptor", a_num, a_num);
  get_keyval(conf, "donor", d_num, a_num);  // the reference's default, src/...coordnums.cpp:349
  if ((a_num == -1) || (d_num == -1)) {
    return error_code |
           cvmodule->error("Error: either acceptor or donor undefined.\n", COLVARS_INPUT_ERROR);
  }
  register_atom_group(new cvm::atom_group);
  {
    colvarproxy *const p = cvmodule->proxy;
    auto modify_atom = atom_groups[0]->get_atom_modifier();
    modify_atom.add_atom(cvm::atom_group::init_atom_from_proxy(p, a_num));
    modify_atom.add_atom(cvm::atom_group::init_atom_from_proxy(p, d_num));
  }
  group1 = atom_groups[0];  // acceptor = row, donor = column of the 1 x 1 problem
  num_pairs = 1;

  cvm::real r0 = r0_vec[0];
  if (get_keyval(conf, "cutoff", r0, r0)) r0_vec = cvm::rvector(r0, r0, r0);
  get_keyval(conf, "expNumer", en, en);
  get_keyval(conf, "expDenom", ed, ed);
  if (error_code != COLVARS_OK) return error_code;

  // odd / non-positive exponents are rejected behind the C ABI with the reference's messages
  b200cv_config cfg;
  b200cv_config_defaults(&cfg);
  cfg.kind = B200CV_COORDNUM;
  cfg.n1 = 1;
  cfg.n2 = 1;
  cfg.cutoff3[0] = r0_vec.x;
  cfg.cutoff3[1] = r0_vec.y;
  cfg.cutoff3[2] = r0_vec.z;
  cfg.exp_numer = en;
  cfg.exp_denom = ed;
  cfg.device = read_shard_env().device;  // one pair: never sharded
  int rc = b200cv_create(&cfg, &engine);
  if (rc != B200CV_OK) {
    return error_code | cvmodule->error(std::string(b200cv_last_error(nullptr)) + "\n", rc);
  }
  int const id1 = group1->id(0), id2 = group1->id(1);
  double const m1 = group1->mass(0), m2 = group1->mass(1);
  rc = b200cv_set_group(engine, 1, &id1, &m1, 1);
  if (rc == B200CV_OK) rc = b200cv_set_group(engine, 2, &id2, &m2, 1);
  if (rc != B200CV_OK) return error_code | b200_error(rc);
  return error_code;
}

void hbond_b200::evaluate(int flags)
{
  static_cast<boundaries_view const &>(boundary_conditions).push(engine);
  cvm::atom_group *g = atom_groups[0];
  double const p1[3] = {g->pos_x(0), g->pos_y(0), g->pos_z(0)};
  double const p2[3] = {g->pos_x(1), g->pos_y(1), g->pos_z(1)};
  double g1[3] = {0.0, 0.0, 0.0}, g2[3] = {0.0, 0.0, 0.0};
  double value = 0.0;
  int const rc = b200cv_eval_host(engine, p1, p2, flags, g1, g2, &value);
  if (rc != B200CV_OK) {
    b200_error(rc);
    return;
  }
  if (flags & B200CV_EF_GRADIENTS) {
    g->grad_x(0) += g1[0];
    g->grad_y(0) += g1[1];
    g->grad_z(0) += g1[2];
    g->grad_x(1) += g2[0];
    g->grad_y(1) += g2[1];
    g->grad_z(1) += g2[2];
  } else {
    x.real_value = value;
  }
}

// value without gradients, then the gradients by a second evaluation whose value is dropped
// (src/colvarcomp_coordnums.cpp:404-469)
void hbond_b200::calc_value() { evaluate(B200CV_EF_NULL); }

void hbond_b200::calc_gradients() { evaluate(B200CV_EF_GRADIENTS); }

// ---- distancePairs (src/colvarcomp_distances.cpp:466-548) ----------------------------------------

distancepairs_b200::~distancepairs_b200()
{
  if (engine) b200cv_destroy(engine);
}

int distancepairs_b200::init(std::string const &conf)
{
  int error_code = colvar::distance_pairs::init(conf);
  if (error_code != COLVARS_OK || !group1 || !group2) return error_code;
  b200cv_config cfg;
  b200cv_config_defaults(&cfg);
  cfg.kind = B200CV_DISTANCEPAIRS;
  cfg.n1 = (int)group1->size();
  cfg.n2 = (int)group2->size();
  cfg.device = read_shard_env().device;
  int rc = b200cv_create(&cfg, &engine);
  if (rc != B200CV_OK) {
    return cvmodule->error(std::string(b200cv_last_error(nullptr)) + "\n", rc);
  }
  rc = set_engine_group(engine, 1, group1);
  if (rc == B200CV_OK) rc = set_engine_group(engine, 2, group2);
  if (rc != B200CV_OK) {
    return cvmodule->error(std::string(b200cv_last_error(engine)) + "\n", rc);
  }
  return error_code;
}

void distancepairs_b200::calc_value()
{
  size_t const n = group1->size() * group2->size();
  x.vector1d_value.resize(n);
  static_cast<boundaries_view const &>(boundary_conditions).push(engine);
  std::vector<double> d(n);
  int const rc = b200cv_distance_pairs_host(engine, &group1->pos_x(0), &group2->pos_x(0), d.data());
  if (rc != B200CV_OK) {
    cvmodule->error(std::string(b200cv_last_error(engine)) + "\n", rc);
    return;
  }
  for (size_t k = 0; k < n; k++) x.vector1d_value[k] = d[k];
}

void distancepairs_b200::apply_force(colvarvalue const &force)
{
  size_t const n1 = group1->size(), n2 = group2->size();
  std::vector<double> f(n1 * n2), f1(3 * n1), f2(3 * n2);
  for (size_t k = 0; k < n1 * n2; k++) f[k] = force[k];
  int const rc = b200cv_distance_pairs_force_host(engine, f.data(), f1.data(), f2.data());
  if (rc != B200CV_OK) {
    cvmodule->error(std::string(b200cv_last_error(engine)) + "\n", rc);
    return;
  }
  auto group1_force_obj = group1->get_group_force_object();
  auto group2_force_obj = group2->get_group_force_object();
  for (size_t i = 0; i < n1; i++) {
    group1_force_obj.add_atom_force(i, cvm::rvector(f1[i], f1[n1 + i], f1[2 * n1 + i]));
  }
  for (size_t j = 0; j < n2; j++) {
    group2_force_obj.add_atom_force(j, cvm::rvector(f2[j], f2[n2 + j], f2[2 * n2 + j]));
  }
}

// ---- alpha (src/colvarcomp_protein.cpp:18-330) ---------------------------------------------------

alpha_b200::~alpha_b200()
{
  if (engine) b200cv_destroy(engine);
}

int alpha_b200::init(std::string const &conf)
{
  // keywords, index groups / residue range, the angle and h_bond terms: the reference's own
  int error_code = colvar::alpha_angles::init(conf);
  if (error_code != COLVARS_OK || hb.empty()) return error_code;
  // the h_bond terms as a batch: pair k = (acceptor of term k, donor of term k)
  b200cv_config cfg;
  b200cv_config_defaults(&cfg);
  cfg.kind = B200CV_COORDNUM;
  cfg.n1 = cfg.n2 = (int)hb.size();
  cfg.cutoff3[0] = cfg.cutoff3[1] = cfg.cutoff3[2] = r0;
  cfg.exp_numer = en;
  cfg.exp_denom = ed;
  cfg.device = read_shard_env().device;  // a handful of pairs: never sharded
  int rc = b200cv_create(&cfg, &engine);
  if (rc != B200CV_OK) {
    return cvmodule->error(std::string(b200cv_last_error(nullptr)) + "\n", rc);
  }
  std::vector<int> acc(hb.size()), don(hb.size());
  for (size_t k = 0; k < hb.size(); k++) {
    acc[k] = hb[k]->atom_groups[0]->id(0);
    don[k] = hb[k]->atom_groups[0]->id(1);
  }
  rc = b200cv_set_group(engine, 1, acc.data(), nullptr, (int)acc.size());
  if (rc == B200CV_OK) rc = b200cv_set_group(engine, 2, don.data(), nullptr, (int)don.size());
  if (rc != B200CV_OK) {
    return cvmodule->error(std::string(b200cv_last_error(engine)) + "\n", rc);
  }
  return error_code;
}

// all hydrogen-bond terms in one call; with gradients they land in the terms' own atom groups
// (atom 0 = acceptor gets -G, atom 1 = donor +G: src/colvarcomp_coordnums.cpp:447-469)
int alpha_b200::evaluate_hbonds(int flags, double &sum)
{
  size_t const n = hb.size();
  std::vector<double> p1(3 * n), p2(3 * n), g1(3 * n, 0.0), g2(3 * n, 0.0);
  for (size_t k = 0; k < n; k++) {
    cvm::atom_group const &g = *hb[k]->atom_groups[0];
    p1[k] = g.pos_x(0); p1[n + k] = g.pos_y(0); p1[2 * n + k] = g.pos_z(0);
    p2[k] = g.pos_x(1); p2[n + k] = g.pos_y(1); p2[2 * n + k] = g.pos_z(1);
  }
  int const rc = b200cv_eval_pairs_host(engine, p1.data(), p2.data(), flags, g1.data(),
                                        g2.data(), &sum);
  if (rc != B200CV_OK) {
    return cvmodule->error(std::string(b200cv_last_error(engine)) + "\n", rc);
  }
  if (flags & B200CV_EF_GRADIENTS) {
    for (size_t k = 0; k < n; k++) {
      cvm::atom_group &g = *hb[k]->atom_groups[0];
      g.grad_x(0) += g1[k]; g.grad_y(0) += g1[n + k]; g.grad_z(0) += g1[2 * n + k];
      g.grad_x(1) += g2[k]; g.grad_y(1) += g2[n + k]; g.grad_z(1) += g2[2 * n + k];
    }
  }
  return COLVARS_OK;
}

void alpha_b200::calc_value()
{
  x.real_value = 0.0;
  // angle terms: the reference's objects and the reference's weighting
  // (src/colvarcomp_protein.cpp:187-210)
  if (theta.size()) {
    cvm::real const theta_norm = (1.0 - hb_coeff) / cvm::real(theta.size());
    for (size_t i = 0; i < theta.size(); i++) {
      theta[i]->calc_value();
      cvm::real const t = (theta[i]->value().real_value - theta_ref) / theta_tol;
      cvm::real const f = (1.0 - (t * t)) / (1.0 - (t * t * t * t));
      x.real_value += theta_norm * f;
    }
  }
  // hydrogen-bond terms (:213-231): same weight for all, so their sum is all that is needed
  if (hb.size() && engine) {
    double sum = 0.0;
    if (evaluate_hbonds(B200CV_EF_NULL, sum) == COLVARS_OK) {
      x.real_value += (hb_coeff / cvm::real(hb.size())) * sum;
    }
  }
}

void alpha_b200::calc_gradients()
{
  for (size_t i = 0; i < theta.size(); i++) theta[i]->calc_gradients();
  if (hb.size() && engine) {
    double sum = 0.0;
    evaluate_hbonds(B200CV_EF_GRADIENTS, sum);
  }
}

void coordnum_b200::calc_value()
{
  // compute_coordnum (src/colvarcomp_coordnums.cpp:251-271): flags for this step
  int flags = is_enabled(f_cvc_gradient) ? B200CV_EF_GRADIENTS : B200CV_EF_NULL;
  if (tolerance > 0 && function_type() != "groupCoord") {
    flags |= B200CV_EF_USE_PAIRLIST;
    if (cvmodule->step_relative() % pairlist_freq == 0) flags |= B200CV_EF_REBUILD_PAIRLIST;
  }
  static_cast<boundaries_view const &>(boundary_conditions).push(engine);
  // group arrays are SoA x..x y..y z..z (src/colvaratoms.h:556-584); read_data() has
  // already gathered (and fitted) them, gradients accumulate on top of its zeros
  const double *p2 = group2 ? &group2->pos_x(0) : nullptr;
  double *g2 = group2 ? &group2->grad_x(0) : nullptr;
  x.real_value = 0.0;
  double dummy_pos[3], dummy_grad[3] = {0.0, 0.0, 0.0};
  if (dummy2) {
    // the reference loops over group2->size() == 0 atoms unless group2CenterOnly is on
    // (its default for a dummy group2, src/colvarcomp_coordnums.cpp:74-75,190-194)
    if (!b_group2_center_only) return;
    // CPU proxy: the group's centre of mass is the dummy position (src/colvaratoms.cpp:1376)
    // and follows set_dummy_pos(); under `smp gpu` the reference does not refresh it for
    // components on the CPU-buffer route, so fall back to the parsed keyword there
    bool const gpu_mode =
        cvmodule->proxy->get_smp_mode() == colvarproxy_smp::smp_mode_t::gpu;
    cvm::atom_pos const c = gpu_mode ? dummy2_pos : group2->center_of_mass();
    dummy_pos[0] = c.x;
    dummy_pos[1] = c.y;
    dummy_pos[2] = c.z;
    p2 = dummy_pos;
    g2 = dummy_grad;  // a dummy atom takes no gradient
  }
  int const rc = b200cv_eval_host(engine, &group1->pos_x(0), p2, flags, &group1->grad_x(0), g2,
                                  &x.real_value);
  if (rc != B200CV_OK) b200_error(rc);
}

void coordnum_b200::calc_gradients()
{
  // produced by calc_value(), as in the reference (src/colvarcomp_coordnums.cpp:316-319)
}


#if defined(COLVARS_CUDA)

bool coordnum_b200::has_gpu_implementation() const
{
  return cvmodule->proxy->get_smp_mode() == colvarproxy_smp::smp_mode_t::gpu && !dummy2;
}

namespace {

// Record whatever `enqueue` puts on `stream` as nodes of `graph`
template <typename F>
int capture_into(cudaGraph_t &graph, cudaStream_t stream, F enqueue)
{
  if (cudaStreamBeginCaptureToGraph(stream, graph, nullptr, nullptr, 0,
                                    cudaStreamCaptureModeRelaxed) != cudaSuccess) {
    return COLVARS_ERROR;
  }
  int const rc = enqueue();
  cudaGraph_t same = nullptr;
  if (cudaStreamEndCapture(stream, &same) != cudaSuccess) return COLVARS_ERROR;
  return rc == B200CV_OK ? COLVARS_OK : COLVARS_ERROR;
}

}  // namespace

int coordnum_b200::add_calc_value_node(
    cudaGraph_t &graph, std::unordered_map<std::string, cudaGraphNode_t> &nodes_map)
{
  (void)nodes_map;
  if (!capture_stream &&
      cudaStreamCreateWithFlags(&capture_stream, cudaStreamNonBlocking) != cudaSuccess) {
    return cvmodule->error("Error: cannot create a CUDA stream.\n", COLVARS_ERROR);
  }
  // the scheduler never calls cvc::read_data() on this route, so copy the cell the way it does
  // (src/colvarcomp.cpp:457-465); it is frozen into the graph
  if (is_enabled(f_cvc_pbc_minimum_image)) {
    boundary_conditions = cvmodule->proxy->get_system_boundaries();
  } else {
    boundary_conditions.reset();
  }
  static_cast<boundaries_view const &>(boundary_conditions).push(engine);
  static_cast<boundaries_view const &>(boundary_conditions).numbers(frozen_cell);
  dynamic_cell = graph_stale = false;  // a fresh capture: current cell, current addresses
  if (char const *f = std::getenv("B200CV_SHIM_FAKE_OVERFLOW_STEP")) fake_overflow_step = std::atoi(f);
  auto &b1 = group1->get_gpu_atom_group()->get_gpu_buffers();
  const double *p2 = group2 ? group2->get_gpu_atom_group()->get_gpu_buffers().d_atoms_pos : nullptr;
  int flags = is_enabled(f_cvc_gradient) ? B200CV_EF_GRADIENTS : B200CV_EF_NULL;
  if (tolerance > 0 && function_type() != "groupCoord") {
    // pair list inside the graph: rebuild sweep and list walk are both captured, the engine's
    // mode word (set in calc_value_after_gpu for the next step) picks one on the device.
    // The list cannot grow under a graph, so size it now on the positions the read_data graph
    // has just produced (src/colvar_gpu_calc.cpp:249-251 synchronises before this point).
    flags |= B200CV_EF_USE_PAIRLIST;
    if (!sharded &&
        b200cv_pairlist_reserve(engine, b1.d_atoms_pos, p2, 2.0, capture_stream) != B200CV_OK) {
      return b200_error(COLVARS_ERROR);
    }
  }
  // gradients stay in the engine's buffers until add_calc_gradients_node hands them over:
  // debug_gradients_gpu relaunches this graph many times (src/colvarcomp.cpp:986-996)
  int const rc = capture_into(graph, capture_stream, [&]() {
    // host node first: is the frozen cell still the proxy's cell?
    if (cudaLaunchHostFunc(capture_stream, &coordnum_b200::cell_check, this) != cudaSuccess) {
      return (int)B200CV_ERROR;
    }
    // a sharded component keeps its collective out of the scheduler's graph: it is evaluated
    // on the proxy's stream in calc_value_after_gpu (device-resident all the same)
    if (sharded) return (int)B200CV_OK;
    return b200cv_eval_async(engine, b1.d_atoms_pos, p2, flags, nullptr, nullptr, capture_stream);
  });
  if (rc != COLVARS_OK) return b200_error(COLVARS_ERROR);
  return COLVARS_OK;
}

void coordnum_b200::current_cell(double out[12])
{
  if (is_enabled(f_cvc_pbc_minimum_image)) {
    cvm::system_boundary_conditions const now = cvmodule->proxy->get_system_boundaries();
    static_cast<boundaries_view const &>(now).numbers(out);
  } else {
    for (int k = 0; k < 12; k++) out[k] = 0.0;
  }
}

// runs as a node of the calc_value graph (no CUDA calls in here): one word in mapped pinned
// memory tells the kernels behind it whether to run
void CUDART_CB coordnum_b200::cell_check(void *self)
{
  coordnum_b200 *c = static_cast<coordnum_b200 *>(self);
  double now[12];
  c->current_cell(now);
  bool differs = false;
  for (int k = 0; k < 12; k++) differs = differs || (now[k] != c->frozen_cell[k]);
  int const skip = (differs || c->dynamic_cell || c->graph_stale || c->sharded) ? 1 : 0;
  b200cv_async_set_skip(c->engine, skip);
  c->launch_skipped.store(skip);
}

// this step outside the graph: the proxy's cell of now, synchronous (queues may grow)
int coordnum_b200::evaluate_now()
{
  if (is_enabled(f_cvc_pbc_minimum_image)) {
    boundary_conditions = cvmodule->proxy->get_system_boundaries();
  } else {
    boundary_conditions.reset();
  }
  static_cast<boundaries_view const &>(boundary_conditions).push(engine);
  int flags = is_enabled(f_cvc_gradient) ? B200CV_EF_GRADIENTS : B200CV_EF_NULL;
  if (tolerance > 0 && function_type() != "groupCoord") {
    flags |= B200CV_EF_USE_PAIRLIST;
    if (cvmodule->step_relative() % pairlist_freq == 0) flags |= B200CV_EF_REBUILD_PAIRLIST;
  }
  auto &b1 = group1->get_gpu_atom_group()->get_gpu_buffers();
  const double *p2 = group2 ? group2->get_gpu_atom_group()->get_gpu_buffers().d_atoms_pos : nullptr;
  // gradients stay in the engine's buffers: the calc_gradients graph hands them over as usual
  int rc = b200cv_eval(engine, b1.d_atoms_pos, p2, flags, nullptr, nullptr,
                       cvmodule->proxy->get_default_stream());
  if (rc == B200CV_OK) rc = b200cv_value(engine, &x.real_value);
  if (rc != B200CV_OK) return b200_error(rc);
  return COLVARS_OK;
}

int coordnum_b200::calc_value_after_gpu()
{
  // the scheduler has synchronised the stream; the scalar sits in mapped pinned memory
  bool const skipped = launch_skipped.load() != 0;
  if (!skipped) {
    int rc = b200cv_value(engine, &x.real_value);
    if (rc == B200CV_OK && fake_overflow_step >= 0 &&
        cvmodule->step_relative() == (cvm::step_number)fake_overflow_step) {
      rc = B200CV_MEMORY_ERROR;
      fake_overflow_step = -1;
    }
    if (rc == B200CV_MEMORY_ERROR) {
      // a queue of the captured evaluation overflowed (it cannot grow under the graph that
      // holds its address): do this step directly -- that path grows the queues -- and keep
      // the captured kernels switched off from now on: they hold the old addresses
      cvmodule->log("b200cv: " + name + ": " + std::string(b200cv_last_error(engine)) +
                    " -- evaluating this step directly; the component stays outside the "
                    "captured graph until the graphs are built again\n");
      graph_stale = true;
      return evaluate_now();
    }
    if (rc != B200CV_OK) return b200_error(rc);
  } else {
    if (!dynamic_cell && !graph_stale && !sharded) {
      dynamic_cell = true;
      cvmodule->log("b200cv: " + name + ": the periodic cell has changed since the graph was "
                    "captured; evaluating outside the captured graph from now on\n");
    }
    int const erc = evaluate_now();
    if (erc != COLVARS_OK) return erc;
  }
  if (tolerance > 0) {
    // compute_coordnum's rebuild test (src/colvarcomp_coordnums.cpp:255) for the NEXT launch
    b200cv_pairlist_set_rebuild(engine, (cvmodule->step_relative() + 1) % pairlist_freq == 0);
  }
  return COLVARS_OK;
}

int coordnum_b200::add_calc_gradients_node(
    cudaGraph_t &graph, std::unordered_map<std::string, cudaGraphNode_t> &nodes_map)
{
  (void)nodes_map;
  auto &b1 = group1->get_gpu_atom_group()->get_gpu_buffers();
  double *g2 = group2 ? group2->get_gpu_atom_group()->get_gpu_buffers().d_atoms_grad : nullptr;
  int const rc = capture_into(graph, capture_stream, [&]() {
    return b200cv_accumulate_gradients(engine, b1.d_atoms_grad, g2, capture_stream);
  });
  if (rc != COLVARS_OK) return b200_error(COLVARS_ERROR);
  return COLVARS_OK;
}

#endif  // COLVARS_CUDA
