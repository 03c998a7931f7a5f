"""Harness for the reference's coordnum regression cases.

Re-creates what `run_colvars_test -c <case>/test.in -t trajectory.xyz --force` does around the
coordination-number component, so that either the CPU oracle or the CUDA path can be run on
the same 5 frames and compared with the reference's golden `test_out.colvars.traj` and
`test_out_forces_<k>.dat` files:

* stub-proxy slot assignment (misc_interfaces/stubs/colvarproxy_stub.cpp:142-165): proxy slots
  are handed out in order of first request -- group1 atoms, then group2 atoms, then
  every remaining atom number 1..N -- while positions are filled in FILE order
  (src/colvarmodule.cpp:2438-2452), i.e. slot k holds XYZ line k;
* harmonic restraint: f = -k (x - x0) / w^2 (src/colvarbias_restraint.cpp), applied through
  cvc::apply_force -> atom_group::apply_colvar_force;
* traj columns printed with %.14e, forces with %.12e (colvarproxy_stub.cpp:176-190).
"""
import re
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden" / "colvars_regression"

CASES = [
    "coordnum", "coordnum-pairlist", "coordnum-aniso", "coordnum-aniso-pairlist",
    "coordnum-group2centeronly", "selfcoordnum", "selfcoordnum-pairlist",
    "groupcoord", "groupcoord-aniso",
    # widening: distanceInv rides the same pair sweep
    "distanceinv",
]

# colvar::h_bond (src/colvarcomp_coordnums.cpp:323-469): ONE pair through compute_pair_coordnum
# (acceptor = first atom, donor = second, r0 = 3.3, 6/8 by default).  Its golden pins the pair
# function with general exponents; it is evaluated as a 1 x 1 coordNum behind the C ABI
# (hbond_b200 in the host shim).
PAIR_CASES = ["hbond"]

KIND = {"coordNum": 0, "selfCoordNum": 1, "groupCoord": 2, "distanceInv": 3}


@dataclass
class CvcConfig:
    """Keywords of a coordNum/selfCoordNum/groupCoord block
    (src/colvarcomp_coordnums.cpp:46-156)."""
    function_type: str = "coordNum"
    group1: list = field(default_factory=list)  # 1-based atom numbers
    group2: list = field(default_factory=list)
    cutoff3: tuple = (4.0, 4.0, 4.0)
    exp_numer: int = 6
    exp_denom: int = 12
    group1_center_only: bool = False
    group2_center_only: bool = False
    tolerance: float = 0.0
    pairlist_freq: int = 100
    exponent: int = 6  # distanceInv (src/colvarcomp_distances.cpp:383)
    # harmonic bias
    centers: float = 0.0
    force_constant: float = 1.0
    width: float = 1.0

    @property
    def kind(self):
        return KIND[self.function_type]


def read_index_file(path):
    groups, name = {}, None
    for line in Path(path).read_text().splitlines():
        m = re.match(r"\s*\[\s*(\S+)\s*\]", line)
        if m:
            name = m.group(1)
            groups[name] = []
        elif name and line.strip():
            groups[name] += [int(t) for t in line.split()]
    return groups


def _yes(v):
    return v.lower() in ("yes", "on", "true")


def parse_test_in(path, index_groups):
    txt = Path(path).read_text()
    cfg = CvcConfig()
    m = re.search(r"\b(coordNum|selfCoordNum|groupCoord|distanceInv|hBond)\s*\{", txt)
    cfg.function_type = m.group(1)
    is_hbond = cfg.function_type == "hBond"

    def kv(key, default=None):
        mm = re.search(r"^\s*" + key + r"\s+(.+?)\s*$", txt, re.M | re.I)
        return mm.group(1) if mm else default

    for gname in ("group1", "group2"):
        mm = re.search(gname + r"\s*\{\s*indexGroup\s+(\S+)\s*\}", txt)
        if mm:
            setattr(cfg, gname, list(index_groups[mm.group(1)]))
    c3 = kv("cutoff3")
    if c3:
        cfg.cutoff3 = tuple(abs(float(t)) for t in re.findall(r"[-+0-9.eE]+", c3))
    c = kv("cutoff")
    if c and not c3:
        cfg.cutoff3 = (float(c),) * 3
    cfg.exp_numer = int(kv("expNumer", 6))
    cfg.exp_denom = int(kv("expDenom", 8 if is_hbond else 12))
    if is_hbond:
        # acceptor first, donor second (src/colvarcomp_coordnums.cpp:356-361)
        cfg.function_type = "coordNum"
        cfg.group1 = [int(kv("acceptor"))]
        cfg.group2 = [int(kv("donor"))]
        if not c:
            cfg.cutoff3 = (3.3, 3.3, 3.3)
    cfg.exponent = int(kv("exponent", 6))
    cfg.tolerance = float(kv("tolerance", 0.0))
    cfg.pairlist_freq = int(kv("pairListFrequency", 100))
    if cfg.function_type == "coordNum":
        cfg.group1_center_only = _yes(kv("group1CenterOnly", "no"))
        cfg.group2_center_only = _yes(kv("group2CenterOnly", "no"))
    if cfg.function_type == "groupCoord":
        cfg.group1_center_only = cfg.group2_center_only = True
    cfg.centers = float(kv("centers", 0.0))
    cfg.force_constant = float(kv("forceConstant", 1.0))
    cfg.width = float(kv("width", 1.0))
    return cfg


def read_xyz_frames(path):
    lines = Path(path).read_text().splitlines()
    frames, k = [], 0
    while k < len(lines) and lines[k].strip():
        n = int(lines[k].split()[0])
        xyz = np.array([[float(t) for t in lines[k + 2 + a].split()[1:4]] for a in range(n)])
        frames.append(xyz)
        k += n + 2
    return frames


def stub_slots(cfg, natoms):
    """Proxy slot of each requested atom number, in order of first request."""
    slot_of = {}
    for num in list(cfg.group1) + list(cfg.group2) + list(range(1, natoms + 1)):
        if num not in slot_of:
            slot_of[num] = len(slot_of)
    idx1 = np.array([slot_of[a] for a in cfg.group1], dtype=np.int32)
    idx2 = np.array([slot_of[a] for a in cfg.group2], dtype=np.int32)
    return idx1, idx2


def load_case(case):
    d = GOLDEN / f"{case}_harmonic-fixed"
    groups = read_index_file(GOLDEN / "index.ndx")
    cfg = parse_test_in(d / "test.in", groups)
    frames = read_xyz_frames(GOLDEN / "trajectory.xyz")
    traj = np.loadtxt(d / "AutoDiff" / "test_out.colvars.traj", comments="#")
    forces = [np.loadtxt(d / "AutoDiff" / f"test_out_forces_{k}.dat") for k in range(len(frames))]
    return cfg, frames, traj, forces


def harmonic_force(cfg, x):
    """colvarbias_restraint_harmonic: f = -k/w^2 (x - x0)."""
    return -cfg.force_constant / (cfg.width * cfg.width) * (x - cfg.centers)


def fmt_matches(value, golden, digits):
    """True when `value` prints to the same %.{digits}e text as the golden number, or
    is within the reference comparer's tolerance (tests/input_files/compare_test.sh:
    relative 1e-6 with absolute fallback 1e-12)."""
    if f"{value:.{digits}e}" == f"{golden:.{digits}e}":
        return True
    return abs(value - golden) <= max(1e-6 * abs(golden), 1e-12)


def run_oracle_case(cfg, frames):
    """The whole per-frame pipeline on the CPU oracle.  Returns [(value, fa, forces(N,3))]."""
    import ctypes as C
    from . import oracle_lib as O
    L = O.lib()
    natoms = frames[0].shape[0]
    idx1, idx2 = stub_slots(cfg, natoms)
    p = O.make_params(cfg.cutoff3, cfg.exp_numer, cfg.exp_denom, cfg.tolerance)
    n1, n2 = len(idx1), len(idx2)
    if cfg.kind == 1:
        npairs = n1 * (n1 - 1) // 2
    else:
        npairs = (1 if cfg.group1_center_only else n1) * (1 if cfg.group2_center_only else n2)
    pl = np.ones(npairs, dtype=np.uint8) if cfg.tolerance > 0 else None
    out = []
    for step, xyz in enumerate(frames):
        proxy_pos = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1)
        pos1 = np.zeros(3 * n1)
        L.cvo_read_positions(n1, idx1.ctypes.data_as(C.POINTER(C.c_int)), O._ptr(proxy_pos),
                             O._ptr(pos1))
        rebuild = pl is not None and (step % cfg.pairlist_freq == 0)
        if cfg.kind == 3:
            pos2 = np.zeros(3 * n2)
            L.cvo_read_positions(n2, idx2.ctypes.data_as(C.POINTER(C.c_int)), O._ptr(proxy_pos),
                                 O._ptr(pos2))
            v, g1, g2 = O.distance_inv(O.unsoa(pos1, n1), O.unsoa(pos2, n2), cfg.exponent)
        elif cfg.kind == 1:
            v, g1 = O.selfcoordnum(p, O.unsoa(pos1, n1), pairlist=pl, rebuild=rebuild)
            g2 = np.zeros((0, 3))
        else:
            pos2 = np.zeros(3 * n2)
            L.cvo_read_positions(n2, idx2.ctypes.data_as(C.POINTER(C.c_int)), O._ptr(proxy_pos),
                                 O._ptr(pos2))
            v, g1, g2 = O.coordnum(p, O.unsoa(pos1, n1), O.unsoa(pos2, n2), pairlist=pl,
                                   rebuild=rebuild, g1_center_only=cfg.group1_center_only,
                                   g2_center_only=cfg.group2_center_only)
        fa = harmonic_force(cfg, v)
        pf = np.zeros(3 * natoms)
        L.cvo_apply_colvar_force(n1, idx1.ctypes.data_as(C.POINTER(C.c_int)), O._ptr(O.soa(g1)),
                                 fa, None, O._ptr(pf))
        if cfg.kind != 1:
            L.cvo_apply_colvar_force(n2, idx2.ctypes.data_as(C.POINTER(C.c_int)),
                                     O._ptr(O.soa(g2)), fa, None, O._ptr(pf))
        out.append((v, fa, pf.reshape(natoms, 3)))
    return out


def check_against_golden(results, traj, forces):
    """Compare [(value, fa, forces)] with the golden traj/forces at print precision."""
    bad = []
    for k, (v, fa, pf) in enumerate(results):
        if not fmt_matches(v, traj[k, 1], 14):
            bad.append(f"step {k}: value {v:.14e} != {traj[k, 1]:.14e}")
        if not fmt_matches(fa, traj[k, 2], 14):
            bad.append(f"step {k}: fa {fa:.14e} != {traj[k, 2]:.14e}")
        g = forces[k]
        for a in range(g.shape[0]):
            for d in range(3):
                if not fmt_matches(pf[a, d], g[a, d], 12):
                    bad.append(f"step {k}: force[{a},{d}] {pf[a, d]:.12e} != {g[a, d]:.12e}")
    return bad
