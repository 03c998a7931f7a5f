"""ctypes binding of oracle/liboracle.so (the CPU restatement; TEST INFRASTRUCTURE ONLY).

Builds the library on first use with the recipe in oracle/Makefile.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"

EF_NULL = 0
EF_GRADIENTS = 1
EF_USE_PAIRLIST = 1 << 9
EF_REBUILD_PAIRLIST = 1 << 10

BC_NON_PERIODIC, BC_MIXED, BC_PBC_ORTHOGONAL, BC_PBC_TRICLINIC, BC_UNSUPPORTED = range(5)


class Boundaries(C.Structure):
    _fields_ = [
        ("type", C.c_int),
        ("unit_cell", (C.c_double * 3) * 3),
        ("reciprocal_cell", (C.c_double * 3) * 3),
        ("periodic", C.c_int * 3),
    ]


class FitState(C.Structure):
    _fields_ = [
        ("center", C.c_int), ("rotate", C.c_int), ("center_origin", C.c_int),
        ("q", C.c_double * 4),
        ("R", (C.c_double * 3) * 3),
        ("eigval", C.c_double * 4),
        ("eigvec", (C.c_double * 4) * 4),
        ("S", (C.c_double * 4) * 4),
        ("ref_pos_cog", C.c_double * 3),
    ]


class Params(C.Structure):
    _fields_ = [
        ("inv_r0", C.c_double * 3),
        ("inv_r0sq", C.c_double * 3),
        ("en", C.c_int),
        ("ed", C.c_int),
        ("tolerance", C.c_double),
        ("tolerance_l2_max", C.c_double),
        ("bc", Boundaries),
    ]


_lib = None
_dp = C.POINTER(C.c_double)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    so = ORACLE_DIR / "liboracle.so"
    srcs = [ORACLE_DIR / "coordnum_oracle.c", ORACLE_DIR / "coordnum_oracle.h",
            ORACLE_DIR / "fit_oracle.c"]
    if (not so.exists()) or any(s.exists() and s.stat().st_mtime > so.stat().st_mtime
                                for s in srcs):
        subprocess.run(["make", "-C", str(ORACLE_DIR), "liboracle.so"], check=True,
                       capture_output=True)
    L = C.CDLL(str(so))
    L.cvo_integer_power.restype = C.c_double
    L.cvo_integer_power.argtypes = [C.c_double, C.c_int]
    L.cvo_switching_function.restype = C.c_double
    L.cvo_switching_function.argtypes = [C.c_double, _dp, C.c_int, C.c_int, C.c_double, C.c_int]
    L.cvo_compute_tolerance_l2_max.restype = C.c_double
    L.cvo_compute_tolerance_l2_max.argtypes = [C.c_int, C.c_int, C.c_double]
    L.cvo_update_cutoffs.argtypes = [C.POINTER(Params), C.c_double, C.c_double, C.c_double]
    L.cvo_boundaries_reset.argtypes = [C.POINTER(Boundaries)]
    L.cvo_set_boundaries.argtypes = [C.POINTER(Boundaries), C.c_int, C.c_int, C.c_int, _dp, _dp,
                                     _dp]
    L.cvo_position_distance.argtypes = [C.POINTER(Boundaries), _dp, _dp, _dp]
    L.cvo_coordnum.restype = C.c_double
    L.cvo_coordnum.argtypes = [C.POINTER(Params), C.c_size_t, _dp, C.c_size_t, _dp, C.c_int,
                               C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, C.c_void_p, C.c_int]
    L.cvo_selfcoordnum.restype = C.c_double
    L.cvo_selfcoordnum.argtypes = [C.POINTER(Params), C.c_size_t, _dp, _dp, C.c_void_p, C.c_int]
    L.cvo_coordnum_omp.restype = C.c_double
    L.cvo_coordnum_omp.argtypes = [C.POINTER(Params), C.c_size_t, _dp, C.c_size_t, _dp, _dp, _dp,
                                   C.c_void_p, C.c_int, C.c_int]
    L.cvo_selfcoordnum_omp.restype = C.c_double
    L.cvo_selfcoordnum_omp.argtypes = [C.POINTER(Params), C.c_size_t, _dp, _dp, C.c_void_p,
                                       C.c_int, C.c_int]
    L.cvo_read_positions.argtypes = [C.c_size_t, C.POINTER(C.c_int), _dp, _dp]
    L.cvo_center_of_geometry.argtypes = [C.c_size_t, _dp, _dp]
    L.cvo_center_of_mass.argtypes = [C.c_size_t, _dp, _dp, C.c_double, _dp]
    L.cvo_set_weighted_gradient.argtypes = [C.c_size_t, _dp, _dp, _dp]
    L.cvo_apply_colvar_force.argtypes = [C.c_size_t, C.POINTER(C.c_int), _dp, C.c_double, _dp,
                                         _dp]
    L.cvo_distance_inv.restype = C.c_double
    L.cvo_distance_inv.argtypes = [C.POINTER(Boundaries), C.c_int, C.c_size_t, _dp, C.c_size_t,
                                   _dp, _dp, _dp, C.c_int]
    L.cvo_apply_roto_translation.argtypes = [C.POINTER(FitState), C.c_size_t, _dp, _dp, C.c_size_t,
                                             _dp, _dp]
    L.cvo_calc_fit_gradients.argtypes = [C.POINTER(FitState), C.c_size_t, _dp, _dp, C.c_size_t,
                                         _dp, _dp, _dp]
    L.cvo_distance_pairs.argtypes = [C.POINTER(Boundaries), C.c_size_t, _dp, C.c_size_t, _dp, _dp,
                                     _dp, _dp, _dp]
    L.cvo_distance_pairs.restype = None
    _lib = L
    return L


class Fit:
    """centerToReference / rotateToReference of one group (atom_group::parse_fitting_options,
    center_ref_pos: src/colvaratoms.cpp:815-934,1186-1211)."""

    def __init__(self, ref_pos, center=True, rotate=True):
        ref = np.asarray(ref_pos, dtype=np.float64)
        self.nref = len(ref)
        # center_ref_pos: sequential sum, divide, subtract
        cog = np.zeros(3)
        for r in ref:
            cog += r
        cog /= self.nref
        self.ref_centered = soa(ref - cog)
        self.fs = FitState()
        self.fs.center, self.fs.rotate, self.fs.center_origin = int(center), int(rotate), 0
        for d in range(3):
            self.fs.ref_pos_cog[d] = cog[d]

    def apply(self, pos_main, pos_fit=None):
        """Returns (fitted main (n,3), main unrotated SoA, fitted fit-group positions or None)."""
        L = lib()
        n = len(pos_main)
        pm = soa(pos_main)
        un = np.zeros(3 * n)
        if pos_fit is None:
            L.cvo_apply_roto_translation(C.byref(self.fs), n, _ptr(pm), _ptr(un), 0, None,
                                         _ptr(self.ref_centered))
            return unsoa(pm, n), un, None
        nf = len(pos_fit)
        pf = soa(pos_fit)
        L.cvo_apply_roto_translation(C.byref(self.fs), n, _ptr(pm), _ptr(un), nf, _ptr(pf),
                                     _ptr(self.ref_centered))
        return unsoa(pm, n), un, unsoa(pf, nf)

    def fit_gradients(self, grad_main, main_unrot_soa, nfit):
        L = lib()
        n = len(grad_main)
        out = np.zeros(3 * nfit)
        L.cvo_calc_fit_gradients(C.byref(self.fs), n, _ptr(soa(grad_main)), _ptr(main_unrot_soa),
                                 nfit, None, _ptr(self.ref_centered), _ptr(out))
        return unsoa(out, nfit)

    @property
    def q(self):
        return np.array(list(self.fs.q))

    @property
    def R(self):
        return np.array([[self.fs.R[a][b] for b in range(3)] for a in range(3)])


def make_params(r0=(4.0, 4.0, 4.0), en=6, ed=12, tolerance=0.0, boundaries=None):
    """coordnum::coordnum / init (src/colvarcomp_coordnums.cpp:18-43,122-141)."""
    L = lib()
    p = Params()
    L.cvo_update_cutoffs(C.byref(p), float(r0[0]), float(r0[1]), float(r0[2]))
    p.en, p.ed = int(en), int(ed)
    p.tolerance = float(tolerance)
    p.tolerance_l2_max = 1.0e20
    if tolerance > 0:
        p.tolerance_l2_max = L.cvo_compute_tolerance_l2_max(p.en, p.ed, p.tolerance)
    L.cvo_boundaries_reset(C.byref(p.bc))
    if boundaries is not None:
        per, A, B, Cc = boundaries
        A, B, Cc = (np.ascontiguousarray(v, dtype=np.float64) for v in (A, B, Cc))
        L.cvo_set_boundaries(C.byref(p.bc), int(per[0]), int(per[1]), int(per[2]), _ptr(A),
                             _ptr(B), _ptr(Cc))
    return p


def soa(pos):
    """(n,3) -> SoA x..x y..y z..z"""
    return np.ascontiguousarray(np.asarray(pos, dtype=np.float64).T).reshape(-1)


def unsoa(a, n):
    return np.ascontiguousarray(a.reshape(3, n).T)


def coordnum(p, pos1, pos2, *, gradients=True, pairlist=None, rebuild=False,
             g1_center_only=False, g2_center_only=False, mass1=None, mass2=None, omp=0):
    """Reference-order evaluation.  pos1/pos2 are (n,3).  Returns (value, grad1 (n1,3),
    grad2 (n2,3)); pairlist (uint8 array of num_pairs) is updated in place on rebuild."""
    L = lib()
    n1, n2 = len(pos1), len(pos2)
    p1, p2 = soa(pos1), soa(pos2)
    g1 = np.zeros(3 * n1)
    g2 = np.zeros(3 * n2)
    flags = EF_GRADIENTS if gradients else EF_NULL
    plp = None
    if pairlist is not None:
        flags |= EF_USE_PAIRLIST | (EF_REBUILD_PAIRLIST if rebuild else 0)
        assert pairlist.dtype == np.uint8 and pairlist.flags.c_contiguous
        plp = pairlist.ctypes.data_as(C.c_void_p)
    if omp and not (g1_center_only or g2_center_only):
        v = L.cvo_coordnum_omp(C.byref(p), n1, _ptr(p1), n2, _ptr(p2), _ptr(g1), _ptr(g2), plp,
                               flags, int(omp))
        return v, unsoa(g1, n1), unsoa(g2, n2)
    m1 = np.ones(n1) if mass1 is None else np.asarray(mass1, dtype=np.float64)
    m2 = np.ones(n2) if mass2 is None else np.asarray(mass2, dtype=np.float64)
    com1, com2 = np.zeros(3), np.zeros(3)
    # atom_group::update_total_mass: std::accumulate, strictly left to right
    # (src/colvaratoms.cpp:1001).  np.add.accumulate is sequential; Python's sum() is not
    # (compensated summation since 3.12) and numpy's sum() is pairwise.
    tm1 = float(np.add.accumulate(m1)[-1])
    tm2 = float(np.add.accumulate(m2)[-1])
    L.cvo_center_of_mass(n1, _ptr(p1), _ptr(m1), tm1, _ptr(com1))
    L.cvo_center_of_mass(n2, _ptr(p2), _ptr(m2), tm2, _ptr(com2))
    w1 = m1 / tm1
    w2 = m2 / tm2
    v = L.cvo_coordnum(C.byref(p), n1, _ptr(p1), n2, _ptr(p2), int(g1_center_only),
                       int(g2_center_only), _ptr(com1), _ptr(com2), _ptr(w1), _ptr(w2),
                       _ptr(g1), _ptr(g2), plp, flags)
    return v, unsoa(g1, n1), unsoa(g2, n2)


def selfcoordnum(p, pos, *, gradients=True, pairlist=None, rebuild=False, omp=0):
    L = lib()
    n = len(pos)
    ps = soa(pos)
    g = np.zeros(3 * n)
    flags = EF_GRADIENTS if gradients else EF_NULL
    plp = None
    if pairlist is not None:
        flags |= EF_USE_PAIRLIST | (EF_REBUILD_PAIRLIST if rebuild else 0)
        plp = pairlist.ctypes.data_as(C.c_void_p)
    if omp:
        v = L.cvo_selfcoordnum_omp(C.byref(p), n, _ptr(ps), _ptr(g), plp, flags, int(omp))
    else:
        v = L.cvo_selfcoordnum(C.byref(p), n, _ptr(ps), _ptr(g), plp, flags)
    return v, unsoa(g, n)


def switching_function(l2, en=6, ed=12, tol=0.0, flags=EF_GRADIENTS):
    L = lib()
    d = C.c_double(0.0)
    f = L.cvo_switching_function(float(l2), C.byref(d), en, ed, float(tol), flags)
    return f, d.value


def distance_inv(pos1, pos2, exponent=6, boundaries=None, omp=0):
    """distance_inv::calc_value (src/colvarcomp_distances.cpp:405-456).  Returns
    (value, grad1 (n1,3), grad2 (n2,3))."""
    L = lib()
    p = make_params(boundaries=boundaries)
    n1, n2 = len(pos1), len(pos2)
    g1, g2 = np.zeros(3 * n1), np.zeros(3 * n2)
    v = L.cvo_distance_inv(C.byref(p.bc), int(exponent), n1, _ptr(soa(pos1)), n2,
                           _ptr(soa(pos2)), _ptr(g1), _ptr(g2), int(omp))
    return v, unsoa(g1, n1), unsoa(g2, n2)


def distance_pairs(pos1, pos2, force=None, boundaries=None):
    """distance_pairs::calc_value / apply_force (src/colvarcomp_distances.cpp:486-548).  Returns
    the distances (n1 * n2,) and, with a force vector, the forces (n1,3), (n2,3) on the atoms."""
    L = lib()
    p = make_params(boundaries=boundaries)
    n1, n2 = len(pos1), len(pos2)
    dist = np.zeros(n1 * n2)
    if force is None:
        L.cvo_distance_pairs(C.byref(p.bc), n1, _ptr(soa(pos1)), n2, _ptr(soa(pos2)), _ptr(dist),
                             None, None, None)
        return dist
    f1, f2 = np.zeros(3 * n1), np.zeros(3 * n2)
    force = np.ascontiguousarray(force, dtype=np.float64)
    L.cvo_distance_pairs(C.byref(p.bc), n1, _ptr(soa(pos1)), n2, _ptr(soa(pos2)), _ptr(dist),
                         _ptr(force), _ptr(f1), _ptr(f2))
    return dist, unsoa(f1, n1), unsoa(f2, n2)

