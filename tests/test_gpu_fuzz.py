"""Seeded fuzz of the whole configuration space of the path against the oracle: component kind
x group sizes x iso / anisotropic cut-off x exponent pair x pair list (tolerance, frequency) x
boundary conditions x centre-only flags x masses, three steps each with moving atoms.  Every
combination the keyword parser accepts (src/colvarcomp_coordnums.cpp:46-156) must give the
reference's numbers: values and gradients to 1e-12 (1e-11 where a centre of mass feeds the
reference's ill-conditioned quotient), pair lists bit-exact.
"""
import numpy as np
import pytest

from . import oracle_lib as O
from .test_gpu_parity import RTOL, _cb, close_arr

pytestmark = pytest.mark.gpu

EXPONENTS = [(6, 12), (6, 12), (4, 8), (6, 8), (2, 4), (8, 16), (10, 20), (6, 10), (12, 24)]
CELLS = {
    "none": None,
    "ortho": ((1, 1, 1), (22.0, 0, 0), (0, 25.0, 0), (0, 0, 20.0)),
    "triclinic": ((1, 1, 1), (22.0, 0, 0), (6.0, 21.0, 0), (2.0, -4.0, 23.0)),
    "mixed": ((1, 0, 1), (22.0, 0, 0), (0, 22.0, 0), (0, 0, 22.0)),
}


def draw(seed):
    rng = np.random.default_rng(1000 + seed)
    kind = rng.choice(["coordNum", "coordNum", "selfCoordNum", "groupCoord", "distanceInv"])
    n1 = int(rng.choice([1, 2, 31, 33, 64, 200, 700]))
    n2 = 0 if kind == "selfCoordNum" else int(rng.choice([1, 5, 32, 97, 400, 1500]))
    if kind == "selfCoordNum":
        n1 = max(n1, 2)
    cfg = dict(kind=kind, n1=n1, n2=n2, seed=seed)
    cfg["cutoff3"] = (4.0,) * 3 if rng.random() < 0.5 else tuple(rng.uniform(2.5, 6.0, 3))
    cfg["en"], cfg["ed"] = EXPONENTS[rng.integers(len(EXPONENTS))]
    cfg["tol"] = float(rng.choice([0.0, 0.0, 0.01, 0.001, 0.05])) if kind in ("coordNum",
                                                                                "selfCoordNum") else 0.0
    cfg["freq"] = int(rng.choice([1, 2, 3]))
    cfg["cell"] = str(rng.choice(list(CELLS)))
    cfg["g1c"] = bool(kind == "coordNum" and rng.random() < 0.2)
    cfg["g2c"] = bool(kind == "coordNum" and rng.random() < 0.2)
    cfg["masses"] = bool(rng.random() < 0.5)
    cfg["exponent"] = int(rng.choice([2, 4, 6, 8]))
    return cfg, rng


@pytest.mark.parametrize("seed", range(120))
def test_random_configuration_matches_the_oracle(seed):
    cb = _cb()
    cfg, rng = draw(seed)
    kind, n1, n2 = cfg["kind"], cfg["n1"], cfg["n2"]
    L = 20.0
    pos = rng.random((n1 + n2, 3)) * L
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32) if n2 else None
    m1 = rng.uniform(1.0, 16.0, n1) if cfg["masses"] else None
    m2 = rng.uniform(1.0, 16.0, n2) if (cfg["masses"] and n2) else None
    cell = CELLS[cfg["cell"]]
    if kind == "distanceInv":
        cv = cb.CoordNum(kind, i1, i2, exponent=cfg["exponent"], mass1=m1, mass2=m2)
    else:
        cv = cb.CoordNum(kind, i1, i2, cutoff3=cfg["cutoff3"], exp_numer=cfg["en"],
                         exp_denom=cfg["ed"], tolerance=cfg["tol"], pairlist_freq=cfg["freq"],
                         group1_center_only=cfg["g1c"], group2_center_only=cfg["g2c"], mass1=m1,
                         mass2=m2)
    if cell:
        cv.set_boundaries(*cell)
    p = O.make_params(cfg["cutoff3"], cfg["en"], cfg["ed"], cfg["tol"], boundaries=cell)
    centre = kind == "groupCoord" or cfg["g1c"] or cfg["g2c"]
    g1c = cfg["g1c"] or kind == "groupCoord"
    g2c = cfg["g2c"] or kind == "groupCoord"
    use_pl = cfg["tol"] > 0
    if kind == "selfCoordNum":
        npairs = n1 * (n1 - 1) // 2
    else:
        npairs = (1 if g1c else n1) * (1 if g2c else n2)
    pl = np.ones(max(npairs, 1), dtype=np.uint8) if use_pl else None
    tol = 1e-11 if centre else RTOL
    for step in range(3):
        frame = pos + rng.normal(0, 0.05, pos.shape) * step
        v = cv.calc_host(np.ascontiguousarray(frame), step=step)
        rebuild = use_pl and step % cfg["freq"] == 0
        if kind == "distanceInv":
            ov, og1, og2 = O.distance_inv(frame[i1], frame[i2], cfg["exponent"], boundaries=cell)
        elif kind == "selfCoordNum":
            ov, og1 = O.selfcoordnum(p, frame[i1], pairlist=pl, rebuild=rebuild)
            og2 = None
        else:
            ov, og1, og2 = O.coordnum(p, frame[i1], frame[i2], pairlist=pl, rebuild=rebuild,
                                      g1_center_only=g1c, g2_center_only=g2c, mass1=m1, mass2=m2)
        scale = max(abs(ov), 1e-9 * max(npairs, 1))  # a value that is (nearly) zero: absolute bar
        assert abs(v - ov) <= tol * scale, (cfg, step, v, ov)
        for got, want in ((cv.gradients(1), og1), (cv.gradients(2) if n2 else None, og2)):
            if want is None:
                continue
            if np.max(np.abs(want)) < 1e-200:
                assert np.max(np.abs(got)) < 1e-200, cfg
            else:
                close_arr(got, want, tol)
    if use_pl and not centre:
        # list contents after the last rebuild, bit-exact canonical indices
        assert np.array_equal(np.sort(cv.pairlist()), np.flatnonzero(pl).astype(np.uint64)), cfg


import os

NSPARSE = int(os.environ.get("B200CV_FUZZ_SPARSE", "16"))


def draw_sparse(seed):
    """a system whose pair list is sparse (well below 2 % of the pairs), so that rebuilds after the
    first go through the cell order: neighbour rows in an open system, cell-ordered entries in a
    periodic cell"""
    rng = np.random.default_rng(77000 + seed)
    kind = str(rng.choice(["selfCoordNum", "coordNum"]))
    if kind == "selfCoordNum":
        n1, n2 = int(rng.integers(1500, 5000)), 0
    else:
        n1, n2 = int(rng.integers(40, 1200)), int(rng.integers(2500, 7000))
    cfg = dict(kind=kind, n1=n1, n2=n2, seed=seed)
    cfg["cutoff3"] = (3.5,) * 3 if rng.random() < 0.5 else tuple(rng.uniform(2.5, 4.0, 3))
    cfg["tol"] = float(rng.choice([0.002, 0.01, 0.03]))
    cfg["freq"] = int(rng.choice([1, 2, 3]))
    cfg["cell"] = str(rng.choice(["none", "none", "ortho", "triclinic"]))
    cfg["shape"] = str(rng.choice(["uniform", "blobs", "slab"]))
    cfg["sigma"] = float(rng.choice([0.02, 0.1, 0.3]))
    return cfg, rng


@pytest.mark.parametrize("seed", range(NSPARSE))
def test_random_sparse_pairlist_configuration_matches_the_oracle(seed):
    cb = _cb()
    cfg, rng = draw_sparse(seed)
    kind, n1, n2 = cfg["kind"], cfg["n1"], cfg["n2"]
    n = n1 + n2
    L = 90.0
    if cfg["shape"] == "uniform":
        pos = rng.random((n, 3)) * L
    elif cfg["shape"] == "blobs":
        centres = rng.random((12, 3)) * L
        pos = centres[rng.integers(12, size=n)] + rng.normal(0, 15.0, (n, 3))
    else:
        pos = rng.random((n, 3)) * np.array([L, L, 40.0])
        pos[:, 2] += 0.25 * L
    if kind == "coordNum":
        rng.shuffle(pos)
    cell = None
    if cfg["cell"] == "ortho":
        cell = ((1, 1, 1), (L, 0, 0), (0, L * 1.1, 0), (0, 0, L * 0.9))
    elif cfg["cell"] == "triclinic":
        cell = ((1, 1, 1), (L, 0, 0), (0.3 * L, L, 0), (0.1 * L, -0.2 * L, L))
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n, dtype=np.int32) if n2 else None
    cv = cb.CoordNum(kind, i1, i2, cutoff3=cfg["cutoff3"], tolerance=cfg["tol"],
                     pairlist_freq=cfg["freq"])
    if cell:
        cv.set_boundaries(*cell)
    p = O.make_params(cfg["cutoff3"], 6, 12, cfg["tol"], boundaries=cell)
    npairs = n1 * (n1 - 1) // 2 if kind == "selfCoordNum" else n1 * n2
    pl = np.ones(npairs, dtype=np.uint8)
    for step in range(2 * cfg["freq"] + 2):
        frame = np.ascontiguousarray(pos + rng.normal(0, cfg["sigma"], pos.shape) * step)
        v = cv.calc_host(frame, step=step)
        rebuild = step % cfg["freq"] == 0
        if kind == "selfCoordNum":
            ov, og1 = O.selfcoordnum(p, frame[i1], pairlist=pl, rebuild=rebuild)
            og2 = None
        else:
            ov, og1, og2 = O.coordnum(p, frame[i1], frame[i2], pairlist=pl, rebuild=rebuild)
        scale = max(abs(ov), 1e-9 * npairs)
        assert abs(v - ov) <= RTOL * scale, (cfg, step, v, ov)
        close_arr(cv.gradients(1), og1)
        if og2 is not None:
            close_arr(cv.gradients(2), og2)
        if rebuild:
            assert np.array_equal(np.sort(cv.pairlist()),
                                  np.flatnonzero(pl).astype(np.uint64)), (cfg, step)
    # the list was sparse, so the later rebuilds went through the cell order: neighbour rows (2) in
    # an open system, cell-ordered entries (1) in a periodic cell
    if pl.sum() * 64 < npairs:
        assert cv.pairlist_stats()["form"] == (2 if cell is None else 1), (cfg, cv.pairlist_stats())
