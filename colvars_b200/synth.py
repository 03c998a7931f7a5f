"""Seeded synthetic solute/solvent coordinates for the benchmark configurations
(BASELINE.json `configs`).  Pure numpy; no compute path here.

* "water O": uniform in a cube of side L = (N2 / 0.0334 A^-3)^(1/3) (number density of
  water; L ~ 144 A for 100 k atoms, ~311 A for 1 M);
* "solute": Gaussian blob centred at L/2 with sigma = 8 A for 1 k atoms, scaled as N^(1/3);
* no two atoms of interacting groups coincide (continuous distributions; l2 == 0 has
  probability zero and is checked for the small cases in tests).
"""
import numpy as np

WATER_DENSITY = 0.0334  # atoms / A^3


def box_length(n_solvent):
    return float((n_solvent / WATER_DENSITY) ** (1.0 / 3.0))


def solvent(n, seed, L=None):
    rng = np.random.default_rng(seed)
    L = box_length(n) if L is None else L
    return rng.random((n, 3)) * L


def solute(n, seed, L, sigma=None):
    rng = np.random.default_rng(seed + 7919)
    if sigma is None:
        sigma = 8.0 * (n / 1000.0) ** (1.0 / 3.0)
    return rng.normal(0.0, sigma, (n, 3)) + 0.5 * L


def coordnum_case(n1, n2, seed=1234):
    """group1 = solute blob, group2 = water O; proxy array = [solute | solvent]."""
    L = box_length(n2)
    p1 = solute(n1, seed, L)
    p2 = solvent(n2, seed, L)
    proxy = np.ascontiguousarray(np.concatenate([p1, p2], axis=0))
    idx1 = np.arange(n1, dtype=np.int32)
    idx2 = np.arange(n1, n1 + n2, dtype=np.int32)
    return proxy, idx1, idx2, L


def selfcoordnum_case(n, seed=1234):
    L = box_length(n)
    proxy = np.ascontiguousarray(solvent(n, seed, L))
    return proxy, np.arange(n, dtype=np.int32), L


def displace(pos, seed, sigma=0.05):
    """Gaussian displacement between steps (so that a pair list matters)."""
    rng = np.random.default_rng(seed)
    return pos + rng.normal(0.0, sigma, pos.shape)


def random_rotation(seed):
    rng = np.random.default_rng(seed)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    q0, q1, q2, q3 = q
    return np.array([
        [q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * (q1 * q2 - q0 * q3), 2 * (q0 * q2 + q1 * q3)],
        [2 * (q0 * q3 + q1 * q2), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * (q2 * q3 - q0 * q1)],
        [2 * (q1 * q3 - q0 * q2), 2 * (q0 * q1 + q2 * q3), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3],
    ])


# pair counts and the algorithmic flop figure of SURVEY.md section 8d / BASELINE.md section 3
FLOP_PER_PAIR_VALUE_GRAD = 39
FLOP_PER_PAIR_VALUE = 19
# distanceInv, exponent 6, value + gradients: 3 differences, d^2 (1 mul + 2 fma), reciprocal
# refinement (3 fma), r^3 (2 mul), F r (1 mul), sum (1 add), six gradient fma = 30 flop
FLOP_PER_PAIR_DISTANCEINV = 30


def num_pairs(kind, n1, n2=0):
    return n1 * (n1 - 1) // 2 if kind == "selfCoordNum" else n1 * n2
