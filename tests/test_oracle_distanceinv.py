"""CPU checks of the distanceInv restatement (oracle/coordnum_oracle.c cvo_distance_inv;
src/colvarcomp_distances.cpp:405-456).  The golden text of the reference's
distanceinv_harmonic-fixed case is covered by test_oracle_golden.py; here: the analytic
gradient against central differences, and the compensated variant against the
reference-order one (they differ only by the reference's own summation error)."""
import numpy as np
import pytest

from . import oracle_lib as O


@pytest.mark.parametrize("exponent", [2, 6, 10])
def test_gradients_are_the_derivative_of_the_value(exponent):
    rng = np.random.default_rng(exponent)
    p1 = rng.random((5, 3)) * 6.0
    p2 = rng.random((9, 3)) * 6.0 + 1.0
    v, g1, g2 = O.distance_inv(p1, p2, exponent)
    h = 1e-6
    for grp, pos, g in ((1, p1, g1), (2, p2, g2)):
        for a in range(len(pos)):
            for d in range(3):
                q = pos.copy()
                q[a, d] += h
                vp = O.distance_inv(q if grp == 1 else p1, p2 if grp == 1 else q, exponent)[0]
                q[a, d] -= 2 * h
                vm = O.distance_inv(q if grp == 1 else p1, p2 if grp == 1 else q, exponent)[0]
                assert abs((vp - vm) / (2 * h) - g[a, d]) <= 1e-7 * max(1.0, np.max(np.abs(g)))


def test_compensated_variant_differs_only_by_the_sequential_summation_error():
    rng = np.random.default_rng(3)
    n1, n2, exponent = 400, 3000, 8
    p1 = rng.random((n1, 3)) * 25.0
    p2 = rng.random((n2, 3)) * 25.0
    v, g1, g2 = O.distance_inv(p1, p2, exponent)
    va, ga1, ga2 = O.distance_inv(p1, p2, exponent, omp=4)
    bound = n1 * n2 * 2.0 ** -53
    assert abs(v - va) <= bound / exponent * abs(va)
    assert np.max(np.abs(g1 - ga1)) <= 2 * bound * np.max(np.abs(ga1))
    assert np.max(np.abs(g2 - ga2)) <= 2 * bound * np.max(np.abs(ga2))
    # exact sum in extended precision from the same pair terms
    d2 = ((p1[:, None, :] - p2[None, :, :]) ** 2).sum(-1)
    s = np.sum((1.0 / d2 ** (exponent // 2)).astype(np.longdouble))
    exact = float((s / (n1 * n2)) ** np.longdouble(-1.0 / exponent))
    assert abs(va - exact) <= 1e-14 * exact
