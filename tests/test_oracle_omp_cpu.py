"""The OpenMP row-parallel checkers of the oracle (used by the BASELINE-size GPU parity tests,
tests/test_gpu_baseline_sizes.py) against the sequential, reference-order oracle that the
golden vectors pin: same pair values, so identical pair-list decisions, gradients equal up to
summation order, scalar equal up to the rounding of the reference's own running sum."""
import numpy as np
import pytest

from . import oracle_lib as O


@pytest.mark.parametrize("tol", [0.0, 0.01])
def test_coordnum_omp_equals_sequential(tol):
    rng = np.random.default_rng(5)
    p1, p2 = rng.random((257, 3)) * 24.0, rng.random((1111, 3)) * 24.0
    p = O.make_params((3.0, 5.0, 4.0), 6, 12, tol)
    kw = {}
    pls = [None, None]
    if tol > 0:
        pls = [np.ones(257 * 1111, dtype=np.uint8) for _ in range(2)]
    for rebuild in ([True, False] if tol > 0 else [False]):
        a = O.coordnum(p, p1, p2, pairlist=pls[0], rebuild=rebuild, **kw)
        b = O.coordnum(p, p1, p2, pairlist=pls[1], rebuild=rebuild, omp=3, **kw)
        assert abs(a[0] - b[0]) <= 1e-13 * abs(a[0])
        for x, y in zip(a[1:], b[1:]):
            assert np.max(np.abs(x - y)) <= 1e-13 * np.max(np.abs(x))
        if tol > 0:
            assert np.array_equal(pls[0], pls[1])
        p1 = p1 + rng.normal(0, 0.05, p1.shape)


@pytest.mark.parametrize("tol", [0.0, 0.02])
def test_selfcoordnum_omp_equals_sequential(tol):
    rng = np.random.default_rng(6)
    pos = rng.random((901, 3)) * 20.0
    p = O.make_params((4.0,) * 3, 6, 12, tol)
    pls = [None, None]
    if tol > 0:
        pls = [np.ones(901 * 900 // 2, dtype=np.uint8) for _ in range(2)]
    for rebuild in ([True, False] if tol > 0 else [False]):
        a = O.selfcoordnum(p, pos, pairlist=pls[0], rebuild=rebuild)
        b = O.selfcoordnum(p, pos, pairlist=pls[1], rebuild=rebuild, omp=3)
        assert abs(a[0] - b[0]) <= 1e-13 * abs(a[0])
        assert np.max(np.abs(a[1] - b[1])) <= 1e-13 * np.max(np.abs(a[1]))
        if tol > 0:
            assert np.array_equal(pls[0], pls[1])
        pos = pos + rng.normal(0, 0.05, pos.shape)


def test_running_double_sum_loses_small_terms_at_scale():
    """Why the checkers sum the scalar in extended precision: a running double sum of many
    positive terms that span orders of magnitude (the reference's `x.real_value += partial`)
    differs from the exact sum by far more than 1e-12 once there are ~1e8 terms.  Pure
    arithmetic, no oracle call: terms (r0/r)^6-like, 2e7 of them."""
    rng = np.random.default_rng(0)
    t = (4.0 / (4.0 + 300.0 * rng.random(20_000_000))) ** 6
    seq = float(np.add.accumulate(t)[-1])          # strictly sequential double sum
    exact = float(np.sum(t.astype(np.longdouble)))
    pairwise = float(np.sum(t))
    assert abs(pairwise - exact) <= 1e-14 * exact
    assert abs(seq - exact) > 1e-13 * exact
