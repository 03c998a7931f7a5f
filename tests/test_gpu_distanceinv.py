"""GPU parity tests of the first widening step: colvar::distance_inv
("distanceInv", src/colvarcomp_distances.cpp:365-470) on the tile sweep of the coordination
numbers -- same all-pairs loop, pair term d^-n, scalar epilogue (sum / N1 N2)^(-1/n).

Checker: oracle/coordnum_oracle.c cvo_distance_inv, pinned to the reference's own
distanceinv_harmonic-fixed golden (tests/test_oracle_golden.py).  Bar: 1e-12 relative.
The regression case itself runs in test_gpu_parity.py::test_reference_regression_cases and,
through the unmodified reference host, in test_gpu_host_shim.py.

One caveat the coordination numbers do not have: the reference adds the N1*N2 terms d^-n --
spanning 15+ orders of magnitude -- one after the other into one double, so terms below half
an ulp of the running sum vanish and ITS result carries an error of up to N1*N2*2^-53 relative
(measured: 1.1e-11 on the sum for 1030 x 2100 atoms, exponent 8).  The CUDA path sums per
lane, per warp, per CTA and is far closer to the exact sum.  Every test therefore checks
(a) against the oracle's compensated variant (what the reference's loop gives without its
summation error) at 1e-12, and (b) against the reference-order oracle at 1e-12 widened by
that a-priori bound.
"""
import numpy as np
import pytest

from . import oracle_lib as O
from .test_gpu_parity import RTOL, _cb, close_arr, close_val, rand_case

pytestmark = pytest.mark.gpu


def check_against_both(v, g1, g2, pos1, pos2, exponent, boundaries=None, sequential=True):
    n1, n2 = len(pos1), len(pos2)
    av, ag1, ag2 = O.distance_inv(pos1, pos2, exponent, boundaries=boundaries, omp=8)
    close_val(v, av)
    close_arr(g1, ag1)
    close_arr(g2, ag2)
    if sequential:
        ov, og1, og2 = O.distance_inv(pos1, pos2, exponent, boundaries=boundaries)
        bound = n1 * n2 * 2.0 ** -53  # on the sum; value: / exponent, gradients: * (n+1)/n
        close_val(v, ov, max(RTOL, bound / exponent))
        close_arr(g1, og1, max(RTOL, 2 * bound))
        close_arr(g2, og2, max(RTOL, 2 * bound))
        return ov, og1, og2
    return av, ag1, ag2


@pytest.mark.parametrize("exponent", [2, 4, 6, 8, 12])
@pytest.mark.parametrize("n1,n2", [(1, 1), (3, 70), (33, 65), (257, 1000), (1030, 2100)])
def test_distanceinv_random_sizes(n1, n2, exponent):
    cb = _cb()
    proxy, i1, i2 = rand_case(n1, n2, seed=5 * n1 + n2 + exponent, L=9.0 + 0.01 * n2)
    cv = cb.CoordNum("distanceInv", i1, i2, exponent=exponent)
    v = cv.calc_host(proxy)
    check_against_both(v, cv.gradients(1), cv.gradients(2), proxy[i1], proxy[i2], exponent)
    # the value is a generalised mean of the pair distances: between the shortest and longest
    d = np.linalg.norm(proxy[i1][:, None, :] - proxy[i2][None, :, :], axis=2)
    assert d.min() * (1 - 1e-12) <= v <= d.max() * (1 + 1e-12)


def test_distanceinv_value_only_and_force_scatter():
    cb = _cb()
    proxy, i1, i2 = rand_case(300, 500, seed=4, L=15.0)
    cv = cb.CoordNum("distanceInv", i1, i2, exponent=6)
    v0 = cv.calc_host(proxy, gradients=False)
    assert np.all(cv.gradients(1) == 0.0) and np.all(cv.gradients(2) == 0.0)
    v = cv.calc_host(proxy, gradients=True)
    assert v == v0
    ov, og1, og2 = check_against_both(v, cv.gradients(1), cv.gradients(2), proxy[i1], proxy[i2],
                                      6)
    pf = np.zeros_like(proxy)
    cv.apply_force_host(-2.5, pf)
    close_arr(pf[i1], -2.5 * og1)
    close_arr(pf[i2], -2.5 * og2)
    # translation invariance: the total force vanishes
    assert np.max(np.abs(pf.sum(0))) <= 1e-10 * np.max(np.abs(pf)) * len(proxy)


@pytest.mark.parametrize("bc", ["ortho", "triclinic", "mixed"])
def test_distanceinv_periodic_boundaries(bc):
    cb = _cb()
    L = 18.0
    n1, n2 = 150, 700
    rng = np.random.default_rng(23)
    proxy = np.ascontiguousarray((rng.random((n1 + n2, 3)) * 3.0 - 1.0) * L)
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32)
    if bc == "ortho":
        per, A, B, Cc = (1, 1, 1), (L, 0, 0), (0, 1.1 * L, 0), (0, 0, 0.9 * L)
    elif bc == "triclinic":
        per, A, B, Cc = (1, 1, 1), (L, 0, 0), (0.3 * L, L, 0), (0.1 * L, -0.2 * L, L)
    else:
        per, A, B, Cc = (1, 1, 0), (L, 0, 0), (0, L, 0), (0, 0, L)
    cv = cb.CoordNum("distanceInv", i1, i2, exponent=4)
    cv.set_boundaries(per, A, B, Cc)
    v = cv.calc_host(proxy)
    check_against_both(v, cv.gradients(1), cv.gradients(2), proxy[i1], proxy[i2], 4,
                       boundaries=(per, A, B, Cc))


def test_distanceinv_input_errors_follow_the_reference():
    cb = _cb()
    i1 = np.arange(4, dtype=np.int32)
    i2 = np.arange(4, 9, dtype=np.int32)
    with pytest.raises(cb.B200cvError, match="odd exponent provided"):
        cb.CoordNum("distanceInv", i1, i2, exponent=3)
    with pytest.raises(cb.B200cvError, match="negative or zero exponent"):
        cb.CoordNum("distanceInv", i1, i2, exponent=0)
    with pytest.raises(cb.B200cvError, match="atoms in common"):
        cb.CoordNum("distanceInv", i1, np.arange(3, 8, dtype=np.int32), exponent=6)


def test_distanceinv_device_resident_graph_step():
    """positions already in HBM: fused read_data + calc_value (CUDA graph replay) and the
    device force scatter, against the host-buffer path of the same handle."""
    import torch
    cb = _cb()
    n1, n2 = 2000, 30000
    proxy, i1, i2 = rand_case(n1, n2, seed=12, L=40.0)
    cv = cb.CoordNum("distanceInv", i1, i2, exponent=6)
    v_host = cv.calc_host(proxy)
    g1, g2 = cv.gradients(1).copy(), cv.gradients(2).copy()
    dpos = torch.tensor(np.ascontiguousarray(proxy.T), device="cuda")  # SoA x..x y..y z..z
    for step in range(3):
        cv.step(dpos, step=step)
        close_val(cv.value(), v_host)
    close_arr(cv.gradients(1), g1)
    close_arr(cv.gradients(2), g2)
    dforce = torch.zeros_like(dpos)
    cv.apply_force(0.75, dforce)
    torch.cuda.synchronize()
    f = dforce.cpu().numpy().T
    close_arr(f[i1], 0.75 * g1)
    close_arr(f[i2], 0.75 * g2)


def test_distanceinv_large_against_parallel_oracle():
    """10k x 300k = 3e9 pairs (many work items per SM) against the OpenMP oracle (compensated
    sum; the sequential loop would take minutes and lose ~1e-11)."""
    cb = _cb()
    n1, n2 = 10000, 300000
    proxy, i1, i2 = rand_case(n1, n2, seed=31, L=160.0)
    cv = cb.CoordNum("distanceInv", i1, i2, exponent=6)
    v = cv.calc_host(proxy)
    check_against_both(v, cv.gradients(1), cv.gradients(2), proxy[i1], proxy[i2], 6,
                       sequential=False)
