"""Parity at the BASELINE.json sizes themselves (configs 3, 4, 5), exactly as bench.py runs
them: device-resident SoA proxy buffers -> read_data -> calc_value (-> fit gradients) ->
apply_force, against the OpenMP row-parallel form of the oracle (same pair arithmetic as the
reference-order oracle, per-row partial sums, scalar summed in long double;
tests/test_oracle_omp_cpu.py pins it to the reference-order oracle).

Bar: pair-list indices bit-exact at every rebuild; values, gradients and forces within 1e-12
(max-norm scaled for arrays, see test_gpu_parity.py).  Fitted groups: geometry 1e-13 against
the oracle's own fit, pair arithmetic 1e-12 on the positions the device produced, fit
gradients 1e-11 through the oracle's chain rule (test_gpu_parity.py::_check_fitted_case
explains the decomposition).
"""
import os

import numpy as np
import pytest

from . import oracle_lib as O

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _cb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import colvars_b200 as cb
    cb.load()
    return cb


def _threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def test_c3_selfcoordnum_50k_pairlist_cycle_with_moving_atoms():
    """BASELINE config 3 as benchmarked: selfCoordNum on 50 k atoms at
    water density, tolerance 1e-3, pairListFrequency 10, 20 steps with Gaussian displacements
    of 0.05 A per step.  Step 0 rebuilds through the all-pairs sweep, step 10 through the cell
    list; the 18 other steps walk the list.  The reference's dense bool list is 1.25 GB."""
    import torch
    from colvars_b200 import synth
    cb = _cb()
    n, tol, freq = 50000, 1e-3, 10
    proxy, i1, L = synth.selfcoordnum_case(n, seed=1234)
    cv = cb.CoordNum("selfCoordNum", i1, cutoff=4.0, tolerance=tol, pairlist_freq=freq)
    p = O.make_params((4.0,) * 3, 6, 12, tol)
    assert cv.tolerance_l2_max() == p.tolerance_l2_max
    pl = np.ones(n * (n - 1) // 2, dtype=np.uint8)
    nthr = _threads()
    pos = proxy.copy()
    d_force = torch.zeros((3, n), dtype=torch.float64, device="cuda")
    worst = 0.0
    launches = []
    for step in range(2 * freq):
        rebuild = step % freq == 0
        d_pos = torch.tensor(np.ascontiguousarray(pos.T), device="cuda")
        cv.read_data(d_pos)
        cv.calc_value(step=step, gradients=True)
        v = cv.value()
        fa = -0.004 * (v - 0.1)
        d_force.zero_()
        cv.apply_force(fa, d_force)
        ov, og = O.selfcoordnum(p, pos, pairlist=pl, rebuild=rebuild, omp=nthr)
        assert abs(v - ov) <= RTOL * abs(ov), (step, v, ov)
        g = cv.gradients(1)
        e = _rel(g, og)
        worst = max(worst, e)
        assert e <= RTOL, (step, e)
        torch.cuda.synchronize()
        assert _rel(d_force.cpu().numpy().T, fa * og) <= RTOL, step
        if rebuild:
            mine = cv.pairlist()
            ref = np.flatnonzero(pl).astype(np.uint64)
            assert mine.shape == ref.shape and np.array_equal(mine, ref), step
            launches.append(cv.last_stats()[0])
        pos = synth.displace(pos, seed=1000 + step, sigma=0.05)
    # the second rebuild went through the cell list (more launches than sweep + exact queue)
    assert launches[1] > launches[0], launches


@pytest.mark.parametrize("kind", ["selfCoordNum", "coordNum"])
def test_rows_longer_than_one_chunk(kind):
    """A core of 1300 atoms inside one list radius in a dilute gas of 24 k: the list is sparse
    (0.2 % of the pairs), but the core atoms have more than a thousand listed partners each --
    such a row is written as several chunks (ROW_BUF = 512 entries each, csrc/rows.cuh: the row's own
    slot and overflow chunks behind the per-row slots).  Value, gradients and the bit-exact
    list over a rebuild through the all-pairs sweep, list steps, a rebuild through the rows
    and more list steps."""
    cb = _cb()
    rng = np.random.default_rng(31)
    core = rng.normal(0.0, 1.6, (1300, 3)) + 150.0
    gas = rng.random((24000, 3)) * 300.0
    if kind == "selfCoordNum":
        pos0 = np.concatenate([core, gas])
        rng.shuffle(pos0)
        n1, n2 = len(pos0), 0
    else:
        g1 = np.concatenate([core[:150], gas[:350]])
        g2 = np.concatenate([core[150:], gas[350:]])
        rng.shuffle(g1)
        rng.shuffle(g2)
        pos0 = np.concatenate([g1, g2])
        n1, n2 = len(g1), len(g2)
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32) if n2 else None
    tol, freq = 0.01, 3
    cv = cb.CoordNum(kind, i1, i2, cutoff=4.0, tolerance=tol, pairlist_freq=freq)
    p = O.make_params((4.0,) * 3, 6, 12, tol)
    npairs = n1 * (n1 - 1) // 2 if kind == "selfCoordNum" else n1 * n2
    pl = np.ones(npairs, dtype=np.uint8)
    nthr = _threads()
    forms = []
    for step in range(2 * freq + 1):
        frame = np.ascontiguousarray(pos0 + rng.normal(0, 0.03, pos0.shape) * step)
        v = cv.calc_host(frame, step=step)
        rebuild = step % freq == 0
        if kind == "selfCoordNum":
            ov, og1 = O.selfcoordnum(p, frame, pairlist=pl, rebuild=rebuild, omp=nthr)
        else:
            ov, og1, og2 = O.coordnum(p, frame[i1], frame[i2], pairlist=pl, rebuild=rebuild,
                                      omp=nthr)
            assert _rel(cv.gradients(2), og2) <= RTOL, step
        assert abs(v - ov) <= RTOL * abs(ov), (step, v, ov)
        assert _rel(cv.gradients(1), og1) <= RTOL, step
        if rebuild:
            assert np.array_equal(np.sort(cv.pairlist()), np.flatnonzero(pl).astype(np.uint64)), step
            forms.append(cv.pairlist_stats())
    assert pl.sum() * 50 < npairs  # sparse
    st = forms[-1]
    assert st["form"] == 2, st                       # neighbour rows
    assert st["row_chunks"] > n1 + 200, st            # ... with overflow chunks (one slot per row)


def test_c4_aniso_fitted_10k_x_500k():
    """BASELINE config 4 as benchmarked: coordNum 10 k x 500 k, cutoff3 (3, 6, 4), group1
    centred and rotated onto its reference positions (fitted on itself, fit gradients on);
    the frame is the reference solute rigidly rotated plus 0.05 A noise.  Whole value and all
    gradients of both groups (5e9 pairs through the OpenMP oracle), fit gradients, forces."""
    import torch
    from colvars_b200 import synth
    cb = _cb()
    n1, n2 = 10000, 500000
    cutoff3 = (3.0, 6.0, 4.0)
    proxy, i1, i2, L = synth.coordnum_case(n1, n2, seed=1234)
    ref = proxy[i1].copy()
    ctr = ref.mean(0)
    proxy[i1] = (ref - ctr) @ synth.random_rotation(7).T + ctr + \
        np.random.default_rng(99).normal(0, 0.05, ref.shape)
    cv = cb.CoordNum("coordNum", i1, i2, cutoff3=cutoff3)
    cv.set_fitting(1, center=True, rotate=True, fitting_group=None, ref_pos=ref,
                   enable_fit_gradients=True)
    d_pos = torch.tensor(np.ascontiguousarray(proxy.T), device="cuda")
    d_force = torch.zeros_like(d_pos)
    cv.read_data(d_pos)
    cv.calc_value(step=0, gradients=True)
    cv.calc_fit_gradients()
    v = cv.value()
    fa = -0.004 * (v - 0.1)
    cv.apply_force(fa, d_force)
    torch.cuda.synchronize()
    # 1. geometry: the oracle's own fit of the same frame
    fit = O.Fit(ref, True, True)
    ofitted, unrot, _ = fit.apply(proxy[i1])
    fitted1 = cv.positions(1)
    assert _rel(fitted1, ofitted) <= 1e-13
    q, qr = cv.rotation(1), fit.q
    assert min(np.abs(q - qr).max(), np.abs(q + qr).max()) < 1e-13
    # 2. pair arithmetic on the device's fitted positions, every pair
    p = O.make_params(cutoff3)
    ov, og1, og2 = O.coordnum(p, fitted1, proxy[i2], omp=_threads())
    assert abs(v - ov) <= RTOL * abs(ov), (v, ov)
    g1, g2 = cv.gradients(1), cv.gradients(2)
    assert _rel(g1, og1) <= RTOL
    assert _rel(g2, og2) <= RTOL
    # 3. fit gradients through the oracle's chain rule on the device's gradients
    ofg = fit.fit_gradients(g1, unrot, n1)
    fg = cv.fit_gradients(1, n1)
    scale = max(np.abs(ofg).max(), np.abs(g1).max())
    assert np.abs(fg - ofg).max() <= 1e-11 * scale
    # 4. forces: group1 gets R^-1 grad + fit gradients, group2 its gradients
    #    (src/colvaratoms.cpp:1665-1703)
    f = d_force.cpu().numpy().T
    exp1 = fa * (g1 @ fit.R + ofg)          # R^-1 g = R^T g, row vectors: g @ R
    assert _rel(f[i1], exp1) <= 1e-11
    assert _rel(f[i2], fa * og2) <= RTOL


def test_c5_whole_value_and_gradients_100k_x_1M():
    """BASELINE config 5 (the headline workload): all 1e11 pairs once through the OpenMP
    oracle (about a minute and a half on 16 cores) against one device-resident step.  The
    checker adds the 1e11 pair values in extended precision: a running double sum -- the
    reference's own `x.real_value += partial` as well -- loses the low bits of every small term
    (a 16-thread double sum of these very pairs is 2.5e-10 away from this one, first GPU run of
    round 2), so at this size "the reference's value" is only defined to ~1e-10 and the bar is
    the correctly rounded sum of the reference's pair values (tests/test_oracle_omp_cpu.py)."""
    import torch
    from colvars_b200 import synth
    cb = _cb()
    n1, n2 = 100000, 1000000
    proxy, i1, i2, L = synth.coordnum_case(n1, n2, seed=1234)
    cv = cb.CoordNum("coordNum", i1, i2, cutoff=4.0)
    d_pos = torch.tensor(np.ascontiguousarray(proxy.T), device="cuda")
    d_force = torch.zeros_like(d_pos)
    cv.read_data(d_pos)
    cv.calc_value(step=0, gradients=True)
    v = cv.value()
    fa = -0.004 * (v - 0.1)
    cv.apply_force(fa, d_force)
    torch.cuda.synchronize()
    p = O.make_params((4.0,) * 3)
    ov, og1, og2 = O.coordnum(p, proxy[i1], proxy[i2], omp=_threads())
    assert abs(v - ov) <= RTOL * abs(ov), (v, ov)
    assert _rel(cv.gradients(1), og1) <= RTOL
    assert _rel(cv.gradients(2), og2) <= RTOL
    f = d_force.cpu().numpy().T
    assert _rel(f[i1], fa * og1) <= RTOL
    assert _rel(f[i2], fa * og2) <= RTOL
