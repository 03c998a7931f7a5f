"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bar (BASELINE.json north_star): pair-list indices bit-exact; values,
gradients and forces within 1e-12 relative in FP64.

Metric for arrays (SURVEY.md section 7 "hard parts"): per-atom gradients are sums of many
signed pair terms; summation order differs from the CPU loop, so the error of an element is
bounded relative to the largest element of the array:  max|a - b| <= RTOL * max|b|.
"""
import numpy as np
import pytest

from . import oracle_lib as O
from . import regression as R

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def _cb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import colvars_b200 as cb
    cb.load()
    return cb


def close_arr(a, b, rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(np.max(np.abs(b)), 1e-300) if b.size else 1.0
    err = np.max(np.abs(a - b)) if b.size else 0.0
    assert err <= rtol * scale, f"max abs err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e})"


def close_val(a, b, rtol=RTOL):
    assert abs(a - b) <= rtol * max(abs(b), 1e-300), f"{a!r} vs {b!r} rel {abs(a - b) / abs(b):.3e}"


def rand_case(n1, n2, seed, L=30.0):
    rng = np.random.default_rng(seed)
    p1 = rng.random((n1, 3)) * L
    p2 = rng.random((n2, 3)) * L
    proxy = np.ascontiguousarray(np.concatenate([p1, p2]))
    return proxy, np.arange(n1, dtype=np.int32), np.arange(n1, n1 + n2, dtype=np.int32)


# ---------------------------------------------------------------------------------------------

def test_reciprocal_is_within_two_ulp():
    cb = _cb()
    for lo, hi in ((1.0, 2.0), (1.0, 1e6), (1.0, 1e40), (1e-200, 1e200)):
        assert cb.selftest_rcp(0, 1 << 20, lo, hi) <= 2.0


@pytest.mark.parametrize("case", R.CASES + R.PAIR_CASES)
def test_reference_regression_cases(case):
    """The reference's own coordnum regression inputs (5 frames of deca-alanine) through
    the host-buffer C-ABI step, against the golden traj/forces (the reference comparer's
    tolerance) AND against the oracle (1e-12)."""
    cb = _cb()
    cfg, frames, traj, forces = R.load_case(case)
    natoms = frames[0].shape[0]
    idx1, idx2 = R.stub_slots(cfg, natoms)
    cv = cb.CoordNum(cfg.function_type, idx1, idx2 if cfg.kind != 1 else None,
                     cutoff3=cfg.cutoff3, exp_numer=cfg.exp_numer, exp_denom=cfg.exp_denom,
                     group1_center_only=cfg.group1_center_only,
                     group2_center_only=cfg.group2_center_only, tolerance=cfg.tolerance,
                     pairlist_freq=cfg.pairlist_freq, exponent=cfg.exponent)
    oracle = R.run_oracle_case(cfg, frames)
    results = []
    for step, xyz in enumerate(frames):
        pos = np.ascontiguousarray(xyz, dtype=np.float64)
        v = cv.calc_host(pos, step=step, gradients=True)
        fa = R.harmonic_force(cfg, v)
        pf = np.zeros((natoms, 3))
        cv.apply_force_host(fa, pf)
        results.append((v, fa, pf))
        ov, ofa, opf = oracle[step]
        close_val(v, ov)
        close_val(fa, ofa)
        close_arr(pf, opf)
    assert not R.check_against_golden(results, traj, forces)


@pytest.mark.parametrize("n1,n2", [(1, 1), (1, 33), (5, 7), (31, 32), (33, 65), (257, 1000),
                                   (1030, 2100)])
def test_coordnum_random_sizes(n1, n2):
    cb = _cb()
    proxy, i1, i2 = rand_case(n1, n2, seed=n1 * 1000 + n2, L=12.0 + 0.01 * n2)
    cv = cb.CoordNum("coordNum", i1, i2, cutoff=4.0)
    v = cv.calc_host(proxy)
    p = O.make_params((4.0, 4.0, 4.0))
    ov, og1, og2 = O.coordnum(p, proxy[i1], proxy[i2])
    close_val(v, ov)
    close_arr(cv.gradients(1), og1)
    close_arr(cv.gradients(2), og2)
    # Newton's third law: the gradients of the two groups cancel
    tot = cv.gradients(1).sum(0) + cv.gradients(2).sum(0)
    assert np.max(np.abs(tot)) <= 1e-11 * max(np.max(np.abs(og1)), 1e-300) * (n1 + n2)


@pytest.mark.parametrize("n", [2, 3, 32, 33, 100, 1025, 2500])
def test_selfcoordnum_random_sizes(n):
    cb = _cb()
    proxy, i1, _ = rand_case(n, 0, seed=77 + n, L=10.0 + 0.005 * n)
    cv = cb.CoordNum("selfCoordNum", i1, cutoff=3.5)
    v = cv.calc_host(proxy)
    p = O.make_params((3.5, 3.5, 3.5))
    ov, og = O.selfcoordnum(p, proxy)
    close_val(v, ov)
    close_arr(cv.gradients(1), og)


def test_anisotropic_cutoff3_and_general_exponents():
    cb = _cb()
    proxy, i1, i2 = rand_case(300, 1500, seed=5, L=25.0)
    for cutoff3, en, ed in (((3.0, 6.0, 4.0), 6, 12), ((4.0, 4.0, 4.0), 8, 10),
                            ((2.5, 3.5, 5.0), 4, 8), ((4.0, 4.0, 4.0), 2, 4),
                            ((3.3, 3.3, 3.3), 6, 8), ((5.0, 4.0, 3.0), 12, 6)):
        cv = cb.CoordNum("coordNum", i1, i2, cutoff3=cutoff3, exp_numer=en, exp_denom=ed)
        v = cv.calc_host(proxy)
        p = O.make_params(cutoff3, en, ed)
        ov, og1, og2 = O.coordnum(p, proxy[i1], proxy[i2])
        close_val(v, ov)
        close_arr(cv.gradients(1), og1)
        close_arr(cv.gradients(2), og2)
        cv.close()


def test_value_only_leaves_gradients_untouched():
    cb = _cb()
    proxy, i1, i2 = rand_case(100, 300, seed=9)
    cv = cb.CoordNum("coordNum", i1, i2)
    v = cv.calc_host(proxy, gradients=False)
    p = O.make_params()
    ov, _, _ = O.coordnum(p, proxy[i1], proxy[i2], gradients=False)
    close_val(v, ov)
    assert not cv.gradients(1).any() and not cv.gradients(2).any()


def test_pairs_at_the_taylor_window_and_coincident_thresholds():
    """Pairs placed exactly at / next to l2 = 1 (the reference's Taylor branch,
    |l2 - 1| < 1e-7) must come out as the reference computes them."""
    cb = _cb()
    r0 = 4.0
    offs = [0.0, 1e-9, -1e-9, 4e-8, -4e-8, 6e-8, 1e-6, -1e-6, 1e-4, -1e-4, 4e-4, -4e-4, 1e-3,
            -1e-3]
    p1 = np.zeros((len(offs), 3))
    p2 = np.zeros((len(offs), 3))
    for k, o in enumerate(offs):
        p1[k] = (100.0 * k, 0.0, 0.0)          # far apart rows: only pair (k,k) is near l2 = 1
        d = np.array([1.0, 2.0, 2.0]) / 3.0    # unit vector
        p2[k] = p1[k] + d * r0 * np.sqrt(1.0 + o)
    proxy = np.ascontiguousarray(np.concatenate([p1, p2]))
    i1 = np.arange(len(offs), dtype=np.int32)
    i2 = i1 + len(offs)
    cv = cb.CoordNum("coordNum", i1, i2, cutoff=r0)
    v = cv.calc_host(proxy)
    ov, og1, og2 = O.coordnum(O.make_params((r0,) * 3), p1, p2)
    close_val(v, ov, 1e-13)
    close_arr(cv.gradients(1), og1, 1e-13)
    close_arr(cv.gradients(2), og2, 1e-13)
    launches, exact = cv.last_stats()
    assert exact == len(offs)  # all of those pairs went through the reference-order path


@pytest.mark.parametrize("kind", ["coordNum", "selfCoordNum"])
@pytest.mark.parametrize("aniso", [False, True])
def test_pairlist_indices_bit_exact_and_values(kind, aniso):
    """Rebuild on steps 0, 3, 6; list walk in between (pairListFrequency 3).  The listed
    set has to equal the reference's dense bool array bit for bit on every rebuild."""
    cb = _cb()
    n1, n2 = (700, 1800) if kind == "coordNum" else (1500, 0)
    proxy, i1, i2 = rand_case(n1, n2, seed=21 + n1, L=22.0)
    cutoff3 = (3.0, 5.0, 4.0) if aniso else (4.0, 4.0, 4.0)
    tol = 0.02
    cv = cb.CoordNum(kind, i1, i2 if kind == "coordNum" else None, cutoff3=cutoff3,
                     tolerance=tol, pairlist_freq=3)
    p = O.make_params(cutoff3, 6, 12, tol)
    assert cv.tolerance_l2_max() == p.tolerance_l2_max
    npairs = n1 * n2 if kind == "coordNum" else n1 * (n1 - 1) // 2
    pl = np.ones(npairs, dtype=np.uint8)
    rng = np.random.default_rng(3)
    pos = proxy.copy()
    for step in range(8):
        rebuild = step % 3 == 0
        v = cv.calc_host(pos, step=step)
        if kind == "coordNum":
            ov, og1, og2 = O.coordnum(p, pos[i1], pos[i2], pairlist=pl, rebuild=rebuild)
            close_arr(cv.gradients(2), og2)
        else:
            ov, og1 = O.selfcoordnum(p, pos, pairlist=pl, rebuild=rebuild)
        close_val(v, ov)
        close_arr(cv.gradients(1), og1)
        if rebuild:
            mine = cv.pairlist()
            ref = np.flatnonzero(pl).astype(np.uint64)
            assert mine.shape == ref.shape and np.array_equal(mine, ref)
        pos = pos + rng.normal(0.0, 0.05, pos.shape)


def test_pairlist_before_first_rebuild_is_all_true():
    """step_relative = 1 on the first call: the reference walks an all-true list."""
    cb = _cb()
    proxy, i1, i2 = rand_case(90, 200, seed=4, L=14.0)
    cv = cb.CoordNum("coordNum", i1, i2, tolerance=0.01, pairlist_freq=5)
    p = O.make_params((4.0,) * 3, 6, 12, 0.01)
    pl = np.ones(90 * 200, dtype=np.uint8)
    v = cv.calc_host(proxy, step=1)
    ov, og1, og2 = O.coordnum(p, proxy[i1], proxy[i2], pairlist=pl, rebuild=False)
    close_val(v, ov)
    close_arr(cv.gradients(1), og1)
    assert len(cv.pairlist()) == 90 * 200


def test_pairlist_thresholds_are_decided_like_the_reference():
    """Pairs straddling l2_max by a few ulp: membership must match the oracle exactly."""
    cb = _cb()
    tol, r0 = 0.01, 4.0
    p = O.make_params((r0,) * 3, 6, 12, tol)
    l2max = p.tolerance_l2_max
    n = 64
    p1 = np.zeros((n, 3))
    p2 = np.zeros((n, 3))
    for k in range(n):
        p1[k] = (200.0 * k, 0.0, 0.0)
        scale = 1.0 + (k - n // 2) * 2.0 ** -52
        p2[k] = p1[k] + np.array([0.0, 0.0, 1.0]) * r0 * np.sqrt(l2max) * scale
    proxy = np.ascontiguousarray(np.concatenate([p1, p2]))
    i1 = np.arange(n, dtype=np.int32)
    i2 = i1 + n
    cv = cb.CoordNum("coordNum", i1, i2, cutoff=r0, tolerance=tol, pairlist_freq=1)
    v = cv.calc_host(proxy, step=0)
    pl = np.ones(n * n, dtype=np.uint8)
    ov, og1, og2 = O.coordnum(p, p1, p2, pairlist=pl, rebuild=True)
    assert np.array_equal(cv.pairlist(), np.flatnonzero(pl).astype(np.uint64))
    assert abs(v - ov) <= 1e-12 * max(abs(ov), 1e-6)


def test_nearly_orthogonal_cell_follows_the_reference():
    """Cell vectors with off-diagonal terms below the reference's tolerance (1e-5,
    src/colvars_system.h:120-140) keep the type "orthogonal" there: fractional coordinates are
    rounded with the full vectors, the lattice combination subtracted with the full vectors, and
    the neighbouring images are NOT searched.  Pairs at half a box length included."""
    cb = _cb()
    L = 20.0
    rng = np.random.default_rng(31)
    n1, n2 = 150, 600
    proxy = (rng.random((n1 + n2, 3)) * 2.0 - 0.5) * L
    # pairs exactly half a cell vector apart (the rounding decides the image)
    proxy[n1:n1 + 20] = proxy[:20] + np.array([0.5 * L, 0.0, 0.0])
    proxy[n1 + 20:n1 + 40] = proxy[20:40] + np.array([0.0, 0.55 * L, 0.0])
    proxy = np.ascontiguousarray(proxy)
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32)
    cell = ((1, 1, 1), (L, 4e-6, -3e-6), (2e-6, 1.1 * L, 5e-6), (-6e-6, 1e-6, 0.9 * L))
    p = O.make_params((4.0,) * 3, boundaries=cell)
    assert p.bc.type == O.BC_PBC_ORTHOGONAL
    for kw in (dict(), dict(tolerance=0.01, pairlist_freq=2)):
        cv = cb.CoordNum("coordNum", i1, i2, cutoff=4.0, **kw)
        cv.set_boundaries(*cell)
        pp = O.make_params((4.0,) * 3, 6, 12, kw.get("tolerance", 0.0), boundaries=cell)
        pl = np.ones(n1 * n2, dtype=np.uint8) if kw else None
        for step in range(3 if kw else 1):
            v = cv.calc_host(proxy, step=step)
            ov, og1, og2 = O.coordnum(pp, proxy[i1], proxy[i2], pairlist=pl, rebuild=step % 2 == 0)
            close_val(v, ov)
            close_arr(cv.gradients(1), og1)
            close_arr(cv.gradients(2), og2)
            if kw and step % 2 == 0:
                assert np.array_equal(cv.pairlist(), np.flatnonzero(pl).astype(np.uint64))
        cv.close()


@pytest.mark.parametrize("bc", ["ortho", "triclinic", "mixed"])
@pytest.mark.parametrize("kind", ["coordNum", "selfCoordNum"])
def test_periodic_boundaries(bc, kind):
    cb = _cb()
    L = 18.0
    n1, n2 = (200, 900) if kind == "coordNum" else (700, 0)
    rng = np.random.default_rng(11)
    proxy = np.ascontiguousarray((rng.random((n1 + n2, 3)) * 3.0 - 1.0) * L)  # outside the cell too
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32)
    if bc == "ortho":
        per, A, B, Cc = (1, 1, 1), (L, 0, 0), (0, 1.1 * L, 0), (0, 0, 0.9 * L)
    elif bc == "triclinic":
        per, A, B, Cc = (1, 1, 1), (L, 0, 0), (0.3 * L, L, 0), (0.1 * L, -0.2 * L, L)
    else:
        per, A, B, Cc = (1, 1, 0), (L, 0, 0), (0, L, 0), (0, 0, L)
    cv = cb.CoordNum(kind, i1, i2 if kind == "coordNum" else None, cutoff=4.0)
    cv.set_boundaries(per, A, B, Cc)
    v = cv.calc_host(proxy)
    p = O.make_params((4.0,) * 3, boundaries=(per, A, B, Cc))
    if kind == "coordNum":
        ov, og1, og2 = O.coordnum(p, proxy[i1], proxy[i2])
        close_arr(cv.gradients(2), og2)
    else:
        ov, og1 = O.selfcoordnum(p, proxy)
    close_val(v, ov)
    close_arr(cv.gradients(1), og1)


@pytest.mark.parametrize("g1c,g2c,kind", [(True, False, "coordNum"), (False, True, "coordNum"),
                                          (True, True, "coordNum"), (True, True, "groupCoord")])
def test_center_only_variants_and_groupcoord(g1c, g2c, kind):
    cb = _cb()
    proxy, i1, i2 = rand_case(150, 400, seed=8, L=9.0)
    rng = np.random.default_rng(2)
    m1, m2 = rng.random(150) * 15 + 1, rng.random(400) * 15 + 1
    cv = cb.CoordNum(kind, i1, i2, cutoff3=(3.0, 6.0, 4.0), group1_center_only=g1c,
                     group2_center_only=g2c, mass1=m1, mass2=m2)
    v = cv.calc_host(proxy)
    p = O.make_params((3.0, 6.0, 4.0))
    ov, og1, og2 = O.coordnum(p, proxy[i1], proxy[i2], g1_center_only=g1c, g2_center_only=g2c,
                              mass1=m1, mass2=m2)
    close_val(v, ov)
    close_arr(cv.gradients(1), og1)
    close_arr(cv.gradients(2), og2)


def test_device_resident_step_and_force_scatter():
    """colvarproxy_gpu layout: SoA planes with stride = #proxy atoms, shuffled indices,
    shared proxy atoms untouched, forces accumulated on top of existing ones."""
    import torch
    cb = _cb()
    rng = np.random.default_rng(31)
    P = 5000
    allpos = rng.random((P, 3)) * 30.0
    perm = rng.permutation(P)
    i1, i2 = perm[:400].astype(np.int32), perm[400:2400].astype(np.int32)
    cv = cb.CoordNum("coordNum", i1, i2)
    d_pos = torch.tensor(np.ascontiguousarray(allpos.T), device="cuda")
    f0 = rng.random((3, P))
    d_force = torch.tensor(f0, device="cuda")
    cv.read_data(d_pos)
    cv.calc_value(step=0)
    v = cv.value()
    cv.apply_force(-0.37, d_force)
    torch.cuda.synchronize()
    p = O.make_params()
    ov, og1, og2 = O.coordnum(p, allpos[i1], allpos[i2])
    close_val(v, ov)
    expect = f0.T.copy()
    expect[i1] += -0.37 * og1
    expect[i2] += -0.37 * og2
    close_arr(d_force.cpu().numpy().T, expect)
    com, cog = cv.centers(2)
    close_arr(cog, allpos[i2].mean(0), 1e-13)


def test_bare_array_eval_accumulates_into_callers_gradients():
    import torch
    cb = _cb()
    proxy, i1, i2 = rand_case(333, 777, seed=12, L=15.0)
    cv = cb.CoordNum("coordNum", i1, i2)
    pos1 = torch.tensor(np.ascontiguousarray(proxy[i1].T), device="cuda")
    pos2 = torch.tensor(np.ascontiguousarray(proxy[i2].T), device="cuda")
    g1 = torch.ones((3, 333), dtype=torch.float64, device="cuda")
    g2 = torch.zeros((3, 777), dtype=torch.float64, device="cuda")
    cv.eval(pos1, pos2, cb.EF_GRADIENTS, g1, g2)
    v = cv.value()
    torch.cuda.synchronize()
    ov, og1, og2 = O.coordnum(O.make_params(), proxy[i1], proxy[i2])
    close_val(v, ov)
    close_arr(g1.cpu().numpy().T, og1 + 1.0)
    close_arr(g2.cpu().numpy().T, og2)


def test_input_errors_follow_the_reference():
    cb = _cb()
    with pytest.raises(cb.B200cvError, match="odd exponent"):
        cb.CoordNum("coordNum", [0], [1], exp_numer=5)
    with pytest.raises(cb.B200cvError, match="negative exponent"):
        cb.CoordNum("coordNum", [0], [1], exp_numer=-2)
    with pytest.raises(cb.B200cvError, match="share a common atom"):
        cb.CoordNum("coordNum", [0, 1, 2], [5, 2])
    with pytest.raises(cb.B200cvError, match="pairlistfrequency"):
        cb.CoordNum("coordNum", [0], [1], tolerance=0.1, pairlist_freq=0)
    with pytest.raises(cb.B200cvError, match="cutoff"):
        cb.CoordNum("coordNum", [0], [1], cutoff=3.0, cutoff3=(1, 2, 3))


def test_config2_full_size_against_parallel_oracle():
    """BASELINE config 2 at full size (1 k x 100 k, 1e8 pairs): value and gradients
    against the OpenMP oracle, plus size-independent properties."""
    cb = _cb()
    from colvars_b200 import synth
    proxy, i1, i2, L = synth.coordnum_case(1000, 100000, seed=1234)
    cv = cb.CoordNum("coordNum", i1, i2, cutoff=4.0)
    v = cv.calc_host(proxy)
    g1, g2 = cv.gradients(1), cv.gradients(2)
    ov, og1, og2 = O.coordnum(O.make_params(), proxy[i1], proxy[i2], omp=8)
    close_val(v, ov)
    close_arr(g1, og1)
    close_arr(g2, og2)
    # translation invariance and Newton's third law
    assert np.max(np.abs(g1.sum(0) + g2.sum(0))) <= 1e-9
    v2 = cv.calc_host(proxy + np.array([3.0, -7.0, 11.0]))
    close_val(v2, v, 1e-11)
    # a rigid rotation of everything leaves the isotropic value unchanged
    Rm = synth.random_rotation(5)
    v3 = cv.calc_host(np.ascontiguousarray(proxy @ Rm.T))
    close_val(v3, v, 1e-11)


@pytest.mark.parametrize("name", __import__("tests.refvec", fromlist=["NAMES"]).NAMES)
def test_cuda_path_matches_reference_vectors(name):
    """Outputs of the UNMODIFIED reference (tests/golden/ref_vectors, produced by
    oracle/_ref/ref_driver) on periodic cells, centre-only variants with unequal masses,
    general exponents and multi-frame pair lists -- compared directly with the CUDA path.

    Centre-only variants: the COM is a parallel sum here and a sequential one in the
    reference (a few ulp apart); the reference's quotient form amplifies input ulps by
    ~5e-17/h^2 for pairs with l2 = 1 + h, so those cases are held to 1e-11 instead of 1e-12."""
    from . import refvec
    cb = _cb()
    c = refvec.load(name)
    n1, n2 = c["n1"], c["n2"]
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32)
    masses = c.get("masses")
    self_ = c["kind"] == "selfCoordNum"
    cv = cb.CoordNum(c["kind"], i1, None if self_ else i2, cutoff3=tuple(c["cutoff3"]),
                     exp_numer=c["en"], exp_denom=c["ed"],
                     group1_center_only=c["g1c"] and c["kind"] == "coordNum",
                     group2_center_only=c["g2c"] and c["kind"] == "coordNum",
                     tolerance=c["tol"], pairlist_freq=c["freq"],
                     mass1=None if masses is None else masses[:n1],
                     mass2=None if masses is None or self_ else masses[n1:])
    b = refvec.boundaries(c)
    if b is not None:
        cv.set_boundaries(*b)
    nfit = c["nfit"]
    if c["fitted"]:
        cv.set_fitting(1, center=bool(c["fit_center"]), rotate=bool(c["fit_rotate"]),
                       fitting_group=np.arange(n1 + n2, n1 + n2 + nfit, dtype=np.int32)
                       if nfit else None, ref_pos=c["fit_ref"], enable_fit_gradients=True)
    com_case = c["g1c"] or c["g2c"]
    # centre-only cases feed ulp-level differences of the COM into the reference's
    # ill-conditioned quotient near l2 = 1 (see the docstring): 1e-11
    rtol = 1e-11 if com_case else RTOL
    if c["fitted"]:
        return _check_fitted_case(cb, cv, c)
    for k in range(c["nframes"]):
        pos = np.ascontiguousarray(c["frames"][k])
        v = cv.calc_host(pos, step=k)
        close_val(v, float(c[f"value_{k}"]), rtol)
        close_arr(cv.gradients(1), c[f"grad1_{k}"], rtol)
        if not self_:
            close_arr(cv.gradients(2), c[f"grad2_{k}"], rtol)
        pf = np.zeros_like(pos)
        cv.apply_force_host(refvec.harmonic_fa(v), pf)
        close_arr(pf, c[f"forces_{k}"], rtol)
        if int(c[f"npl_{k}"]) >= 0 and k % c["freq"] == 0:
            assert np.array_equal(cv.pairlist(), c[f"pl_{k}"])


def _check_fitted_case(cb, cv, c):
    """Fitted groups.  The rotation comes from a different (equally convergent) eigen-solver
    than the reference's, so fitted positions agree to ~1e-14 but not bitwise -- and the
    reference's quotient form turns input ulps into ~5e-17/h^2 relative noise for pairs with
    l2 = 1 + h (4e-11 on one of these gradients).  So parity is checked in two exact steps
    plus one loose end-to-end comparison:
      1. fitted positions and quaternion vs the reference            (1e-13)
      2. value / gradients / fit gradients / forces vs the ORACLE fed with the positions the
         device actually produced                                      (1e-12 / 1e-11)
      3. everything vs the reference vectors                          (1e-9)"""
    from . import refvec
    n1, n2, nfit = c["n1"], c["n2"], c["nfit"]
    self_ = c["kind"] == "selfCoordNum"
    p = O.make_params(tuple(c["cutoff3"]), c["en"], c["ed"], c["tol"])
    fit = O.Fit(c["fit_ref"], bool(c["fit_center"]), bool(c["fit_rotate"]))
    for k in range(c["nframes"]):
        pos = np.ascontiguousarray(c["frames"][k])
        v = cv.calc_host(pos, step=k)
        fitted1 = cv.positions(1)
        g1 = cv.gradients(1)
        # 1. geometry against the reference
        close_arr(fitted1, c[f"fitted1_{k}"], 1e-13)
        if bool(c["fit_rotate"]):
            q, qr = cv.rotation(1), c[f"q_{k}"]
            assert min(np.abs(q - qr).max(), np.abs(q + qr).max()) < 1e-13
        # 2. pair arithmetic against the oracle on the device's own fitted positions
        if self_:
            ov, og1 = O.selfcoordnum(p, fitted1)
        else:
            ov, og1, og2 = O.coordnum(p, fitted1, pos[n1:n1 + n2])
            close_arr(cv.gradients(2), og2)
        close_val(v, ov)
        close_arr(g1, og1)
        # fit gradients from the device's gradients, through the oracle's chain rule
        _, unrot, _ = fit.apply(pos[:n1], pos[n1 + n2:] if nfit else None)
        nf = nfit if nfit else n1
        ofg = fit.fit_gradients(g1, unrot, nf)
        fg = cv.fit_gradients(1, nf)
        scale = max(np.abs(ofg).max(), np.abs(g1).max())
        assert np.abs(fg - ofg).max() <= 1e-11 * scale
        # 3. end to end against the reference
        close_val(v, float(c[f"value_{k}"]), 1e-9)
        close_arr(g1, c[f"grad1_{k}"], 1e-9)
        assert np.abs(fg - c[f"fitgrad_{k}"]).max() <= 1e-9 * scale
        pf = np.zeros_like(pos)
        cv.apply_force_host(refvec.harmonic_fa(v), pf)
        close_arr(pf, c[f"forces_{k}"], 1e-9)


def test_fitted_group_device_resident_matches_host_path():
    """Same fitted configuration through read_data / calc_value / calc_fit_gradients /
    apply_force on device-resident SoA proxy buffers."""
    import torch
    from . import refvec
    cb = _cb()
    c = refvec.load("fit_group_center_rotate")
    n1, n2, nfit = c["n1"], c["n2"], c["nfit"]
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32)
    cv = cb.CoordNum("coordNum", i1, i2, cutoff3=tuple(c["cutoff3"]))
    cv.set_fitting(1, center=True, rotate=True,
                   fitting_group=np.arange(n1 + n2, n1 + n2 + nfit, dtype=np.int32),
                   ref_pos=c["fit_ref"])
    pos = np.ascontiguousarray(c["frames"][0])
    d_pos = torch.tensor(np.ascontiguousarray(pos.T), device="cuda")
    d_force = torch.zeros_like(d_pos)
    cv.read_data(d_pos)
    cv.calc_value(step=0)
    cv.calc_fit_gradients()
    v = cv.value()
    cv.apply_force(refvec.harmonic_fa(v), d_force)
    torch.cuda.synchronize()
    close_val(v, float(c["value_0"]), 1e-9)
    close_arr(d_force.cpu().numpy().T, c["forces_0"], 1e-9)


def test_fused_graph_step_equals_the_plain_calls():
    """b200cv_step (one CUDA-graph launch) against read_data + calc_value on the same inputs:
    pair-list cycle (rebuild / walk graphs), changing cell (re-capture), fitted group."""
    import torch
    from . import refvec
    cb = _cb()
    rng = np.random.default_rng(41)
    pos = np.ascontiguousarray(rng.random((900 + 2100, 3)) * 21.0)
    i1 = np.arange(900, dtype=np.int32)
    i2 = np.arange(900, 3000, dtype=np.int32)
    d_pos = torch.tensor(np.ascontiguousarray(pos.T), device="cuda")
    for kw in (dict(), dict(tolerance=0.02, pairlist_freq=3), dict(cutoff3=(3.0, 5.0, 4.0))):
        a = cb.CoordNum("coordNum", i1, i2, **kw)
        b = cb.CoordNum("coordNum", i1, i2, **kw)
        for step in range(7):
            d_pos += 0.01
            a.read_data(d_pos)
            a.calc_value(step=step)
            va = a.value()
            b.step(d_pos, step=step)
            vb = b.value()
            close_val(vb, va, 1e-13)
            close_arr(b.gradients(1), a.gradients(1), 1e-13)
            close_arr(b.gradients(2), a.gradients(2), 1e-13)
            if step == 3:  # cell change -> the graphs are re-captured
                for cv in (a, b):
                    cv.set_boundaries((1, 1, 1), (40.0, 0, 0), (0, 41.0, 0), (0, 0, 39.0))
        if "tolerance" in kw:
            assert np.array_equal(a.pairlist(), b.pairlist())
        a.close()
        b.close()
    # fitted group through the graph
    c = refvec.load("fit_group_center_rotate")
    n1, n2, nfit = c["n1"], c["n2"], c["nfit"]
    cv = cb.CoordNum("coordNum", np.arange(n1, dtype=np.int32),
                     np.arange(n1, n1 + n2, dtype=np.int32), cutoff3=tuple(c["cutoff3"]))
    cv.set_fitting(1, center=True, rotate=True,
                   fitting_group=np.arange(n1 + n2, n1 + n2 + nfit, dtype=np.int32),
                   ref_pos=c["fit_ref"])
    d = torch.tensor(np.ascontiguousarray(c["frames"][0].T), device="cuda")
    d_force = torch.zeros_like(d)
    for _ in range(3):
        cv.step(d, step=0)
        v = cv.value()
    cv.apply_force(refvec.harmonic_fa(v), d_force)
    torch.cuda.synchronize()
    close_val(v, float(c["value_0"]), 1e-9)
    close_arr(d_force.cpu().numpy().T, c["forces_0"], 1e-9)


def test_config5_full_size_properties():
    """BASELINE config 5 at full size (100 k x 1 M, 1e11 pairs): too big for the oracle as a
    whole, so (a) 96 random group1 rows against the oracle (their gradients are sums over all
    1 M group2 atoms), (b) additivity over a split of group1 (value and group2 gradients of the
    halves add up to the whole), (c) Newton's third law, (d) translation invariance."""
    cb = _cb()
    from colvars_b200 import synth
    n1, n2 = 100000, 1000000
    proxy, i1, i2, L = synth.coordnum_case(n1, n2, seed=1234)
    cv = cb.CoordNum("coordNum", i1, i2, cutoff=4.0)
    v = cv.calc_host(proxy)
    g1, g2 = cv.gradients(1), cv.gradients(2)
    # (a) random rows against the oracle
    rows = np.sort(np.random.default_rng(8).choice(n1, 96, replace=False))
    ov, og1, _ = O.coordnum(O.make_params(), proxy[i1][rows], proxy[i2], omp=8)
    close_arr(g1[rows], og1)
    # (c) third law; the sums run over 1e5 / 1e6 terms of size <= ~1
    assert np.max(np.abs(g1.sum(0) + g2.sum(0))) <= 1e-7
    # (b) additivity
    half = n1 // 2 + 137
    va = cb.CoordNum("coordNum", i1[:half], i2, cutoff=4.0)
    vb = cb.CoordNum("coordNum", i1[half:], i2, cutoff=4.0)
    xa = va.calc_host(proxy)
    xb = vb.calc_host(proxy)
    close_val(xa + xb, v, 1e-12)
    close_arr(va.gradients(2) + vb.gradients(2), g2)
    close_arr(np.concatenate([va.gradients(1), vb.gradients(1)]), g1)
    # (d) translation
    v2 = cv.calc_host(proxy + np.array([1.5, -2.5, 0.75]))
    close_val(v2, v, 1e-11)


@pytest.mark.parametrize("kind,center", [("selfCoordNum", False), ("coordNum", False),
                                         ("coordNum", True)])
def test_pairlist_inside_one_captured_graph(kind, center):
    """`smp gpu` route with a pair list: ONE captured graph (b200cv_eval_async +
    b200cv_accumulate_gradients) is replayed for rebuild and list-walk steps alike; the mode
    word set by the host between launches decides on the device which branch works.  Checked
    per step against the synchronous path of a second handle and against the oracle."""
    import torch
    cb = _cb()
    rng = np.random.default_rng(5)
    n1, n2 = (1500, 0) if kind == "selfCoordNum" else (400, 3000)
    L = 22.0
    pos = rng.random((n1 + n2, 3)) * L
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32) if n2 else None
    freq, tol, nsteps = 3, 0.01, 8
    kw = dict(cutoff=4.0, tolerance=tol, pairlist_freq=freq)
    if center:
        kw["group2_center_only"] = True
    cv = cb.CoordNum(kind, i1, i2, **kw)       # captured
    ref = cb.CoordNum(kind, i1, i2, **kw)      # synchronous twin
    p1 = torch.tensor(np.ascontiguousarray(pos[i1].T), device="cuda")
    p2 = torch.tensor(np.ascontiguousarray(pos[i2].T), device="cuda") if n2 else None
    g1 = torch.zeros_like(p1)
    g2 = torch.zeros_like(p2) if n2 else None
    flags = cb.EF_GRADIENTS | cb.EF_USE_PAIRLIST
    side = torch.cuda.Stream()
    cv.pairlist_reserve(p1, p2, headroom=2.0, stream=side)
    side.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        cv.eval_async(p1, p2, flags, stream=side)
        cv.accumulate_gradients(g1, g2, stream=side)
    p = O.make_params((4.0,) * 3, tolerance=tol)
    npairs = n1 * (n1 - 1) // 2 if kind == "selfCoordNum" else (n1 if center else n1 * n2)
    pl = np.ones(npairs, dtype=np.uint8)
    for step in range(nsteps):
        frame = pos + rng.normal(0, 0.03, pos.shape) * step
        p1.copy_(torch.tensor(np.ascontiguousarray(frame[i1].T)))
        if n2:
            p2.copy_(torch.tensor(np.ascontiguousarray(frame[i2].T)))
            g2.zero_()
        g1.zero_()
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()  # the reference's scheduler syncs before calc_value_after_gpu
        v = cv.value()
        # what the host shim does in calc_value_after_gpu: mode of the NEXT launch
        cv.pairlist_set_rebuild((step + 1) % freq == 0)
        rebuild = step % freq == 0
        rf = flags | (cb.EF_REBUILD_PAIRLIST if rebuild else 0)
        rg1 = torch.zeros_like(p1)
        rg2 = torch.zeros_like(p2) if n2 else None
        ref.eval(p1, p2, rf, rg1, rg2)
        rv = ref.value()
        torch.cuda.synchronize()
        assert v == rv or abs(v - rv) <= 1e-13 * abs(rv)
        close_arr(g1.cpu().numpy(), rg1.cpu().numpy(), 1e-13)
        if n2:
            close_arr(g2.cpu().numpy(), rg2.cpu().numpy(), 1e-13)
        if kind == "selfCoordNum":
            ov, og1 = O.selfcoordnum(p, frame[i1], pairlist=pl, rebuild=rebuild)
        else:
            ov, og1, og2 = O.coordnum(p, frame[i1], frame[i2], pairlist=pl, rebuild=rebuild,
                                      g2_center_only=center)
            close_arr(g2.cpu().numpy().T, og2, 1e-11 if center else RTOL)
        close_val(v, ov, 1e-11 if center else RTOL)
        close_arr(g1.cpu().numpy().T, og1, 1e-11 if center else RTOL)
        if not center:
            assert np.array_equal(np.sort(cv.pairlist()), np.sort(ref.pairlist()))


def test_duplicate_atoms_in_a_group_are_rejected():
    cb = _cb()
    with pytest.raises(cb.B200cvError, match="twice in a group"):
        cb.CoordNum("coordNum", np.array([0, 1, 1], dtype=np.int32),
                    np.array([5, 6], dtype=np.int32))


@pytest.mark.parametrize("form", ["entries", "rows"])
def test_captured_pairlist_overflow_is_an_error_not_a_truncation(form, monkeypatch):
    """A captured graph holds the list's address, so the list cannot grow under it: when a
    rebuild finds more pairs than were reserved, the step's value must come back as an error --
    and reserving again on the new positions plus a re-capture must recover.  Both sparse
    forms: packed entries in cell order, neighbour rows."""
    import torch
    cb = _cb()
    if form == "entries":
        monkeypatch.setenv("B200CV_NO_ROWS", "1")
    rng = np.random.default_rng(9)
    n = 6000
    pos = rng.random((n, 3)) * 60.0
    i1 = np.arange(n, dtype=np.int32)
    cv = cb.CoordNum("selfCoordNum", i1, cutoff=4.0, tolerance=0.01, pairlist_freq=2)
    p1 = torch.tensor(np.ascontiguousarray(pos.T), device="cuda")
    g1 = torch.zeros_like(p1)
    side = torch.cuda.Stream()

    def capture():
        cv.pairlist_reserve(p1, None, headroom=1.0, stream=side)
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            cv.eval_async(p1, None, cb.EF_GRADIENTS | cb.EF_USE_PAIRLIST, stream=side)
            cv.accumulate_gradients(g1, None, stream=side)
        return graph

    graph = capture()
    graph.replay()
    torch.cuda.synchronize()
    v0 = cv.value()  # fits: same positions as the reservation
    assert v0 > 0
    # the same atoms at a third of the box size: ~27x as many pairs within range.  The list was
    # allocated with at least 2^20 entries, so check that the new count really exceeds it
    p1.mul_(1.0 / 3.0)
    cv.pairlist_set_rebuild(True)
    g1.zero_()
    graph.replay()
    torch.cuda.synchronize()
    dense = cb.CoordNum("selfCoordNum", i1, cutoff=4.0, tolerance=0.01, pairlist_freq=2)
    gd = torch.zeros_like(p1)
    dense.eval(p1, None, cb.EF_GRADIENTS | cb.EF_USE_PAIRLIST | cb.EF_REBUILD_PAIRLIST, gd, None)
    vd = dense.value()
    assert len(dense.pairlist()) > (1 << 20)
    with pytest.raises(cb.B200cvError, match="pair-list overflow in a captured evaluation"):
        cv.value()
    # recover: size the list on the positions of now, capture again
    graph = capture()
    cv.pairlist_set_rebuild(True)
    g1.zero_()
    graph.replay()
    torch.cuda.synchronize()
    close_val(cv.value(), vd, 1e-13)
    close_arr(g1.cpu().numpy(), gd.cpu().numpy(), 1e-13)
    assert np.array_equal(cv.pairlist(), dense.pairlist())


def test_handles_are_independent_across_host_threads():
    """`smp cvcs` runs calc_value() of different components on different OpenMP threads
    (src/colvarproxy.cpp:299-315): the C ABI has to be re-entrant per handle.  Four threads,
    each with its own handle (ctypes drops the GIL during the calls), against serial results."""
    import threading
    cb = _cb()
    cases = []
    for k, kind in enumerate(["coordNum", "selfCoordNum", "coordNum", "distanceInv"]):
        n1, n2 = (300 + 50 * k, 0 if kind == "selfCoordNum" else 900 + 100 * k)
        proxy, i1, i2 = rand_case(n1, n2, seed=70 + k, L=14.0)
        kw = dict(tolerance=0.01, pairlist_freq=2) if k == 2 else {}
        cv = cb.CoordNum(kind, i1, i2 if n2 else None, **kw)
        serial = []
        for step in range(6):
            serial.append((cv.calc_host(proxy + 0.01 * step, step=step), cv.gradients(1).copy()))
        cv.close()
        cases.append((kind, i1, i2 if n2 else None, kw, proxy, serial))
    errors = []

    def work(case):
        kind, i1, i2, kw, proxy, serial = case
        try:
            cv = cb.CoordNum(kind, i1, i2, **kw)
            for rep in range(3):
                for step in range(6):
                    v = cv.calc_host(proxy + 0.01 * step, step=step)
                    g = cv.gradients(1)
                    sv, sg = serial[step]
                    # atomics reorder the gradient sums between runs: 1e-13, not bitwise
                    assert abs(v - sv) <= 1e-13 * abs(sv), (kind, step, v, sv)
                    assert np.max(np.abs(g - sg)) <= 1e-13 * np.max(np.abs(sg)), (kind, step)
            cv.close()
        except Exception as e:  # noqa: BLE001 - reported in the main thread
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(c,)) for c in cases]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("form", ["rows", "cells"])
@pytest.mark.parametrize("kind,aniso", [("selfCoordNum", False), ("coordNum", False),
                                        ("coordNum", True), ("selfCoordNum", True)])
def test_sparse_pairlist_rebuild_through_the_cell_list(kind, aniso, form, monkeypatch):
    """Once a rebuild has shown the list to be sparse (< 2 % of the pairs), later rebuilds only
    visit pairs in neighbouring cells; everything else is exactly zero in the reference as well
    (l2 > tolerance_l2_max returns before anything is evaluated), so value, gradients and list
    must not change.  Two forms of the sparse list in an open system: neighbour rows
    (csrc/rows.cu, the default) and packed entries in cell order (csrc/cells.cu, what periodic
    cells use)."""
    cb = _cb()
    if form == "cells":
        monkeypatch.setenv("B200CV_NO_ROWS", "1")
    rng = np.random.default_rng(77)
    n1, n2 = (3000, 0) if kind == "selfCoordNum" else (700, 5000)
    L = 70.0  # ~0.01 atoms / A^3: a few listed pairs per atom
    pos = rng.random((n1 + n2, 3)) * L
    pos[:5] += 3.0 * L  # a few atoms far outside the box of the others
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32) if n2 else None
    cutoff3 = (3.0, 5.0, 4.0) if aniso else (4.0, 4.0, 4.0)
    tol, freq = 0.01, 2
    cv = cb.CoordNum(kind, i1, i2, cutoff3=cutoff3, tolerance=tol, pairlist_freq=freq)
    p = O.make_params(cutoff3, tolerance=tol)
    npairs = n1 * (n1 - 1) // 2 if kind == "selfCoordNum" else n1 * n2
    pl = np.ones(npairs, dtype=np.uint8)
    used_cells = 0
    for step in range(7):
        frame = pos + rng.normal(0, 0.15, pos.shape) * step
        v = cv.calc_host(np.ascontiguousarray(frame), step=step)
        rebuild = step % freq == 0
        if kind == "selfCoordNum":
            ov, og1 = O.selfcoordnum(p, frame[i1], pairlist=pl, rebuild=rebuild)
        else:
            ov, og1, og2 = O.coordnum(p, frame[i1], frame[i2], pairlist=pl, rebuild=rebuild)
            close_arr(cv.gradients(2), og2)
        close_val(v, ov)
        close_arr(cv.gradients(1), og1)
        if rebuild:
            assert np.array_equal(np.sort(cv.pairlist()), np.flatnonzero(pl).astype(np.uint64))
            assert pl.sum() * 50 < npairs  # sparse, as intended
            if cv.last_stats()[0] >= 6:  # grid, keys, sort, offsets, gather, pairs, final sum
                used_cells += 1
    assert used_cells == 3  # every rebuild after the first one


@pytest.mark.parametrize("form", ["rows", "cells"])
@pytest.mark.parametrize("kind", ["selfCoordNum", "coordNum"])
def test_sparse_pairlist_of_a_droplet_with_far_outliers(kind, form, monkeypatch):
    """A dense droplet and a few atoms hundreds of Angstrom away: the bounding box is almost
    empty, the grid's cells are sized for its mean density, so the droplet sits in a handful of
    cells with hundreds of atoms each -- the counting sort's rank pass and the row build work on
    long cells, duplicated coordinates included (equal keys: ties go by atom index)."""
    cb = _cb()
    if form == "cells":
        monkeypatch.setenv("B200CV_NO_ROWS", "1")
    rng = np.random.default_rng(2024)
    n1, n2 = (2600, 0) if kind == "selfCoordNum" else (500, 2600)
    n = n1 + n2
    pos = rng.normal(0.0, 14.0, (n, 3))            # the droplet
    far = rng.choice(n, 40, replace=False)
    pos[far] = rng.random((40, 3)) * 3000.0 - 1500.0  # the outliers
    dup = rng.choice(np.setdiff1d(np.arange(n), far), 16, replace=False)
    pos[dup[:8]] = pos[dup[8:]] + np.array([0.0, 2.5, 0.0])  # same x: equal fine keys
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n, dtype=np.int32) if n2 else None
    tol, freq = 0.01, 2
    cv = cb.CoordNum(kind, i1, i2, cutoff=3.0, tolerance=tol, pairlist_freq=freq)
    p = O.make_params((3.0,) * 3, tolerance=tol)
    npairs = n1 * (n1 - 1) // 2 if kind == "selfCoordNum" else n1 * n2
    pl = np.ones(npairs, dtype=np.uint8)
    for step in range(5):
        frame = np.ascontiguousarray(pos + rng.normal(0, 0.1, pos.shape) * step)
        v = cv.calc_host(frame, step=step)
        rebuild = step % freq == 0
        if kind == "selfCoordNum":
            ov, og1 = O.selfcoordnum(p, frame[i1], pairlist=pl, rebuild=rebuild)
        else:
            ov, og1, og2 = O.coordnum(p, frame[i1], frame[i2], pairlist=pl, rebuild=rebuild)
            close_arr(cv.gradients(2), og2)
        close_val(v, ov)
        close_arr(cv.gradients(1), og1)
        if rebuild:
            assert np.array_equal(np.sort(cv.pairlist()), np.flatnonzero(pl).astype(np.uint64))
    assert pl.sum() * 50 < npairs  # sparse: the later rebuilds went through the cell order


def test_cell_list_rebuild_inside_the_fused_step_graph():
    """b200cv_step keeps one captured graph per flag set; a rebuild exists in two forms (tile
    sweep for the first one, cell list once the list is known to be sparse) and both have to
    replay correctly, the cell sort included."""
    import torch
    cb = _cb()
    rng = np.random.default_rng(5)
    n1, n2 = 600, 4000
    pos = np.ascontiguousarray(rng.random((n1 + n2, 3)) * 70.0)
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32)
    kw = dict(cutoff=4.0, tolerance=0.01, pairlist_freq=2)
    a = cb.CoordNum("coordNum", i1, i2, **kw)   # plain calls
    b = cb.CoordNum("coordNum", i1, i2, **kw)   # fused graph
    d_pos = torch.tensor(np.ascontiguousarray(pos.T), device="cuda")
    cells = 0
    for step in range(9):
        d_pos += 0.02 * torch.sin(d_pos + step)
        a.read_data(d_pos)
        a.calc_value(step=step)
        va = a.value()
        b.step(d_pos, step=step)
        vb = b.value()
        close_val(vb, va, 1e-13)
        close_arr(b.gradients(1), a.gradients(1), 1e-13)
        close_arr(b.gradients(2), a.gradients(2), 1e-13)
        if step % 2 == 0:
            assert np.array_equal(np.sort(a.pairlist()), np.sort(b.pairlist()))
            cells += int(a.last_stats()[0] >= 6)
    assert cells == 4
    p = O.make_params(tolerance=0.01)
    frame = d_pos.cpu().numpy().T
    pl = np.zeros(n1 * n2, dtype=np.uint8)
    pl[a.pairlist()] = 1
    ov, og1, og2 = O.coordnum(p, frame[i1], frame[i2], pairlist=pl, rebuild=False)
    close_val(va, ov)
    close_arr(a.gradients(1), og1)


@pytest.mark.parametrize("cell", ["open", "octahedron"])
@pytest.mark.parametrize("kind", ["selfCoordNum", "coordNum"])
def test_captured_pairlist_graph_on_a_sparse_system_uses_the_cell_list(kind, cell):
    """The `smp gpu` form (one captured graph, mode word) with a sparse list: the rebuild branch
    of the graph is the gated cell-list pipeline; replays for rebuild and walk steps against the
    synchronous path and the oracle.  Open system, and a triclinic cell (fractional binning, image
    search in the cell kernel and in the walk, all gated)."""
    import torch
    cb = _cb()
    rng = np.random.default_rng(15)
    n1, n2 = (2500, 0) if kind == "selfCoordNum" else (500, 4000)
    pos = rng.random((n1 + n2, 3)) * 70.0
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32) if n2 else None
    freq, tol = 3, 0.01
    kw = dict(cutoff=4.0, tolerance=tol, pairlist_freq=freq)
    cv = cb.CoordNum(kind, i1, i2, **kw)
    boundaries = None
    if cell != "open":
        boundaries = _scaled_cell(cell, 75.0)
        cv.set_boundaries(*boundaries)
    p1 = torch.tensor(np.ascontiguousarray(pos[i1].T), device="cuda")
    p2 = torch.tensor(np.ascontiguousarray(pos[i2].T), device="cuda") if n2 else None
    g1 = torch.zeros_like(p1)
    g2 = torch.zeros_like(p2) if n2 else None
    flags = cb.EF_GRADIENTS | cb.EF_USE_PAIRLIST
    side = torch.cuda.Stream()
    cv.pairlist_reserve(p1, p2, headroom=2.0, stream=side)
    side.synchronize()
    launches_before = cb.kernel_launches()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        cv.eval_async(p1, p2, flags, stream=side)
        cv.accumulate_gradients(g1, g2, stream=side)
    # fetch, grid, keys, sort (>= 1), offsets, pairs, shuffle, walk, final sum, repacks ...
    assert cb.kernel_launches() - launches_before >= 10
    p = O.make_params((4.0,) * 3, tolerance=tol, boundaries=boundaries)
    npairs = n1 * (n1 - 1) // 2 if kind == "selfCoordNum" else n1 * n2
    pl = np.ones(npairs, dtype=np.uint8)
    for step in range(7):
        frame = pos + rng.normal(0, 0.1, pos.shape) * step
        p1.copy_(torch.tensor(np.ascontiguousarray(frame[i1].T)))
        g1.zero_()
        if n2:
            p2.copy_(torch.tensor(np.ascontiguousarray(frame[i2].T)))
            g2.zero_()
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        v = cv.value()
        cv.pairlist_set_rebuild((step + 1) % freq == 0)
        rebuild = step % freq == 0
        if kind == "selfCoordNum":
            ov, og1 = O.selfcoordnum(p, frame[i1], pairlist=pl, rebuild=rebuild)
        else:
            ov, og1, og2 = O.coordnum(p, frame[i1], frame[i2], pairlist=pl, rebuild=rebuild)
            close_arr(g2.cpu().numpy().T, og2)
        close_val(v, ov)
        close_arr(g1.cpu().numpy().T, og1)
        if rebuild:
            assert np.array_equal(np.sort(cv.pairlist()), np.flatnonzero(pl).astype(np.uint64))


@pytest.mark.parametrize("kind", ["selfCoordNum", "coordNum"])
def test_sparse_pairlist_rebuild_through_the_cell_list_periodic(kind):
    """Orthorhombic periodic cell: atoms are wrapped into the unit cell for binning, neighbour
    cells wrap around, candidates go through the reference-order minimum-image pair function."""
    cb = _cb()
    rng = np.random.default_rng(78)
    n1, n2 = (3000, 0) if kind == "selfCoordNum" else (700, 5000)
    edges = np.array([70.0, 75.0, 65.0])
    cell = ((1, 1, 1), (edges[0], 0, 0), (0, edges[1], 0), (0, 0, edges[2]))
    pos = (rng.random((n1 + n2, 3)) * 3.0 - 1.0) * edges  # images in the neighbouring cells too
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32) if n2 else None
    tol, freq = 0.01, 2
    cv = cb.CoordNum(kind, i1, i2, cutoff3=(4.0, 3.5, 4.5), tolerance=tol, pairlist_freq=freq)
    cv.set_boundaries(*cell)
    p = O.make_params((4.0, 3.5, 4.5), tolerance=tol, boundaries=cell)
    npairs = n1 * (n1 - 1) // 2 if kind == "selfCoordNum" else n1 * n2
    pl = np.ones(npairs, dtype=np.uint8)
    used_cells = 0
    for step in range(5):
        frame = pos + rng.normal(0, 0.15, pos.shape) * step
        v = cv.calc_host(np.ascontiguousarray(frame), step=step)
        rebuild = step % freq == 0
        if kind == "selfCoordNum":
            ov, og1 = O.selfcoordnum(p, frame[i1], pairlist=pl, rebuild=rebuild)
        else:
            ov, og1, og2 = O.coordnum(p, frame[i1], frame[i2], pairlist=pl, rebuild=rebuild)
            close_arr(cv.gradients(2), og2)
        close_val(v, ov)
        close_arr(cv.gradients(1), og1)
        if rebuild:
            assert np.array_equal(np.sort(cv.pairlist()), np.flatnonzero(pl).astype(np.uint64))
            used_cells += int(cv.last_stats()[0] >= 6)
    assert used_cells == 2
    # a cell too small for three cells per axis falls back to the sweep (and stays correct)
    small = ((1, 1, 1), (20.0, 0, 0), (0, 75.0, 0), (0, 0, 65.0))
    cv.set_boundaries(*small)
    p2 = O.make_params((4.0, 3.5, 4.5), tolerance=tol, boundaries=small)
    v = cv.calc_host(np.ascontiguousarray(pos), step=0)
    pl2 = np.ones(npairs, dtype=np.uint8)
    if kind == "selfCoordNum":
        ov, og1 = O.selfcoordnum(p2, pos[i1], pairlist=pl2, rebuild=True)
    else:
        ov, og1, _ = O.coordnum(p2, pos[i1], pos[i2], pairlist=pl2, rebuild=True)
    assert cv.last_stats()[0] < 6
    close_val(v, ov)
    close_arr(cv.gradients(1), og1)


# ---- mixed / triclinic cells on the tile sweep (src/colvars_system.h:161-219) ----------------

_SQ2, _SQ6 = np.sqrt(2.0), np.sqrt(6.0)
GENERAL_CELLS = {
    # GROMACS rhombic dodecahedron and truncated octahedron as triclinic cells: strongly skewed,
    # a large share of the reduced vectors has a shorter neighbouring image
    "dodecahedron": ((1, 1, 1), (1.0, 0, 0), (0, 1.0, 0), (0.5, 0.5, 0.5 * _SQ2)),
    "octahedron": ((1, 1, 1), (1.0, 0, 0), (1 / 3, 2 * _SQ2 / 3, 0), (-1 / 3, _SQ2 / 3, _SQ6 / 3)),
    "skewed": ((1, 1, 1), (1.0, 0, 0), (0.45, 0.9, 0), (-0.4, 0.42, 0.8)),
    # two periodic axes, the second one off-diagonal: type "mixed", where the reference leaves
    # the reciprocal vectors of off-diagonal axes at zero and relies on the image search alone
    "mixed-skewed": ((1, 1, 0), (1.0, 0, 0), (0.4, 1.0, 0), (0, 0, 1.0)),
    "mixed-z": ((0, 0, 1), (1.0, 0, 0), (0, 1.0, 0), (0, 0, 1.0)),
}


def _scaled_cell(name, L):
    per, A, B, Cc = GENERAL_CELLS[name]
    return per, tuple(L * np.array(A)), tuple(L * np.array(B)), tuple(L * np.array(Cc))


@pytest.mark.parametrize("cell", list(GENERAL_CELLS))
@pytest.mark.parametrize("kind,tol", [("coordNum", 0.0), ("selfCoordNum", 0.0),
                                      ("coordNum", 0.01), ("selfCoordNum", 0.001)])
def test_general_cells_through_the_tile_sweep(cell, kind, tol):
    """Values, gradients and pair lists in mixed / triclinic cells (image search of the sweep,
    general_image) against the oracle's reference-order position_distance; atoms up to two
    cells outside the unit cell."""
    cb = _cb()
    L = 21.0
    per, A, B, Cc = _scaled_cell(cell, L)
    n1, n2 = (300, 2100) if kind == "coordNum" else (1100, 0)
    rng = np.random.default_rng(97)
    frac = rng.random((n1 + n2, 3)) * 4.0 - 2.0
    proxy = np.ascontiguousarray(frac @ np.array([A, B, Cc]))
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32)
    cv = cb.CoordNum(kind, i1, i2 if kind == "coordNum" else None, cutoff3=(4.0, 5.0, 3.5),
                     tolerance=tol, pairlist_freq=2)
    cv.set_boundaries(per, A, B, Cc)
    p = O.make_params((4.0, 5.0, 3.5), 6, 12, tol, boundaries=(per, A, B, Cc))
    npairs = n1 * n2 if kind == "coordNum" else n1 * (n1 - 1) // 2
    pl = np.ones(npairs, dtype=np.uint8) if tol > 0 else None
    for step in range(3):
        frame = np.ascontiguousarray(proxy + rng.normal(0, 0.04, proxy.shape) * step)
        v = cv.calc_host(frame, step=step)
        rebuild = tol > 0 and step % 2 == 0
        if kind == "coordNum":
            ov, og1, og2 = O.coordnum(p, frame[i1], frame[i2], pairlist=pl, rebuild=rebuild)
            close_arr(cv.gradients(2), og2)
        else:
            ov, og1 = O.selfcoordnum(p, frame, pairlist=pl, rebuild=rebuild)
        close_val(v, ov)
        close_arr(cv.gradients(1), og1)
        if tol > 0:
            assert np.array_equal(cv.pairlist(), np.flatnonzero(pl).astype(np.uint64))


@pytest.mark.parametrize("cell", ["dodecahedron", "skewed", "mixed-skewed"])
def test_general_cells_atoms_far_outside_the_unit_cell(cell):
    """Unwrapped coordinates, tens of cells away from each other: the integers of the fractional
    rounding are large, the single-precision decisions of general_image lose bits there and
    must hand over (exact-pair queue) instead of deciding wrongly."""
    cb = _cb()
    L = 18.0
    per, A, B, Cc = _scaled_cell(cell, L)
    n1, n2 = 256, 2048
    rng = np.random.default_rng(1234)
    frac = rng.random((n1 + n2, 3)) * 80.0 - 40.0
    proxy = np.ascontiguousarray(frac @ np.array([A, B, Cc]))
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32)
    cv = cb.CoordNum("coordNum", i1, i2, cutoff=5.0)
    cv.set_boundaries(per, A, B, Cc)
    p = O.make_params((5.0,) * 3, boundaries=(per, A, B, Cc))
    v = cv.calc_host(proxy)
    ov, og1, og2 = O.coordnum(p, proxy[i1], proxy[i2])
    close_val(v, ov)
    close_arr(cv.gradients(1), og1)
    close_arr(cv.gradients(2), og2)
    # ... and it is a hand-over of a small share of the pairs, not of all of them (2.4 % of the
    # pairs are in the queue anyway: |l2 - 1| < 2^-4)
    assert cv.last_stats()[1] < 0.04 * n1 * n2, cv.last_stats()


def test_general_cell_image_ties_follow_the_reference():
    """Pairs whose reduced vector lies on (or an ulp or two beside) a plane where two images are
    equally long: the reference keeps the zero shift, then the first candidate in loop order.
    A wrong image flips the sign of the pair's gradient, so the gradients tell."""
    cb = _cb()
    L = 16.0
    per, A, B, Cc = _scaled_cell("skewed", L)
    lat = np.array([A, B, Cc])
    dirs = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, -1, 0), (1, 0, 1), (1, 0, -1),
            (0, 1, 1), (0, 1, -1), (1, 1, 1), (1, 1, -1), (1, -1, 1), (1, -1, -1)]
    rng = np.random.default_rng(5)
    p1, p2 = [], []
    for d in dirs:
        s = np.array(d, dtype=float) @ lat
        for sign in (1.0, -1.0):
            for ulps in (-2, -1, 0, 1, 2):
                # half a lattice vector plus a component inside the bisector plane
                t = rng.normal(size=3)
                t -= s * (t @ s) / (s @ s)
                t *= 0.8 / np.linalg.norm(t)
                base = rng.random(3) * L
                vec = sign * 0.5 * s * (1.0 + ulps * 2.0 ** -52) + t
                p1.append(base)
                p2.append(base + vec)
    p1, p2 = np.array(p1), np.array(p2)
    n = len(p1)
    proxy = np.ascontiguousarray(np.concatenate([p1, p2]))
    i1 = np.arange(n, dtype=np.int32)
    i2 = i1 + n
    # a cut-off of the order of the cell, so that the half-lattice pairs weigh in
    cv = cb.CoordNum("coordNum", i1, i2, cutoff=9.0)
    cv.set_boundaries(per, A, B, Cc)
    v = cv.calc_host(proxy)
    p = O.make_params((9.0,) * 3, boundaries=(per, A, B, Cc))
    ov, og1, og2 = O.coordnum(p, p1, p2)
    close_val(v, ov)
    close_arr(cv.gradients(1), og1)
    close_arr(cv.gradients(2), og2)


def test_general_cell_sweep_equals_the_dense_reference_order_kernel(monkeypatch):
    """B200CV_DENSE=1 keeps the one-pair-per-thread kernel that runs the reference's loop for
    every pair: same value and gradients as the sweep's image search, on a full tile grid."""
    cb = _cb()
    L = 40.0
    per, A, B, Cc = _scaled_cell("octahedron", L)
    n1, n2 = 1500, 5000
    rng = np.random.default_rng(3)
    proxy = np.ascontiguousarray((rng.random((n1 + n2, 3)) * 3.0 - 1.0) @ np.array([A, B, Cc]))
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32)
    out = []
    for dense in (False, True):
        if dense:
            monkeypatch.setenv("B200CV_DENSE", "1")
        else:
            monkeypatch.delenv("B200CV_DENSE", raising=False)
        cv = cb.CoordNum("coordNum", i1, i2, cutoff=4.0)
        cv.set_boundaries(per, A, B, Cc)
        v = cv.calc_host(proxy)
        out.append((v, cv.gradients(1).copy(), cv.gradients(2).copy(), cv.last_stats()))
        cv.close()
    (vs, g1s, g2s, st_s), (vd, g1d, g2d, st_d) = out
    close_val(vs, vd)
    close_arr(g1s, g1d)
    close_arr(g2s, g2d)


@pytest.mark.parametrize("cell", ["dodecahedron", "octahedron", "skewed"])
@pytest.mark.parametrize("kind", ["selfCoordNum", "coordNum"])
def test_sparse_pairlist_rebuild_through_the_cell_list_triclinic(kind, cell):
    """Triclinic cells: atoms are binned by their wrapped fractional coordinates, a fractional
    cell is at least one (largest) cut-off wide perpendicular to its faces, every candidate goes
    through the reference's own position_distance (27-image search); list walks use the image
    search of the sweep.  Values, gradients and bit-exact lists over rebuild and list steps."""
    cb = _cb()
    rng = np.random.default_rng(79)
    n1, n2 = (3000, 0) if kind == "selfCoordNum" else (700, 5000)
    per, A, B, Cc = _scaled_cell(cell, 80.0)
    lat = np.array([A, B, Cc])
    pos = (rng.random((n1 + n2, 3)) * 3.0 - 1.0) @ lat  # images in the neighbouring cells too
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32) if n2 else None
    tol, freq = 0.01, 2
    cv = cb.CoordNum(kind, i1, i2, cutoff3=(4.0, 3.5, 4.5), tolerance=tol, pairlist_freq=freq)
    cv.set_boundaries(per, A, B, Cc)
    p = O.make_params((4.0, 3.5, 4.5), tolerance=tol, boundaries=(per, A, B, Cc))
    npairs = n1 * (n1 - 1) // 2 if kind == "selfCoordNum" else n1 * n2
    pl = np.ones(npairs, dtype=np.uint8)
    used_cells = 0
    for step in range(5):
        frame = pos + rng.normal(0, 0.15, pos.shape) * step
        v = cv.calc_host(np.ascontiguousarray(frame), step=step)
        rebuild = step % freq == 0
        if kind == "selfCoordNum":
            ov, og1 = O.selfcoordnum(p, frame[i1], pairlist=pl, rebuild=rebuild)
        else:
            ov, og1, og2 = O.coordnum(p, frame[i1], frame[i2], pairlist=pl, rebuild=rebuild)
            close_arr(cv.gradients(2), og2)
        close_val(v, ov)
        close_arr(cv.gradients(1), og1)
        if rebuild:
            assert np.array_equal(np.sort(cv.pairlist()), np.flatnonzero(pl).astype(np.uint64))
            used_cells += int(cv.last_stats()[0] >= 6)
    assert used_cells == 2
