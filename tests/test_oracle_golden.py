"""Pin the CPU oracle (oracle/coordnum_oracle.c) against the reference's own golden
files for the coordination-number path: the 9 coordnum / groupcoord
regression cases, hbond (the single-pair user of the same pair function) and distanceinv, 5 frames each:
value + applied force (%.14e) and 104x3 per-atom forces (%.12e).

The bar here is stricter than the reference's comparer (spiff, rel 1e-6): the oracle has
to print the SAME TEXT as the reference did, i.e. be bit-compatible at print precision.
"""
import pytest

from . import regression as R


@pytest.mark.parametrize("case", R.CASES + R.PAIR_CASES)
def test_oracle_reproduces_reference_golden_text(case):
    cfg, frames, traj, forces = R.load_case(case)
    assert len(frames) == 5 and frames[0].shape == (104, 3)
    results = R.run_oracle_case(cfg, frames)
    mismatches = []
    for k, (v, fa, pf) in enumerate(results):
        if f"{v:.14e}" != f"{traj[k, 1]:.14e}":
            mismatches.append(("value", k, v, traj[k, 1]))
        if f"{fa:.14e}" != f"{traj[k, 2]:.14e}":
            mismatches.append(("fa", k, fa, traj[k, 2]))
        g = forces[k]
        assert g.shape == pf.shape
        for a in range(g.shape[0]):
            for d in range(3):
                if f"{pf[a, d]:.12e}" != f"{g[a, d]:.12e}":
                    mismatches.append(("force", k, a, d, pf[a, d], g[a, d]))
    assert not mismatches, mismatches[:5]


def test_tolerance_l2_max_matches_reference_log():
    """Newton solve of coordnum::compute_tolerance_l2_max (src/colvarcomp_coordnums.cpp:162-184):
    F_pairlist(l2_max) ~ 0 and l2_max > 1 for the regression tolerance 0.01."""
    from . import oracle_lib as O
    l2max = O.lib().cvo_compute_tolerance_l2_max(6, 12, 0.01)
    f, _ = O.switching_function(l2max, 6, 12, 0.01, O.EF_USE_PAIRLIST | O.EF_GRADIENTS)
    assert 1.0 < l2max < 10.0
    assert abs(f) < 1e-6
