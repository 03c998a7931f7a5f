"""SURVEY.md section 8f, second halves of ranks 2 and 3, at the C ABI:
 * the batch of single pairs behind colvar::alpha_angles' hydrogen-bond terms
   (b200cv_eval_pairs_host), against the oracle's pair function pair by pair;
 * colvar::distance_pairs (b200cv_distance_pairs_host / _force_host), against the oracle's
   restatement (pinned to the reference's golden in tests/test_oracle_distancepairs.py),
   bit for bit on the distances -- same operation order -- and 1e-12 on the forces (atomics).
The same components through the unmodified host: tests/test_gpu_host_shim.py."""
import numpy as np
import pytest

from . import oracle_lib as O

pytestmark = pytest.mark.gpu


def _cb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import colvars_b200 as cb
    cb.load()
    return cb


@pytest.mark.parametrize("n,en,ed,r0", [(1, 6, 8, 3.3), (6, 6, 8, 3.3), (77, 4, 10, 3.5),
                                        (1000, 6, 12, 4.0)])
def test_pair_batch_equals_the_single_pairs(n, en, ed, r0):
    cb = _cb()
    rng = np.random.default_rng(n)
    p1 = rng.random((n, 3)) * 30.0
    p2 = p1 + rng.normal(0.0, 0.6 * r0, (n, 3))
    p2[: n // 3] = p1[: n // 3] + r0 * (1.0 + 1e-9 * rng.normal(size=(n // 3, 1))) * np.array([0.6, 0.0, 0.8])
    i1 = np.arange(n, dtype=np.int32)
    i2 = np.arange(n, 2 * n, dtype=np.int32)
    cv = cb.CoordNum("coordNum", i1, i2, cutoff=r0, exp_numer=en, exp_denom=ed)
    g1, g2 = np.zeros(3 * n), np.zeros(3 * n)
    v0 = cv.eval_pairs_host(O.soa(p1), O.soa(p2), flags=cb.EF_NULL)
    v = cv.eval_pairs_host(O.soa(p1), O.soa(p2), flags=cb.EF_GRADIENTS, grad1=g1, grad2=g2)
    p = O.make_params((r0,) * 3, en, ed)
    ov, og1, og2 = 0.0, np.zeros((n, 3)), np.zeros((n, 3))
    for k in range(n):
        f, a, b = O.coordnum(p, p1[k:k + 1], p2[k:k + 1])
        ov += f
        og1[k], og2[k] = a[0], b[0]
    assert abs(v - ov) <= 1e-13 * abs(ov) and abs(v0 - ov) <= 1e-13 * abs(ov)
    # one pair per atom: no summation at all, the gradients are the reference's bit for bit
    assert np.array_equal(O.unsoa(g1, n), og1) and np.array_equal(O.unsoa(g2, n), og2)


@pytest.mark.parametrize("n1,n2", [(1, 1), (4, 4), (33, 70), (300, 1000)])
@pytest.mark.parametrize("cell", [None, "ortho", "triclinic"])
def test_distance_pairs_against_the_oracle(n1, n2, cell):
    cb = _cb()
    rng = np.random.default_rng(100 * n1 + n2)
    L = 15.0
    p1 = (rng.random((n1, 3)) * 2.0 - 0.5) * L
    p2 = (rng.random((n2, 3)) * 2.0 - 0.5) * L
    if n1 > 1 and n2 > 1:
        p2[1] = p1[1]  # coincident atoms: the reference's unit() of a null vector is (1, 0, 0)
    bc = {None: None, "ortho": ((1, 1, 1), (L, 0, 0), (0, 1.2 * L, 0), (0, 0, 0.9 * L)),
          "triclinic": ((1, 1, 1), (L, 0, 0), (0.3 * L, L, 0), (0.1 * L, -0.2 * L, L))}[cell]
    cv = cb.CoordNum("distancePairs", np.arange(n1, dtype=np.int32),
                     np.arange(n1, n1 + n2, dtype=np.int32))
    if bc is not None:
        cv.set_boundaries(*bc)
    d = cv.distance_pairs_host(O.soa(p1), O.soa(p2))
    force = rng.normal(size=n1 * n2)
    f1, f2 = cv.distance_pairs_force_host(force)
    od, of1, of2 = O.distance_pairs(p1, p2, force=force, boundaries=bc)
    assert np.array_equal(d, od)
    scale = max(np.abs(of1).max(), np.abs(of2).max())
    assert np.abs(f1 - of1).max() <= 1e-12 * scale and np.abs(f2 - of2).max() <= 1e-12 * scale
