"""The oracle's restatement of colvar::distance_pairs (src/colvarcomp_distances.cpp:486-548)
against the reference's own regression golden distancepairs_linear-distancepairs: the 16
distances of group1 x group2 per frame (trajectory file) and the atom forces under the linear
bias of that input (force on every component = -forceConstant / width = -0.002)."""
import re

import numpy as np

from . import oracle_lib as O
from . import regression as R

CASE = R.GOLDEN / "distancepairs_linear-distancepairs"


def golden_values():
    rows = []
    for line in (CASE / "AutoDiff" / "test_out.colvars.traj").read_text().splitlines():
        if line.startswith("#") or not line.strip():
            continue
        rows.append([float(v) for v in re.findall(r"[-+]?\d\.\d+e[-+]\d+", line)])
    return np.array(rows)


def groups():
    """Proxy slots of the two groups: the reference's stub proxy hands out slots in order of first
    request and fills them with the rows of the xyz file in that order (tests/regression.py,
    stub_slots), so group1 reads rows 0-3 and group2 rows 4-7."""
    g = R.read_index_file(R.GOLDEN / "index.ndx")
    n1, n2 = len(g["group1"]), len(g["group2"])
    return np.arange(n1), np.arange(n1, n1 + n2)


def test_distance_pairs_values_and_forces_match_the_reference_golden():
    frames = R.read_xyz_frames(R.GOLDEN / "trajectory.xyz")
    i1, i2 = groups()
    gold = golden_values()
    # value vector, then the applied-force vector (outputAppliedForce on)
    assert gold.shape == (5, 2 * len(i1) * len(i2))
    gold = gold[:, :len(i1) * len(i2)]
    for k in range(5):
        xyz = frames[k]
        force = np.full(len(i1) * len(i2), -0.001 / 0.5)
        d, f1, f2 = O.distance_pairs(xyz[i1], xyz[i2], force=force)
        # the trajectory file prints 15 significant digits
        assert np.max(np.abs(d - gold[k]) / gold[k]) < 5e-15
        gf = np.loadtxt(CASE / "AutoDiff" / f"test_out_forces_{k}.dat")
        mine = np.zeros_like(xyz)
        mine[i1] += f1
        mine[i2] += f2
        assert np.max(np.abs(mine - gf)) <= 1e-12 * np.max(np.abs(gf)) + 1e-15
