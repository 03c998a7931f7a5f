"""The geometric invariant behind the cell-list rebuild of the pair list (colvars_b200/csrc/
cells.cu), restated in numpy with the same formulas (grid_kernel, cell_coord): cells are at
least one cut-off radius wide per axis, so two atoms whose cell indices differ by more than one
along any axis are farther apart than the cut-off along that axis, i.e. outside
tolerance_l2_max.  Also checks the clamping of query atoms outside the grid."""
import numpy as np
import pytest

CELL_MAX_DIM = 128


def grid_of(pos2, cut):
    lo, hi = pos2.min(0), pos2.max(0)
    ext = np.maximum(hi - lo, 0.0)
    cell = np.maximum(cut * (1.0 + 1.0e-9), ext / CELL_MAX_DIM * (1.0 + 1.0e-9))
    cell[~(cell > 0)] = 1.0
    dim = np.clip(np.floor(ext / cell).astype(int) + 1, 1, CELL_MAX_DIM)
    return lo, 1.0 / cell, dim


def cell_coord(p, origin, inv_cell, dim):
    c = np.floor((p - origin) * inv_cell)
    return np.minimum(np.maximum(c, -2.0), dim + 1.0).astype(int)


@pytest.mark.parametrize("seed", range(6))
def test_pairs_within_the_cutoff_are_in_neighbouring_cells(seed):
    rng = np.random.default_rng(seed)
    n1, n2 = 300, 2000
    L = float(rng.choice([5.0, 40.0, 900.0]))
    pos2 = rng.random((n2, 3)) * L * np.array([1.0, 0.3, 2.0])
    pos1 = (rng.random((n1, 3)) * 1.6 - 0.3) * L  # some query atoms outside the grid
    cut = rng.uniform(2.0, 9.0, 3)
    # a few query atoms right next to grid atoms (also across cell faces and at the grid's edge)
    pos1[:100] = pos2[:100] + rng.uniform(-1.0, 1.0, (100, 3)) * cut
    origin, inv_cell, dim = grid_of(pos2, cut)
    c2 = np.clip(cell_coord(pos2, origin, inv_cell, dim), 0, dim - 1)
    c1 = cell_coord(pos1, origin, inv_cell, dim)
    d = np.abs(pos1[:, None, :] - pos2[None, :, :])              # (n1, n2, 3)
    within = np.all(d <= cut[None, None, :], axis=2)             # necessary for l2 <= l2_max
    # the kernel visits cells [max(c-1, 0), min(c+1, dim-1)] per axis
    lo = np.maximum(c1 - 1, 0)[:, None, :]
    hi = np.minimum(c1 + 1, dim - 1)[:, None, :]
    visited = np.all((c2[None, :, :] >= lo) & (c2[None, :, :] <= hi), axis=2)
    assert np.all(visited[within]), "a pair inside the cut-off box would be skipped"
    assert within.sum() > 0
    assert np.prod(dim) <= CELL_MAX_DIM ** 3
