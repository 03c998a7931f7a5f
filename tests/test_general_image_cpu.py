"""CPU restatement (numpy) of the image search the tile sweep runs in mixed / triclinic cells
(colvars_b200/csrc/pair_math.cuh, general_image), checked against the oracle's
position_distance (src/colvars_system.h:161-219, pinned to the unmodified reference by
tests/test_oracle_ref_vectors.py):

* wherever the search takes a decision, the image vector is BIT-identical to the reference's;
* the decisions it declines (another candidate, or the zero shift, within the margin) are a
  vanishing share of the pairs -- those go to the exact-pair queue on the GPU.

The arithmetic here is IEEE double without contraction; the kernel contracts the three dot
products of step 3 into FMAs, which moves the candidates' scores by an ulp -- nine orders of
magnitude below the margin.

Since round 2 the kernel takes the DECISIONS (the integers of the fractional rounding, the image)
in single precision with wider margins and does the arithmetic with the decided values in double
precision in the reference's order (general_image: general_wrap-like steps 1-2, step 3);
general_image_f32 below restates that with numpy float32 (no FMA: errors a little larger than
the kernel's) and is held to the same two statements, also for vectors tens of cells long."""
import ctypes as C

import numpy as np
import pytest

from . import oracle_lib as O

SQ2, SQ6 = np.sqrt(2.0), np.sqrt(6.0)
CELLS = {
    "dodecahedron": ((1, 1, 1), (1.0, 0, 0), (0, 1.0, 0), (0.5, 0.5, 0.5 * SQ2)),
    "octahedron": ((1, 1, 1), (1.0, 0, 0), (1 / 3, 2 * SQ2 / 3, 0), (-1 / 3, SQ2 / 3, SQ6 / 3)),
    "skewed": ((1, 1, 1), (1.0, 0, 0), (0.45, 0.9, 0), (-0.4, 0.42, 0.8)),
    "mixed-skewed": ((1, 1, 0), (1.0, 0, 0), (0.4, 1.0, 0), (0, 0, 1.0)),
    "mixed-z": ((0, 0, 1), (1.0, 0, 0), (0, 1.0, 0), (0, 0, 1.0)),
    "mixed-xz": ((1, 0, 1), (22.0 / 21, 0, 0), (0, 22.0 / 21, 0), (0, 0, 22.0 / 21)),
}
DIRS = np.array([(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, -1, 0), (1, 0, 1), (1, 0, -1),
                 (0, 1, 1), (0, 1, -1), (1, 1, 1), (1, 1, -1), (1, -1, 1), (1, -1, -1)], float)


def tables(bc):
    """fill_general_cell_tables (colvars_b200/csrc/api.cu)"""
    per = list(bc.periodic)
    cell = np.array([list(v) for v in bc.unit_cell])
    shift = np.zeros((13, 3))
    n2 = np.full(13, 1e300)
    max_n = 0.0
    for k, (ix, iy, iz) in enumerate(DIRS):
        allowed = all(i == 0 or p for i, p in zip((ix, iy, iz), per))
        s = (ix * cell[0] + iy * cell[1]) + iz * cell[2]
        if allowed and s @ s > 0:
            shift[k] = s
            n2[k] = s @ s
            max_n = max(max_n, s @ s)
    return cell, shift, n2, np.ldexp(max_n, -26)


def general_image(bc, d):
    """d: (n, 3) raw differences p2 - p1 -> (image vectors, ambiguous mask)"""
    cell, shift, n2, margin = tables(bc)
    recip = np.array([list(v) for v in bc.reciprocal_cell])
    dx, dy, dz = d[:, 0].copy(), d[:, 1].copy(), d[:, 2].copy()
    sh = [np.floor(((recip[a, 0] * dx + recip[a, 1] * dy) + recip[a, 2] * dz) + 0.5)
          for a in range(3)]
    dx = dx - ((sh[0] * cell[0, 0] + sh[1] * cell[1, 0]) + sh[2] * cell[2, 0])
    dy = dy - ((sh[0] * cell[0, 1] + sh[1] * cell[1, 1]) + sh[2] * cell[2, 1])
    dz = dz - ((sh[0] * cell[0, 2] + sh[1] * cell[1, 2]) + sh[2] * cell[2, 2])
    p = [2.0 * (dx * cell[a, 0] + dy * cell[a, 1] + dz * cell[a, 2]) for a in range(3)]
    q = np.stack([DIRS[k, 0] * p[0] + DIRS[k, 1] * p[1] + DIRS[k, 2] * p[2] for k in range(13)])
    m = n2[:, None] - np.abs(q)
    best = m.min(axis=0)
    idx = m.argmin(axis=0)
    cnt = (m < best + margin).sum(axis=0)
    shifted = best < margin
    ambiguous = shifted & ((cnt >= 2) | (best > -margin))
    s = np.where(shifted[:, None], shift[idx], 0.0)
    t = dx * s[:, 0] + dy * s[:, 1] + dz * s[:, 2]
    sgn = np.where(t > 0, -1.0, 1.0)
    return np.stack([dx + sgn * s[:, 0], dy + sgn * s[:, 1], dz + sgn * s[:, 2]], axis=1), ambiguous


def general_image_f32(bc, d):
    """general_image of pair_math.cuh, round 2: decisions in float32, arithmetic in float64"""
    f32 = np.float32
    cell, shift, n2, _ = tables(bc)
    recip = np.array([list(v) for v in bc.reciprocal_cell])
    max_n = max([v for v in n2 if v < 1e299], default=0.0)
    marginf = f32(np.ldexp(max_n, -15))
    n2f = np.minimum(n2, 3.0e38).astype(f32)
    recipf, cell2f = recip.astype(f32), (2.0 * cell).astype(f32)
    dx, dy, dz = d[:, 0].copy(), d[:, 1].copy(), d[:, 2].copy()
    fx, fy, fz = dx.astype(f32), dy.astype(f32), dz.astype(f32)
    amb = np.zeros(len(d), bool)
    sh = []
    with np.errstate(invalid="ignore", over="ignore"):
        for a in range(3):
            t0, t1, t2 = recipf[a, 0] * fx, recipf[a, 1] * fy, recipf[a, 2] * fz
            u = (t0 + t1) + t2
            mag = (np.abs(t0) + np.abs(t1)) + np.abs(t2)
            r = np.rint(u)
            dev = np.abs(u - r)
            amb |= ~((mag * f32(2.0 ** -19) + dev) < f32(0.5))
            sh.append(r.astype(np.float64))
        dx = dx - ((sh[0] * cell[0, 0] + sh[1] * cell[1, 0]) + sh[2] * cell[2, 0])
        dy = dy - ((sh[0] * cell[0, 1] + sh[1] * cell[1, 1]) + sh[2] * cell[2, 1])
        dz = dz - ((sh[0] * cell[0, 2] + sh[1] * cell[1, 2]) + sh[2] * cell[2, 2])
        gx, gy, gz = dx.astype(f32), dy.astype(f32), dz.astype(f32)
        p = [(gx * cell2f[a, 0] + gy * cell2f[a, 1]) + gz * cell2f[a, 2] for a in range(3)]
        pab, pamb = p[0] + p[1], p[0] - p[1]
        q = np.stack([p[0], p[1], p[2], pab, pamb, p[0] + p[2], p[0] - p[2], p[1] + p[2],
                      p[1] - p[2], pab + p[2], pab - p[2], pamb + p[2], pamb - p[2]])
        v = n2f[:, None] - np.abs(q)
        # the index rides in the four lowest mantissa bits
        m = ((v.view(np.int32) & ~np.int32(15)) | np.arange(13, dtype=np.int32)[:, None]).view(f32)
        best = m.min(axis=0)
        cnt = (m < best + marginf).sum(axis=0)
        shifted = best < marginf
        amb |= (shifted & ((cnt >= 2) | (best > -marginf))) | np.isnan(best)
        idx = np.where(shifted, best.view(np.int32) & 15, 13)
    tab = np.concatenate([shift, np.zeros((1, 3))])
    s = tab[idx]
    t = dx * s[:, 0] + dy * s[:, 1] + dz * s[:, 2]
    sgn = np.where(t > 0, -1.0, 1.0)
    return np.stack([dx + sgn * s[:, 0], dy + sgn * s[:, 1], dz + sgn * s[:, 2]], axis=1), amb


@pytest.mark.parametrize("span", [2.0, 40.0])
@pytest.mark.parametrize("cell", list(CELLS))
def test_single_precision_decisions_are_the_references_wherever_taken(cell, span):
    """span: the atoms lie within +- span cells of the origin (40: unwrapped trajectories; the
    margins grow with the size of the terms and hand more pairs over)"""
    per, A, B, Cc = CELLS[cell]
    L = 21.0
    A, B, Cc = (L * np.array(v) for v in (A, B, Cc))
    p = O.make_params(boundaries=(per, A, B, Cc))
    lib = O.lib()
    rng = np.random.default_rng(4242)
    n = 60000
    lat = np.array([A, B, Cc])
    p1 = (rng.random((n, 3)) * 2 * span - span) @ lat
    p2 = (rng.random((n, 3)) * 2 * span - span) @ lat
    mine, ambiguous = general_image_f32(p.bc, p2 - p1)
    ref = np.zeros((n, 3))
    out = np.zeros(3)
    dp = C.POINTER(C.c_double)
    for k in range(n):
        a, b = np.ascontiguousarray(p1[k]), np.ascontiguousarray(p2[k])
        lib.cvo_position_distance(C.byref(p.bc), a.ctypes.data_as(dp), b.ctypes.data_as(dp),
                                  out.ctypes.data_as(dp))
        ref[k] = out
    decided = ~ambiguous
    assert ambiguous.mean() < (3e-3 if span < 10 else 2e-2), ambiguous.mean()
    assert np.array_equal(mine[decided], ref[decided])


@pytest.mark.parametrize("cell", list(CELLS))
def test_image_search_is_the_references_wherever_it_decides(cell):
    per, A, B, Cc = CELLS[cell]
    L = 21.0
    A, B, Cc = (L * np.array(v) for v in (A, B, Cc))
    p = O.make_params(boundaries=(per, A, B, Cc))
    lib = O.lib()
    rng = np.random.default_rng(42)
    n = 40000
    lat = np.array([A, B, Cc])
    p1 = (rng.random((n, 3)) * 4 - 2) @ lat
    p2 = (rng.random((n, 3)) * 4 - 2) @ lat
    mine, ambiguous = general_image(p.bc, p2 - p1)
    ref = np.zeros((n, 3))
    out = np.zeros(3)
    dp = C.POINTER(C.c_double)
    for k in range(n):
        a, b = np.ascontiguousarray(p1[k]), np.ascontiguousarray(p2[k])
        lib.cvo_position_distance(C.byref(p.bc), a.ctypes.data_as(dp), b.ctypes.data_as(dp),
                                  out.ctypes.data_as(dp))
        ref[k] = out
    decided = ~ambiguous
    assert ambiguous.mean() < 1e-4
    assert np.array_equal(mine[decided], ref[decided])
    # and a fair share of the vectors does take a neighbouring image in the skewed cells
    if cell in ("dodecahedron", "octahedron", "skewed"):
        assert 0.05 < np.mean(np.any(mine != (p2 - p1), axis=1))


def test_image_ties_are_declined_or_decided_like_the_reference():
    """a vector on a bisector plane: if the two tied images are the shortest ones the search
    declines; if a third image is shorter still it decides, and then as the reference does"""
    per, A, B, Cc = CELLS["skewed"]
    L = 16.0
    A, B, Cc = (L * np.array(v) for v in (A, B, Cc))
    p = O.make_params(boundaries=(per, A, B, Cc))
    lib = O.lib()
    lat = np.array([A, B, Cc])
    rng = np.random.default_rng(1)
    d = []
    for k in range(13):
        s = DIRS[k] @ lat
        for sign in (1.0, -1.0):
            t = rng.normal(size=3)
            t -= s * (t @ s) / (s @ s)
            d.append(sign * 0.5 * s + 0.3 * t / np.linalg.norm(t))
    d = np.array(d)
    mine, ambiguous = general_image(p.bc, d)
    assert ambiguous.sum() >= 6
    dp = C.POINTER(C.c_double)
    zero, out = np.zeros(3), np.zeros(3)
    for k in np.flatnonzero(~ambiguous):
        b = np.ascontiguousarray(d[k])
        lib.cvo_position_distance(C.byref(p.bc), zero.ctypes.data_as(dp), b.ctypes.data_as(dp),
                                  out.ctypes.data_as(dp))
        assert np.array_equal(mine[k], out)
