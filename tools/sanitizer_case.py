import numpy as np, sys
sys.path.insert(0, ".")
import colvars_b200 as cb
from tests import oracle_lib as O
rng = np.random.default_rng(1)
pos = np.ascontiguousarray(rng.random((1500 + 2600, 3)) * 20.0)
i1 = np.arange(1500, dtype=np.int32); i2 = np.arange(1500, 4100, dtype=np.int32)
for kw in (dict(), dict(tolerance=0.02, pairlist_freq=2), dict(cutoff3=(3.0, 5.0, 4.0), exp_numer=8, exp_denom=10)):
    cv = cb.CoordNum("coordNum", i1, i2, **kw)
    for step in range(3):
        v = cv.calc_host(pos, step=step)
    f = np.zeros_like(pos); cv.apply_force_host(0.1, f)
    print("coordNum", kw, v)
    cv.close()
cv = cb.CoordNum("selfCoordNum", i1, tolerance=0.01, pairlist_freq=2)
for step in range(3):
    v = cv.calc_host(pos, step=step)
print("self", v)
cv = cb.CoordNum("coordNum", i1[:40], i2[:300], cutoff=4.0)
cv.set_fitting(1, center=True, rotate=True, fitting_group=None, ref_pos=pos[:40] + 1.0)
v = cv.calc_host(pos, step=0); f = np.zeros_like(pos); cv.apply_force_host(0.1, f)
print("fit", v)
print("SANITIZER-SCRIPT-DONE")
