"""calc_value times of one coordNum problem under the four kinds of boundary conditions
(src/colvars_system.h:161-219): non-periodic, orthorhombic, triclinic, mixed (x and y periodic)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import colvars_b200 as cb
from colvars_b200 import synth

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n2 = int(sys.argv[2]) if len(sys.argv) > 2 else 100000


def time_calc(cv, d_pos, reps=8):
    s = torch.cuda.current_stream()
    ms = []
    for r in range(reps):
        cv.read_data(d_pos, s)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(s)
        cv.calc_value(step=r, gradients=True, stream=s)
        e1.record(s)
        torch.cuda.synchronize()
        if r >= 3:
            ms.append(e0.elapsed_time(e1))
    return float(np.median(ms)) * 1e3


proxy, i1, i2, L = synth.coordnum_case(n1, n2, seed=1234)
d_pos = torch.tensor(np.ascontiguousarray(proxy.T), device="cuda")
cells = {
    "none": None,
    "ortho": ((1, 1, 1), (L, 0, 0), (0, L, 0), (0, 0, L)),
    "triclinic": ((1, 1, 1), (L, 0, 0), (0.3 * L, 0.95 * L, 0), (-0.2 * L, 0.25 * L, 0.9 * L)),
    "mixed-xy": ((1, 1, 0), (L, 0, 0), (0, L, 0), (0, 0, L)),
    "dodecahedron": ((1, 1, 1), (L, 0, 0), (0, L, 0), (0.5 * L, 0.5 * L, 0.5 * 2 ** 0.5 * L)),
}
import os
if len(sys.argv) > 3 and sys.argv[3] == "dense":
    os.environ["B200CV_DENSE"] = "1"  # the one-pair-per-thread reference-order kernel
for name, c in cells.items():
    cv = cb.CoordNum("coordNum", i1, i2, cutoff=4.0)
    if c is not None:
        cv.set_boundaries(*c)
    t = time_calc(cv, d_pos)
    print(f"{name:10s} {n1} x {n2}: {t:10.1f} us  {n1 * n2 / t / 1e3:8.2f} Gpair/s  "
          f"value {cv.value():.12e}", flush=True)
    cv.close()
