"""Per-step calc_value times of the selfCoordNum 50 k pair-list case (BASELINE config 3), to see
what a rebuild costs through the all-pairs sweep and through the cell list (B200CV_CELLS=1)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import colvars_b200 as cb
from colvars_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
proxy, i1, L = synth.selfcoordnum_case(N, seed=1234)
cv = cb.CoordNum("selfCoordNum", i1, cutoff=4.0, tolerance=1e-3, pairlist_freq=10)
if len(sys.argv) > 2 and sys.argv[2] == "tric":
    # rhombic dodecahedron: cell list in fractional coordinates, image search per candidate
    cv.set_boundaries((1, 1, 1), (L, 0, 0), (0, L, 0), (0.5 * L, 0.5 * L, 0.5 * 2 ** 0.5 * L))
d_pos = torch.tensor(np.ascontiguousarray(proxy.T), device="cuda")
s = torch.cuda.current_stream()
out = []
for step in range(21 if N > 50000 else 41):
    cv.read_data(d_pos, s)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(s)
    cv.calc_value(step=step, gradients=True, stream=s)
    e1.record(s)
    v = cv.value()
    torch.cuda.synchronize()
    out.append((step, e0.elapsed_time(e1) * 1e3, cv.last_stats()[0]))
print("rebuild steps (us, launches):", [(s_, round(t), k) for s_, t, k in out if s_ % 10 == 0])
print("list steps median us:", round(float(np.median([t for s_, t, _ in out if s_ % 10]))))
print("listed pairs:", len(cv.pairlist()), "value", v)
