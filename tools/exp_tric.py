"""A few steps of the triclinic workload of bench.py (coordNum 10k x 100k in a rhombic
dodecahedron) for ncu captures of the general-cell sweep."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
import colvars_b200 as cb
from colvars_b200 import synth

w = bench.WORKLOADS["tric"]
proxy, i1, i2 = bench.make_inputs(w)
cv = cb.CoordNum(w["kind"], i1, i2, cutoff3=w["cutoff3"], tolerance=w["tol"], pairlist_freq=w["freq"])
L = synth.box_length(w["n2"])
cv.set_boundaries((1, 1, 1), (L, 0, 0), (0, L, 0), (0.5 * L, 0.5 * L, 0.5 * 2 ** 0.5 * L))
d_pos = torch.tensor(np.ascontiguousarray(proxy.T), device="cuda")
s = torch.cuda.current_stream()
for step in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    cv.read_data(d_pos, s)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(s)
    cv.calc_value(step=step, gradients=True, stream=s)
    e1.record(s)
    v = cv.value()
    torch.cuda.synchronize()
    print(step, "calc_value ms", e0.elapsed_time(e1), "value", v, "launches, exact pairs", cv.last_stats())
