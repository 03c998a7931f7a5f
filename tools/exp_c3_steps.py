"""A few steps of the selfCoordNum 50 k pair-list case (BASELINE config 3) for a kernel launch
list (run under `ncu --metrics gpu__time_duration.sum`): steps 0..12, i.e. the all-pairs rebuild,
nine list steps, the sparse rebuild and two more list steps.  Prints the form of the list."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import colvars_b200 as cb
from colvars_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 13
proxy, i1, L = synth.selfcoordnum_case(N, seed=1234)
cv = cb.CoordNum("selfCoordNum", i1, cutoff=4.0, tolerance=1e-3, pairlist_freq=10)
pos = proxy.copy()
s = torch.cuda.current_stream()
for step in range(nsteps):
    d_pos = torch.tensor(np.ascontiguousarray(pos.T), device="cuda")
    cv.read_data(d_pos, s)
    cv.calc_value(step=step, gradients=True, stream=s)
    v = cv.value()
    if step % 10 == 0:
        st = cv.pairlist_stats()
        print(step, st, "launches", cv.last_stats()[0], flush=True)
    pos = synth.displace(pos, seed=1000 + step, sigma=0.05)
print("value", v)
