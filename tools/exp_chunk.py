"""calc_value time of a configuration for the chunk given in B200CV_CHUNK (tuning experiment)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
import colvars_b200 as cb
from colvars_b200 import synth
n1, n2 = int(sys.argv[1]), int(sys.argv[2])
s = torch.cuda.current_stream()
proxy, i1, i2, _ = synth.coordnum_case(n1, n2, seed=1234)
cv = cb.CoordNum("coordNum", i1, i2, cutoff=4.0)
d_pos = torch.tensor(np.ascontiguousarray(proxy.T), device="cuda")
cv.read_data(d_pos, s)
items, chunk = cb.plan_items("coordNum", n1, n2)
ts = []
for r in range(25):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(s)
    cv.calc_value(step=r, gradients=True, stream=s)
    e1.record(s)
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"{n1} x {n2}: chunk {chunk:5d} ({chunk // 32:2d} tiles) items {len(items):6d}: "
      f"calc_value {np.median(ts[5:]):8.1f} us", flush=True)
