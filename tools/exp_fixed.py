"""Fixed cost of one wave of the tile sweep: 296 work items of 1, 2, 4, 10 tiles each."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import colvars_b200 as cb
from colvars_b200 import synth

s = torch.cuda.current_stream()
for n2 in (9472, 18944, 37888, 94720):
    proxy, i1, i2, L = synth.coordnum_case(1000, n2, seed=1234)
    cv = cb.CoordNum("coordNum", i1, i2, cutoff=4.0)
    d_pos = torch.tensor(np.ascontiguousarray(proxy.T), device="cuda")
    cv.read_data(d_pos, s)
    items, chunk = cb.plan_items("coordNum", 1000, n2)
    ts = []
    for r in range(20):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(s)
        cv.calc_value(step=r, gradients=True, stream=s)
        e1.record(s)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"n2 {n2:6d}: {len(items)} items x {chunk // 32} tiles: calc_value {np.median(ts[5:]):.1f} us",
          flush=True)
