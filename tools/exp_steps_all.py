"""calc_value of the selfCoordNum pair-list case (BASELINE config 3 at N atoms), every step:
GPU time (events), host time to enqueue, host time until the value is there."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import colvars_b200 as cb
from colvars_b200 import synth
N = int(sys.argv[1])
proxy, i1, L = synth.selfcoordnum_case(N, seed=1234)
cv = cb.CoordNum("selfCoordNum", i1, cutoff=4.0, tolerance=1e-3, pairlist_freq=10)
d_pos = torch.tensor(np.ascontiguousarray(proxy.T), device="cuda")
s = torch.cuda.current_stream()
out = []
import time
for step in range(31):
    cv.read_data(d_pos, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.perf_counter()
    e0.record(s)
    cv.calc_value(step=step, gradients=True, stream=s)
    e1.record(s)
    t1 = time.perf_counter()
    v = cv.value()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out.append((step, round(e0.elapsed_time(e1) * 1e3), round((t1 - t0) * 1e6), round((t2 - t0) * 1e6)))
print("step, gpu us, host enqueue us, host total us")
print(out)
lst = [g for st, g, _, _ in out if st > 10 and st % 10]
reb = [g for st, g, _, _ in out if st > 10 and st % 10 == 0]
print("list steps gpu us: median", float(np.median(lst)), "min", min(lst), "; rebuild steps:", reb)
