"""Timing experiments on the single-wave 1k x 100k configuration (BASELINE config 2):
calc_value with and without the gradient flush, and a few neighbouring shapes."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import colvars_b200 as cb
from colvars_b200 import synth


def time_calc(cv, d_pos, gradients, reps=30):
    s = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms = []
    for r in range(reps + 3):
        flush.fill_(r & 0xFF)
        cv.read_data(d_pos, s)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(s)
        cv.calc_value(step=r, gradients=gradients, stream=s)
        e1.record(s)
        torch.cuda.synchronize()
        if r >= 3:
            ms.append(e0.elapsed_time(e1))
    return float(np.median(ms)) * 1e3


for n1, n2 in ((1000, 100000), (1024, 100000), (1000, 94720), (2000, 100000), (4000, 100000),
               (500, 100000)):
    proxy, i1, i2, L = synth.coordnum_case(n1, n2, seed=1234)
    cv = cb.CoordNum("coordNum", i1, i2, cutoff=4.0)
    d_pos = torch.tensor(np.ascontiguousarray(proxy.T), device="cuda")
    tg = time_calc(cv, d_pos, True)
    tv = time_calc(cv, d_pos, False)
    items, chunk = cb.plan_items("coordNum", n1, n2)
    pairs = n1 * n2
    print(f"{n1:6d} x {n2:7d}: items {len(items):5d} chunk {chunk:5d}  value+grad {tg:8.1f} us "
          f"({pairs / tg / 1e6:7.1f} Gpair/s)   value only {tv:8.1f} us", flush=True)
    cv.close()

# host-buffer (e2e) path of the 1k x 100k case: where the wall time goes
import time
proxy, i1, i2, L = synth.coordnum_case(1000, 100000, seed=1234)
cv = cb.CoordNum("coordNum", i1, i2, cutoff=4.0)
h_pos = torch.tensor(proxy).pin_memory()
h_force = torch.zeros_like(h_pos).pin_memory()
tc, ta = [], []
for r in range(30):
    t0 = time.perf_counter()
    v = cv.calc_host(h_pos, step=r)
    t1 = time.perf_counter()
    cv.apply_force_host(-0.004 * (v - 0.1), h_force)
    t2 = time.perf_counter()
    if r >= 5:
        tc.append(t1 - t0)
        ta.append(t2 - t1)
print(f"e2e 1k x 100k: calc_host {np.median(tc) * 1e6:.1f} us, apply_force_host "
      f"{np.median(ta) * 1e6:.1f} us", flush=True)

# is the first wave slower because of a cold start?  five evaluations back to back (no L2
# flush, no gap), an event between each
proxy, i1, i2, L = synth.coordnum_case(1000, 100000, seed=1234)
cv = cb.CoordNum("coordNum", i1, i2, cutoff=4.0)
d_pos = torch.tensor(np.ascontiguousarray(proxy.T), device="cuda")
s = torch.cuda.current_stream()
cv.read_data(d_pos, s)
for rep in range(3):
    evs = [torch.cuda.Event(True) for _ in range(7)]
    evs[0].record(s)
    for k in range(6):
        cv.calc_value(step=k, gradients=True, stream=s)
        evs[k + 1].record(s)
    torch.cuda.synchronize()
    print("back-to-back calc_value us:", " ".join(f"{evs[k].elapsed_time(evs[k + 1]) * 1e3:.1f}"
                                                    for k in range(6)), flush=True)
