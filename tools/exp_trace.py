"""Per-CTA timeline of one sweep launch (debug build libb200cv_trace.so, B200CV_LIB override)."""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import colvars_b200 as cb
from colvars_b200 import synth

L = cb.load()
s = torch.cuda.current_stream()
for n1, n2 in ((1000, 94720), (1000, 100000)):
    proxy, i1, i2, _ = synth.coordnum_case(n1, n2, seed=1234)
    cv = cb.CoordNum("coordNum", i1, i2, cutoff=4.0)
    d_pos = torch.tensor(np.ascontiguousarray(proxy.T), device="cuda")
    cv.read_data(d_pos, s)
    items, chunk = cb.plan_items("coordNum", n1, n2)
    for r in range(4):
        cv.calc_value(step=r, gradients=True, stream=s)
        torch.cuda.synchronize()
    n = len(items)
    buf = np.zeros(6 * n, dtype=np.uint64)
    rc = L.b200cv_trace_dump(buf.ctypes.data_as(C.c_void_p), n)
    assert rc == 0
    t = buf.reshape(n, 6).astype(np.int64)
    items = [items[k] for k in t[:, 5]]  # the item each CTA actually took
    t0, t1, t2, t3 = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
    smid = t[:, 4]
    base = t0.min()
    print(f"{n} items x {chunk // 32} tiles: kernel span {(t3.max() - base) / 1e3:.1f} us")
    print(f"  CTA start  min/med/max {((t0 - base).min()) / 1e3:.1f} {np.median(t0 - base) / 1e3:.1f} {(t0 - base).max() / 1e3:.1f} us")
    print(f"  prologue   med/max {np.median(t1 - t0) / 1e3:.2f} {(t1 - t0).max() / 1e3:.2f} us")
    print(f"  tile loop  min/med/max {(t2 - t1).min() / 1e3:.1f} {np.median(t2 - t1) / 1e3:.1f} {(t2 - t1).max() / 1e3:.1f} us")
    print(f"  epilogue   med/max {np.median(t3 - t2) / 1e3:.2f} {(t3 - t2).max() / 1e3:.2f} us")
    print(f"  CTA end    min/med/max {(t3 - base).min() / 1e3:.1f} {np.median(t3 - base) / 1e3:.1f} {(t3 - base).max() / 1e3:.1f} us")
    cnt = np.bincount(smid, minlength=148)
    print(f"  CTAs per SM: min {cnt[:148].min()} max {cnt.max()}  (SMs used {np.count_nonzero(cnt)})")
    loop = (t2 - t1) / 1e3
    full = np.array([it[2] == chunk for it in items])
    tiles = np.array([(it[2] + 31) // 32 for it in items])
    for k in np.unique(tiles):
        m = tiles == k
        print(f"  items of {k:2d} tiles: {m.sum():4d}, blocks {np.flatnonzero(m).min()}..{np.flatnonzero(m).max()}, "
              f"tile loop med {np.median(loop[m]):.1f} us = {np.median(loop[m]) / k:.2f} us/tile, "
              f"end med {np.median((t3 - base)[m]) / 1e3:.1f}")
    print("  tile-loop histogram (full items):", np.percentile(loop[full], [0, 10, 25, 50, 75, 90, 100]).round(1))
    # per SM: the two co-resident CTAs
    per_sm = {}
    for b in range(n):
        per_sm.setdefault(int(smid[b]), []).append((b, loop[b]))
    pairs = [v for v in per_sm.values() if len(v) == 2]
    d = np.array([[a[1], b[1]] for a, b in pairs])
    print("  co-resident pairs: mean |difference|", np.abs(d[:, 0] - d[:, 1]).mean().round(1),
          "us; mean of slower", d.max(1).mean().round(1), "faster", d.min(1).mean().round(1))
    slow = np.argsort(-loop)[:8]
    print("  slowest CTAs (block, sm, us):", [(int(b), int(smid[b]), round(float(loop[b]), 1)) for b in slow])
    fast = np.argsort(loop)[:8]
    print("  fastest CTAs (block, sm, us):", [(int(b), int(smid[b]), round(float(loop[b]), 1)) for b in fast])
    cv.close()
