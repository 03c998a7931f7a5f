"""Host-side logic of the multi-GPU path, on CPU: two `gloo` ranks take the work-item
tables the library plans for them (b200cv_plan_items -- the same code that feeds the
sweep kernel), evaluate exactly those items with the CPU oracle, all-reduce
[value | grad1 | grad2] like b200cv_calc_value does with NCCL, and must reproduce the
unsharded oracle.  Also: the tables of all ranks tile the pair space exactly once."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["REPO"])
import colvars_b200 as cb
from tests import oracle_lib as O

kind = sys.argv[1]
n1, n2 = int(sys.argv[2]), int(sys.argv[3])
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(5)
pos = rng.random((n1 + n2, 3)) * 25.0
p = O.make_params((4.0, 4.0, 4.0))
items, chunk = cb.plan_items(kind, n1, n2, rank, world)
self_ = kind == "selfCoordNum"
p1 = pos[:n1]
p2 = p1 if self_ else pos[n1:]
g1 = np.zeros((n1, 3))
g2 = np.zeros((len(p2), 3))
value = 0.0
IB = 1024
for i0, j0, jn, diag in items:
    i1 = min(i0 + IB, n1)
    if not self_:
        v, a, b = O.coordnum(p, p1[i0:i1], p2[j0:j0 + jn])
        value += v
        g1[i0:i1] += a
        g2[j0:j0 + jn] += b
    else:
        # rows i0..i1 against columns j0..j0+jn, pairs with j > i only
        for i in range(i0, i1):
            lo = max(j0, i + 1)
            if lo >= j0 + jn:
                continue
            v, a, b = O.coordnum(p, p1[i:i + 1], p2[lo:j0 + jn])
            value += v
            g1[i] += a[0]
            g1[lo:j0 + jn] += b
buf = torch.from_numpy(np.concatenate([[value], g1.ravel(), (g2 if not self_ else g2[:0]).ravel()]))
dist.all_reduce(buf)
if rank == 0:
    if self_:
        ov, og1 = O.selfcoordnum(p, p1)
        og2 = np.zeros((0, 3))
    else:
        ov, og1, og2 = O.coordnum(p, p1, p2)
    out = buf.numpy()
    assert abs(out[0] - ov) <= 1e-12 * abs(ov), (out[0], ov)
    r1 = out[1:1 + 3 * n1].reshape(n1, 3)
    assert np.max(np.abs(r1 - og1)) <= 1e-12 * np.max(np.abs(og1))
    if not self_:
        r2 = out[1 + 3 * n1:].reshape(n2, 3)
        assert np.max(np.abs(r2 - og2)) <= 1e-12 * np.max(np.abs(og2))
    print("SHARD-OK", kind, len(items))
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("kind,n1,n2", [("coordNum", 2300, 700), ("coordNum", 300, 4000),
                                        ("selfCoordNum", 2600, 0)])
def test_two_gloo_ranks_reproduce_the_unsharded_result(kind, n1, n2, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO=str(ROOT), OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                        str(_free_port()), str(script), kind, str(n1), str(n2)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "SHARD-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.parametrize("kind,n1,n2,world", [("coordNum", 100000, 1000000, 8),
                                              ("coordNum", 1000, 100000, 8),
                                              ("coordNum", 5000, 3000, 4),
                                              ("selfCoordNum", 50000, 0, 8),
                                              ("selfCoordNum", 3000, 0, 2)])
def test_item_tables_tile_the_pair_space_once(kind, n1, n2, world):
    import colvars_b200 as cb
    self_ = kind == "selfCoordNum"
    ncols = n1 if self_ else n2
    covered = 0
    seen = set()
    loads = []
    for rank in range(world):
        items, chunk = cb.plan_items(kind, n1, n2, rank, world)
        assert chunk % 32 == 0 and 0 < chunk <= 1024
        pairs = 0
        for i0, j0, jn, diag in items.tolist():
            key = (int(i0), int(j0))
            assert key not in seen
            seen.add(key)
            assert i0 % 1024 == 0 and j0 % 32 == 0 and 0 < jn <= chunk and j0 + jn <= ncols
            rows = min(1024, n1 - i0)
            if not self_:
                pairs += rows * jn
            else:
                # pairs (i, j) with i in [i0, i0+rows), j in [j0, j0+jn), j > i
                i = np.arange(i0, i0 + rows)
                pairs += int(np.clip(j0 + jn - np.maximum(j0, i + 1), 0, None).sum())
                assert bool(diag) == (j0 < i0 + 1024)
        loads.append(pairs)
        covered += pairs
    total = n1 * (n1 - 1) // 2 if self_ else n1 * n2
    assert covered == total
    # balance: no rank more than 15 % above the mean (selfCoordNum is cyclic over the triangle)
    assert max(loads) <= 1.15 * (total / world) + 1024 * 1024
