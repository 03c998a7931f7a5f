"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): one process per GPU through
torchrun, NCCL all-reduce of [value | grad1 | grad2] inside b200cv_calc_value, every rank must
end with the full value and the full gradients of the unsharded oracle."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
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
if kind.endswith("-pairlist"):
    # sparse pair list, rebuilt every 2 steps: the first rebuild is a sharded sweep, the later
    # ones go through the cell list on each rank's rows (csrc/cells.cu)
    kind = kind[:-len("-pairlist")]
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(23)
    pos0 = np.ascontiguousarray(rng.random((n1 + n2, 3)) * 75.0)
    i1 = np.arange(n1, dtype=np.int32)
    i2 = np.arange(n1, n1 + n2, dtype=np.int32)
    self_ = kind == "selfCoordNum"
    cv = cb.CoordNum(kind, i1, None if self_ else i2, cutoff=4.0, tolerance=0.01, pairlist_freq=2,
                     device=int(os.environ["LOCAL_RANK"]))
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(cb.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    cv.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    p = O.make_params(tolerance=0.01)
    npairs = n1 * (n1 - 1) // 2 if self_ else n1 * n2
    pl = np.ones(npairs, dtype=np.uint8)
    cells_used = 0
    for step in range(5):
        pos = pos0 + 0.1 * step * np.sin(pos0 + step)
        d_pos = torch.tensor(np.ascontiguousarray(pos.T), device=dev)
        cv.read_data(d_pos)
        cv.calc_value(step=step)
        v = cv.value()
        rebuild = step % 2 == 0
        if self_:
            ov, og1 = O.selfcoordnum(p, pos, pairlist=pl, rebuild=rebuild)
        else:
            ov, og1, og2 = O.coordnum(p, pos[i1], pos[i2], pairlist=pl, rebuild=rebuild)
            g2 = cv.gradients(2)
            assert np.max(np.abs(g2 - og2)) <= 1e-12 * np.max(np.abs(og2)), (rank, step)
        assert abs(v - ov) <= 1e-12 * abs(ov), (rank, step, v, ov)
        g1 = cv.gradients(1)
        assert np.max(np.abs(g1 - og1)) <= 1e-12 * np.max(np.abs(og1)), (rank, step)
        if rebuild and cv.last_stats()[0] >= 6:
            cells_used += 1
    assert cells_used == 2, cells_used
    # the ranks' lists together are the reference's list
    mine = torch.tensor(cv.pairlist().astype(np.int64), device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.numel()], device=dev))
    total = int(sum(int(x.item()) for x in sizes))
    assert total == int(pl.sum()), (total, int(pl.sum()))
    print(f"MULTI-OK rank {rank} {kind} pair list value {v:.15e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
rng = np.random.default_rng(17)
pos = np.ascontiguousarray(rng.random((n1 + n2, 3)) * 40.0)
i1 = np.arange(n1, dtype=np.int32)
i2 = np.arange(n1, n1 + n2, dtype=np.int32)
self_ = kind == "selfCoordNum"
cv = cb.CoordNum(kind, i1, None if self_ else i2, cutoff=4.0, device=int(os.environ["LOCAL_RANK"]))
idt = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(cb.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
cv.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
d_pos = torch.tensor(np.ascontiguousarray(pos.T), device=dev)
d_force = torch.zeros_like(d_pos)
for step in range(2):
    cv.read_data(d_pos)
    cv.calc_value(step=step)
    v = cv.value()
cv.apply_force(0.5, d_force)
torch.cuda.synchronize()
p = O.make_params()
if kind == "distanceInv":
    # the sum over ranks comes before the non-linear epilogue (DESIGN.md section 9)
    ov, og1, og2 = O.distance_inv(pos[i1], pos[i2], 6, omp=4)
elif self_:
    ov, og1 = O.selfcoordnum(p, pos, omp=4)
    og2 = None
else:
    ov, og1, og2 = O.coordnum(p, pos[i1], pos[i2], omp=4)
assert abs(v - ov) <= 1e-12 * abs(ov), (rank, v, ov)
g1 = cv.gradients(1)
assert np.max(np.abs(g1 - og1)) <= 1e-12 * np.max(np.abs(og1)), rank
expect = np.zeros_like(pos)
expect[i1] += 0.5 * og1
if not self_:
    g2 = cv.gradients(2)
    assert np.max(np.abs(g2 - og2)) <= 1e-12 * np.max(np.abs(og2)), rank
    expect[i2] += 0.5 * og2
f = d_force.cpu().numpy().T
assert np.max(np.abs(f - expect)) <= 1e-12 * np.max(np.abs(expect)), rank
print(f"MULTI-OK rank {rank} {kind} value {v:.15e}", flush=True)
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("kind,n1,n2", [("coordNum", 5000, 20000), ("coordNum", 300, 30000),
                                        ("selfCoordNum", 9000, 0),
                                        ("distanceInv", 3000, 40000),
                                        ("selfCoordNum-pairlist", 4000, 0),
                                        ("coordNum-pairlist", 900, 5000)])
def test_sharded_run_matches_oracle_on_every_rank(kind, n1, n2, tmp_path):
    import torch
    ngpu = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ngpu < 4 else 4
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO=str(ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), str(script), kind, str(n1), str(n2)],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and r.stdout.count("MULTI-OK") == world, \
        r.stdout[-2000:] + r.stderr[-3000:]
