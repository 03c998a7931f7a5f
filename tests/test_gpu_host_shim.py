"""The drop-in test: the UNMODIFIED Colvars library (parser, colvar, atom groups, harmonic
bias, stub proxy, trajectory writer) with only the three coordination-number components
replaced by colvars_b200/host/coordnum_b200 (-> C ABI -> CUDA).  Same command line, same
output files and the same golden files as the reference's own functional tests
(tests/input_files/<case>/AutoDiff, run through tests/functional/run_colvars_test); tolerance
of the reference's comparer (tests/input_files/compare_test.sh: relative 1e-6, absolute
1e-12).  debugGradients is on in most of these inputs, so the reference's own finite-difference
check of our analytical gradients also has to stay below its 1e-3 threshold.

The runner binary is built in the build container (it needs the reference headers) and
travels to the GPU box prebuilt."""
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from . import regression as R

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
RUNNER = ROOT / "colvars_b200" / "host" / "_build" / "run_colvars_b200"
RUNNER_GPU = ROOT / "colvars_b200" / "host" / "_build" / "run_colvars_b200_gpu"


def _close(a, b):
    return np.all(np.abs(a - b) <= np.maximum(1e-6 * np.abs(b), 1e-12))


EXTRA = R.GOLDEN.parent / "colvars_regression_extra"
EXTRA_CASES = sorted(p.name[:-len("_harmonic-fixed")] for p in EXTRA.glob("*_harmonic-fixed"))


@pytest.mark.parametrize("route", ["cpu-proxy", "smp-gpu"])
@pytest.mark.parametrize("case", R.CASES + R.PAIR_CASES + ["extra:" + c for c in EXTRA_CASES])
def test_reference_host_with_b200_components(case, route, tmp_path):
    """route = cpu-proxy: CPU proxy arrays, components call b200cv_eval_host.
    route = smp-gpu: the reference built with COLVARS_CUDA, device-resident proxy, the
    reference's graph scheduler; components without a pair list contribute graph nodes
    (b200cv_eval_async / b200cv_accumulate_gradients; with a pair list the same graph holds
    the rebuild sweep and the list walk); only a dummy-atom group2 takes the reference's
    CPU-buffer route."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    RUNNER = RUNNER_GPU if route == "smp-gpu" else globals()["RUNNER"]
    if not RUNNER.exists():
        pytest.fail(f"{RUNNER} missing: build it with __graft_entry__.build() where "
                    "/root/reference exists")
    # "extra:" cases: configurations the reference's coordnum tests do not exercise (dummy-atom
    # group2, fitted group1, general exponents ...); goldens made by running the unmodified
    # reference (tests/golden/make_extra_regression.py)
    golden_root = R.GOLDEN
    if case.startswith("extra:"):
        case = case[len("extra:"):]
        golden_root = EXTRA
    for f in ("trajectory.xyz", "index.ndx"):
        shutil.copy(golden_root / f, tmp_path / f)
    d = tmp_path / f"{case}_harmonic-fixed"
    d.mkdir()
    shutil.copy(golden_root / f"{case}_harmonic-fixed" / "test.in", d / "test.in")
    r = subprocess.run([str(RUNNER), "-c", f"{d.name}/test.in", "-t", "trajectory.xyz", "-o",
                        f"{d.name}/test_out", "--force"], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    gold = golden_root / f"{case}_harmonic-fixed" / "AutoDiff"
    traj = np.loadtxt(d / "test_out.colvars.traj", comments="#")
    gtraj = np.loadtxt(gold / "test_out.colvars.traj", comments="#")
    assert traj.shape == gtraj.shape and _close(traj, gtraj), (traj, gtraj)
    for k in range(5):
        f = np.loadtxt(d / f"test_out_forces_{k}.dat")
        g = np.loadtxt(gold / f"test_out_forces_{k}.dat")
        assert f.shape == g.shape and _close(f, g)
    # beyond the reference comparer's 1e-6: the cpu-proxy route must print (nearly) the same
    # text; on the smp-gpu route the reference's own GPU kernels (parallel COM sums, its GPU
    # eigen-solver for fitted groups) feed positions that differ by ulps, so compare to 1e-9
    if route == "cpu-proxy":
        same = sum(a == b for a, b in zip((d / "test_out.colvars.traj").read_text().split(),
                                          (gold / "test_out.colvars.traj").read_text().split()))
        total = len((gold / "test_out.colvars.traj").read_text().split())
        assert same >= total - 2, (same, total)
    else:
        assert np.all(np.abs(traj - gtraj) <= 1e-9 * np.maximum(np.abs(gtraj), 1e-3))
    if "debugGradients" in (d / "test.in").read_text():
        assert "Max gradient error" in r.stdout
    if route == "smp-gpu":
        # gather / COM / fit / scatter graphs run on this repository's kernels, not on the
        # archive member built from src/cuda/colvaratoms_kernel.cu
        assert "atom-group kernels: b200" in r.stdout
        assert "optimal-rotation kernels: b200" in r.stdout
        on_graph = "graph route: yes" in r.stdout
        # one acceptor-donor pair (hBond), the pair batch of alpha and a dummy-atom group2 stay
        # on the CPU-buffer route
        assert on_graph == ("dummy" not in case and case != "hbond" and
                            not case.startswith("alpha")), r.stdout[-500:]


@pytest.mark.parametrize("case", ["coordnum", "selfcoordnum-pairlist", "coordnum-aniso"])
def test_periodic_cell_through_the_reference_host(case, tmp_path):
    """The reference's functional tests are non-periodic.  Here the same inputs run in a small
    orthorhombic cell (minimum images matter for most pairs) three times: the reference's own
    components, the B200 components on CPU-proxy buffers, and on the `smp gpu` graph route
    (cell taken from the proxy at capture time) -- all through the unmodified host."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    for f in ("trajectory.xyz", "index.ndx"):
        shutil.copy(R.GOLDEN / f, tmp_path / f)
    d = tmp_path / "case"
    d.mkdir()
    shutil.copy(R.GOLDEN / f"{case}_harmonic-fixed" / "test.in", d / "test.in")
    cell = ["--cell", "3.1", "2.7", "2.9"]
    runs = {"ref": [str(RUNNER), "--reference-components"], "cpu": [str(RUNNER)],
            "gpu": [str(RUNNER_GPU)]}
    out = {}
    for tag, cmd in runs.items():
        r = subprocess.run(cmd + cell + ["-c", "case/test.in", "-t", "trajectory.xyz", "-o",
                                         f"case/{tag}", "--force"], cwd=tmp_path,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, tag + r.stdout[-2000:] + r.stderr[-2000:]
        out[tag] = (np.loadtxt(d / f"{tag}.colvars.traj", comments="#"),
                    [np.loadtxt(d / f"{tag}_forces_{k}.dat") for k in range(5)])
    # the cell matters: not the non-periodic golden
    gold = np.loadtxt(R.GOLDEN / f"{case}_harmonic-fixed" / "AutoDiff" / "test_out.colvars.traj",
                      comments="#")
    assert np.max(np.abs(out["ref"][0][:, 1] - gold[:, 1])) > 1e-3
    for tag in ("cpu", "gpu"):
        assert np.allclose(out[tag][0], out["ref"][0], rtol=1e-11, atol=1e-13), tag
        for k in range(5):
            assert np.allclose(out[tag][1][k], out["ref"][1][k], rtol=1e-9, atol=1e-13), (tag, k)


@pytest.mark.parametrize("route", ["cpu-proxy", "smp-gpu"])
@pytest.mark.parametrize("case", R.CASES + R.PAIR_CASES)
def test_sharded_components_through_the_reference_host(case, route, tmp_path):
    """Two processes, one per GPU, each running the whole unmodified host on the same
    trajectory; the environment (B200CV_NCCL_ID_FILE, RANK, WORLD_SIZE, LOCAL_RANK) makes every
    coordination-number component shard its group1 rows over the two ranks and all-reduce
    [value | gradients] (coordnum_b200.cpp, join_shards).  Both ranks have to reproduce the
    reference's golden files; nothing else in the host knows about the sharding."""
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    runner = RUNNER_GPU if route == "smp-gpu" else RUNNER
    if not runner.exists():
        pytest.fail(f"{runner} missing: build it with __graft_entry__.build()")
    import os
    procs = []
    for rank in range(2):
        wd = tmp_path / f"rank{rank}"
        wd.mkdir()
        for f in ("trajectory.xyz", "index.ndx"):
            shutil.copy(R.GOLDEN / f, wd / f)
        d = wd / f"{case}_harmonic-fixed"
        d.mkdir()
        shutil.copy(R.GOLDEN / f"{case}_harmonic-fixed" / "test.in", d / "test.in")
        env = dict(os.environ, B200CV_NCCL_ID_FILE=str(tmp_path / "nccl_id"), RANK=str(rank),
                   WORLD_SIZE="2", LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen(
            [str(runner), "-c", f"{d.name}/test.in", "-t", "trajectory.xyz", "-o",
             f"{d.name}/test_out", "--force"], cwd=wd, env=env, stdout=subprocess.PIPE,
            stderr=subprocess.STDOUT, text=True))
    try:
        outs = [p.communicate(timeout=120)[0] for p in procs]
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    gold = R.GOLDEN / f"{case}_harmonic-fixed" / "AutoDiff"
    gtraj = np.loadtxt(gold / "test_out.colvars.traj", comments="#")
    for rank in range(2):
        assert procs[rank].returncode == 0, outs[rank][-3000:]
        d = tmp_path / f"rank{rank}" / f"{case}_harmonic-fixed"
        traj = np.loadtxt(d / "test_out.colvars.traj", comments="#")
        assert traj.shape == gtraj.shape and _close(traj, gtraj), (rank, traj, gtraj)
        for k in range(5):
            f = np.loadtxt(d / f"test_out_forces_{k}.dat")
            g = np.loadtxt(gold / f"test_out_forces_{k}.dat")
            assert f.shape == g.shape and _close(f, g), (rank, k)
        if case != "hbond":
            assert "sharded, rank %d of 2" % rank in outs[rank], outs[rank][-1500:]


@pytest.mark.parametrize("late", [False, True])
@pytest.mark.parametrize("case", ["coordnum", "selfcoordnum-pairlist", "coordnum-aniso"])
def test_cell_that_changes_every_step_on_the_graph_route(case, late, tmp_path):
    """cvc::read_data copies the periodic cell from the proxy every step
    (src/colvarcomp.cpp:457-465); a captured graph freezes the one it was built with.  Here the
    cell grows by 1 % per frame (a barostat).  The graph carries a host node that notices the
    change and switches the captured kernels off; the component then evaluates that step and
    the following ones directly with the proxy's cell of now.  late: the new cell arrives only through
    wait_for_extra_info_ready(), between the read_data and the calc_value graphs, the way
    NAMD's CudaGlobalMaster delivers its lattice.  Reference: the unmodified components on the
    CPU with the same cells."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    for f in ("trajectory.xyz", "index.ndx"):
        shutil.copy(R.GOLDEN / f, tmp_path / f)
    d = tmp_path / "case"
    d.mkdir()
    shutil.copy(R.GOLDEN / f"{case}_harmonic-fixed" / "test.in", d / "test.in")
    cell = ["--cell", "3.1", "2.7", "2.9", "--cell-drift", "0.01"]
    runs = {"ref": [str(RUNNER), "--reference-components"] + cell,
            "gpu": [str(RUNNER_GPU)] + cell + (["--late-lattice"] if late else [])}
    out = {}
    for tag, cmd in runs.items():
        r = subprocess.run(cmd + ["-c", "case/test.in", "-t", "trajectory.xyz", "-o",
                                  f"case/{tag}", "--force"], cwd=tmp_path,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, tag + r.stdout[-2000:] + r.stderr[-2000:]
        out[tag] = (np.loadtxt(d / f"{tag}.colvars.traj", comments="#"),
                    [np.loadtxt(d / f"{tag}_forces_{k}.dat") for k in range(5)], r.stdout)
    assert "the periodic cell has changed since the graph was captured" in out["gpu"][2]
    # the drift matters: frame 4 is not what the fixed cell gives
    fixed = subprocess.run([str(RUNNER), "--reference-components", "--cell", "3.1", "2.7", "2.9",
                            "-c", "case/test.in", "-t", "trajectory.xyz", "-o", "case/fixed"],
                           cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert fixed.returncode == 0
    ftraj = np.loadtxt(d / "fixed.colvars.traj", comments="#")
    assert abs(ftraj[4, 1] - out["ref"][0][4, 1]) > 1e-6
    assert np.allclose(out["gpu"][0], out["ref"][0], rtol=1e-11, atol=1e-13)
    for k in range(5):
        assert np.allclose(out["gpu"][1][k], out["ref"][1][k], rtol=1e-9, atol=1e-13), k


@pytest.mark.parametrize("case", ["coordnum-pairlist", "selfcoordnum-pairlist", "coordnum"])
def test_overflow_of_a_captured_evaluation_recovers_on_the_graph_route(case, tmp_path):
    """A queue of a captured evaluation cannot grow under the graph that holds its address;
    the C ABI reports the overflow (tests/test_gpu_parity.py::test_captured_pairlist_overflow...).
    The component recovers from it: it evaluates that step outside the graph -- where queues
    grow -- and keeps its captured kernels (they hold the old addresses) switched off until the
    graphs are built again.  The da-ala system is too
    small to overflow anything, so the overflow of step 2 is injected
    (B200CV_SHIM_FAKE_OVERFLOW_STEP); everything downstream of the report is the real path."""
    import os
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    for f in ("trajectory.xyz", "index.ndx"):
        shutil.copy(R.GOLDEN / f, tmp_path / f)
    d = tmp_path / f"{case}_harmonic-fixed"
    d.mkdir()
    shutil.copy(R.GOLDEN / f"{case}_harmonic-fixed" / "test.in", d / "test.in")
    r = subprocess.run([str(RUNNER_GPU), "-c", f"{d.name}/test.in", "-t", "trajectory.xyz", "-o",
                        f"{d.name}/test_out", "--force"], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300,
                       env=dict(os.environ, B200CV_SHIM_FAKE_OVERFLOW_STEP="2"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "evaluating this step directly" in r.stdout
    gold = R.GOLDEN / f"{case}_harmonic-fixed" / "AutoDiff"
    traj = np.loadtxt(d / "test_out.colvars.traj", comments="#")
    gtraj = np.loadtxt(gold / "test_out.colvars.traj", comments="#")
    assert traj.shape == gtraj.shape and _close(traj, gtraj), (traj, gtraj)
    for k in range(5):
        f = np.loadtxt(d / f"test_out_forces_{k}.dat")
        g = np.loadtxt(gold / f"test_out_forces_{k}.dat")
        assert f.shape == g.shape and _close(f, g), k


@pytest.mark.parametrize("route", ["cpu-proxy", "smp-gpu"])
def test_distance_pairs_through_the_reference_host(route, tmp_path):
    """colvar::distance_pairs replaced by distancepairs_b200: the reference's own regression
    input distancepairs_linear-distancepairs (a 16-component vector under a linear bias) and its
    golden files, through the unmodified host on both routes."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    runner = RUNNER_GPU if route == "smp-gpu" else RUNNER
    for f in ("trajectory.xyz", "index.ndx"):
        shutil.copy(R.GOLDEN / f, tmp_path / f)
    name = "distancepairs_linear-distancepairs"
    d = tmp_path / name
    d.mkdir()
    shutil.copy(R.GOLDEN / name / "test.in", d / "test.in")
    r = subprocess.run([str(runner), "-c", f"{name}/test.in", "-t", "trajectory.xyz", "-o",
                        f"{name}/test_out", "--force"], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    gold = R.GOLDEN / name / "AutoDiff"
    import re
    num = re.compile(r"[-+]?\d\.\d+e[-+]\d+")

    def rows(path):
        return np.array([[float(v) for v in num.findall(line)]
                         for line in path.read_text().splitlines()
                         if line.strip() and not line.startswith("#")])
    traj, gtraj = rows(d / "test_out.colvars.traj"), rows(gold / "test_out.colvars.traj")
    assert traj.shape == gtraj.shape == (5, 32) and _close(traj, gtraj)
    for k in range(5):
        f = np.loadtxt(d / f"test_out_forces_{k}.dat")
        g = np.loadtxt(gold / f"test_out_forces_{k}.dat")
        assert f.shape == g.shape and _close(f, g), k

