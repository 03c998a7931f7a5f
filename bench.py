#!/usr/bin/env python3
"""bench.py -- coordNum pair-evaluations / second (value + gradient) on B200.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

A "step" is one pass of the hot path over one batch of synthetic coordinates:
read_data (gather + COM) -> calc_value (all-pairs sweep, value + gradients [+ NCCL
all-reduce when sharded]) -> value() -> apply_force (scatter).  Workload (BASELINE.json):

    c5 (default)  coordNum 100 k x 1 M atoms, iso r0 = 4 A, n = 6, m = 12  (1e11 pairs; the case the
                  metric and the >= 50 % FP64-peak target are quoted on; the SAME problem is
                  sharded over N GPUs -> "scaling": "strong")
    c2            coordNum 1 k x 100 k (1e8 pairs)
    c3            selfCoordNum 50 k, tolerance 1e-3, pairListFrequency 10
    c4            coordNum 10 k x 500 k, anisotropic cutoff3 (3, 6, 4)

One JSON line is printed by rank 0.  `value` times device-resident inputs; `e2e` goes
through the host-buffer C-ABI step (H2D of the proxy positions and D2H of value +
gradients inside the timed region; sharded runs move 1/N of both per rank).  roofline:
FP64-compute bound: achieved = pair-evals/s x 39 algorithmic flop, peak =
DFMA microbenchmark run live here.  Positions move between timed steps (Gaussian displacements
of 0.05 A; the fitted solute of c4 also gets a new rigid rotation), outside the timed events.
`parity` (after the timed region, not timed): gradients of sampled group1 rows and group2
columns, value + all gradients of a 1024-row slab run through the same sharding, and the
forces of both routes, against the CPU oracle -- max-norm relative errors.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

if os.environ.get("LOCAL_RANK") is not None and os.environ.get("OMP_NUM_THREADS") == "1":
    # torchrun pins every rank to ONE OpenMP thread by default.  The host-buffer path (`e2e`)
    # scatters forces with OpenMP: give each rank its share of the host cores instead.  Has to
    # happen before any OpenMP runtime is loaded.
    _world = max(1, int(os.environ.get("WORLD_SIZE", "1")))
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or _world) // _world))

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from colvars_b200 import synth  # noqa: E402

WORKLOADS = {
    "c2": dict(kind="coordNum", n1=1000, n2=100000, cutoff3=(4.0, 4.0, 4.0), tol=0.0, freq=100,
               name="coordNum 1k x 100k iso r0=4 n=6 m=12"),
    "c3": dict(kind="selfCoordNum", n1=50000, n2=0, cutoff3=(4.0, 4.0, 4.0), tol=1e-3, freq=10,
               name="selfCoordNum 50k tolerance=1e-3 pairListFrequency=10"),
    "c4": dict(kind="coordNum", n1=10000, n2=500000, cutoff3=(3.0, 6.0, 4.0), tol=0.0, freq=100,
               fit=True,
               name="coordNum 10k x 500k aniso cutoff3=(3,6,4), group1 centerToReference + "
                    "rotateToReference (fitted on itself, fit gradients on)"),
    "c5": dict(kind="coordNum", n1=100000, n2=1000000, cutoff3=(4.0, 4.0, 4.0), tol=0.0,
               freq=100, name="coordNum 100k x 1M iso r0=4 n=6 m=12"),
    # widening: distanceInv on the same tile sweep
    "dinv": dict(kind="distanceInv", n1=10000, n2=500000, cutoff3=(4.0, 4.0, 4.0), tol=0.0,
                 freq=100, exponent=6, flop_per_pair=synth.FLOP_PER_PAIR_DISTANCEINV,
                 name="distanceInv 10k x 500k exponent=6"),
}
# the same pair sweep in a triclinic cell (GROMACS rhombic
# dodecahedron of the solvent cube's edge), image search of the 27 neighbours per pair
WORKLOADS["tric"] = dict(kind="coordNum", n1=10000, n2=100000, cutoff3=(4.0, 4.0, 4.0), tol=0.0,
                         freq=100, cell="dodecahedron",
                         name="coordNum 10k x 100k iso r0=4 n=6 m=12 in a rhombic-dodecahedron "
                              "(triclinic) cell")
for _k, _w in WORKLOADS.items():
    _w["key"] = _k
METRIC = "coordNum pair-evals/sec (value+grad)"
UNIT = "pair-evals/s"


def make_inputs(w):
    if w["kind"] == "selfCoordNum":
        proxy, i1, L = synth.selfcoordnum_case(w["n1"], seed=1234)
        return proxy, i1, None
    proxy, i1, i2, L = synth.coordnum_case(w["n1"], w["n2"], seed=1234)
    return proxy, i1, i2


def pairs_of(w):
    return synth.num_pairs(w["kind"], w["n1"], w["n2"])


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref/ref_driver) when it was built,
# else the plain-C port (oracle/liboracle.so); all host threads, bounded sample.
# ---------------------------------------------------------------------------------------------

def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# group1 rows of the slab the CPU arm is timed on (against all of group2), per workload: the
# same slab for `--impl reference` and for the in-arm `cpu_baseline`
CPU_SLAB_ROWS = {"c5": 960, "c2": 1000, "c4": 1920, "c3": 12000}


def cpu_run(w, proxy, i1, i2, budget_s=12.0, threads=None, one_core=False):
    """Time the CPU path on a slab of group1 rows against all of group2 and scale linearly.
    Returns dict(value=pairs/s, cores, kind, sample)."""
    threads = threads or host_cores()
    ref_driver = ROOT / "oracle" / "_ref" / "ref_driver"
    n1 = w["n1"]
    n2 = w["n2"] if w["kind"] != "selfCoordNum" else w["n1"]
    if ref_driver.exists():
        try:
            out = cpu_run_reference(w, proxy, i1, i2, budget_s, threads, ref_driver)
            if one_core:
                # also the reference's real behaviour, one component on
                # one thread, on 1/threads of the slab
                one = cpu_run_reference(w, proxy, i1, i2, budget_s, 1, ref_driver,
                                        rows_scale=1.0 / threads)
                out["value_1core"] = one["value"]
                out["sample_1core"] = one["sample"]
            return out
        except Exception as e:  # fall through to the port, say why
            sys.stderr.write(f"[bench] oracle/_ref/ref_driver failed ({e}); using the C port\n")
    from tests import oracle_lib as O
    p = O.make_params(w["cutoff3"], 6, 12, 0.0)
    pos2 = proxy[i2] if i2 is not None else proxy[i1]
    rows = min(n1, CPU_SLAB_ROWS.get(w.get("key"), 960))
    t0 = time.perf_counter()
    O.coordnum(p, proxy[i1][:rows], pos2, omp=threads)
    t = time.perf_counter() - t0
    return dict(value=rows * n2 / t, unit=UNIT, cores=threads, kind="port",
                sample=f"{rows} rows of group1 x all {n2} atoms of group2 "
                       f"({rows * n2:.3g} pair-evals, {t:.1f} s), OpenMP over rows, "
                       "scaled linearly")


def cpu_run_reference(w, proxy, i1, i2, budget_s, threads, ref_driver, rows_scale=1.0):
    """oracle/_ref/ref_driver: the unmodified reference library driven through
    colvarmodule::calc() with group1 split into `threads` coordNum components under
    `smp cvcs` (the reference's documented way to use several cores)."""
    n1 = w["n1"]
    n2 = w["n2"] if w["kind"] != "selfCoordNum" else w["n1"]
    # the driver runs 1 warm-up + 2 timed calc() steps on a fixed slab (about 0.5-1 s per calc
    # on 16 cores)
    rows = int(max(min(n1, CPU_SLAB_ROWS.get(w.get("key"), 960) * rows_scale), threads))
    with tempfile.TemporaryDirectory() as td:
        inp = Path(td) / "in.bin"
        if w["kind"] == "selfCoordNum":
            pos1, pos2 = proxy[i1][:rows], np.zeros((0, 3))
        else:
            pos1, pos2 = proxy[i1][:rows], proxy[i2]
        write_driver_input(inp, w, pos1, pos2)
        out = subprocess.run([str(ref_driver), "--in", str(inp), "--threads", str(threads),
                              "--steps", "2", "--time"], capture_output=True, text=True,
                             timeout=600, env=dict(os.environ, OMP_NUM_THREADS=str(threads)))
        if out.returncode != 0:
            raise RuntimeError(out.stderr[-500:])
        res = json.loads(out.stdout.strip().splitlines()[-1])
    npairs = rows * (rows - 1) // 2 if w["kind"] == "selfCoordNum" else rows * n2
    sec = res["seconds_per_step"]
    ncomp = int(res.get("components", 1))
    how = (f"group1 split into {ncomp} components under smp cvcs" if ncomp > 1 else
           "one component = one thread (the reference does not thread a single component)")
    return dict(value=npairs / sec, unit=UNIT, cores=ncomp, kind="reference",
                sample=f"unmodified Colvars library (oracle/_ref), colvarmodule::calc() on "
                       f"{rows} rows of group1 x {n2 if w['kind'] != 'selfCoordNum' else rows} "
                       f"atoms ({npairs:.3g} pair-evals, {sec:.2f} s/step), {how}, scaled "
                       "linearly")


def write_driver_input(path, w, pos1, pos2):
    """Binary case file of oracle/ref_driver.cpp."""
    kind = {"coordNum": 0, "selfCoordNum": 1, "groupCoord": 2}[w["kind"]]
    hdr = np.array([kind, len(pos1), len(pos2), 6, 12, w.get("freq", 100), 0, 0, 0, 0, 0, 0],
                   dtype=np.int32)
    dbl = np.array(list(w["cutoff3"]) + [w.get("tol", 0.0)], dtype=np.float64)
    with open(path, "wb") as f:
        f.write(hdr.tobytes())
        f.write(dbl.tobytes())
        f.write(np.ascontiguousarray(pos1, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(pos2, dtype=np.float64).tobytes())


# ---------------------------------------------------------------------------------------------

class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "20", "-i", str(self.index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path).read().splitlines():
            t = [x.strip() for x in line.split(",")]
            if len(t) < 9:
                continue
            try:
                sm.append(float(t[1]))
                mx.append(float(t[2]))
            except ValueError:
                continue
            for nm, v in zip(names, t[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)),
                       reasons=sorted(reasons), samples=len(sm),
                       sm_mhz_min=float(min(sm)))
        return out


def config_block(w, npairs, P, world, last_value, timing):
    return {"workload": w["name"], "pairs_per_step": npairs, "proxy_atoms": int(P),
            "sharding": (f"group1 blocks over {world} GPU(s), group2 replicated, ncclAllReduce "
                         "of [value|grad1|grad2]") if world > 1 else "none",
            "l2": "256 MiB buffer written between timed iterations (L2 flush)",
            "positions": "Gaussian displacements of 0.05 A between timed steps (outside the "
                         "timed events)",
            "timing": timing, "last_value": last_value}


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    proxy, i1, i2 = make_inputs(w)
    steps = max(args.steps, 1)
    vals = []
    info = None
    for k in range(args.warmup + steps):
        # every step times the same fixed slab as the b200 arm's cpu_baseline (CPU_SLAB_ROWS);
        # the last one also measures the one-core rate
        info = cpu_run(w, proxy, i1, i2, one_core=(k == args.warmup + steps - 1))
        if k >= args.warmup:
            vals.append(info["value"])
    value = float(np.mean(vals))
    npairs = pairs_of(w)
    P = proxy.shape[0]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * npairs / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # same key set as the b200 arm's config (the driver compares them)
        "config": config_block(w, npairs, P, 1, None,
                               timing="host wall clock inside oracle/_ref/ref_driver around "
                                      "colvarmodule::calc(), bounded slab scaled linearly to the "
                                      "full step (ms_per_step is that extrapolation)"),
        "cpu_baseline": dict(info, value=value),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


class Mover:
    """Moves the proxy positions on the device between timed steps: Gaussian
    displacements of sigma = 0.05 A per step for every atom; a fitted solute (c4) additionally
    gets a fresh small rigid rotation about its centre, so that the fit has new work every
    frame.  torch is plumbing here (random numbers + one small matmul), outside the timed events."""

    def __init__(self, torch, d_pos, i1, fitted, seed=4321, sigma=0.05):
        self.torch, self.d_pos, self.sigma, self.fitted = torch, d_pos, sigma, fitted
        self.gen = torch.Generator(device=d_pos.device)
        self.gen.manual_seed(seed)
        self.i1 = torch.tensor(np.asarray(i1, dtype=np.int64), device=d_pos.device)

    def step(self):
        t = self.torch
        self.d_pos += self.sigma * t.randn(self.d_pos.shape, generator=self.gen,
                                           dtype=self.d_pos.dtype, device=self.d_pos.device)
        if self.fitted:
            w = 0.02 * t.randn(3, generator=self.gen, dtype=self.d_pos.dtype,
                               device=self.d_pos.device)
            K = t.zeros((3, 3), dtype=self.d_pos.dtype, device=self.d_pos.device)
            K[0, 1], K[0, 2], K[1, 2] = -w[2], w[1], -w[0]
            K = K - K.T
            R = t.linalg.matrix_exp(K)
            x = self.d_pos[:, self.i1]
            c = x.mean(dim=1, keepdim=True)
            self.d_pos[:, self.i1] = R @ (x - c) + c


def quick_workload(cb, torch, key, device, fp64_tflops, steps=10, warmup=3):
    """Device-resident step of another BASELINE configuration: K timed steps, CUDA events, L2
    flushed and positions moved between steps.  Returns a small dict (pair-evals/s, ms/step,
    roofline fraction)."""
    w = WORKLOADS[key]
    proxy, i1, i2 = make_inputs(w)
    dev = torch.device("cuda", device)
    cv = cb.CoordNum(w["kind"], i1, i2, cutoff3=w["cutoff3"], tolerance=w["tol"],
                     pairlist_freq=w["freq"], device=device, exponent=w.get("exponent"))
    if w.get("cell") == "dodecahedron":
        L = synth.box_length(w["n2"])
        cv.set_boundaries((1, 1, 1), (L, 0, 0), (0, L, 0), (0.5 * L, 0.5 * L, 0.5 * 2 ** 0.5 * L))
    fitted = bool(w.get("fit"))
    if fitted:
        ref = proxy[i1].copy()
        cv.set_fitting(1, center=True, rotate=True, fitting_group=None, ref_pos=ref)
        ctr = ref.mean(0)
        proxy[i1] = (ref - ctr) @ synth.random_rotation(7).T + ctr + \
            np.random.default_rng(99).normal(0, 0.05, ref.shape)
    d_pos = torch.tensor(np.ascontiguousarray(proxy.T), device=dev)
    d_force = torch.zeros_like(d_pos)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    mover = Mover(torch, d_pos, i1, fitted)
    if w["tol"] > 0:
        # two full pair-list cycles, after a warm-up that holds the first rebuild (all-pairs
        # sweep: the list is not known to be sparse yet) and the first sparse rebuild
        # (one-time costs: lazy loading of the sort and list kernels, scratch allocation)
        steps = 2 * w["freq"]
        warmup = w["freq"] + 3
    ev, kev = [], []
    v = 0.0
    for s in range(warmup + steps):
        flush.fill_(s & 0xFF)
        mover.step()
        e0, k0, k1, e1 = (torch.cuda.Event(True) for _ in range(4))
        e0.record(stream)
        cv.read_data(d_pos, stream)
        k0.record(stream)
        cv.calc_value(step=s, gradients=True, stream=stream)
        k1.record(stream)
        if fitted:
            cv.calc_fit_gradients(stream)
        v = cv.value()
        cv.apply_force(-0.004 * (v - 0.1), d_force, stream)
        e1.record(stream)
        if s >= warmup:
            ev.append((e0, e1))
            kev.append((s, k0, k1))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    kms = [(s, a.elapsed_time(b)) for s, a, b in kev]
    npairs = pairs_of(w)
    out = {"workload": w["name"], "pairs_per_step": npairs, "steps": steps, "warmup": warmup,
           "value": npairs * len(ms) / (sum(ms) * 1e-3), "unit": UNIT,
           "ms_per_step": float(np.mean(ms)), "last_value": v,
           "positions": "moved by 0.05 A (Gaussian) before every step"}
    if w["tol"] > 0:
        reb = [m for s, m in kms if s % w["freq"] == 0]
        lst = [m for s, m in kms if s % w["freq"] != 0]
        st = cv.pairlist_stats()
        out["pairlist"] = {"rebuild_ms": float(np.mean(reb)), "list_ms": float(np.mean(lst)),
                           "listed_pairs": int(st["pairs"]),
                           "form": {0: "entries", 1: "entries in cell order",
                                    2: "neighbour rows"}.get(st["form"], "none"),
                           "row_entries": int(st["row_entries"]),
                           "extras": int(st["entries"])}
        kern = float(np.mean(reb))
    else:
        kern = float(np.mean([m for _, m in kms]))
    # the same step as one CUDA-graph launch (b200cv_step)
    if not (w["tol"] > 0):
        gev = []
        for s in range(warmup + steps):
            flush.fill_(s & 0xFF)
            mover.step()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(stream)
            cv.step(d_pos, step=s, gradients=True, stream=stream)
            v = cv.value()
            cv.apply_force(-0.004 * (v - 0.1), d_force, stream)
            e1.record(stream)
            if s >= warmup:
                gev.append((e0, e1))
        torch.cuda.synchronize()
        out["graph_step_ms"] = float(np.mean([a.elapsed_time(b) for a, b in gev]))
    out["calc_value_ms"] = kern
    flop = w.get("flop_per_pair", synth.FLOP_PER_PAIR_VALUE_GRAD)
    out["flop_per_pair"] = flop
    out["roofline_frac"] = npairs / (kern * 1e-3) * flop / 1e12 / fp64_tflops
    if fitted:
        out["validated"] = ("tests/test_gpu_baseline_sizes.py::test_c4_aniso_fitted_10k_x_500k: "
                            "geometry 1e-13 vs the oracle's fit, every pair 1e-12 on the device's "
                            "fitted positions, fit gradients / forces 1e-11")
    if w.get("cell"):
        # the 39-flop contract figure is the open-system pair; in a triclinic cell the reference
        # adds a 27-image search per pair (~600 flop as written, ~230 instructions here), so a
        # fraction by either count would mislead -- the absolute rate is the number
        out["roofline_frac"] = None
        out["note"] = ("pair-evals/s in a triclinic cell (image search per pair); open system on "
                       "the same kernel: see c4 / the headline")
    if w["tol"] > 0:
        # a pair list exists to NOT evaluate most pairs (rebuilds go through a cell grid, the
        # other steps walk the list): pair-evals/s counts the reference's pairs, and a flop
        # roofline over them would mean nothing
        out["roofline_frac"] = None
        out["note"] = ("value counts the reference's N(N-1)/2 pairs per step; rebuilds "
                       "evaluate grid candidates only, list steps the listed pairs")
    cv.close()
    return out


def ncu_traffic(key, world):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this
    workload (profiles/*_sweep_*ncu_raw.csv of the newest round), or (None, None)."""
    if key != "c5" or world != 1:
        return None, None
    import csv
    import glob
    for path in sorted(glob.glob(str(ROOT / "profiles" / "r0*_c5_sweep_*ncu_raw.csv")), reverse=True):
        try:
            # `ncu --page raw --csv`: one column per metric, a row of units, then one row per launch
            rows = list(csv.reader(open(path)))
            hdr, units = rows[0], rows[1]
            ki = hdr.index("Kernel Name")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = 0.0
            for r in rows[2:]:
                if len(r) <= ki or "sweep_kernel" not in r[ki]:
                    continue
                for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    c = hdr.index(name)
                    tot += float(r[c].replace(",", "")) * scale.get(units[c], 1.0)
                break
            if tot > 0:
                return tot, str(Path(path).relative_to(ROOT))
        except Exception:
            continue
    return None, None


def parity_block(cb, torch, dist, w, proxy, i1, i2, cv, dev, world, rank, local_rank, k_force):
    """After the timed region (not timed): the sharded run against the CPU oracle on the
    workload's initial positions.  (1) gradients of 256 group1 rows spread over all ranks'
    blocks and of 64 group2 columns of the benchmarked component, and the forces both routes
    scatter for them; (2) a 1024-row slab of group1 against all of group2 as a second component
    with the same sharding: value, all gradients of both groups.  Max-norm relative errors.
    Collective: every rank takes part, rank 0 runs the oracle and returns the dict."""
    from tests import oracle_lib as O
    if w["kind"] != "coordNum" or w["tol"] > 0 or w.get("fit"):
        return None
    n1, n2 = len(i1), len(i2)
    nthr = host_cores()
    p = O.make_params(w["cutoff3"], 6, 12, 0.0)

    def rel(a, b):
        return float(np.max(np.abs(np.asarray(a) - b)) / max(np.max(np.abs(b)), 1e-300))

    d_pos = torch.tensor(np.ascontiguousarray(proxy.T), device=dev)
    d_force = torch.zeros_like(d_pos)
    cv.read_data(d_pos)
    cv.calc_value(step=1, gradients=True)
    v = cv.value()
    fa = k_force * (v - 0.1)
    cv.apply_force(fa, d_force)
    torch.cuda.synchronize()
    # host-buffer route: the ranks' force arrays add up to the full force
    h_pos = torch.tensor(proxy).pin_memory()
    h_force = torch.zeros_like(h_pos).pin_memory()
    vh = cv.calc_host(h_pos, step=1)
    cv.apply_force_host(k_force * (vh - 0.1), h_force)
    hf = h_force.to(dev)
    if world > 1:
        dist.all_reduce(hf)
    # the slab as a component of its own, sharded like the benchmarked one
    rows_slab = np.unique(np.linspace(0, n1 - 1, 1024).astype(np.int64))
    slab = cb.CoordNum(w["kind"], i1[rows_slab], i2, cutoff3=w["cutoff3"], device=local_rank)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(cb.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        slab.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    slab.read_data(d_pos)
    slab.calc_value(step=1, gradients=True)
    sv = slab.value()
    out = None
    if rank == 0:
        rows = np.unique(np.linspace(0, n1 - 1, 256).astype(np.int64))
        cols = np.unique(np.linspace(0, n2 - 1, 64).astype(np.int64))
        pos1, pos2 = proxy[i1], proxy[i2]
        _, og1, _ = O.coordnum(p, pos1[rows], pos2, omp=nthr)
        _, _, og2 = O.coordnum(p, pos1, pos2[cols], omp=nthr)
        g1, g2 = cv.gradients(1), cv.gradients(2)
        f = d_force.cpu().numpy().T
        fh = hf.cpu().numpy()
        errs = {
            "grad1_rows": rel(g1[rows], og1), "grad2_cols": rel(g2[cols], og2),
            "force_rows": rel(f[i1[rows]], fa * og1), "force_cols": rel(f[i2[cols]], fa * og2),
            "e2e_force_rows": rel(fh[i1[rows]], k_force * (vh - 0.1) * og1),
            "e2e_force_cols": rel(fh[i2[cols]], k_force * (vh - 0.1) * og2),
            "e2e_value_vs_device": abs(vh - v) / abs(v),
        }
        sov, sog1, sog2 = O.coordnum(p, pos1[rows_slab], pos2, omp=nthr)
        errs["slab_value"] = abs(sv - sov) / abs(sov)
        errs["slab_grad1"] = rel(slab.gradients(1), sog1)
        errs["slab_grad2"] = rel(slab.gradients(2), sog2)
        out = {"max_rel": max(errs.values()), "tolerance": 1e-12, "errors": errs,
               "rows": int(len(rows)), "cols": int(len(cols)), "slab_rows": int(len(rows_slab)),
               "oracle": "oracle/liboracle.so (OpenMP row-parallel form, scalar summed in "
                         "extended precision), max|a-b| / max|b| per array",
               "positions": "the workload's initial frame"}
    else:
        slab.gradients(1)  # same calls on every rank (they synchronise the stream only)
    slab.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5",
                    choices=sorted(k for k in WORKLOADS if k not in ("dinv", "tric")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-others", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]

    if args.impl == "reference":
        return run_reference_arm(args, w)

    import torch
    import torch.distributed as dist
    import colvars_b200 as cb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the b200 arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    proxy, i1, i2 = make_inputs(w)
    P = proxy.shape[0]
    npairs = pairs_of(w)
    cv = cb.CoordNum(w["kind"], i1, i2, cutoff3=w["cutoff3"], tolerance=w["tol"],
                     pairlist_freq=w["freq"], device=local_rank)
    fitted = bool(w.get("fit"))
    if fitted:
        # refPositions = the solute as generated; the frame that is evaluated is that solute
        # rigidly rotated about its centre plus 0.05 A noise, so the fit has real work to undo
        ref = proxy[i1].copy()
        cv.set_fitting(1, center=True, rotate=True, fitting_group=None, ref_pos=ref,
                       enable_fit_gradients=True)
        rng = np.random.default_rng(99)
        ctr = ref.mean(0)
        proxy[i1] = (ref - ctr) @ synth.random_rotation(7).T + ctr + rng.normal(0, 0.05, ref.shape)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(cb.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        cv.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)

    d_pos = torch.tensor(np.ascontiguousarray(proxy.T), device=dev)  # SoA planes, stride P
    d_force = torch.zeros_like(d_pos)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2
    stream = torch.cuda.current_stream()
    mover = Mover(torch, d_pos, i1, fitted)  # same seed on every rank: identical frames
    k_force = -0.004

    def step(s):
        mover.step()
        cv.read_data(d_pos, stream)
        cv.calc_value(step=s, gradients=True, stream=stream)
        if fitted:
            cv.calc_fit_gradients(stream)
        v = cv.value()
        cv.apply_force(k_force * (v - 0.1), d_force, stream)
        return v

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # FP64 roofline denominator, measured live (sustained, ~250 ms of pure DFMA)
    fp64_tflops, _ = cb.fp64_peak(local_rank, 250.0)

    # clocks are sampled from before the warm-up (nvidia-smi needs ~0.1 s to deliver its first
    # sample) through the end of the timed region; only samples under load are kept
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    for s in range(args.warmup):
        step(s)
    barrier()
    ev = [(torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True),
           torch.cuda.Event(True)) for _ in range(args.steps)]
    launches0 = cb.kernel_launches()  # counted by the library itself, one per kernel launch
    value = 0.0
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.fill_(s & 0xFF)  # evict L2 between timed iterations (outside the events)
        mover.step()           # new frame (outside the events)
        e0, ek0, ek1, e1 = ev[s]
        e0.record(stream)
        cv.read_data(d_pos, stream)
        ek0.record(stream)
        cv.calc_value(step=args.warmup + s, gradients=True, stream=stream)
        ek1.record(stream)
        if fitted:
            cv.calc_fit_gradients(stream)
        value = cv.value()
        cv.apply_force(k_force * (value - 0.1), d_force, stream)
        e1.record(stream)
    launches = cb.kernel_launches() - launches0
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else {}
    step_ms = [a.elapsed_time(d) for a, _, _, d in ev]
    kern_ms = [b.elapsed_time(c) for _, b, c, _ in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    kern_tot = torch.tensor([sum(kern_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_tot, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    kern_ms_avg = float(kern_tot.item()) / args.steps
    # rebuild / list-walk split for pair-list workloads
    split = None
    if w["tol"] > 0:
        reb = [m for s, m in enumerate(kern_ms) if (args.warmup + s) % w["freq"] == 0]
        lst = [m for s, m in enumerate(kern_ms) if (args.warmup + s) % w["freq"] != 0]
        split = {"rebuild_ms": float(np.mean(reb)) if reb else None,
                 "list_ms": float(np.mean(lst)) if lst else None}

    # ---- end-to-end through the host-buffer C-ABI step -----------------------------------
    e2e = None
    if not args.no_e2e:
        h_pos = torch.tensor(proxy).pin_memory()
        h_force = torch.zeros_like(h_pos).pin_memory()
        hf = h_force.numpy()
        for s in range(2):
            v = cv.calc_host(h_pos, step=s)
            cv.apply_force_host(k_force * (v - 0.1), h_force)
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            v = cv.calc_host(h_pos, step=args.warmup + s)
            cv.apply_force_host(k_force * (v - 0.1), h_force)
        barrier()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        # bytes each rank moves per step: its block of the AoS position array up (the other
        # blocks arrive GPU to GPU over NVLink), its share of every group's gradients down
        per = (P + world - 1) // world
        share = sum((n * (rank + 1)) // world - (n * rank) // world
                    for n in (len(i1), len(i2) if i2 is not None else 0))
        e2e = {"value": npairs * args.steps / float(t_e2e.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(min(per, P) * 24), "d2h_bytes_per_step": int(share * 24 + 8),
               "bytes_are": "per rank" if world > 1 else "total",
               "ms_per_step": 1e3 * float(t_e2e.item()) / args.steps,
               "timing": "host wall clock around calc_host + apply_force_host (each call "
                         "synchronises), max over ranks"}
        if world > 1:
            e2e["exchange"] = ("each rank uploads 1/N of the position array, ncclAllGather over "
                               "NVLink; forces: rank r scatters its 1/N share of every group")
        assert hf is not None

    parity = None
    if not args.no_parity:
        try:
            parity = parity_block(cb, torch, dist, w, proxy, i1, i2, cv, dev, world, rank,
                                  local_rank, k_force)
        except Exception as e:  # never lose the bench line to the checker
            parity = {"error": str(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    pairs_per_s = npairs * args.steps / (total_ms * 1e-3)
    roof_note = None
    if split and not split["rebuild_ms"]:
        roof_note = ("no pair-list rebuild fell inside the timed region: list-walk steps do not "
                     "evaluate every pair, the FP64 roofline does not apply to them")
    if split and split["rebuild_ms"]:
        # pair-list workloads: only rebuild steps evaluate every pair; the roofline is quoted on
        # those, `value` stays the whole-job rate (list-walk steps count all pairs they stand for)
        kern_ms_avg = split["rebuild_ms"]
        roof_note = "roofline quoted on pair-list REBUILD steps only (full sweep)"
    kern_pairs_per_s = npairs / (kern_ms_avg * 1e-3)
    achieved_tflops = kern_pairs_per_s * synth.FLOP_PER_PAIR_VALUE_GRAD / 1e12
    if split and not split["rebuild_ms"]:
        achieved_tflops = float("nan")
    peak = fp64_tflops * world
    traffic, traffic_src = ncu_traffic(args.workload, world)
    line = {
        "metric": METRIC, "value": pairs_per_s, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_block(w, npairs, P, world, value,
                               timing="CUDA events on the launching stream per step, summed; "
                                      "max over ranks"),
        "roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved_tflops / peak,
                     # DRAM bytes per sweep launch, read from the committed ncu capture named in
                     # traffic_source (null when there is none for this workload / rank count)
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": "sweep_kernel (+ exact queue with the final sum) = calc_value",
                     "kernel_ms": kern_ms_avg,
                     "flop_per_pair": synth.FLOP_PER_PAIR_VALUE_GRAD,
                     "peak_source": "DFMA microbenchmark run live in this process "
                                    "(b200cv_fp64_peak, 250 ms sustained) x n_gpus; "
                                    "MEASURED_PEAKS.json has no FP64 entry"},
        # for the driver: what MEASURED_PEAKS.json lacks (dense FP64 FMA, one GPU, TFLOP/s)
        "measured_peak_fp64": {"tflops": fp64_tflops, "how": "b200cv_fp64_peak: 16 dependent DFMA "
                               "chains per thread, 8 CTAs x 256 threads per SM, 250 ms"},
        "clocks": clocks,
        "gpu_launches": int(launches),
        "wall_s": t_wall,
    }
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        hbm_peak, which = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, which = 6650.0, "fallback (B200_PROFILING.md)"
    if line["roofline"]["traffic"]:
        gbs = line["roofline"]["traffic"] / (kern_ms_avg * 1e-3) / 1e9
        # why the bound is not "hbm": DRAM traffic of the dominant kernel against the copy peak
        line["roofline"]["hbm_check"] = {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                         "frac": gbs / hbm_peak, "peak_source": which}
    if split:
        line["pairlist"] = split
        line["roofline"]["note"] = roof_note
    if e2e:
        line["e2e"] = e2e
    if parity is not None:
        line["parity"] = parity
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_run(w, proxy, i1, i2, one_core=True)
        except Exception as e:
            line["cpu_baseline"] = {"error": str(e)[:200]}
    if world == 1 and args.workload == "c5" and not args.no_others:
        # the other single-GPU configurations of BASELINE.json, device-resident step only
        # (same timing rules; not the headline): BASELINE config 2 is the 1-GPU case
        others = {}
        for key in ("c2", "c4", "c3", "dinv", "tric"):
            try:
                others[key] = quick_workload(cb, torch, key, local_rank, fp64_tflops)
            except Exception as e:
                others[key] = {"error": str(e)[:200]}
        line["other_workloads"] = others
    if line["roofline"]["achieved"] != line["roofline"]["achieved"]:  # NaN -> null
        line["roofline"]["achieved"] = line["roofline"]["frac"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
