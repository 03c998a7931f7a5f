#!/usr/bin/env python3
"""Generate golden vectors from the UNMODIFIED reference (oracle/_ref/ref_driver, built by
oracle/Makefile from /root/reference) for inputs the reference's own regression tests do not
cover: periodic cells, group1CenterOnly, unequal masses, larger groups,
general exponents and the CONTENTS of the pair list.  Run in the build container:

    make -C oracle ref && python tests/golden/make_ref_vectors.py

Writes tests/golden/ref_vectors/<case>.npz (inputs + reference outputs, a few 100 kB in all);
tests/test_oracle_ref_vectors.py pins the C oracle to them on any machine.
"""
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
OUT = HERE / "ref_vectors"
DRIVER = ROOT / "oracle" / "_ref" / "ref_driver"

KIND = {"coordNum": 0, "selfCoordNum": 1, "groupCoord": 2}


def write_case(path, c, frames):
    flags = (1 if c.get("g1c") else 0) | (2 if c.get("g2c") else 0)
    if c.get("cell") is not None:
        flags |= 4
    if c.get("masses") is not None:
        flags |= 8
    if c.get("fit"):
        flags |= 16 | (32 if c["fit"]["center"] else 0) | (64 if c["fit"]["rotate"] else 0)
    nfit = c.get("nfit", 0)
    hdr = np.array([KIND[c["kind"]], c["n1"], c["n2"], c["en"], c["ed"], c["freq"], flags,
                    len(frames), nfit, 0, 0, 0], dtype=np.int32)
    dbl = np.array(list(c["cutoff3"]) + [c["tol"]], dtype=np.float64)
    with open(path, "wb") as f:
        f.write(hdr.tobytes())
        f.write(dbl.tobytes())
        if flags & 4:
            f.write(np.array(list(c["periodic"]) + [0], dtype=np.int32).tobytes())
            f.write(np.asarray(c["cell"], dtype=np.float64).tobytes())
        if flags & 8:
            f.write(np.asarray(c["masses"], dtype=np.float64).tobytes())
        if flags & 16:
            f.write(np.ascontiguousarray(c["fit"]["ref"], dtype=np.float64).tobytes())
        for fr in frames:
            f.write(np.ascontiguousarray(fr, dtype=np.float64).tobytes())


def read_result(path, c, nframes):
    n1, n2 = c["n1"], c["n2"]
    nfit = c.get("nfit", 0)
    P = n1 + n2 + nfit
    nref = (nfit or n1) if c.get("fit") else 0
    raw = open(path, "rb").read()
    off = 0
    out = []
    for _ in range(nframes):
        v, fa = np.frombuffer(raw, np.float64, 2, off)
        off += 16
        g1 = np.frombuffer(raw, np.float64, 3 * n1, off).reshape(n1, 3)
        off += 24 * n1
        g2 = np.frombuffer(raw, np.float64, 3 * n2, off).reshape(n2, 3)
        off += 24 * n2
        fo = np.frombuffer(raw, np.float64, 3 * P, off).reshape(P, 3)
        off += 24 * P
        npl = int(np.frombuffer(raw, np.int64, 1, off)[0])
        off += 8
        idx = np.frombuffer(raw, np.uint64, max(npl, 0), off)
        off += 8 * max(npl, 0)
        rec = dict(value=v, fa=fa, grad1=g1.copy(), grad2=g2.copy(), forces=fo.copy(),
                   npl=npl, pl=idx.copy())
        if nref:
            rec["q"] = np.frombuffer(raw, np.float64, 4, off).copy()
            off += 32
            rec["fitgrad"] = np.frombuffer(raw, np.float64, 3 * nref, off).reshape(nref, 3).copy()
            off += 24 * nref
            rec["fitted1"] = np.frombuffer(raw, np.float64, 3 * n1, off).reshape(n1, 3).copy()
            off += 24 * n1
        out.append(rec)
    return out


def run_reference(c, frames, threads=1):
    with tempfile.TemporaryDirectory() as td:
        inp, outp = Path(td) / "in.bin", Path(td) / "out.bin"
        write_case(inp, c, frames)
        r = subprocess.run([str(DRIVER), "--in", str(inp), "--out", str(outp), "--threads",
                            str(threads)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stdout[-2000:] + r.stderr[-2000:])
        return read_result(outp, c, len(frames))


def rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    q0, q1, q2, q3 = q
    return np.array([
        [q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * (q1 * q2 - q0 * q3), 2 * (q0 * q2 + q1 * q3)],
        [2 * (q0 * q3 + q1 * q2), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * (q2 * q3 - q0 * q1)],
        [2 * (q1 * q3 - q0 * q2), 2 * (q0 * q1 + q2 * q3), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3]])


def cases():
    rng = np.random.default_rng(20261019)

    def pos(n, L):
        return rng.random((n, 3)) * L

    base = dict(en=6, ed=12, freq=100, tol=0.0, cutoff3=(4.0, 4.0, 4.0), cell=None, masses=None)
    out = {}
    out["coordnum_iso_96x300"] = dict(base, kind="coordNum", n1=96, n2=300, L=14.0)
    out["coordnum_aniso_exp8_10"] = dict(base, kind="coordNum", n1=70, n2=260, L=13.0,
                                         cutoff3=(3.0, 5.5, 4.2), en=8, ed=10)
    out["coordnum_pairlist_freq2"] = dict(base, kind="coordNum", n1=80, n2=240, L=16.0, tol=0.02,
                                          freq=2, nframes=5)
    out["selfcoordnum_pairlist_freq3"] = dict(base, kind="selfCoordNum", n1=220, n2=0, L=15.0,
                                              tol=0.005, freq=3, nframes=5, cutoff3=(3.5,) * 3)
    out["selfcoordnum_iso_257"] = dict(base, kind="selfCoordNum", n1=257, n2=0, L=12.0)
    out["coordnum_pbc_ortho"] = dict(base, kind="coordNum", n1=60, n2=250, L=30.0,
                                     periodic=(1, 1, 1),
                                     cell=[11.0, 0, 0, 0, 12.5, 0, 0, 0, 10.0])
    out["coordnum_pbc_triclinic"] = dict(base, kind="coordNum", n1=50, n2=200, L=30.0,
                                         periodic=(1, 1, 1),
                                         cell=[12.0, 0, 0, 3.5, 11.0, 0, 1.0, -2.5, 10.0])
    out["selfcoordnum_pbc_mixed"] = dict(base, kind="selfCoordNum", n1=150, n2=0, L=25.0,
                                         periodic=(1, 0, 1),
                                         cell=[10.0, 0, 0, 0, 10.0, 0, 0, 0, 9.0])
    out["coordnum_group1centeronly_masses"] = dict(base, kind="coordNum", n1=40, n2=160, L=7.0,
                                                   g1c=True, masses=True)
    out["coordnum_group2centeronly_pairlist"] = dict(base, kind="coordNum", n1=120, n2=30, L=9.0,
                                                     g2c=True, masses=True, tol=0.05, freq=2,
                                                     nframes=4)
    out["groupcoord_masses_aniso"] = dict(base, kind="groupCoord", n1=33, n2=47, L=5.0,
                                          cutoff3=(3.0, 6.0, 4.0), masses=True)
    # fitting (config 4): group1 centred / rotated onto reference positions
    out["fit_self_center_rotate_aniso"] = dict(base, kind="coordNum", n1=60, n2=300, L=14.0,
                                               cutoff3=(3.0, 6.0, 4.0), nframes=3,
                                               fit=dict(center=True, rotate=True))
    out["fit_group_center_rotate"] = dict(base, kind="coordNum", n1=50, n2=250, L=13.0, nfit=40,
                                          nframes=3, fit=dict(center=True, rotate=True))
    out["fit_group_center_only"] = dict(base, kind="coordNum", n1=45, n2=200, L=12.0, nfit=30,
                                        fit=dict(center=True, rotate=False))
    out["fit_self_rotate_only_selfcoordnum"] = dict(base, kind="selfCoordNum", n1=120, n2=0,
                                                    L=11.0, cutoff3=(3.0, 5.0, 4.0),
                                                    fit=dict(center=False, rotate=True))
    for name, c in out.items():
        P = c["n1"] + c["n2"] + c.get("nfit", 0)
        L = c.pop("L")
        nf = c.pop("nframes", 1)
        f0 = pos(P, L)
        frames = [f0]
        for _ in range(nf - 1):
            frames.append(frames[-1] + rng.normal(0, 0.08, f0.shape))
        if c["masses"] is True:
            c["masses"] = rng.random(P) * 15.0 + 1.0
        if c.get("fit"):
            n1, n2, nfit = c["n1"], c["n2"], c.get("nfit", 0)
            fitrows = slice(n1 + n2, P) if nfit else slice(0, n1)
            # reference structure = frame 0 of the fitted atoms, moved and rotated elsewhere;
            # later frames: rigid rotation + noise of those atoms (and of group1)
            Rm = rot(rng)
            c["fit"]["ref"] = (f0[fitrows] - f0[fitrows].mean(0)) @ Rm.T + np.array([3.0, -2.0, 5.0])
            new = []
            for fr in frames:
                fr = fr.copy()
                Rk = rot(rng)
                ctr = fr[fitrows].mean(0)
                for rows in ({(0, n1), (fitrows.start, fitrows.stop)}):
                    a, b = rows
                    fr[a:b] = (fr[a:b] - ctr) @ Rk.T + ctr + rng.normal(0, 0.05, (b - a, 3))
                new.append(fr)
            frames = new
        yield name, c, frames


def main():
    if not DRIVER.exists():
        sys.exit("build the reference first: make -C oracle ref")
    OUT.mkdir(exist_ok=True)
    for name, c, frames in cases():
        res = run_reference(c, frames)
        arrs = dict(kind=c["kind"], n1=c["n1"], n2=c["n2"], en=c["en"], ed=c["ed"], freq=c["freq"],
                    tol=c["tol"], cutoff3=np.array(c["cutoff3"]), g1c=bool(c.get("g1c")),
                    g2c=bool(c.get("g2c")), frames=np.array(frames))
        if c["cell"] is not None:
            arrs["cell"] = np.array(c["cell"])
            arrs["periodic"] = np.array(c["periodic"])
        if c["masses"] is not None:
            arrs["masses"] = np.array(c["masses"])
        if c.get("fit"):
            arrs["nfit"] = c.get("nfit", 0)
            arrs["fit_center"] = c["fit"]["center"]
            arrs["fit_rotate"] = c["fit"]["rotate"]
            arrs["fit_ref"] = np.array(c["fit"]["ref"])
        for k, r in enumerate(res):
            arrs[f"value_{k}"] = r["value"]
            arrs[f"fa_{k}"] = r["fa"]
            arrs[f"grad1_{k}"] = r["grad1"]
            arrs[f"grad2_{k}"] = r["grad2"]
            arrs[f"forces_{k}"] = r["forces"]
            arrs[f"npl_{k}"] = r["npl"]
            arrs[f"pl_{k}"] = r["pl"]
            if "q" in r:
                arrs[f"q_{k}"] = r["q"]
                arrs[f"fitgrad_{k}"] = r["fitgrad"]
                arrs[f"fitted1_{k}"] = r["fitted1"]
        np.savez_compressed(OUT / f"{name}.npz", **arrs)
        print(f"{name}: {len(frames)} frame(s), value[0] = {res[0]['value']:.15g}, "
              f"npl[0] = {res[0]['npl']}")


if __name__ == "__main__":
    main()
