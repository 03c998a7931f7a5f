#!/usr/bin/env python3
"""Copy the reference's own regression fixtures for the coordination-number path
into tests/golden/colvars_regression/ (run once in the build container, where
/root/reference exists; the copies are committed so the GPU box needs nothing else).

Source: /root/reference/tests/input_files/<case>_harmonic-fixed/{test.in,AutoDiff/*}
plus trajectory.xyz and index.ndx.  These are DATA fixtures
(inputs + golden outputs), not reference source code.
"""
import shutil
from pathlib import Path

REF = Path("/root/reference/tests/input_files")
OUT = Path(__file__).resolve().parent / "colvars_regression"
CASES = [
    "coordnum", "coordnum-pairlist", "coordnum-aniso", "coordnum-aniso-pairlist",
    "coordnum-group2centeronly", "selfcoordnum", "selfcoordnum-pairlist",
    "groupcoord", "groupcoord-aniso",
    # widening: same pair-sweep skeleton, different pair function
    "distanceinv",
    # the single-pair user of the same pair function (colvar::h_bond)
    "hbond",
]


# cases whose directory is not <name>_harmonic-fixed: distancePairs,
# a vector-valued component under a linear bias
OTHER_DIRS = ["distancepairs_linear-distancepairs"]


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    for name in OTHER_DIRS:
        d = OUT / name
        (d / "AutoDiff").mkdir(parents=True, exist_ok=True)
        shutil.copy(REF / name / "test.in", d / "test.in")
        for g in sorted((REF / name / "AutoDiff").iterdir()):
            if g.name.endswith(".colvars.traj") or "_forces_" in g.name:
                shutil.copy(g, d / "AutoDiff" / g.name)
    for f in ("trajectory.xyz", "index.ndx"):
        shutil.copy(REF / f, OUT / f)
    for c in CASES:
        d = OUT / f"{c}_harmonic-fixed"
        (d / "AutoDiff").mkdir(parents=True, exist_ok=True)
        shutil.copy(REF / f"{c}_harmonic-fixed" / "test.in", d / "test.in")
        for g in sorted((REF / f"{c}_harmonic-fixed" / "AutoDiff").iterdir()):
            if g.name.endswith(".colvars.traj") or "_forces_" in g.name:
                shutil.copy(g, d / "AutoDiff" / g.name)
    print("collected", len(CASES), "cases into", OUT)


if __name__ == "__main__":
    main()
