#!/usr/bin/env python3
"""Extra functional cases for the host shim, in the format of the reference's regression
tests: configurations the reference's own coordnum tests do not exercise (dummy-atom group2,
fitted group1, general exponents, group1CenterOnly), run through the UNMODIFIED reference
(oracle/_ref/run_colvars_test, built by oracle/Makefile) on the reference's trajectory.xyz to
produce the golden files.  Run in the build container; outputs are committed.

    make -C oracle ref && python tests/golden/make_extra_regression.py
"""
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
BASE = HERE / "colvars_regression"
OUT = HERE / "colvars_regression_extra"
RUNNER = ROOT / "oracle" / "_ref" / "run_colvars_test"

HEAD = """colvarsTrajFrequency 1
colvarsRestartFrequency 10
indexFile index.ndx

colvar {
    name one
    outputAppliedForce on
    width 0.5
"""
TAIL = """}

harmonic {
    colvars        one
    centers        0.1
    forceConstant  0.001
}
"""

CASES = {
    "coordnum-dummy": """    coordNum {
        cutoff 6.0
        group1 { indexGroup group1 }
        group2 { dummyAtom (1.0, -2.0, 3.0) }
    }
""",
    "coordnum-dummy-atoms": """    coordNum {
        debugGradients on
        cutoff3 (5.0, 7.0, 6.0)
        group1CenterOnly off
        group1 { indexGroup group1 }
        group2 { dummyAtom (1.0, -2.0, 3.0) }
    }
""",
    "coordnum-group1centeronly-exp": """    coordNum {
        debugGradients on
        expNumer 8
        expDenom 10
        group1CenterOnly on
        group1 { indexGroup group1 }
        group2 { indexGroup group2 }
    }
""",
    "coordnum-fitted": """    coordNum {
        debugGradients on
        cutoff3 (3.0, 6.0, 4.0)
        group1 {
            indexGroup group1
            centerToReference yes
            rotateToReference yes
            fittingGroup { indexGroup group2 }
            refPositions (1.0, 0.0, 0.5) (2.5, 1.0, 0.0) (0.0, 2.0, 1.0) (1.5, 3.0, 2.5)
        }
        group2 { indexGroup group2 }
    }
""",
    # colvar::alpha_angles: six hydrogen-bond terms O(i)-N(i+4) --
    # the batch of single pairs -- next to eight Calpha angle terms, from the index groups
    # prot_CA / prot_N / prot_O of the reference's index.ndx
    "alpha": """    alpha {
        prefix prot_
        hBondCoeff 0.5
    }
""",
    "alpha-hbonds-only": """    alpha {
        prefix prot_
        hBondCoeff 1.0
        hBondCutoff 3.5
        hBondExpNumer 4
        hBondExpDenom 10
    }
""",
    "selfcoordnum-aniso-exp": """    selfCoordNum {
        debugGradients on
        cutoff3 (3.0, 6.0, 4.0)
        expNumer 4
        expDenom 8
        group1 { indexGroup group1 }
    }
""",
}


def main():
    if not RUNNER.exists():
        raise SystemExit("build the reference first: make -C oracle ref")
    OUT.mkdir(exist_ok=True)
    for f in ("trajectory.xyz", "index.ndx"):
        shutil.copy(BASE / f, OUT / f)
    for name, body in CASES.items():
        d = OUT / f"{name}_harmonic-fixed"
        (d / "AutoDiff").mkdir(parents=True, exist_ok=True)
        (d / "test.in").write_text(HEAD + body + TAIL)
        r = subprocess.run([str(RUNNER), "-c", f"{d.name}/test.in", "-t", "trajectory.xyz", "-o",
                            f"{d.name}/AutoDiff/test_out", "--force"], cwd=OUT,
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(f"{name}: reference failed\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
        for junk in list((d / "AutoDiff").glob("*.state")) + list((d / "AutoDiff").glob("*.BAK")):
            junk.unlink()
        print(name, (d / "AutoDiff" / "test_out.colvars.traj").read_text().splitlines()[1])


if __name__ == "__main__":
    main()
