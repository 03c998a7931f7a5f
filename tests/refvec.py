"""Loader/runner for tests/golden/ref_vectors/*.npz: outputs of the unmodified reference
(oracle/_ref/ref_driver, see tests/golden/make_ref_vectors.py) on synthetic inputs."""
from pathlib import Path

import numpy as np

DIR = Path(__file__).resolve().parent / "golden" / "ref_vectors"
NAMES = sorted(p.stem for p in DIR.glob("*.npz"))


def load(name):
    z = np.load(DIR / f"{name}.npz", allow_pickle=False)
    c = {k: z[k] for k in z.files}
    c["kind"] = str(c["kind"])
    for k in ("n1", "n2", "en", "ed", "freq"):
        c[k] = int(c[k])
    c["tol"] = float(c["tol"])
    c["g1c"], c["g2c"] = bool(c["g1c"]), bool(c["g2c"])
    if c["kind"] == "groupCoord":  # hard-coded (src/colvarcomp_coordnums.cpp:78-82)
        c["g1c"] = c["g2c"] = True
    c["nframes"] = c["frames"].shape[0]
    c["fitted"] = "fit_ref" in c
    c["nfit"] = int(c["nfit"]) if c["fitted"] else 0
    return c


def boundaries(c):
    if "cell" not in c:
        return None
    cell = c["cell"].reshape(3, 3)
    return tuple(int(v) for v in c["periodic"]), cell[0], cell[1], cell[2]


def harmonic_fa(v):
    # ref_driver's bias: harmonic centers 0.1 forceConstant 0.001, colvar width 0.5
    return -0.001 / 0.25 * (v - 0.1)


def run_oracle(c):
    """Run the C oracle over the frames of a vector case -> list of dicts like the file."""
    from . import oracle_lib as O
    n1, n2 = c["n1"], c["n2"]
    p = O.make_params(tuple(c["cutoff3"]), c["en"], c["ed"], c["tol"], boundaries(c))
    masses = c.get("masses")
    m1 = None if masses is None else masses[:n1]
    m2 = None if masses is None else masses[n1:]
    self_ = c["kind"] == "selfCoordNum"
    if self_:
        npairs = n1 * (n1 - 1) // 2
    else:
        npairs = (1 if c["g1c"] else n1) * (1 if c["g2c"] else n2)
    use_pl = c["tol"] > 0 and c["kind"] != "groupCoord"
    pl = np.ones(npairs, dtype=np.uint8) if use_pl else None
    out = []
    nfit = c["nfit"]
    fit = None
    if c["fitted"]:
        fit = O.Fit(c["fit_ref"], bool(c["fit_center"]), bool(c["fit_rotate"]))
    for step in range(c["nframes"]):
        pos = c["frames"][step]
        rebuild = use_pl and step % c["freq"] == 0
        p1 = pos[:n1]
        extra = {}
        if fit is not None:
            # cvc::read_data -> calc_apply_roto_translation (src/colvaratoms.cpp:1127-1173)
            p1, unrot, _ = fit.apply(pos[:n1], pos[n1 + n2:] if nfit else None)
        if self_:
            v, g1 = O.selfcoordnum(p, p1, pairlist=pl, rebuild=rebuild)
            g2 = np.zeros((0, 3))
        else:
            v, g1, g2 = O.coordnum(p, p1, pos[n1:n1 + n2], pairlist=pl, rebuild=rebuild,
                                   g1_center_only=c["g1c"], g2_center_only=c["g2c"], mass1=m1,
                                   mass2=m2)
        fa = harmonic_fa(v)
        if fit is None:
            forces = np.concatenate([fa * g1, fa * g2])
        else:
            # apply_colvar_force (src/colvaratoms.cpp:1665-1703): R^-1 grad on the main
            # group, fit gradients on the group used for the fit
            nf = nfit if nfit else n1
            fg = fit.fit_gradients(g1, unrot, nf)
            f1 = fa * (g1 @ fit.R if fit.fs.rotate else g1)   # rows: R^T g
            forces = np.concatenate([f1, fa * g2, np.zeros((nfit, 3))])
            if nfit:
                forces[n1 + n2:] += fa * fg
            else:
                forces[:n1] += fa * fg
            extra = dict(q=fit.q, fitgrad=fg, fitted1=p1)
        out.append(dict(value=v, fa=fa, grad1=g1, grad2=g2, forces=forces,
                        pl=None if pl is None else np.flatnonzero(pl).astype(np.uint64), **extra))
    return out
