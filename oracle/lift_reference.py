"""oracle/lift_reference.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Executes the reference's OWN enrichment functions without copying them: the
reference script cannot be imported (it imports AMUSE at module top,
al26_nbody.py:19-24), so the FunctionDef nodes are lifted from the reference file by
AST at run time and exec'd with only {numba, numpy} in scope.

  calc_wind_abs        al26_nbody.py:642-702   (numba njit, parallel)
  calc_eta_disk_sne    al26_nbody.py:1291-1334
  calc_intersection    al26_nbody.py:1156-1190 (interloper path, out of scope; cheap pin)

Run as a script in the BUILD container (where /root/reference exists) to regenerate
the committed fixtures:

    python oracle/lift_reference.py            # writes tests/golden/enrich_golden.npz

/root/reference does not exist on the GPU box; nothing at test/bench time on the box
calls `lift()`.  bench.py's CPU arm uses oracle/enrich_oracle.py (kind "port") there.
"""
import ast
import os
import sys

import numpy as np

REFERENCE_FILE = "/root/reference/al26_nbody.py"
WANTED = ("calc_wind_abs", "calc_eta_disk_sne", "calc_intersection")


def available():
    return os.path.exists(REFERENCE_FILE)


def lift(names=WANTED, path=REFERENCE_FILE):
    """Return {name: function} lifted from the reference file."""
    import numba as nb
    with open(path) as fh:
        tree = ast.parse(fh.read(), filename=path)
    picked = [node for node in tree.body if isinstance(node, ast.FunctionDef) and node.name in names]
    mod = ast.Module(body=picked, type_ignores=[])
    ns = {"nb": nb, "np": np}
    exec(compile(mod, path, "exec"), ns)
    return {n: ns[n] for n in names if n in ns}


PLOT_FILE = "/root/reference/plotting/al26_plot.py"


def lift_plotting(names=("local_densities_numba",), path=PLOT_FILE):
    """Lift numba functions of the reference's post-processing script (plotting/al26_plot.py:324-359)."""
    from numba import njit, prange
    with open(path) as fh:
        tree = ast.parse(fh.read(), filename=path)
    picked = [node for node in tree.body if isinstance(node, ast.FunctionDef) and node.name in names]
    ns = {"np": np, "njit": njit, "prange": prange}
    exec(compile(ast.Module(body=picked, type_ignores=[]), path, "exec"), ns)
    return {n: ns[n] for n in names if n in ns}


def make_wind_case(rng, n, n_hm, frac_lm=0.5, length_km=3.0e13, bubble_km=3.0856775814913e12):
    """A seeded synthetic input set for calc_wind_abs in the reference's units."""
    x, y, z = (rng.normal(0.0, length_km, n) for _ in range(3))
    vx, vy, vz = (rng.normal(0.0, 1.0, n) for _ in range(3))
    ids = rng.permutation(n)
    hm_id = np.sort(ids[:n_hm]).astype(np.int64)
    n_lm = int(frac_lm * n)
    lm_id = np.sort(ids[n_hm:n_hm + n_lm]).astype(np.int64)
    mdot = np.zeros(n)
    mdot[hm_id] = 10.0 ** rng.uniform(14.0, 17.0, n_hm)  # kg/s
    wr26 = np.zeros(n)
    wr60 = np.zeros(n)
    wr26[hm_id] = 10.0 ** rng.uniform(-7.0, -4.0, n_hm)
    wr60[hm_id] = 10.0 ** rng.uniform(-9.0, -6.0, n_hm)
    rdisk = np.full(n, 100.0 * 149597870.691)  # 100 au in km
    # a few discs right next to sources so the local model has hits
    for k, hm in enumerate(hm_id):
        near = lm_id[(k * 7) % len(lm_id)] if len(lm_id) else None
        if near is not None:
            x[near], y[near], z[near] = x[hm] + 0.3 * bubble_km, y[hm] - 0.2 * bubble_km, z[hm] + 0.1 * bubble_km
    return dict(x=x, y=y, z=z, vx=vx, vy=vy, vz=vz, mdot=mdot, wr26=wr26, wr60=wr60, rdisk=rdisk,
                lm_id=lm_id, hm_id=hm_id, bubble_km=np.float64(bubble_km))


def generate(out_path):
    fns = lift()
    cwa = fns["calc_wind_abs"]
    out = {}
    # SURVEY 8(c) golden vector
    x = np.array([0.0, 1e12, 5e12]); y = np.zeros(3); z = np.zeros(3)
    vx = np.array([0.0, 3.0, 3.0]); vy = np.array([0.0, 4.0, 4.0]); vz = np.zeros(3)
    mdot = np.array([1e15, 0.0, 0.0]); wr = np.array([1e-4, 0.0, 0.0])
    rd = np.array([0.0, 1.5e10, 1.5e10])
    lm = np.array([1, 2], dtype=np.int64); hm = np.array([0], dtype=np.int64)
    out["tiny_global"] = cwa(lm, hm, x, y, z, vx, vy, vz, mdot, wr, rd, 0.0, 3e12, 1e11)
    out["tiny_local"] = cwa(lm, hm, x, y, z, vx, vy, vz, mdot, wr, rd, 3e12, 3e12, 1e11)
    out["eta_sne_100_206264p806"] = np.float64(fns["calc_eta_disk_sne"](100.0, 206264.806))
    out["intersection"] = np.float64(fns["calc_intersection"](-1, 0, 0, 1, 0, 0, 0, 0.05, 0, 0, 0.05, 0, 0.1))
    # seeded cases
    dt_s = 0.01 * 1.0e6 * 365.242199 * 86400.0
    for tag, (seed, n, n_hm) in {"a": (0, 257, 3), "b": (1, 2048, 19), "c": (2, 5000, 64)}.items():
        c = make_wind_case(np.random.default_rng(seed), n, n_hm)
        rvir = 2.0 * float(c["bubble_km"]) * 10.0
        for k, v in c.items():
            out[f"{tag}_{k}"] = v
        out[f"{tag}_dt_s"] = np.float64(dt_s)
        out[f"{tag}_rvir_km"] = np.float64(rvir)
        args = (c["lm_id"], c["hm_id"], c["x"], c["y"], c["z"], c["vx"], c["vy"], c["vz"], c["mdot"])
        out[f"{tag}_g26"] = cwa(*args, c["wr26"], c["rdisk"], 0.0, rvir, dt_s)
        out[f"{tag}_g60"] = cwa(*args, c["wr60"], c["rdisk"], 0.0, rvir, dt_s)
        out[f"{tag}_l26"] = cwa(*args, c["wr26"], c["rdisk"], float(c["bubble_km"]), float(c["bubble_km"]), dt_s)
        out[f"{tag}_l60"] = cwa(*args, c["wr60"], c["rdisk"], float(c["bubble_km"]), float(c["bubble_km"]), dt_s)
    # calc_intersection (interloper path): seeded straight-line pairs, some grazing the 0.1 pc sphere
    rng = np.random.default_rng(9)
    m = 400
    a_old = rng.normal(0, 1.0, (3,)); a_new = a_old + rng.normal(0, 0.3, (3,))
    b_old = a_old[:, None] + rng.normal(0, 0.15, (3, m)); b_new = b_old + rng.normal(0, 0.2, (3, m))
    b_new[:, :20] = b_old[:, :20]  # discs that do not move
    fr = np.array([fns["calc_intersection"](a_old[0], a_old[1], a_old[2], a_new[0], a_new[1], a_new[2],
                                            b_old[0, i], b_old[1, i], b_old[2, i], b_new[0, i], b_new[1, i], b_new[2, i], 0.1)
                   for i in range(m)])
    out["isect_a_old"], out["isect_a_new"], out["isect_b_old"], out["isect_b_new"], out["isect_frac"] = a_old, a_new, b_old, b_new, fr
    # local densities (plotting/al26_plot.py:324-359): 10-nearest-neighbour density of every star
    ld = lift_plotting()["local_densities_numba"]
    rng = np.random.default_rng(21)
    for tag, nn in (("ld_a", 64), ("ld_b", 1500)):
        px, py, pz = (rng.normal(0, 1.0, nn) * (1.0 + 3.0 * (rng.random(nn) < 0.2)) for _ in range(3))
        pm = 10.0 ** rng.uniform(-2, 2, nn)
        out[tag + "_x"], out[tag + "_y"], out[tag + "_z"], out[tag + "_m"] = px, py, pz, pm
        out[tag + "_rho"] = ld(px, py, pz, pm)
    np.savez_compressed(out_path, **out)
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "..", "tests", "golden", "enrich_golden.npz")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    if not available():
        sys.exit("reference file not present; fixtures can only be regenerated in the build container")
    res = generate(dst)
    print("wrote", os.path.normpath(dst), "with", len(res), "arrays")
    print("tiny_global", res["tiny_global"], "tiny_local", res["tiny_local"])
    print("eta_sne", repr(float(res["eta_sne_100_206264p806"])), "intersection", float(res["intersection"]))
