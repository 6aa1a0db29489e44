"""dE/E at 1 Myr (the second half of BASELINE.json's metric), gravity only, outer step 0.01 Myr as in the
reference loop (al26_nbody.py:786,833): E = K + U, dE = (E0 - E)/E exactly as plotting/al26_plot.py:297-299.

    python scripts/energy_drift.py --n 1000 --t-myr 1.0 [--cpu]     # --cpu: also run the CPU oracle on the same ICs
    python scripts/energy_drift.py --n 100000 --t-myr 0.05 --cpu    # N = 1e5: a bounded horizon the CPU can afford (~2.4e12 pairs)
Prints one JSON line per run (GPU first)."""
import argparse, importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000)
ap.add_argument("--t-myr", type=float, default=1.0)
ap.add_argument("--dt-myr", type=float, default=0.01)
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--model", default="plummer")
ap.add_argument("--cpu", action="store_true")
ap.add_argument("--no-gpu", action="store_true")
args = ap.parse_args()

pkg = importlib.import_module("26al-nbody_b200")
U = pkg.units
c = pkg.ic.cluster(args.n, seed=args.seed, model=args.model)
cv = U.nbody_to_si(1.0 | U.pc, float(c["m_msun"].sum()) | U.MSun)
p = [c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")]
nsteps = int(round(args.t_myr / args.dt_myr))
times = [cv.time_to_nbody((k + 1) * args.dt_myr | U.Myr) for k in range(nsteps)]


def run(name, g, energies):
    k0, u0, _ = energies()
    t0 = time.perf_counter()
    steps = pairs = 0
    trace = []
    for k, t in enumerate(times):
        a, b = g.evolve(t)
        steps += a; pairs += b
        if (k + 1) % max(1, nsteps // 10) == 0:
            k1, u1, _ = energies()
            trace.append(((k + 1) * args.dt_myr, ((k0 + u0) - (k1 + u1)) / (k1 + u1)))
    wall = time.perf_counter() - t0
    k1, u1, _ = energies()
    print(json.dumps({"impl": name, "n": args.n, "model": args.model, "t_myr": args.t_myr, "outer_dt_myr": args.dt_myr,
                      "t_nbody": times[-1], "dE_over_E": ((k0 + u0) - (k1 + u1)) / (k1 + u1), "E0": k0 + u0,
                      "virial_ratio_end": k1 / abs(u1), "block_steps": steps, "pairs": pairs, "wall_s": wall,
                      "pairs_per_s": pairs / wall, "trace": trace}), flush=True)


if not args.no_gpu:
    g = pkg.GravityCore()
    g.commit(*p)
    run("b200", g, g.energies)
if args.cpu:
    from oracle import hermite as H
    H.prefer_native()
    H.use_all_cores()
    o = H.HermiteOracle(args.n)
    o.commit(*p)
    run("cpu-oracle x%d" % H.num_threads(), o, o.energies)
