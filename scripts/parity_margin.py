"""How much margin do the gravity parity tolerances have, per pair form?  Run once per library (AL26_LIB):
 (1) one force evaluation vs the long-double oracle: max relative acc / jerk error, Plummer and fractal ICs;
 (2) tests/test_gpu_gravity.py::test_initialize_and_stepwise_parity's end-state error (60 block steps, then the
     synchronisation step) for every step mode -- the quantity the test bounds by 1e-10.
Prints numbers only; asserts nothing.  Round 2 used it to compare two builds of the force work item (the shipped pair
arithmetic and a rejected "split jerk" form, profiles/r02_kforce_exploration/README.md); the experimental build is no
longer in the sources, the script stays as the margin report for whatever library AL26_LIB names."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("26al-nbody_b200")
from oracle import hermite as H

def vec_rel(a, b):
    a, b = np.stack(a), np.stack(b)
    return np.max(np.linalg.norm(a - b, axis=0) / np.linalg.norm(b, axis=0))

def make(n, seed, model):
    c = pkg.ic.cluster(n, seed=seed, model=model, require_massive=False)
    return [c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")]

lib = os.path.basename(os.environ.get("AL26_LIB", "product"))
ctx = pkg.Context(0)
for n, model in ((1000, "plummer"), (512, "fractal"), (4096, "fractal")):
    p = make(n, n + 1, model)
    g = pkg.GravityCore(ctx=ctx); g.set_time(0.0)
    a = g.force(*p); o = H.force(*p)
    print(f"{lib} force {model} N={n}: acc err {vec_rel(a[:3], o[:3]):.3e} jerk err {vec_rel(a[3:6], o[3:6]):.3e}", flush=True)
for (mode, fuse), name in zip([(1, 32), (1, 0), (0, 32), (2, 32), (3, 32)], ["loop-fused", "loop", "graph", "engine", "chip"]):
    for n, model in ((256, "plummer"), (1000, "plummer"), (512, "fractal")):
        ctx.set_step_mode(mode); ctx.set_fuse_max(fuse)
        g = pkg.GravityCore(ctx=ctx); g.set_time(0.0)
        p = make(n, n + 1, model)
        o = H.HermiteOracle(n); o.commit(*p); g.commit(*p)
        o.initialize(); g.initialize()
        o.begin(0.02); g.begin(0.02)
        same = True
        for step in range(60):
            oi, ot = o.get_active()
            nd_o, fin_o = o.advance(1); nd_g, fin_g = g.advance(1)
            if fin_o: break
            same = same and np.array_equal(g.get_last_active(), oi) and np.array_equal(g.get_timesteps()[1], o.get_timesteps()[1])
        o.advance(-1); g.advance(-1); o.finish(); g.finish()
        gs, os_ = g.get_state(), o.get_state()
        print(f"{lib} {name} {model} N={n}: integer work identical {same}; end state pos err {vec_rel(gs[1:4], os_[1:4]):.3e} vel err {vec_rel(gs[4:7], os_[4:7]):.3e}", flush=True)
ctx.set_step_mode(-1); ctx.set_fuse_max(-1)
