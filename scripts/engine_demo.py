"""Small-N evolve for profiling the cluster engine (ncu launch list / ncu --set full -k regex:k_engine):
    python scripts/engine_demo.py [N] [span]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("26al-nbody_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
span = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0 ** -7
ctx = pkg.Context(0)
c = pkg.ic.cluster(n, seed=0)
g = pkg.GravityCore(ctx=ctx)
g.set_time(0.0)
g.commit(*[c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")])
g.evolve(2.0 ** -9)
steps, pairs = g.evolve(2.0 ** -9 + span)
ms, nl = g.last_device_ms()
ne, cs = ctx.engine_steps()
print(f"N={n}: {steps} block steps, {pairs:.3e} pairs, {ms:.3f} ms, {ms * 1e3 / steps:.2f} us per block step, "
      f"{nl} launches; engine took {ne} steps since commit (cluster of {cs})")
ctx.close()
