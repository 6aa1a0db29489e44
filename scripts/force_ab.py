"""A/B aid: force-kernel rate by block size and one evolve at N=1e5 for the library named by AL26_LIB (default: the product)."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("26al-nbody_b200")
ctx = pkg.Context(0)
n = int(os.environ.get("AB_N", "100000"))
c = pkg.ic.cluster(n, seed=0)
for v in [int(a) for a in os.environ.get("AB_VARIANTS", "0").split(",")]:
    ctx.set_force_variant(v)
    g = pkg.GravityCore(ctx=ctx)
    g.set_time(0.0)
    g.commit(*[c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")])
    row = []
    for na, reps in ((0, 5), (20000, 10), (5000, 20), (1000, 40), (300, 40), (64, 40)):
        ms, pairs = g.bench_force(reps, n_act=na)
        row.append(f"{na or n}:{pairs/ms*1e-6:.1f}G")
    steps, pairs = g.evolve(2.0 ** -5)
    ms, _ = g.last_device_ms()
    print(f"{os.environ.get('AL26_LIB', 'product')} variant {v}: " + " ".join(row) + f" | evolve 2^-5: {steps} steps {ms:.1f} ms {pairs/ms*1e-6:.1f} Gpairs/s", flush=True)
