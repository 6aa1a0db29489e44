"""A few enrichment steps on BASELINE config 5's discs for ncu captures: AL26_SOURCES massive stars, mode AL26_MODE."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("26al-nbody_b200")
ctx = pkg.Context(0)
n_hm, mode = int(os.environ.get("AL26_SOURCES", "16")), int(os.environ.get("AL26_MODE", "0"))
n = 1_000_000 + n_hm
pc_km = 3.08567758128e13
rng = np.random.default_rng(5)
mass = np.full(n, 1.0); hm = np.arange(0, n, n // n_hm)[:n_hm]; mass[hm] = 20.0
wr26 = np.zeros(n); wr60 = np.zeros(n); sn = np.zeros(n)
wr26[hm], wr60[hm], sn[hm] = 1e-5, 1e-7, 1e26
mdot = np.zeros(n); mdot[hm] = 1e16
pv = np.concatenate([rng.normal(0, pc_km, (3, n)), rng.normal(0, 1.0, (3, n))])
e = pkg.EnrichCore(ctx=ctx); e.set_mode(mode)
e.commit(np.full(n, 1.49597870691e10), rng.exponential(2.885, n), np.ones(n), np.zeros(n), wr26, wr60, sn, sn)
f26, f60 = pkg.decay_fractions(0.01)
for k in range(4):
    e.step(mass, mdot, pv, 0.01 * 1e6 * 365.242199 * 86400, 0.01 * (k + 1), 0.1 * pc_km, 2.0 * pc_km, f26, f60)
    print(k, e.last_kernel_ms())
