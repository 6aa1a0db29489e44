"""Enrichment disc kernel in its three modes on BASELINE config 5 (1e6 discs; 1000 / 190 / 16 massive stars):
kernel time per outer step (CUDA events around the two launches), achieved HBM rate on SURVEY 8(d)'s 210 B per
disc-update, and the largest relative difference of every inventory row against mode 0.  Development aid + the
numbers quoted in DESIGN.md; writes one JSON line per (sources, mode)."""
import importlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("26al-nbody_b200")
ctx = pkg.Context(0)
n_disc = 1_000_000
pc_km = 3.08567758128e13
rng = np.random.default_rng(5)
for n_hm in (1000, 190, 16):
    n = n_disc + n_hm
    mass = np.full(n, 1.0)
    hm = np.arange(0, n, n // n_hm)[:n_hm]
    mass[hm] = 20.0
    wr26 = np.zeros(n); wr60 = np.zeros(n); sn = np.zeros(n)
    wr26[hm], wr60[hm], sn[hm] = 1e-5, 1e-7, 1e26
    mdot = np.zeros(n); mdot[hm] = 1e16
    mdot[hm[1]] = 0.0  # one supernova in the first step
    pv = np.concatenate([rng.normal(0, 1.0 * pc_km, (3, n)), rng.normal(0, 1.0, (3, n))])
    rd = np.full(n, 1.49597870691e10)
    tau = rng.exponential(2.885, n)
    f26, f60 = pkg.decay_fractions(0.01)
    dt_s = 0.01 * 1e6 * 365.242199 * 86400
    ref = None
    for mode in (0, 1, 2):
        e = pkg.EnrichCore(ctx=ctx)
        e.set_mode(mode)
        e.commit(rd, tau, np.ones(n), np.zeros(n), wr26, wr60, sn, sn)
        ker = []
        for k in range(6):
            e.step(mass, mdot, pv, dt_s, 0.01 * (k + 1), 0.1 * pc_km, 2.0 * pc_km, f26, f60)
            ker.append(e.last_kernel_ms())
        inv = e.get()[0]
        if ref is None:
            ref = inv
        diff = {}
        for r, name in enumerate(pkg.ROWS):
            nz = ref[r] != 0
            same_support = bool(np.array_equal(inv[r] != 0, nz))
            diff[name] = (float(np.max(np.abs(inv[r][nz] / ref[r][nz] - 1.0))) if nz.any() else 0.0) if same_support else None
        ms = float(np.median(ker[2:]))
        print(json.dumps({"sources": n_hm, "mode": mode, "kernel_ms": ms, "first_ms": ker[0],
                          "disc_updates_per_s": n / (ms * 1e-3), "pairs_per_s": n_hm * float(n_disc) / (ms * 1e-3),
                          "hbm_gbs_210B": n * 210.0 / (ms * 1e-3) / 1e9, "frac_hbm_6547.8": n * 210.0 / (ms * 1e-3) / 1e9 / 6547.8,
                          "local_hits": int(np.count_nonzero(inv[0])), "max_rel_diff_vs_mode0": diff,
                          "table_cta_us": {k: round(v / 1965.0, 2) for k, v in e.profile().items()}}), flush=True)
        e.set_mode(0)
ctx.close()
