"""GPU: `number_of_workers=k` from ONE process (the reference's `ph4(converter, number_of_workers=workers)`,
al26_nbody.py:57,1711-1720) -- an al26_group of k GPUs, one host thread per GPU inside the library -- against the
same calls on one GPU and against the oracle.  Skipped on a box with a single GPU."""
import numpy as np
import pytest

from oracle import enrich_oracle as eo
from oracle import hermite as H

pytestmark = pytest.mark.gpu


def _workers(pkg):
    k = min(pkg.device_count(), 8)
    if k < 2:
        pytest.skip("needs at least 2 GPUs")
    return k if k in (2, 4, 8) else 2


def test_number_of_workers_gravity_surface(pkg):
    k = _workers(pkg)
    U = pkg.units
    n = 4099  # not a multiple of the GPU count: ownership is i % k
    c = pkg.ic.cluster(n, seed=11)
    cv = U.nbody_to_si(1.0 | U.pc, float(c["m_msun"].sum()) | U.MSun)
    cl = pkg.Particles(n)
    cl.key = np.arange(n, dtype=np.uint64)
    cl.mass = c["m_msun"] | U.MSun
    for a in ("x", "y", "z"):
        setattr(cl, a, cv.length_to_si(c[a]))
    for a in ("vx", "vy", "vz"):
        setattr(cl, a, cv.speed_to_si(c[a]))
    out = {}
    for workers in (1, k):
        g = pkg.B200Gravity(cv, number_of_workers=workers)
        g.particles.add_particles(cl)
        e0 = g.kinetic_energy.value_in(U.J) + g.potential_energy.value_in(U.J)
        rv0 = g.virial_radius().value_in(U.pc)
        t1 = cv.time_to_si(0.02)
        g.evolve_model(t1)
        work1 = (g.last_steps, g.last_pairs)
        g.particles.mass = cl.mass * 0.99            # the per-step mass channel (:874)
        g.evolve_model(cv.time_to_si(0.03))
        x = np.stack([g.particles.x.value_in(U.pc), g.particles.y.value_in(U.pc), g.particles.z.value_in(U.pc)])
        e1 = g.kinetic_energy.value_in(U.J) + g.potential_energy.value_in(U.J)
        out[workers] = dict(work=[work1, (g.last_steps, g.last_pairs)], x=x, e0=e0, e1=e1, rv0=rv0,
                            t=g.model_time.value_in(U.Myr), m=g.particles.mass.value_in(U.MSun))
        g.stop()
    a, b = out[1], out[k]
    assert a["work"] == b["work"]                                  # integer work: block steps and pair counts
    assert b["e0"] == pytest.approx(a["e0"], rel=1e-12) and b["rv0"] == pytest.approx(a["rv0"], rel=1e-12)
    assert np.max(np.abs(a["x"] - b["x"])) < 1e-9 and b["e1"] == pytest.approx(a["e1"], rel=1e-9)
    assert a["t"] == b["t"] and np.array_equal(a["m"], b["m"])
    # and against the oracle
    o = H.HermiteOracle(n); o.commit(*[c[q] for q in ("m", "x", "y", "z", "vx", "vy", "vz")])
    assert o.evolve(0.02) == b["work"][0]


def test_group_enrichment_is_bit_equal_to_one_gpu(pkg):
    k = _workers(pkg)
    n = 8191
    c = pkg.ic.cluster(n, seed=4)
    mass = c["m_msun"]
    hm = np.nonzero(mass >= 13.0)[0]
    wr = np.zeros(n); wr[hm] = 1e-5
    sn = np.zeros(n); sn[hm] = 1e26
    mdot = np.zeros(n); mdot[hm] = 1e16; mdot[hm[0]] = 0.0
    alive = (mass >= 0.1) & (mass <= 3.0)
    rd = np.full(n, 1.49597870691e10)
    rng = np.random.default_rng(1)
    pv = np.concatenate([rng.normal(0, 3e13, (3, n)), rng.normal(0, 1, (3, n))])
    f26, f60 = pkg.decay_fractions(0.01)
    st = eo.EnrichState(rd, c["tau_disk_myr"], alive, np.zeros(n, bool), wr, wr, sn, sn)
    grp = pkg.Group(k)
    try:
        e = pkg.EnrichCore(ctx=grp)
        e.commit(rd, c["tau_disk_myr"], alive, np.zeros(n), wr, wr, sn, sn)
        for step in range(1, 4):
            ev = e.step(mass, mdot, pv, 3.15e11, 0.01 * step, 3.0857e12, 6e13, f26, f60)
            ev_o = eo.enrich_step(st, mass, mdot, *pv, 3.15e11, 0.01 * step, 3.0857e12, 6e13, f26, f60)
            assert ev.tolist() == ev_o
        inv, fin, al, kk = e.get()
        assert np.array_equal(inv, st.inv) and np.array_equal(fin, st.fin)
        assert np.array_equal(al, st.disk_alive) and np.array_equal(kk, st.kicked)
    finally:
        grp.close()
