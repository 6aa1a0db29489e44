"""GPU: the reference-facing surface end to end -- B200Gravity (AMUSE-style members), channels, the
outer-loop re-host and the device enrichment -- against the same loop driven by the CPU oracles."""
import numpy as np
import pytest

from oracle import enrich_oracle as eo
from oracle import hermite as H

pytestmark = pytest.mark.gpu


def test_b200gravity_surface(pkg):
    U = pkg.units
    cl, cv = pkg.driver.init_cluster("plummer", 400, 1.0 | U.pc, seed=5)
    g = pkg.B200Gravity(cv, number_of_workers=8)
    try:
        g.particles.add_particles(cl)
        assert len(g.particles) == 400 and np.array_equal(g.particles.key, cl.key)
        assert g.model_time.value_in(U.Myr) == 0.0
        assert np.allclose(g.particles.x.value_in(U.pc), cl.x.value_in(U.pc), rtol=1e-14)
        assert g.particles[3].key == cl[3].key and g.particles[3].x.value_in(U.km) == pytest.approx(cl[3].x.value_in(U.km), rel=1e-14)
        e0 = g.kinetic_energy + g.potential_energy
        assert e0.value_in(U.J) < 0.0
        rv = g.virial_radius()
        assert rv.value_in(U.pc) == pytest.approx(cl.virial_radius().value_in(U.pc), rel=1e-11)  # device vs host O(N^2)
        snap = g.particles.copy()
        g.evolve_model(0.01 | U.Myr)
        assert g.model_time.value_in(U.Myr) == pytest.approx(0.01, rel=1e-14)
        assert not np.array_equal(snap.x.value_in(U.km), g.particles.x.value_in(U.km))  # the copy is detached
        e1 = g.kinetic_energy + g.potential_energy
        assert abs((e1 - e0).value_in(U.J) / e0.value_in(U.J)) < 1e-7
        # mass channel in, full channel out
        st = pkg.StellarStub()
        st.particles.add_particles(cl)
        st.evolve_model(0.01 | U.Myr)
        st.particles.new_channel_to(g.particles).copy_attributes(["mass"])
        g.particles.new_channel_to(cl).copy()
        assert np.allclose(cl.mass.value_in(U.MSun), st.particles.mass.value_in(U.MSun), rtol=1e-14)
        g.model_time = 3.0 | U.Myr  # the setter used after a checkpoint reload (:1736)
        g.evolve_model(3.01 | U.Myr)
        assert g.model_time.value_in(U.Myr) == pytest.approx(3.01, rel=1e-14)
        assert g.parameters.timestep_parameter == 0.14 and g.parameters.epsilon_squared.value_in(U.m ** 2) == 0.0
    finally:
        g.stop()


@pytest.mark.parametrize("model,n", [("plummer", 600), ("fractal", 500)])
def test_outer_loop_matches_oracle_loop(pkg, model, n):
    U = pkg.units
    steps = 12
    st_kw = dict(lifetime_factor=0.004)  # massive stars die within the test: SN events at known steps
    cluster, gravity, stellar, enrich, hist = pkg.driver.run(nstars=n, t_f=10.0 | U.Myr, model=model, seed=3,
                                                              max_outer_steps=steps, stellar=pkg.StellarStub(**st_kw))
    try:
        # the same loop with the CPU oracles
        cl, cv = pkg.driver.init_cluster(model, n, 1.0 | U.pc, seed=3, stellar=pkg.StellarStub(**st_kw))
        o = H.HermiteOracle(n)
        o.commit(cv.mass_to_nbody(cl.mass), *[cv.length_to_nbody(getattr(cl, a)) for a in "xyz"],
                 *[cv.speed_to_nbody(getattr(cl, a)) for a in ("vx", "vy", "vz")])
        sref = pkg.StellarStub(**st_kw)
        sref.particles.add_particles(cl)
        est = eo.EnrichState(cl.r_disk.value_in(U.km), cl.tau_disk.value_in(U.Myr), cl.disk_alive, cl.kicked,
                             cl.wind_ratio_26al, cl.wind_ratio_60fe, cl.sn_yield_26al.value_in(U.kg),
                             cl.sn_yield_60fe.value_in(U.kg))
        mass_class = np.array(cl.mass.value_in(U.MSun))
        dt = (10.0 | U.Myr) / 1000
        t = 0.0 | U.s
        events = []
        for k in range(steps):
            _, _, s = o.energies()
            mt = float(np.sum(o.get_state()[0]))
            rvir_km = cv.length_to_si(mt * mt / (2.0 * s)).value_in(U.km)
            t = t + dt
            o.evolve(cv.time_to_nbody(t))
            sref.evolve_model(t)
            o.set_mass(cv.mass_to_nbody(sref.particles.mass))
            stt = o.get_state()
            pos = [stt[i] * cv.km_per_length for i in (1, 2, 3)]
            vel = [stt[i] * cv.kms_per_speed for i in (4, 5, 6)]
            mdot = -np.asarray(sref.particles.wind_mass_loss_rate.value_in(U.kg / U.s))
            f26, f60 = eo.decay_fractions(float(dt.value_in(U.Myr)))
            ev = eo.enrich_step(est, mass_class, mdot, *pos, *vel, float(dt.value_in(U.s)), float(t.value_in(U.Myr)),
                                float((0.1 | U.pc).value_in(U.km)), float(rvir_km), f26, f60)
            events.append(ev)
            mass_class = np.array(sref.particles.mass.value_in(U.MSun))
            assert hist[k]["virial_radius_pc"] == pytest.approx(rvir_km / (1.0 | U.pc).value_in(U.km), rel=1e-9)
        assert [h["sn_events"] for h in hist] == events          # SN event lists: bit-exact
        assert sum(len(e) for e in events) >= 1
        inv, fin, alive, kicked = enrich.get()
        assert np.array_equal(alive, est.disk_alive) and np.array_equal(kicked, est.kicked)  # flags: bit-exact
        scale = np.max(np.abs(est.inv), axis=1, keepdims=True) + 1e-300
        assert np.max(np.abs(inv - est.inv) / scale) < 1e-10      # per-disc masses: 1e-10 relative
        assert np.max(np.abs(fin - est.fin) / scale) < 1e-10
        dead = [i for e in events for i in e]
        if np.any(cl.sn_yield_26al.value_in(U.kg)[dead] > 0):   # stars above 25 Msun have SN yield 0 (:460-461)
            assert np.any(inv[pkg.ROW["sne26"]] > 0)
        xs = gravity.particles.x.value_in(U.km)
        assert np.max(np.abs(xs - o.get_state()[1] * cv.km_per_length)) / np.max(np.abs(xs)) < 1e-9
        # cluster columns were pulled from the device
        assert np.array_equal(cluster.mass_26al_global.value_in(U.kg), inv[pkg.ROW["global26"]])
        assert set(hist[0]["timings"]) >= {"grav", "stel", "discs", "step"}
    finally:
        gravity.stop()


def test_checkpoint_resume_continues_the_run(pkg, tmp_path):
    """`-r base -nc k` (al26_nbody.py:1641-1656, :1734-1737): a run resumed from its second checkpoint goes on as the
    uninterrupted run did.  The reference's checkpoint holds no forces or timesteps -- the resumed worker starts cold --
    so the uninterrupted run is driven with the re-initialisation policy that does the same at every outer step (1);
    what is left between the two is one ulp in the unit round trip of the checkpointed positions."""
    U = pkg.units
    base = str(tmp_path / "run")
    kw = dict(nstars=600, t_f=10.0 | U.Myr, model="plummer", seed=3, reinit_policy=1)
    stub = lambda: pkg.StellarStub(lifetime_factor=0.004)
    cl_a, grav_a, _, enr_a, hist_a = pkg.driver.run(max_outer_steps=14, stellar=stub(), checkpoint_base=base, **kw)
    try:
        inv_a, fin_a, alive_a, kicked_a = enr_a.get()
        x_a = np.stack([grav_a.particles.x.value_in(U.pc), grav_a.particles.vx.value_in(U.kms)])
        t_a = grav_a.model_time.value_in(U.Myr)
    finally:
        grav_a.stop()
    assert pkg.checkpoint.most_recent_checkpoint(base) == 2          # written after outer steps 1 and 11, plus #0 at start
    cl_b, grav_b, _, enr_b, hist_b = pkg.driver.run(max_outer_steps=3, stellar=stub(), reload=base, n_checkpoint=2, **kw)
    try:
        assert hist_b[0]["t_new_myr"] == pytest.approx(0.12, rel=1e-12)  # the clocks were set from the metadata (:1736)
        assert [h["sn_events"] for h in hist_b] == [h["sn_events"] for h in hist_a[11:14]]
        assert sum(len(h["sn_events"]) for h in hist_a) >= 1
        inv_b, fin_b, alive_b, kicked_b = enr_b.get()
        assert np.array_equal(alive_a, alive_b) and np.array_equal(kicked_a, kicked_b)
        scale = np.max(np.abs(inv_a), axis=1, keepdims=True) + 1e-300
        assert np.max(np.abs(inv_a - inv_b) / scale) < 1e-9 and np.max(np.abs(fin_a - fin_b) / scale) < 1e-9
        assert np.any(inv_a[pkg.ROW["global26"]] > 0)
        x_b = np.stack([grav_b.particles.x.value_in(U.pc), grav_b.particles.vx.value_in(U.kms)])
        assert np.max(np.abs(x_a - x_b)) < 1e-9 * np.max(np.abs(x_a))
        assert grav_b.model_time.value_in(U.Myr) == pytest.approx(t_a, rel=1e-13)
        # the yields book went on in the same CSV: header + one row per save step of both runs
        rows = open(base + "-cluster-yields.csv").read().strip().split("\n")
        assert rows[0].startswith("time,local_26al") and len(rows) == 1 + 3 + 1
    finally:
        grav_b.stop()
