"""GPU parity: the fused enrichment kernel (through the C-ABI) against the golden vectors made from
the reference's own numba kernel, and against the numpy oracle over multi-step runs.  Wind sums are
bit-exact; every integer / flag output (SN event lists, kicked, disk_alive) is bit-exact."""
import numpy as np
import pytest

from conftest import golden_case
from oracle import enrich_oracle as eo

pytestmark = pytest.mark.gpu


def test_survey_golden_three_stars(pkg, ctx):
    e = pkg.EnrichCore(ctx=ctx)
    x = np.array([0.0, 1e12, 5e12]); z = np.zeros(3)
    pv = np.stack([x, z, z, np.array([0.0, 3.0, 3.0]), np.array([0.0, 4.0, 4.0]), z])
    mass = np.array([20.0, 1.0, 1.0]); mdot = np.array([1e15, 0.0, 0.0])
    e.commit(np.array([0.0, 1.5e10, 1.5e10]), np.full(3, 1e9), [0, 1, 1], [1, 0, 0], [1e-4, 0, 0], [0, 0, 0], z, z)
    ev = e.step(mass, mdot, pv, 1e11, 0.01, 3e12, 3e12, 1.0, 1.0)
    inv, fin, alive, kicked = e.get()
    assert len(ev) == 0
    assert inv[pkg.ROW["global26"]].tolist() == [0.0, 3.1249999999999996e16, 3.1249999999999996e16]
    assert inv[pkg.ROW["local26"]].tolist() == [0.0, 3.1249999999999996e16, 0.0]
    assert np.array_equal(fin, inv)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_wind_bit_exact_vs_reference_golden(pkg, ctx, golden, tag):
    c = golden_case(golden, tag)
    n = len(c["x"])
    mass = np.full(n, 5.0)  # neither class
    mass[c["lm_id"]] = 1.0
    mass[c["hm_id"]] = 20.0
    e = pkg.EnrichCore(ctx=ctx)
    e.commit(c["rdisk"], np.full(n, 1e9), np.ones(n), np.ones(n), c["wr26"], c["wr60"], np.zeros(n), np.zeros(n))
    pv = np.stack([c[k] for k in ("x", "y", "z", "vx", "vy", "vz")])
    ev = e.step(mass, c["mdot"], pv, float(c["dt_s"]), 0.01, float(c["bubble_km"]), float(c["rvir_km"]), 1.0, 1.0)
    assert len(ev) == 0
    inv = e.get()[0]
    R = pkg.ROW
    assert np.array_equal(inv[R["global26"]], c["g26"]) and np.array_equal(inv[R["global60"]], c["g60"])
    assert np.array_equal(inv[R["local26"]], c["l26"]) and np.array_equal(inv[R["local60"]], c["l60"])
    assert np.count_nonzero(inv[R["local26"]]) > 0


def _random_problem(n, n_hm, seed):
    rng = np.random.default_rng(seed)
    mass = 10.0 ** rng.uniform(-1.5, 0.9, n)          # some < 0.1, some in (3, 13)
    hm = np.sort(rng.choice(n, n_hm, replace=False))
    mass[hm] = rng.uniform(13.0, 80.0, n_hm)
    wr26 = np.zeros(n); wr60 = np.zeros(n); sn26 = np.zeros(n); sn60 = np.zeros(n)
    wr26[hm] = 10.0 ** rng.uniform(-7, -4, n_hm); wr60[hm] = 10.0 ** rng.uniform(-9, -6, n_hm)
    sn26[hm] = rng.uniform(1e25, 1e26, n_hm) * (mass[hm] <= 25.0)  # SN yield 0 above 25 Msun (:460-461)
    sn60[hm] = rng.uniform(1e24, 1e25, n_hm) * (mass[hm] <= 25.0)
    tau = rng.exponential(0.05, n)
    alive = (mass >= 0.1) & (mass <= 3.0)
    return dict(mass=mass, hm=hm, wr26=wr26, wr60=wr60, sn26=sn26, sn60=sn60, tau=tau, alive=alive,
                rdisk=np.full(n, 1.49597870691e10), rng=rng)


@pytest.mark.parametrize("n,n_hm", [(1000, 7), (5000, 600), (300, 0), (64, 64)])
def test_multi_step_vs_oracle(pkg, ctx, n, n_hm):
    P = _random_problem(n, n_hm, seed=n + n_hm)
    rng = P["rng"]
    e = pkg.EnrichCore(ctx=ctx)
    e.commit(P["rdisk"], P["tau"], P["alive"], np.zeros(n), P["wr26"], P["wr60"], P["sn26"], P["sn60"])
    st = eo.EnrichState(P["rdisk"], P["tau"], P["alive"], np.zeros(n, bool), P["wr26"], P["wr60"], P["sn26"], P["sn60"])
    pos = rng.normal(0, 3e13, (3, n)); vel = rng.normal(0, 1.0, (3, n))
    if n_hm:
        pos[:, ::5] = pos[:, P["hm"][0:1]] + rng.normal(0, 1e12, (3, len(pos[0, ::5])))  # discs inside a local bubble
        pos[:, P["hm"]] += 1.0
    mdot = np.zeros(n); mdot[P["hm"]] = 10.0 ** rng.uniform(14, 17, n_hm)
    dt_myr, dt_s = 0.01, 0.01 * 1e6 * 365.242199 * 86400
    f26, f60 = eo.decay_fractions(dt_myr)
    mass = P["mass"].copy()
    all_events = []
    for step in range(1, 9):
        if n_hm and step in (3, 5):  # two stars die: mdot -> 0
            mdot[P["hm"][step % n_hm]] = 0.0
        if step == 6:  # a remnant drops into the disc-bearing range: it starts receiving deposits
            if n_hm:
                mass[P["hm"][0]] = 1.4
        pos += vel * 1e10
        pv = np.concatenate([pos, vel])
        ev_g = e.step(mass, mdot, pv, dt_s, step * dt_myr, 3.0856775814913e12, 2.1e14, f26, f60)
        ev_o = eo.enrich_step(st, mass, mdot, *pos, *vel, dt_s, step * dt_myr, 3.0856775814913e12, 2.1e14, f26, f60)
        assert ev_g.tolist() == ev_o, f"SN event list differs at step {step}"
        all_events += ev_o
        inv, fin, alive, kicked = e.get()
        assert np.array_equal(alive, st.disk_alive), f"disk_alive differs at step {step}"
        assert np.array_equal(kicked, st.kicked)
        assert np.array_equal(inv, st.inv), f"inventories differ at step {step}"
        assert np.array_equal(fin, st.fin), f"finals differ at step {step}"
    if n_hm >= 7:
        assert len(all_events) == 2
        assert np.any(inv[pkg.ROW["sne26"]] > 0) or np.all(P["sn26"][all_events] == 0)
    if P["alive"].any():
        assert (~alive[P["alive"]]).sum() > 0  # some discs condensed during the run


def test_reads_gravity_state_in_place(pkg, ctx):
    n = 2048
    c = pkg.ic.cluster(n, seed=4)
    g = pkg.GravityCore(ctx=ctx)
    g.commit(c["m"], c["x"], c["y"], c["z"], c["vx"], c["vy"], c["vz"])
    cv = pkg.units.nbody_to_si(1.0 | pkg.units.pc, c["m_msun"].sum() | pkg.units.MSun)
    mass = c["m_msun"]
    hm = np.nonzero(mass >= 13.0)[0]
    wr = np.zeros(n); wr[hm] = 1e-5
    mdot = np.zeros(n); mdot[hm] = 1e16
    args = (np.full(n, 1.5e10), np.full(n, 1e9), np.ones(n), np.zeros(n), wr, wr, np.zeros(n), np.zeros(n))
    e = pkg.EnrichCore(ctx=ctx)
    e.commit(*args)
    e.set_units(cv.km_per_length, cv.kms_per_speed)
    e.step(mass, mdot, None, 3e11, 0.01, 3e12, 3e13, 1.0, 1.0)
    inv_a = e.get()[0]
    st = g.get_state()
    pv = np.stack([st[1] * cv.km_per_length, st[2] * cv.km_per_length, st[3] * cv.km_per_length,
                   st[4] * cv.kms_per_speed, st[5] * cv.kms_per_speed, st[6] * cv.kms_per_speed])
    e.commit(*args)
    e.step(mass, mdot, pv, 3e11, 0.01, 3e12, 3e13, 1.0, 1.0)
    assert np.array_equal(inv_a, e.get()[0])
    assert np.count_nonzero(inv_a[pkg.ROW["global26"]]) == np.count_nonzero((mass >= 0.1) & (mass <= 3.0))


def test_full_size_properties_1e6(pkg, ctx):
    """BASELINE config 5 size (1000 sources x 1e6 discs): linearity + decay properties."""
    n, n_hm = 1_001_000, 1000
    rng = np.random.default_rng(0)
    mass = np.full(n, 1.0); hm = np.arange(0, n, n // n_hm)[:n_hm]; mass[hm] = 20.0
    wr = np.zeros(n); wr[hm] = 1e-5
    mdot = np.zeros(n); mdot[hm] = 1e16
    pv = np.concatenate([rng.normal(0, 3e13, (3, n)), rng.normal(0, 1.0, (3, n))])
    e = pkg.EnrichCore(ctx=ctx)
    e.commit(np.full(n, 1.5e10), np.full(n, 1e9), np.ones(n), np.zeros(n), wr, 2.0 * wr, np.zeros(n), np.zeros(n))
    e.step(mass, mdot, pv, 3e11, 0.01, 3e12, 2e14, 1.0, 1.0)
    inv1 = e.get()[0]
    R = pkg.ROW
    lm = mass == 1.0
    assert np.all(inv1[R["global26"]][lm] > 0) and np.all(inv1[:, ~lm] == 0)
    # 60Fe ratio is exactly twice the 26Al one -> every term doubles exactly
    assert np.array_equal(inv1[R["global60"]], 2.0 * inv1[R["global26"]])
    assert np.array_equal(inv1[R["local60"]], 2.0 * inv1[R["local26"]])
    # a sample of discs against the oracle, bit-exact
    samp = np.sort(rng.choice(np.nonzero(lm)[0], 200, replace=False))
    ref_g = eo.calc_wind_abs(samp, hm, *pv, mdot, wr, np.full(n, 1.5e10), 0.0, 2e14, 3e11)
    ref_l = eo.calc_wind_abs(samp, hm, *pv, mdot, wr, np.full(n, 1.5e10), 3e12, 3e12, 3e11)
    assert np.array_equal(inv1[R["global26"]][samp], ref_g[samp]) and np.array_equal(inv1[R["local26"]][samp], ref_l[samp])
    # second step with no sources alive: pure decay
    e.step(mass, np.zeros(n), pv, 3e11, 0.02, 3e12, 2e14, 0.5, 0.25)
    inv2 = e.get()[0]
    assert np.array_equal(inv2[R["global26"]], inv1[R["global26"]] * 0.5)
    assert np.array_equal(inv2[R["global60"]], inv1[R["global60"]] * 0.25)


def test_interloper_vs_reference_golden_and_oracle(pkg, ctx, golden):
    """AGB interloper: intersection fractions bit-exact against the reference's own
    calc_intersection (golden from the lifted function), deposits bit-exact against the oracle."""
    a0, a1, b0, b1, fr = (golden[k] for k in ("isect_a_old", "isect_a_new", "isect_b_old", "isect_b_new", "isect_frac"))
    m = b0.shape[1]
    n = m + 1                                  # the interloper is the last star (:974)
    old = np.concatenate([b0, a0[:, None]], axis=1)
    new = np.concatenate([b1, a1[:, None]], axis=1)
    mass = np.full(n, 1.0); mass[5] = 8.0      # star 5 bears no disc
    is_int = np.zeros(n, bool); is_int[-1] = True
    rd = np.full(n, 1.49597870691e10)
    e = pkg.EnrichCore(ctx=ctx)
    e.commit(rd, np.full(n, 1e9), np.ones(n), np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n))
    st = eo.EnrichState(rd, np.full(n, 1e9), np.ones(n, bool), np.zeros(n, bool), *(np.zeros(n),) * 4)
    raw = np.zeros((2, n))
    kmpc, rbub, r26, r60, dt = 3.08567758128e13, 0.1 * 3.08567758128e13, 2.0e13, 3.0e10, 3.15e11
    for rep in range(2):                       # two steps: deposits accumulate, raw is never decayed
        e.interloper(mass, old, new, n - 1, rbub, r26, r60, dt)
        frac = eo.interloper_step(st, raw, mass, is_int, old, new, r26, r60, dt, rbub, kmpc)
        assert np.array_equal(frac[:m][mass[:m] == 1.0], fr[mass[:m] == 1.0])  # oracle == reference function
        f26, f60 = 0.9, 0.95
        e.step(mass, np.zeros(n), np.concatenate([new, np.zeros((3, n))]) * kmpc, dt, 0.01 * (rep + 1), 3e12, 3e13, f26, f60, with_agb=True)
        eo.enrich_step(st, mass, np.zeros(n), *(new * kmpc), *np.zeros((3, n)), dt, 0.01 * (rep + 1), 3e12, 3e13, f26, f60, with_agb=True)
    inv, fin, alive, kicked = e.get()
    R = pkg.ROW
    assert np.array_equal(inv, st.inv) and np.array_equal(fin, st.fin)
    assert np.array_equal(e.get_agb_raw(), raw)
    hit = (fr > 0) & np.any(b1 != b0, axis=0)  # a disc that does not move sweeps up nothing (d_disk_trav = 0)
    hit[5] = False
    assert np.array_equal(inv[R["agb26"]][:m] > 0, hit) and inv[R["agb26"]][-1] == 0.0 and raw[0, 5] == 0.0
    assert np.all(raw[0][:m][hit] > inv[R["agb26"]][:m][hit])  # the inventory decays, raw does not


@pytest.mark.parametrize("tag", ["ld_a", "ld_b"])
def test_local_densities_bit_exact_vs_reference_golden(pkg, ctx, golden, tag):
    """plotting/al26_plot.py:324-359: bit-exact against the reference's own numba function"""
    rho = ctx.local_densities(golden[tag + "_x"], golden[tag + "_y"], golden[tag + "_z"], golden[tag + "_m"])
    assert np.array_equal(rho, golden[tag + "_rho"])


def test_local_densities_large_sample_vs_oracle(pkg, ctx):
    from oracle import analysis_oracle as ao
    c = pkg.ic.cluster(20000, seed=3)
    rho = ctx.local_densities(c["x"], c["y"], c["z"], c["m_msun"])
    idx = np.random.default_rng(0).choice(20000, 40, replace=False)
    x, y, z, m = c["x"], c["y"], c["z"], c["m_msun"]
    for i in idx:  # the oracle is O(N) per star
        d = np.sqrt((x[i] - x) * (x[i] - x) + (y[i] - y) * (y[i] - y) + (z[i] - z) * (z[i] - z))
        nr = np.argsort(d, kind="stable")[1:11]
        mass = 0.0
        for j in nr:
            mass += m[j]
        assert rho[i] == mass / (ao.FTP * d[nr[-1]] * d[nr[-1]] * d[nr[-1]])
    with pytest.raises(pkg.Al26Error):
        ctx.local_densities(x[:5], y[:5], z[:5], m[:5])


def _run_modes(pkg, ctx, n, n_hm, seed, steps=6, clustered=False):
    """the same multi-step problem in the three disc-kernel modes; returns {mode: (inv, fin, alive, kicked, events)}"""
    out = {}
    for mode in (0, 1, 2):
        P = _random_problem(n, n_hm, seed=seed)
        rng = P["rng"]
        for s_ in ((2, 4) if n_hm else ()):  # the two stars that die below do have a supernova yield
            P["sn26"][P["hm"][s_ % n_hm]], P["sn60"][P["hm"][s_ % n_hm]] = 5e25, 5e24
        e = pkg.EnrichCore(ctx=ctx)
        e.set_mode(mode)
        try:
            e.commit(P["rdisk"], P["tau"], P["alive"], np.zeros(n), P["wr26"], P["wr60"], P["sn26"], P["sn60"])
            pos = rng.normal(0, 3e13, (3, n)); vel = rng.normal(0, 1.0, (3, n))
            if clustered:  # every disc within a few bubble radii of the sources: many hits per disc (mode 2's fallback)
                pos = rng.normal(0, 2e12, (3, n))
            elif n_hm:
                pos[:, ::5] = pos[:, P["hm"][0:1]] + rng.normal(0, 1e12, (3, len(pos[0, ::5])))
            mdot = np.zeros(n); mdot[P["hm"]] = 10.0 ** rng.uniform(14, 17, n_hm)
            dt_myr, dt_s = 0.01, 0.01 * 1e6 * 365.242199 * 86400
            f26, f60 = eo.decay_fractions(dt_myr)
            events = []
            for step in range(1, steps + 1):
                if n_hm and step in (2, 4):
                    mdot[P["hm"][step % n_hm]] = 0.0
                pos += vel * 1e10
                events.append(e.step(P["mass"], mdot, np.concatenate([pos, vel]), dt_s, step * dt_myr, 3.0856775814913e12,
                                     2.1e14, f26, f60).tolist())
            out[mode] = e.get() + (events,)
        finally:
            e.set_mode(0)
    return out


@pytest.mark.parametrize("n,n_hm,clustered", [(4000, 9, False), (6000, 700, False), (3000, 40, True), (500, 0, False)])
def test_fast_modes_agree_with_the_exact_mode(pkg, ctx, n, n_hm, clustered):
    """Modes 1 (hoisted global sum, expanded bubble test) and 2 (cell-grid pruning) against mode 0, which is itself
    bit-exact against the reference kernel: integer work and the local / SN rows identical, the global rows within the
    north_star's 1e-10 (here: 1e-12, summation order only)."""
    out = _run_modes(pkg, ctx, n, n_hm, seed=n + n_hm, clustered=clustered)
    R = pkg.ROW
    inv0, fin0, alive0, kicked0, ev0 = out[0]
    if n_hm:
        assert np.count_nonzero(inv0[R["local26"]]) > 0 and np.count_nonzero(inv0[R["sne26"]]) > 0
    for mode in (1, 2):
        inv, fin, alive, kicked, ev = out[mode]
        assert ev == ev0 and np.array_equal(alive, alive0) and np.array_equal(kicked, kicked0)
        for row in ("local26", "local60", "sne26", "sne60", "agb26", "agb60"):
            assert np.array_equal(inv[R[row]], inv0[R[row]]), (mode, row)
            assert np.array_equal(fin[R[row]], fin0[R[row]]), (mode, row)
        for row in ("global26", "global60"):
            for a, b in ((inv, inv0), (fin, fin0)):
                nz = b[R[row]] != 0
                assert np.array_equal(a[R[row]] != 0, nz)
                assert np.max(np.abs(a[R[row]][nz] / b[R[row]][nz] - 1.0), initial=0.0) < 1e-12, (mode, row)


def test_capacity_error_leaves_the_state_untouched(pkg, ctx):
    """more massive stars than the source table holds: AL26_ECAP, and neither inventories nor flags have moved"""
    n = 20000
    rng = np.random.default_rng(3)
    mass = np.full(n, 1.0); mass[:9000] = 20.0
    wr = np.where(mass > 13, 1e-5, 0.0)
    pv = np.concatenate([rng.normal(0, 3e13, (3, n)), rng.normal(0, 1.0, (3, n))])
    e = pkg.EnrichCore(ctx=ctx)
    e.commit(np.full(n, 1.5e10), np.full(n, 0.015), np.ones(n), np.zeros(n), wr, wr, wr * 1e30, wr * 1e30)
    ok_mass = mass.copy(); ok_mass[100:9000] = 5.0
    e.step(ok_mass, np.where(ok_mass > 13, 1e16, 0.0), pv, 3e11, 0.01, 3e12, 3e13, 0.99, 0.999)
    before = e.get()
    with pytest.raises(pkg.Al26Error) as ei:
        e.step(mass, np.zeros(n), pv, 3e11, 0.02, 3e12, 3e13, 0.99, 0.999)  # 9000 sources, every one a supernova
    assert ei.value.code == -4
    after = e.get()
    for a, b in zip(before, after):
        assert np.array_equal(a, b)
    # and the context still works
    ev = e.step(ok_mass, np.zeros(n), pv, 3e11, 0.02, 3e12, 3e13, 0.99, 0.999)
    assert len(ev) == 100 and e.get()[3][:100].all()
