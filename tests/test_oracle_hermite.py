"""CPU: analytic known-answer tests that anchor the Hermite oracle (parity against AMUSE ph4 is
unpinned -- SURVEY 8c; these are the anchors named there)."""
import numpy as np
import pytest

from oracle import hermite as H


def two_body(e=0.0, m1=0.6, m2=0.4):
    """Bound two-body orbit, a = 1, G = 1, centre of mass at rest; starts at apocentre."""
    mt = m1 + m2
    r = 1.0 + e
    v = np.sqrt(mt * (1.0 - e) / (1.0 + e))
    x = np.array([-m2 / mt * r, m1 / mt * r]); y = np.zeros(2); z = np.zeros(2)
    vy = np.array([-m2 / mt * v, m1 / mt * v]); vx = np.zeros(2); vz = np.zeros(2)
    return np.array([m1, m2]), x, y, z, vx, vy, vz


def plummer_equal(n, seed):
    import importlib
    ic = importlib.import_module("26al-nbody_b200").ic
    return ic.plummer(n, np.random.default_rng(seed))


def rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b))


def test_force_double_vs_long_double_and_momentum():
    m, x, y, z, vx, vy, vz = plummer_equal(1000, 0)
    m = np.random.default_rng(1).uniform(0.1, 2.0, 1000) / 1000
    fd = H.force(m, x, y, z, vx, vy, vz)
    fl = H.force(m, x, y, z, vx, vy, vz, long_double=True)
    a_d, a_l = np.stack(fd[:3]), np.stack(fl[:3])
    assert np.max(np.linalg.norm(a_d - a_l, axis=0) / np.linalg.norm(a_l, axis=0)) < 1e-12
    j_d, j_l = np.stack(fd[3:6]), np.stack(fl[3:6])
    assert np.max(np.linalg.norm(j_d - j_l, axis=0) / np.linalg.norm(j_l, axis=0)) < 1e-12
    assert rel(fd[6], fl[6]) < 1e-13
    # sum m a = 0, sum m jerk = 0 to rounding
    assert np.max(np.abs((m * a_l).sum(axis=1))) < 1e-13 * np.abs(m * a_l).sum()
    assert np.max(np.abs((m * j_l).sum(axis=1))) < 1e-13 * np.abs(m * j_l).sum()


def test_force_two_body_analytic_and_softening():
    m = np.array([2.0, 3.0])
    x = np.array([0.0, 2.0]); o = np.zeros(2)
    vx = np.array([0.0, 0.0]); vy = np.array([0.0, 1.0])
    ax, ay, az, jx, jy, jz, pot = H.force(m, x, o, o, vx, vy, o)
    assert ax[0] == pytest.approx(3.0 / 4.0) and ax[1] == pytest.approx(-2.0 / 4.0)
    assert jy[0] == pytest.approx(3.0 / 8.0) and jy[1] == pytest.approx(-2.0 / 8.0)
    assert pot[0] == pytest.approx(-1.5) and pot[1] == pytest.approx(-1.0)
    ax2 = H.force(m, x, o, o, vx, vy, o, eps2=5.0)[0]
    assert ax2[0] == pytest.approx(3.0 * 2.0 / 27.0)
    # coincident particles (r2 == 0, eps2 == 0) are masked, not NaN
    out = H.force(np.ones(3), np.array([0.0, 0.0, 1.0]), np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(3))
    assert np.all(np.isfinite(np.stack(out))) and out[0][0] == pytest.approx(1.0)


@pytest.mark.parametrize("e", [0.0, 0.5, 0.9])
def test_kepler_orbit(e):
    m, *ps = two_body(e)
    o = H.HermiteOracle(2, eta=0.05)
    o.commit(m, *ps)
    k0, u0, _ = o.energies()
    period = 2.0 * np.pi
    o.evolve(period)
    k1, u1, _ = o.energies()
    assert abs((k1 + u1) - (k0 + u0)) / abs(k0 + u0) < 2e-7
    st = o.get_state()
    for a, b in zip(st[1:], ps):
        assert np.max(np.abs(a - b)) < 2e-4  # back at the start after one period
    # angular momentum
    L0 = np.sum(m * (ps[0] * ps[4] - ps[1] * ps[3])); L1 = np.sum(st[0] * (st[1] * st[5] - st[2] * st[4]))
    assert abs(L1 - L0) / abs(L0) < 1e-7
    assert o.get_time() == period


def test_fourth_order_convergence():
    m, *ps = two_body(0.5)
    errs = []
    for eta in (0.08, 0.04):
        o = H.HermiteOracle(2, eta=eta)
        o.commit(m, *ps)
        k0, u0, _ = o.energies()
        o.evolve(2.0 * np.pi)
        k1, u1, _ = o.energies()
        errs.append(abs((k1 + u1) - (k0 + u0)))
    # dt ~ eta; block quantisation makes the ratio lumpy: 4th order => ~16x, accept >= 6x
    assert errs[0] / errs[1] > 6.0


def test_block_step_invariants_and_stepwise_equals_evolve():
    m, *ps = plummer_equal(256, 4)
    a = H.HermiteOracle(256); a.commit(m, *ps)
    b = H.HermiteOracle(256); b.commit(m, *ps)
    t_end = 0.03
    b.begin(t_end)
    steps = 0
    while True:
        idx, tn = b.get_active()
        t, dt = b.get_timesteps()
        e = np.log2(dt)
        assert np.all(e == np.round(e)) and np.all(dt <= 2.0 ** -5)
        assert np.all(np.mod(t, dt) == 0.0)  # every particle sits on its own ladder
        nd, fin = b.advance(1)
        if fin:
            break
        steps += nd
        t2, dt2 = b.get_timesteps()
        assert np.all(t2[idx] == tn)
        grew = dt2[idx] > dt[idx]
        assert np.all(np.mod(tn, dt2[idx][grew]) == 0.0)  # doubling only when commensurate
        assert np.all((dt2[idx] == dt[idx]) | (dt2[idx] == 2 * dt[idx]) | (dt2[idx] == dt[idx] / 2))
    b.finish()
    ns, npairs = a.evolve(t_end)
    assert ns == steps + 1  # + the synchronisation step
    for u, v in zip(a.get_state(), b.get_state()):
        assert np.array_equal(u, v)
    assert np.all(b.get_timesteps()[0] == 0.0) and b.get_time() == t_end


def test_plummer_virial_identities():
    m, *ps = plummer_equal(4000, 7)
    o = H.HermiteOracle(4000); o.commit(m, *ps)
    k, u, s = o.energies()
    assert k == pytest.approx(0.25, rel=0.06) and u == pytest.approx(-0.5, rel=0.06)
    assert 1.0 / (2.0 * s) == pytest.approx(1.0, rel=0.06)  # R_vir = M^2 / (2 S)
    assert u == pytest.approx(-s, rel=1e-14)                # eps2 = 0


def test_set_mass_reinitialises_and_time_setter():
    m, *ps = plummer_equal(64, 2)
    o = H.HermiteOracle(64); o.commit(m, *ps)
    o.evolve(0.01)
    o.set_mass(m)  # marks dirty: forces are recomputed from the corrected positions
    o.initialize()
    a0 = o.get_acc_jerk()[0].copy()
    o.set_mass(2.0 * m)
    o.set_time(5.0)
    o.initialize()
    assert np.allclose(o.get_acc_jerk()[0], 2.0 * a0, rtol=1e-12)
    o.evolve(5.01)
    assert o.get_time() == 5.01


def figure_eight():
    """Chenciner & Montgomery (2000) figure-eight choreography, G = m = 1 (the author's informal
    three-body check, notes.md:19); initial conditions from Simo's table, period 6.32591398."""
    r1 = np.array([0.97000436, -0.24308753, 0.0])
    v3 = np.array([-0.93240737, -0.86473146, 0.0])
    pos = np.stack([r1, -r1, np.zeros(3)])
    vel = np.stack([-0.5 * v3, -0.5 * v3, v3])
    return np.ones(3), pos[:, 0].copy(), pos[:, 1].copy(), pos[:, 2].copy(), vel[:, 0].copy(), vel[:, 1].copy(), vel[:, 2].copy()


def test_figure_eight_three_body_returns_after_one_period():
    m, *ps = figure_eight()
    o = H.HermiteOracle(3, eta=0.02)
    o.commit(m, *ps)
    k0, u0, _ = o.energies()
    assert k0 + u0 == pytest.approx(-1.287146, rel=1e-5)  # the choreography's energy
    T = 6.32591398
    o.evolve(T / 3.0)                                      # after T/3 the bodies have cycled places: 1 -> 2 -> 3 -> 1 (up to labelling)
    st = o.get_state()
    third = np.stack(st[1:4]).T
    start = np.stack(ps[:3]).T
    d = [min(np.linalg.norm(third[i] - start[j]) for j in range(3)) for i in range(3)]
    assert max(d) < 2e-5
    o.evolve(T)
    st = o.get_state()
    for a, b in zip(st[1:], ps):
        assert np.max(np.abs(a - b)) < 5e-5                # back at the start after one period
    k1, u1, _ = o.energies()
    assert abs((k1 + u1) - (k0 + u0)) / abs(k0 + u0) < 1e-8


def test_mass_update_policy_and_softened_self_pair():
    """Mass-only update: policy 0 keeps the synchronisation step's timesteps, policy 1 re-derives initial ones; with
    softening the potential excludes the self pair."""
    import importlib
    pkg = importlib.import_module("26al-nbody_b200")
    n = 200
    c = pkg.ic.cluster(n, seed=3, require_massive=False)
    p = [c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")]
    out = {}
    for policy in (0, 1):
        o = H.HermiteOracle(n); o.commit(*p); o.set_reinit_policy(policy)
        o.evolve(0.01)
        dt_sync = o.get_timesteps()[1].copy()
        o.set_mass(p[0] * 0.95)
        o.initialize()
        dt_now = o.get_timesteps()[1]
        if policy == 0:
            assert np.array_equal(dt_now, dt_sync)
        else:
            assert np.all(dt_now <= 2.0 ** -5) and not np.array_equal(dt_now, dt_sync)
        out[policy] = o.evolve(0.02)
    assert out[0][0] < out[1][0]  # keeping the timesteps saves the climb back up the ladder
    pot = H.force(*p, eps2=0.01)[6]
    i = 7
    want = -sum(p[0][j] / np.sqrt((p[1][i] - p[1][j]) ** 2 + (p[2][i] - p[2][j]) ** 2 + (p[3][i] - p[3][j]) ** 2 + 0.01)
                for j in range(n) if j != i)
    assert pot[i] == pytest.approx(want, rel=1e-13)
