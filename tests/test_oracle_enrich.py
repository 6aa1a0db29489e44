"""CPU: the enrichment oracle against the reference's own functions (golden fixtures made by
oracle/lift_reference.py from /root/reference/al26_nbody.py) and SURVEY 8(c) vectors."""
import numpy as np
import pytest

from conftest import golden_case
from oracle import enrich_oracle as eo
from oracle import lift_reference as lr


def _wind(c, wr, limit, radius):
    return eo.calc_wind_abs(c["lm_id"], c["hm_id"], c["x"], c["y"], c["z"], c["vx"], c["vy"], c["vz"], c["mdot"],
                            c[wr], c["rdisk"], limit, radius, float(c["dt_s"]))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_wind_bit_exact_vs_reference_golden(golden, tag):
    c = golden_case(golden, tag)
    b, rv = float(c["bubble_km"]), float(c["rvir_km"])
    assert np.array_equal(_wind(c, "wr26", 0.0, rv), c["g26"])
    assert np.array_equal(_wind(c, "wr60", 0.0, rv), c["g60"])
    assert np.array_equal(_wind(c, "wr26", b, b), c["l26"])
    assert np.array_equal(_wind(c, "wr60", b, b), c["l60"])
    assert np.count_nonzero(c["l26"]) > 0  # the local model has hits in the fixture


def test_survey_golden_vectors(golden):
    # calc_wind_abs on 3 stars
    x = np.array([0.0, 1e12, 5e12]); y = np.zeros(3); z = np.zeros(3)
    vx = np.array([0.0, 3.0, 3.0]); vy = np.array([0.0, 4.0, 4.0]); vz = np.zeros(3)
    mdot = np.array([1e15, 0.0, 0.0]); wr = np.array([1e-4, 0.0, 0.0]); rd = np.array([0.0, 1.5e10, 1.5e10])
    lm, hm = np.array([1, 2]), np.array([0])
    g = eo.calc_wind_abs(lm, hm, x, y, z, vx, vy, vz, mdot, wr, rd, 0.0, 3e12, 1e11)
    l = eo.calc_wind_abs(lm, hm, x, y, z, vx, vy, vz, mdot, wr, rd, 3e12, 3e12, 1e11)
    assert g.tolist() == [0.0, 3.1249999999999996e16, 3.1249999999999996e16]
    assert l.tolist() == [0.0, 3.1249999999999996e16, 0.0]
    assert np.array_equal(g, golden["tiny_global"]) and np.array_equal(l, golden["tiny_local"])
    assert eo.calc_eta_disk_sne(100.0, 206264.806) == 1.0283188385492955e-08
    assert float(golden["eta_sne_100_206264p806"]) == 1.0283188385492955e-08
    assert float(golden["intersection"]) == 0.0859375
    f26, f60 = eo.decay_fractions(0.01)
    # np.exp may differ by 1 ulp between SIMD dispatch paths of numpy (AVX-512 vs libm)
    assert f26 == pytest.approx(0.99037925616650468, rel=3e-16) and f60 == pytest.approx(0.99733760048885856, rel=3e-16)


@pytest.mark.skipif(not lr.available(), reason="reference file only exists in the build container")
def test_wind_bit_exact_vs_lifted_reference_live():
    fn = lr.lift(("calc_wind_abs", "calc_eta_disk_sne"))
    c = lr.make_wind_case(np.random.default_rng(11), 1500, 23)
    dt = 3.15e11
    args = (c["lm_id"], c["hm_id"], c["x"], c["y"], c["z"], c["vx"], c["vy"], c["vz"], c["mdot"])
    b = float(c["bubble_km"])
    for wr in ("wr26", "wr60"):
        assert np.array_equal(fn["calc_wind_abs"](*args, c[wr], c["rdisk"], 0.0, 7.0 * b, dt),
                              eo.calc_wind_abs(*args, c[wr], c["rdisk"], 0.0, 7.0 * b, dt))
        assert np.array_equal(fn["calc_wind_abs"](*args, c[wr], c["rdisk"], b, b, dt),
                              eo.calc_wind_abs(*args, c[wr], c["rdisk"], b, b, dt))
    for r, d in ((100.0, 206264.806), (1.5e10, 3.3e12), (7.0, 0.3)):
        assert fn["calc_eta_disk_sne"](r, d) == eo.calc_eta_disk_sne(r, d)


def test_classify_and_empty_sets():
    m = np.array([0.05, 0.1, 3.0, 3.0000001, 12.999, 13.0, 150.0])
    hm, lm = eo.classify(m)
    assert hm.tolist() == [5, 6] and lm.tolist() == [1, 2]
    n = 4
    z = np.zeros(n)
    assert np.array_equal(eo.calc_wind_abs(np.array([], dtype=int), np.array([0]), z, z, z, z, z, z, z, z, z, 0.0, 1.0, 1.0), z)


def test_sqrt_threshold_is_exact():
    rng = np.random.default_rng(5)
    for radius in (3.0856775814913e12, 1.0, 0.1, 12345.678):
        q = eo.sqrt_threshold(radius)
        d2 = np.concatenate([q * (1 + rng.uniform(-1e-15, 1e-15, 2000)), [q, np.nextafter(q, 0), np.nextafter(q, np.inf)]])
        assert np.array_equal(radius <= np.sqrt(d2), d2 >= q)


def _state(n, rng, n_hm=5):
    mass = rng.uniform(0.05, 5.0, n)
    hm = rng.choice(n, n_hm, replace=False)
    mass[hm] = rng.uniform(13.0, 60.0, n_hm)
    wr26 = np.zeros(n); wr60 = np.zeros(n); sn26 = np.zeros(n); sn60 = np.zeros(n)
    wr26[hm] = 1e-5; wr60[hm] = 1e-7; sn26[hm] = 1e26; sn60[hm] = 3e25
    st = eo.EnrichState(np.full(n, 1.5e10), rng.exponential(2.885, n), (mass >= 0.1) & (mass <= 3.0),
                        np.zeros(n, bool), wr26, wr60, sn26, sn60)
    return st, mass, hm


def test_enrich_step_semantics():
    """deposit -> decay -> condense order, one-shot SN, *_final freeze."""
    rng = np.random.default_rng(3)
    n = 300
    st, mass, hm = _state(n, rng)
    pos = rng.normal(0, 3e13, (3, n)); vel = rng.normal(0, 1.0, (3, n))
    mdot = np.zeros(n); mdot[hm] = 1e16
    f26, f60 = eo.decay_fractions(0.01)
    ev = eo.enrich_step(st, mass, mdot, *pos, *vel, 3.15e11, 0.01, 3e12, 6e13, f26, f60)
    assert ev == []
    lm = eo.classify(mass)[1]
    assert np.all(st.inv[eo.GLOBAL26, lm] > 0) and np.all(st.inv[eo.SNE26] == 0)
    alive_before = st.disk_alive.copy()
    # star hm[0] dies: mdot == 0 -> SN once
    mdot[hm[0]] = 0.0
    ev = eo.enrich_step(st, mass, mdot, *pos, *vel, 3.15e11, 0.02, 3e12, 6e13, f26, f60)
    assert ev == [int(hm[0])] and st.kicked[hm[0]]
    assert np.all(st.inv[eo.SNE26, lm] > 0)
    ev = eo.enrich_step(st, mass, mdot, *pos, *vel, 3.15e11, 0.03, 3e12, 6e13, f26, f60)
    assert ev == []
    # condensed discs: flag cleared exactly when tau < t_new, finals frozen at the last step with tau >= t_new
    gone = alive_before & (st.tau_disk < 0.03) & np.isin(np.arange(n), lm)
    assert np.array_equal(st.disk_alive[lm], (alive_before & ~gone)[lm])
    still = lm[st.disk_alive[lm]]
    assert np.array_equal(st.fin[eo.GLOBAL26, still], st.inv[eo.GLOBAL26, still])
    for i in lm[gone[lm]]:
        assert st.fin[eo.GLOBAL26, i] != st.inv[eo.GLOBAL26, i] or st.inv[eo.GLOBAL26, i] == 0


def test_intersection_bit_exact_vs_reference_golden(golden):
    """calc_intersection (al26_nbody.py:1156-1190): restatement == the reference's own function"""
    a0, a1, b0, b1, fr = (golden[k] for k in ("isect_a_old", "isect_a_new", "isect_b_old", "isect_b_new", "isect_frac"))
    mine = np.array([eo.calc_intersection(*a0, *a1, *b0[:, i], *b1[:, i], 0.1) for i in range(b0.shape[1])])
    assert np.array_equal(mine, fr) and np.count_nonzero(fr) > 20
    assert eo.calc_intersection(-1, 0, 0, 1, 0, 0, 0, 0.05, 0, 0, 0.05, 0, 0.1) == 0.0859375


@pytest.mark.parametrize("tag", ["ld_a", "ld_b"])
def test_local_density_oracle_bit_exact_vs_reference_golden(golden, tag):
    from oracle import analysis_oracle as ao
    rho = ao.local_densities(golden[tag + "_x"], golden[tag + "_y"], golden[tag + "_z"], golden[tag + "_m"])
    assert np.array_equal(rho, golden[tag + "_rho"])
