"""CPU: host-side logic around the hot path -- stellar provider stub, yield precompute, init_cluster."""
import os

import numpy as np
import pytest


def test_stellar_stub_surface_and_lifecycle(pkg):
    U = pkg.units
    P = pkg.Particles(4)
    P.mass = np.array([0.5, 2.0, 13.0, 40.0]) | U.MSun
    st = pkg.StellarStub(wind_loss_fraction=0.2)
    st.particles.add_particles(P)
    assert len(st.particles) == 4 and np.array_equal(st.particles.key, P.key)
    life = pkg.stellar.approx_lifespan_myr(np.array([13.0, 40.0]))
    assert life[0] == pytest.approx(1e4 * 13.0 ** -2.5) and life[1] < life[0]
    st.evolve_model(0.5 * life[1] | U.Myr)
    m = st.particles.mass.value_in(U.MSun)
    mdot = -st.particles.wind_mass_loss_rate.value_in(U.kg / U.s)
    assert m[0] == 0.5 and m[1] == 2.0 and mdot[0] == 0.0 and mdot[1] == 0.0   # low-mass stars do not evolve
    assert m[3] == pytest.approx(40.0 * (1 - 0.2 * 0.5)) and mdot[3] > 0 and mdot[2] > 0
    st.evolve_model(1.01 * life[1] | U.Myr)                                     # the 40 Msun star dies
    m = st.particles.mass.value_in(U.MSun)
    mdot = -st.particles.wind_mass_loss_rate.value_in(U.kg / U.s)
    assert mdot[3] == 0.0 and m[3] == 1.4 and mdot[2] > 0                       # mdot == 0 is the script's SN signal (:946-949)
    assert st.model_time.value_in(U.Myr) == pytest.approx(1.01 * life[1])
    # channel to another set copies by index
    Q = pkg.Particles(4, keys=P.key.copy()); Q.mass = np.zeros(4) | U.MSun
    st.particles.new_channel_to(Q).copy_attributes(["mass"])
    assert np.array_equal(Q.mass.value_in(U.MSun), m)


def test_yield_interpolation(pkg):
    Y = pkg.YieldTables.synthetic()
    m, y = Y.wind["Al26"]
    f = pkg.stellar.calc_slr_yield
    assert f(m[2], m, y) == pytest.approx(y[2], rel=1e-12)          # Akima in log10 passes through the nodes
    assert y[1] < f(17.0, m, y) < y[2]
    assert f(12.9, m, y) == 0.0 and f(121.0, m, y) == 0.0            # zero outside the table (:460-461)
    ms, ys = Y.sne["Al26"]
    assert f(30.0, ms, ys) == 0.0                                     # SN table stops at 25 Msun
    wr26, wr60, sn26, sn60 = Y.star_yields(20.0, 4.0)
    assert wr26 == pytest.approx(f(20.0, m, y) / 4.0) and sn26 > 0 and sn60 > 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/limongi-chieffi-2018"), reason="reference tables only in the build container")
def test_yield_tables_from_reference_dir(pkg):
    Y = pkg.YieldTables.from_reference_dir("/root/reference/limongi-chieffi-2018")
    assert Y.wind["Al26"][0].tolist() == [13.0, 15.0, 20.0, 25.0, 30.0, 40.0, 60.0, 80.0, 120.0]
    assert Y.sne["Al26"][0].max() == 25.0                             # sne-yields.csv:1
    assert pkg.stellar.calc_slr_yield(26.0, *Y.sne["Fe60"]) == 0.0
    assert 1e-9 < pkg.stellar.calc_slr_yield(60.0, *Y.wind["Al26"]) < 1e-2


def test_init_cluster_columns(pkg):
    U = pkg.units
    cl, cv = pkg.driver.init_cluster("plummer", 800, 1.0 | U.pc, seed=2)
    m = cl.mass.value_in(U.MSun)
    assert len(cl) == 800 and m.max() >= 13.0
    hm, lm = m >= 13.0, (m >= 0.1) & (m <= 3.0)
    assert np.array_equal(cl.disk_alive, lm) and not cl.kicked.any()
    assert np.all(cl.r_disk.value_in(U.au) == 100.0) and np.all(cl.radius.value_in(U.au) == 0.0)
    assert np.all(cl.wind_ratio_26al[~hm] == 0) and np.all(cl.wind_ratio_26al[hm] > 0)
    assert np.all(cl.sn_yield_26al.value_in(U.MSun)[hm & (m > 25.0)] == 0.0)
    assert np.all(cl.mass_26al_local.value_in(U.kg) == 0) and np.all(cl.mass_60fe_sne_final.value_in(U.kg) == 0)
    assert cv.mass_si == pytest.approx(m.sum() * U.MSUN_KG) and cv.length_si == pytest.approx(U.PARSEC_M)
    x = cv.length_to_nbody(cl.x)
    assert abs(np.mean(x)) < 0.2 and 0.3 < np.std(x) < 2.0
    with pytest.raises(ValueError):
        pkg.driver.init_cluster("king", 10, 1.0 | U.pc)


def test_yields_book_and_csv_format(pkg, tmp_path):
    """the reference's cluster-yields.csv format and Yields dict keys (al26_nbody.py:125-264)"""
    U = pkg.units
    cl = pkg.Particles(3)
    kg = U.MSUN_KG
    cl.mass_26al_local = np.array([1.0, 2.0, 3.0]) * 1e-9 * kg | U.kg
    cl.mass_26al_global = np.array([1.0, 1.0, 1.0]) * 1e-10 * kg | U.kg
    cl.mass_26al_sne = np.zeros(3) | U.kg
    cl.mass_60fe_local = np.array([4.0, 0.0, 0.0]) * 1e-12 * kg | U.kg
    cl.mass_60fe_global = np.zeros(3) | U.kg
    cl.mass_60fe_sne = np.array([0.0, 5.0, 0.0]) * 1e-11 * kg | U.kg
    cl.mass_26al_local_final = cl.mass_26al_local
    base = str(tmp_path / "run")
    y = pkg.Yields(base)
    y.update_state(0.0 | U.Myr, cl)
    cl.mass_26al_local = np.array([2.0, 2.0, 3.0]) * 1e-9 * kg | U.kg
    y.update_state(0.1 | U.Myr, cl)
    lines = open(base + "-cluster-yields.csv").read().splitlines()
    assert lines[0] == "time,local_26al,global_26al,sne_26al,local_60fe,global_60fe,sne_60fe"
    assert lines[1] == "0.000000e+00,6.000000e-09,3.000000e-10,0.000000e+00,4.000000e-12,0.000000e+00,5.000000e-11"
    assert lines[2].startswith("1.000000e-01,7.000000e-09,") and len(lines) == 3
    assert y.time == [0.0, 0.1] and len(y.local_26al) == 2 and y.agb_26al[0] == [0.0, 0.0, 0.0]
    assert y.local_26al_final == pytest.approx([1e-9, 2e-9, 3e-9]) or y.local_26al_final == pytest.approx([2e-9, 2e-9, 3e-9])
    path = y.marinate(base + "-yields")
    z = pkg.Yields(base)
    z.plate(path)
    assert z.time == y.time and z.sum_local_26al == pytest.approx(y.sum_local_26al) and z.first_write is False
    assert np.allclose(z.local_26al, y.local_26al)


def test_checkpoint_round_trip_and_most_recent(pkg, tmp_path):
    """State checkpoints (al26_nbody.py:347-439): every cluster column, the converter and the metadata come back
    bit-equal; the newest file of a series is found the way the script finds it (:295-318)."""
    U = pkg.units
    ck = pkg.checkpoint
    cl, cv = pkg.driver.init_cluster("plummer", 300, 1.0 | U.pc, seed=2)
    cl.mass_26al_local = np.random.default_rng(0).uniform(0, 1e20, 300) | U.kg
    cl.disk_alive = np.random.default_rng(1).random(300) < 0.5
    base = str(tmp_path / "sim-test")
    md = ck.Metadata(t_f=10.0 | U.Myr, model="plummer", nstars=300, cluster_radius=1.0 | U.pc, filename=base)
    yb = pkg.Yields(base)
    yb.update_state(0.0 | U.Myr, cl)
    ck.save_checkpoint(base, 0, cl, cv, yb, md)
    md.update(0.1 | U.Myr)
    assert md.most_recent_checkpoint == 1 and md.completion == pytest.approx(0.01)
    yb.update_state(0.1 | U.Myr, cl)
    ck.save_checkpoint(base, md.most_recent_checkpoint, cl, cv, yb, md)
    assert ck.most_recent_checkpoint(base) == 1
    cl2, cv2, yb2, md2 = ck.load_checkpoint(base, 1)
    assert set(cl2.attribute_names()) == set(cl.attribute_names()) and np.array_equal(cl2.key, cl.key)
    for name in cl.attribute_names():
        a, b = getattr(cl, name), getattr(cl2, name)
        if hasattr(a, "value_in"):
            assert a.unit == b.unit and np.array_equal(a.number, b.number), name
        else:
            assert np.array_equal(a, b), name
    assert cv2.length_si == cv.length_si and cv2.time_to_nbody(1.0 | U.Myr) == cv.time_to_nbody(1.0 | U.Myr)
    assert float(md2.time.value_in(U.Myr)) == 0.1 and md2.most_recent_checkpoint == 1 and md2.filename == base
    assert yb2.time == yb.time and yb2.sum_local_26al == yb.sum_local_26al
    with pytest.raises(IOError):
        ck.most_recent_checkpoint(str(tmp_path / "no-such-run"))
    with pytest.raises(IOError):
        ck.load_checkpoint(base, 7)
