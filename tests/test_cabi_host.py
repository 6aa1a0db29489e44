"""CPU: the C-ABI library loads and exports every symbol include/al26_b200.h declares, fails loudly
without a GPU, and the host-side shim logic (units, particles, channels, ICs) behaves."""
import ctypes
import importlib
import os
import re

import numpy as np
import pytest


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.load()
    text = open(pkg.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(al26_[a-z0-9_]+)\s*\(", text))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    lib_mod = importlib.import_module("26al-nbody_b200._lib")
    assert declared == set(lib_mod.SIGNATURES), "ctypes prototypes and header disagree"
    assert L.al26_version() == 200


def test_oracle_is_not_reachable_from_the_product(pkg):
    root = os.path.dirname(pkg.__file__)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "libhermite_oracle" not in src, f


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(pkg):
    with pytest.raises(pkg.Al26Error) as ei:
        pkg.Context(0)
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(pkg.Al26Error):
        pkg.GravityCore()
    with pytest.raises(pkg.Al26Error):
        pkg.EnrichCore()


def test_units_and_converter(pkg):
    U = pkg.units
    assert (1.0 | U.pc).value_in(U.km) == pytest.approx(3.08567758128e13)
    assert (100.0 | U.au).value_in(U.km) == pytest.approx(1.49597870691e10)
    assert (0.01 | U.Myr).value_in(U.s) == pytest.approx(0.01e6 * 365.242199 * 86400)
    q = (3.0 | U.km) * 2.0 + (1.0 | U.m)
    assert q.value_in(U.m) == pytest.approx(6001.0)
    assert ((2.0 | U.kms) * (5.0 | U.s)).value_in(U.km) == pytest.approx(10.0)
    with pytest.raises(ValueError):
        (1.0 | U.kg).value_in(U.m)
    assert (13.0 | U.MSun) >= (13.0 | U.MSun) and not ((2.9 | U.MSun) > (3.0 | U.MSun))
    cv = U.nbody_to_si(1.0 | U.pc, 339.0 | U.MSun)
    assert cv.time_si / (1e6 * U.YR_S) == pytest.approx(0.81, rel=0.02)  # ~0.81 Myr per N-body time unit
    x = cv.length_to_si(np.array([1.0, 2.0]))
    assert np.allclose(cv.length_to_nbody(x), [1.0, 2.0])
    assert cv.time_to_nbody(cv.time_to_si(0.125)) == pytest.approx(0.125)
    cv2 = U.nbody_to_si(339.0 | U.MSun, 1.0 | U.pc)  # argument order is free, as in AMUSE
    assert cv2.time_si == cv.time_si


def test_particles_and_channels(pkg):
    U = pkg.units
    P = pkg.Particles
    a = P(5)
    a.mass = np.arange(1.0, 6.0) | U.MSun
    a.x = np.zeros(5) | U.pc
    a.flag = np.zeros(5, dtype=bool)
    b = P(5, keys=a.key.copy())
    b.mass = np.zeros(5) | U.MSun
    a.new_channel_to(b).copy_attributes(["mass"])
    assert np.array_equal(b.mass.value_in(U.MSun), a.mass.value_in(U.MSun))
    b.mass[2] = 10.0 | U.MSun
    assert a.mass.value_in(U.MSun)[2] == 3.0  # copies, not views
    a[1].flag = True
    a[3].new_attr = 7.5 | U.kg
    assert a.flag.tolist() == [False, True, False, False, False] and a.new_attr.value_in(U.kg)[3] == 7.5
    c = a.copy()
    c.x = np.ones(5) | U.pc
    assert a.x.value_in(U.pc)[0] == 0.0
    assert all(a[i].key == b[i].key for i in range(5))
    with pytest.raises(ValueError):
        a.new_channel_to(P(4)).copy()
    # virial radius of an equal-mass pair at separation d is 2 d ... M^2/(2 m1 m2/d) = (2m)^2 d/(2 m^2) = 2d
    p = P(2)
    p.mass = np.ones(2) | U.kg
    p.x = np.array([0.0, 3.0]) | U.m; p.y = np.zeros(2) | U.m; p.z = np.zeros(2) | U.m
    assert p.virial_radius().value_in(U.m) == pytest.approx(6.0)


def test_initial_conditions(pkg):
    ic = pkg.ic
    rng = np.random.default_rng(0)
    m = ic.maschberger_masses(200000, rng)
    assert 0.01 <= m.min() and m.max() <= 150.0 and m.max() >= 13.0
    # SURVEY section 6 (reference sampler): 0.186 % >= 13 Msun, 48.2 % in [0.1, 3], mean 0.339 Msun
    assert np.mean(m >= 13.0) == pytest.approx(0.00186, rel=0.2)
    assert np.mean((m >= 0.1) & (m <= 3.0)) == pytest.approx(0.482, rel=0.03)
    assert m.mean() == pytest.approx(0.339, rel=0.1)
    mm, x, y, z, vx, vy, vz = ic.plummer(20000, np.random.default_rng(1))
    assert abs(x.mean()) < 1e-12 and abs(vx.mean()) < 1e-12 and mm.sum() == pytest.approx(1.0)
    k = 0.5 * np.sum(mm * (vx * vx + vy * vy + vz * vz))
    assert k == pytest.approx(0.25, rel=0.05)
    f = ic.fractal(300, np.random.default_rng(2), 1.6)
    kf = 0.5 * np.sum(f[0] * (f[4] ** 2 + f[5] ** 2 + f[6] ** 2))
    uf = ic._potential_energy_numpy(*f[:4])
    assert uf == pytest.approx(-0.5, rel=1e-9) and kf / abs(uf) == pytest.approx(0.5, rel=1e-9)
    tau = ic.disk_lifetimes(100000, np.random.default_rng(3))
    assert tau.mean() == pytest.approx(2.885, rel=0.02)
    c = ic.cluster(500, seed=3)
    assert c["m"].sum() == pytest.approx(1.0) and len(c["x"]) == 500
    with pytest.raises(ValueError):
        ic.cluster(10, model="king")


def test_imf_sampler_against_the_lifted_reference_sampler(pkg):
    """`ic.maschberger_masses` (inverse CDF) pinned to the reference's own rejection sampler
    (al26_nbody.py:1375-1410), lifted from the reference file when it is mounted: same density function, and what the
    sampler accepts is p(m) - p(m_upper) -- the inverse-CDF draw thinned accordingly -- by a two-sample KS test, an
    analytic one-sample KS test, and the thinning's signature at the top end."""
    from scipy import stats
    ic = pkg.ic
    m_lo, m_hi = 0.01, 150.0
    g_lo, g_hi = ic._maschberger_aux(m_lo), ic._maschberger_aux(m_hi)
    p_hi = ic.maschberger_pdf(m_hi)
    # analytic CDF of the accepted density q(m) = (p(m) - p_hi) / (1 - p_hi (m_hi - m_lo))
    cdf = lambda m: (((ic._maschberger_aux(m) - g_lo) / (g_hi - g_lo)) - p_hi * (m - m_lo)) / (1.0 - p_hi * (m_hi - m_lo))
    mine = ic.maschberger_masses(400000, np.random.default_rng(7), require_massive=False)
    assert stats.kstest(mine, cdf).pvalue > 1e-3
    # the thinning is visible where it matters: stars above 100 Msun are ~45 % rarer than pure Maschberger would make them
    pure = ic.maschberger_masses(4_000_000, np.random.default_rng(8), require_massive=False, as_reference=False)
    thinned = ic.maschberger_masses(4_000_000, np.random.default_rng(8), require_massive=False)
    ratio = np.sum(thinned > 100.0) / np.sum(pure > 100.0)
    grid = np.linspace(100.0, 150.0, 20001)
    pm = ic.maschberger_pdf(grid)
    want = np.sum(pm - p_hi) / np.sum(pm)
    assert 0.35 < want < 0.6 and ratio == pytest.approx(want, abs=0.12)
    assert np.array_equal(pure[:1000] == thinned[:1000], np.ones(1000, bool)) or np.mean(pure == thinned) > 0.999  # same parent stream
    lift = pytest.importorskip("oracle.lift_reference")
    if not lift.available():
        pytest.skip("reference file not mounted")
    ref = lift.lift(names=("maschberger", "gen_mass_numba"))
    mg = 10.0 ** np.linspace(-2, np.log10(150.0), 200)
    ref_pdf = np.array([ref["maschberger"](float(v), g_lo, g_hi) for v in mg])
    assert np.max(np.abs(ref_pdf / ic.maschberger_pdf(mg) - 1.0)) < 1e-13
    theirs = ref["gen_mass_numba"](g_lo, g_hi, p_hi, ic.maschberger_pdf(m_lo), m_lo, m_hi, 30000)
    assert stats.ks_2samp(theirs, mine[:200000]).pvalue > 1e-4
    assert stats.kstest(theirs, cdf).pvalue > 1e-4


def test_decay_fractions_literal(pkg):
    from oracle import enrich_oracle as eo
    f26, f60 = pkg.decay_fractions(0.01)
    assert (f26, f60) == eo.decay_fractions(0.01)  # same call as the reference (np.exp)
    # SURVEY 8(c) golden; np.exp may differ by 1 ulp between SIMD dispatch paths
    assert f26 == pytest.approx(0.99037925616650468, rel=3e-16) and f60 == pytest.approx(0.99733760048885856, rel=3e-16)


def test_force_work_decomposition_invariants(pkg):
    """host-side check of the table the force and corrector kernels share (al26_internal.cuh: choose_jsplit,
    make_decomp): every block size maps to a decomposition that covers all i and all j, fits the partial
    buffers, fills the grid when there is enough work, and never cuts items finer than a TMA-friendly chunk."""
    lib = importlib.import_module("26al-nbody_b200._lib")
    rng = np.random.default_rng(0)
    for variant in range(4):
        for n_tot in (1, 31, 1000, 10_000, 100_000, 1_000_000):
            acts = sorted({1, 2, 16, 17, 32, 33, 2047, 2048, 2049, n_tot, max(1, n_tot // 2), max(1, n_tot // 3)}
                          | set(int(v) for v in rng.integers(1, n_tot + 1, 12)))
            for n_act in acts:
                if n_act > n_tot:
                    continue
                d = lib.decomposition(n_act, n_tot, variant=variant)
                assert d["ti"] == 32 * d["ipt"] and d["ipt"] >= 1
                assert d["n_itiles"] * d["ti"] >= n_act > (d["n_itiles"] - 1) * d["ti"]      # all i, no empty tile
                assert d["n_jsplit"] * d["jchunk"] >= n_tot > (d["n_jsplit"] - 1) * d["jchunk"]  # all j, no empty chunk
                assert d["jchunk"] % 8 == 0
                assert d["slot_stride"] == d["n_itiles"] * d["ti"]
                assert d["n_jsplit"] * d["slot_stride"] <= d["part_capacity"]
                items = d["n_itiles"] * d["n_jsplit"]
                if n_act * n_tot >= 64 * 64 * d["grid"]:
                    assert items >= 0.5 * d["grid"], (variant, n_tot, n_act, d)              # the grid is used
                assert items <= 64 * d["grid"] + d["n_itiles"]
                if d["ipt"] > 1:
                    assert n_act >= 2048
    with pytest.raises(lib.Al26Error):
        lib.decomposition(5, 3)


def test_cluster_engine_plan(pkg):
    """host-side check of the cluster engine's capacity plan (hermite_engine.cu: engine_plan): the smallest cluster
    whose CTAs hold their share of the particles in shared memory; 8 CTAs while that fits, 16 up to ~1.3e4 particles."""
    lib = importlib.import_module("26al-nbody_b200._lib")
    last_cs = 8
    for n in (1, 2, 7, 8, 9, 1000, 3000, 6000, 6592, 6593, 7000, 10_000, 13_000, 13_184, 13_185, 14_000, 20_000, 100_000):
        cs, p, b = lib.engine_plan(n)
        if cs == 0:
            assert (p, b) == (0, 0) and n > 13_184
            last_cs = 99
            continue
        assert cs in (8, 16) and cs >= last_cs            # monotone in n
        last_cs = cs
        assert p % 8 == 0 and cs * p >= n > cs * (p - 8) - cs  # every particle has a home, no CTA over-sized
        assert b <= 232448 and b == lib.engine_plan(cs * p)[2]
        if cs == 16:  # 8 CTAs really were too few
            assert n > 6592
    assert lib.engine_plan(10_000)[0] == 16 and lib.engine_plan(1000)[0] == 8   # BASELINE configs 2 and 1
    assert lib.engine_plan(10_000, 100_000)[0] == 0      # a device with less shared memory: no engine
    with pytest.raises(lib.Al26Error):
        lib.engine_plan(0)


def test_chip_engine_plan(pkg):
    """host-side check of the chip engine's capacity plan (hermite_chip.cu: chip_plan): contiguous chunks of ceil(n / CTAs)
    particles, 208 B each behind the fixed part, within the shared memory a block may opt in to"""
    from importlib import import_module
    lib = import_module("26al-nbody_b200._lib")
    p, b, m = lib.chip_plan(100_000)                      # BASELINE config 3 on the 148 SMs of a B200
    assert p == 680 and 148 * p >= 100_000 and b <= 232448
    fixed = b - 208 * p
    assert 0 < fixed < 80 * 1024 and m > 0
    for n in (1, 2, 147, 148, 149, 4096, 20_000, 110_000):
        p, b, _ = lib.chip_plan(n)
        assert p % 8 == 0 and 148 * p >= n and (p - 8) * 148 < n + 8 * 148 and b == fixed + 208 * p
    assert lib.chip_plan(1_000_000)[0] == 0                # config 4 does not fit the chip: the grid-wide kernels take it
    assert lib.chip_plan(100_000, 148, 150_000)[0] == 0   # a device with less shared memory per block
    assert lib.chip_plan(100_000, 300)[0] == 0            # more CTAs than the scheduling warp handles
    with pytest.raises(pkg.Al26Error):
        lib.chip_plan(0)
