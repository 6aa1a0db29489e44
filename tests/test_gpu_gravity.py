"""GPU parity: the sm_100a gravity path (through the C-ABI) against the CPU oracle on identical
seeded inputs.  Tolerances (BASELINE.json north_star): acc/jerk <= 1e-12 relative (summation-order
differences only); integer work (active sets, timesteps on the dyadic ladder, step counts) bit-exact."""
import numpy as np
import pytest

from oracle import hermite as H

pytestmark = pytest.mark.gpu

TOL = 1e-12


def vec_rel(a, b):
    """max_i |a_i - b_i| / |b_i| over 3-vectors given as lists of component arrays."""
    a, b = np.stack(a), np.stack(b)
    return np.max(np.linalg.norm(a - b, axis=0) / np.linalg.norm(b, axis=0))


def make(pkg, n, seed, model="plummer"):
    c = pkg.ic.cluster(n, seed=seed, model=model, require_massive=False)
    return [c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")]


@pytest.fixture(params=[(1, 32), (1, 0), (0, 32), (2, 32), (3, 32)], ids=["loop-fused", "loop", "graph", "engine", "chip"])
def grav(pkg, ctx, request):
    """The ways of driving block steps: the persistent cooperative loop kernel with and without its fused
    small-step path, the CUDA graph, and the graph with the cluster engine / the chip engine in front of every block
    step."""
    mode, fuse = request.param
    ctx.set_step_mode(mode)
    ctx.set_fuse_max(fuse)
    g = pkg.GravityCore(ctx=ctx)  # the session shares one context: reset its clock and parameters
    g.set_time(0.0)
    yield g
    ctx.set_step_mode(-1)
    ctx.set_fuse_max(-1)


@pytest.mark.parametrize("n", [2, 3, 33, 257, 1000, 4096])
@pytest.mark.parametrize("eps2", [0.0, 1e-4])
def test_force_parity(pkg, grav, n, eps2):
    p = make(pkg, n, seed=n)
    out = grav.force(*p, eps2=eps2)
    ref = H.force(*p, eps2=eps2, long_double=True)
    assert vec_rel(out[:3], ref[:3]) < TOL
    assert vec_rel(out[3:6], ref[3:6]) < TOL
    assert np.max(np.abs(out[6] - ref[6]) / np.abs(ref[6])) < TOL
    m = p[0]
    assert np.max(np.abs((m * np.stack(out[:3])).sum(axis=1))) < 1e-13 * np.abs(m * np.stack(out[:3])).sum()


def test_force_subset_ragged_and_coincident(pkg, grav):
    p = make(pkg, 3000, seed=5)
    rng = np.random.default_rng(0)
    for k in (1, 7, 31, 32, 33, 64, 65, 700, 2999):
        idx = np.sort(rng.choice(3000, k, replace=False)).astype(np.int32)
        out = grav.force(*p, idx=idx)
        ref = H.force(*p, idx=idx, long_double=True)
        assert vec_rel(out[:3], ref[:3]) < TOL and vec_rel(out[3:6], ref[3:6]) < TOL
    # coincident pair with eps2 = 0 is masked like the self-interaction
    p[1][10], p[2][10], p[3][10] = p[1][11], p[2][11], p[3][11]
    out = grav.force(*p)
    ref = H.force(*p, long_double=True)
    assert np.all(np.isfinite(np.stack(out)))
    assert vec_rel(out[:3], ref[:3]) < TOL
    # large active set crossing the two-i-per-thread threshold
    p = make(pkg, 5000, seed=6)
    out = grav.force(*p)
    ref = H.force(*p, long_double=True)
    assert vec_rel(out[:3], ref[:3]) < TOL and vec_rel(out[3:6], ref[3:6]) < TOL


@pytest.mark.parametrize("variant", range(4))
def test_force_variants_parity(pkg, ctx, variant):
    """every compiled force-kernel configuration, big blocks, ragged blocks and tiny (lane-split) blocks"""
    ctx.set_force_variant(variant)
    try:
        g = pkg.GravityCore(ctx=ctx)
        p = make(pkg, 6000, seed=variant)
        rng = np.random.default_rng(variant)
        for k in (6000, 4500, 100, 16, 9, 5, 2, 1):
            idx = np.sort(rng.choice(6000, k, replace=False)).astype(np.int32)
            out = g.force(*p, idx=idx)
            ref = H.force(*p, idx=idx, long_double=True)
            assert vec_rel(out[:3], ref[:3]) < TOL and vec_rel(out[3:6], ref[3:6]) < TOL, (variant, k)
            assert np.max(np.abs(out[6] - ref[6]) / np.abs(ref[6])) < TOL
    finally:
        ctx.set_force_variant(0)


def test_force_is_deterministic(pkg, grav):
    p = make(pkg, 2500, seed=9)
    a = grav.force(*p)
    b = grav.force(*p)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)


@pytest.mark.parametrize("n,model", [(256, "plummer"), (1000, "plummer"), (512, "fractal")])
def test_initialize_and_stepwise_parity(pkg, grav, n, model):
    p = make(pkg, n, seed=n + 1, model=model)
    o = H.HermiteOracle(n)
    o.commit(*p)
    grav.commit(*p)
    o.initialize(); grav.initialize()
    ga, oa = grav.get_acc_jerk(), o.get_acc_jerk()
    assert vec_rel(ga[:3], oa[:3]) < TOL and vec_rel(ga[3:6], oa[3:6]) < TOL
    assert np.array_equal(grav.get_timesteps()[1], o.get_timesteps()[1])  # initial dt on the ladder: bit-exact
    gi, gt = grav.get_active(); oi, ot = o.get_active()
    assert np.array_equal(gi, oi) and gt == ot
    t_end = 0.02
    o.begin(t_end); grav.begin(t_end)
    for step in range(60):
        oi, ot = o.get_active()
        nd_o, fin_o = o.advance(1)
        nd_g, fin_g = grav.advance(1)
        assert fin_o == fin_g and nd_o == nd_g
        if fin_o:
            break
        assert np.array_equal(grav.get_last_active(), oi), f"active set differs at block step {step}"
        gt_, gdt = grav.get_timesteps(); ot_, odt = o.get_timesteps()
        assert np.array_equal(gt_, ot_) and np.array_equal(gdt, odt), f"ladder differs at block step {step}"
    o.advance(-1); grav.advance(-1)
    o.finish(); grav.finish()
    gs, os_ = grav.get_state(), o.get_state()
    assert np.array_equal(gs[0], os_[0])
    assert vec_rel(gs[1:4], os_[1:4]) < 1e-10 and vec_rel(gs[4:7], os_[4:7]) < 1e-10
    assert grav.get_time() == t_end


@pytest.mark.parametrize("n,model,t_end", [(100_000, "plummer", 2.0 ** -10), (10_000, "fractal", 2.0 ** -6)])
def test_stepwise_parity_at_baseline_sizes(pkg, grav, n, model, t_end):
    """BASELINE configs 3 (N = 1e5 Plummer) and 2 (N = 1e4 fractal D = 1.6, eps = 0) against the ORACLE, block step by
    block step: active sets and the dyadic ladder bit-exact, acc / jerk <= 1e-12, >= 60 block steps, then the rest of
    the call and the synchronisation step."""
    p = make(pkg, n, seed=n + 7, model=model)
    o = H.HermiteOracle(n)
    o.commit(*p)
    grav.commit(*p)
    o.initialize(); grav.initialize()
    ga, oa = grav.get_acc_jerk(), o.get_acc_jerk()
    assert vec_rel(ga[:3], oa[:3]) < TOL and vec_rel(ga[3:6], oa[3:6]) < TOL
    assert np.max(np.abs(ga[6] - oa[6]) / np.abs(oa[6])) < TOL
    assert np.array_equal(grav.get_timesteps()[1], o.get_timesteps()[1])
    o.begin(t_end); grav.begin(t_end)
    compared = 0
    for step in range(64):
        oi, ot = o.get_active()
        nd_o, fin_o = o.advance(1)
        nd_g, fin_g = grav.advance(1)
        assert fin_o == fin_g and nd_o == nd_g
        if fin_o:
            break
        assert np.array_equal(grav.get_last_active(), oi), f"active set differs at block step {step}"
        gt_, gdt = grav.get_timesteps(); ot_, odt = o.get_timesteps()
        assert np.array_equal(gt_, ot_) and np.array_equal(gdt, odt), f"ladder differs at block step {step}"
        compared += 1
    assert compared >= 60
    ga, oa = grav.get_acc_jerk(), o.get_acc_jerk()  # the forces the last correctors stored
    assert vec_rel(ga[:3], oa[:3]) < TOL and vec_rel(ga[3:6], oa[3:6]) < TOL
    o.advance(-1); grav.advance(-1)
    o.finish(); grav.finish()
    assert o.counters()[0] >= 61
    gs, os_ = grav.get_state(), o.get_state()
    if model == "plummer":
        assert vec_rel(gs[1:4], os_[1:4]) < 1e-10 and vec_rel(gs[4:7], os_[4:7]) < 1e-10
        assert np.array_equal(grav.get_timesteps()[1], o.get_timesteps()[1])  # timesteps after the synchronisation step
    else:
        # eps = 0 sub-virial fractal: its hard binaries turn hundreds of orbits within the call and amplify the
        # summation-order differences of the force (1e-15 relative) by many orders of magnitude -- two ph4 runs with
        # different worker counts would differ the same way.  The block-by-block checks above are the parity statement.
        assert vec_rel(gs[1:4], os_[1:4]) < 1e-5
    assert grav.get_time() == t_end


@pytest.mark.parametrize("policy", [0, 1])
def test_mass_update_policies_match_the_oracle(pkg, grav, policy):
    """al26_nbody.py:874 feeds new masses in after every outer step.  Policy 0 (default): forces recomputed, timesteps
    of the synchronisation step kept; policy 1: initial timesteps again.  Both against the oracle, integer work exact."""
    n = 700
    p = make(pkg, n, seed=21)
    o = H.HermiteOracle(n); o.commit(*p); o.set_reinit_policy(policy)
    grav.commit(*p); grav.set_reinit_policy(policy)
    try:
        assert grav.evolve(0.01) == o.evolve(0.01)
        dt_sync = grav.get_timesteps()[1].copy()
        assert np.array_equal(dt_sync, o.get_timesteps()[1])
        a_old = np.stack(grav.get_acc_jerk()[:3])
        m2 = p[0] * np.random.default_rng(2).uniform(0.9, 1.0, n)
        grav.set_mass(m2); o.set_mass(m2)
        grav.initialize(); o.initialize()
        ga, oa = grav.get_acc_jerk(), o.get_acc_jerk()
        assert vec_rel(ga[:3], oa[:3]) < TOL and vec_rel(ga[3:6], oa[3:6]) < TOL
        assert not np.array_equal(np.stack(ga[:3]), a_old)  # the forces did see the new masses
        dt_now = grav.get_timesteps()[1]
        assert np.array_equal(dt_now, o.get_timesteps()[1])
        if policy == 0:
            assert np.array_equal(dt_now, dt_sync)
        else:
            assert np.all(dt_now <= 2.0 ** -5) and not np.array_equal(dt_now, dt_sync)
        w_g, w_o = grav.evolve(0.02), o.evolve(0.02)
        assert w_g == w_o
        gs, os_ = grav.get_state(), o.get_state()
        assert np.array_equal(gs[0], m2) and vec_rel(gs[1:4], os_[1:4]) < 1e-9
    finally:
        grav.set_reinit_policy(0)


def test_softened_potential_excludes_the_self_pair(pkg, grav):
    """eps2 > 0: the masked rsqrt lets the self pair through (r^2 + eps2 > 0); its -m_i/eps is taken out of pot again"""
    rng = np.random.default_rng(4)
    n, eps2 = 5, 0.25
    m = rng.uniform(0.5, 1.5, n)
    x, y, z = rng.normal(size=(3, n))
    v = rng.normal(size=(3, n))
    pot = grav.force(m, x, y, z, *v, eps2=eps2)[6]
    want = np.array([-sum(m[j] / np.sqrt((x[i] - x[j]) ** 2 + (y[i] - y[j]) ** 2 + (z[i] - z[j]) ** 2 + eps2)
                          for j in range(n) if j != i) for i in range(n)])
    assert np.max(np.abs(pot - want) / np.abs(want)) < 1e-14
    g2 = grav
    g2.set_params(eps2=eps2)
    try:
        g2.commit(m, x, y, z, *v)
        g2.initialize()
        assert np.max(np.abs(g2.get_acc_jerk()[6] - want) / np.abs(want)) < 1e-14
        k, u, _ = g2.energies()
        assert u == pytest.approx(0.5 * np.sum(m * want), rel=1e-13)
    finally:
        g2.set_params()


def test_evolve_matches_oracle_counts_and_energy(pkg, grav):
    n = 1000
    p = make(pkg, n, seed=3)
    o = H.HermiteOracle(n); o.commit(*p)
    grav.commit(*p)
    k0, u0, s0 = grav.energies()
    ok0, ou0, os0 = o.energies()
    assert k0 == pytest.approx(ok0, rel=1e-13) and u0 == pytest.approx(ou0, rel=1e-12) and s0 == pytest.approx(os0, rel=1e-12)
    steps_g = pairs_g = steps_o = pairs_o = 0
    for k in range(1, 4):
        a, b = grav.evolve(0.0125 * k)
        c, d = o.evolve(0.0125 * k)
        steps_g += a; pairs_g += b; steps_o += c; pairs_o += d
    assert (steps_g, pairs_g) == (steps_o, pairs_o)  # integer work: bit-exact
    gs, os_ = grav.get_state(), o.get_state()
    assert vec_rel(gs[1:4], os_[1:4]) < 1e-9 and vec_rel(gs[4:7], os_[4:7]) < 1e-9
    k1, u1, _ = grav.energies()
    ok1, ou1, _ = o.energies()
    de_g, de_o = ((k0 + u0) - (k1 + u1)) / (k1 + u1), ((ok0 + ou0) - (ok1 + ou1)) / (ok1 + ou1)
    assert abs(de_g) < 1e-5 and abs(de_g) <= 2.0 * abs(de_o) + 1e-9  # dE/E no worse than the CPU path


def test_fused_small_steps_are_taken_and_agree(pkg, ctx):
    """Loop kernel: small block steps go through the fused path (one barrier, last CTA corrects); the
    trajectory agrees with the three-barrier path to rounding and the integer work is identical."""
    n = 2000
    p = make(pkg, n, seed=11)
    out = {}
    ctx.set_step_mode(1)
    try:
        for fuse in (32, 8, 0):
            ctx.set_fuse_max(fuse)
            g = pkg.GravityCore(ctx=ctx)
            g.set_time(0.0)
            g.commit(*p)
            steps, pairs = g.evolve(0.03125)
            h = ctx.block_histogram()
            out[fuse] = (steps, pairs, g.get_state(), g.get_timesteps()[1], ctx.fused_steps(), sum(h[:5]), sum(h[:6]))
    finally:
        ctx.set_step_mode(-1)
        ctx.set_fuse_max(-1)
    assert out[0][4] == 0 and out[32][4] > 0 and 0 < out[8][4] <= out[32][4]
    assert out[32][5] <= out[32][4] <= out[32][6]  # every block of < 32 particles (log2 bins 0-4), plus n_act == 32
    for fuse in (32, 8):
        assert out[fuse][:2] == out[0][:2]
        assert np.array_equal(out[fuse][3], out[0][3])
        a, b = out[fuse][2], out[0][2]
        assert vec_rel(a[1:4], b[1:4]) < 1e-11 and vec_rel(a[4:7], b[4:7]) < 1e-11


@pytest.mark.parametrize("n,model", [(1000, "plummer"), (3000, "fractal"), (10000, "plummer")])
def test_cluster_engine_takes_the_small_steps_and_agrees(pkg, ctx, n, model):
    """Step mode 2: the runs of small block steps are taken on chip by one thread-block cluster (8 CTAs, or 16 when
    the particles need them); identical integer work, trajectory to rounding, same energy error."""
    p = make(pkg, n, seed=5, model=model)
    out = {}
    try:
        for mode in (2, 0):
            ctx.set_step_mode(mode)
            g = pkg.GravityCore(ctx=ctx, eps2=1e-6 if model == "fractal" else 0.0)
            g.set_time(0.0)
            g.commit(*p)
            e0 = sum(g.energies()[:2])
            work = [g.evolve(0.0078125 * k) for k in range(1, 4)]  # the engine restarts with every call
            h = ctx.block_histogram()
            out[mode] = dict(work=work, state=g.get_state(), dt=g.get_timesteps()[1], hist=h, eng=ctx.engine_steps(),
                             de=(e0 - sum(g.energies()[:2])) / e0)
    finally:
        ctx.set_step_mode(-1)
    n_eng, cs = out[2]["eng"]
    assert cs == (8 if n <= 7000 else 16) and out[0]["eng"] == (0, 0)
    small = sum(out[2]["hist"][:5])  # blocks of < 32 particles (log2 bins 0-4)
    assert small <= n_eng <= sum(out[2]["hist"][:6]) and n_eng > 0
    assert out[2]["work"] == out[0]["work"] and out[2]["hist"] == out[0]["hist"]
    assert np.array_equal(out[2]["dt"], out[0]["dt"])
    a, b = out[2]["state"], out[0]["state"]
    assert vec_rel(a[1:4], b[1:4]) < 1e-10 and vec_rel(a[4:7], b[4:7]) < 1e-10
    assert abs(out[2]["de"] - out[0]["de"]) < 1e-9 + 1e-3 * abs(out[0]["de"])


@pytest.mark.parametrize("n,model,chip_max", [(1000, "plummer", 32), (20000, "plummer", 32), (20000, "fractal", 256),
                                               (100_000, "plummer", 32), (100_000, "plummer", 200)])
def test_chip_engine_takes_the_small_steps_and_agrees(pkg, ctx, n, model, chip_max):
    """Step mode 3: the runs of small block steps are taken by the chip engine (one CTA per SM, the particle set resident
    in shared memory, single-writer mail instead of grid barriers); identical integer work, trajectory to rounding, same
    energy error as the plain graph -- and, where the oracle is affordable, the oracle's integer work."""
    p = make(pkg, n, seed=5, model=model)
    span = 2.0 ** -7 if n <= 20000 else 2.0 ** -10
    out = {}
    try:
        for mode in (3, 0):
            ctx.set_step_mode(mode)
            ctx.set_chip_max(chip_max)
            g = pkg.GravityCore(ctx=ctx, eps2=1e-6 if model == "fractal" else 0.0)
            g.set_time(0.0)
            g.commit(*p)
            e0 = sum(g.energies()[:2])
            work = [g.evolve(span * k) for k in range(1, 4)]  # the engine restarts with every call
            h = ctx.block_histogram()
            out[mode] = dict(work=work, state=g.get_state(), dt=g.get_timesteps()[1], hist=h, chip=ctx.chip_steps(),
                             de=(e0 - sum(g.energies()[:2])) / e0)
    finally:
        ctx.set_step_mode(-1)
        ctx.set_chip_max(-1)
    n_chip, n_ctas, mx = out[3]["chip"]
    assert n_ctas > 0 and mx == chip_max and out[0]["chip"][:2] == (0, 0)
    hist = out[3]["hist"]
    lo = sum(hist[b] for b in range(32) if (2 << b) - 1 <= chip_max)  # bins that lie entirely within the engine's limit
    hi = sum(hist[b] for b in range(32) if (1 << b) <= chip_max)
    assert lo <= n_chip <= hi and n_chip > 0
    assert out[3]["work"] == out[0]["work"] and out[3]["hist"] == out[0]["hist"]
    assert np.array_equal(out[3]["dt"], out[0]["dt"])
    a, b = out[3]["state"], out[0]["state"]
    assert np.array_equal(a[0], b[0])
    assert vec_rel(a[1:4], b[1:4]) < 1e-10 and vec_rel(a[4:7], b[4:7]) < 1e-10
    assert abs(out[3]["de"] - out[0]["de"]) < 1e-9 + 1e-3 * abs(out[0]["de"])
    if n <= 20000 and model == "plummer":
        o = H.HermiteOracle(n); o.commit(*p)
        assert [o.evolve(span * k) for k in range(1, 4)] == out[3]["work"]
        assert np.array_equal(o.get_timesteps()[1], out[3]["dt"])
        os_ = o.get_state()
        assert vec_rel(a[1:4], os_[1:4]) < 1e-10 and vec_rel(a[4:7], os_[4:7]) < 1e-10


def test_set_mass_and_time_setter(pkg, grav):
    n = 300
    p = make(pkg, n, seed=8)
    o = H.HermiteOracle(n); o.commit(*p)
    grav.commit(*p)
    grav.evolve(0.01); o.evolve(0.01)
    m2 = p[0] * np.random.default_rng(1).uniform(0.9, 1.0, n)  # mass loss fed in every outer step (:874)
    grav.set_mass(m2); o.set_mass(m2)
    grav.set_time(7.0); o.set_time(7.0)
    a = grav.evolve(7.01); b = o.evolve(7.01)
    assert a == b and grav.get_time() == 7.01
    gs, os_ = grav.get_state(), o.get_state()
    assert np.array_equal(gs[0], m2)
    assert vec_rel(gs[1:4], os_[1:4]) < 1e-9


def test_kepler_known_answer_on_gpu(pkg, grav):
    from test_oracle_hermite import two_body
    m, *ps = two_body(0.5)
    grav.set_params(eta=0.05)
    grav.commit(m, *ps)
    k0, u0, _ = grav.energies()
    grav.evolve(2.0 * np.pi)
    k1, u1, _ = grav.energies()
    assert abs((k1 + u1) - (k0 + u0)) / abs(k0 + u0) < 2e-7
    st = grav.get_state()
    for a, b in zip(st[1:], ps):
        assert np.max(np.abs(a - b)) < 2e-4
    grav.set_params()


def test_error_codes_and_edge_cases(pkg, ctx):
    g = pkg.GravityCore(ctx=ctx)
    g.set_time(0.0)
    p = make(pkg, 64, seed=1)
    g.commit(*p)
    with pytest.raises(pkg.Al26Error) as ei:
        g.evolve(-1.0)
    assert ei.value.code == -3
    with pytest.raises(pkg.Al26Error) as ei:
        g.set_mass(np.ones(63))
    assert ei.value.code == -1
    with pytest.raises(pkg.Al26Error) as ei:
        g.set_params(eta=-1.0)
    assert ei.value.code == -1
    assert g.evolve(0.0) == (0, 0)  # t_end == model time: nothing to do
    # timesteps live on the dyadic ladder: the hook refuses anything else, and a floor below 2^-300
    g.initialize()
    t, dt = g.get_timesteps()
    g.set_timesteps(t, dt)
    bad = dt.copy(); bad[3] *= 1.5
    with pytest.raises(pkg.Al26Error) as ei:
        g.set_timesteps(t, bad)
    assert ei.value.code == -1
    with pytest.raises(pkg.Al26Error) as ei:
        g.set_params(dt_min=1e-200)
    assert ei.value.code == -1
    with pytest.raises(pkg.Al26Error) as ei:
        ctx.set_step_mode(4)
    assert ei.value.code == -1
    with pytest.raises(pkg.Al26Error) as ei:
        ctx.set_chip_max(257)
    assert ei.value.code == -1
    # a single particle moves on a straight line
    g.commit(np.ones(1), np.zeros(1), np.zeros(1), np.zeros(1), np.array([1.0]), np.zeros(1), np.zeros(1))
    g.set_time(0.0)
    g.evolve(0.25)
    st = g.get_state()
    assert st[1][0] == pytest.approx(0.25, rel=1e-15) and st[4][0] == 1.0


def test_fp64_microbenchmarks(ctx):
    """The roofline denominators: the DFMA-only peak, and the same chains with one independent MUFU.RSQ64H per 32
    DFMAs (the force kernel's ratio) -- the 64-bit rsqrt takes no FP64 issue slots."""
    peak, mixed = ctx.fp64_peak_tflops(), ctx.fp64_with_rsqrt_tflops()
    assert 20.0 < peak < 45.0
    assert 0.95 < mixed / peak < 1.03


def test_full_size_properties_1e5(pkg, grav):
    """BASELINE config 3 size: size-independent properties instead of an O(N^2) CPU reference."""
    n = 100_000
    p = make(pkg, n, seed=42)
    grav.commit(*p)
    grav.initialize()
    a = grav.get_acc_jerk()
    m = p[0]
    for comps in (a[:3], a[3:6]):
        v = m * np.stack(comps)
        assert np.max(np.abs(v.sum(axis=1))) < 1e-12 * np.abs(v).sum()  # sum m a = sum m jerk = 0
    k, u, s = grav.energies()
    assert u == pytest.approx(0.5 * np.sum(m * a[6]), rel=1e-12)       # U = 1/2 sum m_i phi_i
    assert u == pytest.approx(-s, rel=1e-14)
    # a random sample of rows against the long-double oracle
    idx = np.sort(np.random.default_rng(0).choice(n, 64, replace=False)).astype(np.int32)
    ref = H.force(*p, idx=idx, long_double=True)
    assert vec_rel([c[idx] for c in a[:3]], ref[:3]) < TOL and vec_rel([c[idx] for c in a[3:6]], ref[3:6]) < TOL
    steps, pairs = grav.evolve(2.0 ** -12)
    assert steps >= 1 and pairs >= n * n
    t, dt = grav.get_timesteps()
    assert np.all(t == 0.0) and np.all(np.log2(dt) == np.round(np.log2(dt)))
    k1, u1, _ = grav.energies()
    assert abs((k1 + u1) - (k + u)) / abs(k + u) < 1e-8


def test_full_size_properties_1e6(pkg, ctx):
    """BASELINE config 4 size (N = 1e6): one full force evaluation, size-independent checks."""
    n = 1_000_000
    c = pkg.ic.cluster(n, seed=7, require_massive=False)
    p = [c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")]
    g = pkg.GravityCore(ctx=ctx)
    g.set_time(0.0)
    g.commit(*p)
    g.initialize()
    a = g.get_acc_jerk()
    m = p[0]
    for comps in (a[:3], a[3:6]):
        v = m * np.stack(comps)
        assert np.max(np.abs(v.sum(axis=1))) < 1e-11 * np.abs(v).sum()  # sum m a = sum m jerk = 0
    idx = np.sort(np.random.default_rng(1).choice(n, 48, replace=False)).astype(np.int32)
    ref = H.force(*p, idx=idx, long_double=True)
    assert vec_rel([c_[idx] for c_ in a[:3]], ref[:3]) < TOL and vec_rel([c_[idx] for c_ in a[3:6]], ref[3:6]) < TOL
    t, dt = g.get_timesteps()
    assert np.all(np.log2(dt) == np.round(np.log2(dt))) and dt.max() <= 2.0 ** -5


def test_figure_eight_known_answer_on_gpu(pkg, grav):
    """the author's informal three-body check (notes.md:19) as a known answer: period 6.32591398, E = -1.287146"""
    from test_oracle_hermite import figure_eight
    m, *ps = figure_eight()
    grav.set_params(eta=0.02)
    grav.commit(m, *ps)
    k0, u0, _ = grav.energies()
    assert k0 + u0 == pytest.approx(-1.287146, rel=1e-5)
    grav.evolve(6.32591398)
    st = grav.get_state()
    for a, b in zip(st[1:], ps):
        assert np.max(np.abs(a - b)) < 5e-5
    k1, u1, _ = grav.energies()
    assert abs((k1 + u1) - (k0 + u0)) / abs(k0 + u0) < 1e-8
    grav.set_params()
