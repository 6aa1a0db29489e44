"""Synthetic initial conditions for the configurations of BASELINE.json.

Host-side numpy restatements of what the reference gets from AMUSE / numba at start-up:
  maschberger_masses   Maschberger (2013) L3 IMF, mu = 0.2, alpha = 2.3, beta = 1.4 on [0.01, 150] Msun
                       (/root/reference/al26_nbody.py:1375-1446), by INVERSE CDF instead of the
                       reference's uniform-proposal rejection sampler (~1e3x faster), thinned to the density that
                       sampler really accepts, p(m) - p(150 Msun) (KS-tested against the lifted reference sampler:
                       tests/test_cabi_host.py); at least one star >= 13 Msun unless disabled (:1427-1435)
  plummer              Henon-unit Plummer sphere (new_plummer_model, :1520): M = 1, E = -1/4,
                       equal masses 1/N which the caller then overwrites without rescaling (:1530)
  fractal              Goodwin & Whitworth (2004) box fractal of dimension D (new_fractal_cluster_model,
                       :1523-1526), virial ratio 0.5
  disk_lifetimes       tau ~ Exp(mean 2.885 Myr) (:1218-1236)
All seeded through numpy.random.Generator; fp64.
"""
import numpy as np

MU, ALPHA, BETA = 0.2, 2.3, 1.4  # al26_nbody.py:1380-1382


def _maschberger_aux(m):
    return (1.0 + (m / MU) ** (1.0 - ALPHA)) ** (1.0 - BETA)  # al26_nbody.py:1394


def maschberger_pdf(m, m_lower=0.01, m_upper=150.0):
    """the normalised L3 density the reference evaluates (al26_nbody.py:1375-1385)"""
    g_lo, g_hi = _maschberger_aux(m_lower), _maschberger_aux(m_upper)
    A = (((1.0 - ALPHA) * (1.0 - BETA)) / MU) * (1.0 / (g_hi - g_lo))
    return A * ((m / MU) ** (-ALPHA)) * ((1.0 + (m / MU) ** (1.0 - ALPHA)) ** (-BETA))


def maschberger_masses(n, rng, m_lower=0.01, m_upper=150.0, require_massive=True, as_reference=True):
    """n masses by inverse CDF.  as_reference: draw from what the reference's sampler REALLY draws from -- its
    rejection test compares p(m) with a uniform deviate on [p(m_upper), p(m_lower)] (generate_masses, :1418-1421;
    gen_mass_numba, :1405-1407), not on [0, p(m_lower)], so the accepted density is p(m) - p(m_upper): the Maschberger
    law thinned by 1 - p(m_upper)/p(m), which takes out 39 % of the stars at 100 Msun, 8 % at 50 and 0.4 % at 13 Msun
    (1.2e-4 of all stars).  The thinning deviates come from a child generator, so the caller's stream -- the positions
    drawn after the masses -- does not depend on it."""
    g_lo, g_hi = _maschberger_aux(m_lower), _maschberger_aux(m_upper)

    def draw(k, gen):
        g = gen.random(k) * (g_hi - g_lo) + g_lo
        return MU * (g ** (1.0 / (1.0 - BETA)) - 1.0) ** (1.0 / (1.0 - ALPHA))

    thin = rng.spawn(1)[0] if as_reference else None
    p_floor = maschberger_pdf(m_upper, m_lower, m_upper)
    while True:
        m = draw(n, rng)
        if as_reference:
            todo = np.nonzero(thin.random(n) * maschberger_pdf(m, m_lower, m_upper) < p_floor)[0]
            while todo.size:  # redraw the thinned stars until they pass the same test
                m[todo] = draw(todo.size, thin)
                todo = todo[thin.random(todo.size) * maschberger_pdf(m[todo], m_lower, m_upper) < p_floor]
        if not require_massive or m.max() >= 13.0:
            return m


def plummer(n, rng, cutoff_mass_fraction=0.999):
    """Positions / velocities of a Plummer sphere in Henon units (G = M = 1, E = -1/4, r_vir = 1),
    centre of mass at rest at the origin.  Aarseth, Henon & Wielen (1974)."""
    u = rng.random(n) * cutoff_mass_fraction
    u = np.maximum(u, 1e-12)
    r = 1.0 / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    cth = rng.uniform(-1.0, 1.0, n)
    sth = np.sqrt(1.0 - cth * cth)
    ph = rng.uniform(0.0, 2.0 * np.pi, n)
    pos = np.stack([r * sth * np.cos(ph), r * sth * np.sin(ph), r * cth])
    # velocity modulus: rejection on g(q) = q^2 (1 - q^2)^3.5
    q = np.zeros(n)
    todo = np.arange(n)
    while todo.size:
        x = rng.random(todo.size)
        y = rng.random(todo.size) * 0.1
        ok = y < x * x * (1.0 - x * x) ** 3.5
        q[todo[ok]] = x[ok]
        todo = todo[~ok]
    v = q * np.sqrt(2.0) * (1.0 + r * r) ** (-0.25)
    cth = rng.uniform(-1.0, 1.0, n)
    sth = np.sqrt(1.0 - cth * cth)
    ph = rng.uniform(0.0, 2.0 * np.pi, n)
    vel = np.stack([v * sth * np.cos(ph), v * sth * np.sin(ph), v * cth])
    scale = 3.0 * np.pi / 16.0  # structural length -> Henon units
    pos *= scale
    vel /= np.sqrt(scale)
    pos -= pos.mean(axis=1, keepdims=True)
    vel -= vel.mean(axis=1, keepdims=True)
    m = np.full(n, 1.0 / n)
    return m, pos[0], pos[1], pos[2], vel[0], vel[1], vel[2]


def fractal(n, rng, fractal_dimension=1.6, virial_ratio=0.5, potential_energy=None):
    """Box fractal (Goodwin & Whitworth 2004): recursively split a cube into 8, keep each child with
    probability 2^(D-3), jitter, inherit the parent's velocity plus a random component that shrinks
    with the generation; prune to a sphere, pick n stars, scale to M = 1, E_kin/|E_pot| = virial_ratio
    with |E_pot| = 1/2 (Henon units).  `potential_energy(m, x, y, z) -> U` may be supplied (e.g. the
    GPU pair reduction) for large n; default is an O(n^2) numpy loop."""
    p_keep = 2.0 ** (fractal_dimension - 3.0)
    pos = np.zeros((1, 3))
    vel = np.zeros((1, 3))
    size = 2.0
    offs = np.array([[i, j, k] for i in (-1, 1) for j in (-1, 1) for k in (-1, 1)], dtype=np.float64)
    gen = 0
    while True:
        gen += 1
        size *= 0.5
        child_pos = (pos[:, None, :] + 0.5 * size * offs[None, :, :]).reshape(-1, 3)
        child_vel = np.repeat(vel, 8, axis=0)
        keep = rng.random(len(child_pos)) < p_keep
        if not keep.any():
            keep[rng.integers(len(keep))] = True
        child_pos = child_pos[keep] + rng.normal(0.0, 0.1 * size, (keep.sum(), 3))
        child_vel = child_vel[keep] + rng.normal(0.0, 1.0, (keep.sum(), 3)) * size
        pos, vel = child_pos, child_vel
        inside = np.sum(pos * pos, axis=1) < 1.0
        if inside.sum() >= 2 * n or gen > 40:
            pos, vel = pos[inside], vel[inside]
            break
    if len(pos) < n:
        raise RuntimeError("fractal generator produced too few particles")
    pick = rng.choice(len(pos), n, replace=False)
    pos, vel = pos[pick], vel[pick]
    m = np.full(n, 1.0 / n)
    pos -= pos.mean(axis=0)
    vel -= vel.mean(axis=0)
    x, y, z = np.ascontiguousarray(pos.T)
    if potential_energy is None:
        potential_energy = _potential_energy_numpy
    u = potential_energy(m, x, y, z)
    lam = abs(u) / 0.5          # scale lengths so that U = -1/2
    pos *= lam
    k = 0.5 * np.sum(m[:, None] * vel * vel)
    vel *= np.sqrt(virial_ratio * 0.5 / k)
    x, y, z = np.ascontiguousarray(pos.T)
    vx, vy, vz = np.ascontiguousarray(vel.T)
    return m, x, y, z, vx, vy, vz


def _potential_energy_numpy(m, x, y, z):
    u = 0.0
    for i in range(len(m) - 1):
        dx, dy, dz = x[i + 1:] - x[i], y[i + 1:] - y[i], z[i + 1:] - z[i]
        u -= m[i] * np.sum(m[i + 1:] / np.sqrt(dx * dx + dy * dy + dz * dz))
    return u


def disk_lifetimes(n, rng, mean_myr=2.885):
    return rng.exponential(mean_myr, n)  # al26_nbody.py:1233-1235


def cluster(n, seed=0, model="plummer", fractal_dimension=1.6, require_massive=True, potential_energy=None):
    """Masses (Msun) + N-body-unit phase space as the reference builds them (init_cluster, :1492-1530):
    equal-mass model first, then masses overwritten WITHOUT rescaling; N-body mass unit = sum(m)."""
    rng = np.random.default_rng(seed)
    m_msun = maschberger_masses(n, rng, require_massive=require_massive)
    if model == "plummer":
        _, x, y, z, vx, vy, vz = plummer(n, rng)
    elif model == "fractal":
        _, x, y, z, vx, vy, vz = fractal(n, rng, fractal_dimension, potential_energy=potential_energy)
    else:
        raise ValueError('Invalid choice of cluster model, must be either "plummer" or "fractal"!')
    m_nbody = m_msun / m_msun.sum()
    return dict(m_msun=m_msun, m=m_nbody, x=x, y=y, z=z, vx=vx, vy=vy, vz=vz,
                tau_disk_myr=disk_lifetimes(n, rng))
