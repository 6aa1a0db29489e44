"""oracle/enrich_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement (fp64, same evaluation order) of the reference's per-outer-step
short-lived-radionuclide enrichment pass, `/root/reference/al26_nbody.py:878-1086`:

  classify   get_high_mass_star_indices          al26_nbody.py:1194-1216
  wind       calc_wind_abs (numba) x4 + accumulate al26_nbody.py:642-702, :897-938
  supernova  event test + deposit                al26_nbody.py:943-967,
             calc_eta_disk_sne :1291-1334, calc_star_distance :1365-1373
  decay      literal 0.693147, t1/2 0.717 / 2.600 al26_nbody.py:1048-1064
  condense   *_final snapshot + disk_alive flag  al26_nbody.py:1071-1086

Pinned: `calc_wind_abs`, `calc_eta_disk_sne` and the decay fractions are checked
bit-for-bit against the reference's own functions, AST-lifted from the reference file
and executed in the build container (oracle/lift_reference.py ->
tests/golden/enrich_golden.npz), and against the golden vectors of SURVEY.md 8(c).
The supernova / condense / classify loops use AMUSE quantities inline in the
reference and cannot be executed without AMUSE; they are restated line for line.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Units follow the reference kernel call exactly
(al26_nbody.py:886-895,904-905): km, km/s, kg/s, s, kg; masses for the
classification in Msun; tau_disk and t_new in Myr.
"""
import math

import numpy as np

# inventory rows (8 = the reference's per-star SLR attributes, al26_nbody.py:1556-1577)
LOCAL26, GLOBAL26, SNE26, AGB26, LOCAL60, GLOBAL60, SNE60, AGB60 = range(8)
NINV = 8

HALF_LIFE_26AL_MYR = 0.717  # al26_nbody.py:1048
HALF_LIFE_60FE_MYR = 2.600  # al26_nbody.py:1049 (the CSV says 2.62; the code is binding)
LN2_LITERAL = 0.693147      # al26_nbody.py:1050-1051


def decay_fractions(dt_myr):
    """al26_nbody.py:1050-1051: np.exp((-dt*0.693147)/half_life)."""
    f26 = float(np.exp((-dt_myr * LN2_LITERAL) / HALF_LIFE_26AL_MYR))
    f60 = float(np.exp((-dt_myr * LN2_LITERAL) / HALF_LIFE_60FE_MYR))
    return f26, f60


def classify(mass_msun):
    """al26_nbody.py:1208-1216: hm = m >= 13; lm = 0.1 <= m <= 3 (disk_alive NOT tested, :1214)."""
    mass_msun = np.asarray(mass_msun, dtype=np.float64)
    hm = np.nonzero(mass_msun >= 13.0)[0].astype(np.int64)
    lm = np.nonzero((mass_msun >= 0.1) & (mass_msun <= 3.0))[0].astype(np.int64)
    return hm, lm


def calc_wind_abs(lm_id, hm_id, x, y, z, vx, vy, vz, mdot, wind_ratio, rdisk,
                  distance_limit, bubble_radius, dt):
    """Restatement of al26_nbody.py:642-702, vectorised over discs, serial over the
    massive stars in ascending list order so the per-disc sum has the reference's
    order.  x**2 -> x*x, x**3 -> x*(x*x), x**0.5 -> sqrt (numba's static powers)."""
    n = len(x)
    out = np.zeros(n)
    lm_id = np.asarray(lm_id, dtype=np.int64)
    if lm_id.size == 0:
        return out
    lx, ly, lz = x[lm_id], y[lm_id], z[lm_id]
    lvx, lvy, lvz = vx[lm_id], vy[lm_id], vz[lm_id]
    r_disk = rdisk[lm_id]
    disk_spd = np.sqrt(lvx * lvx + lvy * lvy + lvz * lvz)
    d_disk_trav = disk_spd * dt
    r3 = bubble_radius * (bubble_radius * bubble_radius)
    eta_bub = 0.75 * (r_disk * r_disk) * d_disk_trav / r3
    acc = np.zeros(lm_id.size)
    for hm in np.asarray(hm_id, dtype=np.int64):
        wind_abs = wind_ratio[hm] * mdot[hm] * eta_bub * dt
        if distance_limit != 0.0:
            dx, dy, dz = lx - x[hm], ly - y[hm], lz - z[hm]
            d_sep = np.sqrt(dx * dx + dy * dy + dz * dz)
            inside = ~(bubble_radius <= d_sep)
            acc = np.where(inside, acc + wind_abs, acc)
        else:
            acc = acc + wind_abs
    out[lm_id] = acc
    return out


def calc_eta_disk_sne(r, d):
    """al26_nbody.py:1326-1334."""
    cos60 = 0.5
    eta_cond = 0.5
    eta_inj = 0.7
    eta_geom = (cos60 * r ** 2) / (4 * d ** 2)
    return eta_cond * eta_inj * eta_geom


class EnrichState:
    """Per-star state the reference keeps as cluster attributes (al26_nbody.py:1543-1603)."""

    def __init__(self, r_disk_km, tau_disk_myr, disk_alive, kicked, wr26, wr60, sn26, sn60):
        n = len(r_disk_km)
        self.n = n
        self.r_disk = np.array(r_disk_km, dtype=np.float64)
        self.tau_disk = np.array(tau_disk_myr, dtype=np.float64)
        self.disk_alive = np.array(disk_alive, dtype=bool)
        self.kicked = np.array(kicked, dtype=bool)
        self.wr26 = np.array(wr26, dtype=np.float64)
        self.wr60 = np.array(wr60, dtype=np.float64)
        self.sn26 = np.array(sn26, dtype=np.float64)
        self.sn60 = np.array(sn60, dtype=np.float64)
        self.inv = np.zeros((NINV, n))
        self.fin = np.zeros((NINV, n))


def enrich_step(st, mass_msun, mdot, x, y, z, vx, vy, vz, dt_s, t_new_myr,
                r_bub_local_km, r_bub_global_km, decay26, decay60, with_agb=False):
    """One outer step of al26_nbody.py:878-1086 (interloper block excluded).
    `mass_msun` is the mass the reference classifies on (cluster.mass as of the
    previous outer step, :767).  Returns the list of supernova events (indices)."""
    hm, lm = classify(mass_msun)
    # winds :883-938
    if len(hm) > 0:
        g26 = calc_wind_abs(lm, hm, x, y, z, vx, vy, vz, mdot, st.wr26, st.r_disk, 0.0, r_bub_global_km, dt_s)
        g60 = calc_wind_abs(lm, hm, x, y, z, vx, vy, vz, mdot, st.wr60, st.r_disk, 0.0, r_bub_global_km, dt_s)
        l26 = calc_wind_abs(lm, hm, x, y, z, vx, vy, vz, mdot, st.wr26, st.r_disk, r_bub_local_km, r_bub_local_km, dt_s)
        l60 = calc_wind_abs(lm, hm, x, y, z, vx, vy, vz, mdot, st.wr60, st.r_disk, r_bub_local_km, r_bub_local_km, dt_s)
        st.inv[GLOBAL26] += g26
        st.inv[GLOBAL60] += g60
        st.inv[LOCAL26] += l26
        st.inv[LOCAL60] += l60
    # supernovae :945-967
    events = []
    for i in hm:
        if mdot[i] == 0.0 and not st.kicked[i]:
            events.append(int(i))
            if lm.size:
                dx, dy, dz = x[lm] - x[i], y[lm] - y[i], z[lm] - z[i]
                d = np.sqrt(dx * dx + dy * dy + dz * dz)
                r = st.r_disk[lm]
                eta = (0.5 * 0.7) * ((0.5 * (r * r)) / (4.0 * (d * d)))
                st.inv[SNE26, lm] += st.sn26[i] * eta
                st.inv[SNE60, lm] += st.sn60[i] * eta
            st.kicked[i] = True
    # decay :1048-1064 (all N stars)
    for row in (LOCAL26, GLOBAL26, SNE26):
        st.inv[row] *= decay26
    for row in (LOCAL60, GLOBAL60, SNE60):
        st.inv[row] *= decay60
    if with_agb:
        st.inv[AGB26] *= decay26
        st.inv[AGB60] *= decay60
    # condense :1071-1086
    rows = list(range(NINV)) if with_agb else [LOCAL26, GLOBAL26, SNE26, LOCAL60, GLOBAL60, SNE60]
    if lm.size:
        alive = st.disk_alive[lm]
        keep = alive & (st.tau_disk[lm] >= t_new_myr)
        gone = alive & (st.tau_disk[lm] < t_new_myr)
        for row in rows:
            st.fin[row, lm[keep]] = st.inv[row, lm[keep]]
        st.disk_alive[lm[gone]] = False
    return events


def calc_intersection(x1o, y1o, z1o, x1n, y1n, z1n, x2o, y2o, z2o, x2n, y2n, z2n, r, n=1024):
    """al26_nbody.py:1156-1190: fraction of n linspace samples of two straight-line paths closer than r."""
    x1i, y1i, z1i = np.linspace(x1o, x1n, n), np.linspace(y1o, y1n, n), np.linspace(z1o, z1n, n)
    x2i, y2i, z2i = np.linspace(x2o, x2n, n), np.linspace(y2o, y2n, n), np.linspace(z2o, z2n, n)
    ri = ((x1i - x2i) ** 2 + (y1i - y2i) ** 2 + (z1i - z2i) ** 2) ** 0.5
    return np.array(ri <= r).sum() / n


def interloper_step(st, raw, mass_msun, is_interloper, pos_old_pc, pos_new_pc, rate26, rate60, dt, r_bub, km_per_pc,
                    r_test_pc=0.1):
    """AGB interloper deposit, al26_nbody.py:985-1028 (the caller has already established that
    interloper_time > 0 and that a rate is positive).  pos_*_pc: (3, n) positions in pc before / after the
    gravity step; rates in kg/s, dt in s, r_bub (interloper_bubble_radius) and r_disk in km.
    `raw` is the (2, n) mass_{26al,60fe}_agb_raw accumulator (never decayed).  Returns the per-disc fractions."""
    hm, lm = classify(mass_msun)
    k = int(np.nonzero(is_interloper)[0][-1])
    frac = np.zeros(len(mass_msun))
    for i in lm:
        if is_interloper[i]:
            continue
        f = calc_intersection(pos_old_pc[0, k], pos_old_pc[1, k], pos_old_pc[2, k], pos_new_pc[0, k], pos_new_pc[1, k],
                              pos_new_pc[2, k], pos_old_pc[0, i], pos_old_pc[1, i], pos_old_pc[2, i], pos_new_pc[0, i],
                              pos_new_pc[1, i], pos_new_pc[2, i], r_test_pc)
        frac[i] = f
        if f != 0.0:
            dx = (pos_new_pc[:, i] - pos_old_pc[:, i]) * km_per_pc
            d_disk_trav = np.sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2])
            d_disk_trav *= f
            eta_bub = 0.75 * (st.r_disk[i] * st.r_disk[i]) * d_disk_trav / (r_bub * (r_bub * r_bub))
            a26 = rate26 * eta_bub * dt
            a60 = rate60 * eta_bub * dt
            st.inv[AGB26, i] += a26
            st.inv[AGB60, i] += a60
            raw[0, i] += a26
            raw[1, i] += a60
    return frac


def sqrt_threshold(radius):
    """Smallest double q with sqrt(q) >= radius, so that `radius <= sqrt(d2)` <=> `d2 >= q`
    for every double d2 (sqrt is correctly rounded and monotone).  Test helper that
    mirrors the host-side computation the product does for the local-bubble test."""
    q = radius * radius
    while math.sqrt(q) >= radius:
        q = math.nextafter(q, -math.inf)
    while math.sqrt(q) < radius:
        q = math.nextafter(q, math.inf)
    return q
