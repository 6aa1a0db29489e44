"""Stellar-evolution provider interface + a parametrised stub, and the per-massive-star SLR yield
precompute.

The reference drives AMUSE's SeBa (third-party, out of scope: SURVEY section 2) through
`.particles.add_particles`, `.evolve_model(t)`, `.particles.mass`, `.particles.wind_mass_loss_rate`,
`.model_time`, `.stop()` (/root/reference/al26_nbody.py:841,892,947,1730-1731,1737,1763).  `StellarStub`
offers exactly that surface so the hot path can be exercised where AMUSE is absent; with AMUSE present a
maintainer passes the real `SeBa()` object instead -- the driver only touches the members above.

Stub physics: main-sequence lifetime 1e10 yr (M/Msun)^-2.5 (the script's own estimate,
al26_nbody.py:480); stars >= 13 Msun lose `wind_loss_fraction` of their mass at a constant rate until
that age, then mdot = 0 (which the script reads as "has gone supernova", :946-949) and the mass drops
to a 1.4 Msun remnant; lower-mass stars do not evolve.

Yield precompute: `calc_slr_yield` is the reference's Akima interpolation in log10(yield) over the
Limongi & Chieffi (2018) tables (al26_nbody.py:444-465, 572-640), zero outside the tabulated mass range
(so the SN yield is 0 above 25 Msun).  The tables themselves are the reference's data files and are
not copied here: pass their directory (`YieldTables.from_reference_dir`), or use
`YieldTables.synthetic()` (smooth power laws of the tables' magnitude) for synthetic benchmarks.
"""
import os

import numpy as np

from . import units as U
from .particles import Particles


def calc_slr_yield(mass_msun, masses_msun, yields_msun):
    """al26_nbody.py:444-465: Akima1D in log10(yield), 0 outside [min, max] of the table."""
    from scipy.interpolate import Akima1DInterpolator
    masses_msun = np.asarray(masses_msun, dtype=np.float64)
    if len(masses_msun) == 0 or mass_msun < masses_msun.min() or mass_msun > masses_msun.max():
        return 0.0
    interp = Akima1DInterpolator(masses_msun, np.log10(np.asarray(yields_msun, dtype=np.float64)))
    return float(10.0 ** interp(mass_msun))


class YieldTables:
    """wind / SN yield tables (Msun) for Al26 and Fe60: dict iso -> (masses, yields)."""

    def __init__(self, wind, sne):
        self.wind, self.sne = wind, sne

    @staticmethod
    def _read(path, isotopes=("Al26", "Fe60")):
        out = {}
        with open(path) as f:
            masses = [float(c[:-1]) for c in f.readline().strip().split(",")[3:]]  # "13m" -> 13.0  (:610-613)
            for line in f:
                data = line.strip().split(",")
                if data[2] in isotopes:
                    out[data[2]] = (np.array(masses), np.array([float(v) for v in data[3:]]))
        return out

    @classmethod
    def from_reference_dir(cls, directory):
        """directory = <reference>/limongi-chieffi-2018 (wind-yields.csv, sne-yields.csv; v=300 km/s, [Fe/H]=0)."""
        return cls(cls._read(os.path.join(directory, "wind-yields.csv")), cls._read(os.path.join(directory, "sne-yields.csv")))

    @classmethod
    def synthetic(cls):
        mw = np.array([13.0, 15.0, 20.0, 25.0, 30.0, 40.0, 60.0, 80.0, 120.0])
        ms = np.array([13.0, 15.0, 20.0, 25.0])  # the SN table stops at 25 Msun (sne-yields.csv:1)
        wind = {"Al26": (mw, 1.0e-8 * (mw / 13.0) ** 4.0), "Fe60": (mw, 1.0e-11 * (mw / 13.0) ** 5.0)}
        sne = {"Al26": (ms, 3.0e-5 * (ms / 13.0) ** 1.5), "Fe60": (ms, 2.0e-5 * (ms / 13.0) ** 2.0)}
        return cls(wind, sne)

    def star_yields(self, mass_msun, total_wind_loss_msun):
        """wind_ratio_26al/60fe and sn_yield_26al/60fe [Msun] of one massive star (al26_nbody.py:1583-1601)."""
        w26 = calc_slr_yield(mass_msun, *self.wind["Al26"])
        w60 = calc_slr_yield(mass_msun, *self.wind["Fe60"])
        return (w26 / total_wind_loss_msun, w60 / total_wind_loss_msun,   # calc_wind_ratio (:440-441)
                calc_slr_yield(mass_msun, *self.sne["Al26"]), calc_slr_yield(mass_msun, *self.sne["Fe60"]))


def approx_lifespan_myr(mass_msun):
    return 1.0e4 * np.asarray(mass_msun, dtype=np.float64) ** -2.5  # (1e10 yr)(Msun/M)^2.5, al26_nbody.py:480


class _StellarParticles:
    """`stellar.particles`: forwards to the provider's particle set; `add_particles` installs it (:1731)."""

    def __init__(self, owner):
        object.__setattr__(self, "_o", owner)

    def add_particles(self, cluster):
        self._o._install(cluster)
        return self

    def __len__(self):
        return len(self._o._p)

    def __getattr__(self, name):
        return getattr(self._o._p, name)

    def __setattr__(self, name, value):
        setattr(self._o._p, name, value)

    def __getitem__(self, i):
        return self._o._p[i]

    def new_channel_to(self, other):
        from .particles import Channel
        return Channel(self, other, default_attributes=self._o._p.attribute_names())


class StellarStub:
    """SeBa-shaped provider of mass(t) and wind_mass_loss_rate(t)."""

    def __init__(self, wind_loss_fraction=0.2, remnant_msun=1.4, lifetime_factor=1.0, number_of_workers=1):
        self.lifetime_factor = lifetime_factor  # < 1 shortens every life (tests: supernovae within a few steps)
        self.wind_loss_fraction = wind_loss_fraction
        self.remnant_msun = remnant_msun
        self._p = Particles(0)
        self.particles = _StellarParticles(self)
        self.model_time = 0.0 | U.Myr

    def total_wind_loss_msun(self, mass_msun):
        """stand-in for calc_total_mass_loss (al26_nbody.py:467-493)"""
        return self.wind_loss_fraction * np.asarray(mass_msun, dtype=np.float64)

    def _install(self, cluster):
        n = len(cluster)
        p = Particles(n, keys=np.array(cluster.key, copy=True))
        # a checkpointed cluster carries evolved masses; the stub's mass(t) is a function of the ZERO-AGE mass, which
        # driver.init_cluster keeps in an extra column so that a resumed run continues exactly
        m0 = getattr(cluster, "zams_mass", None)
        self._m0 = np.array(U.value_in(m0 if m0 is not None else cluster.mass, U.MSun), dtype=np.float64, copy=True)
        self._life = self.lifetime_factor * approx_lifespan_myr(self._m0)
        self._massive = self._m0 >= 13.0
        self._rate = np.where(self._massive, self.wind_loss_fraction * self._m0 / self._life, 0.0)  # Msun / Myr
        p.mass = self._m0 | U.MSun
        p.wind_mass_loss_rate = -(self._rate * 1.0e-6) | U.msolyr  # SeBa's sign: negative while losing mass (:892)
        self._p = p

    def evolve_model(self, t_end):
        t = float(U.value_in(t_end, U.Myr))
        alive = t < self._life
        m = np.where(self._massive, np.where(alive, self._m0 - self._rate * t, self.remnant_msun), self._m0)
        self._p.mass = m | U.MSun
        self._p.wind_mass_loss_rate = -(np.where(alive, self._rate, 0.0) * 1.0e-6) | U.msolyr
        self.model_time = t | U.Myr

    def stop(self):
        pass
