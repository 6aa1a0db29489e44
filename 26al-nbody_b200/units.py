"""Minimal unit algebra + N-body converter for the gravity / enrichment boundary.

Stands in for the slice of `amuse.units` the reference script touches on the hot path
(`units.km`, `units.kg/units.s`, `| myr`, `.value_in(...)`, `nbody_system.nbody_to_si(Rc, Mcluster)`,
/root/reference/al26_nbody.py:19-24,62-74,886-895,1516) so the drop-in can be exercised where AMUSE
is not installed.  Constants follow AMUSE 2023.5 (not CODATA/IAU): they differ in the 4th-6th digit
and only matter for SI <-> N-body conversion; every parity test compares in kernel units.
"""
import math
from fractions import Fraction

import numpy as np

G_SI = 6.67428e-11            # m^3 kg^-1 s^-2 (amuse.units.constants.G)
PARSEC_M = 3.08567758128e16
AU_M = 149597870691.0
MSUN_KG = 1.98892e30
DAY_S = 86400.0
YR_S = 365.242199 * DAY_S     # amuse.units.units.yr


class Unit:
    __slots__ = ("factor", "dims", "name")
    __array_ufunc__ = None  # `ndarray | unit` and `ndarray * unit` defer to __ror__ / __rmul__

    def __init__(self, factor, dims, name=""):
        self.factor = float(factor)
        self.dims = tuple(Fraction(d) for d in dims)  # (length, mass, time)
        self.name = name

    def __mul__(self, o):
        if isinstance(o, Unit):
            return Unit(self.factor * o.factor, [a + b for a, b in zip(self.dims, o.dims)], f"{self.name}*{o.name}")
        return Quantity(o, self)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Unit):
            return Unit(self.factor / o.factor, [a - b for a, b in zip(self.dims, o.dims)], f"{self.name}/{o.name}")
        return NotImplemented

    def __pow__(self, p):
        p = Fraction(p).limit_denominator(12)
        return Unit(self.factor ** float(p), [a * p for a in self.dims], f"{self.name}**{p}")

    def __ror__(self, number):  # `1.0 | units.Myr`
        return Quantity(number, self)

    def __call__(self, number):
        return Quantity(number, self)

    def __repr__(self):
        return self.name or f"Unit({self.factor}, {self.dims})"

    def __eq__(self, o):  # value semantics: a unit that went through a pickle (checkpoint.py) equals the original
        return isinstance(o, Unit) and self.factor == o.factor and self.dims == o.dims

    def __hash__(self):
        return hash((self.factor, self.dims))


class Quantity:
    """A number (scalar or ndarray) with a unit."""
    __array_ufunc__ = None  # ndarray (op) Quantity defers to the reflected operator here

    def __init__(self, number, unit):
        self.number = np.asarray(number, dtype=np.float64) if not np.isscalar(number) else float(number)
        self.unit = unit

    def value_in(self, unit):
        if unit.dims != self.unit.dims:
            raise ValueError(f"incompatible units {self.unit} -> {unit}")
        if unit.factor == self.unit.factor:
            return self.number
        return self.number * (self.unit.factor / unit.factor)

    def in_(self, unit):
        return Quantity(self.value_in(unit), unit)

    def _si(self):
        return self.number * self.unit.factor

    def __add__(self, o):
        return Quantity(self.number + o.value_in(self.unit), self.unit)

    def __sub__(self, o):
        return Quantity(self.number - o.value_in(self.unit), self.unit)

    def __neg__(self):
        return Quantity(-self.number, self.unit)

    def __mul__(self, o):
        if isinstance(o, Quantity):
            return Quantity(self.number * o.number, self.unit * o.unit)
        if isinstance(o, Unit):
            return Quantity(self.number, self.unit * o)
        return Quantity(self.number * o, self.unit)

    __rmul__ = __mul__

    def __radd__(self, o):
        return self.__add__(o)

    def __rtruediv__(self, o):
        return Quantity(o / self.number, self.unit ** -1)

    def __truediv__(self, o):
        if isinstance(o, Quantity):
            u = self.unit / o.unit
            if all(d == 0 for d in u.dims):
                return (self.number / o.number) * u.factor
            return Quantity(self.number / o.number, u)
        return Quantity(self.number / o, self.unit)

    def __pow__(self, p):
        return Quantity(self.number ** p, self.unit ** p)

    def _cmp(self, o):
        return self.number, (o.value_in(self.unit) if isinstance(o, Quantity) else o)

    def __lt__(self, o): a, b = self._cmp(o); return a < b
    def __le__(self, o): a, b = self._cmp(o); return a <= b
    def __gt__(self, o): a, b = self._cmp(o); return a > b
    def __ge__(self, o): a, b = self._cmp(o); return a >= b
    def __eq__(self, o): a, b = self._cmp(o); return a == b
    def __ne__(self, o): a, b = self._cmp(o); return a != b
    __hash__ = None

    def __len__(self):
        return len(self.number)

    def __getitem__(self, i):
        return Quantity(self.number[i], self.unit)

    def __setitem__(self, i, v):
        self.number[i] = v.value_in(self.unit) if isinstance(v, Quantity) else v

    def sum(self):
        return Quantity(np.sum(self.number), self.unit)

    def copy(self):
        return Quantity(np.copy(self.number), self.unit)

    def sqrt(self):
        return self ** Fraction(1, 2)

    def __repr__(self):
        return f"{self.number} {self.unit}"


none = Unit(1.0, (0, 0, 0), "none")
m = Unit(1.0, (1, 0, 0), "m")
kg = Unit(1.0, (0, 1, 0), "kg")
s = Unit(1.0, (0, 0, 1), "s")
km = Unit(1.0e3, (1, 0, 0), "km")
au = AU = Unit(AU_M, (1, 0, 0), "au")
pc = parsec = Unit(PARSEC_M, (1, 0, 0), "parsec")
MSun = msol = Unit(MSUN_KG, (0, 1, 0), "MSun")
day = Unit(DAY_S, (0, 0, 1), "day")
yr = Unit(YR_S, (0, 0, 1), "yr")
Myr = myr = Unit(1.0e6 * YR_S, (0, 0, 1), "Myr")
kms = Unit(1.0e3, (1, 0, -1), "km/s")
J = Unit(1.0, (2, 1, -2), "J")
msolyr = Unit(MSUN_KG / YR_S, (0, 1, -1), "MSun/yr")


def value_in(q, unit):
    """`q.value_in(unit)` for unit-bearing q (ours or AMUSE's); plain numbers pass through."""
    if hasattr(q, "value_in"):
        try:
            return q.value_in(unit)
        except (TypeError, AttributeError):
            # an AMUSE quantity asked for one of OUR units: go through SI numbers
            return _amuse_value_in(q, unit)
    return q


def _amuse_value_in(q, unit):  # pragma: no cover - needs AMUSE
    from amuse.units import units as au_
    base = (au_.m ** float(unit.dims[0])) * (au_.kg ** float(unit.dims[1])) * (au_.s ** float(unit.dims[2]))
    return q.value_in(base) / unit.factor


class nbody_to_si:
    """`nbody_system.nbody_to_si(length, mass)` (al26_nbody.py:1516): G = 1 units."""

    def __init__(self, length, mass):
        a, b = length, mass
        if isinstance(a, Quantity) and a.unit.dims == kg.dims:
            a, b = b, a
        self.length_si = float(value_in(a, m))
        self.mass_si = float(value_in(b, kg))
        self.time_si = math.sqrt(self.length_si ** 3 / (G_SI * self.mass_si))
        self.speed_si = self.length_si / self.time_si
        self.energy_si = self.mass_si * self.speed_si ** 2

    # numbers in N-body units <- quantities
    def length_to_nbody(self, q): return np.asarray(value_in(q, m)) / self.length_si
    def mass_to_nbody(self, q): return np.asarray(value_in(q, kg)) / self.mass_si
    def time_to_nbody(self, q): return float(value_in(q, s)) / self.time_si
    def speed_to_nbody(self, q): return np.asarray(value_in(q, m / s)) / self.speed_si

    # quantities <- numbers in N-body units
    def length_to_si(self, x): return Quantity(np.asarray(x) * self.length_si, m)
    def mass_to_si(self, x): return Quantity(np.asarray(x) * self.mass_si, kg)
    def time_to_si(self, x): return Quantity(x * self.time_si, s)
    def speed_to_si(self, x): return Quantity(np.asarray(x) * self.speed_si, m / s)
    def energy_to_si(self, x): return Quantity(x * self.energy_si, J)

    # factors the enrichment kernel applies when it reads the gravity state in place
    @property
    def km_per_length(self): return self.length_si / 1.0e3
    @property
    def kms_per_speed(self): return self.speed_si / 1.0e3
