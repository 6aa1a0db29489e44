"""A small SoA particle set with channels: the slice of AMUSE's in-memory `Particles` the
reference's hot path relies on (/root/reference/al26_nbody.py:781-783 key check by index,
:831 `.copy()`, :871-876 channels `new_channel_to(...).copy()` / `.copy_attributes([...])`,
:770 `virial_radius()`, :1035 `center_of_mass()`).  Attributes are whole-set vectors (plain numpy
arrays or units.Quantity); `particles[i].attr` reads / writes one element.
"""
import numpy as np

from . import units as U


class Particles:
    def __init__(self, n=0, keys=None):
        object.__setattr__(self, "_n", int(n))
        object.__setattr__(self, "_attrs", {})
        if keys is None:
            # AMUSE keys are random 64-bit integers; index order is what the script relies on
            keys = np.random.default_rng(0xA126).integers(1, 2 ** 63 - 1, size=self._n, dtype=np.int64).astype(np.uint64)
        self._attrs["key"] = np.asarray(keys, dtype=np.uint64)

    def __len__(self):
        return self._n

    def attribute_names(self):
        return tuple(k for k in self._attrs if k != "key")

    def __getattr__(self, name):
        try:
            return object.__getattribute__(self, "_attrs")[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        n = self._n
        if isinstance(value, U.Quantity):
            num = value.number
            if np.ndim(num) == 0:
                num = np.full(n, float(num))
            elif len(num) != n:
                raise ValueError(f"attribute {name}: length {len(num)} != {n}")
            self._attrs[name] = U.Quantity(np.array(num, dtype=np.float64, copy=True), value.unit)
        elif hasattr(value, "value_in"):  # a foreign (AMUSE) quantity: keep as is
            self._attrs[name] = value
        else:
            arr = np.asarray(value)
            if arr.ndim == 0:
                arr = np.full(n, arr[()])
            elif len(arr) != n:
                raise ValueError(f"attribute {name}: length {len(arr)} != {n}")
            self._attrs[name] = np.array(arr, copy=True)

    def __getitem__(self, i):
        return _Particle(self, i)

    def __iter__(self):
        for i in range(self._n):
            yield _Particle(self, i)

    def copy(self):
        p = Particles(self._n, keys=self._attrs["key"].copy())
        for k, v in self._attrs.items():
            if k != "key":
                p._attrs[k] = v.copy()
        return p

    def new_channel_to(self, other):
        return Channel(self, other, default_attributes=self.attribute_names())

    # -- the two set-level reductions the script calls ---------------------------------------
    def center_of_mass(self):
        mass = U.value_in(self.mass, U.kg)
        mt = mass.sum()
        return [U.Quantity(float((mass * U.value_in(getattr(self, a), U.m)).sum() / mt), U.m) for a in ("x", "y", "z")]

    def virial_radius(self):
        """AMUSE particle_attributes.virial_radius: M^2 / (2 sum_{i<j} m_i m_j / r_ij).
        O(N^2) on the host; B200Gravity.virial_radius() is the device path."""
        mass = np.asarray(U.value_in(self.mass, U.kg), dtype=np.float64)
        x, y, z = (np.asarray(U.value_in(getattr(self, a), U.m), dtype=np.float64) for a in ("x", "y", "z"))
        s = 0.0
        for i in range(len(mass) - 1):
            dx, dy, dz = x[i + 1:] - x[i], y[i + 1:] - y[i], z[i + 1:] - z[i]
            s += mass[i] * np.sum(mass[i + 1:] / np.sqrt(dx * dx + dy * dy + dz * dz))
        return U.Quantity(mass.sum() ** 2 / (2.0 * s), U.m)


class _Particle:
    def __init__(self, parent, i):
        object.__setattr__(self, "_p", parent)
        object.__setattr__(self, "_i", i)

    def __getattr__(self, name):
        return getattr(self._p, name)[self._i]

    def __setattr__(self, name, value):
        attrs = self._p._attrs
        if name not in attrs:
            # first write of a new per-particle attribute creates the column (AMUSE behaviour)
            if isinstance(value, U.Quantity):
                attrs[name] = U.Quantity(np.zeros(len(self._p)), value.unit)
            elif isinstance(value, (bool, np.bool_)):
                attrs[name] = np.zeros(len(self._p), dtype=bool)
            else:
                attrs[name] = np.zeros(len(self._p), dtype=np.asarray(value).dtype)
        attrs[name][self._i] = value


class Channel:
    """`a.new_channel_to(b)`: copies attribute vectors by index (sets share index order; the
    reference asserts key equality by index every step, al26_nbody.py:781-783)."""

    def __init__(self, source, target, default_attributes):
        if len(source) != len(target):
            raise ValueError("channel between sets of different length")
        self.source, self.target = source, target
        self.default_attributes = tuple(default_attributes)

    def copy_attributes(self, names):
        for name in names:
            setattr(self.target, name, getattr(self.source, name))

    def copy(self):
        self.copy_attributes(self.default_attributes)
