"""Gravity worker: the AMUSE `GravitationalDynamics` surface the reference script drives
(/root/reference/al26_nbody.py:763,782,831,833,871-876,886-891,957,1101,1709-1728,1736,1762;
plotting/al26_plot.py:281-295), backed by the sm_100a Hermite-4 block-timestep kernels through the
C-ABI (include/al26_b200.h).  Two layers:

  GravityCore   unit-free (N-body units, G = 1) numpy-in / numpy-out wrapper of the C entry
                points; what the parity tests and bench.py call.
  B200Gravity   the reference-facing object: `B200Gravity(converter, number_of_workers=k)`,
                `.particles.add_particles`, `.evolve_model(t)`, `.model_time` (get/set),
                `.particles.{x,y,z,vx,vy,vz,mass,key}`, `.particles.copy()`,
                `.particles.new_channel_to(...)`, `.kinetic_energy`, `.potential_energy`,
                `.parameters.{epsilon_squared,timestep_parameter}`, `.stop()`.

No CPU fallback: constructing either without the CUDA library / a B200 raises.
"""
import ctypes as C

import numpy as np

from . import _lib
from . import units as U
from .particles import Particles


class GravityCore:
    """Hermite-4 block-timestep direct N-body on one GPU (or one rank of a multi-GPU job), or -- given a
    `_lib.Group` -- on several GPUs driven from this one process (the reference-surface methods only; the parity
    hooks below need a single context)."""

    def __init__(self, device=0, ctx=None, eps2=0.0, eta=0.14, dt_max=0.125, dt_min=2.0 ** -40):
        self.ctx = ctx if ctx is not None else _lib.Context(device)
        self.L = self.ctx.L
        self.h = self.ctx.h
        self.grouped = isinstance(self.ctx, _lib.Group)
        self.n = 0
        self.set_params(eps2, eta, dt_max, dt_min)

    def _f(self, name):
        """the C entry point `al26_<name>`, or its `al26_group_<name>` twin when this core drives a group"""
        return getattr(self.L, ("al26_group_" if self.grouped else "al26_") + name)

    # -- reference surface ------------------------------------------------------------------
    def set_params(self, eps2=0.0, eta=0.14, dt_max=0.125, dt_min=2.0 ** -40):
        self.ctx.chk(self._f("grav_set_params")(self.h, eps2, eta, dt_max, dt_min))
        self.params = dict(eps2=eps2, eta=eta, dt_max=dt_max, dt_min=dt_min)

    def commit(self, m, x, y, z, vx, vy, vz):
        arrs = [_lib.f64(a) for a in (m, x, y, z, vx, vy, vz)]
        n = len(arrs[0])
        if any(len(a) != n for a in arrs):
            raise ValueError("commit: arrays differ in length")
        self.ctx.chk(self._f("grav_commit")(self.h, n, *arrs))
        self.n = n
        self.ctx.p2p_connect()  # multi-process peer-memory mode: swap the staging slabs' IPC handles

    def set_mass(self, m):
        self.ctx.chk(self._f("grav_set_mass")(self.h, len(m), _lib.f64(m)))

    def set_reinit_policy(self, policy):
        """mass-only update: 0 = recompute forces, keep timesteps (default, ph4's recommit); 1 = forces + initial timesteps"""
        self.ctx.chk(self._f("grav_set_reinit_policy")(self.h, int(policy)))

    def set_time(self, t):
        self.ctx.chk(self._f("grav_set_time")(self.h, float(t)))

    def get_time(self):
        t = C.c_double(0)
        self.ctx.chk(self._f("grav_get_time")(self.h, C.byref(t)))
        return t.value

    def evolve(self, t_end):
        """Advance to t_end (every particle synchronised there).  Returns (block steps, pairs)."""
        ns, npairs = C.c_int64(0), C.c_int64(0)
        self.ctx.chk(self._f("grav_evolve")(self.h, float(t_end), C.byref(ns), C.byref(npairs)))
        return ns.value, npairs.value

    def get_state(self, out=None):
        out = out if out is not None else [np.empty(self.n) for _ in range(7)]
        self.ctx.chk(self._f("grav_get_state")(self.h, self.n, *out))
        return out

    def energies(self):
        k, u, s = C.c_double(0), C.c_double(0), C.c_double(0)
        self.ctx.chk(self._f("grav_energies")(self.h, C.byref(k), C.byref(u), C.byref(s)))
        return k.value, u.value, s.value

    # -- parity hooks -----------------------------------------------------------------------
    def initialize(self):
        self.ctx.chk(self.L.al26_grav_initialize(self.h))

    def get_acc_jerk(self):
        out = [np.zeros(self.n) for _ in range(7)]
        self.ctx.chk(self.L.al26_grav_get_acc_jerk(self.h, self.n, *out))
        return out

    def get_timesteps(self):
        t, dt = np.zeros(self.n), np.zeros(self.n)
        self.ctx.chk(self.L.al26_grav_get_timesteps(self.h, self.n, t, dt))
        return t, dt

    def set_timesteps(self, t, dt):
        self.ctx.chk(self.L.al26_grav_set_timesteps(self.h, self.n, _lib.f64(t), _lib.f64(dt)))

    def get_active(self):
        idx = np.zeros(max(self.n, 1), dtype=np.int32)
        na, tn = C.c_int64(0), C.c_double(0)
        self.ctx.chk(self.L.al26_grav_get_active(self.h, len(idx), idx, C.byref(na), C.byref(tn)))
        return idx[: na.value].copy(), tn.value

    def get_last_active(self):
        idx = np.zeros(max(self.n, 1), dtype=np.int32)
        na = C.c_int64(0)
        self.ctx.chk(self.L.al26_grav_get_last_active(self.h, len(idx), idx, C.byref(na)))
        return idx[: na.value].copy()

    def begin(self, t_end):
        self.ctx.chk(self.L.al26_grav_dbg_begin(self.h, float(t_end)))

    def advance(self, max_steps=-1):
        nd, fin = C.c_int64(0), C.c_int(0)
        self.ctx.chk(self.L.al26_grav_dbg_advance(self.h, int(max_steps), C.byref(nd), C.byref(fin)))
        return nd.value, bool(fin.value)

    def finish(self):
        self.ctx.chk(self.L.al26_grav_dbg_finish(self.h))

    def profile_steps(self, reps=200):
        """(between begin and finish) mean microseconds of the predict, force and correct kernels per block step"""
        us = (C.c_double * 3)()
        self.ctx.chk(self.L.al26_grav_dbg_profile_steps(self.h, int(reps), us))
        return dict(zip(("predict", "force", "correct"), list(us)))

    def force(self, m, x, y, z, vx, vy, vz, idx=None, eps2=0.0):
        """One K1 evaluation on caller arrays: acc(3), jerk(3), pot on the listed particles."""
        arrs = [_lib.f64(a) for a in (m, x, y, z, vx, vy, vz)]
        n = len(arrs[0])
        idx = np.arange(n, dtype=np.int32) if idx is None else np.ascontiguousarray(idx, dtype=np.int32)
        out = [np.zeros(len(idx)) for _ in range(7)]
        self.ctx.chk(self.L.al26_grav_force(self.h, n, float(eps2), *arrs, len(idx), idx, *out))
        return out

    def bench_force(self, reps=3, n_act=0):
        """Average ms of one K1 launch with the first n_act (0: all) particles active, and its pairs."""
        ms, pairs = C.c_double(0), C.c_int64(0)
        self.ctx.chk(self.L.al26_grav_bench_force_n(self.h, int(n_act), int(reps), C.byref(ms), C.byref(pairs)))
        return ms.value, pairs.value

    def last_device_ms(self):
        return self.ctx.last_device_ms()

    def close(self):
        self.ctx.close()


class _Parameters:
    """`gravity.parameters` (never set by the script; ph4 defaults eps2 = 0, eta = 0.14)."""

    def __init__(self, owner):
        object.__setattr__(self, "_o", owner)

    @property
    def epsilon_squared(self):
        o = self._o
        return U.Quantity(o._core.params["eps2"] * o._len_si ** 2, U.m ** 2)

    @epsilon_squared.setter
    def epsilon_squared(self, q):
        o = self._o
        p = dict(o._core.params)
        p["eps2"] = U.value_in(q, U.m ** 2) / o._len_si ** 2
        o._core.set_params(**p)

    @property
    def timestep_parameter(self):
        return self._o._core.params["eta"]

    @timestep_parameter.setter
    def timestep_parameter(self, eta):
        p = dict(self._o._core.params)
        p["eta"] = float(eta)
        self._o._core.set_params(**p)


class _GravityParticles:
    """Live view of the worker's particle set (`gravity.particles`)."""

    _VEC = {"mass": 0, "x": 1, "y": 2, "z": 3, "vx": 4, "vy": 5, "vz": 6}

    def __init__(self, owner):
        self._o = owner
        self.key = np.zeros(0, dtype=np.uint64)
        self.radius = None

    def __len__(self):
        return self._o._core.n

    def add_particles(self, cluster):
        """al26_nbody.py:1728 -- copies key, mass, radius, x..vz; index order preserved."""
        o = self._o
        cv = o.converter
        m = cv.mass_to_nbody(cluster.mass)
        pos = [cv.length_to_nbody(getattr(cluster, a)) for a in ("x", "y", "z")]
        vel = [cv.speed_to_nbody(getattr(cluster, a)) for a in ("vx", "vy", "vz")]
        o._core.commit(m, *pos, *vel)
        o._mtot = float(np.sum(m))
        self.key = np.array(cluster.key, dtype=np.uint64, copy=True)
        self.radius = getattr(cluster, "radius", None)
        o._cache = None
        return self

    def _state(self):
        o = self._o
        if o._cache is None:
            o._cache = o._core.get_state()
        return o._cache

    def __getattr__(self, name):
        vec = _GravityParticles._VEC
        if name in vec:
            o = self._o
            a = self._state()[vec[name]]
            cv = o.converter
            if name == "mass":
                return cv.mass_to_si(a)
            if name in ("x", "y", "z"):
                return cv.length_to_si(a)
            return cv.speed_to_si(a)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name == "mass":
            o = self._o
            m = o.converter.mass_to_nbody(value)
            o._core.set_mass(m)
            o._mtot = float(np.sum(m))
            o._cache = None
            return
        object.__setattr__(self, name, value)

    def __getitem__(self, i):
        return _ParticleView(self, i)

    def copy(self):
        """Detached snapshot (al26_nbody.py:831)."""
        p = Particles(len(self))
        p.key = self.key.copy()
        for a in ("mass", "x", "y", "z", "vx", "vy", "vz"):
            setattr(p, a, getattr(self, a))
        return p

    def new_channel_to(self, other):
        from .particles import Channel
        return Channel(self, other, default_attributes=("mass", "x", "y", "z", "vx", "vy", "vz"))

    def attribute_names(self):
        return ("mass", "x", "y", "z", "vx", "vy", "vz")


class _ParticleView:
    def __init__(self, parent, i):
        self._p, self._i = parent, i

    def __getattr__(self, name):
        if name == "key":
            return self._p.key[self._i]
        return getattr(self._p, name)[self._i]


class B200Gravity:
    """Drop-in for `ph4(converter, number_of_workers=workers)` (al26_nbody.py:1715-1717).

    `number_of_workers` is the number of GPUs this ONE process drives (the reference's worker count, :57): 1 = one
    context on `device`; k > 1 = an `al26_group` of k GPUs (`devices`, default 0..k-1), one host thread per GPU
    inside the library, peer-memory exchange over NVLink.  A job that is already one process per GPU (torchrun)
    passes its joined `ctx` instead (dist.init_context) and leaves number_of_workers at 1.
    """

    def __init__(self, converter, number_of_workers=1, device=0, ctx=None, devices=None, **_ignored):
        self.converter = converter
        self._len_si = converter.length_si
        number_of_workers = int(number_of_workers)
        if number_of_workers < 1:
            raise ValueError("number_of_workers must be >= 1")
        if ctx is None:
            if number_of_workers > 1 and devices is None:
                # the script hard-codes workers = 8 (:57); like an MPI job on a smaller machine, make do with what is there
                have = _lib.device_count()
                if have < 1:
                    raise _lib.Al26Error(-7, "no CUDA device -- the B200 path has no CPU fallback")
                usable = min(number_of_workers, have, 8)
                if usable < number_of_workers:
                    import warnings
                    warnings.warn(f"number_of_workers={number_of_workers} but {have} GPU(s) visible: using {usable}")
                number_of_workers = usable
            ctx = _lib.Group(number_of_workers, devices) if number_of_workers > 1 else _lib.Context(device)
        self._core = GravityCore(ctx=ctx)
        self._cache = None
        self.particles = _GravityParticles(self)
        self.parameters = _Parameters(self)
        self.number_of_workers = number_of_workers

    # model_time get / set (al26_nbody.py:763,1101,1736)
    @property
    def model_time(self):
        return self.converter.time_to_si(self._core.get_time())

    @model_time.setter
    def model_time(self, t):
        self._core.set_time(self.converter.time_to_nbody(t))

    def evolve_model(self, t_end):
        """al26_nbody.py:833 -- on return every particle is at t_end and model_time == t_end."""
        self.last_steps, self.last_pairs = self._core.evolve(self.converter.time_to_nbody(t_end))
        self._cache = None

    @property
    def kinetic_energy(self):
        return self.converter.energy_to_si(self._core.energies()[0])

    @property
    def potential_energy(self):
        return self.converter.energy_to_si(self._core.energies()[1])

    def virial_radius(self):
        """`cluster.virial_radius()` (al26_nbody.py:770) from the same pair reduction."""
        _, _, s = self._core.energies()
        return self.converter.length_to_si(self._mtot * self._mtot / (2.0 * s))

    def stop(self):
        self._core.close()
