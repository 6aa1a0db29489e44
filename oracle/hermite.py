"""oracle/hermite.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes binding of oracle/libhermite_oracle.so (built from oracle/hermite_oracle.c by
oracle/Makefile): the CPU restatement of the Hermite-4 block-timestep gravity worker
behind `gravity.evolve_model` (al26_nbody.py:833).  PARITY UNPINNED against AMUSE ph4
(see the header of hermite_oracle.c).  Importable only from tests/, smoke() and
bench.py's CPU legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhermite_oracle.so")
_SO_NATIVE = os.path.join(_HERE, "libhermite_oracle_native.so")
_lib = None
_flavour = "x86-64-v3"

_D = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_I32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "hermite_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def prefer_native():
    """bench.py's CPU legs: build the oracle with -march=native ON THE BOX THAT RUNS IT (the portable
    x86-64-v3 build travels with the snapshot; a native build made elsewhere could SIGILL) and use it.
    Must be called before the first oracle call; falls back to the portable build if gcc fails."""
    global _SO, _flavour
    if _lib is not None:
        return _flavour
    try:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "native"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        if os.path.exists(_SO_NATIVE):
            _SO, _flavour = _SO_NATIVE, "native"
    except Exception:
        pass
    return _flavour


def flavour():
    return _flavour


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_params.argtypes = [C.c_void_p] + [C.c_double] * 4
        L.orc_set_long_double.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_reinit_policy.argtypes = [C.c_void_p, C.c_int]
        L.orc_commit.argtypes = [C.c_void_p, C.c_int64] + [_D] * 7
        L.orc_set_mass.argtypes = [C.c_void_p, C.c_int64, _D]
        L.orc_set_time.argtypes = [C.c_void_p, C.c_double]
        L.orc_get_time.restype = C.c_double
        L.orc_get_time.argtypes = [C.c_void_p]
        L.orc_initialize.argtypes = [C.c_void_p]
        L.orc_begin.argtypes = [C.c_void_p, C.c_double]
        L.orc_advance.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        L.orc_finish.argtypes = [C.c_void_p]
        L.orc_evolve.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_get_state.argtypes = [C.c_void_p, C.c_int64] + [_D] * 7
        L.orc_get_acc_jerk.argtypes = [C.c_void_p, C.c_int64] + [_D] * 7
        L.orc_get_timesteps.argtypes = [C.c_void_p, C.c_int64, _D, _D]
        L.orc_set_timesteps.argtypes = [C.c_void_p, C.c_int64, _D, _D]
        L.orc_get_active.argtypes = [C.c_void_p, C.c_int64, _I32, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        L.orc_energies.argtypes = [C.c_void_p] + [C.POINTER(C.c_double)] * 3
        L.orc_force.argtypes = ([C.c_int64, C.c_double] + [_D] * 7 + [C.c_int64, _I32, C.c_int] + [_D] * 7)
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.restype = None
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_counters.restype = None
        L.orc_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def force(m, x, y, z, vx, vy, vz, idx=None, eps2=0.0, long_double=False):
    """acc, jerk, pot on particles `idx` (default all) by direct summation."""
    n = len(m)
    idx = np.arange(n, dtype=np.int32) if idx is None else np.ascontiguousarray(idx, dtype=np.int32)
    out = [np.zeros(len(idx)) for _ in range(7)]
    lib().orc_force(n, eps2, _c(m), _c(x), _c(y), _c(z), _c(vx), _c(vy), _c(vz), len(idx), idx,
                    int(long_double), *out)
    return out


class HermiteOracle:
    """Same call sequence as the product's C-ABI gravity context (include/al26_b200.h)."""

    def __init__(self, n, eps2=0.0, eta=0.14, dt_max=0.125, dt_min=2.0 ** -40, long_double=False):
        self.L = lib()
        self.n = int(n)
        self.h = C.c_void_p(self.L.orc_create(self.n))
        self._chk(self.L.orc_set_params(self.h, eps2, eta, dt_max, dt_min))
        self.L.orc_set_long_double(self.h, int(long_double))

    @staticmethod
    def _chk(rc):
        if rc != 0:
            raise RuntimeError(f"hermite oracle error {rc}")

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def commit(self, m, x, y, z, vx, vy, vz):
        self._chk(self.L.orc_commit(self.h, self.n, _c(m), _c(x), _c(y), _c(z), _c(vx), _c(vy), _c(vz)))

    def set_mass(self, m):
        self._chk(self.L.orc_set_mass(self.h, self.n, _c(m)))

    def set_reinit_policy(self, policy):
        """mass-only update: 0 = recompute forces, keep timesteps (default); 1 = forces + initial timesteps"""
        self._chk(self.L.orc_set_reinit_policy(self.h, int(policy)))

    def set_time(self, t):
        self.L.orc_set_time(self.h, float(t))

    def get_time(self):
        return self.L.orc_get_time(self.h)

    def initialize(self):
        self._chk(self.L.orc_initialize(self.h))

    def begin(self, t_end):
        self._chk(self.L.orc_begin(self.h, float(t_end)))

    def advance(self, max_steps=-1):
        nd, fin = C.c_int64(0), C.c_int(0)
        self._chk(self.L.orc_advance(self.h, max_steps, C.byref(nd), C.byref(fin)))
        return nd.value, bool(fin.value)

    def finish(self):
        self._chk(self.L.orc_finish(self.h))

    def evolve(self, t_end):
        ns, npairs = C.c_int64(0), C.c_int64(0)
        self._chk(self.L.orc_evolve(self.h, float(t_end), C.byref(ns), C.byref(npairs)))
        return ns.value, npairs.value

    def get_state(self):
        out = [np.zeros(self.n) for _ in range(7)]
        self._chk(self.L.orc_get_state(self.h, self.n, *out))
        return out

    def get_acc_jerk(self):
        out = [np.zeros(self.n) for _ in range(7)]
        self._chk(self.L.orc_get_acc_jerk(self.h, self.n, *out))
        return out

    def get_timesteps(self):
        t, dt = np.zeros(self.n), np.zeros(self.n)
        self._chk(self.L.orc_get_timesteps(self.h, self.n, t, dt))
        return t, dt

    def set_timesteps(self, t, dt):
        self._chk(self.L.orc_set_timesteps(self.h, self.n, _c(t), _c(dt)))

    def get_active(self):
        idx = np.zeros(max(self.n, 1), dtype=np.int32)
        na, tn = C.c_int64(0), C.c_double(0)
        self._chk(self.L.orc_get_active(self.h, self.n, idx, C.byref(na), C.byref(tn)))
        return idx[: na.value].copy(), tn.value

    def counters(self):
        """(block steps, pair evaluations) since creation."""
        s, p = C.c_int64(0), C.c_int64(0)
        self.L.orc_counters(self.h, C.byref(s), C.byref(p))
        return s.value, p.value

    def energies(self):
        k, u, s = C.c_double(0), C.c_double(0), C.c_double(0)
        self._chk(self.L.orc_energies(self.h, C.byref(k), C.byref(u), C.byref(s)))
        return k.value, u.value, s.value


def num_threads():
    return lib().orc_num_threads()


def use_all_cores():
    """OpenMP threads = the cores this process may run on (torchrun sets OMP_NUM_THREADS=1 for its children)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_num_threads(int(n))
    return lib().orc_num_threads()
