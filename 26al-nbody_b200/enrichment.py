"""Short-lived-radionuclide enrichment pass: host mirror of the disc routines of the reference's
outer step (/root/reference/al26_nbody.py:878-1086; kernel `calc_wind_abs` :642-702;
`calc_eta_disk_sne` :1291-1334; `get_high_mass_star_indices` :1194-1216), backed by the fused
sm_100a kernel through the C-ABI (include/al26_b200.h).

  EnrichCore       unit-free wrapper (the reference kernel's units: km, km/s, kg/s, s, kg; Msun for
                   the classification; Myr for tau_disk / t_new).  Parity tests and bench.py use it.
  decay_fractions  host-side exp() exactly as al26_nbody.py:1048-1051 (literal 0.693147,
                   half-lives 0.717 / 2.600 Myr).

Inventory rows (AL26_NINV = 8): local26, global26, sne26, agb26, local60, global60, sne60, agb60
= cluster.mass_{26al,60fe}_{local,global,sne,agb} (al26_nbody.py:1556-1577); `fin` holds the
matching *_final columns.  No CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib

NINV = 8
ROWS = ("local26", "global26", "sne26", "agb26", "local60", "global60", "sne60", "agb60")
ROW = {name: i for i, name in enumerate(ROWS)}

HALF_LIFE_26AL_MYR = 0.717   # al26_nbody.py:1048
HALF_LIFE_60FE_MYR = 2.600   # al26_nbody.py:1049
LN2_LITERAL = 0.693147       # al26_nbody.py:1050-1051
R_BUB_LOCAL_WIND_PC = 0.1    # al26_nbody.py:77


def decay_fractions(dt_myr):
    """np.exp((-dt*0.693147)/half_life) for 26Al and 60Fe (al26_nbody.py:1050-1051).  numpy's exp,
    like the reference: it can differ from libm's by 1 ulp depending on the host's SIMD dispatch."""
    return (float(np.exp((-dt_myr * LN2_LITERAL) / HALF_LIFE_26AL_MYR)),
            float(np.exp((-dt_myr * LN2_LITERAL) / HALF_LIFE_60FE_MYR)))


class EnrichCore:
    def __init__(self, device=0, ctx=None):
        self.ctx = ctx if ctx is not None else _lib.Context(device)
        self.L = self.ctx.L
        self.h = self.ctx.h
        self.grouped = isinstance(self.ctx, _lib.Group)  # several GPUs from this process: discs sharded inside the library
        self.n = 0

    def _f(self, name):
        return getattr(self.L, ("al26_group_" if self.grouped else "al26_") + name)

    def commit(self, r_disk_km, tau_disk_myr, disk_alive, kicked, wr26, wr60, sn26_kg, sn60_kg):
        """Per-star attributes of init_cluster (al26_nbody.py:1543-1603); inventories start at 0."""
        n = len(r_disk_km)
        u8 = lambda a: np.ascontiguousarray(np.asarray(a).astype(bool), dtype=np.uint8)
        self.ctx.chk(self._f("enrich_commit")(
            self.h, n, _lib.f64(r_disk_km), _lib.f64(tau_disk_myr), u8(disk_alive), u8(kicked),
            _lib.f64(wr26), _lib.f64(wr60), _lib.f64(sn26_kg), _lib.f64(sn60_kg)))
        self.n = n
        self._sn = np.zeros(8192, dtype=np.int32)

    def set_inventories(self, inv=None, fin=None):
        inv = None if inv is None else _lib.f64(inv)
        fin = None if fin is None else _lib.f64(fin)
        self.ctx.chk(self._f("enrich_set_inventories")(self.h, self.n, _lib.ptr(inv), _lib.ptr(fin)))

    def set_mode(self, mode):
        """0 exact (bit-identical to the reference kernel; default), 1 fast (tolerance 1e-10: hoisted global sum,
        4-instruction pair test), 2 fast + cell-grid pruning (see include/al26_b200.h)"""
        mode = {"exact": 0, "fast": 1, "pruned": 2}.get(mode, mode)
        self.ctx.chk(self._f("enrich_set_mode")(self.h, int(mode)))

    def profile(self):
        """SM cycles of the table-building CTA per phase in the last step (include/al26_b200.h: al26_enrich_profile)"""
        h = (C.c_int64 * 8)()
        self.ctx.chk(self.L.al26_enrich_profile(self.h, h))
        return dict(zip(("sort", "table", "clear", "count", "scan", "scatter"), list(h)[:6]))

    def set_units(self, km_per_length, kms_per_speed):
        self.ctx.chk(self._f("enrich_set_units")(self.h, float(km_per_length), float(kms_per_speed)))

    def step(self, mass_msun, mdot_kg_s, pos_vel, dt_s, t_new_myr, r_bub_local_km, r_bub_global_km,
             decay26, decay60, with_agb=False):
        """One pass of al26_nbody.py:878-1086.  pos_vel: (6, n) array x,y,z [km], vx,vy,vz [km/s],
        or None to read the gravity worker's device state in place.  Returns the ascending list of
        this step's supernova indices."""
        pv = None if pos_vel is None else _lib.f64(pos_vel)
        if pv is not None and pv.shape != (6, self.n):
            raise ValueError(f"pos_vel must have shape (6, {self.n})")
        ne = C.c_int64(0)
        self.ctx.chk(self._f("enrich_step")(
            self.h, self.n, _lib.f64(mass_msun), _lib.f64(mdot_kg_s), _lib.ptr(pv), float(dt_s), float(t_new_myr),
            float(r_bub_local_km), float(r_bub_global_km), float(decay26), float(decay60), int(bool(with_agb)),
            self._sn, len(self._sn), C.byref(ne)))
        return self._sn[: ne.value].copy()

    def interloper(self, mass_msun, pos_old_pc, pos_new_pc, interloper_index, r_bub_km, rate26_kg_s, rate60_kg_s, dt_s,
                   r_test_pc=0.1, km_per_pc=3.08567758128e13):
        """AGB interloper deposit (al26_nbody.py:985-1028); call before step(..., with_agb=True)."""
        po, pn = _lib.f64(pos_old_pc), _lib.f64(pos_new_pc)
        if po.shape != (3, self.n) or pn.shape != (3, self.n):
            raise ValueError(f"positions must have shape (3, {self.n})")
        self.ctx.chk(self.L.al26_enrich_interloper(self.h, self.n, _lib.f64(mass_msun), po, pn, int(interloper_index),
                                                   float(r_test_pc), float(r_bub_km), float(km_per_pc),
                                                   float(rate26_kg_s), float(rate60_kg_s), float(dt_s)))

    def get_agb_raw(self):
        raw = np.zeros((2, self.n))
        self.ctx.chk(self.L.al26_enrich_get_agb_raw(self.h, self.n, raw))
        return raw

    def get(self, want_inv=True, want_fin=True):
        inv = np.zeros((NINV, self.n)) if want_inv else None
        fin = np.zeros((NINV, self.n)) if want_fin else None
        alive = np.zeros(self.n, dtype=np.uint8)
        kicked = np.zeros(self.n, dtype=np.uint8)
        self.ctx.chk(self._f("enrich_get")(self.h, self.n, _lib.ptr(inv), _lib.ptr(fin), _lib.ptr(alive),
                                            _lib.ptr(kicked)))
        return inv, fin, alive.astype(bool), kicked.astype(bool)

    def last_device_ms(self):
        return self.ctx.last_device_ms()

    def last_kernel_ms(self):
        """device time of the three enrichment kernels of the last step (copies excluded)"""
        ms = C.c_double(0)
        self.ctx.chk(self.L.al26_enrich_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def close(self):
        self.ctx.close()
