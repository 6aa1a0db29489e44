"""ctypes binding of libal26b200.so (C-ABI: include/al26_b200.h).

There is NO CPU fallback: if the shared library has not been built, or no sm_100 device is
usable, the calls fail loudly.  Build with `python 26al-nbody_b200/csrc/build.py` (or
`__graft_entry__.build()`).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("AL26_LIB") or os.path.join(_HERE, "csrc", "libal26b200.so")  # AL26_LIB: an instrumented build (csrc/build.py)
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "include", "al26_b200.h"))

_D = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_I32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_U8 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_PD = C.POINTER(C.c_double)
_PI64 = C.POINTER(C.c_int64)
_PINT = C.POINTER(C.c_int)
_VP = C.c_void_p

ERROR_NAMES = {0: "AL26_OK", -1: "AL26_EINVAL", -2: "AL26_ESTATE", -3: "AL26_ETIME", -4: "AL26_ECAP",
               -5: "AL26_ECUDA", -6: "AL26_ENCCL", -7: "AL26_ENODEV"}


class Al26Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


# name -> (restype, argtypes).  Optional (nullable) array arguments are declared c_void_p.
SIGNATURES = {
    "al26_create": (_VP, [C.c_int]),
    "al26_destroy": (None, [_VP]),
    "al26_last_error": (C.c_char_p, [_VP]),
    "al26_version": (C.c_int, []),
    "al26_device_count": (C.c_int, []),
    "al26_device_info": (C.c_int, [_VP, _PINT, _PINT, _PI64, _PI64]),
    "al26_dist_unique_id": (C.c_int, [_VP]),
    "al26_dist_init": (C.c_int, [_VP, C.c_int, C.c_int, _VP]),
    "al26_dist_set_mode": (C.c_int, [_VP, C.c_int]),
    "al26_dist_set_split_min": (C.c_int, [_VP, C.c_int]),
    "al26_dist_p2p_export": (C.c_int, [_VP, _VP]),
    "al26_dist_p2p_import": (C.c_int, [_VP, _VP, C.c_int]),
    "al26_dist_profile": (C.c_int, [_VP, _PI64]),
    "al26_dist_init_local": (C.c_int, [_VP, C.c_int, C.c_int]),
    "al26_dist_p2p_local_slab": (C.c_int, [_VP, C.POINTER(C.c_void_p)]),
    "al26_dist_p2p_attach": (C.c_int, [_VP, C.POINTER(C.c_void_p), _PINT, C.c_int]),
    "al26_group_create": (_VP, [C.c_int, _PINT]),
    "al26_group_destroy": (None, [_VP]),
    "al26_group_last_error": (C.c_char_p, [_VP]),
    "al26_group_size": (C.c_int, [_VP]),
    "al26_group_ctx": (_VP, [_VP, C.c_int]),
    "al26_group_grav_set_params": (C.c_int, [_VP, C.c_double, C.c_double, C.c_double, C.c_double]),
    "al26_group_grav_set_reinit_policy": (C.c_int, [_VP, C.c_int]),
    "al26_group_grav_commit": (C.c_int, [_VP, C.c_int64] + [_D] * 7),
    "al26_group_grav_set_mass": (C.c_int, [_VP, C.c_int64, _D]),
    "al26_group_grav_set_time": (C.c_int, [_VP, C.c_double]),
    "al26_group_grav_get_time": (C.c_int, [_VP, _PD]),
    "al26_group_grav_evolve": (C.c_int, [_VP, C.c_double, _PI64, _PI64]),
    "al26_group_grav_get_state": (C.c_int, [_VP, C.c_int64] + [_D] * 7),
    "al26_group_grav_energies": (C.c_int, [_VP, _PD, _PD, _PD]),
    "al26_group_last_device_ms": (C.c_int, [_VP, _PD, _PI64]),
    "al26_group_enrich_commit": (C.c_int, [_VP, C.c_int64, _D, _D, _U8, _U8, _D, _D, _D, _D]),
    "al26_group_enrich_set_units": (C.c_int, [_VP, C.c_double, C.c_double]),
    "al26_group_enrich_set_mode": (C.c_int, [_VP, C.c_int]),
    "al26_group_enrich_set_inventories": (C.c_int, [_VP, C.c_int64, _VP, _VP]),
    "al26_group_enrich_step": (C.c_int, [_VP, C.c_int64, _D, _D, _VP] + [C.c_double] * 6 + [C.c_int, _I32, C.c_int64, _PI64]),
    "al26_group_enrich_get": (C.c_int, [_VP, C.c_int64, _VP, _VP, _VP, _VP]),
    "al26_grav_set_params": (C.c_int, [_VP, C.c_double, C.c_double, C.c_double, C.c_double]),
    "al26_grav_commit": (C.c_int, [_VP, C.c_int64] + [_D] * 7),
    "al26_grav_set_mass": (C.c_int, [_VP, C.c_int64, _D]),
    "al26_grav_set_reinit_policy": (C.c_int, [_VP, C.c_int]),
    "al26_grav_set_time": (C.c_int, [_VP, C.c_double]),
    "al26_grav_get_time": (C.c_int, [_VP, _PD]),
    "al26_grav_evolve": (C.c_int, [_VP, C.c_double, _PI64, _PI64]),
    "al26_grav_get_state": (C.c_int, [_VP, C.c_int64] + [_D] * 7),
    "al26_grav_energies": (C.c_int, [_VP, _PD, _PD, _PD]),
    "al26_grav_initialize": (C.c_int, [_VP]),
    "al26_grav_get_acc_jerk": (C.c_int, [_VP, C.c_int64] + [_D] * 7),
    "al26_grav_get_timesteps": (C.c_int, [_VP, C.c_int64, _D, _D]),
    "al26_grav_set_timesteps": (C.c_int, [_VP, C.c_int64, _D, _D]),
    "al26_grav_get_active": (C.c_int, [_VP, C.c_int64, _I32, _PI64, _PD]),
    "al26_grav_get_last_active": (C.c_int, [_VP, C.c_int64, _I32, _PI64]),
    "al26_grav_dbg_begin": (C.c_int, [_VP, C.c_double]),
    "al26_grav_dbg_advance": (C.c_int, [_VP, C.c_int64, _PI64, _PINT]),
    "al26_grav_dbg_finish": (C.c_int, [_VP]),
    "al26_grav_dbg_profile_steps": (C.c_int, [_VP, C.c_int, _PD]),
    "al26_grav_force": (C.c_int, [_VP, C.c_int64, C.c_double] + [_D] * 7 + [C.c_int64, _I32] + [_D] * 7),
    "al26_last_device_ms": (C.c_int, [_VP, _PD, _PI64]),
    "al26_enrich_last_kernel_ms": (C.c_int, [_VP, _PD]),
    "al26_grav_bench_force": (C.c_int, [_VP, C.c_int, _PD, _PI64]),
    "al26_grav_bench_force_n": (C.c_int, [_VP, C.c_int64, C.c_int, _PD, _PI64]),
    "al26_dbg_decomposition": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _PI64]),
    "al26_set_force_variant": (C.c_int, [_VP, C.c_int]),
    "al26_set_big_block": (C.c_int, [_VP, C.c_int]),
    "al26_set_decomposition": (C.c_int, [_VP, C.c_int, C.c_double]),
    "al26_set_step_mode": (C.c_int, [_VP, C.c_int]),
    "al26_grav_engine_steps": (C.c_int, [_VP, _PI64, C.POINTER(C.c_int)]),
    "al26_dbg_engine_plan": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "al26_set_chip_max": (C.c_int, [_VP, C.c_int]),
    "al26_grav_chip_steps": (C.c_int, [_VP, _PI64, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "al26_dbg_chip_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), _PI64]),
    "al26_set_fuse_max": (C.c_int, [_VP, C.c_int]),
    "al26_grav_fused_steps": (C.c_int, [_VP, _PI64]),
    "al26_grav_fuse_profile": (C.c_int, [_VP, _PI64]),
    "al26_grav_block_histogram": (C.c_int, [_VP, _PI64]),
    "al26_grav_loop_profile": (C.c_int, [_VP, _PI64]),
    "al26_bench_fp64_peak": (C.c_int, [_VP, _PD]),
    "al26_bench_fp64_with_rsqrt": (C.c_int, [_VP, _PD]),
    "al26_bench_fp64_rate": (C.c_int, [_VP, C.c_int, _PD]),
    "al26_local_densities": (C.c_int, [_VP, C.c_int64, _D, _D, _D, _D, _D]),
    "al26_enrich_commit": (C.c_int, [_VP, C.c_int64, _D, _D, _U8, _U8, _D, _D, _D, _D]),
    "al26_enrich_set_inventories": (C.c_int, [_VP, C.c_int64, _VP, _VP]),
    "al26_enrich_set_units": (C.c_int, [_VP, C.c_double, C.c_double]),
    "al26_enrich_set_mode": (C.c_int, [_VP, C.c_int]),
    "al26_enrich_profile": (C.c_int, [_VP, _PI64]),
    "al26_enrich_step": (C.c_int, [_VP, C.c_int64, _D, _D, _VP] + [C.c_double] * 6 + [C.c_int, _I32, C.c_int64, _PI64]),
    "al26_enrich_interloper": (C.c_int, [_VP, C.c_int64, _D, _D, _D, C.c_int64] + [C.c_double] * 6),
    "al26_enrich_get_agb_raw": (C.c_int, [_VP, C.c_int64, _D]),
    "al26_enrich_get": (C.c_int, [_VP, C.c_int64, _VP, _VP, _VP, _VP]),
}

_lib = None


def load():
    """Load the shared library (no GPU needed for this) and declare every prototype."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} not built: run `python 26al-nbody_b200/csrc/build.py`. "
                "There is no CPU fallback for the B200 hot path.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def device_count():
    return int(load().al26_device_count())


def check(ctx, rc):
    if rc != 0:
        msg = load().al26_last_error(ctx)
        raise Al26Error(rc, msg.decode() if msg else "")


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a):
    """void* of an optional numpy array (None -> NULL)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """Owns one al26_ctx (= one GPU).  Thin, unit-free wrapper used by gravity.py / enrichment.py."""

    def __init__(self, device=0):
        self.L = load()
        h = self.L.al26_create(int(device))
        if not h:
            msg = self.L.al26_last_error(None)
            raise Al26Error(-7, (msg.decode() if msg else "al26_create failed") +
                            " -- the B200 path has no CPU fallback")
        self.h = C.c_void_p(h)
        self.device = int(device)
        self.rank, self.world = 0, 1
        self.dist_mode, self.exchange = None, None

    def close(self):
        if getattr(self, "h", None):
            self.L.al26_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def chk(self, rc):
        check(self.h, rc)

    def device_info(self):
        sm, khz, fr, tot = C.c_int(0), C.c_int(0), C.c_int64(0), C.c_int64(0)
        self.chk(self.L.al26_device_info(self.h, C.byref(sm), C.byref(khz), C.byref(fr), C.byref(tot)))
        return {"sm_count": sm.value, "clock_khz": khz.value, "free_bytes": fr.value, "total_bytes": tot.value}

    def dist_init(self, rank, world, unique_id_bytes, mode="p2p", exchange=None, split_min=0):
        """mode "p2p": peer-memory exchange (default), "nccl": all-gather path.  `exchange(bytes) -> list of
        bytes` all-gathers a 64-byte blob across ranks (dist.py provides one over torch.distributed); it is
        called after every gravity commit in p2p mode."""
        self.chk(self.L.al26_dist_set_mode(self.h, 1 if mode == "p2p" else 0))
        self.chk(self.L.al26_dist_set_split_min(self.h, int(split_min)))
        buf = C.create_string_buffer(bytes(unique_id_bytes), 128) if unique_id_bytes is not None else None
        self.chk(self.L.al26_dist_init(self.h, int(rank), int(world), C.cast(buf, C.c_void_p) if buf else None))
        self.rank, self.world = int(rank), int(world)
        self.dist_mode = mode if world > 1 else None
        self.exchange = exchange

    def p2p_connect(self):
        """after a gravity commit in p2p mode: swap CUDA IPC handles of the staging slabs"""
        if self.world == 1 or self.dist_mode != "p2p":
            return
        if self.exchange is None:
            raise Al26Error(-2, "peer-memory mode needs an exchange function (see dist.init_context)")
        mine = C.create_string_buffer(64)
        self.chk(self.L.al26_dist_p2p_export(self.h, C.cast(mine, C.c_void_p)))
        blobs = self.exchange(mine.raw)
        allh = C.create_string_buffer(b"".join(blobs), 64 * self.world)
        self.chk(self.L.al26_dist_p2p_import(self.h, C.cast(allh, C.c_void_p), self.world))

    def set_force_variant(self, v):
        self.chk(self.L.al26_set_force_variant(self.h, int(v)))

    def set_big_block(self, n_act_min):
        self.chk(self.L.al26_set_big_block(self.h, int(n_act_min)))

    def set_decomposition(self, max_rounds=32, item_overhead_pairs=3200.0):
        self.chk(self.L.al26_set_decomposition(self.h, int(max_rounds), float(item_overhead_pairs)))

    def set_step_mode(self, mode):
        self.chk(self.L.al26_set_step_mode(self.h, int(mode)))

    def engine_steps(self):
        """(block steps taken by the cluster engine since the last commit, its cluster size or 0)"""
        n, cs = C.c_int64(0), C.c_int(0)
        self.chk(self.L.al26_grav_engine_steps(self.h, C.byref(n), C.byref(cs)))
        return n.value, cs.value

    def set_chip_max(self, n_act_max):
        """largest block the chip engine steps (1..256; -1 = default, 0 = engine off)"""
        self.chk(self.L.al26_set_chip_max(self.h, int(n_act_max)))

    def chip_steps(self):
        """(block steps taken by the chip engine since the last commit, its CTAs or 0, its block limit)"""
        n, nc, mx = C.c_int64(0), C.c_int(0), C.c_int(0)
        self.chk(self.L.al26_grav_chip_steps(self.h, C.byref(n), C.byref(nc), C.byref(mx)))
        return n.value, nc.value, mx.value

    def set_fuse_max(self, n_act_max):
        """loop kernels: largest block that takes the fused small-step path (0 = off); before commit"""
        self.chk(self.L.al26_set_fuse_max(self.h, int(n_act_max)))

    def fused_steps(self):
        n = C.c_int64(0)
        self.chk(self.L.al26_grav_fused_steps(self.h, C.byref(n)))
        return n.value

    def fuse_profile(self):
        h = (C.c_int64 * 16)()
        self.chk(self.L.al26_grav_fuse_profile(self.h, h))
        names = ("force_arrive", "wait_partials", "reduce_correct", "-", "wait_release", "fused_total", "scan", "barrier")
        return {k: v for k, v in zip(names, list(h)) if k != "-"}

    def fuse_profile_raw(self):
        h = (C.c_int64 * 16)()
        self.chk(self.L.al26_grav_fuse_profile(self.h, h))
        return list(h)

    def dist_profile(self):
        """peer-memory mode: CTA 0's SM cycles per step category since the last commit (include/al26_b200.h)"""
        h = (C.c_int64 * 12)()
        self.chk(self.L.al26_dist_profile(self.h, h))
        names = ("fused_cycles", "fused_steps", "redundant_cycles", "redundant_steps", "redundant_active",
                 "exch_predict_cycles", "exch_force_cycles", "exch_correct_cycles", "exch_barrier_cycles", "exch_steps",
                 "exch_active", "-")
        return {k: v for k, v in zip(names, list(h)) if k != "-"}

    def block_histogram(self):
        h = (C.c_int64 * 32)()
        self.chk(self.L.al26_grav_block_histogram(self.h, h))
        return list(h)

    def loop_profile(self):
        h = (C.c_int64 * 6)()
        self.chk(self.L.al26_grav_loop_profile(self.h, h))
        return dict(zip(("predict", "bar1", "force", "bar2", "correct", "bar3"), list(h)))

    def fp64_peak_tflops(self):
        tf = C.c_double(0)
        self.chk(self.L.al26_bench_fp64_peak(self.h, C.byref(tf)))
        return tf.value

    def fp64_rate(self, variant):
        """FP64 lane-instructions per second of the operand-pattern microbenchmark `variant` (include/al26_b200.h)"""
        r = C.c_double(0)
        self.chk(self.L.al26_bench_fp64_rate(self.h, int(variant), C.byref(r)))
        return r.value

    def fp64_with_rsqrt_tflops(self):
        """DFMA TFLOP/s left when one independent MUFU.RSQ64H rides along per 32 DFMAs (the force kernel's ratio)"""
        tf = C.c_double(0)
        self.chk(self.L.al26_bench_fp64_with_rsqrt(self.h, C.byref(tf)))
        return tf.value

    def local_densities(self, x_pc, y_pc, z_pc, mass_msun):
        """`local_densities_numba` of the reference's plotting script (plotting/al26_plot.py:324-359) on the GPU."""
        a = [f64(v) for v in (x_pc, y_pc, z_pc, mass_msun)]
        rho = np.zeros(len(a[0]))
        self.chk(self.L.al26_local_densities(self.h, len(rho), *a, rho))
        return rho

    def last_device_ms(self):
        ms, nl = C.c_double(0), C.c_int64(0)
        self.chk(self.L.al26_last_device_ms(self.h, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value


class _GroupRank:
    """One rank's context inside a Group, for the per-context tuning hooks and diagnostics (not owned)."""

    def __init__(self, L, h):
        self.L, self.h = L, C.c_void_p(h)

    def chk(self, rc):
        check(self.h, rc)

    set_fuse_max = Context.set_fuse_max
    set_step_mode = Context.set_step_mode
    dist_profile = Context.dist_profile
    block_histogram = Context.block_histogram
    fused_steps = Context.fused_steps
    device_info = Context.device_info


class Group:
    """Owns one al26_group: n GPUs driven from THIS process, one host thread per GPU inside the library -- the
    counterpart of `number_of_workers=n` behind the reference's single worker object (al26_nbody.py:57,1711-1720).
    Same method names as Context where they make sense; `h` is the group handle, `calls` the group entry points."""

    def __init__(self, n_gpus, device_ids=None):
        self.L = load()
        ids = None if device_ids is None else (C.c_int * int(n_gpus))(*[int(d) for d in device_ids])
        h = self.L.al26_group_create(int(n_gpus), ids)
        if not h:
            msg = self.L.al26_group_last_error(None)
            raise Al26Error(-7, (msg.decode() if msg else "al26_group_create failed") + " -- the B200 path has no CPU fallback")
        self.h = C.c_void_p(h)
        self.n_gpus = int(n_gpus)
        self.rank, self.world = 0, 1  # the group is one logical worker: no per-process rank

    def chk(self, rc):
        if rc != 0:
            msg = self.L.al26_group_last_error(self.h)
            raise Al26Error(rc, msg.decode() if msg else "")

    def rank_ctx(self, r):
        h = self.L.al26_group_ctx(self.h, int(r))
        if not h:
            raise Al26Error(-1, f"rank {r} out of range")
        return _GroupRank(self.L, h)

    def set_chip_max(self, n_act_max):
        """largest block the chip engine steps (1..256; -1 = default, 0 = engine off)"""
        self.chk(self.L.al26_set_chip_max(self.h, int(n_act_max)))

    def chip_steps(self):
        """(block steps taken by the chip engine since the last commit, its CTAs or 0, its block limit)"""
        n, nc, mx = C.c_int64(0), C.c_int(0), C.c_int(0)
        self.chk(self.L.al26_grav_chip_steps(self.h, C.byref(n), C.byref(nc), C.byref(mx)))
        return n.value, nc.value, mx.value

    def set_fuse_max(self, n_act_max):
        for r in range(self.n_gpus):
            self.rank_ctx(r).set_fuse_max(n_act_max)

    def p2p_connect(self):
        pass  # the library wires the peers' slabs itself (al26_group_grav_commit)

    def last_device_ms(self):
        ms, nl = C.c_double(0), C.c_int64(0)
        self.chk(self.L.al26_group_last_device_ms(self.h, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def close(self):
        if getattr(self, "h", None):
            self.L.al26_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def decomposition(n_act, n_tot, sm_count=148, variant=0, big_nact=2048):
    """host-only: the force kernel's work decomposition (dict) for a block of n_act among n_tot particles"""
    out = (C.c_int64 * 8)()
    rc = load().al26_dbg_decomposition(int(n_act), int(n_tot), int(sm_count), int(variant), int(big_nact), out)
    if rc != 0:
        raise Al26Error(rc, "bad arguments")
    return dict(zip(("ipt", "ti", "n_itiles", "n_jsplit", "jchunk", "slot_stride", "part_capacity", "grid"), list(out)))


def chip_plan(n, n_ctas=148, max_smem_per_block=232448):
    """host-only: (particles per CTA or 0, shared-memory bytes per CTA, HBM bytes of mail + partial rows) of the chip engine"""
    p, b, m = C.c_int(0), C.c_int(0), C.c_int64(0)
    rc = load().al26_dbg_chip_plan(int(n), int(n_ctas), int(max_smem_per_block), C.byref(p), C.byref(b), C.byref(m))
    if rc:
        raise Al26Error(rc, "al26_dbg_chip_plan: invalid argument")
    return p.value, b.value, m.value


def engine_plan(n, max_smem_per_block=232448):
    """host-only: (cluster size or 0, particles per CTA, shared-memory bytes per CTA) of the cluster engine for n particles"""
    cs, p, b = C.c_int(0), C.c_int(0), C.c_int(0)
    rc = load().al26_dbg_engine_plan(int(n), int(max_smem_per_block), C.byref(cs), C.byref(p), C.byref(b))
    if rc != 0:
        raise Al26Error(rc, "bad arguments")
    return cs.value, p.value, b.value


def dist_unique_id():
    buf = C.create_string_buffer(128)
    rc = load().al26_dist_unique_id(C.cast(buf, C.c_void_p))
    if rc != 0:
        raise Al26Error(rc, load().al26_last_error(None).decode())
    return buf.raw
