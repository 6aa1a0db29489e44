#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (BASELINE.json):

    metric  : fp64 Hermite pair interactions/s at N=1e5 (BASELINE.json config 3:
              N=100,000 Plummer cluster, Maschberger IMF, Hermite block timesteps)
    step    : one `gravity.evolve_model(t + dt_outer)` with the reference's outer step
              dt_outer = t_f/1000 = 0.01 Myr (al26_nbody.py:786,833) -- hundreds of block steps
    value   : pairs / device time, inputs resident in HBM (CUDA events on the library's stream)
    e2e     : the same step through the reference-facing C-ABI calls with HOST buffers: mass channel
              in (al26_nbody.py:874) -> evolve_model (:833) -> bulk getters out (:876,886-891),
              host<->device copies inside the timed region
    roofline: the force kernel (K1) alone on a full N x N evaluation, 60 flop per pair, against the
              DFMA-microkernel FP64 peak measured on the same GPU in this run
    cpu_baseline / --impl reference: the CPU oracle (a restatement -- AMUSE ph4 itself is not
              installable: no MPI, no network), all host cores, bounded sample of the same workload

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
Under torchrun (N > 1) one rank per GPU; rank 0 prints ONE JSON line.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fp64 Hermite pair interactions/s at N=1e5"  # BASELINE.json's metric; metric_for() labels other sizes
UNIT = "pairs/s"


def metric_for(n):
    return METRIC if n == 100_000 else f"fp64 Hermite pair interactions/s at N={n:.0e}".replace("e+0", "e")


def config_label(n):
    return {100_000: "BASELINE config 3", 1_000_000: "BASELINE config 4", 10_000: "BASELINE config 2 size",
            1_000: "BASELINE config 1 size"}.get(n, "not a BASELINE size")
FLOP_PER_PAIR = 60.0  # acc + jerk + pot, GRAPE counting


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--particles", dest="n", type=int, default=100_000)  # use --particles under torchrun (argparse prefix clash with --nnodes)
    ap.add_argument("--dt-myr", type=float, default=0.01)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-enrich", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer arm (scaling sweeps at N=1e6)")
    ap.add_argument("--dist-mode", default="p2p", choices=["p2p", "nccl"], help="N > 1: peer-memory exchange or NCCL all-gather")
    ap.add_argument("--split-min", type=int, default=0, help="peer-memory mode: exchange only blocks of at least this many active particles (0 = auto)")
    ap.add_argument("--fuse-max", type=int, default=-1, help="loop kernels: largest block on the fused small-step path (0 = off, -1 = library default)")
    ap.add_argument("--step-mode", type=int, default=-1, choices=[-1, 0, 1, 2, 3], help="1 GPU: -1 = library default (graph; + cluster engine when N fits one cluster, + chip engine when it fits the chip), 0 = CUDA graph, 1 = persistent loop kernel, 2 = graph + cluster engine, 3 = graph + chip engine")
    ap.add_argument("--chip-max", type=int, default=-1, help="largest block the chip engine steps (-1 = library default, 0 = engine off)")
    ap.add_argument("--cpu-pairs", type=float, default=1.2e10, help="pair budget of the CPU sample")
    ap.add_argument("--no-config4", action="store_true", help="skip the N=1e6 gravity-only sub-record (BASELINE config 4)")
    ap.add_argument("--config4-dt-myr", type=float, default=5.0e-4, help="outer step of the config-4 evolve (forced full-N sync < 10 %% of its pairs)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the parity sub-record (same call on one GPU inside the job)")
    ap.add_argument("--reinit-policy", type=int, default=0, choices=[0, 1], help="e2e arm: what the per-step mass channel costs (0 = forces only, ph4's recommit; 1 = forces + initial timesteps)")
    return ap.parse_args()


def workload(pkg, n, seed, dt_myr):
    """Config 3: Plummer sphere, r_vir = 1 pc, Maschberger IMF; N-body units from nbody_to_si(Rc, M)."""
    U = pkg.units
    c = pkg.ic.cluster(n, seed=seed, model="plummer")
    cv = U.nbody_to_si(1.0 | U.pc, float(c["m_msun"].sum()) | U.MSun)
    span = cv.time_to_nbody(dt_myr | U.Myr)
    return c, cv, span


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([f.strip() for f in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic_bytes(kernel_csv):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu
    --set full summary under profiles/ (one capture per change; None if the file is not there)."""
    try:
        tot = 0.0
        for line in open(os.path.join(ROOT, "profiles", kernel_csv)):
            f = line.strip().split(",")
            if f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[f[1]]
                tot += float(f[2]) * scale
        return tot or None
    except Exception:
        return None


def cpu_sample(pkg, c, span, pair_budget):
    """The CPU oracle on the same ICs: initial full force + block steps until the pair budget."""
    from oracle import hermite as H
    H.prefer_native()  # -march=native, built on this box
    H.use_all_cores()
    n = len(c["m"])
    o = H.HermiteOracle(n)
    o.commit(*[c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")])
    t0 = time.perf_counter()
    o.begin(span)
    while o.counters()[1] < pair_budget:
        nd, fin = o.advance(1)
        if fin:
            break
    dt = time.perf_counter() - t0
    steps, pairs = o.counters()
    return pairs / dt, pairs, steps, dt, H.num_threads(), H.flavour()


def run_reference(args):
    """--impl reference: the CPU path (oracle port; kind 'port') on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # only the host-side IC / unit helpers are used here: the CUDA library is never loaded on this arm
    pkg = importlib.import_module("26al-nbody_b200")
    c, cv, span = workload(pkg, args.n, args.seed, args.dt_myr)
    from oracle import hermite as H
    flavour = H.prefer_native()  # -march=native, compiled on this box (falls back to the portable x86-64-v3 build)
    H.use_all_cores()  # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core it may run on
    n = args.n
    o = H.HermiteOracle(n)
    o.commit(*[c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")])
    t_end = [span]
    o.begin(t_end[0])  # initial N^2 force: set-up, not timed (the GPU arm's steps do not re-initialise either)
    budget = 2.0e9  # pairs per step: a bounded slice of the same block-step sequence
    outer = [0]

    def one_step():
        s0, p0 = o.counters()
        t0 = time.perf_counter()
        while o.counters()[1] - p0 < budget:
            nd, fin = o.advance(1)
            if fin:  # this outer step is complete: synchronise and start the next one, as the script's loop does
                o.finish()
                t_end[0] += span
                outer[0] += 1
                o.begin(t_end[0])
        s1, p1 = o.counters()
        return p1 - p0, time.perf_counter() - t0, (s0, s1)

    for _ in range(args.warmup):
        one_step()
    pairs = secs = 0.0
    first = last = None
    for _ in range(args.steps):
        a, b, (s0, s1) = one_step()
        pairs += a; secs += b
        first = s0 if first is None else first
        last = s1
    value = pairs / secs
    cores = H.num_threads()
    sample = (f"oracle/hermite_oracle.c (Hermite-4 block-step restatement of ph4, gcc -O3 -march={flavour}, OpenMP x{cores}), same "
              f"N={n} ICs; each step = consecutive block steps until >= {budget:.0e} pairs; the timed steps cover block steps "
              f"{first}..{last} of the run (outer steps 0..{outer[0]} of {args.dt_myr} Myr), the initial N^2 force excluded")
    line = {"impl": "reference", "metric": metric_for(n), "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"N={n} Plummer, Maschberger IMF, r_vir=1 pc, Hermite-4 block timesteps, eta=0.14, eps2=0 ({config_label(n)})",
                       "outer_dt_myr": args.dt_myr},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "block_step_range": [int(first), int(last)],
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


BYTES_PER_DISC_UPDATE = 210.0  # kinematics 48 + r_disk, tau 16 + flags 2 + 6 inventories read 48; 6 inventories + 6 finals + flag written 97


def enrichment_numbers(pkg, ctx, rank, world, do_cpu):
    """BASELINE config 5: 1000 massive stars x 1e6 discs, wind + SN deposit + decay + condense, in the disc kernel's
    three modes (0 exact = bit-identical to the reference kernel, 1 fast = the north_star's 1e-10 tolerance mode,
    2 fast + pruned), plus the HBM-bound regime real clusters live in (190 and 16 massive stars)."""
    n_disc = 1_000_000
    pc_km = 3.08567758128e13
    f26, f60 = pkg.decay_fractions(0.01)
    dt_s = 0.01 * 1e6 * 365.242199 * 86400
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    def problem(n_hm):
        n = n_disc + n_hm
        n -= n % max(world, 1)
        rng = np.random.default_rng(5)
        mass = np.full(n, 1.0)
        hm = np.arange(0, n, n // n_hm)[:n_hm]
        mass[hm] = 20.0
        wr26 = np.zeros(n); wr60 = np.zeros(n); sn = np.zeros(n)
        wr26[hm], wr60[hm], sn[hm] = 1e-5, 1e-7, 1e26
        mdot = np.zeros(n); mdot[hm] = 1e16
        pv = np.concatenate([rng.normal(0, 1.0 * pc_km, (3, n)), rng.normal(0, 1.0, (3, n))])
        return dict(n=n, hm=hm, mass=mass, wr26=wr26, wr60=wr60, sn=sn, mdot=mdot, pv=pv, rd=np.full(n, 1.49597870691e10),
                    tau=rng.exponential(2.885, n))

    def run(P, mode, reps=6):
        e = pkg.EnrichCore(ctx=ctx)
        e.set_mode(mode)
        try:
            n = P["n"]
            e.commit(P["rd"], P["tau"], np.ones(n), np.zeros(n), P["wr26"], P["wr60"], P["sn"], P["sn"])
            dev, wall, ker = [], [], []
            for k in range(reps):
                t0 = time.perf_counter()
                e.step(P["mass"], P["mdot"], P["pv"], dt_s, 0.01 * (k + 1), 0.1 * pc_km, 2.0 * pc_km, f26, f60)
                wall.append(time.perf_counter() - t0)
                dev.append(e.last_device_ms()[0])
                ker.append(e.last_kernel_ms())
            launches = e.last_device_ms()[1]
        finally:
            e.set_mode(0)
        return float(np.median(ker[2:])), float(np.median(dev[2:])), float(np.median(wall[2:])), launches

    def record(P, mode):
        ker_ms, dev_ms, wall_s, launches = run(P, mode)
        n, n_hm = P["n"], len(P["hm"])
        n_loc = n // world
        gbs = n_loc * BYTES_PER_DISC_UPDATE / (ker_ms * 1e-3) / 1e9
        return {"mode": ("exact", "fast", "pruned")[mode], "sources": n_hm, "kernel_ms": ker_ms, "device_ms_incl_h2d": dev_ms,
                "disc_updates_per_s_kernels": n_loc * world / (ker_ms * 1e-3), "disc_updates_per_s_e2e": n_loc * world / wall_s,
                "source_disc_pairs_per_s_kernels": n_hm * float(n_disc) / (ker_ms * 1e-3), "launches_per_step": launches,
                "achieved_hbm_gbs": gbs, "frac_hbm": (gbs / hbm_peak) if hbm_peak else None}

    P = problem(1000)
    modes = [record(P, m) for m in (0, 1, 2)]
    exact, fast, pruned = modes
    out = {"workload": "1000 massive x 1e6 discs, local+global wind, SN, decay, condense (BASELINE config 5)",
           "disc_updates_per_s_e2e": exact["disc_updates_per_s_e2e"], "disc_updates_per_s_kernels": exact["disc_updates_per_s_kernels"],
           "source_disc_pairs_per_s_kernels": exact["source_disc_pairs_per_s_kernels"], "kernel_ms": exact["kernel_ms"],
           "device_ms_incl_h2d": exact["device_ms_incl_h2d"], "h2d_bytes_per_step": 8 * P["n"] * 8,
           "launches_per_step": exact["launches_per_step"],
           "note": "top-level figures = mode 0 (exact: every pair in the reference's order, wind sums bit-identical to the reference's numba kernel); "
                   "modes 1 / 2 are the north_star's 1e-10 tolerance mode (parity-tested against mode 0: local / SN rows identical, global rows <= 1e-12)",
           "modes": modes,
           "roofline": {"bound": "fp64-issue at 1000 sources in modes 0 (15 DP instructions per pair) and 1 (4 per pair); hbm in mode 2 and for few sources",
                        "bytes_per_disc_update": BYTES_PER_DISC_UPDATE, "bytes_basis": "SURVEY 8(d) algorithmic bytes (ncu: 202 MB per launch of 1e6 discs, profiles/)",
                        "hbm_peak_gbs": hbm_peak, "hbm_peak_source": "MEASURED_PEAKS.json" if hbm_peak else "unavailable",
                        "dp_lane_inst_per_s_mode0": 15.0 * 1000 * float(P["n"] // world) / (exact["kernel_ms"] * 1e-3),
                        "dp_lane_inst_per_s_mode1": 4.0 * 1000 * float(P["n"] // world) / (fast["kernel_ms"] * 1e-3),
                        "dp_lane_peak_per_s_nominal": 148 * 64 * 1.965e9}}
    # the HBM-bound regime: a real N=1e5 cluster has ~190 massive stars, an N=1e4 one ~19
    out["few_sources"] = [record(problem(k), m) for k in (190, 16) for m in (0, 2)]
    if do_cpu and rank == 0:
        out["cpu_baseline"] = enrichment_cpu_baseline(P, dt_s, pc_km)
    return out


def enrichment_cpu_baseline(P, dt_s, pc_km):
    """The reference's OWN numba kernel (lifted into the git-ignored oracle/_ref by oracle/make_ref.py), its four calls per
    outer step (al26_nbody.py:897-933) on all host cores, on a bounded sample of the same discs; the numpy port when the
    generated file is not there."""
    hm, pv, mdot, rd = P["hm"], P["pv"], P["mdot"], P["rd"]
    calls = ((P["wr26"], 0.0, 2.0 * pc_km), (P["wr60"], 0.0, 2.0 * pc_km), (P["wr26"], 0.1 * pc_km, 0.1 * pc_km),
             (P["wr60"], 0.1 * pc_km, 0.1 * pc_km))
    ref = None
    try:
        from oracle import make_ref
        ref = make_ref.load()
    except Exception:
        ref = None
    if ref is not None:
        try:
            import numba
            numba.set_num_threads(numba.config.NUMBA_NUM_THREADS)
            lm = np.nonzero(P["mass"] == 1.0)[0]
            small = lm[:2000]
            for wr, lim, rad in calls:  # JIT compile outside the timed region
                ref.calc_wind_abs(small, hm, *pv, mdot, wr, rd, lim, rad, dt_s)
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < 8.0 or reps < 2:
                for wr, lim, rad in calls:
                    ref.calc_wind_abs(lm, hm, *pv, mdot, wr, rd, lim, rad, dt_s)
                reps += 1
            dt = (time.perf_counter() - t0) / reps
            return {"value": len(lm) / dt, "unit": "disc-updates/s", "cores": int(numba.get_num_threads()), "kind": "reference",
                    "sample": f"the reference's own calc_wind_abs (al26_nbody.py:642-702, numba njit parallel) x4 calls per step, "
                              f"{len(lm)} discs x {len(hm)} sources, {reps} repetitions; wind deposit only (the reference's SN / "
                              f"decay / condense loops need AMUSE)"}
        except Exception as ex:  # numba missing or the lifted file unusable on this box
            note = f" (reference kernel unavailable: {type(ex).__name__})"
    else:
        note = " (oracle/_ref not generated)"
    from oracle import enrich_oracle as eo
    ns = 50_000
    lm = np.nonzero(P["mass"] == 1.0)[0][:ns]
    t0 = time.perf_counter()
    for wr, lim, rad in calls:
        eo.calc_wind_abs(lm, hm, *pv, mdot, wr, rd, lim, rad, dt_s)
    dt = time.perf_counter() - t0
    return {"value": ns / dt, "unit": "disc-updates/s", "cores": 1, "kind": "port",
            "sample": f"oracle/enrich_oracle.py (numpy restatement of calc_wind_abs x4), {ns} discs x {len(hm)} sources" + note}


def config4_record(pkg, args, ctx, rank, world, barrier, allmax, allsum):
    """BASELINE config 4: N = 1e6 gravity-only, one evolve call long enough that the forced full-N synchronisation step
    at its end is < 10 % of the pairs (the scaling target of the north_star: >= 6x from 1 to 8 GPUs)."""
    import torch
    n = 1_000_000 - 1_000_000 % world
    c, cv, _ = workload(pkg, n, args.seed, args.dt_myr)
    U = pkg.units
    span = cv.time_to_nbody(args.config4_dt_myr | U.Myr)
    g = pkg.GravityCore(ctx=ctx)
    g.set_time(0.0)
    g.commit(*[c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")])
    k0, u0, _ = g.energies()
    w_steps, w_pairs = g.evolve(span / 64.0)  # warm-up: initial full force + a few block steps + a synchronisation step
    barrier()
    steps, pairs = g.evolve(span / 64.0 + span)
    ms = allmax(g.last_device_ms()[0])
    tot_pairs = allsum(float(pairs))
    k1, u1, _ = g.energies()
    out = {"workload": f"N={n} Plummer, Maschberger IMF, gravity only ({config_label(1_000_000)}); one evolve_model(t + {args.config4_dt_myr} Myr = {span:.6f} N-body) "
                       f"after a warm-up call of 1/64 of that",
           "metric": metric_for(1_000_000), "value": tot_pairs / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "ms": ms,
           "pairs": tot_pairs, "block_steps": steps, "sync_step_pair_share": float(n) * n / tot_pairs,
           "warmup_pairs": allsum(float(w_pairs)), "dE_over_E": ((k0 + u0) - (k1 + u1)) / (k1 + u1)}
    del g
    torch.cuda.empty_cache()
    return out


def parity_record(pkg, args, ctx, g, rank, world, local, c, span, barrier):
    """N > 1: the same calls on ONE GPU inside the same job (rank 0, a second single-GPU context), compared with what
    the multi-GPU job computed: block-step and pair counts (integers, must be equal), positions (summation order differs
    between the decompositions: agreement to rounding), enrichment (sharded discs: bit-equal to the one-GPU result)."""
    import torch
    import torch.distributed as dist
    n = len(c["m"])
    p = [c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")]
    # -- gravity: a fresh commit on both sides, one outer step
    g.commit(*p)
    g.set_time(0.0)
    steps_n, pairs_n = g.evolve(span)
    t = torch.tensor([float(pairs_n)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t)
    pairs_n = int(t.item())
    state_n = g.get_state()
    # -- enrichment, sharded over the ranks, three steps with a supernova
    mass = c["m_msun"]
    hm = np.nonzero(mass >= 13.0)[0]
    wr = np.zeros(n); wr[hm] = 1e-5
    sn = np.zeros(n); sn[hm] = 1e26
    mdot = np.zeros(n); mdot[hm] = 1e16; mdot[hm[0]] = 0.0
    alive = (mass >= 0.1) & (mass <= 3.0)
    rd = np.full(n, 1.49597870691e10)
    rng = np.random.default_rng(1)
    pv = np.concatenate([rng.normal(0, 3e13, (3, n)), rng.normal(0, 1, (3, n))])
    f26, f60 = pkg.decay_fractions(0.01)

    def enrich(context):
        e = pkg.EnrichCore(ctx=context)
        e.commit(rd, c["tau_disk_myr"], alive, np.zeros(n), wr, wr, sn, sn)
        ev = []
        for k in range(1, 4):
            ev += e.step(mass, mdot, pv, 3.15e11, 0.01 * k, 3.0857e12, 6e13, f26, f60).tolist()
        inv, fin, al, kk = e.get()
        return ev, inv, fin, al.astype(np.uint8), kk.astype(np.uint8)

    ev_n, inv_n, fin_n, al_n, _ = enrich(ctx)
    parts = torch.from_numpy(np.concatenate([inv_n.ravel(), fin_n.ravel(), al_n.astype(np.float64)])).cuda()
    dist.all_reduce(parts)  # every rank filled only its own disc slice (zeros elsewhere)
    parts = parts.cpu().numpy()
    out = None
    if rank == 0:
        one = pkg.Context(local)
        g1 = pkg.GravityCore(ctx=one)
        g1.commit(*p)
        steps_1, pairs_1 = g1.evolve(span)
        state_1 = g1.get_state()
        dx = float(max(np.max(np.abs(a - b)) for a, b in zip(state_n[1:4], state_1[1:4])))
        ev_1, inv_1, fin_1, al_1, _ = enrich(one)
        ref = np.concatenate([inv_1.ravel(), fin_1.ravel(), al_1.astype(np.float64)])
        out = {"what": f"the same N={n} outer step and 3 enrichment steps on one GPU inside this job vs the {world}-GPU result",
               "block_steps": [int(steps_1), int(steps_n)], "block_steps_equal": bool(steps_1 == steps_n),
               "pairs": [int(pairs_1), int(pairs_n)], "pairs_equal": bool(pairs_1 == pairs_n),
               "max_abs_dx_nbody": dx, "positions_within_1e-9": bool(dx < 1e-9),
               "position_checksum": [float(np.sum(state_1[1]) + np.sum(state_1[2]) + np.sum(state_1[3])),
                                     float(np.sum(state_n[1]) + np.sum(state_n[2]) + np.sum(state_n[3]))],
               "sn_events_equal": bool(ev_1 == ev_n), "enrichment_bit_equal": bool(np.array_equal(ref, parts))}
        out["ok"] = bool(out["block_steps_equal"] and out["pairs_equal"] and out["positions_within_1e-9"] and
                         out["sn_events_equal"] and out["enrichment_bit_equal"])
        one.close()
    barrier()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    pkg = importlib.import_module("26al-nbody_b200")
    ctx = pkg.Context(local)
    ctx.set_step_mode(args.step_mode)
    ctx.set_fuse_max(args.fuse_max)
    ctx.set_chip_max(args.chip_max)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pkg.dist.init_context(ctx, rank, world, device="cuda", mode=args.dist_mode, split_min=args.split_min)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n = args.n - args.n % world
    c, cv, span = workload(pkg, n, args.seed, args.dt_myr)
    p = [c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")]
    g = pkg.GravityCore(ctx=ctx)
    g.commit(*p)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    # ---- roofline of the dominant kernel (K1 force), timed alone --------------------------------
    # FP64 FMA peak measured on this GPU: the best operand pattern of the issue-rate microbenchmark (a DFMA reading three
    # distinct registers is register-file limited to ~2/3 of the pipe rate on B200; with <= 2 distinct registers it runs
    # at 99 % of 148 SMs x 64 lanes x clock).  The force kernel's mix of both is judged against the HIGHER figure.
    rates = {v: ctx.fp64_rate(v) for v in (0, 1, 4, 5)}
    peak_tf = 2.0 * max(rates[0], rates[1], rates[4]) / 1e12
    f_ms, f_pairs = g.bench_force(5)
    achieved_tf = f_pairs * FLOP_PER_PAIR / (f_ms * 1e-3) / 1e12
    nominal_tf = 148 * 64 * 2 * 1.965e9 / 1e12
    roofline = {"bound": "fp64", "kernel": "k_force (full N x N/P evaluation, timed alone)",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "peak_source": "DFMA issue-rate microkernel measured in this run, best operand pattern (MEASURED_PEAKS.json has no FP64 entry)",
                "fma_tflops_by_operand_pattern": {"x=fma(x,a,b), a b shared": 2e-12 * rates[0], "x=fma(x,imm,b)": 2e-12 * rates[1],
                                                  "x=fma(x,x,b)": 2e-12 * rates[4], "x=fma(x,a_k,b_k), 3 distinct registers": 2e-12 * rates[5]},
                "frac_of_nominal_37.2": achieved_tf / nominal_tf, "flop_per_pair": FLOP_PER_PAIR,
                "pairs_per_launch": f_pairs, "ms_per_launch": f_ms,
                "algorithmic_bytes_per_launch": 64.0 * n, "traffic": ncu_traffic_bytes("r01_k_force_ncu_full.csv") if n == 100_000 and world == 1 else None,
                "traffic_source": "profiles/r01_k_force_ncu_full.csv (ncu --set full, same kernel and launch shape)"}

    # ---- device-resident arm: K evolve calls, CUDA events on the library stream ----------------
    k0, u0, _ = g.energies()
    t_now = 0.0
    for _ in range(args.warmup):
        t_now += span
        g.evolve(t_now)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    wall0 = time.perf_counter()
    tot_ms = tot_pairs = 0.0
    tot_steps = launches = 0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t_now += span
        steps, pairs = g.evolve(t_now)
        ms, nl = g.last_device_ms()
        tot_ms += allmax(ms)
        tot_pairs += allsum(float(pairs))
        tot_steps += steps
        launches += nl
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    k1, u1, _ = g.energies()
    de = ((k0 + u0) - (k1 + u1)) / (k1 + u1)
    value = tot_pairs / (tot_ms * 1e-3)
    n_chip, chip_ctas, chip_max = ctx.chip_steps()
    n_eng, eng_cs = ctx.engine_steps()

    # ---- end-to-end arm: host buffers in / out every step --------------------------------------
    m_host = torch.empty(n, dtype=torch.float64).pin_memory()
    m_host.numpy()[:] = c["m"]
    outs = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(7)]
    outs_np = [o.numpy() for o in outs]

    g.set_reinit_policy(args.reinit_policy)

    def e2e_step():
        nonlocal t_now
        t_now += span
        g.set_mass(m_host.numpy())       # stel_to_grav.copy_attributes(["mass"])   (:874)
        s_, p_ = g.evolve(t_now)         # gravity.evolve_model(t_new)               (:833)
        g.get_state(outs_np)             # grav_to_clus.copy() / bulk getters        (:876,886-891)
        return s_, p_, g.last_device_ms()[1]

    e_launch = 0
    if args.no_e2e:
        e2e = None
    else:
        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        barrier()
        e_pairs = 0.0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            s_, p_, nl = e2e_step()
            e_pairs += p_
            e_launch += nl
        barrier()
        e_wall = allmax(time.perf_counter() - t0)
        e_pairs = allsum(float(e_pairs))
        e2e = {"value": e_pairs / e_wall, "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 7 * 8 * n,
               "ms_per_step": 1e3 * e_wall / args.steps,
               "reinit_policy": args.reinit_policy,
               "note": "set_mass (the reference's per-step mass channel, al26_nbody.py:874: forces recomputed with the new masses"
                       + (", timesteps of the synchronisation step kept -- ph4's recommit" if args.reinit_policy == 0 else " AND initial timesteps again")
                       + ") + evolve + get_state, host buffers in and out"}

    line = {"metric": metric_for(n), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"N={n} Plummer, Maschberger IMF, r_vir=1 pc, Hermite-4 block timesteps, eta=0.14, eps2=0 "
                                   f"({config_label(n)}); step = evolve_model(t + {args.dt_myr} Myr = {span:.5f} N-body)",
                       "outer_dt_myr": args.dt_myr, "parallelism": ("1 GPU" if world == 1 else
                                       f"x{world}: replicated state, owner-computes (i % {world}), corrected particles scattered to peers over NVLink inside the loop kernel"
                                       if args.dist_mode == "p2p" else f"i-partition x{world} + NCCL j all-gather"),
                       "l2": "256 MiB device write between timed steps (state is 20 MB < L2; kernel is FP64-bound)"},
            "block_steps_per_step": tot_steps / args.steps, "pairs_per_step": tot_pairs / args.steps,
            "wall_s_timed_region": wall, "dE_over_E": de, "t_end_nbody": t_now,
            "small_steps": {"chip_engine_ctas": chip_ctas, "chip_engine_max_block": chip_max, "chip_engine_block_steps_since_commit": n_chip,
                            "cluster_engine_size": eng_cs, "cluster_engine_block_steps_since_commit": n_eng},
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches + e_launch), "clocks": clocks}

    if not args.no_enrich:
        line["enrichment"] = enrichment_numbers(pkg, ctx, rank, world, do_cpu=not args.no_cpu)
    if world > 1 and not args.no_parity:
        line["parity"] = parity_record(pkg, args, ctx, g, rank, world, local, c, span, barrier)
    g = None
    if not args.no_config4:
        line["config4"] = config4_record(pkg, args, ctx, rank, world, barrier, allmax, allsum)
    if world == 1 and rank == 0 and not args.no_cpu:
        v, pr, st, secs, cores, flavour = cpu_sample(pkg, c, span, args.cpu_pairs)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"oracle/hermite_oracle.c (ph4-equivalent restatement, not AMUSE ph4; gcc -O3 -march={flavour}), same ICs: initial "
                                          f"N^2 force + {st} block steps = {pr:.3e} pairs in {secs:.1f} s"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
