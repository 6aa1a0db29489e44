"""Outer-loop re-host: `init_cluster` + one outer step `evolve_simulation` + `run`,
in the phase order of /root/reference/al26_nbody.py:704-1113 / :1612-1766, with the two hot phases on the
B200: `gravity.evolve_model` (:833) and the disc routines (:878-1086).

Not optimised host code: it exists so that BASELINE configs 1-2 run end to end through the drop-in
boundary, with the reference's own phase timers (`grav / stel / winds+decay / step`, :796-1109), and
optionally writes the reference's yields book / cluster-yields.csv (yields_io.py) on save steps.  What is NOT
here (out of scope): state checkpoints (pickle), the AGB interloper's set-up and AGB wind
tables (the deposit kernel itself is available: EnrichCore.interloper), plotting.

Differences from the script that are deliberate and do not change results:
  * the virial radius (`cluster.virial_radius()`, an O(N^2) numpy sum every outer step, :770) comes from the
    device pair reduction (`gravity.virial_radius()`), identical to rounding;
  * star classification, SN detection, decay and condense run inside the fused device kernel; the per-star
    inventories stay on the device and are pulled into the cluster columns only on save steps (:1097).
"""
import time

import numpy as np

from . import ic as _ic
from . import units as U
from .enrichment import EnrichCore, ROW, decay_fractions
from .particles import Particles
from .stellar import StellarStub, YieldTables

N_PLOT, STEPS_PER_PLOT = 100, 10        # al26_nbody.py:54-55
R_BUB_LOCAL_WIND = 0.1 | U.pc           # al26_nbody.py:77
INVENTORY_COLUMNS = {                    # cluster column -> inventory row (al26_nbody.py:1556-1577)
    "mass_26al_local": "local26", "mass_26al_global": "global26", "mass_26al_sne": "sne26", "mass_26al_agb": "agb26",
    "mass_60fe_local": "local60", "mass_60fe_global": "global60", "mass_60fe_sne": "sne60", "mass_60fe_agb": "agb60"}


def init_cluster(model, nstars, Rc, yields=None, stellar=None, no_massive_star_requirement=False, r_disk=100.0,
                 seed=0, fractal_dimension=1.6, potential_energy=None):
    """al26_nbody.py:1492-1610, vectorised.  Returns (cluster, converter)."""
    yields = yields or YieldTables.synthetic()
    stellar = stellar or StellarStub()
    c = _ic.cluster(nstars, seed=seed, model=model, fractal_dimension=fractal_dimension,
                    require_massive=not no_massive_star_requirement, potential_energy=potential_energy)
    m = c["m_msun"]
    converter = U.nbody_to_si(Rc, float(m.sum()) | U.MSun)            # :1516
    cl = Particles(nstars)
    cl.mass = m | U.MSun                                               # :1530 (no rescaling)
    for a in ("x", "y", "z"):
        setattr(cl, a, converter.length_to_si(c[a]))
    for a in ("vx", "vy", "vz"):
        setattr(cl, a, converter.speed_to_si(c[a]))
    cl.radius = np.zeros(nstars) | U.au                                # :1542
    cl.kicked = np.zeros(nstars, dtype=bool)                           # :1543
    cl.zams_mass = m | U.MSun  # not in the reference's set: lets the stellar STUB resume from a checkpoint exactly
    cl.r_disk = np.full(nstars, r_disk) | U.au                         # :1547
    cl.tau_disk = c["tau_disk_myr"] | U.Myr                            # :1548
    for col in list(INVENTORY_COLUMNS) + [k + "_final" for k in INVENTORY_COLUMNS]:
        setattr(cl, col, np.zeros(nstars) | U.kg)                      # :1556-1577
    hm = m >= 13.0
    lm = (m >= 0.1) & (m <= 3.0)
    cl.disk_alive = lm.copy()                                          # :1582,1603
    loss = np.where(hm, stellar.total_wind_loss_msun(m), 0.0)
    cl.total_wind_loss = loss | U.MSun                                 # :1583,1604
    wr26 = np.zeros(nstars); wr60 = np.zeros(nstars); sn26 = np.zeros(nstars); sn60 = np.zeros(nstars)
    for i in np.nonzero(hm)[0]:                                        # :1581-1601
        wr26[i], wr60[i], sn26[i], sn60[i] = yields.star_yields(m[i], loss[i])
    cl.wind_ratio_26al, cl.wind_ratio_60fe = wr26, wr60
    cl.sn_yield_26al, cl.sn_yield_60fe = sn26 | U.MSun, sn60 | U.MSun
    return cl, converter


def make_enrichment(gravity, cluster, converter):
    """Commit the per-star enrichment state next to the gravity state (same GPU context)."""
    e = EnrichCore(ctx=gravity._core.ctx)
    e.commit(cluster.r_disk.value_in(U.km), cluster.tau_disk.value_in(U.Myr), cluster.disk_alive, cluster.kicked,
             cluster.wind_ratio_26al, cluster.wind_ratio_60fe,
             cluster.sn_yield_26al.value_in(U.kg), cluster.sn_yield_60fe.value_in(U.kg))
    e.set_units(converter.km_per_length, converter.kms_per_speed)
    return e


def push_inventories(enrich, cluster):
    """cluster columns -> device inventories (resume from a checkpoint, al26_nbody.py:1641-1656): the inverse of
    pull_inventories; disk_alive / kicked go in through make_enrichment's commit."""
    n = len(cluster)
    inv = np.zeros((len(ROW), n))
    fin = np.zeros((len(ROW), n))
    for col, row in INVENTORY_COLUMNS.items():
        inv[ROW[row]] = U.value_in(getattr(cluster, col), U.kg)
        fin[ROW[row]] = U.value_in(getattr(cluster, col + "_final"), U.kg)
    enrich.set_inventories(inv, fin)


def pull_inventories(enrich, cluster):
    """Device inventories -> cluster columns (what the script's yields / checkpoints read, :1097-1105)."""
    inv, fin, alive, kicked = enrich.get()
    for col, row in INVENTORY_COLUMNS.items():
        setattr(cluster, col, inv[ROW[row]] | U.kg)
        setattr(cluster, col + "_final", fin[ROW[row]] | U.kg)
    cluster.disk_alive = alive
    cluster.kicked = kicked
    raw = enrich.get_agb_raw()
    cluster.mass_26al_agb_raw = raw[0] | U.kg                                     # :1560,1572
    cluster.mass_60fe_agb_raw = raw[1] | U.kg


def evolve_simulation(cluster, converter, gravity, stellar, enrich, t_f, save=False, verbose=False, log=print,
                      sync_cluster=True, yields=None, metadata=None):
    """One outer step (al26_nbody.py:704-1113).  Returns (finish, info)."""
    tm = {}
    t0 = time.perf_counter()
    t = gravity.model_time                                                        # :763
    mass_class = np.array(cluster.mass.value_in(U.MSun), copy=True)              # what :767 classifies on
    virial_radius = gravity.virial_radius()                                       # :770 (device)
    if not np.array_equal(cluster.key, gravity.particles.key) or not np.array_equal(cluster.key, stellar.particles.key):
        raise ValueError("Key mismatch between stellar, cluster and gravity particles, simulation cannot continue!")  # :781-783
    dt = t_f / (N_PLOT * STEPS_PER_PLOT)                                          # :786
    t_new = t + dt
    finish = False
    if t_new > t_f:                                                               # :822-825
        t_new = t_f
        dt = t_new - t
        finish = True
    tm["init"] = time.perf_counter() - t0

    t1 = time.perf_counter()
    gravity.evolve_model(t_new)                                                   # :833
    tm["grav"] = time.perf_counter() - t1
    t1 = time.perf_counter()
    stellar.evolve_model(t_new)                                                   # :841
    tm["stel"] = time.perf_counter() - t1

    t1 = time.perf_counter()
    stellar.particles.new_channel_to(gravity.particles).copy_attributes(["mass"])  # :871,874
    if sync_cluster:
        gravity.particles.new_channel_to(cluster).copy()                          # :872,876
    else:
        cluster.mass = stellar.particles.mass
    tm["copy"] = time.perf_counter() - t1

    t1 = time.perf_counter()
    mdot = -np.asarray(stellar.particles.wind_mass_loss_rate.value_in(U.kg / U.s))  # :892
    dt_myr = float(dt.value_in(U.Myr))
    f26, f60 = decay_fractions(dt_myr)                                            # :1050-1051
    sn = enrich.step(mass_class, mdot, None, float(dt.value_in(U.s)), float(t_new.value_in(U.Myr)),
                     float(R_BUB_LOCAL_WIND.value_in(U.km)), float(virial_radius.value_in(U.km)), f26, f60)
    for i in sn:
        log("Star #{} has gone supernova!".format(int(i)))                        # :951
    tm["discs"] = time.perf_counter() - t1
    if save:
        pull_inventories(enrich, cluster)
        if metadata is not None:
            metadata.update(t_new)                                                # :1099
        if yields is not None:
            yields.update_state(gravity.model_time, cluster)                      # :1101
        if metadata is not None:                                                  # :1103-1105
            from .checkpoint import save_checkpoint
            save_checkpoint(metadata.filename, metadata.most_recent_checkpoint, cluster, converter, yields, metadata)
    tm["step"] = time.perf_counter() - t0
    if verbose:
        log("t = {:.3f} Myr: grav {:.3f} s, stel {:.3f} s, discs {:.3f} s, step {:.3f} s".format(
            float(t_new.value_in(U.Myr)), tm["grav"], tm["stel"], tm["discs"], tm["step"]))
    return finish, {"t_new_myr": float(t_new.value_in(U.Myr)), "sn_events": [int(i) for i in sn], "timings": tm,
                    "virial_radius_pc": float(virial_radius.value_in(U.pc)),
                    "block_steps": getattr(gravity, "last_steps", None), "pairs": getattr(gravity, "last_pairs", None)}


def run(nstars=1000, Rc=1.0 | U.pc, t_f=10.0 | U.Myr, model="plummer", seed=0, max_outer_steps=None, verbose=False,
        device=0, fractal_dimension=1.6, yields=None, stellar=None, log=print, yields_file=None,
        epsilon=None, progress=None, step_mode=None, number_of_workers=1, checkpoint_base=None, reload=None,
        n_checkpoint=None, reinit_policy=None):
    """`main()` of the script (al26_nbody.py:1612-1766) with gravity_model == "b200".
    step_mode: None = the library's automatic choice (graph + cluster engine when the particles fit one cluster, else
    the CUDA graph); 0 / 1 / 2 force the graph / the persistent loop kernel / the cluster engine (include/al26_b200.h).
    number_of_workers: GPUs driven from this process (the script's `workers`, al26_nbody.py:57).
    checkpoint_base: write <base>-state-NNNNN / <base>-yields.* on save steps (every 10th outer step, :1755-1758) like
    the script (checkpoint.py); reload = <base> [+ n_checkpoint]: resume from such a checkpoint (`-r base -nc k`)."""
    from .gravity import B200Gravity
    from .gravity import GravityCore
    from . import _lib
    from . import checkpoint as ck
    stellar = stellar or StellarStub()
    ctx = _lib.Group(number_of_workers) if number_of_workers > 1 else _lib.Context(device)
    if step_mode is not None and number_of_workers == 1:
        ctx.set_step_mode(step_mode)
    metadata = None
    ybook = None
    if reload:                                                                    # :1641-1656
        nfile = ck.most_recent_checkpoint(reload) if n_checkpoint is None else n_checkpoint
        cluster, converter, ybook, metadata = ck.load_checkpoint(reload, nfile)
        metadata.update_access_time()
        if t_f is None:
            t_f = metadata.t_f
    else:
        pot = None
        if model == "fractal" and nstars > 2000:
            def pot(m, x, y, z):  # the fractal generator's virial scaling needs U: use the device pair reduction
                g0 = GravityCore(ctx=ctx if number_of_workers == 1 else _lib.Context(device))
                g0.commit(m, x, y, z, np.zeros_like(x), np.zeros_like(x), np.zeros_like(x))
                return g0.energies()[1]
        cluster, converter = init_cluster(model, nstars, Rc, yields=yields, stellar=stellar, seed=seed,
                                          fractal_dimension=fractal_dimension, potential_energy=pot)
    gravity = B200Gravity(converter, number_of_workers=number_of_workers, ctx=ctx)
    if reinit_policy is not None:
        gravity._core.set_reinit_policy(reinit_policy)
    if epsilon is not None:  # the script never sets it (ph4 default 0); sub-virial fractals need it, see DESIGN.md
        gravity.parameters.epsilon_squared = epsilon * epsilon
    gravity.particles.add_particles(cluster)                                      # :1728
    stellar.particles.add_particles(cluster)                                      # :1731
    enrich = make_enrichment(gravity, cluster, converter)
    if reload:                                                                    # :1734-1737
        gravity.model_time = metadata.time
        stellar.model_time = metadata.time
        stellar.evolve_model(metadata.time)  # the STUB's mdot(t) is a function of its clock; SeBa is left as the script leaves it
        push_inventories(enrich, cluster)
    history = []
    n_iter = 0
    finish = False
    if not reload:
        if yields_file is not None or checkpoint_base is not None:                # Yields(filename) + first state (:1739-1745)
            from .yields_io import Yields
            base = checkpoint_base if checkpoint_base is not None else yields_file
            ybook = Yields(base)
            pull_inventories(enrich, cluster)
            ybook.update_state(gravity.model_time, cluster)
        if checkpoint_base is not None:
            metadata = ck.Metadata(t_f=t_f, model=model, nstars=nstars, cluster_radius=Rc, filename=checkpoint_base)
            ck.save_checkpoint(metadata.filename, 0, cluster, converter, ybook, metadata)
    while not finish:                                                             # :1754-1760
        save = (n_iter % 10 == 0)
        finish, info = evolve_simulation(cluster, converter, gravity, stellar, enrich, t_f, save=save,
                                         verbose=verbose, log=log, yields=ybook, metadata=metadata)
        history.append(info)
        if progress is not None:
            progress(n_iter, info)
        n_iter += 1
        if max_outer_steps is not None and n_iter >= max_outer_steps:
            break
    pull_inventories(enrich, cluster)
    return cluster, gravity, stellar, enrich, history
