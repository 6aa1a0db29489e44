"""Run a BASELINE.json configuration end to end through the drop-in boundary (driver.run):
    python scripts/run_config.py 1      # N=1,000 Plummer, 10 Myr
    python scripts/run_config.py 2      # N=10,000 fractal D=1.6, winds + SNe, 20 Myr
Stellar evolution is the parametrised stub (SeBa is out of scope), yields are the synthetic tables."""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("26al-nbody_b200")
U = pkg.units
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else None
kw = {1: dict(nstars=1000, model="plummer", t_f=10.0 | U.Myr), 2: dict(nstars=10000, model="fractal", t_f=20.0 | U.Myr, fractal_dimension=1.6, epsilon=100.0 | U.au)}[cfg]
# config 2: a sub-virial box fractal puts siblings on near-radial orbits; with eps = 0 and no regularisation their
# pericentre passages drive dt to the floor (ph4 would suffer the same), so a disc-sized softening (100 au) is used.
t0 = time.perf_counter()
events = []
def progress(k, info):
    if k % 50 == 0:
        print(f"outer step {k}: t = {info['t_new_myr']:.2f} Myr, {time.perf_counter() - t0:.1f} s, "
              f"{info['block_steps']} block steps, r_vir {info['virial_radius_pc']:.3f} pc", file=sys.stderr, flush=True)
step_mode = int(os.environ["AL26_STEP_MODE"]) if "AL26_STEP_MODE" in os.environ else None  # 2 = graph + cluster engine
cluster, gravity, stellar, enrich, hist = pkg.driver.run(seed=1, max_outer_steps=steps, log=events.append, progress=progress,
                                                         yields_file=f"gpurun_out/config{cfg}", step_mode=step_mode, **kw)
eng_steps, eng_cs = gravity._core.ctx.engine_steps()
wall = time.perf_counter() - t0
tm = {k: float(sum(h["timings"][k] for h in hist)) for k in ("grav", "stel", "copy", "discs", "step")}
inv, fin, alive, kicked = enrich.get()
R = pkg.ROW
lm = (cluster.mass.value_in(U.MSun) <= 3.0)
print(json.dumps({"config": cfg, "step_mode_requested": step_mode, "step_mode_effective": ("graph + cluster engine" if eng_cs else "graph"),
                  "engine_cluster_size": eng_cs, "engine_block_steps": int(eng_steps), "reinit_policy": 0, **{k: (str(v) if not isinstance(v, (int, str)) else v) for k, v in kw.items()},
                  "outer_steps": len(hist), "t_end_myr": hist[-1]["t_new_myr"], "wall_s": wall, "phase_seconds": tm,
                  "block_steps": int(sum(h["block_steps"] for h in hist)), "pairs": float(sum(h["pairs"] for h in hist)),
                  "sn_events": int(sum(len(h["sn_events"]) for h in hist)), "discs_condensed": int((~alive).sum()),
                  "mean_final_26al_local_kg": float(fin[R["local26"]].mean()), "mean_final_26al_global_kg": float(fin[R["global26"]].mean()),
                  "mean_final_26al_sne_kg": float(fin[R["sne26"]].mean()), "virial_radius_pc_end": hist[-1]["virial_radius_pc"]}))
gravity.stop()
