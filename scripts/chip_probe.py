"""Per-block-step cost of the small steps with and without the chip engine (one GPU).
    python scripts/chip_probe.py [N] [outer dt in Myr]
Prints one JSON line per configuration: ms per outer step, block steps, chip-engine steps and CTA 0's cycle split."""
import importlib, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

pkg = importlib.import_module("26al-nbody_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
dt_myr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
c, cv, span = bench.workload(pkg, n, 0, dt_myr)
p = [c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")]
ctx = pkg.Context(0)
for mode, chip_max in ((0, -1), (3, 32), (3, 256)):
    ctx.set_step_mode(mode)
    ctx.set_chip_max(chip_max)
    g = pkg.GravityCore(ctx=ctx)
    g.set_time(0.0)
    g.commit(*p)
    t = 0.0
    ms = []
    for k in range(4):
        t += span
        steps, pairs = g.evolve(t)
        ms.append(g.last_device_ms()[0])
    prof0 = ctx.loop_profile()
    fp = ctx.fuse_profile_raw()
    n_chip, ctas, mx = ctx.chip_steps()
    hist = ctx.block_histogram()
    x = g.get_state()[1]
    print(json.dumps({"n": n, "step_mode": mode, "chip_max": mx, "chip_ctas": ctas, "ms_per_outer_step": ms, "block_steps_last": steps,
                      "pairs_last": pairs, "chip_steps_total": n_chip, "hist": hist[:18],
                      "chip_cycles_per_step_cta0": {k: (v / n_chip if n_chip else 0) for k, v in prof0.items()},
                      "owner_cycles_per_owner_step": {"rows_poll": fp[0] / max(fp[3], 1), "reduce_correct": fp[1] / max(fp[3], 1), "publish": fp[2] / max(fp[3], 1), "owner_steps": fp[3]},
                      "launches": fp[4], "prologue_cycles_per_launch": prof0["bar3"] / max(fp[4], 1),
                      "x_checksum": float(np.sum(x))}), flush=True)
ctx.close()
