"""Where a rank's time goes in peer-memory mode (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_profile.py
N = 1e5 Plummer (BASELINE config 3), one outer step of 0.01 Myr after a warm-up step, for several (fuse_max, split_min)
settings: value (pairs/s, whole job), ms per step, and CTA 0's time split from al26_dist_profile in microseconds per
step category.  One JSON line per setting on rank 0."""
import importlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = importlib.import_module("26al-nbody_b200")
n = int(os.environ.get("AL26_N", "100000"))
n -= n % world
U = pkg.units
c = pkg.ic.cluster(n, seed=0, model="plummer")
cv = U.nbody_to_si(1.0 | U.pc, float(c["m_msun"].sum()) | U.MSun)
span = cv.time_to_nbody(float(os.environ.get("AL26_DT_MYR", "0.01")) | U.Myr)
p = [c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")]
settings = [tuple(int(v) for v in a.split(":")) for a in os.environ.get("AL26_SETTINGS", "0:0,32:0").split(",")]
for st in settings:
    fuse, split = st[0], st[1]
    chip = st[2] if len(st) > 2 else -1  # third field: largest block of the chip engine (-1 = automatic, 0 = engine off)
    ctx = pkg.Context(local)
    ctx.set_fuse_max(fuse)
    ctx.set_chip_max(chip)
    pkg.dist.init_context(ctx, rank, world, device="cuda", mode="p2p", split_min=split)
    g = pkg.GravityCore(ctx=ctx)
    g.commit(*p)
    g.evolve(span)
    p0 = ctx.dist_profile()
    c0, lp0 = ctx.chip_steps(), ctx.loop_profile()
    dist.barrier(); torch.cuda.synchronize()
    steps, pairs = g.evolve(2 * span)
    ms = g.last_device_ms()[0]
    p1 = ctx.dist_profile()
    c1, lp1 = ctx.chip_steps(), ctx.loop_profile()
    t = torch.tensor([ms, float(pairs)], dtype=torch.float64, device="cuda")
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    d = {k: p1[k] - p0[k] for k in p1}
    khz = ctx.device_info()["clock_khz"]
    us = lambda cyc: cyc / (khz * 1e-3)
    if rank == 0:
        n_chip = c1[0] - c0[0]
        out = {"n": n, "world": world, "fuse_max": fuse, "split_min": split, "steps": steps, "ms": float(tmax[0]),
               "chip_engine": {"ctas": c1[1], "max_block": c1[2], "steps": n_chip,
                               "us_per_step_cta0": {k: us(lp1[k] - lp0[k]) / max(n_chip, 1) for k in lp1}},
               "pairs_per_s": float(tsum[1]) / (float(tmax[0]) * 1e-3),
               "fused": {"steps": d["fused_steps"], "ms": us(d["fused_cycles"]) * 1e-3, "us_per_step": us(d["fused_cycles"]) / max(d["fused_steps"], 1)},
               "redundant": {"steps": d["redundant_steps"], "ms": us(d["redundant_cycles"]) * 1e-3, "us_per_step": us(d["redundant_cycles"]) / max(d["redundant_steps"], 1),
                             "mean_active": d["redundant_active"] / max(d["redundant_steps"], 1)},
               "exchanged": {"steps": d["exch_steps"], "mean_active": d["exch_active"] / max(d["exch_steps"], 1),
                             "ms": us(d["exch_predict_cycles"] + d["exch_force_cycles"] + d["exch_correct_cycles"] + d["exch_barrier_cycles"]) * 1e-3,
                             "us_per_step": {k: us(d["exch_" + k + "_cycles"]) / max(d["exch_steps"], 1) for k in ("predict", "force", "correct", "barrier")}}}
        print(json.dumps(out), flush=True)
    dist.barrier()
    ctx.close()
dist.destroy_process_group()
