"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
i-partition + NCCL j all-gather against the single-process CPU oracle on the same ICs."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = importlib.import_module("26al-nbody_b200")
from oracle import hermite as H
from oracle import enrich_oracle as eo

ctx = pkg.Context(local)
ctx.set_fuse_max(int(os.environ.get("AL26_FUSE_MAX", "-1")))
MODE = os.environ.get("AL26_DIST_MODE", "p2p")
pkg.dist.init_context(ctx, rank, world, device="cuda", mode=MODE, split_min=int(os.environ.get("AL26_SPLIT_MIN", "48")))  # 48: block steps ARE exchanged at this N (automatic: ~4600)
ctx.set_chip_max(int(os.environ.get("AL26_CHIP_MAX", "-1")))

n = 4096
c = pkg.ic.cluster(n, seed=7)
p = [c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")]
g = pkg.GravityCore(ctx=ctx)
g.commit(*p)
o = H.HermiteOracle(n); o.commit(*p)
g.initialize(); o.initialize()
i0, i1 = n * rank // world, n * (rank + 1) // world
ga, oa = g.get_acc_jerk(), o.get_acc_jerk()
err = max(np.max(np.abs(a[i0:i1] - b[i0:i1]) / (np.abs(b[i0:i1]) + 1e-300)) for a, b in zip(ga[:3], oa[:3]))
vrel = lambda a, b: np.max(np.linalg.norm(np.stack(a) - np.stack(b), axis=0) / np.linalg.norm(np.stack(b), axis=0))
sl = slice(None) if MODE == "p2p" else slice(i0, i1)  # peer-memory mode: the state is replicated, every rank holds all forces
assert vrel([a[sl] for a in ga[:3]], [b[sl] for b in oa[:3]]) < 1e-12 and vrel([a[sl] for a in ga[3:6]], [b[sl] for b in oa[3:6]]) < 1e-12
assert np.array_equal(g.get_timesteps()[1][i0:i1], o.get_timesteps()[1][i0:i1])
k_g, u_g, s_g = g.energies(); k_o, u_o, s_o = o.energies()
assert abs(k_g - k_o) < 1e-12 * abs(k_o) and abs(u_g - u_o) < 1e-11 * abs(u_o), (k_g, k_o, u_g, u_o)
sg = g.evolve(0.02); so = o.evolve(0.02)
tot_pairs = torch.tensor([float(sg[1])], dtype=torch.float64, device="cuda"); dist.all_reduce(tot_pairs)
assert int(tot_pairs.item()) == so[1], (tot_pairs.item(), so[1])
assert sg[0] == so[0], (sg[0], so[0])
# a dyadic end time: the last block step of the call sits on a coarse level (many active particles -> an exchanged step
# right before the synchronisation step)
sg2 = g.evolve(0.03125); so2 = o.evolve(0.03125)
tot2 = torch.tensor([float(sg2[1])], dtype=torch.float64, device="cuda"); dist.all_reduce(tot2)
assert sg2[0] == so2[0] and int(tot2.item()) == so2[1], (sg2, so2, tot2.item())
assert np.array_equal(g.get_timesteps()[1], o.get_timesteps()[1])
gs, os_ = g.get_state(), o.get_state()   # get_state returns the GLOBAL arrays on every rank
dx = max(np.max(np.abs(a - b)) for a, b in zip(gs[1:4], os_[1:4]))
assert dx < 1e-9, dx
# enrichment: discs partitioned, source table replicated
mass = c["m_msun"]; hm = np.nonzero(mass >= 13.0)[0]
wr = np.zeros(n); wr[hm] = 1e-5; sn = np.zeros(n); sn[hm] = 1e26
mdot = np.zeros(n); mdot[hm] = 1e16; mdot[hm[0]] = 0.0
alive = (mass >= 0.1) & (mass <= 3.0)
rd = np.full(n, 1.49597870691e10)
rng = np.random.default_rng(1)
pv = np.concatenate([rng.normal(0, 3e13, (3, n)), rng.normal(0, 1, (3, n))])
f26, f60 = pkg.decay_fractions(0.01)
e = pkg.EnrichCore(ctx=ctx)
e.commit(rd, c["tau_disk_myr"], alive, np.zeros(n), wr, wr, sn, sn)
st = eo.EnrichState(rd, c["tau_disk_myr"], alive, np.zeros(n, bool), wr, wr, sn, sn)
for k in range(1, 4):
    ev = e.step(mass, mdot, pv, 3.15e11, 0.01 * k, 3.0857e12, 6e13, f26, f60)
    ev_o = eo.enrich_step(st, mass, mdot, *pv, 3.15e11, 0.01 * k, 3.0857e12, 6e13, f26, f60)
    assert ev.tolist() == ev_o
inv, fin, al, kk = e.get()
assert np.array_equal(inv[:, i0:i1], st.inv[:, i0:i1]) and np.array_equal(fin[:, i0:i1], st.fin[:, i0:i1])
assert np.array_equal(al[i0:i1], st.disk_alive[i0:i1]) and np.array_equal(kk, st.kicked)
# the sliced upload of the fast modes (explicit positions, several ranks): same results as the exact mode up to the
# hoisted global sum
for mode in (1, 2):
    e2 = pkg.EnrichCore(ctx=ctx)
    e2.set_mode(mode)
    e2.commit(rd, c["tau_disk_myr"], alive, np.zeros(n), wr, wr, sn, sn)
    for k in range(1, 4):
        ev = e2.step(mass, mdot, pv, 3.15e11, 0.01 * k, 3.0857e12, 6e13, f26, f60)
    inv2 = e2.get()[0]
    R = pkg.ROW
    for row in ("local26", "local60", "sne26", "sne60"):
        assert np.array_equal(inv2[R[row]][i0:i1], st.inv[R[row]][i0:i1]), (mode, row)
    for row in ("global26", "global60"):
        a, b = inv2[R[row]][i0:i1], st.inv[R[row]][i0:i1]
        nz = b != 0
        assert np.array_equal(a != 0, nz) and np.max(np.abs(a[nz] / b[nz] - 1.0), initial=0.0) < 1e-12, (mode, row)
    e2.set_mode(0)
if MODE == "p2p":
    pr = ctx.dist_profile()
    assert pr["exch_steps"] > 10  # init + sync steps alone would be 4
print(f"[{MODE}] rank {rank}/{world}: PASS  acc err {err:.2e}, evolve steps {sg[0]} (oracle {so[0]}), pairs local {sg[1]} total {int(tot_pairs.item())}, dx {dx:.2e}", flush=True)
dist.barrier()
ctx.close()
dist.destroy_process_group()
