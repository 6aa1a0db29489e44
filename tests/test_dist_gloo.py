"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path -- rank slices, the ncclUniqueId
bootstrap, and the distributed block-step PROTOCOL (each rank predicts its j-slice, all-gather, force on
its own active particles, local corrector, min-reduce of the next block time) emulated with
numpy + the oracle's force routine and compared with the single-process oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, queue):
    sys.path.insert(0, ROOT)
    import importlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("26al-nbody_b200")
    from oracle import hermite as H
    try:
        # 1. bootstrap: every rank ends up with rank 0's 128 bytes
        secret = bytes(range(128))
        got = pkg.dist.broadcast_unique_id(lambda: secret, rank, device="cpu")
        assert got == secret
        # 2. slices tile [0, n) in rank order
        n = 250
        i0, i1 = pkg.dist.slice_of(n, rank, world)
        ends = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(ends, torch.tensor([i0, i1]))
        assert ends[0][0] == 0 and ends[-1][1] == n and all(ends[k][1] == ends[k + 1][0] for k in range(world - 1))
        # 3. the distributed block-step protocol vs the single-process oracle
        n = 256
        m, x, y, z, vx, vy, vz = pkg.ic.plummer(n, np.random.default_rng(3))
        ref = H.HermiteOracle(n); ref.commit(m, x, y, z, vx, vy, vz)
        i0, i1 = pkg.dist.slice_of(n, rank, world)
        eta, dt_min, span = 0.14, 2.0 ** -40, 0.02
        X = np.stack([x, y, z])[:, i0:i1].copy(); V = np.stack([vx, vy, vz])[:, i0:i1].copy()
        nl = i1 - i0

        def gather(loc):  # all-gather of the local slice of a (k, nl) array -> (k, n)
            parts = [torch.zeros(loc.shape, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(loc)))
            return np.concatenate([p.numpy() for p in parts], axis=1)

        def force_local(PX, PV, act):  # act: local indices
            idx = (act + i0).astype(np.int32)
            out = H.force(m, PX[0], PX[1], PX[2], PV[0], PV[1], PV[2], idx=idx)
            return np.stack(out[:3]), np.stack(out[3:6])

        PX, PV = gather(X), gather(V)
        A, J = force_local(PX, PV, np.arange(nl))
        sa, sj = (A * A).sum(0), (J * J).sum(0)
        dt0 = np.minimum(eta * 0.0625 * np.sqrt(sa / sj), 2.0 ** -5)
        dt = np.maximum(2.0 ** np.floor(np.log2(dt0)), dt_min)
        D = 2.0 ** np.floor(np.log2(span))
        dt = np.minimum(dt, min(D, 0.125))
        t = np.zeros(nl)
        ref.begin(span)
        nsteps = 0
        while True:
            tn_loc = torch.tensor([np.min(t + dt)])
            dist.all_reduce(tn_loc, op=dist.ReduceOp.MIN)       # the 8-byte min-reduce
            tn = float(tn_loc)
            ridx, rtn = ref.get_active()
            if tn > span:
                assert ref.advance(1)[1]
                break
            assert tn == rtn
            act = np.nonzero(t + dt == tn)[0]
            assert np.array_equal(act + i0, ridx[(ridx >= i0) & (ridx < i1)])  # bit-exact active set, per rank
            s = tn - t
            XP = X + V * s + A * (s * s * 0.5) + J * (s * s * s * (1.0 / 6.0))
            VP = V + A * s + J * (s * s * 0.5)
            PX, PV = gather(XP), gather(VP)                      # the j all-gather
            if len(act):
                A1, J1 = force_local(PX, PV, act)
                h = dt[act]
                da = A[:, act] - A1
                al = -3.0 * da - h * (2.0 * J[:, act] + J1)
                be = 2.0 * da + h * (J[:, act] + J1)
                X[:, act] = XP[:, act] + h * h * (al * (1.0 / 12.0) + be * (1.0 / 20.0))
                V[:, act] = VP[:, act] + h * (al * (1.0 / 3.0) + be * 0.25)
                a2 = (2.0 * al + 6.0 * be) / (h * h); a3 = (6.0 * be) / (h * h * h)
                A[:, act], J[:, act] = A1, J1
                n1, nj, n2, n3 = (A1 * A1).sum(0), (J1 * J1).sum(0), (a2 * a2).sum(0), (a3 * a3).sum(0)
                dtA = eta * np.sqrt((np.sqrt(n1 * n2) + nj) / (np.sqrt(nj * n3) + n2))
                nd = h.copy()
                half = (dtA < h) & (0.5 * h >= dt_min)
                nd[half] = 0.5 * h[half]
                q = tn / (2.0 * h)
                dbl = (dtA >= 2.0 * h) & (2.0 * h <= min(D, 0.125)) & (q == np.floor(q)) & ~(dtA < h)
                nd[dbl] = 2.0 * h[dbl]
                t[act] = tn; dt[act] = nd
            ref.advance(1)
            nsteps += 1
            rt, rdt = ref.get_timesteps()
            assert np.array_equal(rdt[i0:i1], dt) and np.array_equal(rt[i0:i1], t)  # ladder bit-exact
        rs = ref.get_state()
        # positions of particles that are still mid-step are compared after the oracle's own sync; here compare
        # the corrected state of this rank's slice before synchronisation through accelerations
        ra = ref.get_acc_jerk()
        err = np.max(np.abs(np.stack(ra[:3])[:, i0:i1] - A) / (np.abs(A) + 1e-300))
        assert err < 1e-9, err
        queue.put((rank, "ok", nsteps))
    except Exception as e:  # pragma: no cover
        import traceback
        queue.put((rank, "fail: " + traceback.format_exc(), 0))
    finally:
        dist.destroy_process_group()


def test_two_rank_protocol_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, nsteps in res:
        assert status == "ok", f"rank {rank}: {status}"
    assert res[0][2] == res[1][2] and res[0][2] > 5


# ---------------------------------------------------------------------------------------------------
# The DEFAULT multi-GPU protocol (peer-memory mode, csrc/hermite_loop.cu: k_loop_dist), emulated rank by rank:
# replicated state, ownership i % world, exchanged vs redundant block steps by the global active count, corrected
# records "stored" into every rank's double-buffered staging slab under the exchange id, pulled by tag in the next
# predictor pass, next block time = minimum of the ranks' mailbox candidates.  gloo is the transport that stands in for
# the NVLink stores; everything else follows the kernel's bookkeeping (xid, parity, prev_exch, split_min).
# ---------------------------------------------------------------------------------------------------
def _correct(eta, dt_min, Dmax, tn, h, XP, VP, A0, J0, A1, J1):
    da = A0 - A1
    al = -3.0 * da - h * (2.0 * J0 + J1)
    be = 2.0 * da + h * (J0 + J1)
    X1 = XP + h * h * (al * (1.0 / 12.0) + be * (1.0 / 20.0))
    V1 = VP + h * (al * (1.0 / 3.0) + be * 0.25)
    a2 = (2.0 * al + 6.0 * be) / (h * h); a3 = (6.0 * be) / (h * h * h)
    n1, nj, n2, n3 = (A1 * A1).sum(0), (J1 * J1).sum(0), (a2 * a2).sum(0), (a3 * a3).sum(0)
    dtA = eta * np.sqrt((np.sqrt(n1 * n2) + nj) / (np.sqrt(nj * n3) + n2))
    nd = h.copy()
    half = (dtA < h) & (0.5 * h >= dt_min)
    nd[half] = 0.5 * h[half]
    q = tn / (2.0 * h)
    dbl = (dtA >= 2.0 * h) & (2.0 * h <= Dmax) & (q == np.floor(q)) & ~(dtA < h)
    nd[dbl] = 2.0 * h[dbl]
    return X1, V1, nd


def _worker_p2p(rank, world, port, queue):
    sys.path.insert(0, ROOT)
    import importlib
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("26al-nbody_b200")
    from oracle import hermite as H
    try:
        n, split_min = 192, 6
        m, x, y, z, vx, vy, vz = pkg.ic.plummer(n, np.random.default_rng(5))
        ref = H.HermiteOracle(n); ref.commit(m, x, y, z, vx, vy, vz)
        eta, dt_min, span = 0.14, 2.0 ** -40, 0.03
        X = np.stack([x, y, z]); V = np.stack([vx, vy, vz])
        owner = np.arange(n) % world

        def force(PX, PV, idx):
            out = H.force(m, PX[0], PX[1], PX[2], PV[0], PV[1], PV[2], idx=idx.astype(np.int32))
            return np.stack(out[:3]), np.stack(out[3:6])

        A, J = force(X, V, np.arange(n))  # the init step is an exchanged one in the kernel; identical on every rank here
        dt0 = np.minimum(eta * 0.0625 * np.sqrt((A * A).sum(0) / (J * J).sum(0)), 2.0 ** -5)
        D = min(2.0 ** np.floor(np.log2(span)), 0.125)
        dt = np.minimum(np.maximum(2.0 ** np.floor(np.log2(dt0)), dt_min), D)
        t = np.zeros(n)
        slab = [dict(X=np.zeros((3, n)), V=np.zeros((3, n)), A=np.zeros((3, n)), J=np.zeros((3, n)), t=np.zeros(n), dt=np.zeros(n),
                     tag=np.zeros(n, dtype=np.uint64)) for _ in range(2)]  # two parities
        xid, prev_exch = 0, False
        ref.begin(span)
        n_exch = n_red = 0
        tn = float(np.min(t + dt))
        while tn <= span:
            # ---- predictor pass: pull what the peers staged during exchange `xid`, then predict everything
            if prev_exch:
                sv = slab[xid & 1]
                got = sv["tag"] == xid
                X[:, got], V[:, got], A[:, got], J[:, got] = sv["X"][:, got], sv["V"][:, got], sv["A"][:, got], sv["J"][:, got]
                t[got], dt[got] = sv["t"][got], sv["dt"][got]
            ridx, rtn = ref.get_active()
            assert tn == rtn
            act = np.nonzero(t + dt == tn)[0]
            assert np.array_equal(act, ridx)                                   # bit-exact active set on EVERY rank
            own = act[owner[act] == rank]
            s = tn - t
            XP = X + V * s + A * (s * s * 0.5) + J * (s * s * s * (1.0 / 6.0))
            VP = V + A * s + J * (s * s * 0.5)
            rest_min = np.min(np.where(t + dt == tn, np.inf, t + dt))          # non-active particles: folded by the scheduler pass
            exchange = len(act) >= split_min                                   # same decision on every rank (global count)
            todo = own if exchange else act
            cand = rest_min
            if len(todo):
                A1, J1 = force(XP, VP, todo)
                X1, V1, nd = _correct(eta, dt_min, D, tn, dt[todo], XP[:, todo], VP[:, todo], A[:, todo], J[:, todo], A1, J1)
                cand = min(cand, float(np.min(tn + nd)))
            if exchange:
                this_id = xid + 1
                rec = (todo, X1, V1, A1, J1, nd) if len(todo) else (todo, None, None, None, None, None)
                allrec = [None] * world
                dist.all_gather_object(allrec, rec)                            # the NVLink stores into every rank's slab ...
                sv = slab[this_id & 1]
                for idx, rx, rv, ra, rj, rd in allrec:
                    if len(idx):
                        sv["X"][:, idx], sv["V"][:, idx], sv["A"][:, idx], sv["J"][:, idx] = rx, rv, ra, rj
                        sv["t"][idx], sv["dt"][idx], sv["tag"][idx] = tn, rd, this_id
                cands = [None] * world
                dist.all_gather_object(cands, cand)                            # ... and the mailbox of the cross-GPU barrier
                tn_next = min(cands)
                xid, prev_exch = this_id, True
                n_exch += 1
            else:
                if len(todo):
                    X[:, todo], V[:, todo], A[:, todo], J[:, todo] = X1, V1, A1, J1
                    t[todo], dt[todo] = tn, nd
                tn_next = cand
                prev_exch = False
                n_red += 1
            ref.advance(1)
            # the rank's effective state (local + what is still staged) against the oracle: ladder bit-exact
            te, dte = t.copy(), dt.copy()
            if prev_exch:
                got = slab[xid & 1]["tag"] == xid
                te[got], dte[got] = slab[xid & 1]["t"][got], slab[xid & 1]["dt"][got]
            rt, rdt = ref.get_timesteps()
            assert np.array_equal(rt, te) and np.array_equal(rdt, dte)
            tn = tn_next
        assert ref.advance(1)[1]
        if prev_exch:
            got = slab[xid & 1]["tag"] == xid
            A[:, got] = slab[xid & 1]["A"][:, got]
        ra = np.stack(ref.get_acc_jerk()[:3])
        assert np.max(np.abs(ra - A) / (np.abs(A) + 1e-300)) < 1e-9
        queue.put((rank, "ok", (n_exch, n_red)))
    except Exception:  # pragma: no cover
        import traceback
        queue.put((rank, "fail: " + traceback.format_exc(), (0, 0)))
    finally:
        dist.destroy_process_group()


def test_peer_memory_protocol_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_worker_p2p, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, counts in res:
        assert status == "ok", f"rank {rank}: {status}"
    assert res[0][2] == res[1][2] and res[0][2][0] > 3 and res[0][2][1] > 3  # both kinds of block step were taken
