"""CPU: the scheduling protocol of the chip engine (26al-nbody_b200/csrc/hermite_chip.cu), restated in numpy and driven by
the Hermite oracle.  The kernel itself is parity-tested on the GPU (tests/test_gpu_gravity.py, step mode 3); this test pins
the LOGIC its CTAs rely on, chunk by chunk, on the host:

  * every chunk publishes a header {min(t + dt) over the chunk, how many of its particles attain it} and those particles;
  * the block time is the minimum of the headers, the owners are the chunks that attain it, the active set is the owners'
    published particles, numbered chunk by chunk -- it must be exactly the oracle's active set of that block step;
  * after the step ONLY the owners' headers change (the engine re-reads nothing else);
  * the launch stops at the first block with more than `chip_max` particles, which the grid-wide kernels then take;
  * the block time the chunk is predicted to ahead of time -- min(other chunks' minima, t + dt of the step's active particles
    with unchanged timesteps) -- is the true next block time unless a front runner changes its timestep.
Stands in for ph4's scheduler behind gravity.evolve_model (al26_nbody.py:833), like the oracle it is checked against."""
import importlib

import numpy as np
import pytest

from oracle import hermite as H


def headers(t, dt, n_ctas):
    """per chunk: (min(t + dt) or +inf for an empty chunk, indices of the particles that attain it)"""
    n = len(t)
    per = (n + n_ctas - 1) // n_ctas
    out = []
    for c in range(n_ctas):
        j0, j1 = min(c * per, n), min((c + 1) * per, n)
        if j1 == j0:
            out.append((np.inf, np.empty(0, np.int64)))
            continue
        tn = t[j0:j1] + dt[j0:j1]
        m = tn.min()
        out.append((m, j0 + np.nonzero(tn == m)[0]))
    return out


@pytest.mark.parametrize("n,n_ctas,chip_max,model", [(300, 7, 32, "plummer"), (64, 148, 4, "plummer"), (500, 16, 256, "fractal")])
def test_chunk_headers_reproduce_the_oracle_schedule(n, n_ctas, chip_max, model):
    pkg = importlib.import_module("26al-nbody_b200")
    c = pkg.ic.cluster(n, seed=n, model=model, require_massive=False)
    p = [c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")]
    o = H.HermiteOracle(n, eps2=1e-6 if model == "fractal" else 0.0)
    o.commit(*p)
    o.initialize()
    o.begin(0.0625)
    t, dt = o.get_timesteps()
    hd = headers(t, dt, n_ctas)
    taken = handed_over = guessed = guess_right = 0
    for _ in range(400):
        mins = np.array([h[0] for h in hd])
        tn = mins.min()
        owners = np.nonzero(mins == tn)[0]
        act = np.concatenate([hd[k][1] for k in owners])  # slot numbering: chunk by chunk
        oi, ot = o.get_active()
        if ot > 0.0625 or len(oi) == 0:
            break
        assert tn == ot and np.array_equal(np.sort(act), np.sort(oi)), "headers and oracle disagree on the block"
        assert len(set(act.tolist())) == len(act)
        # the engine's ahead-of-time guess of the NEXT block time
        rest = np.delete(mins, owners)
        guess = min(rest.min() if len(rest) else np.inf, (tn + dt[act]).min())
        nd, fin = o.advance(1)
        if fin:
            break
        if len(act) > chip_max:
            handed_over += 1  # a block for the grid-wide kernels; the engine would reload its chunks afterwards
        else:
            taken += 1
        t2, dt2 = o.get_timesteps()
        changed = np.nonzero((t2 != t) | (dt2 != dt))[0]
        assert set(changed.tolist()) <= set(act.tolist()), "a block step touched a particle outside its active set"
        hd2 = headers(t2, dt2, n_ctas)
        for k in range(n_ctas):
            if k not in owners:  # "the other chunks' minima cannot have changed"
                assert hd2[k][0] == hd[k][0] and np.array_equal(hd2[k][1], hd[k][1])
        nxt = min(h[0] for h in hd2)
        guessed += 1
        guess_right += (guess == nxt)
        if guess != nxt:  # only a timestep change of an active particle can make the guess wrong
            assert np.any(dt2[act] != dt[act])
        t, dt, hd = t2, dt2, hd2
    assert taken + handed_over == guessed and guessed > (50 if n >= 300 else 8)
    assert guess_right >= 0.7 * guessed
    if chip_max < 32:
        assert handed_over > 0
