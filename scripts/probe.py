"""GPU probe: force-kernel rate, FP64 peak, evolve timings vs span (development aid, not the bench)."""
import importlib, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("26al-nbody_b200")
ctx = pkg.Context(0)
print("device", ctx.device_info())
print("fp64 peak TF/s (dfma microkernel):", ctx.fp64_peak_tflops())
if os.environ.get("PROBE_KERNELS"):
    for n in (1000, 10000, 100000):
        c = pkg.ic.cluster(n, seed=0)
        ctx.set_step_mode(0)
        g = pkg.GravityCore(ctx=ctx)
        g.set_time(0.0)
        g.commit(*[c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")])
        g.evolve(2.0 ** -9)
        g.begin(2.0 ** -9 + 2.0 ** -5)
        pr = g.profile_steps(400)
        g.advance(-1); g.finish()
        print(f"N={n}: per-kernel us (individual launches, events in between): " + ", ".join(f"{k} {v:.2f}" for k, v in pr.items()), flush=True)
    sys.exit(0)
if os.environ.get("PROBE_DECOMP"):
    n = 100000
    c = pkg.ic.cluster(n, seed=0)
    for mr, ov in ((16, 3200.0), (8, 3200.0), (32, 3200.0), (64, 3200.0), (32, 1000.0), (16, 8000.0), (32, 8000.0)):
        ctx.set_decomposition(mr, ov)
        g = pkg.GravityCore(ctx=ctx)
        g.set_time(0.0)
        g.commit(*[c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")])
        row = []
        for na, reps in ((0, 3), (20000, 5), (5000, 10), (1000, 20), (300, 20), (100, 20)):
            ms, pairs = g.bench_force(reps, n_act=na)
            row.append(f"{na or n}:{pairs/ms*1e-6:.0f}G")
        steps, pairs = g.evolve(2.0 ** -5)
        ms, _ = g.last_device_ms()
        print(f"max_rounds {mr} overhead {ov}: " + " ".join(row) + f" | evolve 2^-5: {ms:.1f} ms {pairs/ms*1e-6:.1f} G", flush=True)
    ctx.set_decomposition()
    sys.exit(0)
if os.environ.get("PROBE_LATENCY"):
    sizes = [int(a) for a in os.environ.get("PROBE_SIZES", "1000,10000,100000").split(",")]
    combos = [tuple(int(b) for b in a.split(":")) for a in os.environ.get("PROBE_COMBOS", "1:32,1:8,1:0,0:0").split(",")]
    for n in sizes:
        c = pkg.ic.cluster(n, seed=0)
        for mode, fuse in combos:
            ctx.set_step_mode(mode)
            ctx.set_fuse_max(fuse)
            g = pkg.GravityCore(ctx=ctx)
            g.set_time(0.0)
            g.commit(*[c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")])
            g.evolve(2.0 ** -9)
            steps, pairs = g.evolve(2.0 ** -9 + 2.0 ** -5)
            ms, _ = g.last_device_ms()
            nf = ctx.fused_steps()
            ne, cs = ctx.engine_steps()
            print(f"N={n} mode={('graph', 'loop', 'graph+engine')[mode]} fuse_max={fuse}: {steps} steps, {ms*1e3/steps:.2f} us/step, {pairs/ms*1e-6:.1f} Gpairs/s, fused {nf}, engine {ne} (cluster {cs})", flush=True)
            if mode == 2 and ne:
                pr = ctx.loop_profile()
                print("    engine, CTA 0 cycles per engine step (predict+list / sync+count / force / sync / correct+min / sync): " + ", ".join(f"{v / ne:.0f}" for v in pr.values()), flush=True)
            if mode and nf:
                tot = max(sum(ctx.block_histogram()), 1)
                pr = ctx.fuse_profile()
                print("    ns per step: " + ", ".join(f"{k} {v / (tot if k in ('scan', 'barrier') else nf):.0f}" for k, v in pr.items()), flush=True)
    ctx.set_step_mode(-1)
    ctx.set_fuse_max(-1)
    sys.exit(0)
if os.environ.get("PROBE_BIGBLOCK"):
    n = 100000
    c = pkg.ic.cluster(n, seed=0)
    for thr in (2048, 1024, 512, 256, 128, 64):
        ctx.set_big_block(thr)
        g = pkg.GravityCore(ctx=ctx)
        g.set_time(0.0)
        g.commit(*[c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")])
        row = []
        for na, reps in ((64, 20), (100, 20), (200, 20), (300, 20), (500, 20), (800, 20), (1000, 20), (1500, 20), (2000, 20)):
            ms, pairs = g.bench_force(reps, n_act=na)
            row.append(f"{na}:{pairs/ms*1e-6:.0f}G")
        t_end = 2.0 ** -5
        steps, pairs = g.evolve(t_end)
        ms, _ = g.last_device_ms()
        print(f"big_block {thr}: " + " ".join(row) + f" | evolve 2^-5: {ms:.1f} ms, {pairs/ms*1e-6:.1f} G", flush=True)
    ctx.set_big_block(2048)
    sys.exit(0)
if os.environ.get("PROBE_VARIANTS"):
    n = 100000
    c = pkg.ic.cluster(n, seed=0)
    for v in range(4):
        ctx.set_force_variant(v)
        g = pkg.GravityCore(ctx=ctx)
        g.commit(*[c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")])
        row = []
        for na, reps in ((0, 3), (20000, 5), (5000, 10), (1000, 20), (300, 20), (32, 20), (8, 20), (1, 20)):
            ms, pairs = g.bench_force(reps, n_act=na)
            row.append(f"{na or n}:{ms*1e3:.1f}us/{pairs/ms*1e-6:.0f}G")
        print(f"variant {v}: " + "  ".join(row), flush=True)
    ctx.set_force_variant(0)
for n in [int(a) for a in (sys.argv[1:] or ["100000"])]:
    c = pkg.ic.cluster(n, seed=0)
    g = pkg.GravityCore(ctx=ctx)
    g.set_time(0.0)
    g.commit(*[c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")])
    ms, pairs = g.bench_force(3)
    print(f"N={n} full force: {ms:.3f} ms, {pairs/ms*1e-6:.1f} Gpairs/s, {pairs*60/ms*1e-9:.2f} TF/s(60 flop)")
    for na in (1, 8, 32, 100, 300, 500, 1000, 2000, 2048, 5000, 20000):
        if na > n: break
        ms, pairs = g.bench_force(20, n_act=na)
        print(f"   n_act={na:6d}: {ms*1e3:8.1f} us/launch (incl. ~2 us reset launch), {pairs/ms*1e-6:7.1f} Gpairs/s")
    t0 = time.perf_counter(); g.initialize(); print("initialize wall", time.perf_counter() - t0)
    t, dt = g.get_timesteps()
    e, cnt = np.unique(np.log2(dt), return_counts=True)
    print("initial dt ladder:", dict(zip(e.astype(int).tolist(), cnt.tolist())))
    for mode in (1, 0):
        ctx.set_step_mode(mode)
        g.set_time(0.0)
        g.commit(*[c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")])
        g.initialize()
        k0, u0, s0 = g.energies()
        tnow = 0.0
        print(" step mode", "persistent loop" if mode else "graph")
        for lg in (-14, -12, -10, -8, -7, -5):
            span = 2.0 ** lg
            tnow += span
            t0 = time.perf_counter()
            steps, pairs = g.evolve(tnow)
            wall = time.perf_counter() - t0
            ms, nl = g.last_device_ms()
            print(f"  span 2^{lg}: {steps} block steps, {pairs:.3e} pairs, dev {ms:.2f} ms, wall {wall*1e3:.2f} ms, "
                  f"{pairs/ms*1e-6:.1f} Gpairs/s, {ms*1e3/max(steps,1):.1f} us/step, mean n_act {pairs/n/max(steps,1):.0f}, launches {nl}")
            if wall > 60: break
        print("  block-size histogram (log2 bins):", {b: h for b, h in enumerate(ctx.block_histogram()) if h})
        if mode:
            pr = ctx.loop_profile(); nst = max(sum(ctx.block_histogram()), 1)
            print("  loop kernel, CTA 0 cycles per block step:", {k: round(v / nst) for k, v in pr.items()})
        k1, u1, _ = g.energies()
        print("  dE/E", ((k0 + u0) - (k1 + u1)) / (k1 + u1), "t=", tnow)
    ctx.set_step_mode(0)
