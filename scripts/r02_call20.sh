#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c20_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c20_pytest.log
cat > /tmp/chip_ncu.py <<'PY'
import importlib, sys, os
sys.path.insert(0, os.getcwd())
import bench
pkg = importlib.import_module("26al-nbody_b200")
c, cv, span = bench.workload(pkg, 100000, 0, 0.0005)
p = [c[k] for k in ("m", "x", "y", "z", "vx", "vy", "vz")]
ctx = pkg.Context(0)
ctx.set_step_mode(3); ctx.set_chip_max(32)
g = pkg.GravityCore(ctx=ctx); g.commit(*p)
print(g.evolve(span), ctx.chip_steps())
ctx.close()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chip -s 20 -c 2 -o gpurun_out/c20_k_chip python /tmp/chip_ncu.py > gpurun_out/c20_ncu.log 2>&1; echo "ncu rc=$?"; tail -5 gpurun_out/c20_ncu.log
