#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
AL26_LIB=$PWD/26al-nbody_b200/csrc/libal26b200_timing.so PROBE_LATENCY=1 PROBE_SIZES=100000,10000 PROBE_COMBOS=1:32 timeout 300 python scripts/probe.py > gpurun_out/c5_fuse_timing.log 2>&1
# launch list of a short graph-mode evolve at N=1e5 (per-kernel durations of the block-step kernels)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/c5_launches_n1e5.csv python -c "
import importlib,sys
sys.path.insert(0,'.')
pkg=importlib.import_module('26al-nbody_b200')
ctx=pkg.Context(0); ctx.set_step_mode(0)
c=pkg.ic.cluster(100000,seed=0)
g=pkg.GravityCore(ctx=ctx); g.commit(*[c[k] for k in ('m','x','y','z','vx','vy','vz')])
print(g.evolve(2.0**-9))
" > gpurun_out/c5_ncu_launches.log 2>&1
tail -3 gpurun_out/c5_pytest.log; cat gpurun_out/c5_fuse_timing.log | cut -c1-300
