#!/bin/bash
# A/B of the split-jerk pair form (no potential in block steps): 4 kernel configurations x {release counters, per-tile barrier}; gravity parity
mkdir -p gpurun_out
AB_VARIANTS=0,1,2,3 timeout 300 python scripts/force_ab.py > gpurun_out/c28_ab.log 2>&1
AB_VARIANTS=0,1,2,3 AL26_LIB=$PWD/26al-nbody_b200/csrc/libal26b200_tb.so timeout 300 python scripts/force_ab.py >> gpurun_out/c28_ab.log 2>&1
timeout 600 python -m pytest tests/test_gpu_gravity.py -m gpu -x -q > gpurun_out/c28_pytest.log 2>&1
cat gpurun_out/c28_ab.log; tail -15 gpurun_out/c28_pytest.log
