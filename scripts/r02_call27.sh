#!/bin/bash
# A/B: per-tile block barrier (tb) vs per-stage release counters (product) in the force work item; gravity parity with the new ring
mkdir -p gpurun_out
timeout 200 python scripts/force_ab.py > gpurun_out/c27_ab.log 2>&1
AL26_LIB=$PWD/26al-nbody_b200/csrc/libal26b200_tb.so timeout 200 python scripts/force_ab.py >> gpurun_out/c27_ab.log 2>&1
timeout 200 python scripts/force_ab.py >> gpurun_out/c27_ab.log 2>&1
timeout 600 python -m pytest tests/test_gpu_gravity.py -m gpu -x -q > gpurun_out/c27_pytest.log 2>&1
cat gpurun_out/c27_ab.log; tail -5 gpurun_out/c27_pytest.log
