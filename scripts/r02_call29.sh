#!/bin/bash
# parity margin per pair form (classic vs split) + the whole gravity suite WITHOUT -x on the split form
mkdir -p gpurun_out
timeout 300 python scripts/parity_margin.py > gpurun_out/c29_margin_split.log 2>&1
AL26_LIB=$PWD/26al-nbody_b200/csrc/libal26b200_classic.so timeout 300 python scripts/parity_margin.py > gpurun_out/c29_margin_classic.log 2>&1
timeout 900 python -m pytest tests/test_gpu_gravity.py -m gpu -q > gpurun_out/c29_pytest.log 2>&1
cat gpurun_out/c29_margin_classic.log gpurun_out/c29_margin_split.log; grep -E "FAILED|passed|failed" gpurun_out/c29_pytest.log
