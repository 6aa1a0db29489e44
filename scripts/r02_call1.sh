#!/bin/bash
# round-2 GPU call 1: tests, probes, quick bench, ncu of the FP64 peak microkernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
nproc >> gpurun_out/c1_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
PROBE_KERNELS=1 timeout 300 python scripts/probe.py > gpurun_out/c1_probe_kernels.log 2>&1
PROBE_LATENCY=1 PROBE_SIZES=100000 PROBE_COMBOS=1:32,1:0,0:0 timeout 300 python scripts/probe.py > gpurun_out/c1_probe_latency.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/c1_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/c1_bench.log
timeout 300 ncu --set full --clock-control none -k regex:k_dfma_peak -c 2 -o gpurun_out/c1_dfma_peak -f python -c "
import importlib,sys
sys.path.insert(0,'.')
pkg=importlib.import_module('26al-nbody_b200')
c=pkg.Context(0); print(c.fp64_peak_tflops())" > gpurun_out/c1_ncu_dfma.log 2>&1
tail -3 gpurun_out/c1_pytest.log; tail -2 gpurun_out/c1_bench.log | cut -c1-600
