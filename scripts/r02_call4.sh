#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
timeout 300 python scripts/enrich_modes.py > gpurun_out/c4_enrich_modes.log 2>&1
AL26_LIB=$PWD/26al-nbody_b200/csrc/libal26b200_timing.so PROBE_LATENCY=1 PROBE_SIZES=100000,10000 PROBE_COMBOS=1:32 timeout 300 python scripts/probe.py > gpurun_out/c4_fuse_timing.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/c4_bench1.log 2>&1; echo "rc=$?" >> gpurun_out/c4_bench1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/c4_bench2.log 2>&1; echo "rc=$?" >> gpurun_out/c4_bench2.log
tail -5 gpurun_out/c4_pytest.log; tail -3 gpurun_out/c4_bench1.log | cut -c1-300; tail -3 gpurun_out/c4_bench2.log | cut -c1-300
