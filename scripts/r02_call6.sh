#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.log
PROBE_LATENCY=1 PROBE_SIZES=100000,10000 PROBE_COMBOS=1:32,1:0,0:0 timeout 300 python scripts/probe.py > gpurun_out/c6_probe_latency.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/dist_check.py > gpurun_out/c6_dist_check_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/c6_dist_check_p2p.log
AL26_SETTINGS=0:0,32:0 timeout 900 $TR --master-port 29512 scripts/dist_profile.py > gpurun_out/c6_dist_profile.log 2>&1; echo "rc=$?" >> gpurun_out/c6_dist_profile.log
tail -3 gpurun_out/c6_pytest.log; cat gpurun_out/c6_probe_latency.log | grep "N="; tail -2 gpurun_out/c6_dist_check_p2p.log | cut -c1-200; grep "^{" gpurun_out/c6_dist_profile.log | cut -c1-1000
