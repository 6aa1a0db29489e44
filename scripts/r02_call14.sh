#!/bin/bash
# 2 GPUs: chip engine chained with the peer-memory loop kernel -- group tests, parity against the oracle, time split with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -q -x > gpurun_out/c14_pytest_group.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c14_pytest_group.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/dist_check.py > gpurun_out/c14_dist_check_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/c14_dist_check_p2p.log
AL26_SETTINGS=0:0:0,0:0:-1,0:0:32 timeout 600 $TR --master-port 29512 scripts/dist_profile.py > gpurun_out/c14_dist_profile.log 2>&1; echo "rc=$?" >> gpurun_out/c14_dist_profile.log
tail -3 gpurun_out/c14_pytest_group.log; tail -4 gpurun_out/c14_dist_check_p2p.log | cut -c1-300; grep "^{\|rc=" gpurun_out/c14_dist_profile.log | cut -c1-1200
