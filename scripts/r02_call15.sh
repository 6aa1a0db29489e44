#!/bin/bash
# 2 GPUs: one loop-kernel step per chained launch
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/dist_check.py > gpurun_out/c15_dist_check_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/c15_dist_check_p2p.log
AL26_SETTINGS=0:0:-1,0:0:256 timeout 600 $TR --master-port 29512 scripts/dist_profile.py > gpurun_out/c15_dist_profile.log 2>&1; echo "rc=$?" >> gpurun_out/c15_dist_profile.log
tail -4 gpurun_out/c15_dist_check_p2p.log | cut -c1-300; grep "^{\|rc=" gpurun_out/c15_dist_profile.log | cut -c1-700
