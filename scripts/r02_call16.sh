#!/bin/bash
# 2 GPUs: chained chip engine with the pull of the last exchange; parity with exchanged block steps and a dyadic end time, with and without the engine
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
AL26_CHIP_MAX=0 timeout 600 $TR --master-port 29510 scripts/dist_check.py > gpurun_out/c16_dist_check_nochip.log 2>&1; echo "rc=$?" >> gpurun_out/c16_dist_check_nochip.log
timeout 600 $TR --master-port 29511 scripts/dist_check.py > gpurun_out/c16_dist_check_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/c16_dist_check_p2p.log
AL26_SETTINGS=0:0:-1 timeout 600 $TR --master-port 29512 scripts/dist_profile.py > gpurun_out/c16_dist_profile.log 2>&1; echo "rc=$?" >> gpurun_out/c16_dist_profile.log
tail -4 gpurun_out/c16_dist_check_nochip.log | cut -c1-300; tail -4 gpurun_out/c16_dist_check_p2p.log | cut -c1-300; grep "^{\|rc=" gpurun_out/c16_dist_profile.log | cut -c1-900
