#!/bin/bash
# 8 GPUs: group tests, multi-GPU parity (exchanged block steps, dyadic end time, chip engine chained), the peer-memory time split with and
# without the chip engine, and the bench line with parity + config 4
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/c21_ngpu.txt
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -q -x > gpurun_out/c21_pytest_group.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c21_pytest_group.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/dist_check.py > gpurun_out/c21_dist_check_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/c21_dist_check_p2p.log
AL26_SETTINGS=0:0:0,0:0:-1 timeout 600 $TR --master-port 29512 scripts/dist_profile.py > gpurun_out/c21_dist_profile.log 2>&1; echo "rc=$?" >> gpurun_out/c21_dist_profile.log
timeout 1200 $TR --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/c21_bench8.log 2>&1; echo "rc=$?" >> gpurun_out/c21_bench8.log
tail -3 gpurun_out/c21_pytest_group.log; tail -3 gpurun_out/c21_dist_check_p2p.log | cut -c1-200; grep "^{" gpurun_out/c21_dist_profile.log | cut -c1-1100; tail -2 gpurun_out/c21_bench8.log | cut -c1-400
