#!/bin/bash
# 2 GPUs: what the cooperative launch attribute costs per chained launch
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
AL26_PLAIN_LAUNCH=1 AL26_SETTINGS=0:0:-1 timeout 600 $TR --master-port 29512 scripts/dist_profile.py > gpurun_out/c17_dist_profile_plain.log 2>&1; echo "rc=$?" >> gpurun_out/c17_dist_profile_plain.log
grep "^{\|rc=" gpurun_out/c17_dist_profile_plain.log | cut -c1-400
AL26_PLAIN_LAUNCH=1 timeout 300 python scripts/chip_probe.py 100000 0.01 2>&1 | cut -c1-200 > gpurun_out/c17_probe_plain.log; cat gpurun_out/c17_probe_plain.log
