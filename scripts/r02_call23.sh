#!/bin/bash
# N GPUs (argument): the bench line with parity + config 4, final code of the round
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1200 $TR --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/c23_bench$N.log 2>&1; echo "rc=$?" >> gpurun_out/c23_bench$N.log
tail -2 gpurun_out/c23_bench$N.log | cut -c1-300
