#!/bin/bash
# 2 GPUs: tests first (1 GPU), then multi-GPU parity and the time split of the peer-memory loop
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
timeout 300 python scripts/enrich_modes.py > gpurun_out/c3_enrich_modes.log 2>&1
timeout 120 python -c "
import importlib,sys
sys.path.insert(0,'.')
pkg=importlib.import_module('26al-nbody_b200')
c=pkg.Context(0)
nom=148*64*1.965e9
for v in range(6): r=c.fp64_rate(v); print('fp64 variant',v,'%.4e lane-inst/s'%r,'%.4f of nominal'%(r/nom))
" > gpurun_out/c3_fp64_rates.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/dist_check.py > gpurun_out/c3_dist_check_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/c3_dist_check_p2p.log
AL26_DIST_MODE=nccl timeout 600 $TR --master-port 29513 scripts/dist_check.py > gpurun_out/c3_dist_check_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/c3_dist_check_nccl.log
AL26_SETTINGS=0:0,32:0,32:64,0:64 timeout 900 $TR --master-port 29512 scripts/dist_profile.py > gpurun_out/c3_dist_profile.log 2>&1; echo "rc=$?" >> gpurun_out/c3_dist_profile.log
tail -4 gpurun_out/c3_pytest.log; tail -3 gpurun_out/c3_dist_check_p2p.log; tail -2 gpurun_out/c3_dist_check_nccl.log; grep "^{" gpurun_out/c3_dist_profile.log | cut -c1-900
