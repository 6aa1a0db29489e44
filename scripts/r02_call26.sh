#!/bin/bash
# 1 GPU, final code of the round: full GPU test suite, smoke, both bench arms, ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c26_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c26_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c26_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c26_smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/c26_bench_reference.json 2> gpurun_out/c26_bench_reference.err
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/c26_bench_b200.json 2> gpurun_out/c26_bench_b200.err; echo "rc=$?" >> gpurun_out/c26_bench_b200.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/c26_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-config4 --no-e2e --dt-myr 0.0002 > gpurun_out/c26_ncu_launches.log 2>&1
tail -3 gpurun_out/c26_pytest.log; tail -2 gpurun_out/c26_smoke.log; cut -c1-300 gpurun_out/c26_bench_reference.json; cut -c1-300 gpurun_out/c26_bench_b200.json
