#!/bin/bash
# 1 GPU: final-ish records for profiles/: bench lines (both arms), launch list, ncu captures, dE/E at N=1e5 vs the CPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c10_pytest.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/c10_bench_reference.json 2> gpurun_out/c10_bench_reference.err
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/c10_bench_b200.json 2> gpurun_out/c10_bench_b200.err; echo "rc=$?" >> gpurun_out/c10_bench_b200.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/c10_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-config4 --no-e2e --dt-myr 0.0002 > gpurun_out/c10_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_force -s 1 -c 1 -o gpurun_out/c10_k_force -f python bench.py --steps 1 --warmup 1 --no-cpu --no-config4 --no-e2e --no-enrich --dt-myr 0.0002 > gpurun_out/c10_ncu_force.log 2>&1
AL26_SOURCES=1000 AL26_MODE=0 timeout 300 ncu --set full --clock-control none -k regex:k_enrich -s 4 -c 2 -o gpurun_out/c10_enrich1000_m0 -f python scripts/enrich_ncu_probe.py > gpurun_out/c10_ncu_enrich.log 2>&1
AL26_SOURCES=1000 AL26_MODE=2 timeout 300 ncu --set full --clock-control none -k regex:k_enrich -s 4 -c 2 -o gpurun_out/c10_enrich1000_m2 -f python scripts/enrich_ncu_probe.py >> gpurun_out/c10_ncu_enrich.log 2>&1
AL26_SOURCES=16 AL26_MODE=0 timeout 300 ncu --set full --clock-control none -k regex:k_enrich -s 4 -c 2 -o gpurun_out/c10_enrich16_m0 -f python scripts/enrich_ncu_probe.py >> gpurun_out/c10_ncu_enrich.log 2>&1
timeout 1500 python scripts/energy_drift.py --n 100000 --t-myr 0.05 --cpu > gpurun_out/c10_energy_drift_n1e5.jsonl 2> gpurun_out/c10_energy_drift.err
tail -3 gpurun_out/c10_pytest.log; cut -c1-300 gpurun_out/c10_bench_reference.json; cut -c1-300 gpurun_out/c10_bench_b200.json; cut -c1-400 gpurun_out/c10_energy_drift_n1e5.jsonl
