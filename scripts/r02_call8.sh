#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_enrich.py tests/test_gpu_driver.py -m gpu -q -x > gpurun_out/c8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c8_pytest.log
timeout 300 python scripts/enrich_modes.py > gpurun_out/c8_enrich_modes.log 2>&1
AL26_SOURCES=16 AL26_MODE=0 timeout 300 ncu --set full --clock-control none -k regex:k_enrich -s 4 -c 2 -o gpurun_out/c8_enrich16_m0 -f python scripts/enrich_ncu_probe.py > gpurun_out/c8_ncu_enrich.log 2>&1
tail -3 gpurun_out/c8_pytest.log; cut -c1-330 gpurun_out/c8_enrich_modes.log
