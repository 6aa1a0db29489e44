set -x
timeout 900 python -m pytest tests/test_gpu_gravity.py -x -q -k "chip" > gpurun_out/c13_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/c13_pytest.log
timeout 600 python scripts/chip_probe.py 100000 0.01 > gpurun_out/c13_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/c13_probe.log | cut -c1-1200
