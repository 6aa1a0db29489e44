#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
timeout 600 python scripts/enrich_modes.py > gpurun_out/c2_enrich_modes.log 2>&1
tail -25 gpurun_out/c2_pytest.log; cat gpurun_out/c2_enrich_modes.log | cut -c1-400
