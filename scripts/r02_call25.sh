#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gravity.py -x -q -k "chip" > gpurun_out/c25_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c25_pytest.log
timeout 300 python scripts/chip_probe.py 100000 0.01 > gpurun_out/c25_probe.log 2>&1; cut -c1-150 gpurun_out/c25_probe.log; python - <<'PY'
import json
for l in open('gpurun_out/c25_probe.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['step_mode'], d['chip_max'], d['chip_cycles_per_step_cta0'], d['owner_cycles_per_owner_step'], d['launches'], d['prologue_cycles_per_launch'])
PY
