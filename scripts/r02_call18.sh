#!/bin/bash
mkdir -p gpurun_out
AL26_CARVEOUT=1 timeout 300 python scripts/chip_probe.py 100000 0.01 2>&1 | cut -c1-200 > gpurun_out/c18_probe_carveout.log; cat gpurun_out/c18_probe_carveout.log
