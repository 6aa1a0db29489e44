#!/bin/bash
mkdir -p gpurun_out
for tag in "" _nif _nifs; do
  echo "== lib$tag" >> gpurun_out/c7_inline_ab.log
  AL26_LIB=$PWD/26al-nbody_b200/csrc/libal26b200$tag.so PROBE_LATENCY=1 PROBE_SIZES=100000 PROBE_COMBOS=1:32,1:0 timeout 300 python scripts/probe.py 2>&1 | grep "N=" >> gpurun_out/c7_inline_ab.log
done
AL26_SOURCES=16 AL26_MODE=0 timeout 300 ncu --set full --clock-control none -k regex:k_enrich -s 4 -c 4 -o gpurun_out/c7_enrich16_m0 -f python scripts/enrich_ncu_probe.py > gpurun_out/c7_ncu_enrich.log 2>&1
AL26_SOURCES=1000 AL26_MODE=2 timeout 300 ncu --set full --clock-control none -k regex:k_enrich -s 4 -c 4 -o gpurun_out/c7_enrich1000_m2 -f python scripts/enrich_ncu_probe.py >> gpurun_out/c7_ncu_enrich.log 2>&1
cat gpurun_out/c7_inline_ab.log
