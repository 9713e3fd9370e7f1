#!/bin/bash
mkdir -p gpurun_out
for w in 0 7 8 9; do echo "== wpc $w"; timeout 200 python scripts/trace.py MortarMayhem-v0 4096 $w 2>&1 | grep -v "memset\|copy 1 GiB"; done > gpurun_out/r02o_trace.log 2>&1
cat gpurun_out/r02o_trace.log
