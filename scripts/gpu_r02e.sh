#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02e_ss_launches.csv python scripts/prof_rollout.py SearingSpotlights-v0 8192 20 4 > gpurun_out/r02e.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02e_ss_step_launches.csv python scripts/prof_rollout.py SearingSpotlights-v0 8192 20 1 step >> gpurun_out/r02e.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02e_ss_launches.csv /dev/null 2>/dev/null | head -12
python scripts/ncu_summary.py gpurun_out/r02e_ss_step_launches.csv /dev/null 2>/dev/null | head -12
