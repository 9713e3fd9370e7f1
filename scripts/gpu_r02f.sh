#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mg_ss_raster -s 2 -c 1 -f -o gpurun_out/r02f_raster python scripts/prof_rollout.py SearingSpotlights-v0 8192 20 2 > gpurun_out/r02f.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mg_ss_logic -s 2 -c 1 -f -o gpurun_out/r02f_logic python scripts/prof_rollout.py SearingSpotlights-v0 8192 20 2 >> gpurun_out/r02f.log 2>&1
tail -3 gpurun_out/r02f.log; ls -la gpurun_out/*.ncu-rep
