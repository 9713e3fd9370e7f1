#!/bin/bash
# usage: runall.sh TAG  -> gpu tests + bench + latency
TAG=$1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest.log
for e in MortarMayhem-v0 SearingSpotlights-v0 Endless-SearingSpotlights-v0 Endless-MysteryPath-v0 MysteryPath-v0 Endless-MortarMayhem-v0; do timeout 200 python bench.py --env $e --steps 20000 --no-cpu --e2e-steps 16 2>&1 | tail -1 >> gpurun_out/${TAG}_bench.log; done
for e in "MortarMayhem-v0 4096" "SearingSpotlights-v0 8192" "Endless-MysteryPath-v0 8192" "MysteryPath-v0 8192" "Endless-MortarMayhem-v0 8192"; do timeout 100 python scripts/step_latency.py $e; done > gpurun_out/${TAG}_latency.log 2>&1
cat gpurun_out/${TAG}_pytest.log
