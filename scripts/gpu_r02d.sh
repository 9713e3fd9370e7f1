#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02d_pytest.log
for tw in 2 3 4; do
  MG_SS_TW=$tw timeout 300 python bench.py --env SearingSpotlights-v0 --steps 20 --warmup 5 --no-cpu --no-configs 2>&1 | tail -1 > gpurun_out/r02d_ss_tw$tw.json
done
MG_SS_FUSED=1 timeout 300 python bench.py --env SearingSpotlights-v0 --steps 20 --warmup 5 --no-cpu --no-configs 2>&1 | tail -1 > gpurun_out/r02d_ss_fused.json
timeout 300 python bench.py --env Endless-SearingSpotlights-v0 --steps 20 --warmup 5 --no-cpu --no-configs 2>&1 | tail -1 > gpurun_out/r02d_ess.json
timeout 300 python bench.py --env SearingSpotlights-v0 --steps 128 --warmup 5 --no-cpu --no-configs 2>&1 | tail -1 > gpurun_out/r02d_ss_k128.json
cat gpurun_out/r02d_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02d_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.1f frac %.3f single %.1f us" % (d["value"], d["roofline"]["frac"], d["single_step"]["us_per_step"]))
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-400:])
PY
