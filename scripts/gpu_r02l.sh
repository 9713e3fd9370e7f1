#!/bin/bash
# round-2 GPU call: sanitizers on the new kernels, ncu full capture of the K=20 MM rollout (DRAM traffic), final bench line
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py MortarMayhem-v0 MysteryPath-v0 SearingSpotlights-v0 Endless-MortarMayhem-v0 > gpurun_out/r02l_san_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02l_san_$tool.log | tail -1)"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mg_env_kernel -s 5 -c 1 -f -o gpurun_out/r02l_mm python scripts/prof_rollout.py MortarMayhem-v0 4096 20 8 > gpurun_out/r02l_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:mg_obs_delta -s 2 -c 1 -f -o gpurun_out/r02l_codec python - > gpurun_out/r02l_ncu2.log 2>&1 <<'PY'
import os, sys
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "endless-memory-gym_b200")]
import torch, memory_gym_b200 as mg
env = mg.make("MortarMayhem-v0", num_envs=4096, device="cuda:0"); env.reset(seed=0)
pipe = env.host_pipe(depth=1)
g = torch.Generator().manual_seed(0)
for t in range(70):
    pipe.step(torch.randint(0, 3, (4096, 2), generator=g).pin_memory())
print("ok")
PY
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
tail -c 600 gpurun_out/r02l_bench.json
