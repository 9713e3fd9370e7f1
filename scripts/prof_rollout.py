"""Rollout launches only (for `ncu -k mg_env_kernel -s <skip> -c 1`): python scripts/prof_rollout.py ENV N K [LAUNCHES] [step]
The first mg_env_kernel launch is the reset; launches 2.. are K-step rollouts (or, with `step`, single-step launches)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "endless-memory-gym_b200")]
import torch
import memory_gym_b200 as mg

env_id, n, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 8
single = len(sys.argv) > 5 and sys.argv[5] == "step"
env = mg.make(env_id, num_envs=n, device="cuda:0")
g = torch.Generator(device="cuda:0").manual_seed(1234)
acts = torch.randint(0, env.spec_.action_nvec[0], (K,) + tuple(env._actions_dev.shape), generator=g, device="cuda:0")
ring = torch.empty((K, n, 84, 84, 3), dtype=torch.uint8, device="cuda:0")
rew = torch.empty((K, n), dtype=torch.float64, device="cuda:0")
dn = torch.empty((K, n), dtype=torch.bool, device="cuda:0")
env.reset(seed=0)
for _ in range(launches):
    (env.step_many if single else env.rollout)(acts, ring, rew, dn)
torch.cuda.synchronize()
print("ok")
