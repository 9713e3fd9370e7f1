"""Measurement helper: us per single-step launch (mg_*_step_many, one launch per step, PDL) in steady state."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "endless-memory-gym_b200")]
import torch
import memory_gym_b200 as mg

dev = torch.device("cuda:0")
env_id = sys.argv[1] if len(sys.argv) > 1 else "MortarMayhem-v0"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T = 128
env = mg.make(env_id, num_envs=n, device=dev)
env.reset(seed=0)
g = torch.Generator(device=dev).manual_seed(1)
acts = torch.randint(0, env.spec_.action_nvec[0], (T,) + tuple(env._actions_dev.shape), generator=g, device=dev)
ring = torch.empty((T, n, 84, 84, 3), dtype=torch.uint8, device=dev)
rew = torch.empty((T, n), dtype=torch.float64, device=dev)
dn = torch.empty((T, n), dtype=torch.bool, device=dev)
for _ in range(8):   # de-synchronise the episodes (random lengths) before timing
    env.step_many(acts, ring, rew, dn)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8):
    env.step_many(acts, ring, rew, dn)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (8 * T)
print(f"{env_id} n={n}: {us:.2f} us/step = {n / us:.1f} M env-steps/s")
