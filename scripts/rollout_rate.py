"""Rollout throughput of one env type, device-timed (for same-box A/B runs of two builds: MG_PKG_DIR selects the package).
Usage: python scripts/rollout_rate.py ENV [N] [K] [SECONDS]"""
import os
import sys
sys.path.insert(0, os.environ.get("MG_PKG_DIR") or os.path.join(os.path.dirname(__file__), "..", "endless-memory-gym_b200"))
import torch
import memory_gym_b200 as mg

env_id = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
secs = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4
dev = torch.device("cuda:0")
env = mg.make(env_id, num_envs=n, device=dev)
env.reset(seed=0)
g = torch.Generator(device=dev).manual_seed(3)
pool = torch.randint(0, env.spec_.action_nvec[0], (4, K) + tuple(env._actions_dev.shape), generator=g, device=dev)
ring = torch.empty((K, n, 84, 84, 3), dtype=torch.uint8, device=dev)
rew = torch.empty((K, n), dtype=torch.float64, device=dev)
dn = torch.empty((K, n), dtype=torch.bool, device=dev)
launches = 0
for phase in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = max(8, launches)
    e0.record()
    for i in range(reps):
        env.rollout(pool[i & 3], ring, rew, dn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = int(reps * secs * 1e3 / ms) + 1
print(f"{env_id} n={n} K={K}: {n * K * reps / ms / 1e3:.1f} M env-steps/s ({reps} launches)", flush=True)
