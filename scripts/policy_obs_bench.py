"""Measurement helper: mg_obs_to_policy_bf16 (consumer step) GB/s, observations cold (larger than L2) and warm."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "endless-memory-gym_b200")]
import torch
import memory_gym_b200 as mg

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = mg.make("MortarMayhem-v0", num_envs=n, device=dev)
env.reset(seed=0)
K = 12  # K x 87 MB of observations >> 126 MB L2
obs = torch.randint(0, 256, (K, n, 84, 84, 3), dtype=torch.uint8, device=dev)
out = torch.empty((K, n, 3, 84, 84), dtype=torch.bfloat16, device=dev)
for mode, idx in (("cold (HBM reads)", lambda i: i % K), ("warm (same frames)", lambda i: 0)):
    for i in range(K):
        env.policy_obs(obs[idx(i)], out[idx(i)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    R = 5 * K
    e0.record()
    for i in range(R):
        env.policy_obs(obs[idx(i)], out[idx(i)])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / R
    gb = n * (21168 + 42336) / 1e9
    print(f"{mode}: {us:.1f} us per {n}-env batch = {gb / (us * 1e-6):.0f} GB/s (read u8 + write bf16), {n / us:.0f} M frames/s")
