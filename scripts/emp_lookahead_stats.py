"""Endless-MysteryPath rollouts: how many resets adopt a path generated ahead by the helper warps, how long a helper
lane's turn takes, and the rollout throughput, per launch length. Usage: python scripts/emp_lookahead_stats.py [n]"""
import sys
import os
sys.path.insert(0, os.environ.get("MG_PKG_DIR") or os.path.join(os.path.dirname(__file__), "..", "endless-memory-gym_b200"))
import torch
import memory_gym_b200 as mg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0")
for K in (5, 20, 128):
    env = mg.make("Endless-MysteryPath-v0", num_envs=n, device=dev)
    env.reset(seed=1)
    g = torch.Generator(device=dev).manual_seed(3)
    acts = torch.randint(0, 4, (K, n), generator=g, device=dev)
    ring = torch.empty((min(K, 8), n, 84, 84, 3), dtype=torch.uint8, device=dev)
    rew = torch.empty((K, n), dtype=torch.float64, device=dev)
    dn = torch.empty((K, n), dtype=torch.bool, device=dev)
    launches = max(4, 2560 // K)
    for _ in range(launches // 2):
        env.rollout(acts, ring, rew, dn)
    torch.cuda.synchronize()
    spec0 = env.state[:, 2048:2064].clone().view(torch.int32)
    dones = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(launches):
        env.rollout(acts, ring, rew, dn)
        dones += dn.sum()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    spec1 = env.state[:, 2048:2064].clone().view(torch.int32)
    adopted = int((spec1[:, 2] - spec0[:, 2]).sum())
    tickets = int(((spec1[:, 1] >> 13) - (spec0[:, 1] >> 13)).sum())
    work = spec1[:, 3].float()
    print(f"K={K:4d} launches={launches} {n * K * launches / ms / 1e3:7.1f} M env-steps/s  resets={int(dones)} adopted={adopted} "
          f"({adopted / max(1, int(dones)):.3f}) new chains={tickets} helper turn us: median {work.median().item() / 1e3:.0f} "
          f"p90 {work.quantile(0.9).item() / 1e3:.0f} max {work.max().item() / 1e3:.0f}", flush=True)
