"""Measurement helper (not part of the product): kernel time per step vs batch size / launch geometry.
    python scripts/sweep.py [env_id]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "endless-memory-gym_b200")]
import torch
import memory_gym_b200 as mg

env_id = sys.argv[1] if len(sys.argv) > 1 else "MortarMayhem-v0"
dev = torch.device("cuda:0")
T = 96


def run(n, wpc, reps=6, ring_mb=None):
    env = mg.make(env_id, num_envs=n, device=dev, warps_per_cta=wpc)
    env.reset(seed=0)
    g = torch.Generator(device=dev).manual_seed(1)
    acts = torch.randint(0, env.spec_.action_nvec[0], (T,) + tuple(env._actions_dev.shape), generator=g, device=dev)
    slots = T  # every step's frames go to their own [t] slot of a rollout buffer: nothing is rewritten in L2
    ring = torch.empty((slots, n, 84, 84, 3), dtype=torch.uint8, device=dev)
    rew = torch.empty((T, n), dtype=torch.float64, device=dev)
    dn = torch.empty((T, n), dtype=torch.bool, device=dev)
    env.rollout(acts, ring, rew, dn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        env.rollout(acts, ring, rew, dn)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * T)
    return us, n * 21705 / us / 1e3


sizes = [int(v) for v in os.environ.get("SWEEP_N", "148,1480,4096,8192,32768").split(",")]
for n in sizes:
    us, gbs = run(n, 0)
    print(f"n={n:6d} wpc=auto  {us:8.2f} us/step  {n / us:8.2f} M env-steps/s  {gbs:7.1f} GB/s", flush=True)
for wpc in [int(v) for v in os.environ.get("SWEEP_WPC", "5,6,7,8,9,10").split(",")]:
    us, gbs = run(4096, wpc)
    print(f"n=  4096 wpc={wpc:2d}  {us:8.2f} us/step  {4096 / us:8.2f} M env-steps/s  {gbs:7.1f} GB/s", flush=True)
