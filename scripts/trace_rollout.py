"""Measurement helper: phase timestamps of every env's LAST step inside a rollout launch (mg_rollout_io.trace).
The pipelined Searing-Spotlights rollout kernels carry no stamps (only the launch time is printed for them)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "endless-memory-gym_b200")]
import numpy as np
import torch
import memory_gym_b200 as mg

dev = torch.device("cuda:0")
env_id = sys.argv[1] if len(sys.argv) > 1 else "SearingSpotlights-v0"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
T = int(sys.argv[3]) if len(sys.argv) > 3 else 48
env = mg.make(env_id, num_envs=n, device=dev)
env.reset(seed=0)
g = torch.Generator(device=dev).manual_seed(1)
acts = torch.randint(0, env.spec_.action_nvec[0], (T,) + tuple(env._actions_dev.shape), generator=g, device=dev)
obs = torch.empty((T, n, 84, 84, 3), dtype=torch.uint8, device=dev)
rew = torch.empty((T, n), dtype=torch.float64, device=dev)
dn = torch.empty((T, n), dtype=torch.bool, device=dev)
trace = torch.zeros((n, 16), dtype=torch.int64, device=dev)
for _ in range(3):
    env.rollout(acts, obs, rew, dn)
env._trace_ptr = trace.data_ptr()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
env.rollout(acts, obs, rew, dn)
e1.record()
torch.cuda.synchronize()
print(f"launch: {e0.elapsed_time(e1) * 1e3:.1f} us for {T} steps x {n} envs = {e0.elapsed_time(e1) * 1e3 / T:.2f} us/step")
b = trace.cpu().numpy().astype(np.int64)
for nm, i0, i1 in (("step start -> logic done", 3, 4), ("logic done -> emit start", 4, 5), ("emit (or compose->issue)", 5, 6),
                   ("emit: update base", 8, 9), ("emit: lit + lut", 9, 10), ("emit: masked copy", 10, 11),
                   ("12->13 (EP reset: path generation | SS: spotlights B)", 12, 13),
                   ("13->14 (EP reset: rest | SS: logic C)", 13, 14), ("3->12 (SS: logic A)", 3, 12)):
    if b[:, i0].any() and b[:, i1].any():
        sel = (b[:, i0] > 0) & (b[:, i1] > 0)
        d = (b[sel, i1] - b[sel, i0]) / 1e3
        print(f"  {nm:28s} mean {d.mean():7.2f} us  p50 {np.median(d):7.2f}  p95 {np.percentile(d, 95):7.2f}  max {d.max():7.2f}")
