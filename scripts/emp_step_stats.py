"""Endless-MysteryPath single-step launches (env.step): time per step and how many auto-resets adopt a path generated
ahead. Usage: python scripts/emp_step_stats.py [n] [steps]"""
import sys
import os
sys.path.insert(0, os.environ.get("MG_PKG_DIR") or os.path.join(os.path.dirname(__file__), "..", "endless-memory-gym_b200"))
import torch
import memory_gym_b200 as mg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 600
dev = torch.device("cuda:0")
env = mg.make("Endless-MysteryPath-v0", num_envs=n, device=dev)
env.reset(seed=1)
g = torch.Generator(device=dev).manual_seed(3)
acts = torch.randint(0, 4, (T, n), generator=g, device=dev)
for phase in ("cold (no chains yet)", "warm"):
    spec0 = env.state[:, 2048:2112].clone().view(torch.int32)
    dones = torch.zeros((), dtype=torch.int64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for t in range(T):
        _, _, d, _, _ = env.step(acts[t])
        dones += d.sum()
    e1.record()
    torch.cuda.synchronize()
    spec1 = env.state[:, 2048:2112].clone().view(torch.int32)
    adopted = int((spec1[:, 2] - spec0[:, 2]).sum())
    chains = int(((spec1[:, 1] >> 13) - (spec0[:, 1] >> 13)).sum())
    ahead = ((spec1[:, 0] >> 4) - spec1[:, 1]) & 255
    print(f"{phase}: {e0.elapsed_time(e1) * 1e3 / T:.1f} us per step (host loop included), resets={int(dones)} adopted={adopted} "
          f"({adopted / max(1, int(dones)):.3f}) new chains={chains} missed: empty {int((spec1[:, 14] - spec0[:, 14]).sum())} "
          f"stale {int((spec1[:, 15] - spec0[:, 15]).sum())}; episodes ahead now: "
          f"{[int((ahead == k).sum()) for k in range(9)]}", flush=True)
# device-timed: T single-step launches queued back to back (programmatic dependent launch), no host work in between
ring = torch.empty((4, n, 84, 84, 3), dtype=torch.uint8, device=dev)
rew = torch.empty((T, n), dtype=torch.float64, device=dev)
dn = torch.empty((T, n), dtype=torch.bool, device=dev)
env.step_many(acts, ring, rew, dn)
spec0 = env.state[:, 2048:2112].clone().view(torch.int32)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
env.step_many(acts, ring, rew, dn)
e1.record()
torch.cuda.synchronize()
spec1 = env.state[:, 2048:2112].clone().view(torch.int32)
print(f"step_many: {e0.elapsed_time(e1) * 1e3 / T:.1f} us per single-step launch (device-timed), resets={int(dn.sum())} "
      f"adopted={int((spec1[:, 2] - spec0[:, 2]).sum())} missed {int((spec1[:, 14] - spec0[:, 14]).sum())}", flush=True)
# the bench's shape: blocks of 20 launches per library call
for slots in (4, 20):
    ring = torch.empty((slots, n, 84, 84, 3), dtype=torch.uint8, device=dev)
    B = 20
    for rep in range(2):
        torch.cuda.synchronize()
        e0.record()
        for b in range(T // B):
            env.step_many(acts[b * B:(b + 1) * B], ring, rew[:B], dn[:B])
        e1.record()
        torch.cuda.synchronize()
    print(f"step_many in blocks of {B}, ring of {slots}: {e0.elapsed_time(e1) * 1e3 / (T // B * B):.1f} us per launch", flush=True)
