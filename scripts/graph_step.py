"""Experiment: one CUDA graph per env step (H2D of the actions from pinned memory, the step kernel, D2H of reward+done)
replayed from the host — the launch-bound policy-in-the-loop shape. Checks the replayed steps against plain step()."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "endless-memory-gym_b200")]
import torch
import memory_gym_b200 as mg

dev = torch.device("cuda:0")
env_id = sys.argv[1] if len(sys.argv) > 1 else "MortarMayhem-v0"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T = 400
a = mg.make(env_id, num_envs=n, device=dev)
b = mg.make(env_id, num_envs=n, device=dev)
a.reset(seed=0)
b.reset(seed=0)
g = torch.Generator(device=dev).manual_seed(1)
acts = torch.randint(0, a.spec_.action_nvec[0], (T,) + tuple(a._actions_dev.shape), generator=g, device=dev)
acts_h = acts.cpu().pin_memory()
step = a.capture_step()          # StepGraph: .actions_host (pinned), .results_host (pinned), .obs, .replay()
torch.cuda.synchronize()
ok = True
for t in range(60):
    step.actions_host.copy_(acts_h[t])
    step.replay()
    torch.cuda.synchronize()
    o, r, d, _, _ = b.step(acts[t])
    ok &= bool(torch.equal(step.obs, o)) and bool(torch.equal(step.results_host[:8 * n].view(torch.float64), r.cpu()))
print("graph replay == step():", ok)
t0 = time.perf_counter()
for t in range(60, T):
    step.actions_host.copy_(acts_h[t])
    step.replay()
    step.done_event.synchronize()     # the host has this step's reward + done
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{env_id} n={n}: {dt / (T - 60) * 1e6:.1f} us per step from the host = {n * (T - 60) / dt / 1e6:.1f} M env-steps/s")
