"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "endless-memory-gym_b200")]
import torch
import memory_gym_b200 as mg

dev = torch.device("cuda:0")
ids = sys.argv[1:] or ["MortarMayhem-v0", "Endless-MortarMayhem-v0", "MortarMayhemB-v0", "MysteryPath-v0", "Endless-MysteryPath-v0",
                       "SearingSpotlights-v0", "Endless-SearingSpotlights-v0"]
for env_id in ids:
    n, T = 300, 24
    opts = {"max_steps": 10} if env_id in ("MysteryPath-v0", "SearingSpotlights-v0", "Endless-SearingSpotlights-v0") else None
    env = mg.make(env_id, num_envs=n, device=dev, final_obs=True)
    env.reset(seed=0, options=opts)
    g = torch.Generator(device=dev).manual_seed(1)
    acts = torch.randint(0, env.spec_.action_nvec[0], (T,) + tuple(env._actions_dev.shape), generator=g, device=dev)
    for t in range(T):
        env.step(acts[t])
    ring = torch.zeros((4, n, 84, 84, 3), dtype=torch.uint8, device=dev)
    rew = torch.zeros((T, n), dtype=torch.float64, device=dev)
    dn = torch.zeros((T, n), dtype=torch.bool, device=dev)
    env.rollout(acts, ring, rew, dn)
    env.policy_obs()
    # round 2: zero-copy host actions / results, the delta transport (HostPipe), the ground-truth render
    a_host = acts.cpu().pin_memory()
    for mode in ("patch", "direct"):
        pipe = env.host_pipe(threads=4, depth=2, mode=mode)
        for t in range(6):
            pipe.submit(a_host[t])
            if t:
                pipe.collect()
        pipe.collect()
    if env_id not in ("MortarMayhemB-v0", "Endless-MysteryPath-v0"):
        dbg = mg.make(env_id, num_envs=40, device=dev, render_mode="debug_rgb_array")
        dbg.reset(seed=0, options=opts)
        for t in range(12):
            dbg.step(acts[t][:40].contiguous())
    torch.cuda.synchronize()
    print("ran", env_id, "dones", int(dn.sum()), flush=True)
