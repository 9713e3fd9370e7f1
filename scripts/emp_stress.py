"""Endless-MysteryPath, look-ahead on vs off: two envs run the same random mix of launch shapes (rollouts of random
length, single steps, step_many, masked resets with new seeds, full resets, state snapshots restored later) — one
with helper warps (paths generated ahead, chains in global memory), one without (MG_FLAG_NO_HELPERS: every reset
generates inline). Observations, rewards, dones and records must be identical after every operation.
Usage: python scripts/emp_stress.py [n] [operations] [seed]"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "endless-memory-gym_b200"))
import numpy as np
import torch
import memory_gym_b200 as mg
from memory_gym_b200 import _lib as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
ops = int(sys.argv[2]) if len(sys.argv) > 2 else 300
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
a = mg.make("Endless-MysteryPath-v0", num_envs=n, device=dev)
b = mg.make("Endless-MysteryPath-v0", num_envs=n, device=dev)
b._knobs |= L.MG_FLAG_NO_HELPERS
b._io.flags |= L.MG_FLAG_NO_HELPERS
rng = np.random.default_rng(seed)
g = torch.Generator(device=dev).manual_seed(seed)
for e in (a, b):
    e.reset(seed=11)
snap = None
steps = resets = 0
for op in range(ops):
    kind = rng.choice(["rollout", "rollout", "rollout", "step", "step_many", "mask_reset", "full_reset", "snapshot", "restore"],
                      p=[0.3, 0.15, 0.1, 0.15, 0.1, 0.08, 0.04, 0.04, 0.04])
    if kind in ("rollout", "step_many"):
        K = int(rng.choice([1, 2, 3, 5, 8, 13, 20, 33, 64]))
        acts = torch.randint(0, 4, (K, n), generator=g, device=dev)
        out = []
        for e in (a, b):
            ring = torch.zeros((min(K, 4), n, 84, 84, 3), dtype=torch.uint8, device=dev)
            rew = torch.zeros((K, n), dtype=torch.float64, device=dev)
            dn = torch.zeros((K, n), dtype=torch.bool, device=dev)
            (e.rollout if kind == "rollout" else e.step_many)(acts, ring, rew, dn)
            out.append((ring, rew, dn))
        assert all(torch.equal(x, y) for x, y in zip(*out)), (op, kind, K)
        steps += K
        resets += int(out[0][2].sum())
    elif kind == "step":
        for _ in range(int(rng.integers(1, 12))):
            acts = torch.randint(0, 4, (n,), generator=g, device=dev)
            ra, rb = a.step(acts), b.step(acts)
            assert torch.equal(ra[0], rb[0]) and torch.equal(ra[1], rb[1]) and torch.equal(ra[2], rb[2]), (op, kind)
            steps += 1
            resets += int(ra[2].sum())
    elif kind == "mask_reset":
        mask = torch.rand(n, generator=g, device=dev) < 0.2
        s = int(rng.integers(0, 1000))
        ra, rb = a.reset(seed=s, mask=mask), b.reset(seed=s, mask=mask)
        assert torch.equal(ra[0], rb[0]), (op, kind)
    elif kind == "full_reset":
        s = int(rng.integers(0, 1000)) if rng.random() < 0.5 else None
        ra, rb = a.reset(seed=s), b.reset(seed=s)
        assert torch.equal(ra[0], rb[0]), (op, kind)
    elif kind == "snapshot":
        snap = (a.state_dict(), b.state_dict())
    elif kind == "restore" and snap is not None:
        a.load_state_dict(snap[0])
        b.load_state_dict(snap[1])
    assert torch.equal(a.records, b.records), (op, kind)
torch.cuda.synchronize()
w = a.state[:, 2048:2112].clone().view(torch.int32)
print(f"ok: {ops} operations, {steps} steps, {resets} auto-resets, adopted {int(w[:, 2].sum())}, missed empty {int(w[:, 14].sum())} "
      f"stale {int(w[:, 15].sum())}; without helpers: adopted {int(b.state[:, 2048:2112].clone().view(torch.int32)[:, 2].sum())}")
