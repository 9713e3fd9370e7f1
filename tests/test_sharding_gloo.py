"""N>1 path on CPU: two processes (gloo, 127.0.0.1), each owning a contiguous shard of the env batch
(`make_sharded` semantics: global env g is seeded seed+g), gathered and compared with the one-shard run.
The shards run on the hostsim backend (the kernel's logic headers compiled for the CPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(env_id, total, world, rank, steps, acts):
    import memory_gym_b200 as mg
    from hostsim_util import HostSimBackend
    n = total // world
    env = mg.MemoryGymVecEnv(env_id, n, device="cpu", env_offset=rank * n, _backend=HostSimBackend())
    obs, _ = env.reset(seed=42)
    out = [obs.clone()]
    rews, dones = [], []
    for t in range(steps):
        o, r, d, _, _ = env.step(acts[t, rank * n:(rank + 1) * n].contiguous())
        rews.append(r.clone())
        dones.append(d.clone())
    out.append(o.clone())
    return out, torch.stack(rews), torch.stack(dones)


def _worker(rank, world, port, env_id, total, steps, acts, result_path):
    for p in (ROOT, os.path.join(ROOT, "endless-memory-gym_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    (obs0, obs1), rew, dn = _run(env_id, total, world, rank, steps, acts)
    gathered = {}
    for name, t in (("obs0", obs0), ("obs1", obs1), ("rew", rew), ("dn", dn.to(torch.uint8))):
        bufs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(bufs, t.contiguous())
        gathered[name] = torch.cat(bufs, dim=0 if name.startswith("obs") else 1)
    # off the step path: episode statistics reduced across shards (what a trainer would log)
    stats = torch.tensor([float(dn.sum()), float(rew.sum())], dtype=torch.float64)
    dist.all_reduce(stats)
    if rank == 0:
        torch.save({**gathered, "stats": stats}, result_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("env_id", ["MortarMayhem-v0", "Endless-MortarMayhem-v0"])
def test_two_shards_equal_one(env_id, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    total, steps, world = 12, 70, 2
    rng = np.random.default_rng(3)
    acts = torch.as_tensor(rng.integers(0, 3, (steps, total, 2)))
    (obs0, obs1), rew, dn = _run(env_id, total, 1, 0, steps, acts)
    path = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(world, _free_port(), env_id, total, steps, acts, path), nprocs=world, join=True)
    got = torch.load(path)
    assert torch.equal(got["obs0"], obs0) and torch.equal(got["obs1"], obs1)
    assert torch.equal(got["rew"], rew) and torch.equal(got["dn"].bool(), dn)
    assert got["stats"][0] == float(dn.sum()) and dn.any()
