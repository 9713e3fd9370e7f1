"""memory_gym_b200 — B200-native batched Memory Gym (Mortar Mayhem, Mystery Path, Searing Spotlights).

Drop-in for the hot path of MarcoMeter/endless-memory-gym: the same env ids, reset-option dicts,
action encodings and ``[x][y][c]`` uint8 observations, for N parallel instances per GPU.

    import memory_gym_b200 as mg
    env = mg.make("MortarMayhem-v0", num_envs=4096, device="cuda:0")
    obs, info = env.reset(seed=0)
    obs, reward, done, truncated, info = env.step(actions)   # int64[N, 2] on the device

``mg.make(id)`` with the default ``num_envs=None`` returns the single-instance wrapper that follows the
reference protocol (numpy obs, manual reset after done).
"""
from __future__ import annotations

from . import _lib
from ._lib import LIB_PATH, MemGymError
from .specs import NOT_BUILT, SPECS, spec
from .vec_env import HostPipe, MemoryGymEnv, MemoryGymVecEnv

__all__ = ["make", "make_sharded", "MemoryGymVecEnv", "MemoryGymEnv", "HostPipe", "MemGymError", "SPECS", "NOT_BUILT", "spec",
           "LIB_PATH", "registered_ids"]


def registered_ids():
    return sorted(SPECS)


def make(env_id, num_envs=None, device="cuda", **kwargs):
    """``num_envs=None`` -> single-instance reference-protocol env; otherwise the batched engine."""
    if num_envs is None:
        return MemoryGymEnv(env_id, device=device, **kwargs)
    return MemoryGymVecEnv(env_id, num_envs, device=device, **kwargs)


def make_sharded(env_id, total_envs, rank, world_size, device=None, local_rank=None, **kwargs):
    """One shard of an env batch split over ``world_size`` GPUs (one process per GPU).

    Envs are independent, so a shard is just a contiguous slice ``[rank*n, (rank+1)*n)`` of the global
    batch; global env ``g`` is seeded with ``seed + g`` whatever the number of shards, so the
    concatenated result is identical to a single-GPU run. No collective is on the step path.

    ``device`` defaults to ``cuda:<local_rank>``; ``local_rank`` defaults to ``$LOCAL_RANK`` (torchrun) and,
    without it, to ``rank`` modulo the number of visible GPUs. The library launches on the device that owns
    the env records, so the caller's current CUDA device does not matter."""
    if total_envs % world_size:
        raise ValueError("total_envs must be divisible by world_size")
    n = total_envs // world_size
    if device is None:
        if local_rank is None:
            import os
            import torch
            env_lr = os.environ.get("LOCAL_RANK")
            local_rank = int(env_lr) if env_lr is not None else rank % max(1, torch.cuda.device_count())
        device = f"cuda:{int(local_rank)}"
    return MemoryGymVecEnv(env_id, n, device=device, env_offset=rank * n, **kwargs)


def _register_with_gymnasium():
    """Register the reference's own id strings (memory_gym/__init__.py:13-60) with gymnasium when it is importable:
    ``gym.make("MortarMayhem-v0")`` then returns the single-instance, reference-protocol env of this package. An id that
    is already registered (the reference package was imported first) is left alone; ``B200-<id>`` always names this
    implementation."""
    try:
        from gymnasium.envs.registration import register, registry
    except Exception:  # noqa: BLE001  (gymnasium is absent from the build image)
        return []
    done = []
    for env_id in SPECS:
        for gid in (env_id, "B200-" + env_id):
            if gid not in registry:
                register(id=gid, entry_point=lambda env_id=env_id, **kw: MemoryGymEnv(env_id, **kw))
                done.append(gid)
    return done


_register_with_gymnasium()
