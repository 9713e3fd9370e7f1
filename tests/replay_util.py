"""Shared parity drivers: replay golden reference trajectories / oracle roll-outs through the engine.

Used by the CPU tier (hostsim backend: the kernels' logic headers compiled for the host) and by the
GPU tier (the real C-ABI on cuda) — same assertions, different backend."""
import numpy as np
import torch

import memory_gym_b200 as mg
from golden_util import action_for, crc, load_case, visual
from oracle import memgym_oracle as O


def make_single(env_id, device, backend):
    return mg.MemoryGymEnv(env_id, device=device, _backend=backend)


def replay_golden_single(name, device, backend, max_envs=None):
    """Reference protocol (step until done, reset(seed=None) continues the stream), one env at a time."""
    case = load_case(name)
    n_envs, steps = case["done"].shape
    infos = {(r["env"], r["t"]): r for r in case["infos"]}
    for e in range(n_envs if max_envs is None else min(n_envs, max_envs)):
        env = make_single(case["env_id"], device, backend)
        obs, info = env.reset(seed=case["seed0"] + e, options=case["options"])
        assert np.array_equal(visual(obs), case["frames"][case["frame_map"][(e, -1)]]), (name, e, "reset frame")
        if "gt" in case:
            assert np.array_equal(info["ground_truth"], case["gt_reset"][e]), (name, e, "reset ground truth")
        for t in range(steps):
            obs, reward, done, trunc, info = env.step(action_for(case, e, t))
            assert trunc is False
            assert float(reward) == case["reward"][e, t], (name, e, t, reward, case["reward"][e, t])
            assert bool(done) == bool(case["done"][e, t]), (name, e, t)
            if (e, t) in case["frame_map"]:
                want = case["frames"][case["frame_map"][(e, t)]]
                assert np.array_equal(visual(obs), want), (name, e, t, int((visual(obs) != want).sum()))
            assert crc(obs) == int(case["crc"][e, t]), (name, e, t)
            if "gt" in case:
                assert np.array_equal(info["ground_truth"], case["gt"][e, t]), (name, e, t, "ground truth")
            if done:
                ref = infos[(e, t)]
                assert float(info["reward"]) == ref["reward_seq_sum"], (name, e, t)
                for k, v in ref.items():
                    if k in ("env", "t", "reward_seq_sum"):
                        continue
                    assert np.array_equal(np.asarray(info[k], dtype=np.float64), np.asarray(v)), (name, k, e, t)
                obs, _ = env.reset(options=case["options"])
                assert crc(obs) == int(case["reset_crc"][e, t]), (name, e, t, "reset after done")


def replay_golden_batched(name, device, backend):
    """All golden envs of a case as ONE batch with auto-reset + final_obs: obs must be the reference's
    reset frame on done steps, final_observation its terminal frame."""
    case = load_case(name)
    n_envs, steps = case["done"].shape
    env = mg.MemoryGymVecEnv(case["env_id"], n_envs, device=device, auto_reset=True, final_obs=True, _backend=backend)
    seeds = [case["seed0"] + e for e in range(n_envs)]
    def to_np(o):  # Dict observations (MortarMayhemB) -> per-env dicts
        if isinstance(o, dict):
            vis, vec = o["visual_observation"].cpu().numpy(), o["vector_observation"].cpu().numpy()
            return [{"visual_observation": vis[i], "vector_observation": vec[i]} for i in range(vis.shape[0])]
        return o.cpu().numpy()

    obs, _ = env.reset(seed=seeds, options=case["options"])
    obs = to_np(obs)
    for e in range(n_envs):
        assert np.array_equal(visual(obs[e]), case["frames"][case["frame_map"][(e, -1)]]), (name, e)
    act_shape = env._actions_dev.shape
    infos = {(r["env"], r["t"]): r for r in case["infos"]}
    for t in range(steps):
        a = torch.as_tensor(case["action"][:, t].reshape(act_shape), device=device)
        obs, reward, done, trunc, info = env.step(a)
        obs, reward, done = to_np(obs), reward.cpu().numpy(), done.cpu().numpy()
        fin = info["final_observation"].cpu().numpy()
        ep = info["episode"].cpu().numpy()
        assert not trunc.any()
        assert np.array_equal(reward, case["reward"][:, t]), (name, t)
        assert np.array_equal(done, case["done"][:, t].astype(bool)), (name, t)
        if "gt" in case:  # done envs: ground truth of the terminal step; obs/ground_truth belong to the new episode
            gt, fgt = info["ground_truth"].cpu().numpy(), info["final_ground_truth"].cpu().numpy()
            for e in range(n_envs):
                assert np.array_equal(fgt[e] if done[e] else gt[e], case["gt"][e, t]), (name, e, t, "ground truth")
        for e in range(n_envs):
            if done[e]:
                if not env.task_b:  # (the terminal frame's CRC of a Dict obs also covers the old command vector)
                    assert crc(fin[e]) == int(case["crc"][e, t]), (name, e, t, "final obs")
                assert crc(obs[e]) == int(case["reset_crc"][e, t]), (name, e, t, "auto-reset obs")
                ref = infos[(e, t)]
                assert ep[e, 0] == ref["reward_seq_sum"] and ep[e, 1] == ref["length"], (name, e, t)
                for k, col in env.spec_.info_keys.items():
                    if k in ref:
                        assert ep[e, col] == np.asarray(ref[k], dtype=np.float64), (name, k, e, t)
            else:
                assert crc(obs[e]) == int(case["crc"][e, t]), (name, e, t)


def rollout_vs_oracle(env_id, options, n_envs, steps, seed0, device, backend, act_seed=7, warps_per_cta=0):
    """Batched engine with auto-reset vs n_envs independent oracle instances (manual reset)."""
    env = mg.MemoryGymVecEnv(env_id, n_envs, device=device, auto_reset=True, final_obs=True, _backend=backend,
                             warps_per_cta=warps_per_cta)
    oras = [O.make(env_id) for _ in range(n_envs)]
    def to_np(o):
        if isinstance(o, dict):
            vis, vec = o["visual_observation"].cpu().numpy(), o["vector_observation"].cpu().numpy()
            return [{"visual_observation": vis[i], "vector_observation": vec[i]} for i in range(vis.shape[0])]
        return o.cpu().numpy()

    def same(a, b):
        if isinstance(a, dict):
            return np.array_equal(a["visual_observation"], b["visual_observation"]) and \
                np.array_equal(a["vector_observation"], b["vector_observation"])
        return np.array_equal(a, b)

    obs, _ = env.reset(seed=seed0, options=options)
    obs = to_np(obs)
    for e, o in enumerate(oras):
        want, _ = o.reset(seed=seed0 + e, options=options)
        assert same(obs[e], want), (env_id, e, "reset")
    rng = np.random.default_rng(act_seed)
    nvec = env.spec_.action_nvec
    episodes = 0
    for t in range(steps):
        a = rng.integers(0, nvec[0], size=env._actions_dev.shape)
        obs, reward, done, trunc, info = env.step(torch.as_tensor(a, device=device))
        obs, reward, done = to_np(obs), reward.cpu().numpy(), done.cpu().numpy()
        fin, ep = info["final_observation"].cpu().numpy(), info["episode"].cpu().numpy()
        for e, o in enumerate(oras):
            w_obs, w_r, w_d, _, w_info = o.step(a[e] if len(nvec) > 1 else int(a[e]))
            assert float(w_r) == reward[e], (env_id, e, t, w_r, reward[e])
            assert bool(w_d) == bool(done[e]), (env_id, e, t)
            if w_d:
                episodes += 1
                assert np.array_equal(fin[e], visual(w_obs)), (env_id, e, t, "terminal frame")
                assert ep[e, 0] == float(w_info["reward"]) and ep[e, 1] == w_info["length"]
                w_obs, _ = o.reset(options=options)
            assert same(obs[e], w_obs), (env_id, e, t)
    return episodes
