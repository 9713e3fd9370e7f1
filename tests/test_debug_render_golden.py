"""`debug_rgb_array` (the reference's ground-truth render, `_build_debug_surface` + `transform.scale`) against renders of
the UNMODIFIED reference (tests/golden/ref_dbg_*.npz, oracle/gen_golden_debug.py): the oracle restatement here, the
library on the GPU in test_gpu_parity.py."""
import json
import os
import zlib

import numpy as np
import pytest

from oracle import memgym_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DBG_CASES = ["mm", "mm_opts", "mmgrid", "emm", "mp", "mp_opts", "mpgrid", "ss", "ss_opts", "ss_black", "ess",
             "ess_nocoin", "emp", "emp_opts"]


def load_dbg(name):
    z = np.load(os.path.join(GOLDEN, f"ref_dbg_{name}.npz"))
    case = {k: z[k] for k in z.files}
    case["options"] = json.loads(str(z["options_json"]))
    case["env_id"] = str(z["env_id"])
    case["seed0"] = int(z["seed0"])
    case["frame_map"] = {(int(e), int(t)): i for i, (e, t) in enumerate(z["frame_idx"])}
    return case


def check_frame(case, e, t, fr, want_crc):
    """`fr`: uint8[336][336][3] ([y][x][c]). Full-frame CRC, and a pixel-level diff where the golden kept the frame."""
    key = (e, t)
    if key in case["frame_map"]:
        want = case["frames"][case["frame_map"][key]]
        got = fr[::4, ::4]
        if not np.array_equal(got, want):
            bad = np.argwhere((got != want).any(-1))
            raise AssertionError(f"debug frame env {e} t {t}: {len(bad)} px differ, first (y, x) = {bad[:5].tolist()}, "
                                 f"got {got[tuple(bad[0])]} want {want[tuple(bad[0])]}")
    assert zlib.crc32(np.ascontiguousarray(fr).tobytes()) == int(want_crc), (e, t)


@pytest.mark.parametrize("name", DBG_CASES)
def test_oracle_debug_render_matches_reference(name):
    case = load_dbg(name)
    n_envs, steps = case["done"].shape
    for e in range(n_envs):
        env = O.make(case["env_id"], render_mode="debug_rgb_array")
        env.reset(seed=case["seed0"] + e, options=case["options"])
        check_frame(case, e, -1, env.render(), case["first_dbg_crc"][e])
        for t in range(steps):
            a = case["action"][e, t]
            obs, reward, done, _, _ = env.step(int(a[0]) if a.shape[0] == 1 else a)
            assert zlib.crc32(np.ascontiguousarray(obs).tobytes()) == int(case["crc"][e, t])
            assert bool(done) == bool(case["done"][e, t])
            check_frame(case, e, t, env.render(), case["dbg_crc"][e, t])
            if done:
                env.reset(options=case["options"])
                check_frame(case, e, -2 - t, env.render(), case["reset_dbg_crc"][e, t])


def replay_debug_single(name, device, backend, max_envs=None):
    """Reference protocol through the engine (single-instance wrapper): render() after every reset / step."""
    import memory_gym_b200 as mg
    case = load_dbg(name)
    n_envs, steps = case["done"].shape
    for e in range(n_envs if max_envs is None else min(n_envs, max_envs)):
        env = mg.MemoryGymEnv(case["env_id"], device=device, render_mode="debug_rgb_array", _backend=backend)
        env.reset(seed=case["seed0"] + e, options=case["options"])
        check_frame(case, e, -1, env.render(), case["first_dbg_crc"][e])
        for t in range(steps):
            a = case["action"][e, t]
            obs, reward, done, _, _ = env.step(int(a[0]) if a.shape[0] == 1 else a)
            assert zlib.crc32(np.ascontiguousarray(obs).tobytes()) == int(case["crc"][e, t]), (name, e, t)
            assert bool(done) == bool(case["done"][e, t])
            check_frame(case, e, t, env.render(), case["dbg_crc"][e, t])
            if done:
                env.reset(options=case["options"])
                check_frame(case, e, -2 - t, env.render(), case["reset_dbg_crc"][e, t])


def replay_debug_batched(name, device, backend):
    """All golden envs as one auto-resetting batch: on done steps the debug frame is the reset frame's."""
    import torch
    import memory_gym_b200 as mg
    case = load_dbg(name)
    n_envs, steps = case["done"].shape
    env = mg.MemoryGymVecEnv(case["env_id"], n_envs, device=device, render_mode="debug_rgb_array", _backend=backend)
    env.reset(seed=[case["seed0"] + e for e in range(n_envs)], options=case["options"])
    dbg = env.render().cpu().numpy()
    for e in range(n_envs):
        check_frame(case, e, -1, dbg[e], case["first_dbg_crc"][e])
    for t in range(steps):
        acts = torch.as_tensor(case["action"][:, t]).to(device)
        _, _, done, _, _ = env.step(acts[:, 0].contiguous() if acts.shape[1] == 1 else acts.contiguous())
        dbg, done = env.render().cpu().numpy(), done.cpu().numpy()
        for e in range(n_envs):
            assert bool(done[e]) == bool(case["done"][e, t]), (name, e, t)
            if done[e]:
                check_frame(case, e, -2 - t, dbg[e], case["reset_dbg_crc"][e, t])
            else:
                check_frame(case, e, t, dbg[e], case["dbg_crc"][e, t])


@pytest.fixture(scope="module")
def hostsim():
    import hostsim_util
    return hostsim_util.HostSimBackend()


@pytest.mark.parametrize("name", DBG_CASES)
def test_hostsim_debug_render_matches_reference(name, hostsim):
    """The kernels' own headers (mg_debug.h + logic), compiled for the CPU by tests/hostsim."""
    replay_debug_single(name, "cpu", hostsim, max_envs=2)
    replay_debug_batched(name, "cpu", hostsim)


@pytest.mark.gpu
@pytest.mark.parametrize("name", DBG_CASES)
def test_gpu_debug_render_matches_reference(name):
    replay_debug_single(name, "cuda:0", None, max_envs=2)
    replay_debug_batched(name, "cuda:0", None)


def test_debug_render_mode_surface():
    import memory_gym_b200 as mg
    assert "debug_rgb_array" in mg.MemoryGymVecEnv.metadata["render_modes"]
    import hostsim_util
    be = hostsim_util.HostSimBackend()
    for env_id in ("MortarMayhemB-v0", "MortarMayhemB-Grid-v0"):
        with pytest.raises(mg.MemGymError):
            mg.MemoryGymVecEnv(env_id, 2, device="cpu", render_mode="debug_rgb_array", _backend=be)
    with pytest.raises(ValueError):
        mg.MemoryGymVecEnv("MortarMayhem-v0", 2, device="cpu", render_mode="human", _backend=be)
