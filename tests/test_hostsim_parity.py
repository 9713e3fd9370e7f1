"""CPU tier: the env logic + frame composition headers that the CUDA kernels are built from
(csrc/mg_*_logic.h, mg_frame.h, mg_policy.h), compiled for the host by tests/hostsim, reproduce the
UNMODIFIED reference's golden trajectories and the oracle bit for bit. (The GPU tier repeats the same
assertions through the real C-ABI on cuda: tests/test_gpu_parity.py.)"""
import pytest

from hostsim_util import HostSimBackend
from replay_util import replay_golden_batched, replay_golden_single, rollout_vs_oracle

CASES = ["mm", "mm_opts", "mmgrid", "emm", "emm_opts", "mp", "mp_opts", "ss", "ss_opts", "ss_black",
         "mmb", "mmb_opts", "mmbgrid", "mpgrid", "mpgrid_opts", "ess", "ess_opts", "ess_black", "ess_nocoin", "ss_dim0", "ess_dim0",
         "emp", "emp_opts", "emp_nopast", "emp_cam"]


@pytest.fixture(scope="module")
def backend():
    return HostSimBackend()


@pytest.mark.parametrize("name", CASES)
def test_golden_single(name, backend):
    replay_golden_single(name, "cpu", backend)


@pytest.mark.parametrize("name", CASES)
def test_golden_batched_autoreset(name, backend):
    replay_golden_batched(name, "cpu", backend)


@pytest.mark.parametrize("env_id,options", [
    ("MortarMayhem-v0", None),
    ("MortarMayhem-v0", {"arena_size": 2, "command_count": [1, 2], "explosion_delay": [3], "explosion_duration": [1],
                         "command_show_duration": [1], "command_show_delay": [0], "agent_scale": 0.5}),
    ("MortarMayhem-v0", {"arena_size": 6, "allowed_commands": 4, "agent_speed": 1.0}),
    ("MortarMayhem-Grid-v0", {"arena_size": 3, "command_count": [2]}),
    ("Endless-MortarMayhem-v0", {"explosion_delay": [4], "explosion_duration": [2], "agent_speed": 5.0}),
    ("MysteryPath-v0", {"max_steps": 40, "show_goal": True}),
    ("MysteryPath-v0", {"visual_feedback": False, "agent_scale": 0.3, "max_steps": 64}),
    ("MortarMayhemB-v0", {"command_count": [1, 3], "explosion_delay": [4], "explosion_duration": [2]}),
    ("MortarMayhemB-Grid-v0", {"arena_size": 4, "command_count": [2, 20]}),
    ("MysteryPath-Grid-v0", {"max_steps": 30, "show_origin": True}),
    ("SearingSpotlights-v0", {"max_steps": 60}),
    ("SearingSpotlights-v0", {"max_steps": 50, "coins_visible": True, "exit_visible": True, "agent_visible": True,
                              "num_coins": [2, 4], "agent_health": 3, "spot_damage": 0.7, "initial_spawns": 6,
                              "spot_min_radius": 12.0, "spot_max_radius": 20.0, "spot_min_speed": 0.01,
                              "spot_max_speed": 0.03, "reward_inside_spotlight": -0.1, "reward_outside_spotlight": 0.01,
                              "reward_death": -1.0}),
    ("SearingSpotlights-v0", {"max_steps": 45, "exit_visible": True, "hide_chessboard": True, "light_threshold": 100,
                              "light_dim_off_duration": 3, "show_last_action": False, "visual_feedback": False,
                              "initial_spawn_interval": 3, "spawn_interval_threshold": 2, "num_spawns": 9}),
    ("SearingSpotlights-v0", {"max_steps": 40, "sample_agent_position": False, "light_dim_off_duration": 0,
                              "show_last_positive_reward": False, "agent_visible": True, "coin_scale": 1.0,
                              "exit_scale": 1.0, "agent_scale": 0.4, "agent_speed": 5.0, "num_coins": [3]}),
    ("Endless-MysteryPath-v0", None),
    ("Endless-MysteryPath-v0", {"camera_offset_scale": 0.0, "stamina_level": 9, "reward_fall_off": -0.3,
                                "reward_path_progress_dense": 0.02, "reward_step": -0.001}),
    ("Endless-MysteryPath-v0", {"camera_offset_scale": 1.7, "show_background": True, "show_stamina": True,
                                "agent_speed": 13.0, "max_steps": 30, "agent_scale": 0.5}),
    ("Endless-MysteryPath-v0", {"camera_offset_scale": 9.0, "agent_speed": 7.0, "stamina_level": 35,
                                "show_stamina": True, "visual_feedback": False}),
    ("Endless-SearingSpotlights-v0", {"steps_per_coin": 40}),
    ("Endless-SearingSpotlights-v0", {"max_steps": 55, "coins_visible": True, "agent_visible": True, "agent_health": 3,
                                      "spot_damage": 0.7, "initial_spawns": 6, "spawn_interval": 7,
                                      "spot_min_radius": 12.0, "spot_max_radius": 20.0, "spot_min_speed": 0.01,
                                      "spot_max_speed": 0.03, "reward_inside_spotlight": -0.1,
                                      "reward_outside_spotlight": 0.01, "reward_death": -1.0, "coin_scale": 1.0}),
    ("Endless-SearingSpotlights-v0", {"steps_per_coin": 30, "hide_chessboard": True, "light_threshold": 100,
                                      "light_dim_off_duration": 3, "visual_feedback": False, "coin_show_duration": 20,
                                      "sample_agent_position": False, "agent_scale": 0.4, "agent_speed": 5.0}),
    ("Endless-SearingSpotlights-v0", {"coin_enabled": False, "max_steps": 35, "agent_visible": True,
                                      "coins_visible": True, "light_dim_off_duration": 0}),
])
def test_rollout_vs_oracle(env_id, options, backend):
    episodes = rollout_vs_oracle(env_id, options, n_envs=8, steps=150, seed0=1000, device="cpu", backend=backend)
    assert episodes > 0


@pytest.mark.parametrize("env_id,opts", [("MortarMayhem-v0", None), ("MysteryPath-v0", {"max_steps": 50}),
                                         ("SearingSpotlights-v0", {"max_steps": 60}),
                                         ("SearingSpotlights-v0", {"max_steps": 45, "agent_visible": True,
                                                                   "light_threshold": 90, "num_coins": [3]}),
                                         ("Endless-SearingSpotlights-v0", {"steps_per_coin": 40, "agent_speed": 5.0}),
                                         ("Endless-MysteryPath-v0", {"show_stamina": True}),
                                         ("Endless-MysteryPath-v0", {"show_background": True, "agent_speed": 5.0})])
def test_rollout_equals_step_loop_hostsim(env_id, opts, backend):
    """rollout() (env-major: all T steps of an env back to back, persistent per-warp scratch — exercises the
    incremental frame paths; slot t % ring, per-step reward/done/info rows) == T step() calls."""
    import numpy as np
    import torch
    import memory_gym_b200 as mg
    n, T = 6, 130
    a = mg.MemoryGymVecEnv(env_id, n, device="cpu", _backend=backend)
    b = mg.MemoryGymVecEnv(env_id, n, device="cpu", _backend=backend)
    nvec = mg.spec(env_id).action_nvec
    acts = torch.as_tensor(np.random.default_rng(1).integers(0, nvec[0], (T, n, 2) if len(nvec) > 1 else (T, n)))
    a.reset(seed=9, options=opts)
    b.reset(seed=9, options=opts)
    ring = torch.zeros((4, n, 84, 84, 3), dtype=torch.uint8)
    rew, dn = torch.zeros((T, n), dtype=torch.float64), torch.zeros((T, n), dtype=torch.bool)
    inf = torch.zeros((T, n, mg._lib.MG_INFO_COLS), dtype=torch.float64)
    full = torch.zeros((T, n, 84, 84, 3), dtype=torch.uint8)
    a.rollout(acts, full, rew, dn, inf)
    a2 = mg.MemoryGymVecEnv(env_id, n, device="cpu", _backend=backend)
    a2.reset(seed=9, options=opts)
    a2.rollout(acts, ring, torch.zeros_like(rew), torch.zeros_like(dn), None)
    for t in range(T):
        o, r, d, _, info = b.step(acts[t])
        assert torch.equal(r, rew[t]) and torch.equal(d, dn[t])
        assert torch.equal(info["episode"][d], inf[t][d])
        assert torch.equal(o, full[t]), (t, int((o != full[t]).sum()))
        if t >= T - 4:
            assert torch.equal(o, ring[t % 4])
    assert dn.any() and torch.equal(a.records, b.records)
    if env_id == "Endless-MysteryPath-v0":
        # the rollout's resets adopted paths generated ahead by the "helper lane" (MgEPSpecRec: hword, tword, adopted),
        # the step loop generated all of its paths inline: same bytes either way
        words = a.state[:, 2048:2060].contiguous().view(torch.int32)
        assert (words[:, 2] >= 1).all() and int(words[:, 2].sum()) >= 2 * n and ((words[:, 0] >> 13) >= 1).all()
        assert backend.sim.mgsim_ep_resumed() > 10  # searches stopped mid-way were parked and resumed
        assert (b.state[:, 2048:] == 0).all()


def test_emp_lookahead_chain_outlives_its_counters_hostsim(backend):
    """Endless-MysteryPath: a look-ahead chain lives for hundreds of episodes, its produced / consumed counters are
    8 bits wide. Past the wrap the helper must still continue from the previous episode's stream (not from the chain's
    base): no reset may ever find an episode generated from another stream, and the bytes stay those of the step loop."""
    import numpy as np
    import torch
    import memory_gym_b200 as mg
    env_id, n, T = "Endless-MysteryPath-v0", 2, 9000
    a = mg.MemoryGymVecEnv(env_id, n, device="cpu", _backend=backend)
    b = mg.MemoryGymVecEnv(env_id, n, device="cpu", _backend=backend)
    acts = torch.as_tensor(np.random.default_rng(5).integers(0, 4, (T, n)))
    a.reset(seed=3)
    b.reset(seed=3)
    ring = torch.zeros((4, n, 84, 84, 3), dtype=torch.uint8)
    rew, dn = torch.zeros((T, n), dtype=torch.float64), torch.zeros((T, n), dtype=torch.bool)
    inf = torch.zeros((T, n, mg._lib.MG_INFO_COLS), dtype=torch.float64)
    a.rollout(acts, ring, rew, dn, inf)
    for t in range(T):
        o, r, d, _, info = b.step(acts[t])
        assert torch.equal(r, rew[t]) and torch.equal(d, dn[t]), t
        assert torch.equal(info["episode"][d], inf[t][d]), t
    assert torch.equal(o, ring[(T - 1) % 4]) and torch.equal(a.records, b.records)
    words = a.state[:, 2048:2112].contiguous().view(torch.int32)
    assert (dn.sum(0) > 280).all()                  # more episodes per env than the counters can count
    assert (words[:, 2] > 256).all()                # ... most of them adopted from ONE long-lived chain
    assert (words[:, 15] == 0).all(), words[:, 12:16]  # missed_stale: never an episode from the wrong stream
