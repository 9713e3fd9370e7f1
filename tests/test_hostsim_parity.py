"""CPU tier: the env logic + frame composition headers that the CUDA kernels are built from
(csrc/mg_*_logic.h, mg_frame.h, mg_policy.h), compiled for the host by tests/hostsim, reproduce the
UNMODIFIED reference's golden trajectories and the oracle bit for bit. (The GPU tier repeats the same
assertions through the real C-ABI on cuda: tests/test_gpu_parity.py.)"""
import pytest

from hostsim_util import HostSimBackend
from replay_util import replay_golden_batched, replay_golden_single, rollout_vs_oracle

CASES = ["mm", "mm_opts", "mmgrid", "emm", "emm_opts", "mp", "mp_opts"]


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
])
def test_rollout_vs_oracle(env_id, options, backend):
    episodes = rollout_vs_oracle(env_id, options, n_envs=8, steps=150, seed0=1000, device="cpu", backend=backend)
    assert episodes > 0


def test_rollout_equals_step_loop_hostsim(backend):
    """rollout() (slot t % ring, per-step reward/done/info rows) == T step() calls, on the host harness."""
    import numpy as np
    import torch
    import memory_gym_b200 as mg
    n, T = 6, 130
    a = mg.MemoryGymVecEnv("MortarMayhem-v0", n, device="cpu", _backend=backend)
    b = mg.MemoryGymVecEnv("MortarMayhem-v0", n, device="cpu", _backend=backend)
    acts = torch.as_tensor(np.random.default_rng(1).integers(0, 3, (T, n, 2)))
    a.reset(seed=9)
    b.reset(seed=9)
    ring = torch.zeros((4, n, 84, 84, 3), dtype=torch.uint8)
    rew, dn = torch.zeros((T, n), dtype=torch.float64), torch.zeros((T, n), dtype=torch.bool)
    inf = torch.zeros((T, n, 8), dtype=torch.float64)
    a.rollout(acts, ring, rew, dn, inf)
    for t in range(T):
        o, r, d, _, info = b.step(acts[t])
        assert torch.equal(r, rew[t]) and torch.equal(d, dn[t])
        assert torch.equal(info["episode"][d], inf[t][d])
        if t >= T - 4:
            assert torch.equal(o, ring[t % 4])
    assert dn.any() and torch.equal(a.state, b.state)
