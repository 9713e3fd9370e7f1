"""GPU tier (-m gpu): the real product path — libmemgym_b200.so through its C ABI on cuda:0 — against
(a) golden trajectories of the UNMODIFIED reference, (b) the CPU oracle on seeded roll-outs,
(c) size-independent properties at BASELINE.json's full batch sizes. Bit-exact everywhere."""
import numpy as np
import pytest
import torch

import memory_gym_b200 as mg
from replay_util import replay_golden_batched, replay_golden_single, rollout_vs_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = ["mm", "mm_opts", "mmgrid", "emm", "emm_opts", "mp", "mp_opts", "ss", "ss_opts", "ss_black",
         "mmb", "mmb_opts", "mmbgrid", "mpgrid", "mpgrid_opts", "ess", "ess_opts", "ess_black", "ess_nocoin", "ss_dim0", "ess_dim0",
         "emp", "emp_opts", "emp_nopast", "emp_cam"]


@pytest.mark.parametrize("name", CASES)
def test_golden_single(name):
    replay_golden_single(name, DEV, None, max_envs=3)


@pytest.mark.parametrize("name", CASES)
def test_golden_batched_autoreset(name):
    replay_golden_batched(name, DEV, None)


@pytest.mark.parametrize("env_id,options,n_envs,wpc", [
    ("MortarMayhem-v0", None, 40, 0),
    ("MortarMayhem-v0", None, 333, 10),   # several envs per warp: exercises patch / un-patch of the frame buffer
    ("MortarMayhem-v0", {"arena_size": 2, "command_count": [1, 2], "explosion_delay": [3], "explosion_duration": [1],
                         "command_show_duration": [1], "command_show_delay": [0], "agent_scale": 0.5}, 64, 3),
    ("MortarMayhem-Grid-v0", {"arena_size": 3, "command_count": [2]}, 64, 0),
    ("Endless-MortarMayhem-v0", {"explosion_delay": [4], "explosion_duration": [2], "agent_speed": 5.0}, 64, 2),
    ("MysteryPath-v0", {"max_steps": 40, "show_goal": True, "show_origin": True}, 64, 1),
    ("MysteryPath-v0", {"visual_feedback": False, "agent_scale": 0.3, "max_steps": 64}, 64, 0),
    ("MortarMayhemB-v0", {"command_count": [1, 3], "explosion_delay": [4], "explosion_duration": [2]}, 64, 0),
    ("MortarMayhemB-Grid-v0", {"arena_size": 4, "command_count": [2, 20]}, 64, 3),
    ("MysteryPath-Grid-v0", {"max_steps": 30, "show_origin": True}, 64, 0),
    ("SearingSpotlights-v0", {"max_steps": 60}, 64, 0),
    ("SearingSpotlights-v0", {"max_steps": 50, "coins_visible": True, "exit_visible": True, "agent_visible": True,
                              "num_coins": [2, 4], "agent_health": 3, "spot_damage": 0.7, "initial_spawns": 6,
                              "spot_min_radius": 12.0, "spot_max_radius": 20.0, "spot_min_speed": 0.01,
                              "spot_max_speed": 0.03, "reward_inside_spotlight": -0.1, "reward_outside_spotlight": 0.01,
                              "reward_death": -1.0}, 48, 2),
    ("SearingSpotlights-v0", {"max_steps": 45, "exit_visible": True, "hide_chessboard": True, "light_threshold": 100,
                              "light_dim_off_duration": 3, "show_last_action": False, "visual_feedback": False,
                              "initial_spawn_interval": 3, "spawn_interval_threshold": 2, "num_spawns": 9}, 48, 1),
    ("SearingSpotlights-v0", {"max_steps": 40, "sample_agent_position": False, "light_dim_off_duration": 0,
                              "show_last_positive_reward": False, "agent_visible": True, "coin_scale": 1.0,
                              "exit_scale": 1.0, "agent_scale": 0.4, "agent_speed": 5.0, "num_coins": [3]}, 300, 8),
    ("Endless-MysteryPath-v0", None, 64, 0),
    ("Endless-MysteryPath-v0", {"camera_offset_scale": 0.0, "stamina_level": 9, "reward_fall_off": -0.3,
                                "reward_path_progress_dense": 0.02, "reward_step": -0.001}, 64, 3),
    ("Endless-MysteryPath-v0", {"camera_offset_scale": 1.7, "show_background": True, "show_stamina": True,
                                "agent_speed": 13.0, "max_steps": 30, "agent_scale": 0.5}, 300, 9),
    ("Endless-MysteryPath-v0", {"camera_offset_scale": 9.0, "agent_speed": 7.0, "stamina_level": 35,
                                "show_stamina": True, "visual_feedback": False}, 48, 1),
    ("Endless-SearingSpotlights-v0", {"steps_per_coin": 40}, 64, 0),
    ("Endless-SearingSpotlights-v0", {"max_steps": 55, "coins_visible": True, "agent_visible": True, "agent_health": 3,
                                      "spot_damage": 0.7, "initial_spawns": 6, "spawn_interval": 7,
                                      "spot_min_radius": 12.0, "spot_max_radius": 20.0, "spot_min_speed": 0.01,
                                      "spot_max_speed": 0.03, "reward_inside_spotlight": -0.1,
                                      "reward_outside_spotlight": 0.01, "reward_death": -1.0, "coin_scale": 1.0}, 48, 2),
    ("Endless-SearingSpotlights-v0", {"steps_per_coin": 30, "hide_chessboard": True, "light_threshold": 100,
                                      "light_dim_off_duration": 3, "visual_feedback": False, "coin_show_duration": 20,
                                      "sample_agent_position": False, "agent_scale": 0.4, "agent_speed": 5.0}, 300, 8),
    ("Endless-SearingSpotlights-v0", {"coin_enabled": False, "max_steps": 35, "agent_visible": True,
                                      "coins_visible": True, "light_dim_off_duration": 0}, 48, 1),
])
def test_rollout_vs_oracle(env_id, options, n_envs, wpc):
    steps = 70 if n_envs > 100 else 150
    episodes = rollout_vs_oracle(env_id, options, n_envs, steps, seed0=4000, device=DEV, backend=None, warps_per_cta=wpc)
    assert episodes > 0


def _random_actions(env, steps, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    nvec = env.spec_.action_nvec
    return torch.randint(0, nvec[0], (steps,) + tuple(env._actions_dev.shape), generator=g, device=DEV)


@pytest.mark.parametrize("env_id,n,options", [("MortarMayhem-v0", 4096, None),
                                              ("MysteryPath-v0", 8192, {"max_steps": 48}),
                                              ("Endless-MortarMayhem-v0", 8192, None),
                                              ("SearingSpotlights-v0", 8192, {"max_steps": 40}),
                                              ("Endless-SearingSpotlights-v0", 8192, {"steps_per_coin": 40}),
                                              ("Endless-MysteryPath-v0", 8192, None)])
def test_full_size_properties(env_id, n, options):
    """At BASELINE.json's batch sizes: (1) bit-identical to the same envs run in small shards with other
    launch geometries (1 / 3 warps per CTA) — env i's result depends only on seed0+i and its actions;
    (2) deterministic; (3) a sample of envs matches the CPU oracle."""
    steps = 64
    env = mg.make(env_id, num_envs=n, device=DEV)
    acts = _random_actions(env, steps, 1234)
    obs0 = env.reset(seed=0, options=options)[0].clone()
    crcs, rew, dn = [], [], []
    for t in range(steps):
        o, r, d, _, _ = env.step(acts[t])
        crcs.append(o.view(n, -1).to(torch.int64).mul_(torch.arange(1, 21169, device=DEV) % 251 + 1).sum(1))
        rew.append(r.clone())
        dn.append(d.clone())
    crcs, rew, dn = torch.stack(crcs), torch.stack(rew), torch.stack(dn)
    assert dn.any()
    # (2) determinism
    env2 = mg.make(env_id, num_envs=n, device=DEV, warps_per_cta=7)
    assert torch.equal(env2.reset(seed=0, options=options)[0], obs0)
    for t in range(steps):
        o, r, d, _, _ = env2.step(acts[t])
        assert torch.equal(r, rew[t]) and torch.equal(d, dn[t])
    assert torch.equal(o, env.obs)
    # (1) shards with other geometries
    for lo, cnt, wpc in ((0, 37, 1), (n // 2 - 5, 150, 3), (n - 64, 64, 10)):
        sh = mg.make(env_id, num_envs=cnt, device=DEV, env_offset=lo, warps_per_cta=wpc)
        assert torch.equal(sh.reset(seed=0, options=options)[0], obs0[lo:lo + cnt])
        for t in range(steps):
            o, r, d, _, _ = sh.step(acts[t, lo:lo + cnt].contiguous())
            c = o.view(cnt, -1).to(torch.int64).mul_(torch.arange(1, 21169, device=DEV) % 251 + 1).sum(1)
            assert torch.equal(c, crcs[t, lo:lo + cnt]), (lo, t)
            assert torch.equal(r, rew[t, lo:lo + cnt]) and torch.equal(d, dn[t, lo:lo + cnt])
    # (3) oracle on a sample
    from oracle import memgym_oracle as O
    a_cpu = acts.cpu().numpy()
    for e in (0, 1, n // 3, n - 1):
        ora = O.make(env_id)
        ora.reset(seed=e, options=options)
        for t in range(steps):
            w_obs, w_r, w_d, _, _ = ora.step(a_cpu[t, e])
            assert float(w_r) == float(rew[t, e]) and bool(w_d) == bool(dn[t, e]), (e, t)
            if w_d:
                w_obs, _ = ora.reset(options=options)
        assert np.array_equal(env.obs[e].cpu().numpy(), w_obs), e


def test_masked_reset_and_no_autoreset():
    env = mg.make("MortarMayhem-v0", num_envs=16, device=DEV, auto_reset=False)
    env.reset(seed=3)
    acts = _random_actions(env, 80, 5)
    for t in range(3000):
        obs, r, d, _, _ = env.step(acts[t % 80])
        if d.all():  # finished envs stay finished (done keeps being reported) until they are reset
            break
    assert d.all()
    before = obs.clone()
    mask = torch.zeros(16, dtype=torch.bool, device=DEV)
    mask[::2] = True
    obs, _ = env.reset(mask=mask)
    assert torch.equal(obs[1::2], before[1::2]) and not torch.equal(obs[::2], before[::2])
    obs, r, d, _, _ = env.step(acts[0])
    assert d[1::2].all() and not d[::2].any()


def test_state_dict_roundtrip():
    env = mg.make("Endless-MortarMayhem-v0", num_envs=32, device=DEV)
    env.reset(seed=11)
    acts = _random_actions(env, 60, 9)
    for t in range(20):
        env.step(acts[t])
    sd = env.state_dict()
    tail = [tuple(x.clone() for x in env.step(acts[t])[:3]) for t in range(20, 60)]
    env2 = mg.make("Endless-MortarMayhem-v0", num_envs=32, device=DEV)
    env2.load_state_dict(sd)
    for t in range(20, 60):
        o, r, d, _, _ = env2.step(acts[t])
        assert torch.equal(o, tail[t - 20][0]) and torch.equal(r, tail[t - 20][1]) and torch.equal(d, tail[t - 20][2])


def test_abi_rejects_bad_arguments():
    import ctypes as C
    from memory_gym_b200 import _lib as L
    lib = L.lib()
    spec = mg.spec("MortarMayhem-v0")
    p = spec.to_struct(spec.process_reset_params(None))
    io = L.StepIO()
    assert lib.mg_mm_step(C.byref(p), None, None, C.byref(io), 4, None) == -1
    assert b"NULL" in lib.mg_last_error()


@pytest.mark.parametrize("env_id", ["MortarMayhem-v0", "MortarMayhem-Grid-v0", "MysteryPath-v0", "SearingSpotlights-v0",
                                    "Endless-SearingSpotlights-v0", "Endless-MysteryPath-v0"])
def test_rollout_equals_step_loop(env_id):
    """mg_*_rollout (ONE launch, every warp runs all T steps of its envs) and mg_*_step_many (T launches
    enqueued by one call, programmatic dependent launch between them) give exactly what T separate
    step() calls give."""
    n, T = 500, 90
    opts = {"max_steps": 40} if env_id in ("MysteryPath-v0", "SearingSpotlights-v0", "Endless-SearingSpotlights-v0") else None
    a = mg.make(env_id, num_envs=n, device=DEV)
    b = mg.make(env_id, num_envs=n, device=DEV)
    acts = _random_actions(a, T, 77)
    a.reset(seed=5, options=opts)
    b.reset(seed=5, options=opts)
    ring = torch.zeros((3, n, 84, 84, 3), dtype=torch.uint8, device=DEV)
    rew = torch.zeros((T, n), dtype=torch.float64, device=DEV)
    dn = torch.zeros((T, n), dtype=torch.bool, device=DEV)
    inf = torch.zeros((T, n, mg._lib.MG_INFO_COLS), dtype=torch.float64, device=DEV)
    a.rollout(acts, ring, rew, dn, inf)
    c = mg.make(env_id, num_envs=n, device=DEV, warps_per_cta=3)
    c.reset(seed=5, options=opts)
    ring2, rew2, dn2, inf2 = torch.zeros_like(ring), torch.zeros_like(rew), torch.zeros_like(dn), torch.zeros_like(inf)
    c.step_many(acts, ring2, rew2, dn2, inf2)
    assert torch.equal(ring, ring2) and torch.equal(rew, rew2) and torch.equal(dn, dn2) and torch.equal(c.records, a.records)
    assert torch.equal(inf[dn], inf2[dn])
    for t in range(T):
        o, r, d, _, info = b.step(acts[t])
        assert torch.equal(r, rew[t]) and torch.equal(d, dn[t]), t
        assert torch.equal(info["episode"][d], inf[t][d]), t
        if t >= T - 3:
            assert torch.equal(o, ring[t % 3]), t
    assert dn.any() and torch.equal(a.records, b.records)


@pytest.mark.parametrize("env_id", ["MortarMayhem-v0", "Endless-MysteryPath-v0", "SearingSpotlights-v0"])
def test_rollout_scheduling_modes(env_id):
    """A rollout gives the same bytes whether the envs are handed to the warps through the global work counter
    (mg_rollout_io.sched) or through each CTA's shared-memory counter (sched = NULL), launch after launch
    (the global counter is re-armed by the kernel itself)."""
    n, T = 3000, 40
    a = mg.make(env_id, num_envs=n, device=DEV)
    b = mg.make(env_id, num_envs=n, device=DEV)
    b._sched = None  # CTA-local scheduling
    acts = _random_actions(a, T, 9)
    a.reset(seed=3)
    b.reset(seed=3)
    bufs = []
    for env in (a, b):
        ring = torch.zeros((T, n, 84, 84, 3), dtype=torch.uint8, device=DEV)
        rew = torch.zeros((T, n), dtype=torch.float64, device=DEV)
        dn = torch.zeros((T, n), dtype=torch.bool, device=DEV)
        for _ in range(3):
            env.rollout(acts, ring, rew, dn)
        bufs.append((ring, rew, dn))
    assert all(torch.equal(x, y) for x, y in zip(*bufs)) and torch.equal(a.records, b.records)
    assert bufs[0][2].any()
    if a._sched is not None:
        assert int(a._sched.abs().sum()) == 0  # re-armed


def test_results_buffer_aliases_reward_and_done():
    """`results` = reward (f64[N]) followed by done (u8[N]) in one allocation: one copy moves both to the host."""
    n = 257
    env = mg.make("MortarMayhem-Grid-v0", num_envs=n, device=DEV)
    env.reset(seed=1)
    acts = _random_actions(env, 60, 5)
    for t in range(60):
        _, r, d, _, _ = env.step(acts[t])
        raw = env.results.cpu().numpy()
        assert np.array_equal(raw[:8 * n].view(np.float64), r.cpu().numpy())
        assert np.array_equal(raw[8 * n:].astype(bool), d.cpu().numpy())
    assert d.dtype == torch.bool and r.dtype == torch.float64


@pytest.mark.parametrize("n", [1, 3, 300, 4096])
def test_policy_obs_matches_torch(n):
    """mg_obs_to_policy_bf16 == obs.permute(0, 3, 1, 2).float() / 255 -> bf16, bit for bit (fp tolerance: none)."""
    env = mg.make("SearingSpotlights-v0", num_envs=n, device=DEV)
    env.reset(seed=11)
    acts = _random_actions(env, 12, 2)
    for t in range(12):
        obs = env.step(acts[t])[0]
    want = (obs.permute(0, 3, 1, 2).to(torch.float32) / 255.0).to(torch.bfloat16).contiguous()
    got = env.policy_obs()
    assert got.shape == (n, 3, 84, 84) and got.dtype == torch.bfloat16
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    rnd = torch.randint(0, 256, (n, 84, 84, 3), dtype=torch.uint8, device=DEV)  # every byte value
    out = torch.empty((n, 3, 84, 84), dtype=torch.bfloat16, device=DEV)
    env.policy_obs(rnd, out)
    want = (rnd.permute(0, 3, 1, 2).to(torch.float32) / 255.0).to(torch.bfloat16).contiguous()
    assert torch.equal(out.view(torch.int16), want.view(torch.int16))


def test_emp_rollout_with_and_without_helper_warps(monkeypatch):
    """Endless-MysteryPath rollouts run helper warps that generate the next episode's path ahead of time; a reset
    adopts their result only when the env's stream still is what the helper started from. Same bytes either way."""
    n, T = 2500, 120
    outs = []
    for helpers in ("1", "0"):
        monkeypatch.setenv("MG_HELPERS", helpers)
        env = mg.make("Endless-MysteryPath-v0", num_envs=n, device=DEV)
        acts = _random_actions(env, T, 31)
        env.reset(seed=8)
        ring = torch.zeros((T, n, 84, 84, 3), dtype=torch.uint8, device=DEV)
        rew = torch.zeros((T, n), dtype=torch.float64, device=DEV)
        dn = torch.zeros((T, n), dtype=torch.bool, device=DEV)
        env.rollout(acts, ring, rew, dn)
        env.rollout(acts, ring, rew, dn)
        torch.cuda.synchronize()
        outs.append((ring, rew, dn, env.records.clone()))
    assert all(torch.equal(x, y) for x, y in zip(*outs))
    assert int(outs[0][2].sum()) > 3 * n  # several episodes per env: the hand-over was exercised


@pytest.mark.parametrize("options", [None, {"camera_offset_scale": 1.7, "show_background": True, "show_stamina": True,
                                             "agent_speed": 13.0, "max_steps": 30, "agent_scale": 0.5}])
def test_emp_rollout_launches_vs_oracle(options):
    """Endless-MysteryPath through ROLLOUT launches of a full-size batch (20 steps each; from the second episode on most
    resets adopt paths the helper warps generated ahead, one env per lane): a sample of the envs against independent
    oracle instances, every frame, reward and done."""
    from oracle import memgym_oracle as O
    n, m, K, launches = 4096, 16, 20, 24
    env = mg.make("Endless-MysteryPath-v0", num_envs=n, device=DEV)
    obs, _ = env.reset(seed=300, options=options)
    sample = [7 + 255 * i for i in range(m)]  # spread over the CTAs / helper warps
    oras = [O.make("Endless-MysteryPath-v0") for _ in sample]
    for e, o in zip(sample, oras):
        assert np.array_equal(obs[e].cpu().numpy(), o.reset(seed=300 + e, options=options)[0])
    rng = np.random.default_rng(12)
    ring = torch.zeros((K, n, 84, 84, 3), dtype=torch.uint8, device=DEV)
    rew = torch.zeros((K, n), dtype=torch.float64, device=DEV)
    dn = torch.zeros((K, n), dtype=torch.bool, device=DEV)
    idx = torch.as_tensor(sample, device=DEV)
    spec0 = env.state[idx, 2048:2060].clone().view(torch.int32)
    episodes = 0
    for launch in range(launches):
        a = rng.integers(0, 4, size=(K, n))
        env.rollout(torch.as_tensor(a, device=DEV), ring, rew, dn)
        got, r, d = ring[:, idx].cpu().numpy(), rew[:, idx].cpu().numpy(), dn[:, idx].cpu().numpy()
        for t in range(K):
            for j, (e, o) in enumerate(zip(sample, oras)):
                w_obs, w_r, w_d, _, _ = o.step(int(a[t][e]))
                assert float(w_r) == r[t][j] and bool(w_d) == bool(d[t][j]), (launch, t, e)
                if w_d:
                    episodes += 1
                    w_obs, _ = o.reset(options=options)
                assert np.array_equal(got[t][j], w_obs), (launch, t, e)
    adopted = int((env.state[idx, 2048:2060].clone().view(torch.int32)[:, 2] - spec0[:, 2]).sum())
    assert episodes > 5 * m and adopted > episodes // 3, (episodes, adopted)  # the adopted paths are the reference's paths


def test_emp_lookahead_survives_a_random_mix_of_operations():
    """scripts/emp_stress.py: rollouts of random length, single steps, step_many, masked re-seeding resets, full resets and
    restored snapshots in random order, on an env with helper warps and on one without: identical bytes after every
    operation (adopted, missing and stale look-ahead episodes all occur)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "emp_stress.py"), "1500", "90", "5"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().splitlines()[-1].startswith("ok: 90 operations"), out.stdout + out.stderr


def test_emp_results_do_not_depend_on_the_lookahead_block():
    """The second half of an Endless-MysteryPath record (the episodes generated ahead) is a cache: a reset adopts an
    episode only if its stream matches. Overwriting that half with random bytes — chains that claim to be active, parked
    searches, counters — must change nothing but the adoption statistics."""
    n, K = 3000, 40
    outs = []
    for scramble in (False, True):
        env = mg.make("Endless-MysteryPath-v0", num_envs=n, device=DEV)
        env.reset(seed=21)
        acts = _random_actions(env, K, 9)
        ring = torch.zeros((K, n, 84, 84, 3), dtype=torch.uint8, device=DEV)
        rew = torch.zeros((K, n), dtype=torch.float64, device=DEV)
        dn = torch.zeros((K, n), dtype=torch.bool, device=DEV)
        g = torch.Generator(device=DEV).manual_seed(77)
        for launch in range(4):
            if scramble and launch in (0, 2):
                env.state[:, 2048:] = torch.randint(0, 256, (n, 2048), generator=g, device=DEV, dtype=torch.uint8)
            env.rollout(acts, ring, rew, dn)
        for t in range(6):
            if scramble and t == 2:
                env.state[:, 2048:] = torch.randint(0, 256, (n, 2048), generator=g, device=DEV, dtype=torch.uint8)
            o, r, d, _, _ = env.step(acts[t])
        torch.cuda.synchronize()
        outs.append((ring.clone(), rew.clone(), dn.clone(), o.clone(), r.clone(), env.records.clone()))
    assert all(torch.equal(x, y) for x, y in zip(*outs))
    assert int(outs[0][2].sum()) > n


def test_emp_resets_adopt_paths_generated_ahead():
    """Endless-MysteryPath: the helper warps keep a chain of episodes ahead of every env (MgEPSpecRec behind the
    record: hword, tword, adopted). In steady state nearly every auto-reset adopts its path from there — in rollouts of
    any launch length and in single-step launches (where a search stopped by the end of a launch is parked and
    resumed by the next) — and the bytes equal the launches without helper warps (the test above, the step-loop tests).
    (One helper lane per env produces an episode in ~300 us: batches of fewer than ~8192 envs per GPU step an env more
    often than that and adopt less — 63 % at 4096 envs.)"""
    n = 8192
    env = mg.make("Endless-MysteryPath-v0", num_envs=n, device=DEV)
    env.reset(seed=5)

    def spec():
        return env.state[:, 2048:2060].clone().view(torch.int32)

    for K, launches in ((20, 24), (5, 60)):
        acts = _random_actions(env, K, 7 + K)
        ring = torch.zeros((min(K, 4), n, 84, 84, 3), dtype=torch.uint8, device=DEV)
        rew = torch.zeros((K, n), dtype=torch.float64, device=DEV)
        dn = torch.zeros((K, n), dtype=torch.bool, device=DEV)
        for _ in range(launches // 2):
            env.rollout(acts, ring, rew, dn)
        before, resets = spec(), 0
        for _ in range(launches):
            env.rollout(acts, ring, rew, dn)
            resets += int(dn.sum())
        adopted = int((spec()[:, 2] - before[:, 2]).sum())
        assert resets > n and adopted > (0.8 if K >= 20 else 0.65) * resets, (K, resets, adopted)
    acts = _random_actions(env, 120, 3)
    before, resets = spec(), 0
    for t in range(120):
        _, _, d, _, _ = env.step(acts[t])
        resets += int(d.sum())
    after = spec()
    adopted = int((after[:, 2] - before[:, 2]).sum())
    assert resets > n and adopted > 0.9 * resets, (resets, adopted)
    ahead = ((after[:, 0] >> 4) - after[:, 1]) & 255
    assert int(ahead.max()) <= 8 and float((ahead >= 5).float().mean()) > 0.7  # the chains stay (nearly) full


@pytest.mark.parametrize("env_id", ["SearingSpotlights-v0", "Endless-SearingSpotlights-v0", "MortarMayhem-v0",
                                    "Endless-MysteryPath-v0"])
def test_rollout_final_obs_equals_step_loop(env_id):
    """Terminal frames (final_obs) written inside a rollout launch — for Searing Spotlights from the deferred
    auto-reset of the pipelined loop — equal what the step loop leaves behind."""
    n, T = 600, 70
    opts = {"max_steps": 25} if "SearingSpotlights" in env_id else None
    a = mg.make(env_id, num_envs=n, device=DEV, final_obs=True)
    b = mg.make(env_id, num_envs=n, device=DEV, final_obs=True)
    acts = _random_actions(a, T, 123)
    a.reset(seed=21, options=opts)
    b.reset(seed=21, options=opts)
    ring = torch.zeros((2, n, 84, 84, 3), dtype=torch.uint8, device=DEV)
    rew = torch.zeros((T, n), dtype=torch.float64, device=DEV)
    dn = torch.zeros((T, n), dtype=torch.bool, device=DEV)
    a.rollout(acts, ring, rew, dn)
    ever = torch.zeros(n, dtype=torch.bool, device=DEV)
    for t in range(T):
        o, r, d, _, _ = b.step(acts[t])
        assert torch.equal(r, rew[t]) and torch.equal(d, dn[t]), t
        ever |= d
    assert ever.any()
    assert torch.equal(o, ring[(T - 1) % 2])
    assert torch.equal(a.final_obs[ever], b.final_obs[ever]) and torch.equal(a.records, b.records)


def test_step_graph_replay_equals_step():
    """capture_step(): one CUDA graph = H2D of the actions + the step kernel + D2H of reward/done; its replays give
    exactly what step() gives, and capturing does not advance the envs."""
    n = 512
    a = mg.make("MortarMayhem-v0", num_envs=n, device=DEV)
    b = mg.make("MortarMayhem-v0", num_envs=n, device=DEV)
    a.reset(seed=4)
    b.reset(seed=4)
    acts = _random_actions(a, 70, 6)
    graph = a.capture_step()
    assert torch.equal(a.records, b.records) and torch.equal(a.obs, b.obs)
    for t in range(70):
        graph.actions_host.copy_(acts[t].cpu())
        graph.replay()
        graph.done_event.synchronize()
        o, r, d, _, _ = b.step(acts[t])
        assert torch.equal(graph.obs, o), t
        assert torch.equal(graph.results_host[:8 * n].view(torch.float64), r.cpu()), t
        assert torch.equal(graph.results_host[8 * n:].bool(), d.cpu()), t
    assert torch.equal(a.records, b.records)


@pytest.mark.parametrize("env_id", ["MortarMayhem-v0", "SearingSpotlights-v0", "Endless-SearingSpotlights-v0",
                                    "Endless-MysteryPath-v0", "MysteryPath-v0"])
def test_rollout_edge_geometries(env_id):
    """Tiny batches and one- or two-step rollouts (more warps and helper warps than envs, a pipeline with nothing to
    overlap, ring of one slot) still equal the step loop."""
    for n, T, ring_n in ((1, 1, 1), (1, 2, 1), (2, 5, 2), (3, 33, 1), (149, 3, 3)):
        a = mg.make(env_id, num_envs=n, device=DEV)
        b = mg.make(env_id, num_envs=n, device=DEV)
        acts = _random_actions(a, T, 50 + n)
        a.reset(seed=n)
        b.reset(seed=n)
        ring = torch.zeros((ring_n, n, 84, 84, 3), dtype=torch.uint8, device=DEV)
        rew = torch.zeros((T, n), dtype=torch.float64, device=DEV)
        dn = torch.zeros((T, n), dtype=torch.bool, device=DEV)
        a.rollout(acts, ring, rew, dn)
        a.rollout(acts, ring, rew, dn)  # a second launch on the same records
        for rep in range(2):
            for t in range(T):
                o, r, d, _, _ = b.step(acts[t])
        assert torch.equal(r, rew[T - 1]) and torch.equal(d, dn[T - 1]), (n, T)
        assert torch.equal(o, ring[(T - 1) % ring_n]) and torch.equal(a.records, b.records), (n, T)


@pytest.mark.parametrize("env_id,opts", [
    ("SearingSpotlights-v0", {"max_steps": 50, "coins_visible": True, "exit_visible": True, "agent_visible": True,
                              "num_coins": [2, 4], "agent_health": 3, "spot_damage": 0.7, "initial_spawns": 6,
                              "spot_min_radius": 12.0, "spot_max_radius": 20.0, "spot_min_speed": 0.01, "spot_max_speed": 0.03}),
    ("SearingSpotlights-v0", {"max_steps": 45, "exit_visible": True, "hide_chessboard": True, "light_threshold": 100,
                              "light_dim_off_duration": 3, "show_last_action": False, "visual_feedback": False,
                              "initial_spawn_interval": 3, "spawn_interval_threshold": 2, "num_spawns": 9}),
    ("SearingSpotlights-v0", {"max_steps": 40, "black_background": True, "num_coins": [3]}),
    ("SearingSpotlights-v0", {"max_steps": 40, "sample_agent_position": False, "light_dim_off_duration": 0,
                              "show_last_positive_reward": False, "agent_visible": True, "agent_scale": 0.4, "agent_speed": 5.0}),
    ("Endless-SearingSpotlights-v0", {"max_steps": 55, "coins_visible": True, "agent_visible": True, "agent_health": 3,
                                      "spot_damage": 0.7, "initial_spawns": 6, "spawn_interval": 7, "coin_scale": 1.0}),
    ("Endless-SearingSpotlights-v0", {"steps_per_coin": 30, "hide_chessboard": True, "light_threshold": 100,
                                      "light_dim_off_duration": 3, "visual_feedback": False, "coin_show_duration": 20}),
    ("Endless-SearingSpotlights-v0", {"coin_enabled": False, "max_steps": 35, "agent_visible": True, "coins_visible": True}),
    ("Endless-SearingSpotlights-v0", {"black_background": True, "steps_per_coin": 25}),
])
def test_pipelined_rollout_with_options_equals_step_loop(env_id, opts):
    """The pipelined Searing-Spotlights rollout kernels under non-default layer orders, dimming ramps, outlines, opaque
    coin layers ...: every frame, reward and done of a rollout equals the step loop (which the oracle tests pin)."""
    n, T = 200, 90
    a = mg.make(env_id, num_envs=n, device=DEV, final_obs=True)
    b = mg.make(env_id, num_envs=n, device=DEV, final_obs=True)
    acts = _random_actions(a, T, 17)
    a.reset(seed=40, options=opts)
    b.reset(seed=40, options=opts)
    full = torch.zeros((T, n, 84, 84, 3), dtype=torch.uint8, device=DEV)
    rew = torch.zeros((T, n), dtype=torch.float64, device=DEV)
    dn = torch.zeros((T, n), dtype=torch.bool, device=DEV)
    a.rollout(acts, full, rew, dn)
    for t in range(T):
        o, r, d, _, _ = b.step(acts[t])
        assert torch.equal(o, full[t]), (t, int((o != full[t]).sum()))
        assert torch.equal(r, rew[t]) and torch.equal(d, dn[t]), t
    assert dn.any() and torch.equal(a.records, b.records) and torch.equal(a.final_obs, b.final_obs)


@pytest.mark.parametrize("env_id,opts", [
    ("Endless-MysteryPath-v0", {"camera_offset_scale": 1.7, "show_background": True, "show_stamina": True,
                                "agent_speed": 13.0, "max_steps": 30, "agent_scale": 0.5}),
    ("Endless-MysteryPath-v0", {"camera_offset_scale": 0.0, "stamina_level": 9, "reward_fall_off": -0.3,
                                "reward_path_progress_dense": 0.02, "reward_step": -0.001}),
    ("Endless-MysteryPath-v0", {"camera_offset_scale": 9.0, "agent_speed": 7.0, "stamina_level": 35,
                                "show_stamina": True, "visual_feedback": False}),
    ("MysteryPath-Grid-v0", {"max_steps": 30, "show_origin": True}),
    ("Endless-MortarMayhem-v0", {"explosion_delay": [4], "explosion_duration": [2], "agent_speed": 5.0}),
    ("MortarMayhemB-v0", {"command_count": [1, 3], "explosion_delay": [4], "explosion_duration": [2]}),
])
def test_rollout_with_options_equals_step_loop(env_id, opts):
    """Rollout launches (helper warps for Endless Mystery Path, global work counter, record on chip) under option
    sets equal the step loop frame by frame."""
    n, T = 200, 100
    a = mg.make(env_id, num_envs=n, device=DEV)
    b = mg.make(env_id, num_envs=n, device=DEV)
    acts = _random_actions(a, T, 19)
    a.reset(seed=60, options=opts)
    b.reset(seed=60, options=opts)
    full = torch.zeros((T, n, 84, 84, 3), dtype=torch.uint8, device=DEV)
    rew = torch.zeros((T, n), dtype=torch.float64, device=DEV)
    dn = torch.zeros((T, n), dtype=torch.bool, device=DEV)
    a.rollout(acts, full, rew, dn)
    for t in range(T):
        o, r, d, _, _ = b.step(acts[t])
        o = o["visual_observation"] if isinstance(o, dict) else o
        assert torch.equal(o, full[t]), (t, int((o != full[t]).sum()))
        assert torch.equal(r, rew[t]) and torch.equal(d, dn[t]), t
    assert dn.any() and torch.equal(a.records, b.records)
    if a.task_b:
        assert torch.equal(a.vector_obs, b.vector_obs)


@pytest.mark.parametrize("env_id,n,steps", [("MysteryPath-v0", 8192, 520), ("SearingSpotlights-v0", 8192, 264)])
def test_default_options_full_episode_at_baseline_size(env_id, n, steps):
    """BASELINE configs 3 and 4 as bench.py runs them — DEFAULT options (MysteryPath max_steps 512, SearingSpotlights
    max_steps 256), 8192 envs, one rollout launch longer than a whole episode: sampled envs equal the CPU oracle
    step for step (observations through a position-weighted checksum per frame, rewards, dones), every env finishes
    at least one episode, and a different launch geometry gives the same bytes."""
    from oracle import memgym_oracle as O
    env = mg.make(env_id, num_envs=n, device=DEV)
    acts = _random_actions(env, steps, 99)
    env.reset(seed=0)
    ring = torch.empty((4, n, 84, 84, 3), dtype=torch.uint8, device=DEV)  # a 4-slot ring: the last frames survive
    rew = torch.empty((steps, n), dtype=torch.float64, device=DEV)
    dn = torch.empty((steps, n), dtype=torch.bool, device=DEV)
    env.rollout(acts, ring, rew, dn)
    torch.cuda.synchronize()
    assert bool(dn.any(0).all())  # max_steps ends every episode inside the launch
    sample = [0, 1, 2, n // 2, n - 1]
    # the same envs, stepped one launch per step in a small shard: every frame is checked against the oracle
    a_cpu = acts.cpu().numpy()
    for e in sample:
        sh = mg.make(env_id, num_envs=1, device=DEV, env_offset=e)
        obs = sh.reset(seed=0)[0]
        ora = O.make(env_id)
        w_obs, _ = ora.reset(seed=e)
        assert np.array_equal(obs[0].cpu().numpy(), w_obs)
        for t in range(steps):
            obs, r, d, _, _ = sh.step(acts[t, e:e + 1].contiguous())
            w_obs, w_r, w_d, _, _ = ora.step(a_cpu[t, e])
            if w_d:
                w_obs, _ = ora.reset()
            assert float(w_r) == float(r[0]) == float(rew[t, e]) and bool(w_d) == bool(d[0]) == bool(dn[t, e]), (e, t)
            assert np.array_equal(obs[0].cpu().numpy(), w_obs), (e, t)
        assert torch.equal(obs[0], ring[(steps - 1) % 4, e]), e  # ... and the big launch ended on the same frame
    env2 = mg.make(env_id, num_envs=n, device=DEV, warps_per_cta=5)
    env2.reset(seed=0)
    ring2, rew2, dn2 = torch.empty_like(ring), torch.empty_like(rew), torch.empty_like(dn)
    env2.rollout(acts, ring2, rew2, dn2)
    assert torch.equal(ring, ring2) and torch.equal(rew, rew2) and torch.equal(dn, dn2)
    assert torch.equal(env.records, env2.records)


def test_zero_copy_host_actions_and_results():
    """A host-side policy loop without copy operations: actions read by the kernel straight from PINNED host memory,
    rewards / dones written straight into pinned host memory (host_results=True). Same bytes as the device-buffer path."""
    n, T = 700, 80
    a = mg.make("MortarMayhem-v0", num_envs=n, device=DEV)
    b = mg.make("MortarMayhem-v0", num_envs=n, device=DEV, host_results=True)
    assert b.reward.device.type == "cpu" and b.results.is_pinned()
    acts = _random_actions(a, T, 5)
    acts_host = acts.cpu().pin_memory()
    assert torch.equal(a.reset(seed=3)[0], b.reset(seed=3)[0])
    for t in range(T):
        oa, ra, da, _, ia = a.step(acts[t])
        ob, rb, db, _, ib = b.step(acts_host[t])
        torch.cuda.synchronize()
        assert torch.equal(oa, ob) and torch.equal(ra.cpu(), rb) and torch.equal(da.cpu(), db)
    assert torch.equal(a.records, b.records) and torch.equal(a.info, b.info)
    # results_out: per-step result buffers (pinned host memory, double-buffered) on an ordinary env
    c = mg.make("MortarMayhem-v0", num_envs=n, device=DEV)
    c.reset(seed=3)
    a.reset(seed=3)
    res = [torch.zeros(9 * n, dtype=torch.uint8).pin_memory() for _ in range(2)]
    for t in range(T):
        oa, ra, da, _, _ = a.step(acts[t])
        oc, rc, dc, _, ic = c.step(acts_host[t], results_out=res[t & 1])
        torch.cuda.synchronize()
        assert rc.data_ptr() == res[t & 1].data_ptr() and ic["_episode"] is dc
        assert torch.equal(oa, oc) and torch.equal(ra.cpu(), rc) and torch.equal(da.cpu(), dc)


def test_launch_follows_the_env_device_not_the_current_device():
    """the library launches on the device that owns the env records. With a second GPU the env lives on
    cuda:1 while the current device stays cuda:0; with one GPU the guard's lookup path still runs on every call."""
    dev = "cuda:1" if torch.cuda.device_count() > 1 else DEV
    torch.cuda.set_device(0)
    env = mg.make("MortarMayhem-v0", num_envs=300, device=dev)
    ref = mg.make("MortarMayhem-v0", num_envs=300, device=DEV)
    acts = _random_actions(ref, 30, 8)
    assert torch.equal(env.reset(seed=1)[0].cpu(), ref.reset(seed=1)[0].cpu())
    for t in range(30):
        o1, r1, d1, _, _ = env.step(acts[t].to(dev))
        o2, r2, d2, _, _ = ref.step(acts[t])
        assert torch.equal(o1.cpu(), o2.cpu()) and torch.equal(r1.cpu(), r2.cpu()) and torch.equal(d1.cpu(), d2.cpu())
    assert torch.cuda.current_device() == 0
    out = env.policy_obs()
    assert out.device == env.obs.device


def test_obs_delta_encoder_matches_stage_format():
    """mg_obs_delta_encode against the numpy statement of the stage layout (tests/test_codec_host.py), and the host
    decoder on the device-written stage: random frames, sparse changes, a forced full resync."""
    import ctypes as C
    from test_codec_host import decode, encode_numpy
    from memory_gym_b200 import _lib as L
    lib = L.lib()
    n = 300
    g = torch.Generator(device=DEV).manual_seed(2)
    prev = torch.randint(0, 256, (n, 21168), dtype=torch.uint8, generator=g, device=DEV)
    host = prev.cpu().numpy().copy()
    stage = torch.zeros(int(lib.mg_obs_delta_stage_bytes(n)), dtype=torch.uint8).pin_memory()
    for step in range(4):
        cur = prev.clone()
        if step != 1:  # step 1: nothing changed at all
            m = torch.rand((n, 21168), generator=g, device=DEV) < (0.002 if step else 0.3)
            cur[m] ^= 0x5A
            cur[0, -1] ^= 1
        prev_np = prev.cpu().numpy().copy()
        full = step == 3
        L.check(lib.mg_obs_delta_encode(cur.data_ptr(), prev.data_ptr(), n, stage.data_ptr(), int(full),
                                        torch.cuda.current_stream().cuda_stream), "encode")
        torch.cuda.synchronize()
        assert torch.equal(prev, cur)  # the device copy of "what the host holds" is up to date
        want = encode_numpy(cur.cpu().numpy(), prev_np, full)
        got = stage.numpy()
        nh = n * L.MG_DELTA_HDR_BYTES
        hdr_w = want[:nh].view(np.uint32).reshape(n, -1)
        hdr_g = got[:nh].view(np.uint32).reshape(n, -1)
        assert np.array_equal(hdr_g[:, :43], hdr_w[:, :43])
        for e in (0, 1, n // 2, n - 1):
            k = 16 * int(hdr_w[e, 0])
            lo = nh + e * L.MG_OBS_BYTES
            assert np.array_equal(got[lo:lo + k], want[lo:lo + k]), (step, e)
        rc, nbytes = decode(got, host, 8)
        assert rc == 0 and np.array_equal(host, cur.cpu().numpy())
        assert nbytes == 16 * int(hdr_w[:, 0].sum()) + nh


@pytest.mark.parametrize("mode", ["direct", "patch"])
@pytest.mark.parametrize("env_id,depth", [("MortarMayhem-v0", 1), ("MortarMayhem-v0", 2), ("SearingSpotlights-v0", 2),
                                          ("Endless-MysteryPath-v0", 3), ("MortarMayhemB-Grid-v0", 2)])
def test_host_pipe_host_arrays_equal_device(env_id, depth, mode):
    """HostPipe: observations / rewards / dones in HOST memory after every step (delta transport over PCIe, zero-copy
    actions and results) equal the device-side step loop bit for bit — with `depth` steps in flight; changed chunks
    written by the GPU straight to their place in pinned host arrays ("direct") or patched in by host threads ("patch")."""
    n, T = 500, 90
    a = mg.make(env_id, num_envs=n, device=DEV)
    b = mg.make(env_id, num_envs=n, device=DEV)
    acts = _random_actions(a, T, 17)
    acts_host = acts.cpu()
    obs_a = a.reset(seed=6)[0]
    b.reset(seed=6)
    pipe = b.host_pipe(threads=6, depth=depth, mode=mode)
    vis = (lambda o: o["visual_observation"]) if a.task_b else (lambda o: o)
    assert torch.equal(pipe.obs, vis(obs_a).cpu())
    want = []
    for t in range(T):
        o, r, d, _, _ = a.step(acts[t])
        want.append((vis(o).cpu(), r.cpu(), d.cpu()))
    sent0 = pipe.total_d2h_bytes
    for t in range(T + depth - 1):
        if t < T:
            pipe.submit(acts_host[t] if t % 2 else acts_host[t].pin_memory())  # pinned and pageable action arrays
        if t >= depth - 1:
            o, r, d, _, _ = pipe.collect()
            wo, wr, wd = want[t - depth + 1]
            assert torch.equal(o, wo) and torch.equal(r, wr) and torch.equal(d, wd), t
    assert 0 < pipe.total_d2h_bytes - sent0 < T * n * 21168  # less than whole frames would have cost
    # resets through the pipe, then more steps
    o2, _ = pipe.reset(seed=9)
    assert torch.equal(o2, vis(a.reset(seed=9)[0]).cpu())
    o, r, d, _, _ = pipe.step(acts_host[0])
    wo, wr, wd, _, _ = a.step(acts[0])
    assert torch.equal(o, vis(wo).cpu()) and torch.equal(r, wr.cpu()) and torch.equal(d, wd.cpu())
