"""The restated CPU oracle replays, bit for bit, trajectories recorded from the UNMODIFIED reference
(tests/golden/ref_*.npz, made by oracle/gen_golden.py from /root/reference on oracle/shims)."""
import numpy as np
import pytest

from oracle import memgym_oracle as O
from golden_util import action_for, crc, load_case, visual

CASES = ["mm", "mm_opts", "mmgrid", "emm", "emm_opts", "mp", "mp_opts", "ss", "ss_opts", "ss_black",
         "mmb", "mmb_opts", "mmbgrid", "mpgrid", "mpgrid_opts", "ess", "ess_opts", "ess_black", "ess_nocoin", "ss_dim0", "ess_dim0",
         "emp", "emp_opts", "emp_nopast", "emp_cam"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_replays_reference(name):
    case = load_case(name)
    if case["env_id"] not in O.ORACLES:
        pytest.skip("oracle for this env not built yet")
    n_envs, steps = case["done"].shape
    infos = {(r["env"], r["t"]): r for r in case["infos"]}
    for e in range(n_envs):
        env = O.make(case["env_id"])
        obs, info = env.reset(seed=case["seed0"] + e, options=case["options"])
        assert np.array_equal(visual(obs), case["frames"][case["frame_map"][(e, -1)]])
        if "gt" in case:
            assert np.array_equal(np.asarray(info["ground_truth"], dtype=np.float64), case["gt_reset"][e])
        for t in range(steps):
            obs, reward, done, trunc, info = env.step(action_for(case, e, t))
            assert trunc is False
            assert float(reward) == case["reward"][e, t], (e, t)
            assert bool(done) == bool(case["done"][e, t]), (e, t)
            if (e, t) in case["frame_map"]:
                want = case["frames"][case["frame_map"][(e, t)]]
                assert np.array_equal(visual(obs), want), (e, t, int((visual(obs) != want).sum()))
            assert crc(obs) == int(case["crc"][e, t]), (e, t)
            if "gt" in case:
                assert np.array_equal(np.asarray(info["ground_truth"], dtype=np.float64), case["gt"][e, t]), (e, t)
            if done:
                ref = infos[(e, t)]
                assert float(info["reward"]) == ref["reward_seq_sum"]
                for k, v in ref.items():
                    if k in ("env", "t", "reward_seq_sum"):
                        continue
                    assert np.array_equal(np.asarray(info[k], dtype=np.float64), np.asarray(v)), (k, e, t)
                obs, _ = env.reset(options=case["options"])
                assert crc(obs) == int(case["reset_crc"][e, t]), (e, t)
