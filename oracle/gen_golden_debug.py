"""Golden `debug_rgb_array` renders from the UNMODIFIED reference (build container only).

    python oracle/gen_golden_debug.py        # writes tests/golden/ref_dbg_<case>.npz

Same harness as gen_golden.py (reference sources from /root/reference on top of oracle/shims), with the envs made in
`render_mode="debug_rgb_array"` and `env.render()` (`_build_debug_surface`, e.g. memory_gym/mortar_mayhem.py:105-136,
397-399) called exactly ONCE after every reset() and every step() — the call pattern the batched engine reproduces
(the reference's render pops a clone of the command queue per call). Stored per step: action, reward, done, CRC32 of the
observation and of the 336x336x3 debug frame, plus a sparse set of debug frames at 84x84 (the 4x nearest-neighbour
scale is checked to be exact before sub-sampling).
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    "mm": ("MortarMayhem-v0", None, 4, 330),
    "mm_opts": ("MortarMayhem-v0", {"arena_size": 4, "allowed_commands": 6, "command_count": [1, 3],
                                     "command_show_duration": [1, 2], "command_show_delay": [0, 1],
                                     "explosion_duration": [2, 4], "explosion_delay": [5, 9],
                                     "visual_feedback": False, "agent_speed": 4.0}, 3, 220),
    "mmgrid": ("MortarMayhem-Grid-v0", None, 3, 260),
    "emm": ("Endless-MortarMayhem-v0", {"explosion_delay": [8], "explosion_duration": [3]}, 4, 420),
    "mp": ("MysteryPath-v0", None, 4, 560),
    "mp_opts": ("MysteryPath-v0", {"max_steps": 90, "show_origin": True, "cardinal_origin_choice": [1, 3],
                                    "visual_feedback": False, "agent_speed": 5.0}, 3, 260),
    "mpgrid": ("MysteryPath-Grid-v0", None, 3, 300),
    "ss": ("SearingSpotlights-v0", None, 4, 420),
    "ss_opts": ("SearingSpotlights-v0", {"max_steps": 120, "initial_spawns": 3, "num_coins": [2, 3], "agent_health": 4,
                                          "light_dim_off_duration": 10, "coins_visible": True, "agent_visible": True,
                                          "hide_chessboard": True, "agent_speed": 4.0}, 3, 300),
    "ss_black": ("SearingSpotlights-v0", {"max_steps": 100, "black_background": True, "exit_visible": True,
                                           "light_threshold": 140, "light_dim_off_duration": 4, "initial_spawns": 6,
                                           "spot_min_speed": 0.008, "spot_max_speed": 0.02, "show_last_action": False,
                                           "num_coins": [2]}, 3, 240),
    "ess": ("Endless-SearingSpotlights-v0", {"steps_per_coin": 80}, 4, 420),
    "emp": ("Endless-MysteryPath-v0", None, 4, 420),
    "emp_opts": ("Endless-MysteryPath-v0", {"max_steps": 150, "show_background": True, "show_stamina": True,
                                             "camera_offset_scale": 2.3, "stamina_level": 12, "agent_speed": 5.0,
                                             "visual_feedback": False}, 3, 300),
    "ess_nocoin": ("Endless-SearingSpotlights-v0", {"coin_enabled": False, "max_steps": 60, "initial_spawns": 6,
                                                     "agent_health": 3, "black_background": True}, 3, 160),
}


def main(argv):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(HERE, "shims"))
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, HERE)
    import gymnasium
    import memory_gym  # noqa: F401
    from gen_golden import goal_seeking_actions

    only = set(argv[1:])
    for name, (env_id, options, n_envs, steps) in CASES.items():
        if only and name not in only:
            continue
        rng = np.random.default_rng(zlib.crc32(("dbg_" + name).encode()))
        rec = {k: [] for k in ("action", "reward", "done", "crc", "dbg_crc", "reset_dbg_crc")}
        frames, frame_idx, first_dbg = [], [], []

        def grab(env, e, t):
            fr = env.render()
            assert fr.shape == (336, 336, 3) and fr.dtype == np.uint8
            sub = fr[::4, ::4]
            assert np.array_equal(np.repeat(np.repeat(sub, 4, 0), 4, 1), fr)
            keep = t < 0 or t % 29 == 0
            if keep:
                frames.append(sub.copy())
                frame_idx.append((e, t))
            return zlib.crc32(np.ascontiguousarray(fr).tobytes())

        for e in range(n_envs):
            env = gymnasium.make(env_id, render_mode="debug_rgb_array")
            env.reset(seed=300 + e, options=options)
            first_dbg.append(grab(env, e, -1))
            cols = {k: [] for k in rec}
            for t in range(steps):
                a = goal_seeking_actions(env_id, env, rng, e)
                obs, reward, done, trunc, info = env.step(a)
                cols["action"].append(np.asarray(a, dtype=np.int64).reshape(-1))
                cols["reward"].append(float(reward))
                cols["done"].append(bool(done))
                cols["crc"].append(zlib.crc32(np.ascontiguousarray(obs).tobytes()))
                cols["dbg_crc"].append(grab(env, e, t))
                rc = 0
                if done:
                    env.reset(options=options)
                    rc = grab(env, e, -2 - t)
                cols["reset_dbg_crc"].append(rc)
            for k in rec:
                rec[k].append(np.asarray(cols[k]))
        out = {k: np.stack(v) for k, v in rec.items()}
        for k in ("crc", "dbg_crc", "reset_dbg_crc"):
            out[k] = out[k].astype(np.uint32)
        out["first_dbg_crc"] = np.asarray(first_dbg, dtype=np.uint32)
        out["frames"] = np.stack(frames)
        out["frame_idx"] = np.asarray(frame_idx, dtype=np.int64)
        out["seed0"] = np.asarray(300)
        out["options_json"] = np.asarray(json.dumps(options))
        out["env_id"] = np.asarray(env_id)
        path = os.path.join(OUT, f"ref_dbg_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, env_id, "episodes:", int(out["done"].sum()), "frames:", len(frames), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(sys.argv)
