"""Golden trajectories from the UNMODIFIED reference (build container only).

    python oracle/gen_golden.py            # writes tests/golden/ref_<env>.npz

Imports ``memory_gym`` from /root/reference on top of ``oracle/shims`` (stand-ins for the missing
pygame / gymnasium wheels, see oracle/shims/README.md), steps every in-scope env id with seeded
random actions and auto-reset, and stores per step: action, reward, done, a CRC32 of the
observation, plus a sparse set of full frames. /root/reference does not exist on the GPU box, so
tests only ever read the committed .npz files.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (env id, options, n_envs (seeds 0..n-1... offset), steps, action kind)
    "mm": ("MortarMayhem-v0", None, 6, 400),
    "mm_opts": ("MortarMayhem-v0", {"arena_size": 4, "allowed_commands": 6, "command_count": [3, 5, 7],
                                     "command_show_duration": [1, 2], "command_show_delay": [0, 2],
                                     "explosion_duration": [2, 4], "explosion_delay": [5, 9],
                                     "visual_feedback": False, "reward_command_failure": -0.5,
                                     "reward_command_success": 0.3, "reward_episode_success": 1.5,
                                     "agent_speed": 4.0}, 4, 300),
    "mmgrid": ("MortarMayhem-Grid-v0", None, 6, 400),
    "emm": ("Endless-MortarMayhem-v0", None, 6, 500),
    "emm_opts": ("Endless-MortarMayhem-v0", {"max_steps": 90, "initial_command_count": 3, "allowed_commands": 5,
                                              "command_show_duration": [2], "command_show_delay": [1, 3],
                                              "explosion_duration": [3], "explosion_delay": [7, 11],
                                              "reward_new_command_success": 0.7, "reward_command_failure": -0.1},
                 4, 400),
    "mp": ("MysteryPath-v0", None, 6, 700),
    "mp_opts": ("MysteryPath-v0", {"max_steps": 120, "show_origin": True, "show_goal": True,
                                    "cardinal_origin_choice": [1, 3], "reward_fall_off": -0.2,
                                    "reward_step": -0.01, "agent_speed": 5.0}, 4, 400),
    "ss": ("SearingSpotlights-v0", None, 6, 600),
    "ss_opts": ("SearingSpotlights-v0", {"max_steps": 200, "initial_spawns": 2, "num_spawns": 8,
                                          "initial_spawn_interval": 12, "spawn_interval_threshold": 3,
                                          "num_coins": [1, 2, 3], "agent_health": 8, "spot_damage": 0.5,
                                          "light_dim_off_duration": 10, "coins_visible": True,
                                          "reward_inside_spotlight": -0.01, "reward_coin": 0.5,
                                          "agent_speed": 4.0}, 4, 450),
    "mmb": ("MortarMayhemB-v0", None, 6, 300),
    "mmb_opts": ("MortarMayhemB-v0", {"arena_size": 3, "allowed_commands": 7, "command_count": [2, 4, 20],
                                       "explosion_duration": [2, 3], "explosion_delay": [4, 6], "visual_feedback": False,
                                       "reward_command_failure": -0.3, "reward_episode_success": 2.0,
                                       "agent_speed": 4.0}, 4, 260),
    "mmbgrid": ("MortarMayhemB-Grid-v0", None, 6, 300),
    "mpgrid": ("MysteryPath-Grid-v0", None, 6, 400),
    "mpgrid_opts": ("MysteryPath-Grid-v0", {"max_steps": 60, "show_origin": True, "show_goal": True,
                                             "cardinal_origin_choice": [0, 2], "reward_fall_off": -0.1,
                                             "reward_path_progress": 0.2, "reward_step": -0.01,
                                             "visual_feedback": False}, 4, 300),
    "ss_black": ("SearingSpotlights-v0", {"max_steps": 150, "black_background": True, "agent_visible": True,
                                           "exit_visible": True, "light_threshold": 140, "light_dim_off_duration": 4,
                                           "initial_spawns": 6, "spot_min_speed": 0.008, "spot_max_speed": 0.02,
                                           "show_last_action": False, "num_coins": [2]}, 4, 320),
    "ess": ("Endless-SearingSpotlights-v0", None, 6, 600),
    "ess_opts": ("Endless-SearingSpotlights-v0", {"max_steps": 300, "steps_per_coin": 70, "initial_spawns": 2,
                                                   "spawn_interval": 15, "agent_health": 6, "spot_damage": 0.5,
                                                   "light_dim_off_duration": 10, "coin_show_duration": 12,
                                                   "reward_inside_spotlight": -0.01, "reward_coin": 0.5,
                                                   "agent_speed": 4.0, "agent_visible": True}, 4, 450),
    "ess_black": ("Endless-SearingSpotlights-v0", {"black_background": True, "coins_visible": True,
                                                    "light_threshold": 140, "light_dim_off_duration": 4,
                                                    "initial_spawns": 5, "spot_min_speed": 0.008,
                                                    "spot_max_speed": 0.02, "steps_per_coin": 90,
                                                    "sample_agent_position": False, "show_last_positive_reward": False},
                  4, 320),
    "ess_nocoin": ("Endless-SearingSpotlights-v0", {"coin_enabled": False, "max_steps": 70, "initial_spawns": 6,
                                                     "spot_min_speed": 0.01, "spot_max_speed": 0.03,
                                                     "agent_health": 3}, 4, 200),
    # light_dim_off_duration == 0: the reset frame shows spotlight_surface as the previous episode's last step left it
    "ss_dim0": ("SearingSpotlights-v0", {"light_dim_off_duration": 0, "light_threshold": 200, "max_steps": 45,
                                          "agent_health": 3, "initial_spawns": 5}, 4, 260),
    "ess_dim0": ("Endless-SearingSpotlights-v0", {"light_dim_off_duration": 0, "max_steps": 40, "black_background": True,
                                                   "agent_visible": True, "steps_per_coin": 30}, 4, 220),
    "emp": ("Endless-MysteryPath-v0", None, 6, 700),
    "emp_opts": ("Endless-MysteryPath-v0", {"max_steps": 150, "show_background": True, "show_stamina": True,
                                             "camera_offset_scale": 2.3, "stamina_level": 12, "reward_fall_off": -0.1,
                                             "reward_path_progress_dense": 0.05, "reward_step": -0.01,
                                             "agent_speed": 5.0, "show_origin": True}, 4, 450),
    "emp_nopast": ("Endless-MysteryPath-v0", {"show_past_path": False, "visual_feedback": False,
                                               "camera_offset_scale": 0.0, "agent_scale": 0.4, "agent_speed": 2.0,
                                               "stamina_level": 40}, 4, 400),
    "emp_cam": ("Endless-MysteryPath-v0", {"camera_offset_scale": 7.0, "show_background": True, "agent_speed": 4.0,
                                            "stamina_level": 30}, 4, 500),
}


def _toward(c, goal, dead=1):
    """MultiDiscrete component steering c toward goal: 1 = negative direction, 2 = positive."""
    if c > goal + dead:
        return 1
    if c < goal - dead:
        return 2
    return 0


def goal_seeking_actions(env_id, env, rng, e):
    """Seeded random actions; env instances with e % 3 != 0 follow a scripted expert (reading the
    reference env's internals) with 12 % random actions, so reward / success / growth branches are
    covered as well as the early-failure ones."""
    if env_id == "Endless-MysteryPath-v0":
        # Discrete(4): 0 no-op, 1 right, 2 up, 3 down — follow current_node.next_node where that is possible
        rand = int(rng.integers(0, 4))
        if e % 3 == 0 or rng.random() < (0.03 if e % 3 == 1 else 0.08):
            return rand
        cy = env.agent.rect.center[1]
        cur, nxt = env.current_node, env.current_node.next_node
        own_y = 12 * cur.y + 6
        if nxt.x > cur.x:
            if abs(cy - own_y) > 3:  # stay clear of the row boundary while moving right
                return 2 if cy > own_y else 3
            return 1
        if nxt.y < cur.y:
            return 2
        if nxt.y > cur.y:
            return 3
        return rand  # a westward link cannot be followed (no "left" action)
    grid = "Grid" in env_id
    rand = int(rng.integers(0, 4)) if grid else rng.integers(0, 3, 2)
    if e % 3 == 0 or rng.random() < 0.12:
        return rand
    cx, cy = env.agent.rect.center
    if "MortarMayhem" in env_id:
        ox = env.arena.rect[0]
        tx, ty = int(env._target_pos[0]), int(env._target_pos[1])
        if grid:
            gx, gy = (int(v) for v in env.agent.grid_position)
            if (gx, gy) == (tx, ty):
                return 0
            want = 3 if tx > gx else 1 if tx < gx else 0 if ty < gy else 2  # E / W / N / S
            face = env.agent.rotation // 90
            if face == want:
                return 3
            return 1 if (want - face) % 4 in (1, 2) else 2
        gx_c, gy_c = ox + 14 * tx + 7, ox + 14 * ty + 7
        if "Endless" in env_id:  # torus: take the short way round
            dx = (gx_c - cx + 42) % 84 - 42
            dy = (gy_c - cy + 42) % 84 - 42
            return np.asarray([_toward(0, dx, 2), _toward(0, dy, 2)])
        return np.asarray([_toward(cx, gx_c, 2), _toward(cy, gy_c, 2)])
    if env_id == "MysteryPath-Grid-v0":
        path = [(n.x, n.y) for n in env.mystery_path.path]
        here = (int(env.agent.grid_position[0]), int(env.agent.grid_position[1]))
        k = path.index(here) if here in path else len(path) - 1
        nx, ny = path[max(k - 1, 0)]
        want = 3 if nx > here[0] else 1 if nx < here[0] else 0 if ny < here[1] else 2  # E / W / N / S
        face = env.agent.rotation // 90
        if (nx, ny) == here:
            return 0
        if face == want:
            return 3
        return 1 if (want - face) % 4 in (1, 2) else 2
    if env_id == "MysteryPath-v0":
        path = [(n.x, n.y) for n in env.mystery_path.path]
        here = (cx // 12, cy // 12)
        k = path.index(here) if here in path else len(path) - 1
        nx, ny = path[max(k - 1, 0)]
        # finish centring on one axis before moving along the other (no corner cutting)
        own_x, own_y = 12 * here[0] + 6, 12 * here[1] + 6
        if nx != here[0]:
            if abs(cy - own_y) > 2:
                return np.asarray([0, _toward(cy, own_y, 1)])
            return np.asarray([_toward(cx, 12 * nx + 6, 1), 0])
        if abs(cx - own_x) > 2:
            return np.asarray([_toward(cx, own_x, 1), 0])
        return np.asarray([0, _toward(cy, 12 * ny + 6, 1)])
    if env_id == "Endless-SearingSpotlights-v0":
        if env.coin is None:
            return rand
        goal = env.coin.location
        return np.asarray([_toward(cx, int(goal[0]), 1), _toward(cy, int(goal[1]), 1)])
    if env_id == "SearingSpotlights-v0":
        goal = env.coins[0].location if env.coins else env.exit.location
        return np.asarray([_toward(cx, int(goal[0]), 1), _toward(cy, int(goal[1]), 1)])
    return rand


def _split(obs):
    """(visual frame, bytes of everything that must match) for plain and Dict observations."""
    if isinstance(obs, dict):
        vis, vec = obs["visual_observation"], obs["vector_observation"]
        return vis, vis.tobytes() + np.asarray(vec, dtype=np.float32).tobytes(), np.asarray(vec, dtype=np.float32)
    return obs, obs.tobytes(), None


def run_case(name, make_env, env_id, options, n_envs, steps, seed0=100, record_gt=False):
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    rec = {k: [] for k in ("action", "reward", "done", "crc", "reset_crc")}
    frames, frame_idx = [], []
    vecs, vec_idx = [], []
    infos = []
    gts, gt0s = [], []  # per-step info["ground_truth"] (envs that publish it every step) / the reset's
    for e in range(n_envs):
        env = make_env(env_id)
        obs, info = env.reset(seed=seed0 + e, options=options)
        gt_l = []
        if "ground_truth" in info:
            gt0s.append(np.asarray(info["ground_truth"], dtype=np.float64))
        vis, blob, vec = _split(obs)
        frames.append(vis.copy())
        frame_idx.append((e, -1))
        if vec is not None:
            vecs.append(vec.copy())
            vec_idx.append((e, -1))
        a_l, r_l, d_l, c_l, rc_l = [], [], [], [], []
        ep_rewards = []
        for t in range(steps):
            a = goal_seeking_actions(env_id, env, rng, e)
            obs, reward, done, trunc, info = env.step(a)
            assert trunc is False
            ep_rewards.append(reward)
            if "ground_truth" in info:
                gt_l.append(np.asarray(info["ground_truth"], dtype=np.float64))
            a_l.append(np.asarray(a, dtype=np.int64).reshape(-1))
            r_l.append(float(reward))
            d_l.append(bool(done))
            vis, blob, vec = _split(obs)
            c_l.append(zlib.crc32(blob))
            if t % 37 == 0 or done:
                frames.append(vis.copy())
                frame_idx.append((e, t))
            rc = 0
            if done:
                seq = 0
                for v in ep_rewards:
                    seq = seq + v
                row = {"env": e, "t": t, "reward_seq_sum": float(seq)}
                for k, v in info.items():
                    if k != "reward":
                        row[k] = np.asarray(v, dtype=np.float64).tolist()
                infos.append(row)
                ep_rewards = []
                obs, info = env.reset(options=options)  # seed=None: the stream continues
                vis, blob, vec = _split(obs)
                rc = zlib.crc32(blob)
                frames.append(vis.copy())
                frame_idx.append((e, -2 - t))
                if vec is not None:
                    vecs.append(vec.copy())
                    vec_idx.append((e, -2 - t))
            rc_l.append(rc)
        for k, v in zip(("action", "reward", "done", "crc", "reset_crc"), (a_l, r_l, d_l, c_l, rc_l)):
            rec[k].append(np.asarray(v))
        if gt_l:
            gts.append(np.stack(gt_l))
    out = {k: np.stack(v) for k, v in rec.items()}
    out["crc"] = out["crc"].astype(np.uint32)
    out["reset_crc"] = out["reset_crc"].astype(np.uint32)
    out["frames"] = np.stack(frames)
    out["frame_idx"] = np.asarray(frame_idx, dtype=np.int64)
    if vecs:
        out["vecs"] = np.stack(vecs)
        out["vec_idx"] = np.asarray(vec_idx, dtype=np.int64)
    if record_gt and gts:
        out["gt"] = np.stack(gts)      # [env][t][k]
        out["gt_reset"] = np.stack(gt0s)
    out["seed0"] = np.asarray(seed0)
    import json
    out["infos_json"] = np.asarray(json.dumps(infos))
    out["options_json"] = np.asarray(json.dumps(options))
    out["env_id"] = np.asarray(env_id)
    return out


def main(argv):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(HERE, "shims"))
    sys.path.insert(0, "/root/reference")
    import gymnasium
    import memory_gym  # noqa: F401  (registers the ids)

    only = set(argv[1:])
    for name, (env_id, options, n_envs, steps) in CASES.items():
        if only and name not in only:
            continue
        out = run_case(name, gymnasium.make, env_id, options, n_envs, steps, record_gt=name.startswith(("ess", "emp")))
        path = os.path.join(OUT, f"ref_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, env_id, "episodes:", int(out["done"].sum()), "frames:", len(out["frames"]),
              os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(sys.argv)
