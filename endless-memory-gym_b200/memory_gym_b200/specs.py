"""Env registry and reset-option handling — the same ids, option keys, defaults and assertions as the
reference, so an `options` dict written for the reference works unchanged.

Reference: ``memory_gym/__init__.py:13-60`` (ids), ``default_reset_parameters`` /
``process_reset_params`` of each env class (``mortar_mayhem.py:21-54``, ``mortar_mayhem_grid.py:21-55``,
``endless_mortar_mayhem.py:21-53``, ``mystery_path.py:21-52``, ``searing_spotlights.py:22-80``,
``endless_searing_spotlights.py:22-75``).
"""
from __future__ import annotations

from . import _lib as L

SCALE = 0.25


class EnvSpec:
    """Static description of one env id: option defaults, spaces, and the C-ABI family behind it."""

    def __init__(self, env_id, family, variant, defaults, action_nvec, state_bytes, info_keys, check=None,
                 has_ground_truth=False, gt_dim=2):
        self.id = env_id
        self.family = family            # "mm" | "mp" | "ss" | "ep": mg_<family>_reset / _step
        self.variant = variant
        self.default_reset_parameters = defaults
        self.action_nvec = action_nvec  # [3, 3] = MultiDiscrete([3, 3]); [4] = Discrete(4)
        self.state_bytes = state_bytes
        # the env record proper; Endless-MysteryPath keeps a 256-byte look-ahead block behind it (MgEPSpecRec)
        self.record_bytes = 2048 if family == "ep" else state_bytes
        self.info_keys = info_keys      # name -> column of the info matrix
        self._check = check
        self.has_ground_truth_info = has_ground_truth
        self.gt_dim = gt_dim            # columns of info["ground_truth"] (ground_truth_space.shape[0])

    def process_reset_params(self, reset_params):
        """Same contract as the reference's ``process_reset_params``: unknown key -> AssertionError,
        missing keys are filled with the defaults, range asserts as in the reference."""
        cloned_params = self.default_reset_parameters.copy()
        if reset_params is not None:
            for k, v in reset_params.items():
                assert k in cloned_params.keys(), "Provided reset parameter (" + str(k) + ") is not valid. Check spelling."
                cloned_params[k] = v
        if self._check is not None:
            self._check(cloned_params)
        return cloned_params

    # ---- options dict -> C struct ------------------------------------------------------------------
    def to_struct(self, p):
        if self.family == "mm":
            s = L.MMParams()
            s.variant = self.variant
            s.arena_size = int(p.get("arena_size", 6))
            s.allowed_commands = int(p["allowed_commands"])
            s.visual_feedback = int(bool(p["visual_feedback"]))
            s.max_steps = int(p.get("max_steps", -1))
            s.initial_command_count = int(p.get("initial_command_count", 1))
            s.command_count = L.ChoiceList.of(p.get("command_count", [1]))
            s.command_show_duration = L.ChoiceList.of(p.get("command_show_duration", [0]))
            s.command_show_delay = L.ChoiceList.of(p.get("command_show_delay", [1]))
            s.explosion_duration = L.ChoiceList.of(p["explosion_duration"])
            s.explosion_delay = L.ChoiceList.of(p["explosion_delay"])
            s.agent_scale = float(p["agent_scale"])
            s.agent_speed = float(p.get("agent_speed", 0.0))
            s.reward_command_failure = float(p["reward_command_failure"])
            s.reward_command_success = float(p["reward_command_success"])
            s.reward_episode_success = float(
                p["reward_new_command_success"] if self.variant == L.MG_MM_ENDLESS else p["reward_episode_success"])
            return s
        if self.family == "mp":
            s = L.MPParams()
            s.variant = self.variant
            s.max_steps = int(p["max_steps"])
            s.show_origin = int(bool(p["show_origin"]))
            s.show_goal = int(bool(p["show_goal"]))
            s.visual_feedback = int(bool(p["visual_feedback"]))
            s.cardinal_origin_choice = L.ChoiceList.of(p["cardinal_origin_choice"])
            s.agent_speed = float(p.get("agent_speed", 0.0))
            for k in ("agent_scale", "reward_goal", "reward_fall_off", "reward_path_progress", "reward_step"):
                setattr(s, k, float(p[k]))
            return s
        if self.family == "ss":
            s = L.SSParams()
            # keys the other Searing-Spotlights variant does not have: neutral values the kernels never read
            absent = {"num_coins": [1], "use_exit": True, "steps_per_coin": 0, "spawn_interval": 0, "coin_enabled": 1,
                      "coin_show_duration": 0}
            for name, ctype in L.SSParams._fields_:
                if name == "variant":
                    s.variant = self.variant
                    continue
                v = p[name] if name in p else absent.get(name, 0)
                if name == "num_coins":
                    s.num_coins = L.ChoiceList.of(v)
                elif ctype is L.C.c_int32:
                    setattr(s, name, int(v))
                else:
                    setattr(s, name, float(v))
            return s
        if self.family == "ep":
            s = L.EPParams()
            for name, ctype in L.EPParams._fields_:
                if name == "pad0":
                    continue
                setattr(s, name, int(p[name]) if ctype is L.C.c_int32 else float(p[name]))
            return s
        raise KeyError(self.family)


def _check_mm(p):
    assert p["allowed_commands"] >= 4 and p["allowed_commands"] <= 9
    assert p["arena_size"] >= 2 and p["arena_size"] <= 6


def _check_mmb(p):
    assert p["allowed_commands"] >= 4 and p["allowed_commands"] <= 9
    assert p["arena_size"] >= 2 and p["arena_size"] <= 6
    assert max(p["command_count"]) <= 20, "20 commands are allowed at maximum"


def _check_emm(p):
    assert p["allowed_commands"] >= 4 and p["allowed_commands"] <= 9


_MM_DEFAULTS = {
    "agent_scale": 1.0 * SCALE,
    "agent_speed": 12.0 * SCALE,
    "arena_size": 5,
    "allowed_commands": 9,
    "command_count": [10],
    "command_show_duration": [3],
    "command_show_delay": [1],
    "explosion_duration": [6],
    "explosion_delay": [18],
    "visual_feedback": True,
    "reward_command_failure": 0.0,
    "reward_command_success": 0.1,
    "reward_episode_success": 0.0,
}

_MMGRID_DEFAULTS = {
    "agent_scale": 1.0 * SCALE,
    "arena_size": 5,
    "allowed_commands": 5,
    "command_count": [10],
    "command_show_duration": [3],
    "command_show_delay": [1],
    "explosion_duration": [2],
    "explosion_delay": [6],
    "visual_feedback": True,
    "reward_command_failure": 0.0,
    "reward_command_success": 0.1,
    "reward_episode_success": 0.0,
}

_MMB_DEFAULTS = {k: v for k, v in _MM_DEFAULTS.items() if k not in ("command_show_duration", "command_show_delay")}
_MMBGRID_DEFAULTS = {k: v for k, v in _MMGRID_DEFAULTS.items() if k not in ("command_show_duration", "command_show_delay")}

_EMM_DEFAULTS = {
    "max_steps": -1,
    "agent_scale": 1.0 * SCALE,
    "agent_speed": 12.0 * SCALE,
    "allowed_commands": 9,
    "initial_command_count": 1,
    "command_show_duration": [3],
    "command_show_delay": [1],
    "explosion_duration": [6],
    "explosion_delay": [18],
    "visual_feedback": True,
    "reward_command_failure": 0.0,
    "reward_command_success": 0.1,
    "reward_new_command_success": 0.0,
}

_MP_DEFAULTS = {
    "max_steps": 512,
    "agent_scale": 1.0 * SCALE,
    "agent_speed": 12.0 * SCALE,
    "cardinal_origin_choice": [0, 1, 2, 3],
    "show_origin": False,
    "show_goal": False,
    "visual_feedback": True,
    "reward_goal": 1.0,
    "reward_fall_off": 0.0,
    "reward_path_progress": 0.1,
    "reward_step": 0.0,
}

_MPGRID_DEFAULTS = {k: v for k, v in _MP_DEFAULTS.items() if k != "agent_speed"}
_MPGRID_DEFAULTS.update({"max_steps": 128, "reward_path_progress": 0.0})

_SS_DEFAULTS = {
    "max_steps": 256,
    "initial_spawns": 4,
    "num_spawns": 30,
    "initial_spawn_interval": 30,
    "spawn_interval_threshold": 10,
    "spawn_interval_decay": 0.95,
    "spot_min_radius": 30.0 * SCALE,
    "spot_max_radius": 55.0 * SCALE,
    "spot_min_speed": 0.0025,
    "spot_max_speed": 0.0075,
    "spot_damage": 1.0,
    "visual_feedback": True,
    "black_background": False,
    "hide_chessboard": False,
    "light_dim_off_duration": 6,
    "light_threshold": 255,
    "num_coins": [1],
    "coin_scale": 1.5 * SCALE,
    "coins_visible": False,
    "use_exit": True,
    "exit_scale": 2.0 * SCALE,
    "exit_visible": False,
    "agent_speed": 12.0 * SCALE,
    "agent_health": 5,
    "agent_scale": 1.0 * SCALE,
    "agent_visible": False,
    "sample_agent_position": True,
    "show_last_action": True,
    "show_last_positive_reward": True,
    "reward_inside_spotlight": 0.0,
    "reward_outside_spotlight": 0.0,
    "reward_death": 0.0,
    "reward_exit": 1.0,
    "reward_max_steps": 0.0,
    "reward_coin": 0.25,
}

_ESS_DEFAULTS = {
    "max_steps": -1,
    "steps_per_coin": 160,
    "initial_spawns": 3,
    "spawn_interval": 50,
    "spot_min_radius": 30.0 * SCALE,
    "spot_max_radius": 55.0 * SCALE,
    "spot_min_speed": 0.0025,
    "spot_max_speed": 0.0075,
    "spot_damage": 1.0,
    "visual_feedback": True,
    "black_background": False,
    "hide_chessboard": False,
    "light_dim_off_duration": 6,
    "light_threshold": 255,
    "coin_enabled": True,
    "coin_scale": 1.5 * SCALE,
    "coin_show_duration": 6,
    "coins_visible": False,
    "agent_speed": 12.0 * SCALE,
    "agent_health": 10,
    "agent_scale": 1.0 * SCALE,
    "agent_visible": False,
    "sample_agent_position": True,
    "show_last_action": True,
    "show_last_positive_reward": True,
    "reward_inside_spotlight": 0.0,
    "reward_outside_spotlight": 0.0,
    "reward_death": 0.0,
    "reward_coin": 0.25,
}

_EMP_DEFAULTS = {
    "max_steps": -1,
    "agent_scale": 1.0 * SCALE,
    "agent_speed": 12.0 * SCALE,
    "show_origin": False,
    "show_past_path": True,
    "show_background": False,
    "show_stamina": False,
    "visual_feedback": True,
    "camera_offset_scale": 5.0,
    "stamina_level": 20,
    "reward_fall_off": 0.0,
    "reward_path_progress": 0.1,
    "reward_path_progress_dense": 0.0,
    "reward_step": 0.0,
}

_MM_INFO = {"reward": L.MG_INFO_REWARD, "length": L.MG_INFO_LENGTH, "success": L.MG_INFO_SUCCESS,
            "commands_completed": L.MG_INFO_A}
_EMM_INFO = {"reward": L.MG_INFO_REWARD, "length": L.MG_INFO_LENGTH, "commands_completed": L.MG_INFO_A,
             "max_command_sequence": L.MG_INFO_B}
_MP_INFO = {"reward": L.MG_INFO_REWARD, "length": L.MG_INFO_LENGTH, "success": L.MG_INFO_SUCCESS,
            "num_fails": L.MG_INFO_A}
_SS_INFO = {"reward": L.MG_INFO_REWARD, "length": L.MG_INFO_LENGTH, "success": L.MG_INFO_SUCCESS,
            "agent_health": L.MG_INFO_A, "coins_collected": L.MG_INFO_B}
_ESS_INFO = {"reward": L.MG_INFO_REWARD, "length": L.MG_INFO_LENGTH, "agent_health": L.MG_INFO_A,
             "coins_collected": L.MG_INFO_B}
_EMP_INFO = {"reward": L.MG_INFO_REWARD, "length": L.MG_INFO_LENGTH, "num_fails": L.MG_INFO_A, "max_x": L.MG_INFO_B,
             "tiles_visited": L.MG_INFO_C}

SPECS = {
    s.id: s for s in (
        EnvSpec("MortarMayhem-v0", "mm", L.MG_MM_CONTINUOUS, _MM_DEFAULTS, [3, 3], L.MG_MM_STATE_BYTES, _MM_INFO,
                _check_mm),
        EnvSpec("MortarMayhem-Grid-v0", "mm", L.MG_MM_GRID, _MMGRID_DEFAULTS, [4], L.MG_MM_STATE_BYTES, _MM_INFO,
                _check_mm),
        EnvSpec("Endless-MortarMayhem-v0", "mm", L.MG_MM_ENDLESS, _EMM_DEFAULTS, [3, 3], L.MG_MM_STATE_BYTES,
                _EMM_INFO, _check_emm, has_ground_truth=True),
        EnvSpec("MysteryPath-v0", "mp", L.MG_MP_CONTINUOUS, _MP_DEFAULTS, [3, 3], L.MG_MP_STATE_BYTES, _MP_INFO),
        EnvSpec("MysteryPath-Grid-v0", "mp", L.MG_MP_GRID, _MPGRID_DEFAULTS, [4], L.MG_MP_STATE_BYTES, _MP_INFO),
        EnvSpec("MortarMayhemB-v0", "mm", L.MG_MM_B, _MMB_DEFAULTS, [3, 3], L.MG_MM_STATE_BYTES, _MM_INFO, _check_mmb),
        EnvSpec("MortarMayhemB-Grid-v0", "mm", L.MG_MM_B_GRID, _MMBGRID_DEFAULTS, [4], L.MG_MM_STATE_BYTES, _MM_INFO,
                _check_mmb),
        EnvSpec("SearingSpotlights-v0", "ss", L.MG_SS_FINITE, _SS_DEFAULTS, [3, 3], L.MG_SS_STATE_BYTES, _SS_INFO),
        EnvSpec("Endless-SearingSpotlights-v0", "ss", L.MG_SS_ENDLESS, _ESS_DEFAULTS, [3, 3], L.MG_SS_STATE_BYTES,
                _ESS_INFO, has_ground_truth=True, gt_dim=4),
        EnvSpec("Endless-MysteryPath-v0", "ep", 0, _EMP_DEFAULTS, [4], L.MG_EP_STATE_BYTES, _EMP_INFO,
                has_ground_truth=True, gt_dim=3),
    )
}

# ids the reference registers that this engine does not accelerate yet
NOT_BUILT = ()


def spec(env_id: str) -> EnvSpec:
    if env_id in SPECS:
        return SPECS[env_id]
    if env_id in NOT_BUILT:
        raise NotImplementedError(f"{env_id} is registered by the reference but not built in this engine yet")
    raise KeyError(f"unknown env id {env_id!r}; known: {sorted(SPECS)}")
