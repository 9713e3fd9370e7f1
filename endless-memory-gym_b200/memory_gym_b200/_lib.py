"""ctypes binding of ``libmemgym_b200.so`` (C ABI declared in ``include/memgym_b200.h``).

The library is the product: every reset/step goes through it and there is no Python or CPU
fallback — if the shared object is missing, or no CUDA device is usable, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmemgym_b200.so")

MG_ABI_VERSION = 7
MG_OBS_BYTES = 84 * 84 * 3
MG_DELTA_HDR_BYTES = 176
MG_DEBUG_DIM = 336
MG_DEBUG_BYTES = 336 * 336 * 3
MG_INFO_COLS = 10
MG_GT_COLS = 4
MG_MAX_CHOICES = 16
MG_MM_MAX_COMMANDS = 152
MG_MM_STATE_BYTES = 256
MG_MP_STATE_BYTES = 128
MG_SS_STATE_BYTES = 1152
MG_EP_STATE_BYTES = 4096
MG_EP_MAX_SEGMENTS = 45
MG_FLAG_AUTORESET = 1
MG_FLAG_STREAM_OBS = 2
MG_FLAG_NO_HELPERS = 4
MG_FLAG_NO_PDL = 8
MG_MM_CONTINUOUS, MG_MM_GRID, MG_MM_ENDLESS, MG_MM_B, MG_MM_B_GRID = 0, 1, 2, 3, 4
MG_MP_CONTINUOUS, MG_MP_GRID = 0, 1
MG_MM_B_VECTOR = 180
(MG_INFO_REWARD, MG_INFO_LENGTH, MG_INFO_SUCCESS, MG_INFO_A, MG_INFO_B, MG_INFO_GT0, MG_INFO_GT1, MG_INFO_GT2,
 MG_INFO_GT3, MG_INFO_C) = range(10)
MG_SS_FINITE, MG_SS_ENDLESS = 0, 1


class ChoiceList(C.Structure):
    _fields_ = [("n", C.c_int32), ("v", C.c_int32 * MG_MAX_CHOICES)]

    @classmethod
    def of(cls, values):
        values = [int(v) for v in values]
        if not 1 <= len(values) <= MG_MAX_CHOICES:
            raise ValueError(f"list-valued reset option must have 1..{MG_MAX_CHOICES} entries, got {len(values)}")
        out = cls()
        out.n = len(values)
        for i, v in enumerate(values):
            out.v[i] = v
        return out


class MMParams(C.Structure):
    _fields_ = [
        ("variant", C.c_int32), ("arena_size", C.c_int32), ("allowed_commands", C.c_int32),
        ("visual_feedback", C.c_int32), ("max_steps", C.c_int32), ("initial_command_count", C.c_int32),
        ("command_count", ChoiceList), ("command_show_duration", ChoiceList), ("command_show_delay", ChoiceList),
        ("explosion_duration", ChoiceList), ("explosion_delay", ChoiceList),
        ("agent_scale", C.c_double), ("agent_speed", C.c_double), ("reward_command_failure", C.c_double),
        ("reward_command_success", C.c_double), ("reward_episode_success", C.c_double),
    ]


class MPParams(C.Structure):
    _fields_ = [
        ("variant", C.c_int32), ("pad0", C.c_int32),
        ("max_steps", C.c_int32), ("show_origin", C.c_int32), ("show_goal", C.c_int32),
        ("visual_feedback", C.c_int32), ("cardinal_origin_choice", ChoiceList),
        ("agent_scale", C.c_double), ("agent_speed", C.c_double), ("reward_goal", C.c_double),
        ("reward_fall_off", C.c_double), ("reward_path_progress", C.c_double), ("reward_step", C.c_double),
    ]


class SSParams(C.Structure):
    _fields_ = [
        ("max_steps", C.c_int32), ("initial_spawns", C.c_int32), ("num_spawns", C.c_int32),
        ("visual_feedback", C.c_int32), ("black_background", C.c_int32), ("hide_chessboard", C.c_int32),
        ("light_dim_off_duration", C.c_int32), ("light_threshold", C.c_int32), ("coins_visible", C.c_int32),
        ("use_exit", C.c_int32), ("exit_visible", C.c_int32), ("agent_visible", C.c_int32),
        ("sample_agent_position", C.c_int32), ("show_last_action", C.c_int32),
        ("show_last_positive_reward", C.c_int32), ("variant", C.c_int32), ("num_coins", ChoiceList),
        ("initial_spawn_interval", C.c_double), ("spawn_interval_threshold", C.c_double),
        ("spawn_interval_decay", C.c_double), ("spot_min_radius", C.c_double), ("spot_max_radius", C.c_double),
        ("spot_min_speed", C.c_double), ("spot_max_speed", C.c_double), ("spot_damage", C.c_double),
        ("coin_scale", C.c_double), ("exit_scale", C.c_double), ("agent_speed", C.c_double),
        ("agent_health", C.c_double), ("agent_scale", C.c_double), ("reward_inside_spotlight", C.c_double),
        ("reward_outside_spotlight", C.c_double), ("reward_death", C.c_double), ("reward_exit", C.c_double),
        ("reward_max_steps", C.c_double), ("reward_coin", C.c_double),
        ("steps_per_coin", C.c_int32), ("spawn_interval", C.c_int32), ("coin_enabled", C.c_int32),
        ("coin_show_duration", C.c_int32),
    ]


class EPParams(C.Structure):
    _fields_ = [
        ("max_steps", C.c_int32), ("show_origin", C.c_int32), ("show_past_path", C.c_int32),
        ("show_background", C.c_int32), ("show_stamina", C.c_int32), ("visual_feedback", C.c_int32),
        ("stamina_level", C.c_int32), ("pad0", C.c_int32),
        ("agent_scale", C.c_double), ("agent_speed", C.c_double), ("camera_offset_scale", C.c_double),
        ("reward_fall_off", C.c_double), ("reward_path_progress", C.c_double),
        ("reward_path_progress_dense", C.c_double), ("reward_step", C.c_double),
    ]


class StepIO(C.Structure):
    _fields_ = [
        ("actions", C.c_void_p), ("obs", C.c_void_p), ("final_obs", C.c_void_p), ("reward", C.c_void_p),
        ("done", C.c_void_p), ("info", C.c_void_p), ("ground_truth", C.c_void_p), ("reset_mask", C.c_void_p),
        ("flags", C.c_uint32), ("warps_per_cta", C.c_uint32), ("vector_obs", C.c_void_p), ("trace", C.c_void_p),
        ("debug_obs", C.c_void_p),
    ]


class RolloutIO(C.Structure):
    _fields_ = [
        ("actions", C.c_void_p), ("obs", C.c_void_p), ("final_obs", C.c_void_p), ("reward", C.c_void_p),
        ("done", C.c_void_p), ("info", C.c_void_p), ("ground_truth", C.c_void_p),
        ("n_steps", C.c_int32), ("ring", C.c_int32), ("info_per_step", C.c_int32), ("flags", C.c_uint32),
        ("warps_per_cta", C.c_uint32), ("pad0", C.c_uint32), ("trace", C.c_void_p), ("vector_obs", C.c_void_p),
        ("sched", C.c_void_p),
    ]


class MemGymError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the shared library once. Missing library = hard error (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MemGymError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C endless-memory-gym_b200/csrc`. memory_gym_b200 has no CPU / Python fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, u64p = C.c_void_p, C.c_int, C.POINTER(C.c_uint64)
    L.mg_abi_version.restype = i32
    L.mg_last_error.restype = C.c_char_p
    L.mg_device_info.argtypes = [C.POINTER(i32)] * 3 + [C.POINTER(C.c_size_t)]
    L.mg_seed_pcg64.argtypes = [vp, i32, vp]
    L.mg_obs_to_policy_bf16.argtypes = [vp, vp, i32, vp]
    L.mg_obs_delta_stage_bytes.restype = C.c_size_t
    L.mg_obs_delta_stage_bytes.argtypes = [i32]
    L.mg_obs_delta_encode.argtypes = [vp, vp, i32, vp, i32, vp]
    L.mg_obs_delta_decode.argtypes = [vp, vp, i32, i32, u64p]
    L.mg_obs_delta_scatter.argtypes = [vp, vp, i32, vp, i32, vp, vp]
    L.mg_selftest_pcg64_host.argtypes = [vp, vp, vp, vp, i32, vp, vp]
    for fam, P in (("mm", MMParams), ("mp", MPParams), ("ss", SSParams), ("ep", EPParams)):
        getattr(L, f"mg_{fam}_check_params").argtypes = [C.POINTER(P)]
        getattr(L, f"mg_{fam}_reset").argtypes = [C.POINTER(P), vp, vp, vp, C.POINTER(StepIO), i32, vp]
        getattr(L, f"mg_{fam}_step").argtypes = [C.POINTER(P), vp, vp, C.POINTER(StepIO), i32, vp]
        getattr(L, f"mg_{fam}_step_many").argtypes = [C.POINTER(P), vp, vp, C.POINTER(StepIO), i32, i32, vp]
        getattr(L, f"mg_{fam}_rollout").argtypes = [C.POINTER(P), vp, vp, C.POINTER(RolloutIO), i32, vp]
        b = getattr(L, f"mg_{fam}_build_assets")
        b.restype = C.c_long
        b.argtypes = [C.POINTER(P), i32, vp, C.c_size_t] if fam == "mm" else [C.POINTER(P), vp, C.c_size_t]
    if L.mg_abi_version() != MG_ABI_VERSION:
        raise MemGymError(f"ABI mismatch: library {L.mg_abi_version()}, binding {MG_ABI_VERSION}")
    _lib = L
    return L


def last_error() -> str:
    return lib().mg_last_error().decode()


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise MemGymError(f"{what}: {last_error()} (code {rc})")
    raise MemGymError(f"{what}: CUDA error {rc}: {last_error()}")
