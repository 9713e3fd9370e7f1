"""Batched Memory-Gym environments: the reference's gymnasium surface with a leading batch dim.

``MemoryGymVecEnv`` mirrors the reference env classes (``reset(seed=, options=)`` /
``step(action)`` / ``action_space`` / ``observation_space`` / ``default_reset_parameters`` /
``process_reset_params`` / ``has_ground_truth_info`` / ``max_episode_steps`` / ``render`` / ``close``,
e.g. ``memory_gym/mortar_mayhem.py:56-91,206-272,274-369``) for N parallel instances whose state
lives in HBM. Every reset/step is ONE kernel launch through the C ABI of ``libmemgym_b200.so``;
PyTorch only provides device memory and the CUDA stream. There is no CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L
from . import spaces
from .specs import spec as _spec


class _CudaBackend:
    """The product path: ctypes calls into libmemgym_b200.so on the caller's current CUDA stream."""

    name = "cuda"

    def __init__(self):
        self.lib = L.lib()
        self._fn = {}

    def fn(self, family, what):
        """``mg_<family>_<what>`` of the library (resolved once)."""
        key = (family, what)
        f = self._fn.get(key)
        if f is None:
            f = self._fn[key] = getattr(self.lib, f"mg_{family}_{what}")
        return f

    def check_device(self, device):
        if device.type != "cuda":
            raise L.MemGymError(f"memory_gym_b200 runs on CUDA devices only (got device={device}); there is no CPU path")
        if not torch.cuda.is_available():
            raise L.MemGymError("no CUDA device available; memory_gym_b200 has no CPU fallback")

    def stream(self, device):
        return torch.cuda.current_stream(device).cuda_stream

    def build_assets(self, family, params):
        fn = getattr(self.lib, f"mg_{family}_build_assets")
        args = (C.byref(params), 0) if family == "mm" else (C.byref(params),)
        size = fn(*args, None, 0)
        if size < 0:
            raise L.MemGymError(f"mg_{family}_build_assets: {L.last_error()}")
        buf = np.zeros(size, dtype=np.uint8)
        got = fn(*args, buf.ctypes.data, size)
        if got != size:
            raise L.MemGymError(f"mg_{family}_build_assets: size changed {size} -> {got}")
        return buf

    def seed(self, seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        out = np.empty((seeds.shape[0], 4), dtype=np.uint64)
        L.check(self.lib.mg_seed_pcg64(seeds.ctypes.data, seeds.shape[0], out.ctypes.data), "mg_seed_pcg64")
        return out

    def check_params(self, family, params):
        L.check(getattr(self.lib, f"mg_{family}_check_params")(C.byref(params)), f"mg_{family}_check_params")

    def reset(self, family, params, assets, state, rng4, io, n, stream):
        L.check(getattr(self.lib, f"mg_{family}_reset")(C.byref(params), assets, state, rng4, C.byref(io), n, stream),
                f"mg_{family}_reset")

    def step(self, family, params, assets, state, io, n, stream):
        rc = self.fn(family, "step")(C.byref(params), assets, state, C.byref(io), n, stream)
        if rc:
            L.check(rc, f"mg_{family}_step")

    def rollout(self, family, params, assets, state, rio, n, stream):
        rc = self.fn(family, "rollout")(C.byref(params), assets, state, C.byref(rio), n, stream)
        if rc:
            L.check(rc, f"mg_{family}_rollout")

    def step_many(self, family, params, assets, state, ios, n_steps, n, stream):
        rc = self.fn(family, "step_many")(C.byref(params), assets, state, ios, n_steps, n, stream)
        if rc:
            L.check(rc, f"mg_{family}_step_many")


class MemoryGymVecEnv:
    """N parallel instances of one Memory-Gym env id.

    * ``reset(seed=None|int|sequence, options=None|dict) -> (obs uint8[N,84,84,3], info)``: an int
      seed seeds env ``i`` with ``seed + env_offset + i``; ``None`` continues every env's stream
      (first use: OS entropy), exactly like ``env.reset()`` in gymnasium 0.29.
    * ``step(actions int64[N,2] | int64[N]) -> (obs, reward float64[N], done bool[N], truncated bool[N], info)``.
      With ``auto_reset`` (default) finished envs are reset inside the same launch: ``obs`` is the new
      episode's first frame, the terminal frame is in ``info["final_observation"]`` (if ``final_obs``)
      and the episode statistics in ``info[<key>]`` (valid where ``done``).
    Returned tensors are views of persistent device buffers that the next call overwrites.
    Observation layout per env is pygame's ``surfarray.array3d``: ``[x][y][c]``.
    """

    metadata = {"render_modes": ["rgb_array", "debug_rgb_array"], "render_fps": 25}

    def __init__(self, env_id, num_envs=1, device="cuda", auto_reset=True, final_obs=False, env_offset=0,
                 warps_per_cta=0, stream_obs=None, host_results=False, render_mode=None, _backend=None):
        self.spec_ = _spec(env_id)
        self.env_id = env_id
        self.num_envs = int(num_envs)
        if self.num_envs < 1:
            raise ValueError("num_envs must be >= 1")
        if render_mode not in (None, "rgb_array", "debug_rgb_array"):
            raise ValueError(f"render_mode {render_mode!r}: the batched engine renders 'rgb_array' and 'debug_rgb_array'")
        if render_mode == "debug_rgb_array" and env_id in ("MortarMayhemB-v0", "MortarMayhemB-Grid-v0"):
            # the reference's B variants have no _build_debug_surface at all
            raise L.MemGymError(f"render_mode='debug_rgb_array' is not available for {env_id}")
        self.render_mode = render_mode
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None and torch.cuda.is_available():
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._be = _backend if _backend is not None else _CudaBackend()
        self._be.check_device(self.device)
        self.auto_reset = bool(auto_reset)
        self.env_offset = int(env_offset)
        self.warps_per_cta = int(warps_per_cta)
        self.default_reset_parameters = self.spec_.default_reset_parameters
        self.has_ground_truth_info = self.spec_.has_ground_truth_info
        nvec = self.spec_.action_nvec
        self.single_action_space = spaces.Discrete(nvec[0]) if len(nvec) == 1 else spaces.MultiDiscrete(nvec)
        self.single_observation_space = spaces.Box(low=0, high=255, shape=(84, 84, 3), dtype=np.uint8)
        self.action_space, self.observation_space = self.single_action_space, self.single_observation_space
        if self.has_ground_truth_info:
            self.ground_truth_space = spaces.Box(low=0.0, high=1.0, shape=(self.spec_.gt_dim,), dtype=np.float32)
        n, dev = self.num_envs, self.device
        self.state = torch.zeros((n, self.spec_.state_bytes), dtype=torch.uint8, device=dev)
        self.obs = torch.zeros((n, 84, 84, 3), dtype=torch.uint8, device=dev)
        self.final_obs = torch.zeros((n, 84, 84, 3), dtype=torch.uint8, device=dev) if final_obs else None
        # reward (f64) and done (u8) share one allocation: `results` moves both to the host with a single copy.
        # host_results=True puts it into PINNED HOST memory instead: the step kernel writes rewards and dones straight
        # over PCIe (zero-copy), `reward` / `done` are CPU tensors that are valid once the step has finished on the
        # stream — no device->host copy operation sits between two step launches (a host-side policy loop).
        self.host_results = bool(host_results)
        if self.host_results:
            self.results = torch.zeros(9 * n, dtype=torch.uint8).pin_memory()
        else:
            self.results = torch.zeros(9 * n, dtype=torch.uint8, device=dev)
        self.reward = self.results[:8 * n].view(torch.float64)
        self.done = self.results[8 * n:]
        self.truncated = torch.zeros(n, dtype=torch.bool, device=dev)
        self.info = torch.zeros((n, L.MG_INFO_COLS), dtype=torch.float64, device=dev)
        # the library writes [n][MG_GT_COLS]; the env publishes the columns its reference counterpart has
        self._gt_full = torch.zeros((n, L.MG_GT_COLS), dtype=torch.float64, device=dev) if self.has_ground_truth_info else None
        self.ground_truth = self._gt_full[:, :self.spec_.gt_dim] if self.has_ground_truth_info else None
        # MortarMayhemB: Dict observation {"visual_observation", "vector_observation"} (mortar_mayhem_b.py:84-97)
        self.task_b = self.spec_.family == "mm" and self.spec_.variant in (L.MG_MM_B, L.MG_MM_B_GRID)
        self.vector_obs = torch.zeros((n, L.MG_MM_B_VECTOR), dtype=torch.float32, device=dev) if self.task_b else None
        self._actions_dev = torch.zeros((n, len(nvec)) if len(nvec) > 1 else (n,), dtype=torch.int64, device=dev)
        self._io = L.StepIO()
        self._io.obs = self.obs.data_ptr()
        # step() hot path: the views it returns are built once (they alias the library's output buffers)
        self._done_bool = self.done.view(torch.bool)
        self._info_views = {"episode": self.info, "_episode": self._done_bool}
        for k_, col in self.spec_.info_keys.items():
            self._info_views[k_] = self.info[:, col]
        if self.final_obs is not None:
            self._info_views["final_observation"] = self.final_obs
        if self.has_ground_truth_info:
            self._info_views["ground_truth"] = self.ground_truth
            self._info_views["final_ground_truth"] = self.info[:, L.MG_INFO_GT0:L.MG_INFO_GT0 + self.spec_.gt_dim]
        self._step_fn = None
        self._results_views = {}
        self._step_c = self._be.fn(self.spec_.family, "step") if hasattr(self._be, "fn") else None
        self._io_ref = C.byref(self._io)
        self._obs_ptr, self._reward_ptr, self._done_ptr = self.obs.data_ptr(), self.reward.data_ptr(), self.done.data_ptr()
        self._params_ref = None
        self._io.final_obs = self.final_obs.data_ptr() if final_obs else None
        self._io.reward = self.reward.data_ptr()
        self._io.done = self.done.data_ptr()
        self._io.info = self.info.data_ptr()
        self._io.ground_truth = self._gt_full.data_ptr() if self.has_ground_truth_info else None
        self._io.vector_obs = self.vector_obs.data_ptr() if self.task_b else None
        # render_mode="debug_rgb_array": every reset / step also paints the ground-truth render of the reference
        # (_build_debug_surface, 336 x 336 x 3 per env, [y][x][c]) — as if render() were called once per reset / step
        self.debug = None
        if self.render_mode == "debug_rgb_array":
            self.debug = torch.zeros((n, L.MG_DEBUG_DIM, L.MG_DEBUG_DIM, 3), dtype=torch.uint8, device=dev)
            self._io.debug_obs = self.debug.data_ptr()

        # rollouts: a global work counter the kernels keep at zero between launches (dynamic env -> warp scheduling
        # across SMs). MG_LOCAL_SCHED=1 switches it off (measurement knob): every SM works through its own share.
        self._sched = None if os.environ.get("MG_LOCAL_SCHED") else torch.zeros(2, dtype=torch.int32, device=dev)
        # stream_obs: write observations with an L2 evict-first policy. Default: on for rollout() (frames are
        # not re-read soon; measured +13 % HBM write throughput), off for step() (the policy reads obs next).
        self._stream_step = L.MG_FLAG_STREAM_OBS if stream_obs else 0
        self._stream_rollout = L.MG_FLAG_STREAM_OBS if (stream_obs is None or stream_obs) else 0
        # measurement knobs, read when the env is made and passed to the library as explicit flags (it reads no
        # environment variable itself): MG_HELPERS=0 = no helper warps, MG_DISABLE_PDL = no programmatic dependent launch
        self._knobs = (L.MG_FLAG_NO_HELPERS if os.environ.get("MG_HELPERS") == "0" else 0) | \
                      (L.MG_FLAG_NO_PDL if os.environ.get("MG_DISABLE_PDL") else 0)
        if self.render_mode == "debug_rgb_array":
            # (Endless-MysteryPath: the render shows the whole path, so an auto-reset generates it inline instead of
            # handing it to helper warps that finish after the frame)
            self._knobs |= L.MG_FLAG_NO_HELPERS
        self._io.flags = (L.MG_FLAG_AUTORESET if self.auto_reset else 0) | self._stream_step | self._knobs
        self._io.warps_per_cta = self.warps_per_cta
        self.reset_params = None
        self._params = None
        self._params_key = None
        self._assets = None
        self._assets_ptr = self._state_ptr = 0
        self._seeded = False
        self.max_episode_steps = None

    @property
    def records(self):
        """The env records proper (`state` minus what the engine keeps behind a record for its own use: the paths
        Endless-MysteryPath rollouts generate ahead of a reset). Two envs with equal records behave identically."""
        return self.state[:, :self.spec_.record_bytes]

    # ---- reference-compatible helpers ----------------------------------------------------------------
    def process_reset_params(self, reset_params):
        return self.spec_.process_reset_params(reset_params)

    def _set_options(self, options):
        p = self.spec_.process_reset_params(options)
        key = repr(sorted(p.items()))
        if key != self._params_key:
            params = self.spec_.to_struct(p)
            self._be.check_params(self.spec_.family, params)
            blob = self._be.build_assets(self.spec_.family, params)
            # 128-byte aligned device copy (TMA bulk loads read the background templates from it)
            raw = torch.empty(blob.shape[0] + 128, dtype=torch.uint8, device=self.device)
            off = (-raw.data_ptr()) % 128
            self._assets = raw[off:off + blob.shape[0]]
            self._assets.copy_(torch.from_numpy(blob))
            self._assets_ptr, self._state_ptr = self._assets.data_ptr(), self.state.data_ptr()
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)  # the kernels' prologue reads the blob before stream order applies
            self._params, self._params_key = params, key
            self._params_ref = C.byref(params)
        self.reset_params = p
        self.max_episode_steps = self._max_episode_steps(p)

    def _max_episode_steps(self, p):
        if self.spec_.family == "mm" and self.spec_.variant != L.MG_MM_ENDLESS:
            # calc_max_episode_steps (pygame_assets.py:420-436), called with delay/duration swapped
            # (mortar_mayhem.py:222-226); informational only — the env has no step limit
            cc = max(p["command_count"])
            sd, sl = max(p.get("command_show_duration", [0])), max(p.get("command_show_delay", [0]))
            ed, el = max(p["explosion_delay"]), max(p["explosion_duration"])
            return (sd + sl) * cc + (ed + el) * cc - el + 1
        return p.get("max_steps")

    def _seed_table(self, seed):
        n = self.num_envs
        if seed is None:
            if self._seeded:
                return None
            seeds = np.frombuffer(os.urandom(8 * n), dtype=np.uint64)  # gymnasium: OS entropy on first use
        elif isinstance(seed, (int, np.integer)):
            seeds = np.uint64(int(seed) + self.env_offset) + np.arange(n, dtype=np.uint64)
        else:
            seeds = np.asarray(list(seed), dtype=np.uint64)
            if seeds.shape != (n,):
                raise ValueError(f"seed sequence must have {n} entries")
        return torch.from_numpy(self._be.seed(seeds)).to(self.device)

    # ---- gymnasium surface ---------------------------------------------------------------------------
    def reset(self, seed=None, return_info=True, options=None, mask=None):
        """Reset all envs (or those selected by ``mask``, a bool/uint8 tensor of shape [N])."""
        self._set_options(options)
        rng4 = self._seed_table(seed)
        self._seeded = True
        mask_t = None
        if mask is not None:
            mask_t = torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()
        self._io.reset_mask = mask_t.data_ptr() if mask_t is not None else None
        self._io.actions = None
        self._io.obs = self.obs.data_ptr()
        self._be.reset(self.spec_.family, self._params, self._assets.data_ptr(), self.state.data_ptr(),
                       rng4.data_ptr() if rng4 is not None else None, self._io, self.num_envs,
                       self._be.stream(self.device))
        self._io.reset_mask = None
        self._keep = (rng4, mask_t)  # keep the inputs alive until the stream has consumed them
        info = {}
        if self.has_ground_truth_info:
            info["ground_truth"] = self.ground_truth
        return self._wrap_obs(self.obs), info

    def _wrap_obs(self, obs):
        if self.task_b:
            return {"visual_observation": obs, "vector_observation": self.vector_obs}
        return obs

    def step(self, actions, obs_out=None, results_out=None):
        """``obs_out``: optional uint8[N,84,84,3] device tensor (e.g. a slot of a rollout buffer) that
        receives this step's observations instead of the env's own ``obs`` buffer. ``results_out``: optional
        uint8[9 N] tensor — device, or PINNED host memory (written by the kernel over PCIe, no copy operation) — that
        receives this step's rewards (float64[N], first 8 N bytes) and dones (uint8[N]) instead of ``env.results``.
        ``actions`` may be a device tensor or a pinned host tensor (read by the kernel in place)."""
        if self._params is None:
            raise L.MemGymError("step() called before reset()")
        if isinstance(actions, torch.Tensor) and actions.dtype == torch.int64 and actions.is_contiguous() and \
                (actions.device == self.device or (actions.device.type == "cpu" and actions.is_pinned())):
            # device tensor, or PINNED host memory: the kernel reads the actions straight from it (zero-copy over PCIe;
            # unified addressing: the host pointer is valid on the device) — no host->device copy operation
            a = actions
        else:
            a = self._actions_dev
            a.copy_(torch.as_tensor(np.asarray(actions) if not isinstance(actions, torch.Tensor) else actions)
                    .reshape(a.shape), non_blocking=True)
        if a.numel() != self._actions_dev.numel():
            raise ValueError(f"actions must have shape {tuple(self._actions_dev.shape)}")
        io = self._io
        io.actions = a.data_ptr()
        obs = self.obs
        if obs_out is not None:
            if obs_out.shape != self.obs.shape or obs_out.dtype != torch.uint8 or obs_out.device != self.device \
                    or not obs_out.is_contiguous():
                raise ValueError("obs_out must be a contiguous uint8[N,84,84,3] tensor on the env's device")
            obs = obs_out
        io.obs = obs.data_ptr()
        reward, done = self.reward, self._done_bool
        if results_out is not None:
            views = self._results_views.get(results_out.data_ptr())
            if views is None:
                n = self.num_envs
                if results_out.dtype != torch.uint8 or results_out.numel() != 9 * n or not results_out.is_contiguous() or \
                        not (results_out.device == self.device or (results_out.device.type == "cpu" and results_out.is_pinned())):
                    raise ValueError("results_out must be a contiguous uint8[9 N] tensor on the env's device or in pinned host memory")
                flat = results_out.view(-1)
                views = (results_out, flat[:8 * n].view(torch.float64), flat[8 * n:].view(torch.bool))
                if len(self._results_views) > 8:
                    self._results_views.clear()
                self._results_views[results_out.data_ptr()] = views
            _, reward, done = views
            io.reward, io.done = reward.data_ptr(), done.data_ptr()
        self._be.step(self.spec_.family, self._params, self._assets_ptr, self._state_ptr, io,
                      self.num_envs, self._be.stream(self.device))
        if results_out is not None:
            io.reward, io.done = self.reward.data_ptr(), self.done.data_ptr()
            info = dict(self._info_views)
            info["_episode"] = done
            return self._wrap_obs(obs), reward, done, self.truncated, info
        return self._wrap_obs(obs), self.reward, self._done_bool, self.truncated, dict(self._info_views)

    def _launch_step_ptrs(self, actions_ptr, reward_ptr, done_ptr, stream):
        """One step launch from raw pointers (validated by the caller): the lean path behind ``HostPipe.submit``."""
        io = self._io
        io.actions, io.obs, io.reward, io.done = actions_ptr, self._obs_ptr, reward_ptr, done_ptr
        rc = self._step_c(self._params_ref, self._assets_ptr, self._state_ptr, self._io_ref, self.num_envs, stream)
        io.reward, io.done = self._reward_ptr, self._done_ptr
        if rc:
            L.check(rc, f"mg_{self.spec_.family}_step")

    def _check_rollout_buffers(self, actions, obs_out, reward_out, done_out, info_out):
        if self._params is None:
            raise L.MemGymError("rollout() called before reset()")
        T = actions.shape[0]
        for t_, name in ((actions, "actions"), (obs_out, "obs_out"), (reward_out, "reward_out"), (done_out, "done_out")):
            if t_.device != self.device or not t_.is_contiguous():
                raise ValueError(f"{name} must be a contiguous tensor on the env's device")
        if actions.dtype != torch.int64 or obs_out.dtype != torch.uint8 or reward_out.dtype != torch.float64 \
                or done_out.element_size() != 1:
            raise ValueError("dtypes: actions int64, obs_out uint8, reward_out float64, done_out bool/uint8")
        if tuple(actions.shape[1:]) != tuple(self._actions_dev.shape) or tuple(obs_out.shape[1:]) != (self.num_envs, 84, 84, 3) \
                or tuple(reward_out.shape) != (T, self.num_envs) or tuple(done_out.shape) != (T, self.num_envs):
            raise ValueError("rollout buffer shapes do not match (T, num_envs)")
        if info_out is not None and (tuple(info_out.shape) != (T, self.num_envs, L.MG_INFO_COLS)
                                     or info_out.dtype != torch.float64 or not info_out.is_contiguous()):
            raise ValueError(f"info_out must be a contiguous float64[T, N, {L.MG_INFO_COLS}] tensor")
        return T

    def rollout(self, actions, obs_out, reward_out, done_out, info_out=None):
        """``T`` consecutive steps with given actions in ONE kernel launch (``mg_*_rollout``): every warp runs
        all T steps of its envs back to back with the env record resident on chip. The reference's
        ``bench.py`` loop shape (pre-sampled random actions) and scripted roll-outs. Always auto-resets.

        ``actions`` int64[T, N, 2] (or [T, N]); ``obs_out`` uint8[R, N, 84, 84, 3], step t writes slot t % R
        (R = T keeps every frame, a small R is a ring); ``reward_out`` float64[T, N]; ``done_out``
        bool/uint8[T, N]; ``info_out`` optional float64[T, N, MG_INFO_COLS] (else ``self.info`` keeps each env's last
        finished episode)."""
        T = self._check_rollout_buffers(actions, obs_out, reward_out, done_out, info_out)
        r = L.RolloutIO()
        r.actions, r.obs = actions.data_ptr(), obs_out.data_ptr()
        r.final_obs = self._io.final_obs
        r.reward, r.done = reward_out.data_ptr(), done_out.data_ptr()
        r.info = info_out.data_ptr() if info_out is not None else self._io.info
        r.ground_truth = self._io.ground_truth
        r.vector_obs = self._io.vector_obs
        r.sched = self._sched.data_ptr() if self._sched is not None else None
        r.n_steps, r.ring, r.info_per_step = T, obs_out.shape[0], int(info_out is not None)
        r.flags, r.warps_per_cta = L.MG_FLAG_AUTORESET | self._stream_rollout | self._knobs, self.warps_per_cta
        r.trace = getattr(self, "_trace_ptr", None)
        self._be.rollout(self.spec_.family, self._params, self._assets.data_ptr(), self.state.data_ptr(), r,
                         self.num_envs, self._be.stream(self.device))

    def step_many(self, actions, obs_out, reward_out, done_out, info_out=None):
        """Same contract as ``rollout`` but as T separate step launches enqueued by one library call
        (``mg_*_step_many``), honouring this env's ``auto_reset`` setting."""
        T = self._check_rollout_buffers(actions, obs_out, reward_out, done_out, info_out)
        key = (T, actions.data_ptr(), obs_out.data_ptr(), obs_out.shape[0], reward_out.data_ptr(), done_out.data_ptr(),
               info_out.data_ptr() if info_out is not None else 0)
        if getattr(self, "_many_key", None) != key:
            ios = (L.StepIO * T)()
            a_stride = actions.stride(0) * 8
            o_stride = obs_out.stride(0)
            for t in range(T):
                io = ios[t]
                io.actions = actions.data_ptr() + t * a_stride
                io.obs = obs_out.data_ptr() + (t % obs_out.shape[0]) * o_stride
                io.final_obs = self._io.final_obs
                io.reward = reward_out.data_ptr() + t * self.num_envs * 8
                io.done = done_out.data_ptr() + t * self.num_envs
                io.info = (info_out.data_ptr() + t * self.num_envs * L.MG_INFO_COLS * 8) if info_out is not None \
                    else self._io.info
                io.ground_truth = self._io.ground_truth
                io.vector_obs = self._io.vector_obs
                io.reset_mask = None
                io.flags = self._io.flags
                io.warps_per_cta = self._io.warps_per_cta
            self._many_key, self._many_ios = key, ios
        self._be.step_many(self.spec_.family, self._params, self._assets.data_ptr(), self.state.data_ptr(),
                           self._many_ios, T, self.num_envs, self._be.stream(self.device))

    def capture_step(self, obs_out=None):
        """One CUDA graph for the whole host-driven step: actions pinned-host -> device, the step kernel, reward + done
        device -> pinned-host. Returns a ``StepGraph``: write the actions into ``.actions_host``, call ``.replay()``
        (asynchronous, one graph launch instead of three stream operations + the library call), wait on
        ``.done_event`` before reading ``.results_host`` (reward f64[N] then done u8[N]); observations are in ``.obs``
        on the device. The env must have been reset; reset() with other options invalidates the graph."""
        if self._params is None:
            raise L.MemGymError("capture_step() called before reset()")
        return StepGraph(self, obs_out)

    def host_pipe(self, threads=None, depth=2, mode="auto"):
        """Observations, rewards and dones in HOST memory every step — the reference's interface (numpy arrays on the
        host) for a whole batch — without shipping whole frames over PCIe: see ``HostPipe``."""
        if self._params is None:
            raise L.MemGymError("host_pipe() called before reset()")
        return HostPipe(self, threads, depth, mode)

    def policy_obs(self, obs=None, out=None):
        """The consumer step (``mg_obs_to_policy_bf16``): ``obs`` uint8[N,84,84,3] (default: this env's observation
        buffer) -> bf16[N,3,84,84] = ``obs.permute(0, 3, 1, 2) / 255``, bit-identical to torch's float32 division and
        cast, by one streaming kernel on the current stream."""
        src = self.obs if obs is None else obs
        if src.dtype != torch.uint8 or src.device != self.device or not src.is_contiguous() or tuple(src.shape[1:]) != (84, 84, 3):
            raise ValueError("obs must be a contiguous uint8[N,84,84,3] tensor on the env's device")
        n = src.shape[0]
        if out is None:
            out = torch.empty((n, 3, 84, 84), dtype=torch.bfloat16, device=self.device)
        elif out.dtype != torch.bfloat16 or out.device != self.device or not out.is_contiguous() or tuple(out.shape) != (n, 3, 84, 84):
            raise ValueError("out must be a contiguous bfloat16[N,3,84,84] tensor on the env's device")
        L.check(self._be.lib.mg_obs_to_policy_bf16(src.data_ptr(), out.data_ptr(), n, self._be.stream(self.device)),
                "mg_obs_to_policy_bf16")
        return out

    def render(self):
        """``rgb_array`` of the reference (mortar_mayhem.py:396): the observation transposed to [y][x][c];
        ``debug_rgb_array`` (:397-399): the ground-truth render uint8[N,336,336,3] of the last reset / step."""
        if self.render_mode == "debug_rgb_array":
            return self.debug
        return self.obs.permute(0, 2, 1, 3)

    def close(self):
        self.state = self.obs = self.final_obs = None

    def state_dict(self):
        """Snapshot of every env's state (PCG64 stream included) — restoring it resumes bit-exactly."""
        return {"env_id": self.env_id, "state": self.state.clone(), "params": dict(self.reset_params or {}),
                "obs": self.obs.clone(), "vector_obs": self.vector_obs.clone() if self.task_b else None}

    def load_state_dict(self, sd):
        if sd["env_id"] != self.env_id or sd["state"].shape != self.state.shape:
            raise ValueError("state_dict does not match this env")
        self._set_options(sd["params"] or None)
        self.state.copy_(sd["state"])
        self.obs.copy_(sd["obs"])
        if self.task_b:
            self.vector_obs.copy_(sd["vector_obs"])
        self._seeded = True


class StepGraph:
    """See ``MemoryGymVecEnv.capture_step``."""

    def __init__(self, env, obs_out=None):
        self.env = env
        n = env.num_envs
        self.obs = env.obs if obs_out is None else obs_out
        self.actions_host = torch.zeros_like(env._actions_dev, device="cpu").pin_memory()
        self.results_host = torch.zeros(9 * n, dtype=torch.uint8).pin_memory()
        self.done_event = torch.cuda.Event()
        self._a_dev = torch.zeros_like(env._actions_dev)
        # the warm-up below runs one real step: everything it touches is put back afterwards
        touched = (env.state, self.obs, env.results, env.info, env.final_obs, env.vector_obs, env._gt_full)
        keep = [(t, t.clone()) for t in touched if t is not None]
        side = torch.cuda.Stream(device=env.device)
        side.wait_stream(torch.cuda.current_stream(env.device))
        with torch.cuda.stream(side):
            self._body()  # warm-up outside the capture (lazy initialisation in the library / driver)
        torch.cuda.current_stream(env.device).wait_stream(side)
        torch.cuda.synchronize(env.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()
        for t, saved in keep:
            t.copy_(saved)
        torch.cuda.synchronize(env.device)

    def _body(self):
        self._a_dev.copy_(self.actions_host, non_blocking=True)
        self.env.step(self._a_dev, obs_out=self.obs)
        self.results_host.copy_(self.env.results, non_blocking=True)

    def replay(self):
        self.graph.replay()
        self.done_event.record()


class HostPipe:
    """Host-buffer stepping for a batched env: ``obs`` (uint8[N,84,84,3]), ``reward`` (float64[N]) and ``done``
    (bool[N]) live in HOST memory and are brought up to date every step, like the arrays the reference returns.

    Transport: the step kernel reads the actions from pinned host memory and writes rewards / dones into pinned host
    memory (zero-copy); a second kernel compares the new frames with the frames the host already holds and sends only
    the 16-byte chunks that changed. Lossless: ``obs`` equals the device observation bit for bit. Two modes:

    * ``mode="direct"`` (``mg_obs_delta_scatter``): the host arrays are pinned and the kernel writes every
      changed chunk to its place in them — the host does nothing but wait for the step's event. One host array per step
      in flight (``depth`` of them, used in turn; ``obs`` is the array of the step collected last and stays untouched
      until ``depth`` further steps have been collected).
    * ``mode="patch"`` (``mg_obs_delta_encode`` / ``mg_obs_delta_decode``): the chunks go densely packed into a pinned
      stage and worker threads of the library patch them into ONE pageable host array, in place.
    * ``mode="auto"`` (default): "patch" while this rank has at least 10 host cores for the patching (measured on
      B200 boxes, MortarMayhem-v0 x 4096 envs per GPU, M env-steps/s whole job, direct / patch: 1 GPU + 16 cores 20 / 29,
      2 GPUs + 24 cores 37 / 47, 4 GPUs + 32 cores 49 / 37, 8 GPUs + 32 cores 48 / 29), else "direct": scattered 16-byte
      writes drain more slowly over PCIe than packed ones, but they cost no host core.

    ``step(actions)`` = ``submit(actions)`` + ``collect()``. A loop whose actions do not depend on the newest
    observation (pre-sampled / scripted actions, the reference's bench.py shape) may keep up to ``depth`` steps in
    flight: ``submit`` step t+1 before ``collect``-ing step t. The views returned for rewards / dones belong to that
    step's slot and are reused ``depth`` steps later."""

    AUTO_MIN_CORES = 10   # "auto": host cores per rank from which patching on the host beats direct writes

    def __init__(self, env, threads=None, depth=2, mode="auto"):
        self.env = env
        n = env.num_envs
        self.depth = int(depth)
        if self.depth < 1:
            raise ValueError("depth must be >= 1")
        if mode not in ("auto", "direct", "patch"):
            raise ValueError("mode must be 'auto', 'direct' or 'patch'")
        # worker threads of the host decoder: the host's cores, shared among the ranks of this node (torchrun)
        ranks_here = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        cores = max(1, (os.cpu_count() or 1) // ranks_here)
        if mode == "auto":
            mode = "patch" if cores >= self.AUTO_MIN_CORES else "direct"
        self.mode = mode
        self.threads = int(threads) if threads else max(1, min(32, cores))
        self._lib = env._be.lib
        if mode == "direct":
            # one pinned host array + the device-side mirror of what it holds, per step in flight
            self._obs_h = [torch.zeros((n, 84, 84, 3), dtype=torch.uint8).pin_memory() for _ in range(self.depth)]
            self._prevs = [torch.zeros((n, 84, 84, 3), dtype=torch.uint8, device=env.device) for _ in range(self.depth)]
            self._counter = torch.zeros(1, dtype=torch.int64, device=env.device)   # payload bytes written so far
            self.obs = self._obs_h[0]
            self._stage = []
        else:
            self.obs = torch.zeros((n, 84, 84, 3), dtype=torch.uint8)            # the caller-facing host array
            self._prev = torch.zeros((n, 84, 84, 3), dtype=torch.uint8, device=env.device)  # what the host holds, on the device
            stage_bytes = int(self._lib.mg_obs_delta_stage_bytes(n))
            self._stage = [torch.zeros(stage_bytes, dtype=torch.uint8).pin_memory() for _ in range(self.depth)]
            self._prev_ptr, self._obs_host_ptr = self._prev.data_ptr(), self.obs.data_ptr()
        self._res = [torch.zeros(9 * n, dtype=torch.uint8).pin_memory() for _ in range(self.depth)]
        self._act = [torch.zeros_like(env._actions_dev, device="cpu").pin_memory() for _ in range(self.depth)]
        self._ev = [torch.cuda.Event() for _ in range(self.depth)]
        self._views = [(r[:8 * n].view(torch.float64), r[8 * n:].view(torch.bool)) for r in self._res]
        self._truncated = torch.zeros(n, dtype=torch.bool)
        self._head = self._tail = 0
        self._bytes = C.c_uint64(0)
        self._bytes_ref = C.byref(self._bytes)
        self._res_ptrs = [(r.data_ptr(), r.data_ptr() + 8 * n) for r in self._res]
        self._stage_ptrs = [t.data_ptr() for t in self._stage]
        self._obs_dev_ptr = env.obs.data_ptr()
        self._pinned_ok = {}   # data_ptr -> numel of action tensors already validated (pinned, int64, contiguous)
        self._n_act = env._actions_dev.numel()
        self.last_d2h_bytes = 0   # patch mode: bytes that crossed the link for the step collected last
        self._total_d2h_bytes = 0
        self._steps = 0
        self.resync()

    @property
    def total_d2h_bytes(self):
        """Bytes that have crossed the link device -> host so far (changed chunks, headers in patch mode, results).
        Direct mode reads a device-side counter: this synchronises."""
        if self.mode == "direct":
            return int(self._counter.item()) + 9 * self.env.num_envs * self._steps
        return self._total_d2h_bytes

    def _encode(self, slot, full, stream=None):
        env = self.env
        cur = torch.cuda.current_stream(env.device)
        st = cur.cuda_stream if stream is None else stream
        if self.mode == "direct":
            rc = self._lib.mg_obs_delta_scatter(self._obs_dev_ptr, self._prevs[slot].data_ptr(), env.num_envs,
                                                self._obs_h[slot].data_ptr(), int(full), self._counter.data_ptr(), st)
        else:
            rc = self._lib.mg_obs_delta_encode(self._obs_dev_ptr, self._prev_ptr, env.num_envs, self._stage_ptrs[slot], int(full), st)
        if rc:
            L.check(rc, "mg_obs_delta_scatter" if self.mode == "direct" else "mg_obs_delta_encode")
        self._ev[slot].record(cur)

    def _decode(self, slot):
        self._ev[slot].synchronize()
        if self.mode == "direct":
            self.obs = self._obs_h[slot]   # (already up to date: the kernel wrote the changed chunks in place)
            return
        rc = self._lib.mg_obs_delta_decode(self._stage_ptrs[slot], self._obs_host_ptr, self.env.num_envs, self.threads,
                                           self._bytes_ref)
        if rc:
            L.check(rc, "mg_obs_delta_decode")
        self.last_d2h_bytes = self._bytes.value + 9 * self.env.num_envs
        self._total_d2h_bytes += self.last_d2h_bytes

    def resync(self):
        """Send the env's current observation as whole frames (after ``env.reset()``, or whenever ``env.obs`` was
        produced outside this pipe)."""
        if self._head != self._tail:
            raise L.MemGymError("resync() with steps still in flight: collect() them first")
        if self.mode == "direct":
            for slot in range(self.depth):   # every host array in the rotation holds the current frame afterwards
                self._encode(slot, True)
            for slot in range(self.depth):
                self._ev[slot].synchronize()
            self.obs = self._obs_h[(self._tail - 1) % self.depth]
            return self.obs
        self._encode(0, True)
        self._decode(0)
        return self.obs

    def reset(self, seed=None, options=None):
        """``env.reset`` + the reset frames into the host array. Returns (obs, info) like the env."""
        if self._head != self._tail:
            raise L.MemGymError("reset() with steps still in flight: collect() them first")
        _, info = self.env.reset(seed=seed, options=options)
        self.resync()
        return self.obs, info

    def submit(self, actions):
        """Enqueue one env step with host-side ``actions`` (int64; pinned memory is read in place)."""
        if self._head - self._tail >= self.depth:
            raise L.MemGymError(f"more than depth={self.depth} steps in flight: collect() first")
        slot = self._head % self.depth
        a = actions
        ptr = a.data_ptr() if isinstance(a, torch.Tensor) else 0
        if self._pinned_ok.get(ptr) != self._n_act:
            if isinstance(a, torch.Tensor) and a.device.type == "cpu" and a.dtype == torch.int64 and a.is_contiguous() \
                    and a.numel() == self._n_act and a.is_pinned():
                if len(self._pinned_ok) > 4096:
                    self._pinned_ok.clear()
                self._pinned_ok[ptr] = self._n_act   # (pinned memory comes from its own allocator: the address stays pinned)
            else:
                self._act[slot].copy_(torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a)
                                      .reshape(self._act[slot].shape))
                a = self._act[slot]
                ptr = a.data_ptr()
        env = self.env
        if env._params is None:
            raise L.MemGymError("submit() called before reset()")
        stream = torch.cuda.current_stream(env.device).cuda_stream
        rp, dp = self._res_ptrs[slot]
        env._launch_step_ptrs(ptr, rp, dp, stream)
        self._keep_actions = a  # (alive until the next submit: the kernel reads it in place)
        self._encode(slot, False, stream)
        self._head += 1

    def collect(self):
        """Wait for the oldest step in flight and bring the host arrays up to date with it. Returns
        (obs, reward, done, truncated, info) with host tensors; ``info`` is the env's device-side episode matrix.
        Direct mode: ``obs`` is this step's own host array (``depth`` arrays are used in turn)."""
        if self._head == self._tail:
            raise L.MemGymError("collect() without a submitted step")
        slot = self._tail % self.depth
        self._decode(slot)
        self._tail += 1
        self._steps += 1
        reward, done = self._views[slot]
        return self.obs, reward, done, self._truncated, {"episode": self.env.info, "_episode": done}

    def step(self, actions):
        self.submit(actions)
        return self.collect()


class MemoryGymEnv:
    """Single-instance view with the reference's exact call protocol (step until done, then reset):
    numpy observations, Python float rewards, dict infos. Thin wrapper over ``MemoryGymVecEnv(num_envs=1,
    auto_reset=False)`` for drop-in use and for the parity tests; throughput comes from the batched class."""

    def __init__(self, env_id, device="cuda", render_mode=None, _backend=None):
        self.vec = MemoryGymVecEnv(env_id, 1, device=device, auto_reset=False, render_mode=render_mode, _backend=_backend)
        self.render_mode = render_mode
        self.action_space, self.observation_space = self.vec.action_space, self.vec.observation_space
        self.default_reset_parameters = self.vec.default_reset_parameters
        self.has_ground_truth_info = self.vec.has_ground_truth_info
        if self.has_ground_truth_info:
            self.ground_truth_space = self.vec.ground_truth_space
        self.metadata = self.vec.metadata

    @property
    def max_episode_steps(self):
        return self.vec.max_episode_steps

    def process_reset_params(self, reset_params):
        return self.vec.process_reset_params(reset_params)

    def reset(self, seed=None, return_info=True, options=None):
        obs, info = self.vec.reset(seed=seed, options=options)
        out = {}
        if self.has_ground_truth_info:
            out["ground_truth"] = info["ground_truth"][0].cpu().numpy()
        return self._np_obs(obs), out

    @staticmethod
    def _np_obs(obs):
        if isinstance(obs, dict):
            return {k: v[0].cpu().numpy() for k, v in obs.items()}
        return obs[0].cpu().numpy()

    def step(self, action):
        a = np.asarray(action, dtype=np.int64).reshape(self.vec._actions_dev.shape)
        obs, reward, done, trunc, info = self.vec.step(a)
        d = bool(done[0].item())
        out = {}
        if d:
            row = info["episode"][0].cpu().numpy()
            for k, col in self.vec.spec_.info_keys.items():
                out[k] = float(row[col])
            out["length"] = int(out["length"])
            if self.has_ground_truth_info:
                out["ground_truth"] = row[L.MG_INFO_GT0:L.MG_INFO_GT0 + self.vec.spec_.gt_dim].copy()
        elif self.has_ground_truth_info:
            out["ground_truth"] = info["ground_truth"][0].cpu().numpy()
        return self._np_obs(obs), float(reward[0].item()), d, False, out

    def render(self):
        return self.vec.render()[0].cpu().numpy()

    def close(self):
        self.vec.close()
