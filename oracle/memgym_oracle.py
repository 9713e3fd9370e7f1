"""CPU oracle: a restatement of the reference Memory-Gym environments (TEST INFRASTRUCTURE ONLY).

Not imported by the product package. Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and only as the checker /
CPU baseline. It travels to the GPU box (``/root/reference`` does not), which is why the env logic
is restated here instead of importing the reference.

PARITY STATUS
* env logic (state machines, RNG draw order, quirks): pinned — ``tests/golden/ref_*.npz`` hold
  trajectories produced by the UNMODIFIED reference sources imported from /root/reference on top
  of ``oracle/shims`` (``oracle/gen_golden.py``); ``tests/test_oracle_vs_reference_golden.py``
  replays them through this module bit-for-bit.
* raster: pinned against real-PyGame GIF known answers (``tests/test_raster_kat.py``) except the
  pygame>=2.1.3 thick circle of the "stay" icon (4 px) — "parity unpinned" for those pixels.
* RNG: the real ``numpy.random.Generator(PCG64(SeedSequence(seed)))`` the reference gets from
  gymnasium 0.29 (``gymnasium/utils/seeding.py``), so draws are the reference's by construction.
* ``info["reward"]`` is the left-to-right IEEE-double sum of the step rewards, i.e. what ``sum()``
  returns on the Python 3.11 the reference documents (README.md:63); Python >= 3.12's ``sum`` is
  compensated and can differ in the last bit.

Each class cites the reference lines it follows. Frames are ``uint8[84][84][3]`` in pygame's
``surfarray.array3d`` layout ``[x][y][c]``.
"""
from __future__ import annotations

import math

import numpy as np

from . import assets
from . import sdl_raster as R

SCALE = 0.25
SCREEN = int(336 * SCALE)


def _frame_to_obs(px: np.ndarray) -> np.ndarray:
    t = px.T
    out = np.empty((px.shape[1], px.shape[0], 3), dtype=np.uint8)
    out[..., 0] = t >> 16
    out[..., 1] = t >> 8
    out[..., 2] = t
    return out


def _paste(frame: np.ndarray, src: R.Surface, x: int, y: int) -> None:
    """Colour-key blit of ``src`` with its top-left at (x, y), clipped (SDL_BlitSurface)."""
    x0, y0 = max(x, 0), max(y, 0)
    x1, y1 = min(x + src.w, frame.shape[1]), min(y + src.h, frame.shape[0])
    if x1 <= x0 or y1 <= y0:
        return
    sv = src.px[y0 - y:y1 - y, x0 - x:x1 - x]
    if src.colorkey is None:
        frame[y0:y1, x0:x1] = sv
    else:
        np.copyto(frame[y0:y1, x0:x1], sv, where=sv != src.colorkey)


class _OracleEnv:
    """Seeding contract of gymnasium 0.29 ``Env.reset`` (seed given -> new PCG64; else continue)."""

    render_mode = None  # "debug_rgb_array": render() returns the ground-truth composite (_build_debug_surface)

    def __init__(self):
        self._rng = None

    def render(self):
        """``render()`` of the reference in ``debug_rgb_array`` mode (e.g. mortar_mayhem.py:397-399): the 84x84 debug
        surface scaled to 336x336 (``transform.scale`` = exact 4x pixel replication), returned as [y][x][c]. Like the
        reference's, a call may consume state (the clone of the command queue): call it once per reset() / step()."""
        assert self.render_mode == "debug_rgb_array"
        obs = _frame_to_obs(self._debug_surface())           # [x][y][c] at 84x84
        return np.repeat(np.repeat(obs.transpose(1, 0, 2), 4, 0), 4, 1)

    def _seed(self, seed):
        if seed is not None:
            self._rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        elif self._rng is None:
            self._rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence()))
        return self._rng

    @classmethod
    def process_reset_params(cls, options):
        params = dict(cls.default_reset_parameters)
        if options is not None:
            for k, v in options.items():
                assert k in params, "Provided reset parameter (" + str(k) + ") is not valid. Check spelling."
                params[k] = v
        cls._check_params(params)
        return params

    @classmethod
    def _check_params(cls, params):
        pass


# --------------------------------------------------------------------------------------------------
# kinematics  (character_controller.py)
# --------------------------------------------------------------------------------------------------
def _velocity_and_rotation(action, rotation):
    """character_controller.py:100-121 — unit direction and the new rotation (degrees)."""
    vx = -1 if action[0] == 1 else (1 if action[0] == 2 else 0)
    vy = -1 if action[1] == 1 else (1 if action[1] == 2 else 0)
    if action[0] == 1:
        rotation = 90
    if action[0] == 2:
        rotation = 270
    if action[1] == 1:
        rotation = 0
    if action[1] == 2:
        rotation = 180
    if vx and vy:
        rotation = {(-1, -1): 45, (-1, 1): 135, (1, -1): 315, (1, 1): 225}[(vx, vy)]
    return vx, vy, rotation


def move_clamped(cx, cy, rotation, action, speed, radius, rect):
    """CharacterController.step (character_controller.py:89-146). ``rect`` = (x, y, w, h)."""
    vx, vy, rotation = _velocity_and_rotation(action, rotation)
    if vx or vy:
        n = math.sqrt(float(vx * vx + vy * vy))
        # normalize() * speed, then int() per component (:124-126)
        cx += int((vx / n) * speed)
        cy += int((vy / n) * speed)
    left, top, right, bottom = rect[0], rect[1], rect[0] + rect[2], rect[1] + rect[3]
    cx = min(cx, right - radius)
    cx = max(cx, left + radius)
    cy = min(cy, bottom - radius)
    cy = max(cy, top + radius)
    return cx, cy, rotation


def move_free(cx, cy, rotation, action, speed):
    """CharacterController.step without a boundary rect (endless_mystery_path.py:299)."""
    vx, vy, rotation = _velocity_and_rotation(action, rotation)
    if vx or vy:
        n = math.sqrt(float(vx * vx + vy * vy))
        cx += int((vx / n) * speed)
        cy += int((vy / n) * speed)
    return cx, cy, rotation


def move_wrapped(cx, cy, rotation, action, speed, radius, rect):
    """ScreenWrapCharacterController.step (character_controller.py:226-283): float velocity, the
    Rect assignment truncates toward zero, then wrap at +-radius/2 beyond the rect."""
    vx, vy, rotation = _velocity_and_rotation(action, rotation)
    fx, fy = float(vx), float(vy)
    if vx or vy:
        n = math.sqrt(fx * fx + fy * fy)
        fx, fy = (fx / n) * speed, (fy / n) * speed
    cx, cy = int(float(cx) + fx), int(float(cy) + fy)
    left, top, right, bottom = rect[0], rect[1], rect[0] + rect[2], rect[1] + rect[3]
    off = radius * 0.5
    x, y = cx, cy
    if x > right + off:
        x = left - off
    if x < left - off:
        x = right + off
    if y > bottom + off:
        y = top - off
    if y < top - off:
        y = bottom + off
    return int(x), int(y), rotation


# --------------------------------------------------------------------------------------------------
# Mortar Mayhem family
# --------------------------------------------------------------------------------------------------
TILE = 56 * SCALE  # 14.0


def calc_max_episode_steps(command_count, show_duration, show_delay, execution_duration, execution_delay):
    """pygame_assets.py:420-436 (informational only; the caller passes delay/duration swapped)."""
    clue = (show_duration + show_delay) * command_count
    act = (execution_duration + execution_delay) * command_count
    return clue + act - execution_delay + 1


class MortarMayhemOracle(_OracleEnv):
    """mortar_mayhem.py:15-369 (MortarMayhemEnv) on CharacterController + MortarArena + Command."""

    env_id = "MortarMayhem-v0"
    grid_locomotion = False
    task_b = False        # MortarMayhemB: no command visualisation, commands are given as a one-hot vector
    spawn_offset = False  # MortarMayhemB: spawn jitter like the endless variant
    default_reset_parameters = {
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

    @classmethod
    def _check_params(cls, p):
        assert p["allowed_commands"] >= 4 and p["allowed_commands"] <= 9
        assert p["arena_size"] >= 2 and p["arena_size"] <= 6

    # ---- geometry --------------------------------------------------------------------------------
    def _setup_arena(self, size):
        self.size = size
        dim = int(TILE * size)
        self.origin = SCREEN // 2 - (dim >> 1)  # arena.rect.center = (42, 42)
        self.arena_rect = (self.origin, self.origin, dim, dim)
        self.tiles_on = False

    def _tile_centre(self, i, j):
        # get_tile_global_position + translate (mortar_mayhem.py:238-242)
        t = SCREEN // 2 - ((int(TILE * self.size)) >> 1) + TILE // 2
        return TILE * i + t, TILE * j + t

    def _normalize(self, cx, cy):
        return ((cx - self.origin) // TILE, (cy - self.origin) // TILE)

    def _agent_scale(self):
        return self.p["agent_scale"]

    # ---- reset -----------------------------------------------------------------------------------
    def reset(self, seed=None, options=None):
        rng = self._seed(seed)
        p = self.p = self.process_reset_params(options)
        self.max_episode_steps = calc_max_episode_steps(
            max(p["command_count"]), max(p.get("command_show_duration", [0])), max(p.get("command_show_delay", [0])),
            max(p["explosion_delay"]), max(p["explosion_duration"]))
        self.rewards = []
        self._setup_arena(p["arena_size"])
        self.sprites = assets.agent_sprites(self._agent_scale())
        self.radius = int(25 * self._agent_scale())
        tile_id = int(rng.integers(0, self.size ** 2))
        i, j = tile_id // self.size, tile_id % self.size
        x, y = self._tile_centre(i, j)
        if self.spawn_offset:  # mortar_mayhem_b.py:167-170
            off = rng.integers(-8 * SCALE, 8 * SCALE, 2)
            x, y = x + int(off[0]), y + int(off[1])
        self.cx, self.cy = int(x), int(y)
        self.rotation = 0
        self.pos = self._normalize(self.cx, self.cy)
        if self.grid_locomotion:
            self.gx, self.gy = int(self.pos[0]), int(self.pos[1])
        # command sequence (:163-183)
        self.n = int(rng.choice(p["command_count"]))
        self.commands = []
        sim = self.pos
        for _ in range(self.n):
            valid = []
            for k in range(p["allowed_commands"]):
                d = assets.COMMAND_DELTAS[k]
                tx, ty = sim[0] + d[0], sim[1] + d[1]
                if 0 <= tx < self.size and 0 <= ty < self.size:
                    valid.append(k)
            k = valid[int(rng.integers(0, len(valid)))]
            self.commands.append(k)
            sim = (sim[0] + assets.COMMAND_DELTAS[k][0], sim[1] + assets.COMMAND_DELTAS[k][1])
        if self.task_b:
            self.show, first = [], None
        else:
            dur = int(rng.choice(p["command_show_duration"]))
            delay = int(rng.choice(p["command_show_delay"]))
            self.show = self._show_list(self.commands, dur, delay)
            self.show_clone = list(self.show)  # "the clone is needed for render()" (:251)
            first = self.show.pop(0)
        d0 = assets.COMMAND_DELTAS[self.commands[0]]
        self.target = (self.pos[0] + d0[0], self.pos[1] + d0[1])
        self.cc = self.cs = self.cv = 0
        self.E = int(rng.choice(p["explosion_duration"]))
        self.D = int(rng.choice(p["explosion_delay"]))
        self.draw_cx, self.draw_cy, self.draw_rot = self.cx, self.cy, 0
        return self._obs(self._render(first, (self.cx, self.cy, 0))), {}

    def _obs(self, frame):
        if not self.task_b:
            return frame
        # _encode_commands_one_hot (mortar_mayhem_b.py:102-131): 20 x 9, slot order stay, right, left, up, down, ...
        slot = {"stay": 0, "right": 1, "left": 2, "up": 3, "down": 4, "right_down": 5, "right_up": 6, "left_down": 7,
                "left_up": 8}
        vec = np.zeros(20 * 9, dtype=np.float32)
        for c, k in enumerate(self.commands):
            vec[9 * c + slot[assets.COMMAND_NAMES[k]]] = 1.0
        return {"visual_observation": frame, "vector_observation": vec}

    @staticmethod
    def _show_list(commands, dur, delay):
        out = []
        for k in commands:
            out += [k] * dur + [-1] * delay
        return out

    # ---- step ------------------------------------------------------------------------------------
    def _move(self, action):
        if self.grid_locomotion:
            # GridCharacterController.step (character_controller.py:177-210)
            if action == 1:
                self.rotation = (self.rotation + 90) % 360
            if action == 2:
                self.rotation = (self.rotation - 90) % 360
            if action == 3:
                face = self.rotation // 90
                last = self.size - 1
                if face == 0 and self.gy > 0:
                    self.gy -= 1
                elif face == 3 and self.gx < last:
                    self.gx += 1
                elif face == 2 and self.gy < last:
                    self.gy += 1
                elif face == 1 and self.gx > 0:
                    self.gx -= 1
                x, y = self._tile_centre(self.gx, self.gy)
                self.cx, self.cy = int(x), int(y)
        else:
            self.cx, self.cy, self.rotation = move_clamped(
                self.cx, self.cy, self.rotation, action, self.p["agent_speed"], self.radius, self.arena_rect)

    def step(self, action):
        p = self.p
        reward = 0
        done = False
        success = 0
        icon = None
        if self.show:
            icon = self.show.pop(0)
            sprite = (self.cx, self.cy, 0)
        else:
            self._move(action)
            sprite = (self.cx, self.cy, self.rotation)
            self.pos = self._normalize(self.cx, self.cy)
            verify = self.cs % self.D == 0 and self.cs > 0
            if verify and not self.tiles_on:
                if self.cc < self.n:
                    self.cc += 1
                    self.tiles_on = True
                    if self.pos == self.target:
                        reward += p["reward_command_success"]
                    else:
                        done = True
                        reward += p["reward_command_failure"]
                if self.cc >= self.n:
                    done = True
                    success = 1
                    reward += p["reward_episode_success"]
                self.cs = 1
            if self.tiles_on:
                if self.cv % self.E == 0 and self.cv > 0:
                    self.tiles_on = False
                    self.cv = 0
                    if self.cc < self.n:
                        d = assets.COMMAND_DELTAS[self.commands[self.cc]]
                        self.target = (self.target[0] + d[0], self.target[1] + d[1])
                else:
                    if not self.pos == self.target:
                        done = True
                        reward = p["reward_command_failure"]
                    self.cv += 1
            else:
                self.cs += 1
        self.rewards.append(reward)
        info = {}
        if done:
            info = {"reward": _seq_sum(self.rewards), "length": len(self.rewards), "success": success,
                    "commands_completed": (self.cc - 1 + success) / self.n}
        self.handle = sprite  # rotated_agent_surface / rect (:291, :295): survives reset(), see _debug_surface
        return self._obs(self._render(icon, sprite)), reward, done, False, info

    # ---- frame -----------------------------------------------------------------------------------
    handle = None      # pose of rotated_agent_surface/rect; None until the first step of the instance (:91)
    show_clone = ()    # _command_visualization_clone

    def _debug_surface(self):
        """_build_debug_surface (mortar_mayhem.py:105-136; same code in mortar_mayhem_grid.py:104-135 and
        endless_mortar_mayhem.py:114-145): arena, the sprite handle of the last step (it is not refreshed by reset():
        right after a reset the agent is still drawn at the previous episode's last pose), the next entry of the
        CLONED command queue while the real queue is non-empty, and a green ring (r = tile_dim // 2, width int(8 *
        SCALE)) on the target tile."""
        icon = None
        if self.show:
            icon = self.show_clone.pop(0)
        sprite = self.handle if self.handle is not None else (self.cx, self.cy, 0)
        surf = R.Surface(SCREEN, SCREEN)
        surf.px[...] = self._compose(icon, sprite)
        tx, ty = self._tile_centre(int(self.target[0]), int(self.target[1]))
        R.circle(surf, (0, 255, 0), (tx, ty), TILE // 2, int(8 * SCALE))
        return surf.px

    def _render(self, icon, sprite):
        return _frame_to_obs(self._compose(icon, sprite))

    def _compose(self, icon, sprite):
        frame = np.zeros((SCREEN, SCREEN), dtype=np.uint32)
        blue, red = assets.mortar_tile(SCALE, False), assets.mortar_tile(SCALE, True)
        redden = self.tiles_on and self.p["visual_feedback"]
        t = int(TILE)
        for i in range(self.size):
            for j in range(self.size):
                tile = red if (redden and not self.target == (i, j)) else blue
                frame[self.origin + t * j:self.origin + t * (j + 1), self.origin + t * i:self.origin + t * (i + 1)] = tile.px
        cx, cy, rot = sprite
        sp = self.sprites[rot // 45]
        _paste(frame, sp, cx - (sp.w >> 1), cy - (sp.h >> 1))
        if icon is not None:
            surf = assets.command_icon("" if icon < 0 else assets.COMMAND_NAMES[icon], SCALE)
            pos = assets.command_blit_pos(SCALE, SCREEN)
            _paste(frame, surf, int(pos[0]), int(pos[1]))
        return frame


def _seq_sum(values):
    total = 0
    for v in values:
        total = total + v
    return total


class GridMortarMayhemOracle(MortarMayhemOracle):
    """mortar_mayhem_grid.py:15-375 (GridMortarMayhemEnv) on GridCharacterController."""

    env_id = "MortarMayhem-Grid-v0"
    grid_locomotion = True
    default_reset_parameters = {
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

    def _agent_scale(self):
        return SCALE  # GridCharacterController(SCALE, ...) ignores agent_scale (mortar_mayhem_grid.py:249)


class MortarMayhemBOracle(MortarMayhemOracle):
    """mortar_mayhem_b.py:15-290 (MortarMayhemTaskBEnv): act-only variant, commands as a one-hot vector."""

    env_id = "MortarMayhemB-v0"
    task_b = True
    spawn_offset = True
    default_reset_parameters = {
        "agent_scale": 1.0 * SCALE,
        "agent_speed": 12.0 * SCALE,
        "arena_size": 5,
        "allowed_commands": 9,
        "command_count": [10],
        "explosion_duration": [6],
        "explosion_delay": [18],
        "visual_feedback": True,
        "reward_command_failure": 0.0,
        "reward_command_success": 0.1,
        "reward_episode_success": 0.0,
    }

    @classmethod
    def _check_params(cls, p):
        assert p["allowed_commands"] >= 4 and p["allowed_commands"] <= 9
        assert p["arena_size"] >= 2 and p["arena_size"] <= 6
        assert max(p["command_count"]) <= 20, "20 commands are allowed at maximum"


class GridMortarMayhemBOracle(MortarMayhemBOracle):
    """mortar_mayhem_b_grid.py:15-290 (GridMortarMayhemTaskBEnv)."""

    env_id = "MortarMayhemB-Grid-v0"
    grid_locomotion = True
    spawn_offset = False
    default_reset_parameters = {
        "agent_scale": 1.0 * SCALE,
        "arena_size": 5,
        "allowed_commands": 5,
        "command_count": [10],
        "explosion_duration": [2],
        "explosion_delay": [6],
        "visual_feedback": True,
        "reward_command_failure": 0.0,
        "reward_command_success": 0.1,
        "reward_episode_success": 0.0,
    }

    def _agent_scale(self):
        return SCALE


class EndlessMortarMayhemOracle(MortarMayhemOracle):
    """endless_mortar_mayhem.py:15-373 (EndlessMortarMayhemEnv) on ScreenWrapCharacterController."""

    env_id = "Endless-MortarMayhem-v0"
    default_reset_parameters = {
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

    def __init__(self):
        super().__init__()
        # pose of the sprite handle (rotated_agent_surface/rect) last returned by a controller. It is
        # never cleared by reset() (:99, :277-278), and after reset it keeps pointing at the OLD
        # controller's rect, i.e. it is frozen at the previous episode's last pose (quirk A.9#3).
        self.stale = None

    @classmethod
    def _check_params(cls, p):
        assert p["allowed_commands"] >= 4 and p["allowed_commands"] <= 9

    def _draw_commands(self, rng, count):
        return [int(v) for v in rng.integers(0, self.p["allowed_commands"], count)]

    def _gt(self):
        return np.asarray([*self.target]) / 5.0

    def reset(self, seed=None, options=None):
        rng = self._seed(seed)
        p = self.p = self.process_reset_params(options)
        self.t = 0
        self.max_episode_steps = p["max_steps"]
        self.rewards = []
        self._setup_arena(6)
        self.sprites = assets.agent_sprites(p["agent_scale"])
        self.radius = int(25 * p["agent_scale"])
        tile_id = int(rng.integers(0, 36))
        off = rng.integers(-8 * SCALE, 8 * SCALE, 2)
        x, y = self._tile_centre(tile_id // 6, tile_id % 6)
        self.cx, self.cy = int(x + int(off[0])), int(y + int(off[1]))
        self.rotation = 0
        self.pos = self._normalize(self.cx, self.cy)
        self.n = p["initial_command_count"]
        self.commands = self._draw_commands(rng, self.n)
        self.dur = int(rng.choice(p["command_show_duration"]))
        self.delay = int(rng.choice(p["command_show_delay"]))
        self.show = self._show_list(self.commands, self.dur, self.delay)
        self.show_clone = list(self.show)
        first = self.show.pop(0)
        d0 = assets.COMMAND_DELTAS[self.commands[0]]
        self.target = ((self.pos[0] + d0[0]) % 6, (self.pos[1] + d0[1]) % 6)
        self.cc = self.cs = self.cv = 0
        self.total_completed = 0
        self.E = int(rng.choice(p["explosion_duration"]))
        self.D = int(rng.choice(p["explosion_delay"]))
        return self._render(first, (self.cx, self.cy, 0)), {"ground_truth": self._gt()}

    def step(self, action):
        p = self.p
        reward = 0
        done = False
        icon = None
        if self.show:
            icon = self.show.pop(0)
            if self.stale is None:
                # first episode of this instance: get_rotated_sprite(0) of the live controller
                self.stale = (self.cx, self.cy, 0)
        else:
            self.cx, self.cy, self.rotation = move_wrapped(
                self.cx, self.cy, self.rotation, action, p["agent_speed"], self.radius, self.arena_rect)
            self.stale = (self.cx, self.cy, self.rotation)
            self.pos = self._normalize(self.cx, self.cy)
            verify = self.cs % self.D == 0 and self.cs > 0
            if verify and not self.tiles_on:
                if self.cc < self.n:
                    self.cc += 1
                    self.tiles_on = True
                    if self.pos == self.target:
                        reward += p["reward_command_success"]
                        self.total_completed += 1
                        if self.cc == self.n:
                            reward += p["reward_new_command_success"]
                    else:
                        done = True
                        reward += p["reward_command_failure"]
                if self.cc >= self.n:
                    new = self._draw_commands(self._rng, 1)
                    self.commands.append(new[0])
                    self.n = len(self.commands)
                    self.cc = self.cs = self.cv = 0
                    self.show = self._show_list(new, self.dur, self.delay)
                    self.show_clone = list(self.show)  # re-cloned (:321): render() now runs one entry AHEAD of step()
                self.cs = 1
            if self.tiles_on:
                if self.cv % self.E == 0 and self.cv > 0:
                    self.tiles_on = False
                    self.cv = 0
                    if self.cc < self.n:
                        d = assets.COMMAND_DELTAS[self.commands[self.cc]]
                        self.target = ((self.target[0] + d[0]) % 6, (self.target[1] + d[1]) % 6)
                else:
                    if not self.pos == self.target:
                        done = True
                        reward = p["reward_command_failure"]
                    self.cv += 1
            else:
                self.cs += 1
        self.t += 1
        if self.t == self.max_episode_steps:
            done = True
        self.rewards.append(reward)
        info = {"ground_truth": self._gt()}
        if done:
            info = {"reward": _seq_sum(self.rewards), "length": len(self.rewards),
                    "commands_completed": self.total_completed,
                    "max_command_sequence": len(self.commands) - 1 if len(self.commands) > 1 else 0,
                    "ground_truth": self._gt()}
        # (within an episode the agent cannot move during a show phase, so the handle's pose equals
        # the live pose; only across a reset does it lag behind)
        self.handle = self.stale
        return self._render(icon, self.stale), reward, done, False, info


# --------------------------------------------------------------------------------------------------
# Mystery Path
# --------------------------------------------------------------------------------------------------
def generate_mystery_path(rng, n, start, end):
    """pygame_assets.py:606-724 (MysteryPath.__init__): walls + noisy A*. Returns [end, ..., start]."""
    wall = [[False] * n for _ in range(n)]
    for i in range(n):
        for j in range(n):
            if 0 < i < n - 2 and 0 < j < n - 2:
                if rng.integers(0, 100) < 33:
                    wall[i][j] = True

    def neighbours(i, j):  # Node.add_neighbors order (:465-472)
        out = []
        if i < n - 1:
            out.append((i + 1, j))
        if i > 0:
            out.append((i - 1, j))
        if j < n - 1:
            out.append((i, j + 1))
        if j > 0:
            out.append((i, j - 1))
        return out

    def diagonals(i, j):  # (:474-481)
        out = []
        if i < n - 1 and j < n - 1:
            out.append((i + 1, j + 1))
        if i > 0 and j > 0:
            out.append((i - 1, j - 1))
        if i < n - 1 and j > 0:
            out.append((i + 1, j - 1))
        if i > 0 and j < n - 1:
            out.append((i - 1, j + 1))
        return out

    near_ends = set(neighbours(*start)) | set(neighbours(*end))
    ring = []
    for i in range(n):
        for j in range(n):
            if i == 0 or i == n - 1 or j == 0 or j == n - 1:
                if (i, j) == tuple(start) or (i, j) == tuple(end) or (i, j) in near_ends:
                    continue
                if any(wall[a][b] for a, b in neighbours(i, j)) or any(wall[a][b] for a, b in diagonals(i, j)):
                    continue
                ring.append((i, j))
    for _ in range(int(rng.choice([4, 8]))):
        if ring:
            node = ring[int(rng.integers(0, len(ring)))]  # rng.choice(outer_nodes)
            wall[node[0]][node[1]] = True
            ring.remove(node)

    start, end = tuple(start), tuple(end)
    g = {start: 0}
    f = {start: 0}
    prev = {start: None}
    open_set, closed = [start], set()
    while True:
        if not open_set:
            raise Exception("No valid path found")
        w = 0
        for k in range(len(open_set)):
            if f[open_set[k]] < f[open_set[0]]:
                w = k
                break
        cur = open_set[w]
        if cur == end:
            path = [end]
            while prev[path[-1]] is not None:
                path.append(prev[path[-1]])
            return path, wall
        open_set.remove(cur)
        closed.add(cur)
        for nb in neighbours(*cur):
            if nb in closed or wall[nb[0]][nb[1]]:
                continue
            cost = g[cur] + int(rng.integers(1, 9))
            new_path = False
            if nb in open_set:
                if cost < g[nb]:
                    new_path = True  # g_cost is NOT updated (reference bug :712)
            else:
                g[nb] = cost
                new_path = True
                open_set.append(nb)
            if new_path:
                h = math.sqrt((nb[0] - end[0]) ** 2 + abs(nb[1] - end[1]) ** 2)
                f[nb] = g[nb] + h
                prev[nb] = cur


class MysteryPathOracle(_OracleEnv):
    """mystery_path.py:15-276 (MysteryPathEnv)."""

    env_id = "MysteryPath-v0"
    default_reset_parameters = {
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
    GRID = 7
    TILE = SCREEN / 7  # 12.0

    def reset(self, seed=None, options=None):
        rng = self._seed(seed)
        p = self.p = self.process_reset_params(options)
        self.t = 0
        self.max_episode_steps = p["max_steps"]
        self.rewards = []
        g = self.GRID
        side = rng.choice(p["cardinal_origin_choice"])
        if side == 0:
            self.start = (0, int(rng.integers(0, g)))
            self.end = (g - 1, int(rng.integers(0, g)))
        elif side == 1:
            self.start = (g - 1, int(rng.integers(0, g)))
            self.end = (0, int(rng.integers(0, g)))
        elif side == 2:
            self.start = (int(rng.integers(0, g)), 0)
            self.end = (int(rng.integers(0, g)), g - 1)
        else:
            self.start = (int(rng.integers(0, g)), g - 1)
            self.end = (int(rng.integers(0, g)), 0)
        self.path, self.walls = generate_mystery_path(rng, g, self.start, self.end)
        self.visited = set()
        self.sprites = assets.agent_sprites(p["agent_scale"])
        self.radius = int(25 * p["agent_scale"])
        self.rotation = 0
        self.cx = int(self.start[0] * self.TILE + self.radius)
        self.cy = int(self.start[1] * self.TILE + self.radius)
        self.off_path = False
        self.num_fails = 0
        self.cross_alpha = 0
        self.cross_x = self.cross_y = 0  # fall_off_rect top-left
        self._after_reset()
        return self._render(cross=False), {}

    def _after_reset(self):
        pass

    def step(self, action):
        p = self.p
        reward = 0
        done = False
        success = 0
        screen = (0, 0, SCREEN, SCREEN)
        if not self.off_path:
            self.cx, self.cy, self.rotation = move_clamped(
                self.cx, self.cy, self.rotation, action, p["agent_speed"], self.radius, screen)
        else:
            self.cx = int(self.start[0] * self.TILE + self.radius)
            self.cy = int(self.start[1] * self.TILE + self.radius)
            self.cx, self.cy, self.rotation = move_clamped(
                self.cx, self.cy, self.rotation, (0, 0), p["agent_speed"], self.radius, screen)
        pos = (self.cx // self.TILE, self.cy // self.TILE)
        if pos == self.end:
            reward += p["reward_goal"]
            done = True
            success = 1
        else:
            on_path = False
            for node in self.path:
                if pos == node:
                    on_path = True
                    if node not in self.visited and node != self.start and node != self.end:
                        reward += p["reward_path_progress"]
                        self.visited.add(node)
                    break
            if not on_path:
                reward += p["reward_fall_off"]
                self.num_fails += 1
                if p["visual_feedback"]:
                    self.cross_alpha = 255
                self.off_path = True
            else:
                self.cross_alpha = 0
                self.off_path = False
            cross = assets.fall_off_cross(SCALE)
            self.cross_x, self.cross_y = self.cx - (cross.w >> 1), self.cy - (cross.h >> 1)
        reward += p["reward_step"]
        self.t += 1
        if self.t == self.max_episode_steps:
            done = True
        self.rewards.append(reward)
        info = {}
        if done:
            info = {"reward": _seq_sum(self.rewards), "length": len(self.rewards), "success": success,
                    "num_fails": self.num_fails}
        return self._render(cross=True), reward, done, False, info

    def _render(self, cross):
        return _frame_to_obs(self._compose(cross, False))

    def _debug_surface(self):
        """_build_debug_surface (mystery_path.py:103-117, mystery_path_grid.py:102-116): the whole path (goal green,
        origin blue, the rest white), the walls in red, the agent and the fall-off cross (its alpha is 0 after reset)."""
        return self._compose(True, True)

    def _compose(self, cross, debug):
        frame = np.zeros((SCREEN, SCREEN), dtype=np.uint32)
        t = int(self.TILE)
        # MysteryPath.draw_to_surface (pygame_assets.py:738-762): path[0] is the goal, path[-1] the origin
        last = len(self.path) - 1
        for k, (i, j) in enumerate(self.path):
            if k == 0 and (self.p["show_goal"] or debug):
                frame[j * t:(j + 1) * t, i * t:(i + 1) * t] = R.map_rgb((0, 255, 0))
            elif k == last and (self.p["show_origin"] or debug):
                frame[j * t:(j + 1) * t, i * t:(i + 1) * t] = R.map_rgb((0, 0, 255))
            elif 0 < k < last and debug:
                frame[j * t:(j + 1) * t, i * t:(i + 1) * t] = R.map_rgb((255, 255, 255))
        if debug:
            for i in range(self.GRID):
                for j in range(self.GRID):
                    if self.walls[i][j]:
                        frame[j * t:(j + 1) * t, i * t:(i + 1) * t] = R.map_rgb((255, 0, 0))
        sp = self.sprites[self.rotation // 45]
        _paste(frame, sp, self.cx - (sp.w >> 1), self.cy - (sp.h >> 1))
        if cross and self.cross_alpha == 255:
            _paste(frame, assets.fall_off_cross(SCALE), self.cross_x, self.cross_y)
        return frame


class GridMysteryPathOracle(MysteryPathOracle):
    """mystery_path_grid.py:15-280 (GridMysteryPathEnv): grid locomotion on the 7x7 path grid."""

    env_id = "MysteryPath-Grid-v0"
    default_reset_parameters = {
        "max_steps": 128,
        "agent_scale": 1.0 * SCALE,
        "cardinal_origin_choice": [0, 1, 2, 3],
        "show_origin": False,
        "show_goal": False,
        "visual_feedback": True,
        "reward_goal": 1.0,
        "reward_fall_off": 0.0,
        "reward_path_progress": 0.0,
        "reward_step": 0.0,
    }

    def process_reset_params(self, options):  # instance-level: adds the two keys the base class reads
        p = dict(type(self).default_reset_parameters)
        if options is not None:
            for k, v in options.items():
                assert k in p, "Provided reset parameter (" + str(k) + ") is not valid. Check spelling."
                p[k] = v
        p = dict(p)
        p["_agent_scale_param"] = p["agent_scale"]
        p["agent_scale"] = SCALE      # GridCharacterController(SCALE, ...) ignores agent_scale (:188)
        p["agent_speed"] = 0.0
        return p

    def _after_reset(self):
        self.gx, self.gy = self.start  # normalize(start * 12 + int(25 * SCALE))
        self.cx, self.cy = 12 * self.gx + 6, 12 * self.gy + 6

    def step(self, action):
        p = self.p
        reward = 0
        done = False
        success = 0
        if self.off_path:  # reset_position(start) then a no-op controller step (:218-221)
            self.gx, self.gy = self.start
            action = 0
        if action == 1:
            self.rotation = (self.rotation + 90) % 360
        if action == 2:
            self.rotation = (self.rotation - 90) % 360
        if action == 3:
            face = self.rotation // 90
            if face == 0 and self.gy > 0:
                self.gy -= 1
            elif face == 3 and self.gx < self.GRID - 1:
                self.gx += 1
            elif face == 2 and self.gy < self.GRID - 1:
                self.gy += 1
            elif face == 1 and self.gx > 0:
                self.gx -= 1
        self.cx, self.cy = 12 * self.gx + 6, 12 * self.gy + 6
        pos = (self.gx, self.gy)
        if pos == self.end:
            reward += p["reward_goal"]
            done = True
            success = 1
        else:
            on_path = False
            for node in self.path:
                if pos == node:
                    on_path = True
                    if node not in self.visited and node != self.start and node != self.end:
                        reward += p["reward_path_progress"]
                        self.visited.add(node)
                    break
            if not on_path:
                reward += p["reward_fall_off"]
                self.num_fails += 1
                if p["visual_feedback"]:
                    self.cross_alpha = 255
                self.off_path = True
            else:
                self.cross_alpha = 0
                self.off_path = False
            cross = assets.fall_off_cross(SCALE)
            self.cross_x, self.cross_y = self.cx - (cross.w >> 1), self.cy - (cross.h >> 1)
        reward += p["reward_step"]
        self.rewards.append(reward)
        self.t += 1
        if self.t == self.max_episode_steps:
            done = True
        info = {}
        if done:
            info = {"reward": _seq_sum(self.rewards), "length": len(self.rewards), "success": success,
                    "num_fails": self.num_fails}
        return self._render(cross=True), reward, done, False, info


# --------------------------------------------------------------------------------------------------
# Searing Spotlights
# --------------------------------------------------------------------------------------------------
class _Spot:
    """pygame_assets.py:61-131 (Spotlight)."""

    def __init__(self, dim, radius, speed, rng):
        self.speed = speed
        self.t = 0
        self.done = False
        diagonal = math.sqrt(math.pow(dim, 2) + math.pow(dim, 2))
        self.spawn_radius = diagonal / 2 + radius
        self.radius = radius
        start = rng.integers(0, 360)
        target = start + 180 + rng.integers(-45, 45)
        offset = target + rng.integers(-135, 135)
        c = dim // 2
        self.spawn = tuple(c + v for v in R.vec_rotate(float(self.spawn_radius), 0.0, float(start)))
        self.target = tuple(c + v for v in R.vec_rotate(float(self.spawn_radius), 0.0, float(target)))
        self.offset = tuple(c + v for v in R.vec_rotate(float(self.spawn_radius), 0.0, float(offset)))
        self.cur = self.spawn

    @staticmethod
    def _lerp(a, b, t):
        t = float(t)
        return (a[0] * (1 - t) + b[0] * t, a[1] * (1 - t) + b[1] * t)

    def advance(self):
        """Spotlight.draw minus the drawing: returns the integer centre used for the circle."""
        lt = self._lerp(self.target, self.offset, self.t)
        self.cur = self._lerp(self.spawn, lt, self.t)
        centre = (int(self.cur[0]), int(self.cur[1]))
        self.t += self.speed
        if self.t >= 1.0:
            self.t = 1.0
            self.done = True
        return centre

    def hits(self, cx, cy, agent_radius):
        dx, dy = float(cx) - self.cur[0], float(cy) - self.cur[1]
        return math.sqrt(dx * dx + dy * dy) <= self.radius + agent_radius


class SearingSpotlightsOracle(_OracleEnv):
    """searing_spotlights.py:15-562 (SearingSpotlightsEnv) on CharacterController, GridPositionSampler
    (pygame_assets.py:7-59), Spotlight, Coin (:133-167), Exit (:169-220), tiled background (:222-239)."""

    env_id = "SearingSpotlights-v0"
    default_reset_parameters = {
        "max_steps": 256, "initial_spawns": 4, "num_spawns": 30, "initial_spawn_interval": 30,
        "spawn_interval_threshold": 10, "spawn_interval_decay": 0.95, "spot_min_radius": 30.0 * SCALE,
        "spot_max_radius": 55.0 * SCALE, "spot_min_speed": 0.0025, "spot_max_speed": 0.0075, "spot_damage": 1.0,
        "visual_feedback": True, "black_background": False, "hide_chessboard": False,
        "light_dim_off_duration": 6, "light_threshold": 255, "num_coins": [1], "coin_scale": 1.5 * SCALE,
        "coins_visible": False, "use_exit": True, "exit_scale": 2.0 * SCALE, "exit_visible": False,
        "agent_speed": 12.0 * SCALE, "agent_health": 5, "agent_scale": 1.0 * SCALE, "agent_visible": False,
        "sample_agent_position": True, "show_last_action": True, "show_last_positive_reward": True,
        "reward_inside_spotlight": 0.0, "reward_outside_spotlight": 0.0, "reward_death": 0.0, "reward_exit": 1.0,
        "reward_max_steps": 0.0, "reward_coin": 0.25,
    }
    BAR_H = int(16 * SCALE)

    def __init__(self):
        super().__init__()
        # the two tiled backgrounds are built once per instance and mutated in place by reset/step
        # (hide_chessboard -> white, black_background -> black): sticky across resets
        self.bg_state = {"blue": "chess", "red": "chess"}
        # spotlight_surface is created once per instance (:124-126) and only repainted by step() (:208-216): after a
        # reset it still shows the discs of the previous episode's last step. They are invisible while the surface's
        # alpha is 0 (light_dim_off_duration > 0) and visible in the reset frame when light_dim_off_duration == 0.
        self.surface_discs = []
        self.handle = None  # pose of rotated_agent_surface / rect (:135, :462): not refreshed by reset()

    # ---- GridPositionSampler -----------------------------------------------------------------------
    def _sample(self, rng, block_radius):
        n = SCREEN
        free = np.flatnonzero(~self.spawn_mask.reshape(-1))
        cell = int(free[int(rng.integers(0, len(free)))])
        y, x = cell // n, cell % n
        self._block(x, y, block_radius)
        return x, y

    def _block(self, x, y, r=21):
        n = SCREEN
        yy, xx = np.ogrid[-y:n - y, -x:n - x]
        self.spawn_mask[xx * xx + yy * yy < r * r] = True

    @staticmethod
    def _clamp_spawn(x, y, offset=30):
        o = int(offset * SCALE)
        x = o if x < o else (SCREEN - o if x > SCREEN - o else x)
        y = o if y < o else (SCREEN - o if y > SCREEN - o else y)
        return x, y

    def _new_spot(self, rng):
        p = self.p
        radius = rng.integers(p["spot_min_radius"], p["spot_max_radius"] + 1)
        speed = rng.uniform(p["spot_min_speed"], p["spot_max_speed"])
        return _Spot(SCREEN, radius, speed, rng)

    # ---- reset -------------------------------------------------------------------------------------
    def reset(self, seed=None, options=None):
        rng = self._seed(seed)
        p = self.p = self.process_reset_params(options)
        self.max_episode_steps = p["max_steps"]
        self.t = 0
        if p["hide_chessboard"]:
            self.bg_state["blue"] = self.bg_state["red"] = "white"
        self.spawn_mask = np.zeros((SCREEN, SCREEN), dtype=bool)
        self.rewards = []
        self.last_action = [0, 0]
        rotation = int(rng.choice([0, 45, 90, 135, 180, 225, 270, 315]))
        self.sprites = assets.agent_sprites(p["agent_scale"])
        self.radius = int(25 * p["agent_scale"])
        self.rotation = rotation
        if p["sample_agent_position"]:
            x, y = self._sample(rng, 28)
            x, y = x + int(rng.integers(2, 4)), y + int(rng.integers(2, 4))
        else:
            x, y = SCREEN // 2, SCREEN // 2
            self._block(x, y)
        self.cx, self.cy = int(x), int(y)
        self.agent_health = p["agent_health"]
        self.hp = self.agent_health
        q = int(SCREEN // 4)
        self.q = q
        # top bar: grey, green health on the first half, action boxes (searing_spotlights.py:372-389)
        self.bar = np.zeros((self.BAR_H, SCREEN), dtype=np.uint32)
        self.bar[:, :] = R.map_rgb((50, 50, 50))
        self.bar[:, 0:q * 2] = R.map_rgb((0, 255, 0))
        self.act_colors = [(120, 120, 120), (116, 1, 113), (255, 94, 14)]
        if p["show_last_action"]:
            self.bar[:, q * 2:q * 3] = R.map_rgb(self.act_colors[0])
            self.bar[:, q * 3:q * 4] = R.map_rgb(self.act_colors[0])
        if p["show_last_positive_reward"]:
            self.last_reward = 0.0
            if p["show_last_action"]:
                self.coin_bar = (int(q * 2.75), int(q * 0.5))
            else:
                self.coin_bar = (int(q * 2), int(q * 2))
        # spotlights: every interval is popped from the wrong end, so intervals[0] applies throughout (quirk A.9#5)
        self.spawn_intervals = []
        initial = p["initial_spawn_interval"]
        for _ in range(p["num_spawns"]):
            self.spawn_intervals.append(int(initial + p["spawn_interval_threshold"]))
            initial = initial * math.pow(p["spawn_interval_decay"], 1)
        self.alpha = 0 if p["light_dim_off_duration"] > 0 else max(0, min(255, p["light_threshold"]))
        self.spots = []
        self.spawn_timer = 0
        for _ in range(p["initial_spawns"]):
            self.spots.append(self._new_spot(rng))
        self.coins_collected = 0
        self.coins = []
        self.num_coins = int(rng.choice(p["num_coins"])) if len(p["num_coins"]) > 0 else 0
        self.coin_radius = int(10 * p["coin_scale"])
        for _ in range(self.num_coins):
            x, y = self._sample(rng, 21)
            x, y = x + int(rng.integers(2, 4)), y + int(rng.integers(2, 4))
            self.coins.append(self._clamp_spawn(x, y))
        x, y = self._sample(rng, 21)
        x, y = x + int(rng.integers(2, 4)), y + int(rng.integers(2, 4))
        self.exit = self._clamp_spawn(x, y)
        self.exit_open = False
        self.exit_radius = 20 / 2 * p["exit_scale"]
        self.bg = "blue"
        if p["black_background"]:
            self.bg_state["blue"] = "black"
        return self._render(self.surface_discs, 0), {}

    # ---- step --------------------------------------------------------------------------------------
    def step(self, action):
        p = self.p
        rng = self._rng
        self.cx, self.cy, self.rotation = move_clamped(
            self.cx, self.cy, self.rotation, action, p["agent_speed"], self.radius,
            (0, self.BAR_H, SCREEN, SCREEN - self.BAR_H))
        q = self.q
        if p["show_last_action"]:
            self.bar[:, q * 2:q * 3] = R.map_rgb(self.act_colors[int(self.last_action[0])])
            self.bar[:, q * 3:q * 4] = R.map_rgb(self.act_colors[int(self.last_action[1])])
            self.last_action = action
        if self.alpha <= p["light_threshold"]:
            if p["light_dim_off_duration"] > 0:
                self.alpha = max(0, min(255, self.alpha + int(255 / p["light_dim_off_duration"])))
            else:
                self.alpha = max(0, min(255, p["light_threshold"]))
        reward = 0.0
        # -- spotlight task (:187-242)
        r = 0.0
        spot_done = False
        self.spawn_timer += 1
        if self.spawn_intervals:
            if self.spawn_timer >= self.spawn_intervals[0]:
                self.spots.append(self._new_spot(rng))
                self.spawn_intervals.pop()
                self.spawn_timer = 0
        discs = []
        hit = 0
        i = 0
        while i < len(self.spots):  # `for spot in list` with list.remove inside: the next element is skipped
            spot = self.spots[i]
            if spot.done:
                self.spots.remove(spot)
            else:
                discs.append((spot.advance(), int(spot.radius)))
                if spot.hits(self.cx, self.cy, self.radius):
                    hit += 1
            i += 1
        if hit > 0:
            self.hp -= p["spot_damage"]
            r += p["reward_inside_spotlight"]
            width = int(SCREEN // 2 * (1 - self.hp / self.agent_health))
            if width > 0:
                self.bar[:, 0:min(width, SCREEN)] = R.map_rgb((255, 0, 0))
            self.bg = "red" if p["visual_feedback"] else "blue"
        else:
            self.bg = "blue"
            r += p["reward_outside_spotlight"]
        if p["black_background"]:
            self.bg_state[self.bg] = "black"
        if self.hp <= 0:
            spot_done = True
            r += p["reward_death"]
        reward += r
        # -- coin task (:288-311)
        if self.num_coins > 0:
            r = 0.0
            i = 0
            while i < len(self.coins):
                c = self.coins[i]
                dx, dy = float(self.cx) - float(c[0]), float(self.cy) - float(c[1])
                if math.sqrt(dx * dx + dy * dy) <= self.coin_radius + self.radius:
                    self.coins.remove(c)
                    r += p["reward_coin"]
                    self.coins_collected += 1
                i += 1
            coins_done = not self.coins
            reward += r
        else:
            coins_done = True
        # -- exit task (:313-330)
        success = 0
        exit_done = False
        if p["use_exit"]:
            r = 0.0
            if coins_done:
                self.exit_open = True
                dx, dy = float(self.cx) - float(self.exit[0]), float(self.cy) - float(self.exit[1])
                if math.sqrt(dx * dx + dy * dy) <= self.exit_radius + self.radius:
                    exit_done = True
                    r = p["reward_exit"]
            reward += r
        done = False
        if spot_done:
            done = True
        elif coins_done:
            if p["use_exit"]:
                if exit_done:
                    done = True
                    success = 1
            elif self.num_coins > 0:
                done = True
                success = 1
        self.t += 1
        if self.t == self.max_episode_steps:
            done = True
        if p["show_last_positive_reward"]:
            x0, w = self.coin_bar
            self.bar[:, x0:x0 + w] = R.map_rgb((255, 255, 0) if self.last_reward > 0 else (50, 50, 50))
            self.last_reward = reward
        self.rewards.append(reward)
        info = {}
        if done:
            info = {"reward": _seq_sum(self.rewards), "length": len(self.rewards),
                    "agent_health": self.hp / self.agent_health,
                    "coins_collected": self.coins_collected / self.num_coins, "success": success}
        self.surface_discs = discs
        self.handle = (self.cx, self.cy, self.rotation)
        return self._render(discs, self.rotation), reward, done, False, info

    # ---- frame (searing_spotlights.py:526-543 / 419-437) ---------------------------------------------
    def _layer_order(self):
        p = self.p
        layers = ["bg", "spot", "bar"]
        sid = 1
        if p["coins_visible"]:
            layers.insert(sid + 1, "coin")
        else:
            layers.insert(sid, "coin")
            sid += 1
        if p["exit_visible"]:
            layers.insert(sid + 2, "exit")
        else:
            layers.insert(sid, "exit")
            sid += 1
        if p["agent_visible"]:
            layers.insert(sid + 3, "agent")
        else:
            layers.insert(sid, "agent")
            sid += 1
        return layers

    DEBUG_LAYERS = ["bg", "spot", "exit", "coin", "agent", "bar"]

    def _debug_surface(self):
        """_build_debug_surface (searing_spotlights.py:157-185, endless_searing_spotlights.py:150-177): background, the
        spotlight surface as step() left it (with its current alpha), then exit, coins, the sprite handle of the last
        step (not refreshed by reset()) and the top bar — everything the agent may not see is drawn above the light."""
        pose = self.handle if self.handle is not None else (self.cx, self.cy, 0)
        obs = self._render(self.surface_discs, pose[2], layers=self.DEBUG_LAYERS, pose=pose, debug=True)
        px = obs.astype(np.uint32)
        return ((px[..., 0] << 16) | (px[..., 1] << 8) | px[..., 2]).T  # back to the [y][x] pixel array

    def _render(self, discs, sprite_rot, layers=None, pose=None, debug=False):
        p = self.p
        dst = R.Surface(SCREEN, SCREEN)
        ax, ay = (self.cx, self.cy) if pose is None else (pose[0], pose[1])
        for layer in (layers if layers is not None else self._layer_order()):
            if layer == "bg":
                st = self.bg_state[self.bg]
                if st == "chess":
                    src = assets.chessboard(SCREEN, (0, 0, 255) if self.bg == "blue" else (255, 0, 0), SCALE)
                    dst.px[...] = src.px
                else:
                    dst.px[...] = R.map_rgb((255, 255, 255)) if st == "white" else 0
            elif layer == "coin":
                if getattr(self, "coin_layer_opaque", False) and not debug:
                    dst.px[...] = 0
                elif self.num_coins > 0:
                    cs = R.Surface(SCREEN, SCREEN)
                    R.fill(cs, 255)
                    cs.colorkey = 255
                    for c in self.coins:
                        R.circle(cs, (255, 255, 0), c, self.coin_radius)
                        R.circle(cs, (255, 165, 0), c, self.coin_radius, int(2 * p["coin_scale"]))
                    R.blit(cs, dst, (0, 0))
            elif layer == "exit":
                es = assets.exit_sprite(p["exit_scale"], self.exit_open)
                R.blit(es, dst, (self.exit[0] - (es.w >> 1), self.exit[1] - (es.h >> 1)))
            elif layer == "agent":
                sp = self.sprites[sprite_rot // 45]
                R.blit(sp, dst, (ax - (sp.w >> 1), ay - (sp.h >> 1)))
            elif layer == "spot":
                ss = R.Surface(SCREEN, SCREEN)
                ss.colorkey = R.map_rgb((255, 0, 0))
                ss.alpha = self.alpha
                for centre, radius in discs:
                    R.circle(ss, (255, 0, 0), centre, radius)
                    if p["black_background"]:
                        R.circle(ss, (255, 255, 255), centre, radius, 1)
                R.blit(ss, dst, (0, 0))
            elif layer == "bar":
                bar = R.Surface(SCREEN, self.BAR_H)
                bar.px[...] = self.bar
                R.blit(bar, dst, (0, 0))
        return R.array3d(dst)


class EndlessSearingSpotlightsOracle(SearingSpotlightsOracle):
    """endless_searing_spotlights.py:15-504 (EndlessSearingSpotlightsEnv): one respawning coin, no exit, constant
    spawn interval, ``steps_per_coin`` time limit, coin layer above the spotlights for ``coin_show_duration`` steps
    after every (re)spawn, ground truth = (agent centre, coin location) / 84 every step."""

    env_id = "Endless-SearingSpotlights-v0"
    default_reset_parameters = {
        "max_steps": -1, "steps_per_coin": 160, "initial_spawns": 3, "spawn_interval": 50,
        "spot_min_radius": 30.0 * SCALE, "spot_max_radius": 55.0 * SCALE, "spot_min_speed": 0.0025,
        "spot_max_speed": 0.0075, "spot_damage": 1.0, "visual_feedback": True, "black_background": False,
        "hide_chessboard": False, "light_dim_off_duration": 6, "light_threshold": 255, "coin_enabled": True,
        "coin_scale": 1.5 * SCALE, "coin_show_duration": 6, "coins_visible": False, "agent_speed": 12.0 * SCALE,
        "agent_health": 10, "agent_scale": 1.0 * SCALE, "agent_visible": False, "sample_agent_position": True,
        "show_last_action": True, "show_last_positive_reward": True, "reward_inside_spotlight": 0.0,
        "reward_outside_spotlight": 0.0, "reward_death": 0.0, "reward_coin": 0.25,
    }

    @classmethod
    def _check_params(cls, p):
        # step() paints the action boxes unconditionally with self.action_colors, which only exists after a
        # reset with show_last_action (endless_searing_spotlights.py:349-354, 421-424): AttributeError otherwise
        assert p["show_last_action"], "the reference crashes in step() without show_last_action"

    def _gt(self):
        coin = np.asarray(self.coin, dtype=np.float64) if self.p["coin_enabled"] else np.zeros(2)
        return np.concatenate((np.asarray([self.cx, self.cy], dtype=np.float64), coin)) / SCREEN

    def _spawn_coin(self, rng):
        """_spawn_coin (:256-269): the sampler is reset, only the OLD coin's surroundings are blocked."""
        self.spawn_mask = np.zeros((SCREEN, SCREEN), dtype=bool)
        if self.coin is not None:
            self._block(self.coin[0], self.coin[1], 28)
        x, y = self._sample(rng, 28)
        x, y = x + int(rng.integers(2, 4)), y + int(rng.integers(2, 4))
        self.coin = self._clamp_spawn(x, y)

    def reset(self, seed=None, options=None):
        rng = self._seed(seed)
        p = self.p = self.process_reset_params(options)
        self.max_episode_steps = p["max_steps"]
        self.t = 0
        self.coin_t = 0
        if p["hide_chessboard"]:
            self.bg_state["blue"] = self.bg_state["red"] = "white"
        self.spawn_mask = np.zeros((SCREEN, SCREEN), dtype=bool)
        self.rewards = []
        self.last_action = [0, 0]
        rotation = int(rng.choice([0, 45, 90, 135, 180, 225, 270, 315]))
        self.sprites = assets.agent_sprites(p["agent_scale"])
        self.radius = int(25 * p["agent_scale"])
        self.rotation = rotation
        if p["sample_agent_position"]:
            x, y = self._sample(rng, 28)
            x, y = x + int(rng.integers(2, 4)), y + int(rng.integers(2, 4))
        else:
            x, y = SCREEN // 2, SCREEN // 2
            self._block(x, y)
        self.cx, self.cy = int(x), int(y)
        self.agent_health = p["agent_health"]
        self.hp = self.agent_health
        q = self.q = int(SCREEN // 4)
        # top bar: a fresh (black) surface, green health half, action boxes (:345-360) — no grey fill here
        self.bar = np.zeros((self.BAR_H, SCREEN), dtype=np.uint32)
        self.bar[:, 0:q * 2] = R.map_rgb((0, 255, 0))
        self.act_colors = [(120, 120, 120), (116, 1, 113), (255, 94, 14)]
        self.bar[:, q * 2:q * 3] = R.map_rgb(self.act_colors[0])
        self.bar[:, q * 3:q * 4] = R.map_rgb(self.act_colors[0])
        if p["show_last_positive_reward"]:
            self.last_reward = 0.0
            self.coin_bar = (int(q * 2.75), int(q * 0.5))
        self.alpha = 0 if p["light_dim_off_duration"] > 0 else max(0, min(255, p["light_threshold"]))
        self.spots = []
        self.spawn_timer = 0
        for _ in range(p["initial_spawns"]):
            self.spots.append(self._new_spot(rng))
        self.coins_collected = 0
        self.coin = None
        self.coin_radius = int(10 * p["coin_scale"])
        if p["coin_enabled"]:
            self._spawn_coin(rng)
        self.bg = "blue"
        if p["black_background"]:
            self.bg_state["blue"] = "black"
        return self._render(self.surface_discs, 0), {"ground_truth": self._gt()}

    def step(self, action):
        p = self.p
        rng = self._rng
        self.cx, self.cy, self.rotation = move_clamped(
            self.cx, self.cy, self.rotation, action, p["agent_speed"], self.radius,
            (0, self.BAR_H, SCREEN, SCREEN - self.BAR_H))
        q = self.q
        self.bar[:, q * 2:q * 3] = R.map_rgb(self.act_colors[int(self.last_action[0])])
        self.bar[:, q * 3:q * 4] = R.map_rgb(self.act_colors[int(self.last_action[1])])
        self.last_action = action
        if self.alpha <= p["light_threshold"]:
            if p["light_dim_off_duration"] > 0:
                self.alpha = max(0, min(255, self.alpha + int(255 / p["light_dim_off_duration"])))
            else:
                self.alpha = max(0, min(255, p["light_threshold"]))
        reward = 0.0
        # -- spotlight task (:179-234): a new spotlight every spawn_interval steps, forever
        r = 0.0
        spot_done = False
        self.spawn_timer += 1
        if self.spawn_timer >= p["spawn_interval"]:
            self.spots.append(self._new_spot(rng))
            self.spawn_timer = 0
        discs = []
        hit = 0
        i = 0
        while i < len(self.spots):
            spot = self.spots[i]
            if spot.done:
                self.spots.remove(spot)
            else:
                discs.append((spot.advance(), int(spot.radius)))
                if spot.hits(self.cx, self.cy, self.radius):
                    hit += 1
            i += 1
        if hit > 0:
            self.hp -= p["spot_damage"]
            r += p["reward_inside_spotlight"]
            width = int(SCREEN // 2 * (1 - self.hp / self.agent_health))
            if width > 0:
                self.bar[:, 0:min(width, SCREEN)] = R.map_rgb((255, 0, 0))
            self.bg = "red" if p["visual_feedback"] else "blue"
        else:
            self.bg = "blue"
            r += p["reward_outside_spotlight"]
        if p["black_background"]:
            self.bg_state[self.bg] = "black"
        if self.hp <= 0:
            spot_done = True
            r += p["reward_death"]
        reward += r
        # -- coin task (:271-292)
        if p["coin_enabled"]:
            r = 0.0
            dx, dy = float(self.cx) - float(self.coin[0]), float(self.cy) - float(self.coin[1])
            if math.sqrt(dx * dx + dy * dy) <= self.coin_radius + self.radius:
                r += p["reward_coin"]
                self.coins_collected += 1
                self.coin_t = 0
                self._spawn_coin(rng)
            reward += r
        done = spot_done
        self.t += 1
        self.coin_t += 1
        if self.coin_t == p["steps_per_coin"] and p["coin_enabled"]:
            done = True
        if self.t == self.max_episode_steps:
            done = True
        if p["show_last_positive_reward"]:
            x0, w = self.coin_bar
            self.bar[:, x0:x0 + w] = R.map_rgb((255, 255, 0) if self.last_reward > 0 else (50, 50, 50))
            self.last_reward = reward
        self.rewards.append(reward)
        info = {"ground_truth": self._gt()}
        if done:
            info = {"reward": _seq_sum(self.rewards), "length": len(self.rewards),
                    "agent_health": self.hp / self.agent_health, "coins_collected": self.coins_collected,
                    "ground_truth": self._gt()}
        self.surface_discs = discs
        self.handle = (self.cx, self.cy, self.rotation)
        return self._render(discs, self.rotation), reward, done, False, info

    # ---- frame (:376-391 / 463-478): the coin layer sits above the spotlights while the coin is fresh ----
    def _layer_order(self):
        p = self.p
        layers = ["bg", "spot", "bar"]
        sid = 1
        if p["coins_visible"] or self.coin_t < p["coin_show_duration"]:
            layers.insert(sid + 1, "coin")
        else:
            layers.insert(sid, "coin")
            sid += 1
        if p["agent_visible"]:
            layers.insert(sid + 3, "agent")
        else:
            layers.insert(sid, "agent")
            sid += 1
        return layers

    DEBUG_LAYERS = ["bg", "spot", "coin", "agent", "bar"]

    def _render(self, discs, sprite_rot, layers=None, pose=None, debug=False):
        # coin_surface is a fresh opaque black surface when the coin is disabled (never filled / colour-keyed);
        # the debug render builds its own colour-keyed coin surface (:157-162)
        self.num_coins = 1
        self.coins = [self.coin] if self.p["coin_enabled"] else []
        self.coin_layer_opaque = not self.p["coin_enabled"]
        return super()._render(discs, sprite_rot, layers, pose, debug)


# --------------------------------------------------------------------------------------------------
# Endless Mystery Path
# --------------------------------------------------------------------------------------------------
class EndlessMysteryPathOracle(_OracleEnv):
    """endless_mystery_path.py:16-444 (EndlessMysteryPathEnv) on EndlessMysteryPath
    (pygame_assets.py:495-604: 7x7 MysteryPath segments joined by transition tiles, 8 columns per segment),
    CharacterController.step without a boundary rect, a camera that follows the agent's x velocity."""

    env_id = "Endless-MysteryPath-v0"
    default_reset_parameters = {
        "max_steps": -1, "agent_scale": 1.0 * SCALE, "agent_speed": 12.0 * SCALE, "show_origin": False,
        "show_past_path": True, "show_background": False, "show_stamina": False, "visual_feedback": True,
        "camera_offset_scale": 5.0, "stamina_level": 20, "reward_fall_off": 0.0, "reward_path_progress": 0.1,
        "reward_path_progress_dense": 0.0, "reward_step": 0.0,
    }
    GRID = 7
    TILE = SCREEN // 7  # 12
    SEG = 8             # columns per segment incl. the transition tile (segment_length)
    ACTIONS = ([0, 0], [2, 0], [0, 1], [0, 2])  # Discrete(4): no-op, right, up, down (:283-291)

    # ---- EndlessMysteryPath.add_path_segment (pygame_assets.py:559-604) -----------------------------
    def _add_segment(self, rng):
        g = self.GRID
        if self.seg_start is None:
            self.seg_start = (0, int(rng.integers(0, g)))
        else:
            self.seg_start = (0, self.seg_end[1])
        self.seg_end = (g - 1, int(rng.integers(0, g)))
        path, _ = generate_mystery_path(rng, g, self.seg_start, self.seg_end)  # [end, ..., start]
        k = len(self.segments)
        shift = 0 if k == 0 else k * g + k
        nodes = [(x + shift, y) for x, y in reversed(path)]
        k += 1
        nodes.append((self.seg_end[0] + k + (k - 1) * g, self.seg_end[1]))  # the transition tile
        self.segments.append(nodes)

    def _node(self, ref):
        return self.segments[ref[0]][ref[1]]

    def _prev(self, ref):
        s, i = ref
        if i > 0:
            return (s, i - 1)
        return (s - 1, len(self.segments[s - 1]) - 1) if s > 0 else None

    def _next(self, ref):
        s, i = ref
        if i + 1 < len(self.segments[s]):
            return (s, i + 1)
        return (s + 1, 0) if s + 1 < len(self.segments) else None

    def _update_direction(self):
        nx = self._next(self.current)
        x = self._node(nx)[0] - self._node(self.current)[0]
        y = self._node(nx)[1] - self._node(self.current)[1]
        if x == 1:
            self.target_direction = np.asarray([1, 0, 0])
        elif y == -1:
            self.target_direction = np.asarray([0, 1, 0])
        elif y == 1:
            self.target_direction = np.asarray([0, 0, 1])
        # a westward link leaves the previous direction in place

    def reset(self, seed=None, options=None):
        rng = self._seed(seed)
        p = self.p = self.process_reset_params(options)
        self.t = 0
        self.max_episode_steps = p["max_steps"]
        self.rewards = []
        self.segments, self.seg_start, self.seg_end = [], None, None
        for _ in range(3):
            self._add_segment(rng)
        self.reward_visited = {self.segments[0][0]}
        self.stamina_visited = set()
        self.start = self.segments[0][0]
        cos = max(0, min(5.5, p["camera_offset_scale"]))
        self.camera_offset = -self.TILE * cos
        self.camera_x = self.camera_offset
        self.bg_scroll = 0
        self.sprites = assets.agent_sprites(p["agent_scale"])
        self.radius = int(25 * p["agent_scale"])
        self.rotation = 270
        self.vx = 0
        self.cx = self.start[0] * self.TILE + self.radius
        self.cy = self.start[1] * self.TILE + self.radius
        self.agent_draw_x = (self.cx - (self.sprites[0].w >> 1)) - self.camera_offset
        self.pos = (self.cx // self.TILE, self.cy // self.TILE)
        self.current = (0, 0)
        self._update_direction()
        self.off_path = False
        self.num_fails = 0
        self.fall_off_locations = []
        self.stamina = p["stamina_level"]
        self.max_x = 0
        self.tiles_visited = 0
        self.max_visible_stamina = p["stamina_level"]
        self.bar_height = int(SCREEN * (1 - (max(self.stamina, self.max_visible_stamina) / self.max_visible_stamina)))
        self.cross_alpha = 0
        self.cross_x = self.cross_y = 0
        return self._render(), {"ground_truth": self.target_direction}

    def step(self, action):
        p = self.p
        rng = self._rng
        action = self.ACTIONS[int(action)]
        reward = 0
        done = False
        if not self.off_path:
            before = self.cx
            self.cx, self.cy, self.rotation = move_free(self.cx, self.cy, self.rotation, action, p["agent_speed"])
            self.vx = self.cx - before
            self.camera_x += float(self.vx)
            self.bg_scroll -= float(self.vx)
            if abs(self.bg_scroll) >= self.TILE:
                remainder = abs(self.bg_scroll) % abs(float(self.vx))
                sign = self.bg_scroll / abs(self.bg_scroll)
                self.bg_scroll = remainder * sign
        else:
            self.cx = self.start[0] * self.TILE + self.radius
            self.cy = self.start[1] * self.TILE + self.radius
            self.vx = 0
            self.camera_x = self.camera_offset
            self.bg_scroll = 0
        self.pos = (self.cx // self.TILE, self.cy // self.TILE)
        seg = int(self.pos[0]) // self.SEG
        nodes = self.segments[seg]
        if seg > len(self.segments) - 2:
            self._add_segment(rng)
        on_path = False
        for i, node in enumerate(nodes):
            if self.pos == node:
                on_path = True
                self.current = (seg, i)
                if node not in self.reward_visited and node != self.start:
                    reward += p["reward_path_progress"]
                    self.tiles_visited += 1
                    self.reward_visited.add(node)
                if node not in self.stamina_visited and node != self.start:
                    reward += p["reward_path_progress_dense"]
                    self.stamina = p["stamina_level"]
                    self.stamina_visited.add(node)
                break
        if not on_path:
            reward += p["reward_fall_off"]
            self.num_fails += 1
            if p["visual_feedback"]:
                self.cross_alpha = 255
            self.off_path = True
            if self.pos[0] < self.max_x:
                done = True
            elif self.pos in self.fall_off_locations:
                done = True
            else:
                self.fall_off_locations.append(self.pos)
            self.stamina_visited = set()
            self.stamina = p["stamina_level"]
        else:
            self.cross_alpha = 0
            self.off_path = False
        # fall_off_rect.center = (cx - camera_x, cy): a 10x10 rect, floats C-cast
        self.cross_x = R.cint(self.cx - self.camera_x) - 5
        self.cross_y = self.cy - 5
        # (`max_x_reached > previous_max_x` can never hold here: fall_off_locations is never cleared, :390-391)
        reward += p["reward_step"]
        self.stamina -= 1
        if self.stamina == 0:
            done = True
        self.t += 1
        if self.t == self.max_episode_steps:
            done = True
        self._update_direction()
        if self.pos[0] > self.max_x and on_path:
            self.max_x = self.pos[0]
        self.rewards.append(reward)
        if done:
            info = {"reward": _seq_sum(self.rewards), "length": len(self.rewards), "num_fails": self.num_fails,
                    "max_x": self.max_x, "tiles_visited": self.tiles_visited, "ground_truth": self.target_direction}
        else:
            info = {"ground_truth": self.target_direction}
        self.bar_height = int(SCREEN * (1 - (min(self.stamina, self.max_visible_stamina) / self.max_visible_stamina)))
        return self._render(), reward, done, False, info

    def _debug_surface(self):
        """_build_debug_surface (endless_mystery_path.py:162-182): scrolling background (if shown), the WHOLE path
        generated so far (EndlessMysteryPath.gen_surface, pygame_assets.py:544-557: white tiles on a colour-keyed surface
        with alpha 200) at -camera_x, the agent, the fall-off cross and — always — the stamina bar."""
        T = self.TILE
        dst = R.Surface(SCREEN, SCREEN)
        self._draw_background(dst)
        n = len(self.segments)
        path = R.Surface((self.GRID * n + n) * T, self.GRID * T)
        path.colorkey = 0
        path.alpha = 200
        for nodes in self.segments:
            for x, y in nodes:
                R.fill_rect(path, R.map_rgb((255, 255, 255)), x * T, y * T, T, T)
        R.blit(path, dst, (-self.camera_x, 0))
        sp = self.sprites[self.rotation // 45]
        R.blit(sp, dst, (self.agent_draw_x, self.cy - (sp.h >> 1)))
        if self.cross_alpha == 255:
            R.blit(assets.fall_off_cross(SCALE), dst, (self.cross_x, self.cross_y))
        bar = R.Surface(int(16 * SCALE), SCREEN)
        R.fill(bar, (0, 255, 0))
        R.rect(bar, (255, 0, 0), (0, 0, 16 * SCALE, self.bar_height))
        R.blit(bar, dst, (SCREEN - 16 * SCALE, 0))
        return dst.px

    def _draw_background(self, dst):
        p = self.p
        T = self.TILE
        if p["show_background"]:
            col = R.Surface(T, T * self.GRID)  # draw_column_tile_surface (pygame_assets.py:801-818)
            for i in range(self.GRID):
                R.fill_rect(col, R.map_rgb((125, 177, 250)), 0, i * T, T, T)
                tile = R.Surface(T, T)
                R.fill(tile, (125, 177, 250))
                R.rect(tile, (210, 210, 210), (0, 0, T, T), 1)
                R.blit(tile, col, (0, i * T))
            for i in range(math.ceil(SCREEN / T) + 2):
                R.blit(col, dst, (i * T + self.bg_scroll - T, 0))

    # ---- _draw_surfaces / _draw_past_path (:111-160) ---------------------------------------------------
    def _render(self):
        p = self.p
        T = self.TILE
        dst = R.Surface(SCREEN, SCREEN)
        if p["show_background"]:
            col = R.Surface(T, T * self.GRID)  # draw_column_tile_surface (pygame_assets.py:801-818)
            for i in range(self.GRID):
                R.fill_rect(col, R.map_rgb((125, 177, 250)), 0, i * T, T, T)
                tile = R.Surface(T, T)
                R.fill(tile, (125, 177, 250))
                R.rect(tile, (210, 210, 210), (0, 0, T, T), 1)
                R.blit(tile, col, (0, i * T))
            for i in range(math.ceil(SCREEN / T) + 2):
                R.blit(col, dst, (i * T + self.bg_scroll - T, 0))
        if p["show_past_path"]:
            x = self.pos[0] - 1
            if x >= 0:
                depth = int(p["camera_offset_scale"])
                past_x = max(0, x - depth)
                ref = self._prev(self.current)
                while x >= past_x and x >= 0:
                    if ref is None:
                        break
                    x, y = self._node(ref)
                    draw_x = x * T - self.camera_x
                    R.rect(dst, (255, 255, 255), (draw_x, y * T, T, T))
                    R.rect(dst, (210, 210, 210), (draw_x, y * T, T, T), 1)
                    if x == past_x:
                        break
                    ref = self._prev(ref)
        sp = self.sprites[self.rotation // 45]
        R.blit(sp, dst, (self.agent_draw_x, self.cy - (sp.h >> 1)))
        if p["show_stamina"]:
            bar = R.Surface(int(16 * SCALE), SCREEN)
            R.fill(bar, (0, 255, 0))
            R.rect(bar, (255, 0, 0), (0, 0, 16 * SCALE, self.bar_height))
            R.blit(bar, dst, (SCREEN - 16 * SCALE, 0))
        if p["visual_feedback"] and self.cross_alpha == 255:
            R.blit(assets.fall_off_cross(SCALE), dst, (self.cross_x, self.cross_y))
        return R.array3d(dst)


ORACLES = {
    "MortarMayhem-v0": MortarMayhemOracle,
    "MortarMayhem-Grid-v0": GridMortarMayhemOracle,
    "Endless-MortarMayhem-v0": EndlessMortarMayhemOracle,
    "MysteryPath-v0": MysteryPathOracle,
    "SearingSpotlights-v0": SearingSpotlightsOracle,
    "MortarMayhemB-v0": MortarMayhemBOracle,
    "MortarMayhemB-Grid-v0": GridMortarMayhemBOracle,
    "MysteryPath-Grid-v0": GridMysteryPathOracle,
    "Endless-SearingSpotlights-v0": EndlessSearingSpotlightsOracle,
    "Endless-MysteryPath-v0": EndlessMysteryPathOracle,
}


def make(env_id, render_mode=None):
    env = ORACLES[env_id]()
    env.render_mode = render_mode
    return env
