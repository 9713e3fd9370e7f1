"""Static drawable assets of Memory Gym, rebuilt with the raster oracle (TEST INFRASTRUCTURE ONLY).

Each builder restates one asset constructor of the reference at an arbitrary ``scale`` (the envs
use SCALE = 0.25; the docs GIFs used 1.0, which is what the known-answer tests exercise):

* ``agent_sprites``      <- character_controller.py:29-75  (CharacterController.create_character_sprites)
* ``command_icon``       <- pygame_assets.py:254-304       (Command.__init__)
* ``mortar_tile``        <- pygame_assets.py:316-343       (MortarTile.__init__ / toggle_color)
* ``fall_off_cross``     <- mystery_path.py:175-182
* ``exit_sprite``        <- pygame_assets.py:170-205       (Exit.__init__ / draw)
* ``chessboard``         <- pygame_assets.py:222-239       (get_tiled_background_surface)

``legacy_thick`` selects the pygame<=2.1.2 thick-circle algorithm (what the docs GIFs were rendered
with); the default is the pygame>=2.1.3 algorithm that the pinned pygame==2.4.0 ships.
"""
from __future__ import annotations

import functools

from . import sdl_raster as R

COMMAND_NAMES = ("right", "down", "left", "up", "stay", "right_down", "right_up", "left_down", "left_up")
# pygame_assets.py:242-252 (Command.COMMANDS), in dict order
COMMAND_DELTAS = ((1, 0), (0, 1), (-1, 0), (0, -1), (0, 0), (1, 1), (1, -1), (-1, 1), (-1, -1))
COMMAND_ANGLES = {"right": 0, "left": 180, "down": 270, "up": 90, "right_down": 315, "right_up": 45,
                  "left_down": 225, "left_up": 135}

BODY = (250, 204, 153)
HAND = (250, 250, 250)
HAND_OUTLINE = (50, 50, 50)
BLUE_KEY = 255  # raw mapped pixel 255 == (0, 0, 255)


@functools.lru_cache(maxsize=None)
def agent_sprites(scale: float, legacy_thick: bool = False):
    """Eight pre-rotated sprites; index k = rotation // 45 (rotated by 360 - 45k degrees)."""
    radius = int(25 * scale)
    hands_x = int(18 * scale)
    hand_y = int(12 * scale)
    hand_r = int(10 * scale)
    outline = int(3 * scale)
    ext = 14
    dim = radius * 2 + hand_r + ext
    c = (dim // 2, dim // 2)
    left = (dim // 2 - hands_x, hand_y + ext // 2)
    right = (dim // 2 + hands_x, hand_y + ext // 2)
    out = []
    for ang in range(360, 0, -45):
        s = R.Surface(dim, dim)
        R.fill(s, BLUE_KEY)
        s.colorkey = BLUE_KEY
        R.circle(s, BODY, c, radius)
        lx, ly = R.vec_rotate(float(left[0] - c[0]), float(left[1] - c[1]), float(ang))
        rx, ry = R.vec_rotate(float(right[0] - c[0]), float(right[1] - c[1]), float(ang))
        lp, rp = (lx + c[0], ly + c[1]), (rx + c[0], ry + c[1])
        R.circle(s, HAND, lp, hand_r)
        R.circle(s, HAND, rp, hand_r)
        R.circle(s, HAND_OUTLINE, lp, hand_r, outline, legacy_thick=legacy_thick)
        R.circle(s, HAND_OUTLINE, rp, hand_r, outline, legacy_thick=legacy_thick)
        out.append(s)
    return tuple(out)


@functools.lru_cache(maxsize=None)
def command_icon(name: str, scale: float, legacy_thick: bool = False):
    """Command icon surface (colour key black). ``""`` is the blank delay frame."""
    assert name in COMMAND_NAMES or name == ""
    dim = 88 * scale
    s = R.Surface(dim, dim)
    s.colorkey = 0
    lw = int(8 * scale)
    white = (255, 255, 255)
    if name == "stay":
        radius = dim // 2 - 4 * scale
        x = dim - 12 * scale
        y = dim // 2 - 8 * scale
        R.circle(s, white, (radius, radius), radius, lw, legacy_thick=legacy_thick)
        R.line(s, white, (0, y), (x, y), lw)
    elif name:
        x1, x2, y1, y2 = 2 * scale, 80 * scale, 40 * scale, 0
        R.line(s, white, (x1, y1), (x2, y1), lw)
        R.line(s, white, (x2, y1), (y1, y2), lw)
        R.line(s, white, (x2, y1), (y1, x2), lw)
        s = R.rotate(s, COMMAND_ANGLES[name])
    return s


def command_blit_pos(scale: float, screen_dim: int):
    """mortar_mayhem.py:363: always the un-rotated icon's top-left, whatever the rotated size."""
    dim = 88 * scale
    return (screen_dim // 2 - dim // 2, screen_dim // 2 - dim // 2)


TILE_BLUE, TILE_LIGHT_BLUE = (21, 43, 77), (29, 60, 107)
TILE_RED, TILE_LIGHT_RED = (81, 18, 26), (112, 24, 36)


@functools.lru_cache(maxsize=None)
def mortar_tile(scale: float, red: bool):
    dim = 56 * scale
    s = R.Surface(dim, dim)
    R.rect(s, TILE_RED if red else TILE_BLUE, (0, 0, dim, dim))
    R.rect(s, TILE_LIGHT_RED if red else TILE_LIGHT_BLUE, (0, 0, dim, dim), int(4 * scale))
    return s


@functools.lru_cache(maxsize=None)
def fall_off_cross(scale: float):
    dim = 40 * scale
    s = R.Surface(dim, dim)
    s.colorkey = 0
    R.line(s, (255, 0, 0), (0, 0), (dim - 1, dim - 1), int(12 * scale))
    R.line(s, (255, 0, 0), (dim - 1, 0), (0, dim - 1), int(12 * scale))
    return s


EXIT_OPEN, EXIT_CLOSED = (48, 141, 70), (55, 55, 55)


@functools.lru_cache(maxsize=None)
def exit_sprite(scale: float, is_open: bool):
    dim = 20 * scale
    s = R.Surface(dim, dim)
    R.fill(s, BLUE_KEY)
    s.colorkey = BLUE_KEY
    corners = (int(10 * scale), int(10 * scale), 0, 0)
    r = (0, 0, s.w, s.h)
    R.rect(s, EXIT_OPEN if is_open else EXIT_CLOSED, r, 0, 0, *corners)
    R.rect(s, 0, r, int(2 * scale), 0, *corners)
    return s


@functools.lru_cache(maxsize=None)
def chessboard(screen_dim: int, tile_color, scale: float):
    s = R.Surface(screen_dim, screen_dim)
    ts = int(50 * scale)
    n = (screen_dim + ts - 1) // ts
    for x in range(n):
        for y in range(n):
            R.rect(s, (255, 255, 255) if (x + y) % 2 == 0 else tile_color, (x * ts, y * ts, ts, ts))
    return s
