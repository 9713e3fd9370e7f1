"""Minimal stand-in for pygame 2.4.0 (see ../README.md). Pixel work -> oracle.sdl_raster."""
import sys as _sys
import types as _types

from oracle import sdl_raster as _R

NOFRAME = 32
QUIT = 256
KEYDOWN = 768
K_UP, K_DOWN, K_LEFT, K_RIGHT, K_w, K_a, K_s, K_d, K_SPACE, K_PAGEUP, K_PAGEDOWN = range(11)


def _two_ints(obj):
    return _R.cint(obj[0]), _R.cint(obj[1])


class Rect:
    """pygame.Rect: four C ints; float arguments are truncated toward zero (pg_IntFromObj)."""

    __slots__ = ("x", "y", "w", "h")

    def __init__(self, *args):
        if len(args) == 1:
            args = tuple(args[0])
        if len(args) == 2:
            args = (*args[0], *args[1])
        self.x, self.y, self.w, self.h = (_R.cint(a) for a in args)

    def copy(self):
        return Rect(self.x, self.y, self.w, self.h)

    def __iter__(self):
        return iter((self.x, self.y, self.w, self.h))

    def __len__(self):
        return 4

    def __getitem__(self, i):
        return (self.x, self.y, self.w, self.h)[i]

    @property
    def center(self):
        return (self.x + (self.w >> 1), self.y + (self.h >> 1))

    @center.setter
    def center(self, v):
        cx, cy = _two_ints(v)
        self.x += cx - (self.x + (self.w >> 1))
        self.y += cy - (self.y + (self.h >> 1))

    @property
    def topleft(self):
        return (self.x, self.y)

    @property
    def bottomright(self):
        return (self.x + self.w, self.y + self.h)

    @property
    def width(self):
        return self.w

    @property
    def height(self):
        return self.h

    def __repr__(self):
        return f"<rect({self.x}, {self.y}, {self.w}, {self.h})>"


class Surface:
    def __init__(self, size, flags=0, _raw=None):
        self._s = _raw if _raw is not None else _R.Surface(size[0], size[1])

    def fill(self, color, rect=None):
        if rect is None:
            _R.fill(self._s, color)
        else:
            x, y, w, h = (_R.cint(v) for v in rect)
            _R.fill_rect(self._s, _R.map_rgb(color), x, y, w, h)

    def set_colorkey(self, color):
        self._s.colorkey = None if color is None else _R.map_rgb(color)

    def set_alpha(self, value):
        self._s.alpha = None if value is None else max(0, min(255, _R.cint(value)))

    def get_alpha(self):
        return self._s.alpha

    def get_size(self):
        return (self._s.w, self._s.h)

    def get_width(self):
        return self._s.w

    def get_height(self):
        return self._s.h

    def get_rect(self, **kwargs):
        r = Rect(0, 0, self._s.w, self._s.h)
        for k, v in kwargs.items():
            setattr(r, k, v)
        return r

    def copy(self):
        return Surface((0, 0), _raw=self._s.copy())

    def blit(self, source, dest, area=None, special_flags=0):
        assert area is None and special_flags == 0
        if isinstance(dest, Rect):
            pos = (dest.x, dest.y)
        else:
            pos = _two_ints(dest)
        _R.blit(source._s, self._s, pos)

    def __bool__(self):
        return True


# ---- submodules -----------------------------------------------------------------------------------
def _mod(name):
    m = _types.ModuleType("pygame." + name)
    _sys.modules["pygame." + name] = m
    return m


draw = _mod("draw")


def _draw_circle(surface, color, center, radius, width=0, *quadrants):
    assert not any(quadrants)
    _R.circle(surface._s, color, center, radius, width)


def _draw_line(surface, color, start_pos, end_pos, width=1):
    _R.line(surface._s, color, start_pos, end_pos, width)


def _draw_rect(surface, color, rect, width=0, border_radius=0, border_top_left_radius=-1,
               border_top_right_radius=-1, border_bottom_left_radius=-1, border_bottom_right_radius=-1):
    _R.rect(surface._s, color, tuple(rect), width, border_radius, border_top_left_radius,
            border_top_right_radius, border_bottom_left_radius, border_bottom_right_radius)


draw.circle, draw.line, draw.rect = _draw_circle, _draw_line, _draw_rect

transform = _mod("transform")
transform.rotate = lambda surface, angle: Surface((0, 0), _raw=_R.rotate(surface._s, angle))


def _scale(surface, size):
    import numpy as np
    w, h = _two_ints(size)
    out = _R.Surface(w, h)
    ys = (np.arange(h) * surface._s.h) // h
    xs = (np.arange(w) * surface._s.w) // w
    out.px[...] = surface._s.px[ys][:, xs]
    return Surface((0, 0), _raw=out)


transform.scale = _scale

surfarray = _mod("surfarray")
surfarray.array3d = lambda surface: _R.array3d(surface._s)
surfarray.pixels3d = surfarray.array3d

display = _mod("display")
_display_surface = [None]


def _set_mode(size, flags=0, *a, **k):
    _display_surface[0] = Surface(size)
    return _display_surface[0]


display.set_mode = _set_mode
display.get_surface = lambda: _display_surface[0]
display.flip = lambda: None
display.set_caption = lambda *a, **k: None

time = _mod("time")
time.Clock = type("Clock", (), {"tick": lambda self, fps=0: 0})
event = _mod("event")
event.set_allowed = lambda *a: None
event.get = lambda: []
key = _mod("key")
key.get_pressed = lambda: {}
image = _mod("image")
image.save = lambda *a: None
sprite = _mod("sprite")
sprite.Sprite = type("Sprite", (), {})

_sdl2 = _mod("_sdl2")
for _n in ("Window", "Texture", "Renderer"):
    setattr(_sdl2, _n, type(_n, (), {}))

from . import math  # noqa: E402  (pygame.math.Vector2)


def init():
    return (5, 0)


def quit():
    return None
