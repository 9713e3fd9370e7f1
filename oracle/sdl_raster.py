"""CPU restatement of the pygame 2.4.0 / SDL2 software-raster arithmetic used by Memory Gym.

TEST INFRASTRUCTURE ONLY. Nothing under ``oracle/`` is imported by the product package
(``endless-memory-gym_b200/``); only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, as the checker.

The reference (``/root/reference/memory_gym``) does all of its pixel work through the third-party
wheel ``pygame==2.4.0`` (``/root/reference/setup.py:35``), which is NOT vendored, not installed in
this image and not installable (no network).  This module therefore restates the published
algorithms of pygame 2.4.0's ``src_c/draw.c``, ``transform.c``, ``surface.c`` and of SDL2's
``SDL_blit_A.c`` / colour-key blit for exactly the call sites the reference has:

====================================  ==============================================================
reference call site                   primitive restated here
====================================  ==============================================================
character_controller.py:58,67-70      ``circle`` (filled midpoint circle, ``draw_circle_filled``)
pygame_assets.py:275                  ``circle`` width>1 (``draw_circle_bresenham``, pygame>=2.1.3)
pygame_assets.py:276,283-285          ``line`` width>1 (``draw_line_width``)
mystery_path.py:180-181               ``line`` width 3
pygame_assets.py:303                  ``rotate`` (``transform.rotate``: rotate90 / 16.16 fixed point)
pygame_assets.py:328-329,342-343      ``rect`` fill + 1-px frame (``draw_rect``)
pygame_assets.py:201-204              ``rect`` rounded (``draw_round_rect`` + ``draw_fillpoly`` +
                                      ``draw_circle_quadrant``)
mortar_mayhem.py:102 (every blit)     ``blit`` opaque / colour-key / colour-key + surface alpha
mortar_mayhem.py:270                  ``array3d`` ([x][y][c] layout)
====================================  ==============================================================

PARITY STATUS: the primitives are pinned against real PyGame renders (``docs/assets/*.gif`` of
the reference, see ``tests/golden/kat_*.npz`` + ``tests/test_raster_kat.py``).  One primitive is
"parity unpinned": the thick circle (``circle`` with 1 < width < radius, used only by the "stay"
command icon) changed algorithm in pygame 2.1.3; the GIFs pre-date that release, so the >=2.1.3
variant implemented as the default here is restated from the published source and is not covered
by a real-PyGame known answer (``thick_circle_legacy`` reproduces the GIF and is kept for the KAT).

Pixels are SDL "mapped" 32-bit XRGB8888 values ``0x00RRGGBB`` held in ``numpy.uint32[h, w]``.
All coordinates are Python ints; ``cint`` is the C ``(int)double`` cast pygame applies to float
arguments (``pg_IntFromObj``).
"""
from __future__ import annotations

import math

import numpy as np


def cint(v) -> int:
    """C-style ``(int)`` cast of a float/int argument (truncation toward zero)."""
    return int(v)


def map_rgb(color) -> int:
    """Colour argument -> mapped pixel. Ints are raw mapped pixels (``fill(255)`` is blue)."""
    if isinstance(color, (int, np.integer)):
        return int(color) & 0xFFFFFFFF
    r, g, b = (int(c) for c in color[:3])
    return (r << 16) | (g << 8) | b


class Surface:
    """A 32-bit software surface without per-pixel alpha (what ``pygame.Surface((w, h))`` is
    after ``display.set_mode`` under the dummy video driver)."""

    __slots__ = ("w", "h", "px", "colorkey", "alpha")

    def __init__(self, w, h):
        self.w, self.h = cint(w), cint(h)
        self.px = np.zeros((self.h, self.w), dtype=np.uint32)
        self.colorkey = None  # mapped pixel or None
        self.alpha = None  # None (no blend mode) or 0..255

    def copy(self):
        s = Surface(self.w, self.h)
        s.px[...] = self.px
        s.colorkey, s.alpha = self.colorkey, self.alpha
        return s


# ------------------------------------------------------------------------------------------------
# pixel / span helpers (clipped to the surface = SDL clip rect)
# ------------------------------------------------------------------------------------------------
def set_px(s: Surface, x: int, y: int, px: int) -> None:
    if 0 <= x < s.w and 0 <= y < s.h:
        s.px[y, x] = px


def hline(s: Surface, px: int, x1: int, y: int, x2: int) -> None:
    """``drawhorzlineclip``: inclusive span, swapped if reversed, clipped."""
    if y < 0 or y >= s.h:
        return
    if x2 < x1:
        x1, x2 = x2, x1
    x1 = max(x1, 0)
    x2 = min(x2, s.w - 1)
    if x2 < 0 or x1 >= s.w or x1 > x2:
        return
    s.px[y, x1:x2 + 1] = px


def vline(s: Surface, px: int, x: int, y1: int, y2: int) -> None:
    if x < 0 or x >= s.w:
        return
    if y2 < y1:
        y1, y2 = y2, y1
    y1 = max(y1, 0)
    y2 = min(y2, s.h - 1)
    if y1 > y2:
        return
    s.px[y1:y2 + 1, x] = px


def fill(s: Surface, color) -> None:
    s.px[...] = map_rgb(color)


def fill_rect(s: Surface, px: int, x: int, y: int, w: int, h: int) -> None:
    """``SDL_FillRect`` of the rect intersected with the surface."""
    x0, y0 = max(x, 0), max(y, 0)
    x1, y1 = min(x + w, s.w), min(y + h, s.h)
    if x1 > x0 and y1 > y0:
        s.px[y0:y1, x0:x1] = px


# ------------------------------------------------------------------------------------------------
# circles  (pygame 2.4.0 draw.c: circle(), draw_circle_filled, draw_circle_bresenham[_thin])
# ------------------------------------------------------------------------------------------------
def _circle_filled(s, px, x0, y0, radius):
    f = 1 - radius
    ddx = 0
    ddy = -2 * radius
    x, y = 0, radius
    while x < y:
        if f >= 0:
            y -= 1
            ddy += 2
            f += ddy
        x += 1
        ddx += 2
        f += ddx + 1
        if f >= 0:
            hline(s, px, x0 - x, y0 + y - 1, x0 + x - 1)
            hline(s, px, x0 - x, y0 - y, x0 + x - 1)
        hline(s, px, x0 - y, y0 + x - 1, x0 + y - 1)
        hline(s, px, x0 - y, y0 - x, x0 + y - 1)


def _circle_thin(s, px, x0, y0, radius):
    f = 1 - radius
    ddx = 0
    ddy = -2 * radius
    x, y = 0, radius
    while x < y:
        if f >= 0:
            y -= 1
            ddy += 2
            f += ddy
        x += 1
        ddx += 2
        f += ddx + 1
        set_px(s, x0 + x - 1, y0 + y - 1, px)
        set_px(s, x0 - x, y0 + y - 1, px)
        set_px(s, x0 + x - 1, y0 - y, px)
        set_px(s, x0 - x, y0 - y, px)
        set_px(s, x0 + y - 1, y0 + x - 1, px)
        set_px(s, x0 + y - 1, y0 - x, px)
        set_px(s, x0 - y, y0 + x - 1, px)
        set_px(s, x0 - y, y0 - x, px)


def _circle_thick(s, px, x0, y0, radius, thickness):
    """``draw_circle_bresenham`` of pygame >= 2.1.3 (outer/inner midpoint scan). Parity unpinned."""
    x, y = 0, radius
    r2 = radius * radius
    d_r2 = 2 * r2
    d1 = r2 * (1.25 - radius)
    dx = 0
    dy = d_r2 * y
    line = True
    ri = radius - thickness + 1
    xi, yi = 0, ri
    q2 = ri * ri
    d_q2 = 2 * q2
    d1i = q2 * (1.25 - ri)
    d2i = 0.0
    dxi = 0
    dyi = d_q2 * yi

    def spans():
        if line:
            hline(s, px, x0 - x, y0 - y, x0 + x - 1)
            hline(s, px, x0 - x, y0 + y - 1, x0 + x - 1)
        else:
            hline(s, px, x0 - x, y0 - y, x0 - xi)
            hline(s, px, x0 - x, y0 + y - 1, x0 - xi)
            hline(s, px, x0 + xi - 1, y0 - y, x0 + x - 1)
            hline(s, px, x0 + xi - 1, y0 + y - 1, x0 + x - 1)

    while dx < dy:
        while d1 < 0:
            x += 1
            dx += d_r2
            d1 += dx + r2
        spans()
        x += 1
        y -= 1
        dx += d_r2
        dy -= d_r2
        d1 += dx - dy + r2
        if line and y < ri:
            line = False
        if not line:
            while d1i < 0:
                xi += 1
                dxi += d_q2
                d1i += dxi + q2
            xi += 1
            yi -= 1
            dxi += d_q2
            dyi -= d_q2
            d1i += dxi - dyi + q2
    d1 = r2 * (x + 0.5) * (x + 0.5) + r2 * (y - 1) * (y - 1) - r2 * r2
    while y >= 0:
        spans()
        if d1 > 0:
            y -= 1
            dy -= d_r2
            d1 += r2 - dy
        else:
            y -= 1
            x += 1
            dx += d_r2
            dy -= d_r2
            d1 += dx - dy + r2
        if line and y < ri:
            line = False
        if not line:
            if dxi < dyi:
                while d1i < 0:
                    xi += 1
                    dxi += d_q2
                    d1i += dxi + q2
                xi += 1
                yi -= 1
                dxi += d_q2
                dyi -= d_q2
                d1i += dxi - dyi + q2
            else:
                if d2i == 0:
                    d2i = q2 * (xi + 0.5) * (xi + 0.5) + q2 * (yi - 1) * (yi - 1) - q2 * q2
                if d2i > 0:
                    yi -= 1
                    dyi -= d_q2
                    d2i += q2 - dyi
                else:
                    yi -= 1
                    xi += 1
                    dxi += d_q2
                    dyi -= d_q2
                    d2i += dxi - dyi + q2


def thick_circle_legacy(s, px, x0, y0, radius, thickness):
    """Thick circle of pygame 2.0.0 .. 2.1.2 (concentric Bresenham). Reproduces the docs GIFs."""
    f = 1 - radius
    ddx = 0
    ddy = -2 * radius
    x, y = 0, radius
    iy = radius - thickness
    jf = 1 - iy
    jdx = 0
    jdy = -2 * iy
    for i in range(thickness):
        y1 = radius - i
        for (a, b) in ((x0, y0 - y1), (x0 - 1, y0 - y1), (x0, y0 + y1 - 1), (x0 - 1, y0 + y1 - 1),
                       (x0 - y1, y0), (x0 - y1, y0 - 1), (x0 + y1 - 1, y0), (x0 + y1 - 1, y0 - 1)):
            set_px(s, a, b, px)
    while x < y:
        if f >= 0:
            y -= 1
            ddy += 2
            f += ddy
        if jf >= 0:
            iy -= 1
            jdy += 2
            jf += jdy
        x += 1
        ddx += 2
        f += ddx + 1
        jdx += 2
        jf += jdx + 1
        if thickness > 1:
            thickness = y - iy
        for i in range(thickness):
            y1 = y - i
            for (a, b) in ((x0 + x - 1, y0 - y1), (x0 - x, y0 - y1), (x0 + x - 1, y0 + y1 - 1),
                           (x0 - x, y0 + y1 - 1), (x0 + y1 - 1, y0 - x), (x0 - y1, y0 - x),
                           (x0 + y1 - 1, y0 + x - 1), (x0 - y1, y0 + x - 1)):
                set_px(s, a, b, px)


def circle(s: Surface, color, center, radius, width=0, legacy_thick=False) -> None:
    """``pygame.draw.circle`` (no quadrant flags)."""
    px = map_rgb(color)
    x0, y0 = cint(center[0]), cint(center[1])
    radius, width = cint(radius), cint(width)
    if radius < 1 or width < 0:
        return
    if width > radius:
        width = radius
    if x0 > s.w + radius or x0 < -radius or y0 > s.h + radius or y0 < -radius:
        return
    if width == 0 or width == radius:
        _circle_filled(s, px, x0, y0, radius)
    elif width == 1:
        _circle_thin(s, px, x0, y0, radius)
    elif legacy_thick:
        thick_circle_legacy(s, px, x0, y0, radius, width)
    else:
        _circle_thick(s, px, x0, y0, radius, width)


def disc_rows(radius: int):
    """Row spans of the filled circle relative to its centre: {dy: (dx_min, dx_max)}.
    Derived by running the primitive itself so tables never drift from it."""
    s = Surface(2 * radius + 4, 2 * radius + 4)
    c = radius + 2
    circle(s, 1, (c, c), radius)
    rows = {}
    for y in range(s.h):
        xs = np.nonzero(s.px[y])[0]
        if xs.size:
            rows[y - c] = (int(xs[0]) - c, int(xs[-1]) - c)
    return rows


# ------------------------------------------------------------------------------------------------
# lines  (draw_line, draw_line_width)
# ------------------------------------------------------------------------------------------------
def _line1(s, px, x1, y1, x2, y2):
    if x1 == x2 and y1 == y2:
        set_px(s, x1, y1, px)
        return
    if y1 == y2:
        d = 1 if x1 < x2 else -1
        for k in range(abs(x1 - x2) + 1):
            set_px(s, x1 + d * k, y1, px)
        return
    if x1 == x2:
        d = 1 if y1 < y2 else -1
        for k in range(abs(y1 - y2) + 1):
            set_px(s, x1, y1 + d * k, px)
        return
    dx, sx = abs(x2 - x1), (1 if x1 < x2 else -1)
    dy, sy = abs(y2 - y1), (1 if y1 < y2 else -1)
    err = _cdiv(dx if dx > dy else -dy, 2)
    while x1 != x2 or y1 != y2:
        set_px(s, x1, y1, px)
        e2 = err
        if e2 > -dx:
            err -= dy
            x1 += sx
        if e2 < dy:
            err += dx
            y1 += sy
    set_px(s, x2, y2, px)


def _cdiv(a: int, b: int) -> int:
    """C integer division (truncation toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def line(s: Surface, color, start, end, width=1) -> None:
    """``pygame.draw.line``; float endpoints are C-cast like pygame 2.4.0's ``pg_TwoIntsFromObj``."""
    px = map_rgb(color)
    x1, y1 = cint(start[0]), cint(start[1])
    x2, y2 = cint(end[0]), cint(end[1])
    width = cint(width)
    if width < 1:
        return
    if width == 1:
        _line1(s, px, x1, y1, x2, y2)
        return
    extra = 1 - (width % 2)
    w2 = width // 2
    thick_in_x = abs(x1 - x2) <= abs(y1 - y2)
    dx, sx = abs(x2 - x1), (1 if x1 < x2 else -1)
    dy, sy = abs(y2 - y1), (1 if y1 < y2 else -1)
    err = _cdiv(dx if dx > dy else -dy, 2)
    if thick_in_x:
        while y1 != y2 + sy:
            if 0 <= y1 < s.h:
                a, b = max(x1 - w2 + extra, 0), min(s.w - 1, x1 + w2)
                if a <= b:
                    s.px[y1, a:b + 1] = px
            e2 = err
            if e2 > -dx:
                err -= dy
                x1 += sx
            if e2 < dy:
                err += dx
                y1 += sy
    else:
        while x1 != x2 + sx:
            if 0 <= x1 < s.w:
                a, b = max(y1 - w2 + extra, 0), min(s.h - 1, y1 + w2)
                if a <= b:
                    s.px[a:b + 1, x1] = px
            e2 = err
            if e2 > -dx:
                err -= dy
                x1 += sx
            if e2 < dy:
                err += dx
                y1 += sy


# ------------------------------------------------------------------------------------------------
# rects  (rect(), draw_rect, draw_round_rect, draw_fillpoly, draw_circle_quadrant)
# ------------------------------------------------------------------------------------------------
def _fillpoly(s, px, xs, ys):
    n = len(xs)
    miny, maxy = min(ys), max(ys)
    if miny == maxy:
        hline(s, px, min(xs), miny, max(xs))
        return
    for y in range(miny, maxy + 1):
        hits = []
        for i in range(n):
            ip = i - 1 if i else n - 1
            y1, y2 = ys[ip], ys[i]
            if y1 < y2:
                x1, x2 = xs[ip], xs[i]
            elif y1 > y2:
                y2, y1 = ys[ip], ys[i]
                x2, x1 = xs[ip], xs[i]
            else:
                continue
            if (y1 <= y < y2) or (y == maxy and y2 == maxy):
                hits.append(_cdiv((y - y1) * (x2 - x1), (y2 - y1)) + x1)
        hits.sort()
        for k in range(0, len(hits) - 1, 2):
            hline(s, px, hits[k], y, hits[k + 1])
    for i in range(n):
        ip = i - 1 if i else n - 1
        y = ys[i]
        if miny < y and ys[ip] == y and y < maxy:
            hline(s, px, xs[i], y, xs[ip])


def _quadrant(s, px, x0, y0, radius, thickness, tr, tl, bl, br):
    f = 1 - radius
    ddx = 0
    ddy = -2 * radius
    x, y = 0, radius
    iy = radius - thickness
    jf = 1 - iy
    jdx = 0
    jdy = -2 * iy
    if radius == 1:
        if tr:
            set_px(s, x0, y0 - 1, px)
        if tl:
            set_px(s, x0 - 1, y0 - 1, px)
        if bl:
            set_px(s, x0 - 1, y0, px)
        if br:
            set_px(s, x0, y0, px)
        return
    if thickness != 0:
        while x < y:
            if f >= 0:
                y -= 1
                ddy += 2
                f += ddy
            if jf >= 0:
                iy -= 1
                jdy += 2
                jf += jdy
            x += 1
            ddx += 2
            f += ddx + 1
            jdx += 2
            jf += jdx + 1
            if thickness > 1:
                thickness = y - iy
            for i in range(thickness):
                y1 = y - i
                if tr:
                    set_px(s, x0 + x - 1, y0 - y1, px)
                    set_px(s, x0 + y1 - 1, y0 - x, px)
                if tl:
                    set_px(s, x0 - x, y0 - y1, px)
                    set_px(s, x0 - y1, y0 - x, px)
                if bl:
                    set_px(s, x0 - x, y0 + y1 - 1, px)
                    set_px(s, x0 - y1, y0 + x - 1, px)
                if br:
                    set_px(s, x0 + x - 1, y0 + y1 - 1, px)
                    set_px(s, x0 + y1 - 1, y0 + x - 1, px)
    else:
        while x < y:
            if f >= 0:
                y -= 1
                ddy += 2
                f += ddy
            x += 1
            ddx += 2
            f += ddx + 1
            if tr:
                for yy in range(y0 - x, y0 + 1):
                    set_px(s, x0 + y - 1, yy, px)
                for yy in range(y0 - y, y0 + 1):
                    set_px(s, x0 + x - 1, yy, px)
            if tl:
                for yy in range(y0 - x, y0 + 1):
                    set_px(s, x0 - y, yy, px)
                for yy in range(y0 - y, y0 + 1):
                    set_px(s, x0 - x, yy, px)
            if bl:
                for yy in range(y0, y0 + x):
                    set_px(s, x0 - y, yy, px)
                for yy in range(y0, y0 + y):
                    set_px(s, x0 - x, yy, px)
            if br:
                for yy in range(y0, y0 + x):
                    set_px(s, x0 + y - 1, yy, px)
                for yy in range(y0, y0 + y):
                    set_px(s, x0 + x - 1, yy, px)


def _round_rect(s, px, x1, y1, x2, y2, radius, width, tl, tr, bl, br):
    if tl < 0:
        tl = radius
    if tr < 0:
        tr = radius
    if bl < 0:
        bl = radius
    if br < 0:
        br = radius
    W, H = x2 - x1 + 1, y2 - y1 + 1
    if tl + tr > W or bl + br > W or tl + bl > H or tr + br > H:
        f32 = np.float32
        with np.errstate(divide="ignore"):
            q = min(f32(W) / f32(tl + tr), f32(H) / f32(tl + bl), f32(W) / f32(bl + br), f32(H) / f32(tr + br))
        tl, tr, bl, br = (int(f32(c) * q) for c in (tl, tr, bl, br))
    if width == 0:
        xs = [x1, x1 + tl, x2 - tr, x2, x2, x2 - br, x1 + bl, x1]
        ys = [y1 + tl, y1, y1, y1 + tr, y2 - br, y2, y2, y2 - bl]
        _fillpoly(s, px, xs, ys)
        _quadrant(s, px, x2 - tr + 1, y1 + tr, tr, 0, 1, 0, 0, 0)
        _quadrant(s, px, x1 + tl, y1 + tl, tl, 0, 0, 1, 0, 0)
        _quadrant(s, px, x1 + bl, y2 - bl + 1, bl, 0, 0, 0, 1, 0)
        _quadrant(s, px, x2 - br + 1, y2 - br + 1, br, 0, 0, 0, 0, 1)
        return
    o = width // 2 - 1 + width % 2
    if x2 - tr == x1 + tl:
        for i in range(width):
            set_px(s, x1 + tl, y1 + i, px)
    else:
        line(s, px, (x1 + tl, y1 + o), (x2 - tr, y1 + o), width)
    if y2 - bl == y1 + tl:
        for i in range(width):
            set_px(s, x1 + i, y1 + tl, px)
    else:
        line(s, px, (x1 + o, y1 + tl), (x1 + o, y2 - bl), width)
    if x2 - br == x1 + bl:
        for i in range(width):
            set_px(s, x1 + bl, y2 - i, px)
    else:
        line(s, px, (x1 + bl, y2 - width // 2), (x2 - br, y2 - width // 2), width)
    if y2 - br == y1 + tr:
        for i in range(width):
            set_px(s, x2 - i, y1 + tr, px)
    else:
        line(s, px, (x2 - width // 2, y1 + tr), (x2 - width // 2, y2 - br), width)
    _quadrant(s, px, x2 - tr + 1, y1 + tr, tr, width, 1, 0, 0, 0)
    _quadrant(s, px, x1 + tl, y1 + tl, tl, width, 0, 1, 0, 0)
    _quadrant(s, px, x1 + bl, y2 - bl + 1, bl, width, 0, 0, 1, 0)
    _quadrant(s, px, x2 - br + 1, y2 - br + 1, br, width, 0, 0, 0, 1)


def rect(s: Surface, color, r, width=0, border_radius=0, tl=-1, tr=-1, bl=-1, br=-1) -> None:
    """``pygame.draw.rect``; ``r`` = (x, y, w, h), floats C-cast."""
    px = map_rgb(color)
    x, y, w, h = (cint(v) for v in r)
    width = cint(width)
    if width < 0:
        return
    if (border_radius <= 0 and tl <= 0 and tr <= 0 and bl <= 0 and br <= 0) or abs(w) < 2 or abs(h) < 2:
        # SDL_IntersectRect with the clip rect; empty (or negative) rects draw nothing
        cx0, cy0 = max(x, 0), max(y, 0)
        cx1, cy1 = min(x + w, s.w), min(y + h, s.h)
        cw, ch = cx1 - cx0, cy1 - cy0
        if w <= 0 or h <= 0 or cw <= 0 or ch <= 0:
            return
        if width > 0 and width * 2 < cw and width * 2 < ch:
            x2, y2 = x + w - 1, y + h - 1
            for i in range(width):
                hline(s, px, x, y + i, x2)
                hline(s, px, x, y2 - i, x2)
            for i in range((y2 - y) - 2 * width + 1):
                hline(s, px, x, y + width + i, x + width - 1)
                hline(s, px, x2 - width + 1, y + width + i, x2)
        else:
            s.px[cy0:cy1, cx0:cx1] = px
        return
    if w < 0:
        x += w
        w = -w
    if h < 0:
        y += h
        h = -h
    if width > w // 2 or width > h // 2:
        width = max(w // 2, h // 2)
    _round_rect(s, px, x, y, x + w - 1, y + h - 1, border_radius, width, tl, tr, bl, br)


# ------------------------------------------------------------------------------------------------
# transform.rotate  (transform.c: surf_rotate, rotate90, rotate)
# ------------------------------------------------------------------------------------------------
def _c_i32(v: float) -> int:
    """(int) cast of a double known to be in int32 range."""
    return int(v)


def rotate(src: Surface, angle) -> Surface:
    """``pygame.transform.rotate`` for 32-bit surfaces; the result keeps colour-key and alpha."""
    angle = float(angle)
    if math.fmod(angle, 90.0) == 0.0:
        turns = int(angle / 90) % 4 if angle >= 0 else -(int(-angle / 90) % 4)
        if turns < 0:
            turns += 4
        if turns == 0:
            out = src.copy()
            if src.colorkey is not None:
                # rotate90 blits src onto a zeroed surface with the colour key active
                out.px[src.px == src.colorkey] = 0
            return out
        # numpy.rot90 on [y][x] arrays is the same counter-clockwise quarter turn:
        # case 1 of rotate90: dst(x, y) = src(w-1-y, x)
        out = Surface(src.h if turns % 2 else src.w, src.w if turns % 2 else src.h)
        out.px[...] = np.rot90(src.px, turns)
        out.colorkey, out.alpha = src.colorkey, src.alpha
        return out
    rad = angle * .01745329251994329
    sa, ca = math.sin(rad), math.cos(rad)
    x, y = src.w, src.h
    cx, cy, sx, sy = ca * x, ca * y, sa * x, sa * y
    nx = int(max(abs(cx + sy), abs(cx - sy), abs(-cx + sy), abs(-cx - sy)))
    ny = int(max(abs(sx + cy), abs(sx - cy), abs(-sx + cy), abs(-sx - cy)))
    dst = Surface(nx, ny)
    dst.colorkey, dst.alpha = src.colorkey, src.alpha
    bg = src.colorkey if src.colorkey is not None else int(src.px[0, 0])
    half = dst.h // 2
    xd = (src.w - dst.w) << 15
    yd = (src.h - dst.h) << 15
    isin = _c_i32(sa * 65536)
    icos = _c_i32(ca * 65536)
    ax = (dst.w << 15) - _c_i32(ca * ((dst.w - 1) << 15))
    ay = (dst.h << 15) - _c_i32(sa * ((dst.w - 1) << 15))
    xmax = (src.w << 16) - 1
    ymax = (src.h << 16) - 1
    for yy in range(dst.h):
        dx = (ax + isin * (half - yy)) + xd
        dy = (ay - icos * (half - yy)) + yd
        for xx in range(dst.w):
            if dx < 0 or dy < 0 or dx > xmax or dy > ymax:
                dst.px[yy, xx] = bg
            else:
                dst.px[yy, xx] = src.px[dy >> 16, dx >> 16]
            dx += icos
            dy += isin
    return dst


# ------------------------------------------------------------------------------------------------
# blit  (SDL_BlitSurface: opaque copy, colour key, colour key + per-surface alpha)
# ------------------------------------------------------------------------------------------------
def blit(src: Surface, dst: Surface, pos) -> None:
    """``dst.blit(src, pos)``: ``pos`` is the top-left (floats C-cast, as Rect/tuple coercion does)."""
    dx, dy = cint(pos[0]), cint(pos[1])
    x0, y0 = max(dx, 0), max(dy, 0)
    x1, y1 = min(dx + src.w, dst.w), min(dy + src.h, dst.h)
    if x1 <= x0 or y1 <= y0:
        return
    sv = src.px[y0 - dy:y1 - dy, x0 - dx:x1 - dx]
    dv = dst.px[y0:y1, x0:x1]
    a = src.alpha
    if a is None or a == 255:
        if src.colorkey is None:
            dv[...] = sv
        else:
            m = sv != src.colorkey
            dv[m] = sv[m]
        return
    if a == 0:
        return
    # BlitNtoNSurfaceAlpha[Key]: per channel d = (Uint8)(((s - d) * A) / 255 + d), C division
    m = np.ones(sv.shape, dtype=bool) if src.colorkey is None else (sv != src.colorkey)
    out = dv.copy()
    res = np.zeros(sv.shape, dtype=np.uint32)
    for shift in (16, 8, 0):
        sc = ((sv >> shift) & 0xFF).astype(np.int64)
        dc = ((dv >> shift) & 0xFF).astype(np.int64)
        num = (sc - dc) * a
        q = np.where(num >= 0, num // 255, -((-num) // 255))
        res |= ((q + dc) & 0xFF).astype(np.uint32) << shift
    out[m] = res[m]
    dv[...] = out


def array3d(s: Surface) -> np.ndarray:
    """``pygame.surfarray.array3d``: uint8 ``[x][y][c]``."""
    out = np.empty((s.w, s.h, 3), dtype=np.uint8)
    t = s.px.T
    out[..., 0] = (t >> 16) & 0xFF
    out[..., 1] = (t >> 8) & 0xFF
    out[..., 2] = t & 0xFF
    return out


# ------------------------------------------------------------------------------------------------
# pygame.math.Vector2 arithmetic used by the reference (math.c)
# ------------------------------------------------------------------------------------------------
def vec_rotate(x: float, y: float, angle: float):
    """``Vector2.rotate`` / ``rotate_ip`` (degrees, epsilon 1e-6 special-casing of quarter turns)."""
    eps = 1e-6
    angle = math.fmod(angle, 360.0)
    if angle < 0:
        angle += 360.0
    if math.fmod(angle + eps, 90.0) < 2 * eps:
        q = int((angle + eps) / 90)
        if q in (0, 4):
            return (x, y)
        if q == 1:
            return (-y, x)
        if q == 2:
            return (-x, -y)
        return (y, -x)
    rad = angle * math.pi / 180.0
    sv, cv = math.sin(rad), math.cos(rad)
    return (cv * x - sv * y, sv * x + cv * y)
