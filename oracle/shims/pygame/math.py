"""pygame.math.Vector2 subset (doubles; restates pygame 2.4.0 src_c/math.c for the used methods)."""
import math as _m

from oracle import sdl_raster as _R


class Vector2:
    __slots__ = ("x", "y")

    def __init__(self, x=0.0, y=None):
        if y is None:
            if isinstance(x, (int, float)):
                x, y = x, x
            else:
                x, y = x[0], x[1]
        self.x, self.y = float(x), float(y)

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return (self.x, self.y)[i]

    def __iter__(self):
        return iter((self.x, self.y))

    def __add__(self, o):
        return Vector2(self.x + float(o[0]), self.y + float(o[1]))

    __radd__ = __add__

    def __sub__(self, o):
        return Vector2(self.x - float(o[0]), self.y - float(o[1]))

    def __rsub__(self, o):
        return Vector2(float(o[0]) - self.x, float(o[1]) - self.y)

    def __mul__(self, k):
        return Vector2(self.x * float(k), self.y * float(k))

    __rmul__ = __mul__

    def __eq__(self, o):
        try:
            return abs(self.x - o[0]) < 1e-6 and abs(self.y - o[1]) < 1e-6
        except Exception:
            return False

    def length(self):
        return _m.sqrt(self.x * self.x + self.y * self.y)

    def normalize(self):
        n = _m.sqrt(self.x * self.x + self.y * self.y)
        if n == 0:
            raise ValueError("Can't normalize Vector of length Zero")
        return Vector2(self.x / n, self.y / n)

    def rotate(self, angle):
        return Vector2(*_R.vec_rotate(self.x, self.y, float(angle)))

    def rotate_ip(self, angle):
        self.x, self.y = _R.vec_rotate(self.x, self.y, float(angle))

    def lerp(self, other, t):
        t = float(t)
        if t < 0 or t > 1:
            raise ValueError("Argument 3 must be in range [0, 1]")
        return Vector2(self.x * (1 - t) + float(other[0]) * t, self.y * (1 - t) + float(other[1]) * t)

    def distance_to(self, other):
        dx = float(other[0]) - self.x
        dy = float(other[1]) - self.y
        return _m.sqrt(dx * dx + dy * dy)

    def __repr__(self):
        return f"<Vector2({self.x}, {self.y})>"
