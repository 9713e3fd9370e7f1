// mg_raster_host.cpp — see mg_raster_host.h. Integer scan conversion restated from pygame 2.4.0's
// published draw.c / transform.c (third-party dependency of the reference, setup.py:35).
#include "mg_raster_host.h"

#include <math.h>
#include <stdlib.h>

#include <algorithm>

namespace mgr {

static inline void set_px(Surface& s, int x, int y, uint32_t px) {
    if (x >= 0 && x < s.w && y >= 0 && y < s.h) s.at(x, y) = px;
}

// drawhorzlineclip: inclusive span, swapped when reversed, clipped to the surface
static void hline(Surface& s, uint32_t px, int x1, int y, int x2) {
    if (y < 0 || y >= s.h) return;
    if (x2 < x1) std::swap(x1, x2);
    x1 = std::max(x1, 0);
    x2 = std::min(x2, s.w - 1);
    for (int x = x1; x <= x2; ++x) s.at(x, y) = px;
}

void fill(Surface& s, uint32_t px) { std::fill(s.px.begin(), s.px.end(), px); }

// ---- circles ----------------------------------------------------------------------------------
static void circle_filled(Surface& s, uint32_t px, int x0, int y0, int radius) {
    int f = 1 - radius, ddx = 0, ddy = -2 * radius, x = 0, y = radius;
    while (x < y) {
        if (f >= 0) {
            y--;
            ddy += 2;
            f += ddy;
        }
        x++;
        ddx += 2;
        f += ddx + 1;
        if (f >= 0) {
            hline(s, px, x0 - x, y0 + y - 1, x0 + x - 1);
            hline(s, px, x0 - x, y0 - y, x0 + x - 1);
        }
        hline(s, px, x0 - y, y0 + x - 1, x0 + y - 1);
        hline(s, px, x0 - y, y0 - x, x0 + y - 1);
    }
}

static void circle_thin(Surface& s, uint32_t px, int x0, int y0, int radius) {
    int f = 1 - radius, ddx = 0, ddy = -2 * radius, x = 0, y = radius;
    while (x < y) {
        if (f >= 0) {
            y--;
            ddy += 2;
            f += ddy;
        }
        x++;
        ddx += 2;
        f += ddx + 1;
        set_px(s, x0 + x - 1, y0 + y - 1, px);
        set_px(s, x0 - x, y0 + y - 1, px);
        set_px(s, x0 + x - 1, y0 - y, px);
        set_px(s, x0 - x, y0 - y, px);
        set_px(s, x0 + y - 1, y0 + x - 1, px);
        set_px(s, x0 + y - 1, y0 - x, px);
        set_px(s, x0 - y, y0 + x - 1, px);
        set_px(s, x0 - y, y0 - x, px);
    }
}

// ring of pygame >= 2.1.3: outer and inner midpoint ellipses scanned together, span per row
namespace {
struct Ring {
    Surface& s;
    uint32_t px;
    int x0, y0;
    long long x, y, xi, yi;
    bool solid;
    void spans() {
        if (solid) {
            hline(s, px, (int)(x0 - x), (int)(y0 - y), (int)(x0 + x - 1));
            hline(s, px, (int)(x0 - x), (int)(y0 + y - 1), (int)(x0 + x - 1));
        } else {
            hline(s, px, (int)(x0 - x), (int)(y0 - y), (int)(x0 - xi));
            hline(s, px, (int)(x0 - x), (int)(y0 + y - 1), (int)(x0 - xi));
            hline(s, px, (int)(x0 + xi - 1), (int)(y0 - y), (int)(x0 + x - 1));
            hline(s, px, (int)(x0 + xi - 1), (int)(y0 + y - 1), (int)(x0 + x - 1));
        }
    }
};
}  // namespace

static void circle_thick(Surface& s, uint32_t px, int x0, int y0, int radius, int thickness) {
    Ring g{s, px, x0, y0, 0, radius, 0, 0, true};
    long long r2 = (long long)radius * radius, d_r2 = 2 * r2;
    double d1 = r2 * (1.25 - radius);
    long long dx = 0, dy = d_r2 * g.y;
    long long ri = radius - thickness + 1;
    g.yi = ri;
    long long q2 = ri * ri, d_q2 = 2 * q2;
    double d1i = q2 * (1.25 - ri), d2i = 0.0;
    long long dxi = 0, dyi = d_q2 * g.yi;
    auto inner_region1 = [&]() {
        while (d1i < 0) {
            g.xi++;
            dxi += d_q2;
            d1i += dxi + q2;
        }
        g.xi++;
        g.yi--;
        dxi += d_q2;
        dyi -= d_q2;
        d1i += dxi - dyi + q2;
    };
    while (dx < dy) {
        while (d1 < 0) {
            g.x++;
            dx += d_r2;
            d1 += dx + r2;
        }
        g.spans();
        g.x++;
        g.y--;
        dx += d_r2;
        dy -= d_r2;
        d1 += dx - dy + r2;
        if (g.solid && g.y < ri) g.solid = false;
        if (!g.solid) inner_region1();
    }
    d1 = r2 * (g.x + 0.5) * (g.x + 0.5) + r2 * (g.y - 1) * (g.y - 1) - (double)r2 * r2;
    while (g.y >= 0) {
        g.spans();
        if (d1 > 0) {
            g.y--;
            dy -= d_r2;
            d1 += r2 - dy;
        } else {
            g.y--;
            g.x++;
            dx += d_r2;
            dy -= d_r2;
            d1 += dx - dy + r2;
        }
        if (g.solid && g.y < ri) g.solid = false;
        if (!g.solid) {
            if (dxi < dyi) {
                inner_region1();
            } else {
                if (d2i == 0) d2i = q2 * (g.xi + 0.5) * (g.xi + 0.5) + q2 * (g.yi - 1) * (g.yi - 1) - (double)q2 * q2;
                if (d2i > 0) {
                    g.yi--;
                    dyi -= d_q2;
                    d2i += q2 - dyi;
                } else {
                    g.yi--;
                    g.xi++;
                    dxi += d_q2;
                    dyi -= d_q2;
                    d2i += dxi - dyi + q2;
                }
            }
        }
    }
}

// ring of pygame 2.0.0 .. 2.1.2: two concentric Bresenham circles, 8-way symmetric set_at
static void circle_thick_legacy(Surface& s, uint32_t px, int x0, int y0, int radius, int thickness) {
    int f = 1 - radius, ddx = 0, ddy = -2 * radius, x = 0, y = radius;
    int iy = radius - thickness, jf = 1 - iy, jdx = 0, jdy = -2 * iy;
    for (int i = 0; i < thickness; ++i) {
        int y1 = radius - i;
        set_px(s, x0, y0 - y1, px);
        set_px(s, x0 - 1, y0 - y1, px);
        set_px(s, x0, y0 + y1 - 1, px);
        set_px(s, x0 - 1, y0 + y1 - 1, px);
        set_px(s, x0 - y1, y0, px);
        set_px(s, x0 - y1, y0 - 1, px);
        set_px(s, x0 + y1 - 1, y0, px);
        set_px(s, x0 + y1 - 1, y0 - 1, px);
    }
    while (x < y) {
        if (f >= 0) {
            y--;
            ddy += 2;
            f += ddy;
        }
        if (jf >= 0) {
            iy--;
            jdy += 2;
            jf += jdy;
        }
        x++;
        ddx += 2;
        f += ddx + 1;
        jdx += 2;
        jf += jdx + 1;
        if (thickness > 1) thickness = y - iy;
        for (int i = 0; i < thickness; ++i) {
            int y1 = y - i;
            set_px(s, x0 + x - 1, y0 - y1, px);
            set_px(s, x0 - x, y0 - y1, px);
            set_px(s, x0 + x - 1, y0 + y1 - 1, px);
            set_px(s, x0 - x, y0 + y1 - 1, px);
            set_px(s, x0 + y1 - 1, y0 - x, px);
            set_px(s, x0 - y1, y0 - x, px);
            set_px(s, x0 + y1 - 1, y0 + x - 1, px);
            set_px(s, x0 - y1, y0 + x - 1, px);
        }
    }
}

void circle(Surface& s, uint32_t px, int x0, int y0, int radius, int width, bool legacy_thick) {
    if (radius < 1 || width < 0) return;
    if (width > radius) width = radius;
    if (x0 > s.w + radius || x0 < -radius || y0 > s.h + radius || y0 < -radius) return;
    if (width == 0 || width == radius)
        circle_filled(s, px, x0, y0, radius);
    else if (width == 1)
        circle_thin(s, px, x0, y0, radius);
    else if (legacy_thick)
        circle_thick_legacy(s, px, x0, y0, radius, width);
    else
        circle_thick(s, px, x0, y0, radius, width);
}

// ---- lines ------------------------------------------------------------------------------------
static void line1(Surface& s, uint32_t px, int x1, int y1, int x2, int y2) {
    if (x1 == x2 && y1 == y2) {
        set_px(s, x1, y1, px);
        return;
    }
    if (y1 == y2) {
        int d = x1 < x2 ? 1 : -1;
        for (int k = 0; k <= abs(x1 - x2); ++k) set_px(s, x1 + d * k, y1, px);
        return;
    }
    if (x1 == x2) {
        int d = y1 < y2 ? 1 : -1;
        for (int k = 0; k <= abs(y1 - y2); ++k) set_px(s, x1, y1 + d * k, px);
        return;
    }
    int dx = abs(x2 - x1), sx = x1 < x2 ? 1 : -1;
    int dy = abs(y2 - y1), sy = y1 < y2 ? 1 : -1;
    int err = (dx > dy ? dx : -dy) / 2;
    while (x1 != x2 || y1 != y2) {
        set_px(s, x1, y1, px);
        int e2 = err;
        if (e2 > -dx) {
            err -= dy;
            x1 += sx;
        }
        if (e2 < dy) {
            err += dx;
            y1 += sy;
        }
    }
    set_px(s, x2, y2, px);
}

void line(Surface& s, uint32_t px, int x1, int y1, int x2, int y2, int width) {
    if (width < 1) return;
    if (width == 1) {
        line1(s, px, x1, y1, x2, y2);
        return;
    }
    int extra = 1 - (width % 2), w2 = width / 2;
    bool thick_in_x = abs(x1 - x2) <= abs(y1 - y2);
    int dx = abs(x2 - x1), sx = x1 < x2 ? 1 : -1;
    int dy = abs(y2 - y1), sy = y1 < y2 ? 1 : -1;
    int err = (dx > dy ? dx : -dy) / 2;
    if (thick_in_x) {
        while (y1 != y2 + sy) {
            if (y1 >= 0 && y1 < s.h)
                for (int x = std::max(x1 - w2 + extra, 0); x <= std::min(s.w - 1, x1 + w2); ++x) s.at(x, y1) = px;
            int e2 = err;
            if (e2 > -dx) {
                err -= dy;
                x1 += sx;
            }
            if (e2 < dy) {
                err += dx;
                y1 += sy;
            }
        }
    } else {
        while (x1 != x2 + sx) {
            if (x1 >= 0 && x1 < s.w)
                for (int y = std::max(y1 - w2 + extra, 0); y <= std::min(s.h - 1, y1 + w2); ++y) s.at(x1, y) = px;
            int e2 = err;
            if (e2 > -dx) {
                err -= dy;
                x1 += sx;
            }
            if (e2 < dy) {
                err += dx;
                y1 += sy;
            }
        }
    }
}

// ---- rects ------------------------------------------------------------------------------------
static void fillpoly(Surface& s, uint32_t px, const int* xs, const int* ys, int n) {
    int miny = ys[0], maxy = ys[0], minx = xs[0], maxx = xs[0];
    for (int i = 1; i < n; ++i) {
        miny = std::min(miny, ys[i]);
        maxy = std::max(maxy, ys[i]);
        minx = std::min(minx, xs[i]);
        maxx = std::max(maxx, xs[i]);
    }
    if (miny == maxy) {
        hline(s, px, minx, miny, maxx);
        return;
    }
    std::vector<int> hits;
    for (int y = miny; y <= maxy; ++y) {
        hits.clear();
        for (int i = 0; i < n; ++i) {
            int ip = i ? i - 1 : n - 1;
            int y1 = ys[ip], y2 = ys[i], x1, x2;
            if (y1 < y2) {
                x1 = xs[ip];
                x2 = xs[i];
            } else if (y1 > y2) {
                y2 = ys[ip];
                y1 = ys[i];
                x2 = xs[ip];
                x1 = xs[i];
            } else {
                continue;
            }
            if ((y >= y1 && y < y2) || (y == maxy && y2 == maxy)) hits.push_back((y - y1) * (x2 - x1) / (y2 - y1) + x1);
        }
        std::sort(hits.begin(), hits.end());
        for (size_t k = 0; k + 1 < hits.size(); k += 2) hline(s, px, hits[k], y, hits[k + 1]);
    }
    for (int i = 0; i < n; ++i) {
        int ip = i ? i - 1 : n - 1;
        int y = ys[i];
        if (miny < y && ys[ip] == y && y < maxy) hline(s, px, xs[i], y, xs[ip]);
    }
}

static void quadrant(Surface& s, uint32_t px, int x0, int y0, int radius, int thickness, bool tr, bool tl, bool bl,
                     bool br) {
    int f = 1 - radius, ddx = 0, ddy = -2 * radius, x = 0, y = radius;
    int iy = radius - thickness, jf = 1 - iy, jdx = 0, jdy = -2 * iy;
    if (radius == 1) {
        if (tr) set_px(s, x0, y0 - 1, px);
        if (tl) set_px(s, x0 - 1, y0 - 1, px);
        if (bl) set_px(s, x0 - 1, y0, px);
        if (br) set_px(s, x0, y0, px);
        return;
    }
    if (thickness != 0) {
        while (x < y) {
            if (f >= 0) {
                y--;
                ddy += 2;
                f += ddy;
            }
            if (jf >= 0) {
                iy--;
                jdy += 2;
                jf += jdy;
            }
            x++;
            ddx += 2;
            f += ddx + 1;
            jdx += 2;
            jf += jdx + 1;
            if (thickness > 1) thickness = y - iy;
            for (int i = 0; i < thickness; ++i) {
                int y1 = y - i;
                if (tr) {
                    set_px(s, x0 + x - 1, y0 - y1, px);
                    set_px(s, x0 + y1 - 1, y0 - x, px);
                }
                if (tl) {
                    set_px(s, x0 - x, y0 - y1, px);
                    set_px(s, x0 - y1, y0 - x, px);
                }
                if (bl) {
                    set_px(s, x0 - x, y0 + y1 - 1, px);
                    set_px(s, x0 - y1, y0 + x - 1, px);
                }
                if (br) {
                    set_px(s, x0 + x - 1, y0 + y1 - 1, px);
                    set_px(s, x0 + y1 - 1, y0 + x - 1, px);
                }
            }
        }
        return;
    }
    while (x < y) {
        if (f >= 0) {
            y--;
            ddy += 2;
            f += ddy;
        }
        x++;
        ddx += 2;
        f += ddx + 1;
        if (tr) {
            for (int yy = y0 - x; yy <= y0; ++yy) set_px(s, x0 + y - 1, yy, px);
            for (int yy = y0 - y; yy <= y0; ++yy) set_px(s, x0 + x - 1, yy, px);
        }
        if (tl) {
            for (int yy = y0 - x; yy <= y0; ++yy) set_px(s, x0 - y, yy, px);
            for (int yy = y0 - y; yy <= y0; ++yy) set_px(s, x0 - x, yy, px);
        }
        if (bl) {
            for (int yy = y0; yy < y0 + x; ++yy) set_px(s, x0 - y, yy, px);
            for (int yy = y0; yy < y0 + y; ++yy) set_px(s, x0 - x, yy, px);
        }
        if (br) {
            for (int yy = y0; yy < y0 + x; ++yy) set_px(s, x0 + y - 1, yy, px);
            for (int yy = y0; yy < y0 + y; ++yy) set_px(s, x0 + x - 1, yy, px);
        }
    }
}

static void round_rect(Surface& s, uint32_t px, int x1, int y1, int x2, int y2, int radius, int width, int tl, int tr,
                       int bl, int br) {
    if (tl < 0) tl = radius;
    if (tr < 0) tr = radius;
    if (bl < 0) bl = radius;
    if (br < 0) br = radius;
    int W = x2 - x1 + 1, H = y2 - y1 + 1;
    if (tl + tr > W || bl + br > W || tl + bl > H || tr + br > H) {
        float q = std::min(std::min((float)W / (float)(tl + tr), (float)H / (float)(tl + bl)),
                           std::min((float)W / (float)(bl + br), (float)H / (float)(tr + br)));
        tl = (int)((float)tl * q);
        tr = (int)((float)tr * q);
        bl = (int)((float)bl * q);
        br = (int)((float)br * q);
    }
    if (width == 0) {
        int xs[8] = {x1, x1 + tl, x2 - tr, x2, x2, x2 - br, x1 + bl, x1};
        int ys[8] = {y1 + tl, y1, y1, y1 + tr, y2 - br, y2, y2, y2 - bl};
        fillpoly(s, px, xs, ys, 8);
        quadrant(s, px, x2 - tr + 1, y1 + tr, tr, 0, true, false, false, false);
        quadrant(s, px, x1 + tl, y1 + tl, tl, 0, false, true, false, false);
        quadrant(s, px, x1 + bl, y2 - bl + 1, bl, 0, false, false, true, false);
        quadrant(s, px, x2 - br + 1, y2 - br + 1, br, 0, false, false, false, true);
        return;
    }
    int o = width / 2 - 1 + width % 2;
    if (x2 - tr == x1 + tl)
        for (int i = 0; i < width; ++i) set_px(s, x1 + tl, y1 + i, px);
    else
        line(s, px, x1 + tl, y1 + o, x2 - tr, y1 + o, width);
    if (y2 - bl == y1 + tl)
        for (int i = 0; i < width; ++i) set_px(s, x1 + i, y1 + tl, px);
    else
        line(s, px, x1 + o, y1 + tl, x1 + o, y2 - bl, width);
    if (x2 - br == x1 + bl)
        for (int i = 0; i < width; ++i) set_px(s, x1 + bl, y2 - i, px);
    else
        line(s, px, x1 + bl, y2 - width / 2, x2 - br, y2 - width / 2, width);
    if (y2 - br == y1 + tr)
        for (int i = 0; i < width; ++i) set_px(s, x2 - i, y1 + tr, px);
    else
        line(s, px, x2 - width / 2, y1 + tr, x2 - width / 2, y2 - br, width);
    quadrant(s, px, x2 - tr + 1, y1 + tr, tr, width, true, false, false, false);
    quadrant(s, px, x1 + tl, y1 + tl, tl, width, false, true, false, false);
    quadrant(s, px, x1 + bl, y2 - bl + 1, bl, width, false, false, true, false);
    quadrant(s, px, x2 - br + 1, y2 - br + 1, br, width, false, false, false, true);
}

void rect(Surface& s, uint32_t px, int x, int y, int w, int h, int width, int border_radius, int tl, int tr, int bl,
          int br) {
    if (width < 0) return;
    if ((border_radius <= 0 && tl <= 0 && tr <= 0 && bl <= 0 && br <= 0) || abs(w) < 2 || abs(h) < 2) {
        int cx0 = std::max(x, 0), cy0 = std::max(y, 0);
        int cx1 = std::min(x + w, s.w), cy1 = std::min(y + h, s.h);
        int cw = cx1 - cx0, ch = cy1 - cy0;
        if (w <= 0 || h <= 0 || cw <= 0 || ch <= 0) return;
        if (width > 0 && width * 2 < cw && width * 2 < ch) {
            int x2 = x + w - 1, y2 = y + h - 1;
            for (int i = 0; i < width; ++i) {
                hline(s, px, x, y + i, x2);
                hline(s, px, x, y2 - i, x2);
            }
            for (int i = 0; i < (y2 - y) - 2 * width + 1; ++i) {
                hline(s, px, x, y + width + i, x + width - 1);
                hline(s, px, x2 - width + 1, y + width + i, x2);
            }
        } else {
            for (int yy = cy0; yy < cy1; ++yy)
                for (int xx = cx0; xx < cx1; ++xx) s.at(xx, yy) = px;
        }
        return;
    }
    if (w < 0) {
        x += w;
        w = -w;
    }
    if (h < 0) {
        y += h;
        h = -h;
    }
    if (width > w / 2 || width > h / 2) width = std::max(w / 2, h / 2);
    round_rect(s, px, x, y, x + w - 1, y + h - 1, border_radius, width, tl, tr, bl, br);
}

// ---- transform.rotate -------------------------------------------------------------------------
Surface rotate(const Surface& src, double angle) {
    if (fmod(angle, 90.0) == 0.0) {
        int turns = angle >= 0 ? ((int)(angle / 90) % 4) : -((int)(-angle / 90) % 4);
        if (turns < 0) turns += 4;
        Surface out((turns % 2) ? src.h : src.w, (turns % 2) ? src.w : src.h);
        out.has_key = src.has_key;
        out.key = src.key;
        for (int y = 0; y < out.h; ++y)
            for (int x = 0; x < out.w; ++x) {
                uint32_t v;
                // counter-clockwise quarter turns
                if (turns == 0)
                    v = src.at(x, y);
                else if (turns == 1)
                    v = src.at(src.w - 1 - y, x);
                else if (turns == 2)
                    v = src.at(src.w - 1 - x, src.h - 1 - y);
                else
                    v = src.at(y, src.h - 1 - x);
                if (turns == 0 && src.has_key && v == src.key) v = 0;  // blitted onto a zeroed surface
                out.at(x, y) = v;
            }
        return out;
    }
    double rad = angle * .01745329251994329;
    double sa = sin(rad), ca = cos(rad);
    double cx = ca * src.w, cy = ca * src.h, sx = sa * src.w, sy = sa * src.h;
    int nx = (int)std::max(std::max(fabs(cx + sy), fabs(cx - sy)), std::max(fabs(-cx + sy), fabs(-cx - sy)));
    int ny = (int)std::max(std::max(fabs(sx + cy), fabs(sx - cy)), std::max(fabs(-sx + cy), fabs(-sx - cy)));
    Surface dst(nx, ny);
    dst.has_key = src.has_key;
    dst.key = src.key;
    uint32_t bg = src.has_key ? src.key : src.at(0, 0);
    int half = dst.h / 2;
    int xd = (src.w - dst.w) << 15, yd = (src.h - dst.h) << 15;
    int isin = (int)(sa * 65536), icos = (int)(ca * 65536);
    int ax = (dst.w << 15) - (int)(ca * ((dst.w - 1) << 15));
    int ay = (dst.h << 15) - (int)(sa * ((dst.w - 1) << 15));
    int xmax = (src.w << 16) - 1, ymax = (src.h << 16) - 1;
    for (int yy = 0; yy < dst.h; ++yy) {
        int dx = (ax + isin * (half - yy)) + xd;
        int dy = (ay - icos * (half - yy)) + yd;
        for (int xx = 0; xx < dst.w; ++xx) {
            if (dx < 0 || dy < 0 || dx > xmax || dy > ymax)
                dst.at(xx, yy) = bg;
            else
                dst.at(xx, yy) = src.at(dx >> 16, dy >> 16);
            dx += icos;
            dy += isin;
        }
    }
    return dst;
}

// pygame.math.Vector2.rotate (degrees; quarter turns special-cased with epsilon 1e-6)
void vec_rotate(double x, double y, double angle, double* ox, double* oy) {
    const double eps = 1e-6;
    angle = fmod(angle, 360.0);
    if (angle < 0) angle += 360.0;
    if (fmod(angle + eps, 90.0) < 2 * eps) {
        int q = (int)((angle + eps) / 90);
        if (q == 0 || q == 4) {
            *ox = x;
            *oy = y;
        } else if (q == 1) {
            *ox = -y;
            *oy = x;
        } else if (q == 2) {
            *ox = -x;
            *oy = -y;
        } else {
            *ox = y;
            *oy = -x;
        }
        return;
    }
    double rad = angle * M_PI / 180.0;
    double sv = sin(rad), cv = cos(rad);
    *ox = cv * x - sv * y;
    *oy = sv * x + cv * y;
}

}  // namespace mgr
