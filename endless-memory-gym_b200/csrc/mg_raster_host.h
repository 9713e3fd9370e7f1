// mg_raster_host.h — host-side software raster used ONLY to build the static sprite atlases.
//
// The reference draws its sprites with pygame 2.4.0 (`pygame.draw.circle/line/rect`,
// `pygame.transform.rotate`) at construction time: character_controller.py:29-75,
// pygame_assets.py:254-304 (Command), :177-205 (Exit), :145-152 (Coin), mystery_path.py:175-182.
// pygame is a third-party wheel (setup.py:35, not vendored); the routines below restate the
// published integer algorithms of its draw.c / transform.c for exactly those call sites.
// Pixels are SDL XRGB8888 mapped values (0x00RRGGBB).
#pragma once
#include <stdint.h>
#include <vector>

namespace mgr {

struct Surface {
    int w = 0, h = 0;
    std::vector<uint32_t> px;
    bool has_key = false;
    uint32_t key = 0;
    Surface() {}
    Surface(int w_, int h_) : w(w_), h(h_), px((size_t)w_ * h_, 0u) {}
    uint32_t& at(int x, int y) { return px[(size_t)y * w + x]; }
    uint32_t at(int x, int y) const { return px[(size_t)y * w + x]; }
};

inline uint32_t rgb(int r, int g, int b) { return ((uint32_t)r << 16) | ((uint32_t)g << 8) | (uint32_t)b; }

void fill(Surface& s, uint32_t px);
void circle(Surface& s, uint32_t px, int x0, int y0, int radius, int width, bool legacy_thick);
void line(Surface& s, uint32_t px, int x1, int y1, int x2, int y2, int width);
void rect(Surface& s, uint32_t px, int x, int y, int w, int h, int width, int border_radius, int tl, int tr,
          int bl, int br);
Surface rotate(const Surface& src, double angle);
void vec_rotate(double x, double y, double angle, double* ox, double* oy);

}  // namespace mgr
