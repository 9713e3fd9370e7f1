// mg_frame.h — frame composition: background template + sparse sprite pixel lists into an
// 84x84x3 byte frame in pygame `surfarray.array3d` layout ([x][y][c], column x = 252 contiguous bytes).
//
// Replaces `_draw_surfaces` + `pygame.surfarray.array3d` (memory_gym/mortar_mayhem.py:93-103,367,
// mystery_path.py:271-274): painter's order background -> arena -> agent -> command / cross, with
// SDL colour-key blits clipped to the screen. A warp keeps ONE frame buffer in shared memory that
// always holds the clean background between env-steps: compose = patch a few hundred bytes,
// (bulk-store the 21168 bytes), restore = un-patch.
//
// `Ctx` supplies the execution context: lane id / lane count, a warp sync, and a full-frame copy.
#pragma once
#include "mg_common.h"

struct MgAssetView {
    const uint8_t* base;           // bytes [0, hdr->stage_bytes) of the blob: global memory, or the kernel's
                                   // shared-memory copy (header, sprite lists, dense bitmaps, tables)
    const MgAssetHeader* hdr;
    const uint32_t* entries;
    const uint8_t (*palette)[4];
    const uint8_t* tmpl[4];        // full-frame templates: always in global memory (TMA source)
};

// `data` = where the stageable prefix lives, `blob` = the whole blob in global memory
MG_HD MgAssetView mg_asset_view2(const void* data, const void* blob) {
    MgAssetView v;
    v.base = (const uint8_t*)data;
    v.hdr = (const MgAssetHeader*)v.base;
    v.entries = (const uint32_t*)(v.base + v.hdr->entries_off);
    v.palette = v.hdr->palette;
    for (int i = 0; i < 4; ++i) v.tmpl[i] = v.hdr->tmpl_off[i] ? (const uint8_t*)blob + v.hdr->tmpl_off[i] : nullptr;
    return v;
}
MG_HD MgAssetView mg_asset_view(const void* blob) { return mg_asset_view2(blob, blob); }

// colour-key blit of a sparse sprite with its top-left at (ax, ay), clipped to the screen
template <class Ctx>
MG_HD void mg_sprite_draw(Ctx& c, uint8_t* buf, const MgAssetView& a, int slot, int ax, int ay) {
    const MgSprite sp = a.hdr->sprite[slot];
    const uint32_t* e = a.entries + sp.first;
    for (int i = c.lane; i < (int)sp.count; i += c.nl) {
        uint32_t v = e[i];
        int x = ax + (int)(v & 255u), y = ay + (int)((v >> 8) & 255u);
        if ((unsigned)x < (unsigned)MG_W && (unsigned)y < (unsigned)MG_W) {
            const uint8_t* col = a.palette[(v >> 16) & 255u];
            uint8_t* o = buf + x * MG_COL_BYTES + y * 3;
            o[0] = col[0];
            o[1] = col[1];
            o[2] = col[2];
        }
    }
}

// undo mg_sprite_draw: put the background back. `bg(x, y)` returns the clean frame's 0x00BBGGRR at a pixel
// (closed form per env type: no memory traffic on the un-patch path).
template <class Ctx, class Bg>
MG_HD void mg_sprite_erase(Ctx& c, uint8_t* buf, const MgAssetView& a, int slot, int ax, int ay, const Bg& bg) {
    const MgSprite sp = a.hdr->sprite[slot];
    const uint32_t* e = a.entries + sp.first;
    for (int i = c.lane; i < (int)sp.count; i += c.nl) {
        uint32_t v = e[i];
        int x = ax + (int)(v & 255u), y = ay + (int)((v >> 8) & 255u);
        if ((unsigned)x < (unsigned)MG_W && (unsigned)y < (unsigned)MG_W) {
            uint32_t col = bg(x, y);
            uint8_t* o = buf + x * MG_COL_BYTES + y * 3;
            o[0] = (uint8_t)col;
            o[1] = (uint8_t)(col >> 8);
            o[2] = (uint8_t)(col >> 16);
        }
    }
}

// MortarArena background (pygame_assets.py:316-343,362-368): size x size tiles of 14 px, 1-px light frame
struct MgArenaBg {
    int origin, extent;
    MG_HD uint32_t operator()(int x, int y) const {
        unsigned u = (unsigned)(x - origin), v = (unsigned)(y - origin);
        if (u >= (unsigned)extent || v >= (unsigned)extent) return 0u;
        unsigned p = u % 14u, q = v % 14u;
        bool edge = p == 0u || p == 13u || q == 0u || q == 13u;
        return edge ? MG_RGB_TILE_EDGE : MG_RGB_TILE_FILL;
    }
};
struct MgBlackBg {
    MG_HD uint32_t operator()(int, int) const { return 0u; }
};

// copy the w x h pixel block at (x0, y0) from a full-frame template (clipped)
template <class Ctx>
MG_HD void mg_block_from(Ctx& c, uint8_t* buf, const uint8_t* src, int x0, int y0, int w, int h) {
    int xa = x0 < 0 ? 0 : x0, ya = y0 < 0 ? 0 : y0;
    int xb = x0 + w > MG_W ? MG_W : x0 + w, yb = y0 + h > MG_W ? MG_W : y0 + h;
    if (xb <= xa || yb <= ya) return;
    int rowb = (yb - ya) * 3, total = (xb - xa) * rowb;
    for (int i = c.lane; i < total; i += c.nl) {
        int off = (xa + i / rowb) * MG_COL_BYTES + ya * 3 + i % rowb;
        buf[off] = src[off];
    }
}

// fill the w x h pixel block at (x0, y0) with one colour (clipped)
template <class Ctx>
MG_HD void mg_block_fill(Ctx& c, uint8_t* buf, int x0, int y0, int w, int h, int r, int g, int b) {
    int xa = x0 < 0 ? 0 : x0, ya = y0 < 0 ? 0 : y0;
    int xb = x0 + w > MG_W ? MG_W : x0 + w, yb = y0 + h > MG_W ? MG_W : y0 + h;
    if (xb <= xa || yb <= ya) return;
    int rows = yb - ya, total = (xb - xa) * rows;
    for (int i = c.lane; i < total; i += c.nl) {
        uint8_t* o = buf + (xa + i / rows) * MG_COL_BYTES + (ya + i % rows) * 3;
        o[0] = (uint8_t)r;
        o[1] = (uint8_t)g;
        o[2] = (uint8_t)b;
    }
}

// ---- Mortar Mayhem family -----------------------------------------------------------------------
// buf holds template 0 (blue arena) on entry
template <class Ctx>
MG_HD void mg_mm_compose(Ctx& c, uint8_t* buf, const MgAssetView& a, const MgMMFrame& f, int origin, int agent_half,
                         int icon_x, int icon_y) {
    if (f.red) {
        // MortarArena.toggle_tiles(target): every tile but the target turns red (pygame_assets.py:385-403)
        c.copy_frame(buf, a.tmpl[1]);
        mg_block_from(c, buf, a.tmpl[0], origin + 14 * f.tx, origin + 14 * f.ty, 14, 14);
        c.sync();
    }
    mg_sprite_draw(c, buf, a, MG_SPR_AGENT0 + f.srot, f.sx - agent_half, f.sy - agent_half);
    if (f.icon >= 0) {
        c.sync();  // the command layer is blitted after (on top of) the agent
        mg_sprite_draw(c, buf, a, MG_SPR_ICON0 + f.icon, icon_x, icon_y);
    }
}

template <class Ctx>
MG_HD void mg_mm_restore(Ctx& c, uint8_t* buf, const MgAssetView& a, const MgMMFrame& f, int origin, int size,
                         int agent_half, int icon_x, int icon_y) {
    if (f.red) {
        c.copy_frame(buf, a.tmpl[0]);
        return;
    }
    const MgArenaBg bg{origin, 14 * size};
    mg_sprite_erase(c, buf, a, MG_SPR_AGENT0 + f.srot, f.sx - agent_half, f.sy - agent_half, bg);
    if (f.icon >= 0) mg_sprite_erase(c, buf, a, MG_SPR_ICON0 + f.icon, icon_x, icon_y, bg);
}

// ---- Mystery Path ------------------------------------------------------------------------------
struct MgMPDraw {
    MgMPFrame f;
    int show_origin, show_goal;  // options
    int ox, oy, gx, gy;          // origin / goal tile indices
};

template <class Ctx>
MG_HD void mg_mp_compose(Ctx& c, uint8_t* buf, const MgAssetView& a, const MgMPDraw& d, int agent_half) {
    // MysteryPath.draw_to_surface (pygame_assets.py:749-758): goal tile first in path order, origin last
    if (d.show_goal) mg_block_fill(c, buf, 12 * d.gx, 12 * d.gy, 12, 12, 0, 255, 0);
    if (d.show_origin) {
        if (d.show_goal) c.sync();
        mg_block_fill(c, buf, 12 * d.ox, 12 * d.oy, 12, 12, 0, 0, 255);
    }
    if (d.show_goal || d.show_origin) c.sync();
    mg_sprite_draw(c, buf, a, MG_SPR_AGENT0 + d.f.srot, d.f.sx - agent_half, d.f.sy - agent_half);
    if (d.f.cross) {
        c.sync();
        mg_sprite_draw(c, buf, a, MG_SPR_CROSS, d.f.cross_x, d.f.cross_y);
    }
}

template <class Ctx>
MG_HD void mg_mp_restore(Ctx& c, uint8_t* buf, const MgAssetView& a, const MgMPDraw& d, int agent_half) {
    const MgBlackBg bg;
    mg_sprite_erase(c, buf, a, MG_SPR_AGENT0 + d.f.srot, d.f.sx - agent_half, d.f.sy - agent_half, bg);
    if (d.f.cross) mg_sprite_erase(c, buf, a, MG_SPR_CROSS, d.f.cross_x, d.f.cross_y, bg);
    if (d.show_goal) mg_block_fill(c, buf, 12 * d.gx, 12 * d.gy, 12, 12, 0, 0, 0);
    if (d.show_origin) mg_block_fill(c, buf, 12 * d.ox, 12 * d.oy, 12, 12, 0, 0, 0);
}

// ---- Searing Spotlights ----------------------------------------------------------------------------
// The frame is a per-pixel function of the state: layers in painter's order
// (searing_spotlights.py:526-542), the spotlight layer being a black surface with colour-keyed discs blended
// with per-surface alpha (SDL BlitNtoNSurfaceAlphaKey: d += trunc((0 - d) * A / 255)).
struct MgSSDrawK {
    int n_layers;
    int layers[6];
    int cb_x, cb_w;
    int show_last_action;
    int has_coin_layer;   // coin_surface is None when num_coins == 0
};

MG_HD bool mg_disc_has(const int8_t* discs, int r, int x0, int y0, int x, int y) {
    int dy = y - y0;
    if (r < 1 || dy < -r || dy >= r) return false;
    const int8_t* row = discs + 2 * (r * (r - 1) + dy + r);
    int dx = x - x0;
    return dx >= (int)row[0] && dx <= (int)row[1];
}

// colour of a dense sprite at screen pixel (x, y) when its top-left is (ax, ay); returns false if transparent
MG_HD bool mg_dense_at(const MgAssetView& a, int slot, int ax, int ay, int x, int y, uint32_t& col) {
    const MgSprite sp = a.hdr->sprite[slot];
    unsigned u = (unsigned)(x - ax), v = (unsigned)(y - ay);
    if (u >= (unsigned)sp.w || v >= (unsigned)sp.h) return false;
    uint8_t pi = a.base[a.hdr->dense_off[slot] + v * sp.w + u];
    if (pi == 255) return false;
    const uint8_t* c = a.palette[pi];
    col = (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16);
    return true;
}

MG_HD bool mg_ss_layer_at(const MgAssetView& a, const MgSSFrame& f, int layer, int x, int y, uint32_t& col) {
    if (layer == 5) {  // MG_SS_L_AGENT
        int half = a.hdr->agent_dim >> 1;
        return mg_dense_at(a, MG_SPR_AGENT0 + f.srot, f.sx - half, f.sy - half, x, y, col);
    }
    if (layer == 4) {  // MG_SS_L_EXIT
        int slot = f.exit_open ? MG_SPR_EXIT_OPEN : MG_SPR_EXIT_CLOSED;
        const MgSprite sp = a.hdr->sprite[slot];
        return mg_dense_at(a, slot, f.exit_x - (sp.w >> 1), f.exit_y - (sp.h >> 1), x, y, col);
    }
    if (layer == 3) {  // MG_SS_L_COIN: coins are drawn in list order, so the last one covering the pixel wins
        const MgSprite sp = a.hdr->sprite[MG_SPR_COIN];
        for (int c = (int)f.n_coins - 1; c >= 0; --c)
            if (mg_dense_at(a, MG_SPR_COIN, f.coins[c][0] - (sp.w >> 1), f.coins[c][1] - (sp.h >> 1), x, y, col))
                return true;
        return false;
    }
    return false;
}

MG_HD uint32_t mg_ss_bg(int kind, int x, int y) {
    if (kind == 2) return 0x00FFFFFFu;
    if (kind == 3) return 0u;
    // get_tiled_background_surface (pygame_assets.py:222-239): 12-px tiles, white where (tx + ty) is even
    if ((((unsigned)x / 12u + (unsigned)y / 12u) & 1u) == 0u) return 0x00FFFFFFu;
    return kind == 0 ? 0x00FF0000u /* (0,0,255) */ : 0x000000FFu /* (255,0,0) */;
}

// top bar (searing_spotlights.py:372-389, 465-468, 222-224, 518-523), y < 4
MG_HD uint32_t mg_ss_bar(const MgSSDrawK& k, const MgSSFrame& f, int x) {
    const uint32_t act[3] = {120u | (120u << 8) | (120u << 16), 116u | (1u << 8) | (113u << 16), 255u | (94u << 8) | (14u << 16)};
    uint32_t c = 50u | (50u << 8) | (50u << 16);
    if (x < 42) c = 0x0000FF00u;  // green health
    if (k.show_last_action && x >= 42) c = act[x < 63 ? f.bar_c0 : f.bar_c1];
    // health lost: red from the left; beyond the health half it only survives until the action boxes are repainted
    if ((x < f.red_w_max && (x < 42 || !k.show_last_action)) || x < f.red_now) c = 0x000000FFu;
    if (f.coin_bar && x >= k.cb_x && x < k.cb_x + k.cb_w) c = f.coin_bar == 1 ? (255u | (255u << 8)) : (50u | (50u << 8) | (50u << 16));
    return c;
}

// Per-frame scene under the spotlight layer, resolved once per frame: among coin < exit < agent (their painter's
// order is fixed by the insert logic of searing_spotlights.py:526-542) the ones BELOW the spotlight layer.
struct MgSSUnder {
    int ag_x0, ag_y0, ag_dim;      // agent sprite box (dim 0 = not below the spotlight layer)
    int ex_x0, ex_y0, ex_w, ex_h;  // exit sprite box (w 0 = not below)
    int co_w, co_h, n_coins;       // coin sprite size (n_coins 0 = none below)
    const uint8_t* ag_bmp;
    const uint8_t* ex_bmp;
    const uint8_t* co_bmp;
    int bg_kind;
    int bar_below;                 // never with the reference's orders; kept for generality
};

MG_HD MgSSUnder mg_ss_scene(const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f, int spot_li) {
    MgSSUnder u;
    u.ag_dim = u.ex_w = u.ex_h = u.n_coins = u.bar_below = 0;
    u.ag_x0 = u.ag_y0 = u.ex_x0 = u.ex_y0 = u.co_w = u.co_h = 0;
    u.ag_bmp = u.ex_bmp = u.co_bmp = a.base;
    u.bg_kind = f.bg_kind;
    for (int lj = 1; lj < spot_li; ++lj) {
        int M = k.layers[lj];
        if (M == 5) {
            u.ag_dim = a.hdr->agent_dim;
            u.ag_x0 = f.sx - (u.ag_dim >> 1);
            u.ag_y0 = f.sy - (u.ag_dim >> 1);
            u.ag_bmp = a.base + a.hdr->dense_off[MG_SPR_AGENT0 + f.srot];
        } else if (M == 4) {
            int slot = f.exit_open ? MG_SPR_EXIT_OPEN : MG_SPR_EXIT_CLOSED;
            const MgSprite sp = a.hdr->sprite[slot];
            u.ex_w = sp.w;
            u.ex_h = sp.h;
            u.ex_x0 = f.exit_x - (sp.w >> 1);
            u.ex_y0 = f.exit_y - (sp.h >> 1);
            u.ex_bmp = a.base + a.hdr->dense_off[slot];
        } else if (M == 3) {
            const MgSprite sp = a.hdr->sprite[MG_SPR_COIN];
            u.co_w = sp.w;
            u.co_h = sp.h;
            u.n_coins = f.n_coins;
            u.co_bmp = a.base + a.hdr->dense_off[MG_SPR_COIN];
        } else if (M == 2) {
            u.bar_below = 1;
        }
    }
    return u;
}

MG_HD uint32_t mg_pal(const MgAssetView& a, uint8_t pi) {
    const uint8_t* c = a.palette[pi];
    return (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16);
}

// colour of what lies UNDER the spotlight layer at (x, y): agent over exit over coins over the background
MG_HD uint32_t mg_ss_under(const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f, const MgSSUnder& u, int x,
                           int y) {
    if (u.bar_below && y < 4) return mg_ss_bar(k, f, x);
    unsigned du = (unsigned)(x - u.ag_x0), dv = (unsigned)(y - u.ag_y0);
    if (du < (unsigned)u.ag_dim && dv < (unsigned)u.ag_dim) {
        uint8_t pi = u.ag_bmp[dv * u.ag_dim + du];
        if (pi != 255) return mg_pal(a, pi);
    }
    du = (unsigned)(x - u.ex_x0), dv = (unsigned)(y - u.ex_y0);
    if (du < (unsigned)u.ex_w && dv < (unsigned)u.ex_h) {
        uint8_t pi = u.ex_bmp[dv * u.ex_w + du];
        if (pi != 255) return mg_pal(a, pi);
    }
    for (int c = u.n_coins - 1; c >= 0; --c) {  // coins are drawn in list order: the last one covering the pixel wins
        du = (unsigned)(x - (f.coins[c][0] - (u.co_w >> 1))), dv = (unsigned)(y - (f.coins[c][1] - (u.co_h >> 1)));
        if (du < (unsigned)u.co_w && dv < (unsigned)u.co_h) {
            uint8_t pi = u.co_bmp[dv * u.co_w + du];
            if (pi != 255) return mg_pal(a, pi);
        }
    }
    return mg_ss_bg(u.bg_kind, x, y);
}

MG_HD uint32_t mg_ss_dim(uint32_t c, uint32_t A) {
    uint32_t r = c & 255u, g = (c >> 8) & 255u, b = (c >> 16) & 255u;
    r -= r * A / 255u;  // SDL: d + trunc((0 - d) * A / 255)
    g -= g * A / 255u;
    b -= b * A / 255u;
    return r | (g << 8) | (b << 16);
}

// colour of the layers ABOVE the spotlight layer at (x, y) (visible coins / exit / agent, top bar); false = none
MG_HD bool mg_ss_over(const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f, int spot_li, int x, int y,
                      uint32_t& col) {
    for (int li = k.n_layers - 1; li > spot_li; --li) {
        int L = k.layers[li];
        if (L == 2) {
            if (y < 4) {
                col = mg_ss_bar(k, f, x);
                return true;
            }
        } else if (mg_ss_layer_at(a, f, L, x, y, col)) {
            return true;
        }
    }
    return false;
}

MG_HD void mg_put_px(uint8_t* buf, int x, int y, uint32_t col) {
    uint8_t* o = buf + x * MG_COL_BYTES + y * 3;
    o[0] = (uint8_t)col;
    o[1] = (uint8_t)(col >> 8);
    o[2] = (uint8_t)(col >> 16);
}

// kind 0: under() dimmed by alpha, 1: under() undimmed, 2: over() where present — over a clipped rectangle
template <class Ctx>
MG_HD void mg_ss_rect(Ctx& c, uint8_t* buf, const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f,
                      const MgSSUnder& u, int spot_li, int kind, int x0, int y0, int w, int h) {
    int xa = x0 < 0 ? 0 : x0, ya = y0 < 0 ? 0 : y0;
    int xb = x0 + w > MG_W ? MG_W : x0 + w, yb = y0 + h > MG_W ? MG_W : y0 + h;
    if (xb <= xa || yb <= ya) return;
    const int rows = yb - ya, total = (xb - xa) * rows;
    const uint64_t inv = ((1ull << 24) + (uint64_t)rows - 1) / (uint64_t)rows;  // exact i / rows for i < 2^13
    for (int i = c.lane; i < total; i += c.nl) {
        int q = (int)(((uint64_t)i * inv) >> 24);
        int x = xa + q, y = ya + (i - q * rows);
        uint32_t col;
        if (kind == 2) {
            if (!mg_ss_over(a, k, f, spot_li, x, y, col)) continue;
        } else {
            col = mg_ss_under(a, k, f, u, x, y);
            if (kind == 0) col = mg_ss_dim(col, f.alpha);
        }
        mg_put_px(buf, x, y, col);
    }
}

// the pixels of one spotlight disc (clipped to the screen), column by column — a column of the frame is
// contiguous memory: undimmed under()
template <class Ctx>
MG_HD void mg_ss_disc(Ctx& c, uint8_t* buf, const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f,
                      const MgSSUnder& u, int x0, int y0, int r) {
    if (r < 1) return;
    const int8_t* cols = (const int8_t*)(a.base + a.hdr->disc_off) + 2 * MG_MAX_DISC_R * (MG_MAX_DISC_R + 1) +
                         2 * r * (r - 1);
    int xa = x0 - r < 0 ? 0 : x0 - r, xb = x0 + r > MG_W ? MG_W : x0 + r;  // [xa, xb) x [ya, yb): box on screen
    int ya = y0 - r < 0 ? 0 : y0 - r, yb = y0 + r > MG_W ? MG_W : y0 + r;
    if (xb <= xa || yb <= ya) return;
    const int rows = yb - ya, total = (xb - xa) * rows;
    const uint64_t inv = ((1ull << 24) + (uint64_t)rows - 1) / (uint64_t)rows;
    for (int i = c.lane; i < total; i += c.nl) {
        int q = (int)(((uint64_t)i * inv) >> 24);
        int x = xa + q, y = ya + (i - q * rows);
        int ci = x - x0 + r, dy = y - y0;
        if (dy < (int)cols[2 * ci] || dy > (int)cols[2 * ci + 1]) continue;
        mg_put_px(buf, x, y, mg_ss_under(a, k, f, u, x, y));
    }
}

template <class Ctx>
MG_HD void mg_ss_compose(Ctx& c, uint8_t* buf, const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f) {
    int spot_li = 1;
    for (int li = 0; li < k.n_layers; ++li)
        if (k.layers[li] == 1) spot_li = li;
    const MgSSUnder u = mg_ss_scene(a, k, f, spot_li);
    // 1. everything the spotlight layer dims: black when the lights are fully off, else under() scaled by alpha
    if (f.alpha == 255) {
        MgVec16* w = reinterpret_cast<MgVec16*>(buf);
        const MgVec16 z = {0u, 0u, 0u, 0u};
        for (int i = c.lane; i < MG_FRAME_VEC16; i += c.nl) w[i] = z;
    } else {
        mg_ss_rect(c, buf, a, k, f, u, spot_li, f.alpha == 0 ? 1 : 0, 0, 0, MG_W, MG_W);
    }
    c.sync();
    // 2. inside the colour-keyed discs the layers underneath show through unchanged
    if (f.alpha != 0)
        for (int d = 0; d < (int)f.n_discs; ++d) mg_ss_disc(c, buf, a, k, f, u, f.disc_x[d], f.disc_y[d], f.disc_r[d]);
    c.sync();
    // 3. layers above the spotlight layer, in painter's order (overlaps resolved per pixel by mg_ss_over)
    for (int li = spot_li + 1; li < k.n_layers; ++li) {
        int L = k.layers[li];
        if (L == 2) {
            mg_ss_rect(c, buf, a, k, f, u, spot_li, 2, 0, 0, MG_W, 4);
        } else if (L == 5) {
            int dim = a.hdr->agent_dim;
            mg_ss_rect(c, buf, a, k, f, u, spot_li, 2, f.sx - (dim >> 1), f.sy - (dim >> 1), dim, dim);
        } else if (L == 4) {
            const MgSprite sp = a.hdr->sprite[MG_SPR_EXIT_OPEN];
            mg_ss_rect(c, buf, a, k, f, u, spot_li, 2, f.exit_x - (sp.w >> 1), f.exit_y - (sp.h >> 1), sp.w, sp.h);
        } else if (L == 3) {
            const MgSprite sp = a.hdr->sprite[MG_SPR_COIN];
            for (int q = 0; q < (int)f.n_coins; ++q)
                mg_ss_rect(c, buf, a, k, f, u, spot_li, 2, f.coins[q][0] - (sp.w >> 1), f.coins[q][1] - (sp.h >> 1), sp.w,
                           sp.h);
        }
    }
}
