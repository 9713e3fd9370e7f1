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

// ---- Endless Mystery Path -----------------------------------------------------------------------
// _draw_surfaces (endless_mystery_path.py:137-160): [scrolling icy columns | black] -> past-path tiles -> agent
// -> stamina bar -> fall-off cross.
// Past-path tiles (:131-133): a white 12x12 rect, then a 1-px (210,210,210) frame. pygame draws the frame only
// when the CLIPPED rect is wider than 2 px, else it fills the clipped rect with the frame colour.
template <class Ctx>
MG_HD void mg_ep_tiles(Ctx& c, uint8_t* buf, const MgEPFrame& f, bool erase) {
    const int units = f.n_tiles * 12;  // one (tile, column) = 36 contiguous bytes per lane iteration
    for (int u = c.lane; u < units; u += c.nl) {
        const int t = u / 12, cl = u - 12 * t;
        const int tx = f.tile_x[t], x = tx + cl;
        if ((unsigned)x >= (unsigned)MG_W) continue;
        const int xa = tx < 0 ? 0 : tx, xb = tx + 12 > MG_W ? MG_W : tx + 12;
        const bool all_edge = !(2 < xb - xa) || cl == 0 || cl == 11;
        uint8_t* o = buf + x * MG_COL_BYTES + (int)f.tile_y[t] * 36;
        for (int v = 0; v < 12; ++v) {
            uint8_t g = erase ? 0 : ((all_edge || v == 0 || v == 11) ? 210 : 255);
            o[3 * v] = g, o[3 * v + 1] = g, o[3 * v + 2] = g;
        }
    }
}

// draw_column_tile_surface (pygame_assets.py:782-818) blitted every 12 px from `off` (<= 0): icy (125,177,250) tiles
// with a 1-px (210,210,210) frame
template <class Ctx>
MG_HD void mg_ep_background(Ctx& c, uint8_t* buf, int off) {
    for (int x = c.lane; x < MG_W; x += c.nl) {
        const unsigned u = (unsigned)(x - off) % 12u;
        const bool edge = u == 0u || u == 11u;
        uint8_t* o = buf + x * MG_COL_BYTES;
        for (int y = 0, v = 0; y < MG_W; ++y) {
            const bool e = edge || v == 0 || v == 11;
            o[0] = e ? 210 : 125, o[1] = e ? 210 : 177, o[2] = e ? 210 : 250;
            o += 3;
            if (++v == 12) v = 0;
        }
    }
}

// stamina_surface (:424-428): 4 x 84, green with `bar_h` red rows from the top, blitted at (80, 0)
template <class Ctx>
MG_HD void mg_ep_bar(Ctx& c, uint8_t* buf, int bar_h, bool erase) {
    for (int i = c.lane; i < 4 * MG_W; i += c.nl) {
        const int x = MG_W - 4 + i / MG_W, y = i % MG_W;
        uint8_t* o = buf + x * MG_COL_BYTES + y * 3;
        o[0] = (!erase && y < bar_h) ? 255 : 0;
        o[1] = (!erase && y >= bar_h) ? 255 : 0;
        o[2] = 0;
    }
}

template <class Ctx>
MG_HD void mg_ep_compose(Ctx& c, uint8_t* buf, const MgAssetView& a, const MgEPFrame& f) {
    if (f.bg_off > -1000) {
        mg_ep_background(c, buf, f.bg_off);
        c.sync();
    }
    if (f.n_tiles) {
        mg_ep_tiles(c, buf, f, false);
        c.sync();
    }
    mg_sprite_draw(c, buf, a, MG_SPR_AGENT0 + f.srot, f.sx, f.sy);
    if (f.bar_h >= 0) {
        c.sync();
        mg_ep_bar(c, buf, f.bar_h, false);
    }
    if (f.cross) {
        c.sync();
        mg_sprite_draw(c, buf, a, MG_SPR_CROSS, f.cross_x, f.cross_y);
    }
}

template <class Ctx>
MG_HD void mg_ep_restore(Ctx& c, uint8_t* buf, const MgAssetView& a, const MgEPFrame& f) {
    if (f.bg_off > -1000) return;  // the next frame repaints every pixel
    const MgBlackBg bg;
    if (f.cross) mg_sprite_erase(c, buf, a, MG_SPR_CROSS, f.cross_x, f.cross_y, bg);
    if (f.bar_h >= 0) mg_ep_bar(c, buf, 0, true);
    mg_sprite_erase(c, buf, a, MG_SPR_AGENT0 + f.srot, f.sx, f.sy, bg);
    if (f.n_tiles) mg_ep_tiles(c, buf, f, true);
}

// ---- Searing Spotlights ----------------------------------------------------------------------------
// The frame is a per-pixel function of the state: layers in painter's order
// (searing_spotlights.py:526-542), the spotlight layer being a black surface with colour-keyed discs blended
// with per-surface alpha (SDL BlitNtoNSurfaceAlphaKey: d += trunc((0 - d) * A / 255)).
struct MgSSDrawK {
    int n_layers;
    const int32_t* layers;  // painter's order of this frame (points into the kernel parameters: no local copy)
    int spot_li, bar_li;  // where the spotlight layer / the top bar sit in that order
    int cb_x, cb_w;
    int show_last_action;
    int has_coin_layer;   // coin_surface is None when num_coins == 0
    int border;           // black_background: every spotlight gets a white 1-px outline (Spotlight.has_border)
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

MG_HD uint32_t mg_pal(const MgAssetView& a, uint8_t pi) {
    const uint8_t* c = a.palette[pi];
    return (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16);
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
        if (f.coin_black) {  // Endless with coin_enabled=False: an opaque black surface
            col = 0u;
            return true;
        }
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
    uint32_t c = 50u | (50u << 8) | (50u << 16);
    if (x < 42) c = 0x0000FF00u;  // green health
    if (k.show_last_action && x >= 42) {
        const int which = x < 63 ? f.bar_c0 : f.bar_c1;  // no action grey, negative purple, positive orange
        c = which == 0 ? (120u | (120u << 8) | (120u << 16)) : (which == 1 ? (116u | (1u << 8) | (113u << 16)) : (255u | (94u << 8) | (14u << 16)));
    }
    // health lost: red from the left; beyond the health half it only survives until the action boxes are repainted
    if ((x < f.red_w_max && (x < 42 || !k.show_last_action)) || x < f.red_now) c = 0x000000FFu;
    if (f.coin_bar && x >= k.cb_x && x < k.cb_x + k.cb_w) c = f.coin_bar == 1 ? (255u | (255u << 8)) : (50u | (50u << 8) | (50u << 16));
    return c;
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

// background colours of this frame (white tile, coloured tile), optionally dimmed by the spotlight layer's alpha
MG_HD void mg_ss_bg_colours(int kind, int alpha, uint32_t& c_even, uint32_t& c_odd) {
    c_even = kind == 3 ? 0u : 0x00FFFFFFu;
    c_odd = kind == 0 ? 0x00FF0000u : (kind == 1 ? 0x000000FFu : c_even);
    if (alpha > 0) {
        c_even = mg_ss_dim(c_even, (uint32_t)alpha);
        c_odd = mg_ss_dim(c_odd, (uint32_t)alpha);
    }
}

// one column run [ylo, yhi] of the tiled background (get_tiled_background_surface, pygame_assets.py:222-239:
// 12-px tiles, the even ones white): contiguous bytes of the [x][y][c] frame
MG_HD void mg_ss_bg_run(uint8_t* buf, int x, int ylo, int yhi, uint32_t c_even, uint32_t c_odd) {
    unsigned ty = (unsigned)ylo / 12u, rem = (unsigned)ylo - ty * 12u;
    unsigned par = ((unsigned)x / 12u + ty) & 1u;
    uint8_t* o = buf + x * MG_COL_BYTES + ylo * 3;
    for (int y = ylo; y <= yhi; ++y) {
        uint32_t c = par ? c_odd : c_even;
        o[0] = (uint8_t)c;
        o[1] = (uint8_t)(c >> 8);
        o[2] = (uint8_t)(c >> 16);
        o += 3;
        if (++rem == 12u) {
            rem = 0;
            par ^= 1u;
        }
    }
}

// ---- base buffer + lit mask: the per-frame cost of Searing Spotlights is a masked 16-byte copy -----------------
// Per warp: `base` = the scene UNDER the spotlight layer, undimmed (tiled background + coins/exit/agent that sit
// below the layer), kept across frames and patched only where the scene changed; `lit` = 84 columns x 96 bits,
// bit y of column x set iff a spotlight disc drawn this step covers (x, y). The frame is then
// out[chunk] = lit ? base : dim(base), 16 bytes at a time, through two small per-CTA tables.
#define MG_SS_CHUNKS MG_FRAME_VEC16  // 1323 16-byte chunks per frame

struct MgSSBaseDesc {  // what `base` currently shows
    int32_t valid;
    int32_t lut_alpha;  // alpha the dim LUT was built for (-1: none)
    uint32_t scene_ver;  // MgSSFrame::scene_ver of the frame `base` was last brought up to date with
    int16_t sx, sy;      // ... and the agent pose it shows
    int32_t srot;
    int32_t env;         // whose scene it is (versions of different envs are unrelated)
};

#define MG_SS_LIT_WORDS 224  // 84 * 84 = 7056 pixel bits (x-major, like the frame bytes) = 220.5 words, padded
struct MgSSWarpMem {
    uint8_t base[21248];
    uint32_t lit[MG_SS_LIT_WORDS];
    uint8_t dimlut[256];
    MgSSBaseDesc desc;
    uint8_t step_tmp[96];  // MgSSStepTmp of the logic (mg_ss_logic.h)
};

struct MgSSCtaTables {
    MgVec16 mask[3][64];                 // byte mask of a 16-byte chunk for (phase = chunk % 3, 6 lit bits)
    uint32_t bar_lit[MG_SS_LIT_WORDS];   // lit bits of rows 0..3 of every column (the top bar is never dimmed)
};

MG_HD void mg_ss_tables_init(MgSSCtaTables* t, int tid, int nthreads) {
    for (int w = tid; w < MG_SS_LIT_WORDS; w += nthreads) {
        uint32_t v = 0;
        for (int b = 0; b < 32; ++b) {
            int p = 32 * w + b;
            if (p < MG_W * MG_W && p % MG_W < 4) v |= 1u << b;
        }
        t->bar_lit[w] = v;
    }
    for (int i = tid; i < 3 * 64; i += nthreads) {
        int phase = i / 64, bits = i % 64;
        uint8_t m[16];
        for (int j = 0; j < 16; ++j) m[j] = ((bits >> ((j + phase) / 3)) & 1) ? 0xFF : 0x00;
        uint32_t w[4];
        for (int q = 0; q < 4; ++q)
            w[q] = (uint32_t)m[4 * q] | ((uint32_t)m[4 * q + 1] << 8) | ((uint32_t)m[4 * q + 2] << 16) |
                   ((uint32_t)m[4 * q + 3] << 24);
        MgVec16 v = {w[0], w[1], w[2], w[3]};
        t->mask[phase][bits] = v;
    }
}

// colour under the spotlight layer at (x, y) IGNORING the agent: exit over coins over the background
MG_HD uint32_t mg_ss_under_no_agent(const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f, int spot_li, int x, int y) {
    uint32_t col;
    for (int lj = spot_li - 1; lj > 0; --lj) {
        int M = k.layers[lj];
        if (M == 5 || M == 2) continue;
        if (mg_ss_layer_at(a, f, M, x, y, col)) return col;
    }
    return mg_ss_bg(f.bg_kind, x, y);
}

MG_HD bool mg_ss_layer_below(const MgSSDrawK& k, int spot_li, int layer) {
    for (int lj = 1; lj < spot_li; ++lj)
        if (k.layers[lj] == layer) return true;
    return false;
}

// The whole tiled background into a frame: a 12-px tile run of a column is 36 bytes = 9 aligned words with a
// period-3 pattern (a column starts 4-byte aligned: 252 = 4 * 63). One (column, tile) unit per lane iteration.
template <class Ctx>
MG_HD void mg_ss_fill_bg(Ctx& c, uint8_t* base, uint32_t c_even, uint32_t c_odd) {
    for (int unit = c.lane; unit < MG_W * 7; unit += c.nl) {
        const int x = unit / 7, t = unit - 7 * x;
        const uint32_t col = ((((unsigned)x / 12u) + (unsigned)t) & 1u) ? c_odd : c_even;
        const uint32_t r = col & 255u, g = (col >> 8) & 255u, b = (col >> 16) & 255u;
        const uint32_t w0 = r | (g << 8) | (b << 16) | (r << 24), w1 = g | (b << 8) | (r << 16) | (g << 24),
                       w2 = b | (r << 8) | (g << 16) | (b << 24);
        uint32_t* w = reinterpret_cast<uint32_t*>(base + x * MG_COL_BYTES + t * 36);
        w[0] = w0, w[1] = w1, w[2] = w2, w[3] = w0, w[4] = w1, w[5] = w2, w[6] = w0, w[7] = w1, w[8] = w2;
    }
}

// the top bar into rows 0..3 of `base` (a column start is 4-byte aligned: 12 bytes = three words per column)
template <class Ctx>
MG_HD void mg_ss_draw_bar(Ctx& c, uint8_t* base, const MgSSDrawK& k, const MgSSFrame& f) {
    for (int x = c.lane; x < MG_W; x += c.nl) {
        uint32_t col = mg_ss_bar(k, f, x), r = col & 255u, g = (col >> 8) & 255u, b = (col >> 16) & 255u;
        uint32_t* w = reinterpret_cast<uint32_t*>(base + x * MG_COL_BYTES);
        w[0] = r | (g << 8) | (b << 16) | (r << 24);
        w[1] = g | (b << 8) | (r << 16) | (g << 24);
        w[2] = b | (r << 8) | (g << 16) | (b << 24);
    }
}

template <class Ctx>
MG_HD void mg_ss_draw_layer(Ctx& c, uint8_t* base, const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f, int L) {
    if (L == 3 && f.coin_black) {
        mg_ss_fill_bg(c, base, 0u, 0u);
        c.sync();
    } else if (L == 3) {
        const MgSprite sp = a.hdr->sprite[MG_SPR_COIN];
        for (int q = 0; q < (int)f.n_coins; ++q) {
            mg_sprite_draw(c, base, a, MG_SPR_COIN, f.coins[q][0] - (sp.w >> 1), f.coins[q][1] - (sp.h >> 1));
            c.sync();
        }
    } else if (L == 4) {
        int slot = f.exit_open ? MG_SPR_EXIT_OPEN : MG_SPR_EXIT_CLOSED;
        const MgSprite sp = a.hdr->sprite[slot];
        mg_sprite_draw(c, base, a, slot, f.exit_x - (sp.w >> 1), f.exit_y - (sp.h >> 1));
        c.sync();
    } else if (L == 5) {
        int half = a.hdr->agent_dim >> 1;
        mg_sprite_draw(c, base, a, MG_SPR_AGENT0 + f.srot, f.sx - half, f.sy - half);
        c.sync();
    } else if (L == 2) {
        mg_ss_draw_bar(c, base, k, f);
        c.sync();
    }
}

// the whole base buffer from scratch (after a reset, a background change, in a non-default layer order — and for
// every frame of a single-step launch): inlined at its one call site, so assets and options stay in registers
template <class Ctx>
MG_HD void mg_ss_rebuild_base(Ctx& c, uint8_t* base, const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f) {
    uint32_t c_even, c_odd;
    mg_ss_bg_colours(f.bg_kind, 0, c_even, c_odd);
    mg_ss_fill_bg(c, base, c_even, c_odd);
    c.sync();
    for (int li = 1; li < k.n_layers; ++li)
        if (k.layers[li] != 1) mg_ss_draw_layer(c, base, a, k, f, k.layers[li]);
}

// out-of-line copy for kernels that are short of registers (the endless variant)
template <class Ctx>
MG_HD_NOINLINE void mg_ss_rebuild_base_ool(Ctx c, uint8_t* base, const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f) {
    mg_ss_rebuild_base(c, base, a, k, f);
}

// Bring `base` up to date with frame f. `base` = every layer EXCEPT the spotlight layer, in painter's order
// (layers above the spotlight layer are never dimmed: their pixels are forced lit in the mask instead).
// `simple` = the reference's default order (coins, exit, agent below the spotlight layer, bar on top): then a
// frame that only moved the agent is a small patch; any other order rebuilds.
template <bool kRebuildOol, class Ctx>
MG_HD void mg_ss_update_base(Ctx& c, MgSSWarpMem* wm, const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f,
                             int spot_li, bool simple) {
    uint8_t* base = wm->base;
    const int half = a.hdr->agent_dim >> 1;
    bool same_rest, same_agent;
    int osx, osy, osrot;  // pose of the agent `base` shows
    {
        const MgSSBaseDesc& d = wm->desc;
        same_rest = simple && d.valid && d.scene_ver == f.scene_ver;
        osx = d.sx, osy = d.sy, osrot = d.srot;
        same_agent = osx == f.sx && osy == f.sy && osrot == f.srot;
    }
    c.sync();  // every thread has read the description: lane 0 may rewrite it (nobody reads it again this frame)
    if (c.lane == 0) {
        MgSSBaseDesc& w = wm->desc;
        w.valid = 1;
        w.scene_ver = f.scene_ver;
        w.sx = f.sx, w.sy = f.sy, w.srot = f.srot;
    }
    if (MG_UNLIKELY(!same_rest)) {
        if constexpr (kRebuildOol)
            mg_ss_rebuild_base_ool(c, base, a, k, f);
        else
            mg_ss_rebuild_base(c, base, a, k, f);
        return;
    }
    // Only the agent moved and the bar changed. In the `simple` order the layers under the spotlights are coins <
    // exit < agent and the bar lies on top of everything, so three sets of pixels are written, pairwise disjoint —
    // no ordering between them, the caller's sync before the copy is the only one needed:
    //   (1) pixels of the old agent that the new agent does not cover: what lies underneath (exit, coins, background)
    //   (2) the new agent        — both without rows y < 4: those belong to
    //   (3) the top bar
    if (!same_agent) {
        if (MG_UNLIKELY(f.coin_black)) {  // an opaque coin layer under the agent (Endless, coin disabled): the generic way
            const MgSprite sp = a.hdr->sprite[MG_SPR_AGENT0 + osrot];
            const uint32_t* e = a.entries + sp.first;
            for (int i = c.lane; i < (int)sp.count; i += c.nl) {
                uint32_t v = e[i];
                int x = osx - half + (int)(v & 255u), y = osy - half + (int)((v >> 8) & 255u);
                if ((unsigned)x < (unsigned)MG_W && (unsigned)y < (unsigned)MG_W)
                    mg_put_px(base, x, y, mg_ss_under_no_agent(a, k, f, spot_li, x, y));
            }
            c.sync();
            mg_sprite_draw(c, base, a, MG_SPR_AGENT0 + f.srot, f.sx - half, f.sy - half);
            c.sync();
        } else {
            const uint32_t* pal = reinterpret_cast<const uint32_t*>(a.palette);
            // sprite boxes and dense bitmaps, hoisted out of the pixel loop (byte stores alias everything)
            const int eslot = f.exit_open ? MG_SPR_EXIT_OPEN : MG_SPR_EXIT_CLOSED;
            const MgSprite esp = a.hdr->sprite[eslot], csp = a.hdr->sprite[MG_SPR_COIN];
            const MgSprite osp = a.hdr->sprite[MG_SPR_AGENT0 + osrot], nsp = a.hdr->sprite[MG_SPR_AGENT0 + f.srot];
            const uint8_t* edense = a.base + a.hdr->dense_off[eslot];
            const uint8_t* cdense = a.base + a.hdr->dense_off[MG_SPR_COIN];
            const uint8_t* ndense = a.base + a.hdr->dense_off[MG_SPR_AGENT0 + f.srot];
            const int ex0 = f.exit_x - (esp.w >> 1), ey0 = f.exit_y - (esp.h >> 1);
            const int nx0 = f.sx - half, ny0 = f.sy - half, ox0 = osx - half, oy0 = osy - half;
            const int ncoins = f.n_coins, bg_kind = f.bg_kind;
            const uint32_t* eo = a.entries + osp.first;
            for (int i = c.lane; i < (int)osp.count; i += c.nl) {
                const uint32_t v = eo[i];
                const int x = ox0 + (int)(v & 255u), y = oy0 + (int)((v >> 8) & 255u);
                if ((unsigned)x >= (unsigned)MG_W || (unsigned)(y - 4) >= (unsigned)(MG_W - 4)) continue;
                unsigned u = (unsigned)(x - nx0), w = (unsigned)(y - ny0);
                if (u < (unsigned)nsp.w && w < (unsigned)nsp.h && ndense[w * nsp.w + u] != 255) continue;  // (2) covers it
                uint32_t col = mg_ss_bg(bg_kind, x, y);
                bool found = false;
                u = (unsigned)(x - ex0), w = (unsigned)(y - ey0);
                if (u < (unsigned)esp.w && w < (unsigned)esp.h) {
                    const uint8_t pi = edense[w * esp.w + u];
                    if (pi != 255) col = pal[pi] & 0x00FFFFFFu, found = true;
                }
                for (int q = ncoins - 1; q >= 0 && !found; --q) {  // the coin drawn last wins
                    u = (unsigned)(x - (f.coins[q][0] - (csp.w >> 1))), w = (unsigned)(y - (f.coins[q][1] - (csp.h >> 1)));
                    if (u < (unsigned)csp.w && w < (unsigned)csp.h) {
                        const uint8_t pi = cdense[w * csp.w + u];
                        if (pi != 255) col = pal[pi] & 0x00FFFFFFu, found = true;
                    }
                }
                mg_put_px(base, x, y, col);
            }
            const uint32_t* en = a.entries + nsp.first;
            for (int i = c.lane; i < (int)nsp.count; i += c.nl) {
                const uint32_t v = en[i];
                const int x = nx0 + (int)(v & 255u), y = ny0 + (int)((v >> 8) & 255u);
                if ((unsigned)x >= (unsigned)MG_W || (unsigned)(y - 4) >= (unsigned)(MG_W - 4)) continue;
                mg_put_px(base, x, y, pal[(v >> 16) & (MG_MAX_PALETTE - 1)]);
            }
        }
    }
    mg_ss_draw_bar(c, base, k, f);  // the bar changes every step and sits on top
}

// lit = pixel bits (index 84 * x + y, the frame's own order) covered by a spotlight disc drawn this step, on top
// of `init` (the rows of layers above the spotlight layer). One (disc, column) run per lane iteration, OR-ed in
// with shared-memory atomics (a run of <= 80 bits touches at most 4 words).
template <class Ctx>
MG_HD void mg_ss_build_lit(Ctx& c, MgSSWarpMem* wm, const MgAssetView& a, const MgSSFrame& f, const uint32_t* init) {
    {
        MgVec16* l16 = reinterpret_cast<MgVec16*>(wm->lit);
        const MgVec16* i16 = reinterpret_cast<const MgVec16*>(init);
        const MgVec16 zero = {0u, 0u, 0u, 0u};
        for (int w = c.lane; w < MG_SS_LIT_WORDS / 4; w += c.nl) l16[w] = init ? i16[w] : zero;
    }
    c.sync();
    const int8_t* cols_all = (const int8_t*)(a.base + a.hdr->disc_off) + 2 * MG_MAX_DISC_R * (MG_MAX_DISC_R + 1);
    // the team's warps take alternate discs; within a warp one lane per disc column
    const int width = c.nl < 32 ? c.nl : 32, subs = c.nl / width, sub = c.lane / width, col0 = c.lane - sub * width;
    for (int d = sub; d < (int)f.n_discs; d += subs) {
        const int r = f.disc_r[d], x0 = f.disc_x[d], y0 = f.disc_y[d];
        if (r < 1) continue;
        const int8_t* cols = cols_all + 2 * r * (r - 1);
        int xa = x0 - r < 0 ? 0 : x0 - r, xb = x0 + r > MG_W ? MG_W : x0 + r;
        for (int x = xa + col0; x < xb; x += width) {
            int ci = x - x0 + r;
            int ylo = y0 + (int)cols[2 * ci], yhi = y0 + (int)cols[2 * ci + 1];
            if (ylo < 0) ylo = 0;
            if (yhi > MG_W - 1) yhi = MG_W - 1;
            if (ylo > yhi) continue;
            int lo = MG_W * x + ylo, hi = MG_W * x + yhi;
            for (int w = lo >> 5; w <= (hi >> 5); ++w) {
                int bl = lo - 32 * w, bh = hi - 32 * w;
                if (bl < 0) bl = 0;
                if (bh > 31) bh = 31;
                c.atomic_or(&wm->lit[w], (0xFFFFFFFFu >> (31 - (bh - bl))) << bl);
            }
        }
    }
}

// The steady state of an episode (alpha == 255: dimmed = black): frame = base & mask(lit). One lane takes a GROUP of
// three chunks = 48 bytes = exactly 16 pixels, so its lit bits are one aligned half-word of the flat bitmap and the
// three chunk phases (0, 1, 2 bytes into a pixel) are compile-time: no index arithmetic, no funnel shift.
template <class Ctx>
MG_HD void mg_ss_masked_copy_black(Ctx& c, uint8_t* out, const MgSSWarpMem* wm, const MgSSCtaTables* t, bool stream) {
    const MgVec16* src = reinterpret_cast<const MgVec16*>(wm->base);
    MgVec16* dst = reinterpret_cast<MgVec16*>(out);
    const uint16_t* lit16 = reinterpret_cast<const uint16_t*>(wm->lit);
    constexpr int kGroups = MG_SS_CHUNKS / 3;  // 441
    static_assert(MG_SS_CHUNKS % 3 == 0, "a frame is a whole number of 16-pixel groups");
    for (int g0 = c.lane; g0 < kGroups; g0 += 2 * c.nl) {  // two groups per lane in flight: loads first, then stores
        MgVec16 o[2][3];
        MG_UNROLL
        for (int u = 0; u < 2; ++u) {
            const int g = g0 + u * c.nl;
            if (g >= kGroups) break;
            const uint32_t b = lit16[g];
            MG_UNROLL
            for (int ph = 0; ph < 3; ++ph) {
                const MgVec16 m = t->mask[ph][(b >> (5 * ph)) & 63u], v = src[3 * g + ph];
                o[u][ph].x = v.x & m.x, o[u][ph].y = v.y & m.y, o[u][ph].z = v.z & m.z, o[u][ph].w = v.w & m.w;
            }
        }
        MG_UNROLL
        for (int u = 0; u < 2; ++u) {
            const int g = g0 + u * c.nl;
            if (g >= kGroups) break;
            MG_UNROLL
            for (int ph = 0; ph < 3; ++ph) c.store16(dst + 3 * g + ph, o[u][ph], stream);
        }
    }
}

// frame = lit ? base : dim(base), 16 bytes at a time, straight to the observation buffer (coalesced 128-bit
// stores: consecutive lanes write consecutive chunks). kLut = false: alpha == 255 (dimmed = black), the steady
// state of an episode; kLut = true: 0 < alpha < 255 through the per-alpha table (the light-dimming ramp).
template <bool kLut, class Ctx>
MG_HD void mg_ss_masked_copy_t(Ctx& c, uint8_t* out, const MgSSWarpMem* wm, const MgSSCtaTables* t, bool stream) {
    const MgVec16* src = reinterpret_cast<const MgVec16*>(wm->base);
    MgVec16* dst = reinterpret_cast<MgVec16*>(out);
    // four chunks per lane in flight: all loads first, then the stores (the loop is latency-, not issue-bound)
    for (int i0 = c.lane; i0 < MG_SS_CHUNKS; i0 += 4 * c.nl) {
        MgVec16 o[4];
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * c.nl;
            if (i >= MG_SS_CHUNKS) break;
            // chunk i = bytes 16 i .. 16 i + 15 = pixels p0 .. p0 + 5, p0 = 16 i / 3 = 5 i + i / 3, starting `phase`
            // = 16 i % 3 = i % 3 bytes into pixel p0
            const unsigned q3 = ((unsigned)i * 43691u) >> 17, phase = (unsigned)i - 3u * q3, p0 = 5u * (unsigned)i + q3;
            const unsigned w = p0 >> 5, sh = p0 & 31u;
            const uint32_t v = c.funnel_r(wm->lit[w], wm->lit[w + 1], sh);  // 6 lit bits from the flat pixel bitmap
            const MgVec16 m = t->mask[phase][(int)(v & 63u)], b = src[i];
            if (!kLut) {
                o[u].x = b.x & m.x, o[u].y = b.y & m.y, o[u].z = b.z & m.z, o[u].w = b.w & m.w;
            } else {
                const uint8_t* lut = wm->dimlut;
                uint32_t bw[4] = {b.x, b.y, b.z, b.w}, mw[4] = {m.x, m.y, m.z, m.w}, ow[4];
                for (int q = 0; q < 4; ++q) {
                    uint32_t dm = (uint32_t)lut[bw[q] & 255u] | ((uint32_t)lut[(bw[q] >> 8) & 255u] << 8) |
                                  ((uint32_t)lut[(bw[q] >> 16) & 255u] << 16) | ((uint32_t)lut[bw[q] >> 24] << 24);
                    ow[q] = (bw[q] & mw[q]) | (dm & ~mw[q]);
                }
                o[u].x = ow[0], o[u].y = ow[1], o[u].z = ow[2], o[u].w = ow[3];
            }
        }
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * c.nl;
            if (i < MG_SS_CHUNKS) c.store16(dst + i, o[u], stream);
        }
    }
}

// the cold variants: alpha == 0 (plain copy, the reset frame) and the dimming ramp
template <class Ctx>
MG_HD_NOINLINE void mg_ss_masked_copy_cold(Ctx c, uint8_t* out, MgSSWarpMem* wm, const MgSSCtaTables* t, int alpha,
                                           bool stream) {
    if (alpha == 0) {
        const MgVec16* src = reinterpret_cast<const MgVec16*>(wm->base);
        MgVec16* dst = reinterpret_cast<MgVec16*>(out);
        for (int i = c.lane; i < MG_SS_CHUNKS; i += c.nl) c.store16(dst + i, src[i], stream);
        return;
    }
    if (wm->desc.lut_alpha != alpha) {
        c.sync();
        for (int v = c.lane; v < 256; v += c.nl) wm->dimlut[v] = (uint8_t)((uint32_t)v - (uint32_t)v * (uint32_t)alpha / 255u);
        if (c.lane == 0) wm->desc.lut_alpha = alpha;
        c.sync();
    }
    mg_ss_masked_copy_t<true>(c, out, wm, t, stream);
}

// non-default layer orders: pixels of the sprites painted ABOVE the spotlight layer are never dimmed (cold)
template <class Ctx>
MG_HD_NOINLINE void mg_ss_lit_over(Ctx c, MgSSWarpMem* wm, const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f,
                                   int spot_li) {
    for (int li = spot_li + 1; li < k.n_layers; ++li) {
        int L = k.layers[li];
        int slot = L == 5 ? MG_SPR_AGENT0 + f.srot : (L == 4 ? (f.exit_open ? MG_SPR_EXIT_OPEN : MG_SPR_EXIT_CLOSED) : MG_SPR_COIN);
        if (L != 3 && L != 4 && L != 5) continue;
        if (L == 3 && f.coin_black) {  // an opaque layer above the spotlights: nothing is dimmed
            for (int w = c.lane; w < MG_SS_LIT_WORDS; w += c.nl) wm->lit[w] = 0xFFFFFFFFu;
            continue;
        }
        const MgSprite sp = a.hdr->sprite[slot];
        const uint32_t* e = a.entries + sp.first;
        int reps = L == 3 ? (int)f.n_coins : 1;
        for (int q = 0; q < reps; ++q) {
            int ax = L == 5 ? f.sx - (a.hdr->agent_dim >> 1) : (L == 4 ? f.exit_x - (sp.w >> 1) : f.coins[q][0] - (sp.w >> 1));
            int ay = L == 5 ? f.sy - (a.hdr->agent_dim >> 1) : (L == 4 ? f.exit_y - (sp.h >> 1) : f.coins[q][1] - (sp.h >> 1));
            for (int i = c.lane; i < (int)sp.count; i += c.nl) {
                int x = ax + (int)(e[i] & 255u), y = ay + (int)((e[i] >> 8) & 255u);
                if ((unsigned)x < (unsigned)MG_W && (unsigned)y < (unsigned)MG_W)
                    c.atomic_or(&wm->lit[(MG_W * x + y) >> 5], 1u << ((MG_W * x + y) & 31));
            }
        }
    }
}

// black_background: every spotlight has a white 1-px outline (pygame_assets.py:112-113) (cold)
template <class Ctx>
MG_HD_NOINLINE void mg_ss_outlines(Ctx c, uint8_t* buf, const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f,
                                   int spot_li) {
    // 3b. black_background: white 1-px outlines on the spotlight surface (pygame_assets.py:112-113). An outline
    //     pixel survives unless a LATER spotlight's disc is drawn over it; it is blended like any non-key pixel:
    //     d + trunc((255 - d) * A / 255) over what lies underneath.
    {
        c.sync();  // the outline bytes land after the vector stores of the other lanes
        const uint16_t* start = (const uint16_t*)(a.base + a.hdr->ring_off);
        const int8_t* ring = (const int8_t*)(start + MG_MAX_DISC_R + 2);
        const int8_t* rows = (const int8_t*)(a.base + a.hdr->disc_off);
        for (int d = 0; d < (int)f.n_discs; ++d) {
            int r = f.disc_r[d], n = (int)start[r + 1] - (int)start[r];
            const int8_t* px = ring + 2 * (int)start[r];
            for (int i = c.lane; i < n; i += c.nl) {
                int x = f.disc_x[d] + (int)px[2 * i], y = f.disc_y[d] + (int)px[2 * i + 1];
                if ((unsigned)x >= (unsigned)MG_W || (unsigned)y >= (unsigned)MG_W) continue;
                bool covered = false;
                for (int e = d + 1; e < (int)f.n_discs && !covered; ++e)
                    covered = mg_disc_has(rows, f.disc_r[e], f.disc_x[e], f.disc_y[e], x, y);
                if (covered) continue;
                uint32_t under = mg_ss_bg(f.bg_kind, x, y), col;
                if (mg_ss_over(a, k, f, spot_li, x, y, col)) continue;  // painted over by a layer above the spotlights
                for (int lj = spot_li - 1; lj > 0; --lj)
                    if (mg_ss_layer_at(a, f, k.layers[lj], x, y, col)) {
                        under = col;
                        break;
                    }
                uint32_t A = f.alpha, cr = under & 255u, cg = (under >> 8) & 255u, cb = (under >> 16) & 255u;
                cr += (255u - cr) * A / 255u;
                cg += (255u - cg) * A / 255u;
                cb += (255u - cb) * A / 255u;
                mg_put_px(buf, x, y, cr | (cg << 8) | (cb << 16));
            }
        }
        c.sync();
    }
}

// One Searing-Spotlights frame -> `out` (the env's 21168 observation bytes in global memory), in two parts so that a
// pipelined rollout can run the first one (the base-buffer patch) on one warp while the other warp already computes
// the next step: mg_ss_emit_base, [barrier], mg_ss_emit_rest.
MG_HD bool mg_ss_simple_order(const MgSSDrawK& k) { return k.spot_li == k.n_layers - 2 && k.bar_li == k.n_layers - 1; }

template <bool kRebuildOol, class Ctx>
MG_HD void mg_ss_emit_base(Ctx& c, const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f, MgSSWarpMem* wm) {
    // 1. every layer but the spotlight layer (kept across frames, patched where it changed)
    c.stamp(8);
    mg_ss_update_base<kRebuildOol>(c, wm, a, k, f, k.spot_li, mg_ss_simple_order(k));
    c.stamp(9);
}

// 2. which pixels stay undimmed: the colour-keyed discs, and whatever is painted above the spotlight layer
template <class Ctx>
MG_HD void mg_ss_emit_lit(Ctx& c, const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f, MgSSWarpMem* wm,
                          const MgSSCtaTables* tables) {
    if (MG_LIKELY(f.alpha != 0)) {
        mg_ss_build_lit(c, wm, a, f, k.bar_li > k.spot_li ? tables->bar_lit : nullptr);
        if (MG_UNLIKELY(!mg_ss_simple_order(k))) {
            c.sync();
            mg_ss_lit_over(c, wm, a, k, f, k.spot_li);
        }
    }
}

// 3. the copy (after a barrier: base buffer and lit bitmap are complete)
template <class Ctx>
MG_HD void mg_ss_emit_copy(Ctx& c, uint8_t* out, const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f,
                           MgSSWarpMem* wm, const MgSSCtaTables* tables, bool stream) {
    // frame = lit ? base : base dimmed by alpha (SDL: d + trunc((0 - d) * A / 255))
    c.stamp(10);
    if (MG_LIKELY(f.alpha == 255))
        mg_ss_masked_copy_black(c, out, wm, tables, stream);
    else
        mg_ss_masked_copy_cold(c, out, wm, tables, f.alpha, stream);
    c.stamp(11);
    // 3b. black_background: white 1-px outlines on the spotlight surface
    if (MG_UNLIKELY(k.border && f.alpha != 0 && a.hdr->ring_off)) mg_ss_outlines(c, out, a, k, f, k.spot_li);
}

template <class Ctx>
MG_HD void mg_ss_emit_rest(Ctx& c, uint8_t* out, const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f,
                           MgSSWarpMem* wm, const MgSSCtaTables* tables, bool stream) {
    mg_ss_emit_lit(c, a, k, f, wm, tables);
    c.sync();
    mg_ss_emit_copy(c, out, a, k, f, wm, tables, stream);
}

template <bool kRebuildOol, class Ctx>
MG_HD void mg_ss_emit(Ctx& c, uint8_t* out, const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f,
                      MgSSWarpMem* wm, const MgSSCtaTables* tables, bool stream) {
    mg_ss_emit_base<kRebuildOol>(c, a, k, f, wm);
    mg_ss_emit_rest(c, out, a, k, f, wm, tables, stream);  // (its first barrier orders the patch before the copy)
}
