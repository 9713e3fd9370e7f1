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
    const MgAssetHeader* hdr;
    const uint32_t* entries;       // sprite pixel lists (the kernel re-points this at a shared-memory copy)
    const uint8_t (*palette)[4];   // ditto
    const uint8_t* tmpl[4];
};

MG_HD MgAssetView mg_asset_view(const void* blob) {
    MgAssetView v;
    const uint8_t* b = (const uint8_t*)blob;
    v.hdr = (const MgAssetHeader*)b;
    v.entries = (const uint32_t*)(b + v.hdr->entries_off);
    v.palette = v.hdr->palette;
    for (int i = 0; i < 4; ++i) v.tmpl[i] = v.hdr->tmpl_off[i] ? b + v.hdr->tmpl_off[i] : nullptr;
    return v;
}

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
