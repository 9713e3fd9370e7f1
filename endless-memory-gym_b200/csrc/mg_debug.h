// mg_debug.h — the ground-truth render (`render_mode="debug_rgb_array"`): `_build_debug_surface` + `transform.scale`
// of the reference (memory_gym/mortar_mayhem.py:105-136, 397-399; mortar_mayhem_grid.py:104-135;
// endless_mortar_mayhem.py:114-145; mystery_path.py:103-117; mystery_path_grid.py:102-116;
// searing_spotlights.py:157-185; endless_searing_spotlights.py:150-177).
//
// A cold path of the reset / step kernels: when mg_step_io.debug_obs is set, the team that just produced an env's
// observation also paints that env's 84x84 debug surface — from the same frame description plus what the reference
// keeps across reset() (sprite handle, cloned command queue, spotlight surface) — scaled 4x (`transform.scale` to
// 336x336 = exact pixel replication) straight into global memory, in the [y][x][c] layout `render()` returns.
// Semantics: as if `env.render()` were called exactly once after every reset() and step() (the reference's render
// pops a clone of the command queue per call).
#pragma once
#include "../../include/memgym_b200.h"
#include "mg_common.h"
#include "mg_frame.h"
#include "mg_mm_logic.h"
#include "mg_mp_logic.h"
#include "mg_ep_logic.h"
#include "mg_ss_logic.h"

#define MG_DEBUG_ROW_WORDS (MG_DEBUG_DIM * 3 / 4)  // 252 32-bit words per output row

// pixel (x, y) of the 84x84 surface -> its 4x4 block (four rows of 12 bytes = three aligned words each)
MG_HD void mg_dbg_put(uint8_t* out, int x, int y, uint32_t col) {
    const uint32_t r = col & 255u, g = (col >> 8) & 255u, b = (col >> 16) & 255u;
    const uint32_t w0 = r | (g << 8) | (b << 16) | (r << 24), w1 = g | (b << 8) | (r << 16) | (g << 24),
                   w2 = b | (r << 8) | (g << 16) | (b << 24);
    uint32_t* o = reinterpret_cast<uint32_t*>(out) + (size_t)(4 * y) * MG_DEBUG_ROW_WORDS + 3 * x;
    for (int q = 0; q < 4; ++q) {
        o[0] = w0, o[1] = w1, o[2] = w2;
        o += MG_DEBUG_ROW_WORDS;
    }
}

// colour-key blit of a sparse sprite with its top-left at (ax, ay), clipped
template <class Ctx>
MG_HD void mg_dbg_sprite(Ctx& c, uint8_t* out, const MgAssetView& a, int slot, int ax, int ay) {
    const MgSprite sp = a.hdr->sprite[slot];
    const uint32_t* e = a.entries + sp.first;
    for (int i = c.lane; i < (int)sp.count; i += c.nl) {
        const uint32_t v = e[i];
        const int x = ax + (int)(v & 255u), y = ay + (int)((v >> 8) & 255u);
        if ((unsigned)x < (unsigned)MG_W && (unsigned)y < (unsigned)MG_W) mg_dbg_put(out, x, y, mg_pal(a, (uint8_t)((v >> 16) & 255u)));
    }
}

// ---- Mortar Mayhem family: arena, sprite handle, next entry of the cloned queue, green ring on the target tile --------
template <class Ctx>
MG_HD_NOINLINE void mg_mm_debug(Ctx c, uint8_t* out, const MgAssetView& a, const MgMMK& k, const MgMMFrame& f,
                                const MgMMState& s) {
    const int origin = k.origin, extent = MG_TILE * k.size;
    for (int p = c.lane; p < MG_W * MG_W; p += c.nl) {
        const int y = p / MG_W, x = p - MG_W * y;  // consecutive lanes: consecutive 12-byte segments of a row
        const unsigned u = (unsigned)(x - origin), v = (unsigned)(y - origin);
        uint32_t col = 0u;
        if (u < (unsigned)extent && v < (unsigned)extent) {
            const unsigned i = u / 14u, j = v / 14u, pu = u - 14u * i, pv = v - 14u * j;
            const bool edge = pu == 0u || pu == 13u || pv == 0u || pv == 13u;
            const bool red = f.red && !((int)i == f.tx && (int)j == f.ty);
            col = red ? (edge ? MG_RGB_TILE_EDGE_ON : MG_RGB_TILE_FILL_ON) : (edge ? MG_RGB_TILE_EDGE : MG_RGB_TILE_FILL);
        }
        mg_dbg_put(out, x, y, col);
    }
    c.sync();
    // rotated_agent_surface / rect: what the last step drew; reset() does not refresh it (the first frames of an
    // instance fall back to get_rotated_sprite(0) = the frame's own sprite)
    const int half = a.hdr->agent_dim >> 1;
    const int ax = s.has_draw ? s.draw_cx : f.sx, ay = s.has_draw ? s.draw_cy : f.sy, arot = s.has_draw ? s.draw_rot : f.srot;
    mg_dbg_sprite(c, out, a, MG_SPR_AGENT0 + arot, ax - half, ay - half);
    c.sync();
    if (s.show_pos < s.show_len) {  // `if self._command_visualization:` pop the clone
        const int icon = mg_mm_show_entry(s, (int)s.show_pos - 1 + (int)s.clone_ahead);
        if (icon >= 0) {
            mg_dbg_sprite(c, out, a, MG_SPR_ICON0 + icon, a.hdr->icon_x, a.hdr->icon_y);
            c.sync();
        }
    }
    const MgSprite ring = a.hdr->sprite[MG_SPR_RING];
    const int tcx = MG_TILE * s.tx + origin + MG_TILE / 2, tcy = MG_TILE * s.ty + origin + MG_TILE / 2;
    mg_dbg_sprite(c, out, a, MG_SPR_RING, tcx - (ring.w >> 1), tcy - (ring.h >> 1));
}

// ---- Mystery Path: whole path (goal green, origin blue, rest white), walls red, agent, fall-off cross ------------------
template <class Ctx>
MG_HD_NOINLINE void mg_mp_debug(Ctx c, uint8_t* out, const MgAssetView& a, const MgMPDraw& d, const MgMPState& s) {
    for (int p = c.lane; p < MG_W * MG_W; p += c.nl) {
        const int y = p / MG_W, x = p - MG_W * y;
        const int i = x / MG_MP_TILE, j = y / MG_MP_TILE, id = i * MG_MP_CELLS + j;
        uint32_t col = 0u;
        if ((s.wall_mask >> id) & 1ull)
            col = 0x000000FFu;  // (255, 0, 0)
        else if ((s.path_mask >> id) & 1ull)
            col = (i == (int)s.ex && j == (int)s.ey) ? 0x0000FF00u : ((i == (int)s.sx && j == (int)s.sy) ? 0x00FF0000u : 0x00FFFFFFu);
        mg_dbg_put(out, x, y, col);
    }
    c.sync();
    const int half = a.hdr->agent_dim >> 1;
    mg_dbg_sprite(c, out, a, MG_SPR_AGENT0 + d.f.srot, d.f.sx - half, d.f.sy - half);
    if (d.f.cross) {
        c.sync();
        mg_dbg_sprite(c, out, a, MG_SPR_CROSS, d.f.cross_x, d.f.cross_y);
    }
}

// ---- Endless Mystery Path: background, the WHOLE path generated so far (white, alpha 200), agent, cross, stamina bar ----
// EndlessMysteryPath.gen_surface (pygame_assets.py:544-557) is colour-keyed and blended with alpha 200 at (-camera_x, 0).
template <class Ctx>
MG_HD_NOINLINE void mg_ep_debug(Ctx c, uint8_t* out, const MgAssetView& a, const MgEPK& k, const MgEPFrame& f,
                                const MgEPState& s) {
    const int x0 = (int)(-s.camera_x);  // blit destination: floats are C-cast
    for (int p = c.lane; p < MG_W * MG_W; p += c.nl) {
        const int y = p / MG_W, x = p - MG_W * y;
        uint32_t r = 0u, g = 0u, b = 0u;
        if (f.bg_off > -1000) {  // icy tiles with a 1-px frame (draw_column_tile_surface), columns starting at bg_off
            const unsigned u = (unsigned)(x - f.bg_off) % 12u, v = (unsigned)y % 12u;
            const bool edge = u == 0u || u == 11u || v == 0u || v == 11u;
            r = edge ? 210u : 125u, g = edge ? 210u : 177u, b = edge ? 210u : 250u;
        }
        const int sxp = x - x0;
        if (sxp >= 0 && mg_ep_on_path(s, sxp / 12, y / 12)) {  // d + trunc((255 - d) * 200 / 255)
            r += (255u - r) * 200u / 255u;
            g += (255u - g) * 200u / 255u;
            b += (255u - b) * 200u / 255u;
        }
        mg_dbg_put(out, x, y, r | (g << 8) | (b << 16));
    }
    c.sync();
    mg_dbg_sprite(c, out, a, MG_SPR_AGENT0 + f.srot, f.sx, f.sy);
    if (f.cross) {
        c.sync();
        mg_dbg_sprite(c, out, a, MG_SPR_CROSS, f.cross_x, f.cross_y);
    }
    c.sync();
    const int bar_h = mg_ep_bar_height(s.stamina, k.p.stamina_level, false);  // always drawn here (:181)
    for (int i = c.lane; i < 4 * MG_W; i += c.nl) {
        const int x = MG_W - 4 + i / MG_W, y = i % MG_W;
        mg_dbg_put(out, x, y, y < bar_h ? 0x000000FFu : 0x0000FF00u);
    }
}

// ---- Searing Spotlights: background, spotlight surface (current alpha), exit, coins, sprite handle, top bar -------------
template <class Ctx>
MG_HD_NOINLINE void mg_ss_debug(Ctx c, uint8_t* out, const MgAssetView& a, const MgSSDrawK& k, const MgSSFrame& f,
                                const MgSSState& s, bool endless) {
    const uint32_t A = f.alpha;
    const int8_t* rows = (const int8_t*)(a.base + a.hdr->disc_off);
    const int nd = A ? (int)f.n_discs : 0;
    for (int p = c.lane; p < MG_W * MG_W; p += c.nl) {
        const int y = p / MG_W, x = p - MG_W * y;
        uint32_t col = mg_ss_bg(f.bg_kind, x, y);
        if (A) {
            bool lit = false;
            for (int q = 0; q < nd && !lit; ++q) lit = mg_disc_has(rows, f.disc_r[q], f.disc_x[q], f.disc_y[q], x, y);
            if (!lit) col = mg_ss_dim(col, A);
        }
        mg_dbg_put(out, x, y, col);
    }
    c.sync();
    if (k.border && A && a.hdr->ring_off) {
        // black_background: a white 1-px outline per spotlight (pygame_assets.py:112-113) unless a later disc covers it
        const uint16_t* start = (const uint16_t*)(a.base + a.hdr->ring_off);
        const int8_t* ring = (const int8_t*)(start + MG_MAX_DISC_R + 2);
        for (int q = 0; q < nd; ++q) {
            const int r = f.disc_r[q], n = (int)start[r + 1] - (int)start[r];
            const int8_t* px = ring + 2 * (int)start[r];
            for (int i = c.lane; i < n; i += c.nl) {
                const int x = f.disc_x[q] + (int)px[2 * i], y = f.disc_y[q] + (int)px[2 * i + 1];
                if ((unsigned)x >= (unsigned)MG_W || (unsigned)y >= (unsigned)MG_W) continue;
                bool covered = false;
                for (int e = q + 1; e < nd && !covered; ++e) covered = mg_disc_has(rows, f.disc_r[e], f.disc_x[e], f.disc_y[e], x, y);
                if (covered) continue;
                const uint32_t under = mg_ss_bg(f.bg_kind, x, y);
                uint32_t cr = under & 255u, cg = (under >> 8) & 255u, cb = (under >> 16) & 255u;
                cr += (255u - cr) * A / 255u;
                cg += (255u - cg) * A / 255u;
                cb += (255u - cb) * A / 255u;
                mg_dbg_put(out, x, y, cr | (cg << 8) | (cb << 16));
            }
            c.sync();  // a later outline may cross an earlier one: keep the reference's draw order
        }
    }
    if (!endless) {
        const int slot = f.exit_open ? MG_SPR_EXIT_OPEN : MG_SPR_EXIT_CLOSED;
        const MgSprite sp = a.hdr->sprite[slot];
        mg_dbg_sprite(c, out, a, slot, f.exit_x - (sp.w >> 1), f.exit_y - (sp.h >> 1));
        c.sync();
    }
    {
        const MgSprite sp = a.hdr->sprite[MG_SPR_COIN];
        for (int q = 0; q < (int)f.n_coins; ++q) {
            mg_dbg_sprite(c, out, a, MG_SPR_COIN, f.coins[q][0] - (sp.w >> 1), f.coins[q][1] - (sp.h >> 1));
            c.sync();
        }
    }
    const int half = a.hdr->agent_dim >> 1;
    const int ax = s.hand_has ? s.hand_x : f.sx, ay = s.hand_has ? s.hand_y : f.sy, arot = s.hand_has ? s.hand_rot : f.srot;
    mg_dbg_sprite(c, out, a, MG_SPR_AGENT0 + arot, ax - half, ay - half);
    c.sync();
    for (int p = c.lane; p < MG_W * 4; p += c.nl) {
        const int y = p / MG_W, x = p - MG_W * y;
        mg_dbg_put(out, x, y, mg_ss_bar(k, f, x));
    }
}
