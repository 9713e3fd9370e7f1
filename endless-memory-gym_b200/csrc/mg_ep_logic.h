// mg_ep_logic.h — Endless Mystery Path: segment generation, reset and the step state machine for ONE env record.
//
// Reference paths replaced:
//   EndlessMysteryPathEnv.reset/step/_draw_past_path   memory_gym/endless_mystery_path.py:192-268, 270-444, 111-135
//   EndlessMysteryPath.add_path_segment                memory_gym/pygame_assets.py:559-604
//   MysteryPath.__init__ (walls + noisy A*)            memory_gym/pygame_assets.py:606-736 (mg_mp_generate)
//   CharacterController.step without a boundary rect   memory_gym/character_controller.py:89-130
// The path is kept per segment as a 49-bit cell mask plus a 2-bit "where is next_node" code per cell, so
// next_node is one lookup and previous_node a test of the (at most four) neighbours.
// Deliberate deviations: the path is capped at MG_EP_MAXSEG segments (360 columns) and fall_off_locations at
// MG_EP_MAXFALL entries; reaching either cap ends the episode (done).
#pragma once
#include <math.h>

#include "../../include/memgym_b200.h"
#include "mg_common.h"
#include "mg_mm_logic.h"  // kinematics
#include "mg_mp_logic.h"  // mg_mp_generate
#include "mg_pcg64.h"

#define MG_EP_SEG 8  // segment_length: 7 columns + the transition tile

struct MgEPK {
    mg_ep_params p;
    int32_t radius;          // int(25 * agent_scale)
    int32_t step_axis;       // int(agent_speed): the only velocities are (v, 0), (0, -v), (0, v)
    int32_t agent_draw_x;    // int(agent.rect.topleft[0] - camera_offset), constant per episode (:222)
    int32_t depth;           // int(camera_offset_scale): how many columns back the past path is drawn
    int32_t agent_half;
    int32_t pad;
    double camera_offset;    // -tile_dim * clamp(camera_offset_scale, 0, 5.5)
};

MG_HD int mg_ep_start_y(const MgEPSegment& g) { return (int)((g.path >> 56) & 7u); }
MG_HD int mg_ep_end_y(const MgEPSegment& g) { return (int)((g.path >> 60) & 7u); }
MG_HD int mg_ep_next_code(const MgEPSegment& g, int cell) { return (int)((g.next[cell >> 5] >> (2 * (cell & 31))) & 3u); }

// (x, y) is a node of the path: transition tiles included
MG_HD bool mg_ep_on_path(const MgEPState& s, int x, int y) {
    if (x < 0 || y < 0 || y >= MG_MP_CELLS) return false;
    const int sg = x / MG_EP_SEG, lx = x % MG_EP_SEG;
    if (sg >= s.n_segments) return false;
    const MgEPSegment& g = s.seg[sg];
    if (lx == MG_EP_SEG - 1) return y == mg_ep_end_y(g);
    return ((g.path >> mg_mp_id(lx, y)) & 1ull) != 0;
}

// next_node of node (x, y); false past the end of the generated path
MG_HD bool mg_ep_next(const MgEPState& s, int x, int y, int& nx, int& ny) {
    const int sg = x / MG_EP_SEG, lx = x % MG_EP_SEG;
    nx = x, ny = y;
    if (lx == MG_EP_SEG - 1) {  // transition tile -> first node of the next segment
        nx = x + 1;
        return sg + 1 < s.n_segments;
    }
    const int code = mg_ep_next_code(s.seg[sg], mg_mp_id(lx, y));
    if (code == 0)
        nx = x + 1;
    else if (code == 1)
        ny = y - 1;
    else if (code == 2)
        ny = y + 1;
    else
        nx = x - 1;
    return true;
}

// previous_node of node (x, y); false for the very first node
MG_HD bool mg_ep_prev(const MgEPState& s, int x, int y, int& px, int& py) {
    const int sg = x / MG_EP_SEG, lx = x % MG_EP_SEG;
    const MgEPSegment& g = s.seg[sg];
    px = x, py = y;
    if (lx == MG_EP_SEG - 1) {  // transition tile <- the segment's end node
        px = x - 1;
        return true;
    }
    if (lx == 0 && y == mg_ep_start_y(g)) {  // first node <- the previous segment's transition tile
        px = x - 1;
        return sg > 0;
    }
    // the neighbour on the path whose next_node is this cell
    const int dx[4] = {-1, 0, 0, 1}, dy[4] = {0, 1, -1, 0}, want[4] = {0, 1, 2, 3};  // it lies west / south / north / east
    for (int q = 0; q < 4; ++q) {
        int ax = lx + dx[q], ay = y + dy[q];
        if (ax < 0 || ax >= MG_MP_CELLS || ay < 0 || ay >= MG_MP_CELLS) continue;
        int cell = mg_mp_id(ax, ay);
        if (((g.path >> cell) & 1ull) && mg_ep_next_code(g, cell) == want[q]) {
            px = sg * MG_EP_SEG + ax;
            py = ay;
            return true;
        }
    }
    return false;
}

// EndlessMysteryPath.add_path_segment: draws the start row (first segment only), the end row, then the A* path.
// Called by every lane of the warp (mg_mp_generate_coop); the stream in the record advances even when it fails.
template <class Ctx>
MG_HD_NOINLINE bool mg_ep_add_segment_coop(Ctx c, MgEPState& s, MgGenMem& m, const MgGenCta& tab) {
    const int ns = s.n_segments;
    if (ns >= MG_EP_MAXSEG) return false;
    const int g7 = MG_MP_CELLS;
    MgRngR r = mg_rr_load(s.rng);
    const int sy = ns == 0 ? mg_rr_integers(r, 0, g7) : mg_ep_end_y(s.seg[ns - 1]);
    const int ey = mg_rr_integers(r, 0, g7);
    c.sync();  // every lane has read the record (stream, segment count) before lane 0 rewrites it below
    uint64_t nx[2] = {0, 0};
    const uint64_t mask = mg_mp_generate_coop(c, r, 0, sy, g7 - 1, ey, tab, nx);
    const uint64_t n0 = nx[0], n1 = nx[1];  // the end node's next is the transition tile: east = 0
    if (c.lane == 0) {
        mg_rr_store(r, s.rng);
        if (mask) {
            MgEPSegment& g = s.seg[ns];
            g.path = mask | ((uint64_t)sy << 56) | ((uint64_t)ey << 60);
            g.reward_visited = 0;
            g.stamina_visited = 0;
            g.next[0] = n0;
            g.next[1] = n1;
            s.n_segments = ns + 1;
        }
    }
    c.sync();
    return mask != 0;
}

MG_HD void mg_ep_update_direction(MgEPState& s) {
    int nx, ny;
    if (!mg_ep_next(s, s.cur_x, s.cur_y, nx, ny)) return;
    if (nx - s.cur_x == 1)
        s.dir = 0;
    else if (ny - (int)s.cur_y == -1)
        s.dir = 1;
    else if (ny - (int)s.cur_y == 1)
        s.dir = 2;
    // a westward link keeps the previous direction
}

MG_HD void mg_ep_ground_truth(const MgEPState& s, double* gt) {
    gt[0] = s.dir == 0 ? 1.0 : 0.0;
    gt[1] = s.dir == 1 ? 1.0 : 0.0;
    gt[2] = s.dir == 2 ? 1.0 : 0.0;
}

// python floor division of ints by 12
MG_HD int mg_ep_floordiv12(int v) { return v >= 0 ? v / 12 : -((11 - v) / 12); }

// frame description: _draw_surfaces + _draw_past_path (:111-160)
MG_HD void mg_ep_frame(const MgEPK& k, const MgEPState& s, int bar_h, MgEPFrame& f) {
    f.sx = k.agent_draw_x;
    f.sy = s.cy - k.agent_half;
    f.srot = s.rot;
    f.cross = s.cross_on;
    f.cross_x = s.cross_x;
    f.cross_y = s.cross_y;
    f.bar_h = k.p.show_stamina ? bar_h : -1;
    f.bg_off = k.p.show_background ? -s.bg_scroll : -1000;  // icy columns start at i * 12 + bg_scroll - 12 (:141-143)
    f.n_tiles = 0;
    if (!k.p.show_past_path) return;
    int x = mg_ep_floordiv12(s.cx) - 1;
    if (x < 0) return;
    const int past_x = x - k.depth > 0 ? x - k.depth : 0;
    int nx = s.cur_x, ny = s.cur_y;
    bool have = mg_ep_prev(s, s.cur_x, s.cur_y, nx, ny);
    while (x >= past_x && x >= 0) {
        if (!have) break;
        x = nx;
        const int y = ny;
        // pygame.draw.rect((draw_x, y * 12, 12, 12)): float coordinates are C-cast
        const int draw_x = (int)((double)(x * 12) - s.camera_x);
        if (draw_x > -12 && draw_x < MG_W && f.n_tiles < MG_EP_MAXTILES) {
            f.tile_x[f.n_tiles] = (int8_t)draw_x;
            f.tile_y[f.n_tiles] = (uint8_t)y;
            f.n_tiles++;
        }
        if (x == past_x) break;
        have = mg_ep_prev(s, x, y, nx, ny);
    }
}

MG_HD int mg_ep_bar_height(int stamina, int max_visible, bool at_reset) {
    // reset uses max(), step uses min() (:252 vs :427)
    int v = at_reset ? (stamina > max_visible ? stamina : max_visible) : (stamina < max_visible ? stamina : max_visible);
    return (int)((double)MG_W * (1.0 - ((double)v / (double)max_visible)));
}

template <class Ctx>
MG_HD void mg_ep_reset_coop(Ctx& c, const MgEPK& k, MgEPState& s, MgEPFrame& f, double* gt, MgGenMem& m,
                                     const MgGenCta& tab) {
    c.stamp(12);
    if (c.lane == 0) {
        s.t = 0;
        s.ep_len = 0;
        s.ep_reward = 0.0;
        s.n_segments = 0;
    }
    c.sync();
    bool ok = true;
    for (int q = 0; q < 3 && ok; ++q) ok = mg_ep_add_segment_coop(c, s, m, tab);
    c.stamp(13);
    if (c.lane == 0) {
        const int sy = ok ? mg_ep_start_y(s.seg[0]) : 0;
        if (ok) s.seg[0].reward_visited |= 1ull << mg_mp_id(0, sy);  // the first node yields no reward (:529)
        s.camera_x = k.camera_offset;
        s.bg_scroll = 0;
        s.rot = 6;  // CharacterController(..., 270): sprite 270 // 45
        s.cx = 0 * MG_MP_TILE + k.radius;
        s.cy = sy * MG_MP_TILE + k.radius;
        s.cur_x = 0;
        s.cur_y = (uint8_t)sy;
        s.dir = 0;
        if (ok) mg_ep_update_direction(s);
        s.off_path = 0;
        s.num_fails = 0;
        s.n_fall = 0;
        s.stamina = k.p.stamina_level;
        s.max_x = 0;
        s.tiles_visited = 0;
        s.cross_on = 0;
        s.cross_x = s.cross_y = 0;
        s.live = ok ? 1 : 0;  // "No valid path found": the reference raises; here the env stays finished
        mg_ep_frame(k, s, mg_ep_bar_height(s.stamina, k.p.stamina_level, true), f);
        if (gt) mg_ep_ground_truth(s, gt);
    }
    c.sync();
    c.stamp(14);
}

// step(), first half (one lane): the move. Returns true when the path needs another segment before the second
// half runs (the warp then generates it cooperatively).
MG_HD bool mg_ep_step_pre(const MgEPK& k, MgEPState& s, int action) {
    const int start_y = mg_ep_start_y(s.seg[0]);
    if (!s.off_path) {
        // Discrete(4) -> [0,0], [2,0], [0,1], [0,2] (:283-291); CharacterController.step, no boundary
        int vx = 0;
        if (action == 1) {
            s.rot = 6;
            vx = k.step_axis;
            s.cx += vx;
        } else if (action == 2) {
            s.rot = 0;
            s.cy -= k.step_axis;
        } else if (action == 3) {
            s.rot = 4;
            s.cy += k.step_axis;
        }
        s.camera_x = s.camera_x + (double)vx;
        s.bg_scroll += vx;
        if (vx > 0 && s.bg_scroll >= MG_MP_TILE) s.bg_scroll = s.bg_scroll % vx;  // always 0: a multiple of vx
    } else {
        s.cx = 0 * MG_MP_TILE + k.radius;
        s.cy = start_y * MG_MP_TILE + k.radius;
        s.camera_x = k.camera_offset;
        s.bg_scroll = 0;
    }
    const int sg = mg_ep_floordiv12(s.cx) / MG_EP_SEG;
    return sg > s.n_segments - 2;
}

// step(), second half (one lane). `capped`: a segment was needed but the record is full
MG_HD bool mg_ep_step_post(const MgEPK& k, MgEPState& s, bool capped, double& reward_out, double* info,
                                    MgEPFrame& f, double* gt) {
    double reward = 0.0;
    bool done = false;
    const int start_y = mg_ep_start_y(s.seg[0]);
    const int px = mg_ep_floordiv12(s.cx), py = mg_ep_floordiv12(s.cy);
    const int sg = px / MG_EP_SEG;
    const bool on_path = mg_ep_on_path(s, px, py) && sg < s.n_segments;
    if (on_path) {
        s.cur_x = px;
        s.cur_y = (uint8_t)py;
        MgEPSegment& g = s.seg[sg];
        const int lx = px % MG_EP_SEG;
        const uint64_t bit = 1ull << (lx == MG_EP_SEG - 1 ? 49 : mg_mp_id(lx, py));
        const bool is_start = px == 0 && py == start_y;
        if (!(g.reward_visited & bit) && !is_start) {
            reward += k.p.reward_path_progress;
            s.tiles_visited++;
            g.reward_visited |= bit;
        }
        if (!(g.stamina_visited & bit) && !is_start) {
            reward += k.p.reward_path_progress_dense;
            s.stamina = k.p.stamina_level;
            g.stamina_visited |= bit;
        }
        s.cross_on = 0;
        s.off_path = 0;
    } else {
        reward += k.p.reward_fall_off;
        s.num_fails++;
        if (k.p.visual_feedback) s.cross_on = 1;
        s.off_path = 1;
        if (px < s.max_x) {
            done = true;
        } else {
            bool found = false;
            for (int q = 0; q < (int)s.n_fall; ++q)
                if (s.fall[q][0] == px && s.fall[q][1] == py) {
                    found = true;
                    break;
                }
            if (found) {
                done = true;
            } else if (s.n_fall < MG_EP_MAXFALL) {
                s.fall[s.n_fall][0] = (int16_t)px;
                s.fall[s.n_fall][1] = (int16_t)py;
                s.n_fall++;
            } else {
                capped = true;
            }
        }
        for (int q = 0; q < s.n_segments; ++q) s.seg[q].stamina_visited = 0;
        s.stamina = k.p.stamina_level;
    }
    // fall_off_rect.center = (cx - camera_x, cy): 10x10 rect
    s.cross_x = (int16_t)((int)((double)s.cx - s.camera_x) - 5);
    s.cross_y = (int16_t)(s.cy - 5);
    reward += k.p.reward_step;
    s.stamina--;
    if (s.stamina == 0) done = true;
    s.t++;
    if (s.t == k.p.max_steps) done = true;
    if (capped) done = true;
    mg_ep_update_direction(s);
    if (px > s.max_x && on_path) s.max_x = px;
    s.ep_reward = s.ep_reward + reward;
    s.ep_len++;
    reward_out = reward;
    mg_ep_frame(k, s, mg_ep_bar_height(s.stamina, k.p.stamina_level, false), f);
    double g3[3];
    mg_ep_ground_truth(s, g3);
    if (gt)
        for (int i = 0; i < 3; ++i) gt[i] = g3[i];
    if (done) {
        s.live = 0;
        if (info) {
            info[MG_INFO_REWARD] = s.ep_reward;
            info[MG_INFO_LENGTH] = (double)s.ep_len;
            info[MG_INFO_SUCCESS] = 0.0;
            info[MG_INFO_A] = (double)s.num_fails;
            info[MG_INFO_B] = (double)s.max_x;
            info[MG_INFO_GT0] = g3[0];
            info[MG_INFO_GT1] = g3[1];
            info[MG_INFO_GT2] = g3[2];
            info[MG_INFO_GT3] = 0.0;
            info[MG_INFO_C] = (double)s.tiles_visited;
        }
    }
    return done;
}
