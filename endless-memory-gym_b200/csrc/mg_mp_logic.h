// mg_mp_logic.h — Mystery Path: per-reset procedural path generation (walls + noisy A*) and the
// step state machine for ONE env record.
//
// Reference paths replaced:
//   MysteryPathEnv.reset/step        memory_gym/mystery_path.py:130-200, 202-276
//   MysteryPath.__init__ / Node      memory_gym/pygame_assets.py:438-481, 606-736
//   CharacterController.step         memory_gym/character_controller.py:89-146
// The order and number of RNG draws is part of the contract.
#pragma once
#include <math.h>

#include "../../include/memgym_b200.h"
#include "mg_common.h"
#include "mg_frame.h"
#include "mg_mm_logic.h"  // kinematics shared with Mortar Mayhem
#include "mg_pcg64.h"

struct MgMPK {
    mg_mp_params p;
    int32_t radius;
    int32_t step_axis, step_diag;
    int32_t pad;
};

#define MG_MP_CELLS 7
#define MG_MP_TILE 12  // screen_dim / grid_dim (mystery_path.py:76-78)
#define MG_MP_NODES (MG_MP_CELLS * MG_MP_CELLS)

MG_HD int mg_mp_id(int i, int j) { return i * MG_MP_CELLS + j; }

// 4-neighbours in Node.add_neighbors order (+x, -x, +y, -y), pygame_assets.py:465-472
MG_HD int mg_mp_neighbours(int i, int j, int* out) {
    int c = 0;
    if (i < MG_MP_CELLS - 1) out[c++] = mg_mp_id(i + 1, j);
    if (i > 0) out[c++] = mg_mp_id(i - 1, j);
    if (j < MG_MP_CELLS - 1) out[c++] = mg_mp_id(i, j + 1);
    if (j > 0) out[c++] = mg_mp_id(i, j - 1);
    return c;
}

// MysteryPath.__init__: returns the path as a 49-bit mask, or 0 if A* ran out of open nodes
// (the reference raises Exception("No valid path found"), pygame_assets.py:724)
// `order` (optional, MG_MP_NODES entries) receives the path's cell ids from the end node back to the start node.
MG_HD_NOINLINE uint64_t mg_mp_generate(MgRng& r, int sx, int sy, int ex, int ey, uint8_t* order = nullptr,
                                       int* n_order = nullptr) {
    const int n = MG_MP_CELLS;
    uint64_t wall = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            if (i > 0 && i < n - 2 && j > 0 && j < n - 2)
                if (mg_integers(r, 0, 100) < 33) wall |= 1ull << mg_mp_id(i, j);
    const int start = mg_mp_id(sx, sy), end = mg_mp_id(ex, ey);
    int nb[4];
    uint64_t near_ends = 0;
    int c = mg_mp_neighbours(sx, sy, nb);
    for (int q = 0; q < c; ++q) near_ends |= 1ull << nb[q];
    c = mg_mp_neighbours(ex, ey, nb);
    for (int q = 0; q < c; ++q) near_ends |= 1ull << nb[q];
    // outer-ring wall candidates (:650-669)
    uint8_t ring[24];
    int nring = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            if (!(i == 0 || i == n - 1 || j == 0 || j == n - 1)) continue;
            int id = mg_mp_id(i, j);
            if (id == start || id == end || ((near_ends >> id) & 1)) continue;
            bool blocked = false;
            for (int di = -1; di <= 1 && !blocked; ++di)
                for (int dj = -1; dj <= 1; ++dj) {
                    if (!di && !dj) continue;
                    int a = i + di, b = j + dj;
                    if (a < 0 || a >= n || b < 0 || b >= n) continue;
                    if ((wall >> mg_mp_id(a, b)) & 1) {
                        blocked = true;
                        break;
                    }
                }
            if (!blocked) ring[nring++] = (uint8_t)id;
        }
    int picks = mg_integers(r, 0, 2) ? 8 : 4;  // choice([4, 8])
    for (int q = 0; q < picks; ++q) {
        if (nring > 0) {
            int w = (int)mg_integers(r, 0, nring);  // choice(outer_nodes); one candidate draws nothing
            wall |= 1ull << ring[w];
            for (int t = w; t + 1 < nring; ++t) ring[t] = ring[t + 1];
            nring--;
        }
    }
    // noisy A* (:681-724)
    int g[MG_MP_NODES];
    double f[MG_MP_NODES];
    int8_t prev[MG_MP_NODES];
    uint8_t open[MG_MP_NODES];
    int nopen = 0;
    uint64_t in_open = 0, closed = 0;
    g[start] = 0;
    f[start] = 0.0;
    prev[start] = -1;
    open[nopen++] = (uint8_t)start;
    in_open |= 1ull << start;
    for (;;) {
        if (nopen == 0) return 0;
        int w = 0;
        for (int q = 0; q < nopen; ++q)
            if (f[open[q]] < f[open[0]]) {  // first node better than the head, not the best one
                w = q;
                break;
            }
        int cur = open[w];
        if (cur == end) {
            uint64_t path = 0;
            int cnt = 0;
            for (int at = end; at >= 0; at = prev[at]) {
                path |= 1ull << at;
                if (order) order[cnt] = (uint8_t)at;
                cnt++;
            }
            if (n_order) *n_order = cnt;
            return path;
        }
        for (int t = w; t + 1 < nopen; ++t) open[t] = open[t + 1];
        nopen--;
        in_open &= ~(1ull << cur);
        closed |= 1ull << cur;
        int cn = mg_mp_neighbours(cur / n, cur % n, nb);
        for (int q = 0; q < cn; ++q) {
            int v = nb[q];
            if (((closed >> v) & 1) || ((wall >> v) & 1)) continue;
            int cost = g[cur] + (int)mg_integers(r, 1, 9);
            bool new_path = false;
            if ((in_open >> v) & 1) {
                if (cost < g[v]) new_path = true;  // g_cost is not updated (reference bug, :712)
            } else {
                g[v] = cost;
                new_path = true;
                open[nopen++] = (uint8_t)v;
                in_open |= 1ull << v;
            }
            if (new_path) {
                int dx = v / n - ex, dy = v % n - ey;
                f[v] = (double)g[v] + sqrt((double)(dx * dx + dy * dy));
                prev[v] = (int8_t)cur;
            }
        }
    }
}

MG_HD void mg_mp_draw(const MgMPK& k, const MgMPState& s, int cross, MgMPDraw& d) {
    d.f.sx = s.cx;
    d.f.sy = s.cy;
    d.f.srot = s.rot;
    d.f.cross = cross;
    d.f.cross_x = s.cross_x;
    d.f.cross_y = s.cross_y;
    d.show_origin = k.p.show_origin;
    d.show_goal = k.p.show_goal;
    d.ox = s.sx;
    d.oy = s.sy;
    d.gx = s.ex;
    d.gy = s.ey;
}

MG_HD_NOINLINE void mg_mp_reset_env(const MgMPK& k, MgMPState& s, MgMPDraw& d) {
    const int g = MG_MP_CELLS;
    MgRng r = s.rng;
    s.t = 0;
    s.ep_len = 0;
    s.ep_reward = 0.0;
    int side = mg_choice(r, k.p.cardinal_origin_choice);
    int a = (int)mg_integers(r, 0, g);
    int b = (int)mg_integers(r, 0, g);
    int sx, sy, ex, ey;
    if (side == 0) {
        sx = 0, sy = a, ex = g - 1, ey = b;
    } else if (side == 1) {
        sx = g - 1, sy = a, ex = 0, ey = b;
    } else if (side == 2) {
        sx = a, sy = 0, ex = b, ey = g - 1;
    } else {
        sx = a, sy = g - 1, ex = b, ey = 0;
    }
    s.sx = (uint8_t)sx;
    s.sy = (uint8_t)sy;
    s.ex = (uint8_t)ex;
    s.ey = (uint8_t)ey;
    s.path_mask = mg_mp_generate(r, sx, sy, ex, ey);
    s.visited = 0;
    s.rot = 0;
    s.cx = (int16_t)(sx * MG_MP_TILE + k.radius);
    s.cy = (int16_t)(sy * MG_MP_TILE + k.radius);
    s.gx = (int8_t)sx;  // Grid: normalize(start * 12 + int(25 * SCALE)) (mystery_path_grid.py:186-188)
    s.gy = (int8_t)sy;
    s.off_path = 0;
    s.num_fails = 0;
    s.cross_on = 0;
    s.cross_x = 0;
    s.cross_y = 0;
    s.live = s.path_mask ? 1 : 0;  // no path: the reference raises; here the env stays finished
    s.rng = r;
    mg_mp_draw(k, s, 0, d);
}

MG_HD bool mg_mp_step_env(const MgMPK& k, MgMPState& s, int a0, int a1, double& reward_out, double* info,
                          MgMPDraw& d) {
    double reward = 0.0;
    bool done = false;
    int success = 0;
    int cx = s.cx, cy = s.cy, rot = s.rot;
    if (k.p.variant == MG_MP_GRID) {
        // GridCharacterController.step (character_controller.py:177-210): 1 / 2 rotate by +-90, 3 forward;
        // cell centres 12 i + 6 (MysteryPath.to_grid, pygame_assets.py:764-778)
        int gx = s.gx, gy = s.gy;
        if (s.off_path) {  // reset_position(start) then agent.step(0) (mystery_path_grid.py:218-221)
            gx = s.sx;
            gy = s.sy;
            a0 = 0;
        }
        if (a0 == 1) rot = (rot + 2) & 7;
        if (a0 == 2) rot = (rot + 6) & 7;
        if (a0 == 3) {
            int face = rot >> 1, last = MG_MP_CELLS - 1;
            if (face == 0 && gy > 0)
                gy--;
            else if (face == 3 && gx < last)
                gx++;
            else if (face == 2 && gy < last)
                gy++;
            else if (face == 1 && gx > 0)
                gx--;
        }
        s.gx = (int8_t)gx;
        s.gy = (int8_t)gy;
        cx = MG_MP_TILE * gx + MG_MP_TILE / 2;
        cy = MG_MP_TILE * gy + MG_MP_TILE / 2;
    } else {
        if (s.off_path) {  // respawn on the origin tile, then a no-op controller step (mystery_path.py:216-220)
            cx = s.sx * MG_MP_TILE + k.radius;
            cy = s.sy * MG_MP_TILE + k.radius;
            a0 = 0;
            a1 = 0;
        }
        mg_move_clamped(cx, cy, rot, a0, a1, k.step_axis, k.step_diag, k.radius, 0, 0, MG_W, MG_W);
    }
    s.cx = (int16_t)cx;
    s.cy = (int16_t)cy;
    s.rot = (uint8_t)rot;
    int px = cx / MG_MP_TILE, py = cy / MG_MP_TILE;  // centre is clamped to [radius, 84 - radius]: non-negative
    if (px == s.ex && py == s.ey) {
        reward += k.p.reward_goal;
        done = true;
        success = 1;
    } else {
        int id = mg_mp_id(px, py);
        bool on_path = px < MG_MP_CELLS && py < MG_MP_CELLS && ((s.path_mask >> id) & 1);
        if (on_path) {
            bool is_start = (px == s.sx && py == s.sy);
            if (!((s.visited >> id) & 1) && !is_start) {
                reward += k.p.reward_path_progress;
                s.visited |= 1ull << id;
            }
            s.cross_on = 0;
            s.off_path = 0;
        } else {
            reward += k.p.reward_fall_off;
            s.num_fails++;
            if (k.p.visual_feedback) s.cross_on = 1;
            s.off_path = 1;
        }
        // fall_off_rect.center = agent centre; the cross is 10x10 -> top-left = centre - 5
        s.cross_x = (int16_t)(cx - 5);
        s.cross_y = (int16_t)(cy - 5);
    }
    reward += k.p.reward_step;
    s.t++;
    if (s.t == k.p.max_steps) done = true;
    s.ep_reward = s.ep_reward + reward;
    s.ep_len++;
    reward_out = reward;
    mg_mp_draw(k, s, s.cross_on, d);
    if (done) {
        s.live = 0;
        if (info) {
            info[MG_INFO_REWARD] = s.ep_reward;
            info[MG_INFO_LENGTH] = (double)s.ep_len;
            info[MG_INFO_SUCCESS] = (double)success;
            info[MG_INFO_A] = (double)s.num_fails;
            info[MG_INFO_B] = 0.0;
            info[MG_INFO_GT0] = 0.0;
            info[MG_INFO_GT1] = 0.0;
            info[7] = info[8] = info[9] = 0.0;
        }
    }
    return done;
}
