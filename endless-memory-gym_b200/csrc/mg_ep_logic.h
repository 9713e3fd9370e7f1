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
#include <stddef.h>

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

// ---- path generation ahead of time ("speculation") ---------------------------------------------------------
// Only add_path_segment draws from the env's stream, so between two resets the stream stands still unless the agent
// walks into a new segment. Rollout launches therefore run HELPER WARPS next to the env teams (mg_kernels.cu): after
// a reset the team posts the stream it ended with, a helper generates the NEXT episode's three segments from it, and
// the next reset adopts them if — and only if — the stream still is what the helper started from (bit-exact either
// way; a stale or unfinished job is simply dropped). The hand-over lives in shared memory only.
enum { MG_SPEC_EMPTY = 0, MG_SPEC_REQUESTED = 1, MG_SPEC_BUSY = 2, MG_SPEC_READY = 3 };
struct MgEPSpec {
    uint32_t state;   // MG_SPEC_*: EMPTY -> (team) REQUESTED -> (helper) BUSY -> READY -> (team) EMPTY
    int32_t env;      // single-step launches: the env whose (already written back) record waits for its path
    MgRng before;     // the stream the three segments were (are to be) generated from
    MgRng after;      // ... and the stream after them
    MgEPSegment seg[3];
};
struct MgEPWarpMem {  // per-team scratch of the Endless Mystery Path kernels
    MgGenMem gen;
    int32_t helpers;  // helper warps in this launch (0: none, resets generate their own paths)
    int16_t offload;  // single-step launch: an auto-reset hands its path generation to a helper (mg_ep_reset_offload)
    int16_t pad;
    int32_t cur_env;  // the env this team works on
    int32_t post_env; // >= 0: its record carries a pending path; post the job once the record is written back
    MgEPSpec spec;
};
// what a helper warp works on: the head of an env record (stream, segment count, the first three segments)
#define MG_EP_HELPER_BYTES 384
static_assert(offsetof(MgEPState, seg) + 3 * sizeof(MgEPSegment) <= MG_EP_HELPER_BYTES, "helper scratch too small");

MG_HD uint32_t mg_vload_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
MG_HD void mg_vstore_u32(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }

// EndlessMysteryPath.add_path_segment: draws the start row (first segment only), the end row, then the A* path.
// Called by every lane of the warp (mg_mp_generate_coop); the stream in the record advances even when it fails.
template <class Ctx>
MG_HD_NOINLINE bool mg_ep_add_segment_coop(Ctx c, MgEPState& s, MgGenMem& m, const MgGenCta& tab, int forced_sy = -1) {
    const int ns = s.n_segments;
    if (ns >= MG_EP_MAXSEG) return false;
    const int g7 = MG_MP_CELLS;
    MgRngR r = mg_rr_load(s.rng);
    // (forced_sy: the start row of the first segment was already drawn, see mg_ep_reset_offload)
    const int sy = ns == 0 ? (forced_sy >= 0 ? forced_sy : mg_rr_integers(r, 0, g7)) : mg_ep_end_y(s.seg[ns - 1]);
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

// Team side of the speculation: true if the three segments of this reset were adopted from a helper
template <class Ctx>
MG_HD bool mg_ep_spec_take(Ctx& c, MgEPState& s, MgEPSpec& sp) {
    for (;;) {
        const uint32_t st = c.shfl(mg_vload_u32(&sp.state), 0);  // lane 0's view, so the warp stays uniform
        if (st == MG_SPEC_BUSY) {  // a helper is at it: it finishes within one generation, sooner than we would
            c.nanosleep(256);
            continue;
        }
        if (st == MG_SPEC_REQUESTED) {  // nobody picked it up: take it back (a helper may win the race)
            uint32_t won = 0;
            if (c.lane == 0) won = c.cas_u32(&sp.state, MG_SPEC_REQUESTED, MG_SPEC_EMPTY) == MG_SPEC_REQUESTED ? 1u : 0u;
            if (c.shfl(won, 0)) return false;
            continue;
        }
        if (st != MG_SPEC_READY) return false;
        c.sync();         // lane 0's observation of READY is ordered before every lane's reads below
        c.fence_block();  // the helper's results are visible after its READY
        const bool match = sp.before.state_lo == s.rng.state_lo && sp.before.state_hi == s.rng.state_hi &&
                           sp.before.inc_lo == s.rng.inc_lo && sp.before.inc_hi == s.rng.inc_hi &&
                           sp.before.buf == s.rng.buf && sp.before.has_buf == s.rng.has_buf;
        if (match) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(sp.seg);
            uint32_t* dst = reinterpret_cast<uint32_t*>(s.seg);
            for (int w = c.lane; w < (int)(3 * sizeof(MgEPSegment) / 4); w += c.nl) dst[w] = src[w];
            if (c.lane == 0) {
                s.rng = sp.after;
                s.n_segments = 3;
            }
        }
        c.sync();
        if (c.lane == 0) mg_vstore_u32(&sp.state, MG_SPEC_EMPTY);
        c.sync();
        return match;
    }
}

// Team side, after a reset: ask for the next episode's segments (the slot is EMPTY here, see mg_ep_spec_take)
template <class Ctx>
MG_HD void mg_ep_spec_request(Ctx& c, const MgEPState& s, MgEPSpec& sp) {
    if (c.lane == 0) {
        sp.before = s.rng;
        c.fence_block();
        mg_vstore_u32(&sp.state, MG_SPEC_REQUESTED);
    }
}

// Helper warp: serve the teams of this CTA until all of them have left. `hs` = this helper's MG_EP_HELPER_BYTES.
template <class Ctx>
MG_HD void mg_ep_helper_loop(Ctx& c, uint8_t* hs_bytes, uint8_t* wextra0, int wextra_stride, int nteams, int hidx,
                             const uint32_t* teams_done, MgGenMem& m, const MgGenCta& tab) {
    MgEPState& hs = *reinterpret_cast<MgEPState*>(hs_bytes);  // only its head is touched (n_segments <= 3)
    int next = hidx % nteams;
    for (;;) {
        if (c.shfl(mg_vload_u32(teams_done), 0) >= (uint32_t)nteams) return;
        MgEPSpec* job = nullptr;
        for (int i = 0; i < nteams && !job; ++i) {
            const int t = (next + i) % nteams;
            MgEPSpec& sp = reinterpret_cast<MgEPWarpMem*>(wextra0 + (size_t)t * wextra_stride)->spec;
            if (c.shfl(mg_vload_u32(&sp.state), 0) != MG_SPEC_REQUESTED) continue;
            uint32_t won = 0;
            if (c.lane == 0) won = c.cas_u32(&sp.state, MG_SPEC_REQUESTED, MG_SPEC_BUSY) == MG_SPEC_REQUESTED ? 1u : 0u;
            if (c.shfl(won, 0)) {
                job = &sp;
                next = (t + 1) % nteams;
            }
        }
        if (!job) {
            c.nanosleep(512);
            continue;
        }
        c.sync();
        c.fence_block();  // the team's `before` is visible after its REQUESTED
        if (c.lane == 0) {
            hs.rng = job->before;
            hs.n_segments = 0;
        }
        c.sync();
        bool ok = true;
        for (int q = 0; q < 3 && ok; ++q) ok = mg_ep_add_segment_coop(c, hs, m, tab);
        if (ok) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(hs.seg);
            uint32_t* dst = reinterpret_cast<uint32_t*>(job->seg);
            for (int w = c.lane; w < (int)(3 * sizeof(MgEPSegment) / 4); w += c.nl) dst[w] = src[w];
            if (c.lane == 0) job->after = hs.rng;
        }
        c.sync();
        if (c.lane == 0) {
            c.fence_block();
            // a failed generation ("No valid path found") is left to the env itself: the slot just empties
            mg_vstore_u32(&job->state, ok ? MG_SPEC_READY : MG_SPEC_EMPTY);
        }
    }
}

// ---- single-step launches: the frame first, the path later ------------------------------------------------------
// The first frame of an episode needs only the start row (the first draw of the first segment); the three A* runs
// are what makes a reset 20x as long as a step. In a single-step launch an auto-reset therefore draws the start
// row, finishes everything the frame needs, and leaves the path to a helper warp, which completes the record in
// global memory before the kernel ends. The team moves on to its next env at once.
template <class Ctx>
MG_HD bool mg_ep_reset_offload(Ctx& c, const MgEPK& k, MgEPState& s, MgEPFrame& f, MgEPWarpMem& wm) {
    if (c.shfl(mg_vload_u32(&wm.spec.state), 0) != MG_SPEC_EMPTY || wm.post_env >= 0) return false;  // slot in use: do it here
    MgRngR r = mg_rr_load(s.rng);
    const int sy = mg_rr_integers(r, 0, MG_MP_CELLS);
    c.sync();
    if (c.lane == 0) {
        mg_rr_store(r, s.rng);
        s.t = 0;
        s.ep_len = 0;
        s.ep_reward = 0.0;
        s.n_segments = 0;
        s.pending = (uint8_t)(1 + sy);
        s.camera_x = k.camera_offset;
        s.bg_scroll = 0;
        s.rot = 6;
        s.cx = 0 * MG_MP_TILE + k.radius;
        s.cy = sy * MG_MP_TILE + k.radius;
        s.cur_x = 0;
        s.cur_y = (uint8_t)sy;
        s.dir = 0;
        s.off_path = 0;
        s.num_fails = 0;
        s.n_fall = 0;
        s.stamina = k.p.stamina_level;
        s.max_x = 0;
        s.tiles_visited = 0;
        s.cross_on = 0;
        s.cross_x = s.cross_y = 0;
        s.live = 1;
        mg_ep_frame(k, s, mg_ep_bar_height(s.stamina, k.p.stamina_level, true), f);  // (no past-path tiles at x = 0)
        wm.post_env = wm.cur_env;
    }
    c.sync();
    return true;
}

// called by the kernel once the team's record is back in global memory
template <class Ctx>
MG_HD void mg_ep_post_offload(Ctx& c, MgEPWarpMem& wm) {
    if (c.lane == 0 && wm.post_env >= 0) {
        wm.spec.env = wm.post_env;
        wm.post_env = -1;
        c.fence_device();  // the record (written by every lane before the team's barrier) is visible device-wide
        mg_vstore_u32(&wm.spec.state, MG_SPEC_REQUESTED);
    }
}

// Helper warp of a single-step launch: completes pending records in global memory until every team has left and no
// job is waiting.
template <class Ctx>
MG_HD void mg_ep_helper_offload_loop(Ctx& c, uint8_t* hs_bytes, uint8_t* wextra0, int wextra_stride, int nteams, int hidx,
                                     const uint32_t* teams_done, MgGenMem& m, const MgGenCta& tab, uint8_t* state,
                                     double* ground_truth) {
    MgEPState& hs = *reinterpret_cast<MgEPState*>(hs_bytes);
    int next = hidx % nteams;
    for (;;) {
        const bool all_left = c.shfl(mg_vload_u32(teams_done), 0) >= (uint32_t)nteams;  // (jobs are posted before a team leaves)
        MgEPSpec* job = nullptr;
        for (int i = 0; i < nteams && !job; ++i) {
            const int t = (next + i) % nteams;
            MgEPSpec& sp = reinterpret_cast<MgEPWarpMem*>(wextra0 + (size_t)t * wextra_stride)->spec;
            if (c.shfl(mg_vload_u32(&sp.state), 0) != MG_SPEC_REQUESTED) continue;
            uint32_t won = 0;
            if (c.lane == 0) won = c.cas_u32(&sp.state, MG_SPEC_REQUESTED, MG_SPEC_BUSY) == MG_SPEC_REQUESTED ? 1u : 0u;
            if (c.shfl(won, 0)) {
                job = &sp;
                next = (t + 1) % nteams;
            }
        }
        if (!job) {
            if (all_left) return;
            c.nanosleep(256);
            continue;
        }
        c.sync();
        c.fence_device();
        MgEPState* g = reinterpret_cast<MgEPState*>(state) + job->env;
        const volatile MgEPState* gv = g;
        const int sy = (int)gv->pending - 1;
        if (c.lane == 0) {
            hs.rng.state_lo = gv->rng.state_lo, hs.rng.state_hi = gv->rng.state_hi;
            hs.rng.inc_lo = gv->rng.inc_lo, hs.rng.inc_hi = gv->rng.inc_hi;
            hs.rng.buf = gv->rng.buf, hs.rng.has_buf = gv->rng.has_buf;
            hs.n_segments = 0;
        }
        c.sync();
        bool ok = true;
        for (int q = 0; q < 3 && ok; ++q) ok = mg_ep_add_segment_coop(c, hs, m, tab, q == 0 ? sy : -1);
        if (ok) {
            if (c.lane == 0) hs.seg[0].reward_visited |= 1ull << mg_mp_id(0, sy);  // the first node yields no reward (:529)
            c.sync();
            const uint32_t* src = reinterpret_cast<const uint32_t*>(hs.seg);
            uint32_t* dst = reinterpret_cast<uint32_t*>(g->seg);
            for (int w = c.lane; w < (int)(3 * sizeof(MgEPSegment) / 4); w += c.nl) dst[w] = src[w];
        }
        if (c.lane == 0) {
            g->rng = hs.rng;
            g->n_segments = hs.n_segments;
            if (ok) {
                hs.cur_x = 0;
                hs.cur_y = (uint8_t)sy;
                hs.dir = 0;
                mg_ep_update_direction(hs);
                g->dir = hs.dir;
                if (ground_truth) mg_ep_ground_truth(hs, ground_truth + MG_GT_COLS * (size_t)job->env);
            } else {
                // "No valid path found": the reference raises; here the env stays finished, in the same state the inline
                // reset leaves behind (node (0, 0), no direction): nothing of the helper's scratch record is read
                // uninitialised
                g->live = 0;
                hs.cur_x = 0;
                hs.cur_y = 0;
                hs.dir = 0;
                g->cur_x = 0;
                g->cur_y = 0;
                g->dir = 0;
                if (ground_truth) mg_ep_ground_truth(hs, ground_truth + MG_GT_COLS * (size_t)job->env);
            }
            g->pending = 0;
        }
        c.sync();
        if (c.lane == 0) {
            c.fence_device();
            mg_vstore_u32(&job->state, MG_SPEC_EMPTY);
        }
    }
}

template <class Ctx>
MG_HD void mg_ep_reset_coop(Ctx& c, const MgEPK& k, MgEPState& s, MgEPFrame& f, double* gt, MgEPWarpMem& wm,
                            const MgGenCta& tab) {
    MgGenMem& m = wm.gen;
    c.stamp(12);
    // (an agent wider than a tile would show past-path tiles in its very first frame: those need the path)
    if (wm.helpers > 0 && wm.offload && k.radius < MG_MP_TILE && mg_ep_reset_offload(c, k, s, f, wm)) return;
    if (c.lane == 0) {
        s.t = 0;
        s.ep_len = 0;
        s.ep_reward = 0.0;
        s.n_segments = 0;
    }
    c.sync();
    bool ok = true;
    const bool spec = wm.helpers > 0 && !wm.offload;
    if (!(spec && mg_ep_spec_take(c, s, wm.spec)))
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
    if (spec) mg_ep_spec_request(c, s, wm.spec);
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
