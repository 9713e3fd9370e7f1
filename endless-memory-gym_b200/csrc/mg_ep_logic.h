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
// walks into a new segment. The three segments of the next episodes can therefore be generated ahead: episode 1 from
// the stream a reset leaves behind, episode 2 from the stream episode 1's segments leave behind, ... A reset adopts
// the oldest episode of this CHAIN if — and only if — the env's stream is what that episode was generated from
// (bit-exact either way); otherwise it generates inline and starts a new chain from the stream it ends with.
//
// Step and rollout launches run HELPER WARPS next to the env teams (mg_kernels.cu) that extend the chains with ONE ENV PER
// LANE (mg_mp_generate_lane): 32 paths per warp-instruction, where the cooperative generator of a reset keeps a
// whole warp busy for ~60 us per episode. The chain lives in the MgEPSpecRec behind the record, in GLOBAL memory,
// so it survives the launch; progress is published segment by segment, and a search cut off by the end of a launch is
// parked there too (MgEPResume). The counters are 8 bits wide and wrap; a chain lives for thousands of episodes.
//   team    consumes: slot[consumed % MG_EP_AHEAD] when produced > consumed and slot.before == stream, consumed + 1;
//           starts a new chain (base = stream, chain id + 1, counters 0) after an inline reset / a mid-episode segment
//   helper  (each env has ONE helper lane per launch: no helper-helper races) produces into slot[produced %
//           MG_EP_AHEAD] while produced - consumed < MG_EP_AHEAD; writes the data first, then advances hword by CAS:
//           a chain the team has replaced meanwhile makes the CAS fail and the data is never looked at
enum { MG_CHAIN_IDLE = 0, MG_CHAIN_ACTIVE = 1, MG_CHAIN_FAILED = 3 };
#define MG_CHAIN_FRESH (1u << 12)  // hword: no episode of this chain is complete yet (the first one starts from `base`)
MG_HD uint32_t mg_chain_id(uint32_t w) { return w >> 13; }
MG_HD uint32_t mg_chain_count(uint32_t w) { return (w >> 4) & 255u; }   // hword: produced
MG_HD uint32_t mg_chain_nseg(uint32_t w) { return (w >> 2) & 3u; }
MG_HD uint32_t mg_chain_consumed(uint32_t t) { return t & 255u; }        // tword
// single-step launches hand an auto-reset's path to a helper warp through shared memory (mg_ep_reset_offload)
enum { MG_SPEC_EMPTY = 0, MG_SPEC_REQUESTED = 1, MG_SPEC_BUSY = 2 };
struct MgEPSpec {
    uint32_t state;   // MG_SPEC_*: EMPTY -> (team) REQUESTED -> (helper) BUSY -> EMPTY
    int32_t env;      // the env whose (already written back) record waits for its path
};
struct MgEPWarpMem {  // per-team scratch of the Endless Mystery Path kernels
    MgGenMem gen;
    int32_t pad;
    int16_t ahead;    // this launch runs helper warps that generate paths ahead (mg_ep_lane_helper_loop): chains are live
    int16_t offload;  // single-step launch with helper warps that finish an auto-reset's path (mg_ep_reset_offload)
    int32_t cur_env;  // the env this team works on
    int32_t post_env; // >= 0: its record carries a pending path; post the job once the record is written back
    uint8_t* state;   // the batch's records (rollouts: where the MgEPSpecRec of an env lives)
    MgEPSpec spec;
};
// what a helper warp of a single-step launch works on: the head of an env record (stream, segment count, the
// first three segments)
#define MG_EP_HELPER_BYTES 384
static_assert(offsetof(MgEPState, seg) + 3 * sizeof(MgEPSegment) <= MG_EP_HELPER_BYTES, "helper scratch too small");
// per-lane scratch of a rollout helper warp: uint16 info[49][32] + uint8 list[49][32] (mg_mp_generate_lane)
#define MG_EP_LANE_WARP_BYTES (MG_MP_NODES * 32 * 3)

MG_HD MgEPSpecRec* mg_ep_spec_of(uint8_t* state, int env) {
    return reinterpret_cast<MgEPSpecRec*>(state + (size_t)env * MG_EP_STATE_BYTES + sizeof(MgEPState));
}

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

MG_HD void mg_ep_stream_put(volatile MgRng& d, const MgRng& s) {
    d.state_lo = s.state_lo, d.state_hi = s.state_hi, d.inc_lo = s.inc_lo, d.inc_hi = s.inc_hi;
    d.buf = s.buf, d.has_buf = s.has_buf;
}

// Team side, at a reset: true if the three segments were adopted from the env's chain. The slot is read as 50
// independent words spread over the lanes (a chain of dependent global loads would cost the team ~1 us each).
template <class Ctx>
MG_HD bool mg_ep_chain_take(Ctx& c, MgEPState& s, MgEPWarpMem& wm) {
    static_assert(sizeof(MgRng) == 40 && offsetof(MgEPSpecSlot, after) == 40 && offsetof(MgEPSpecSlot, seg) == 80, "slot layout");
    MgEPSpecRec* const g = mg_ep_spec_of(wm.state, (int)c.shfl((uint32_t)wm.cur_env, 0));
    const uint32_t h = c.shfl(mg_vload_u32(&g->hword), 0), tword = c.shfl(mg_vload_u32(&g->tword), 0);
    if ((h & 3u) == MG_CHAIN_IDLE || mg_chain_id(h) != mg_chain_id(tword) ||
        ((mg_chain_count(h) - mg_chain_consumed(tword)) & 255u) == 0u) {
        if (c.lane == 0) c.atomic_inc(&g->missed_empty);
        return false;
    }
    MgEPSpecSlot* const sl = &g->slot[mg_chain_consumed(tword) % MG_EP_AHEAD];
    c.fence_device();  // the helper's data is visible after its count
    const volatile uint32_t* src = reinterpret_cast<const volatile uint32_t*>(sl);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(&s.rng);
    uint32_t differs = 0;
    for (int w = c.lane; w < 10; w += c.nl) differs |= src[w] != mine[w] ? 1u : 0u;  // slot.before == stream?
    if (c.reduce_or(differs)) {
        if (c.lane == 0) c.atomic_inc(&g->missed_stale);
        return false;
    }
    c.sync();
    uint32_t* rng = reinterpret_cast<uint32_t*>(&s.rng);
    uint32_t* seg = reinterpret_cast<uint32_t*>(s.seg);
    for (int w = 10 + c.lane; w < 50; w += c.nl) {
        const uint32_t v = src[w];
        if (w < 20)
            rng[w - 10] = v;  // slot.after (same increment)
        else
            seg[w - 20] = v;
    }
    c.sync();  // every lane has read the slot before lane 0 hands it back to the helper
    if (c.lane == 0) {
        s.n_segments = 3;
        c.atomic_inc(&g->adopted);
        c.fence_device();
        mg_vstore_u32(&g->tword, (tword & ~255u) | ((mg_chain_consumed(tword) + 1u) & 255u));
    }
    c.sync();
    return true;
}

// Owner side (one thread): start a new chain from `stream` (what was generated ahead is stale or missing)
template <class Ctx>
MG_HD void mg_ep_chain_start(Ctx& c, MgEPSpecRec* g, const MgRng& stream) {
    const uint32_t chain = (mg_chain_id(mg_vload_u32(&g->tword)) + 1u) & 0x7FFFFu;
    mg_ep_stream_put(g->base, stream);
    mg_vstore_u32(&g->tword, chain << 13);
    c.fence_device();
    mg_vstore_u32(&g->hword, (chain << 13) | MG_CHAIN_FRESH | MG_CHAIN_ACTIVE);
}
// Team side: ... from the env's stream as it is now
template <class Ctx>
MG_HD void mg_ep_chain_restart(Ctx& c, const MgEPState& s, MgEPWarpMem& wm) {
    if (c.lane == 0) mg_ep_chain_start(c, mg_ep_spec_of(wm.state, wm.cur_env), s.rng);
}

// Helper side, ONE LANE: does the chain of record g have room for another segment? (hword as read by the caller)
MG_HD bool mg_ep_chain_wants(MgEPSpecRec* g, uint32_t h) {
    if ((h & 3u) != MG_CHAIN_ACTIVE) return false;
    const uint32_t t = mg_vload_u32(&g->tword);
    if (mg_chain_id(t) != mg_chain_id(h)) return false;  // the team is about to replace the chain
    return ((mg_chain_count(h) - mg_chain_consumed(t)) & 255u) < (uint32_t)MG_EP_AHEAD;
}

// Helper side, ONE LANE: generate up to max_new segments for the chain (hword = h) of record g. Progress is published
// after every segment; it stops when the chain is full, fails, is replaced by the team, or the launch ends (m.stop).
template <class Ctx>
MG_HD void mg_ep_lane_work(Ctx& c, MgEPSpecRec* g, uint32_t h, const double* sqrt_d2, const MgLaneMem& m,
                           int max_new = 1) {
    while (max_new-- > 0 && mg_ep_chain_wants(g, h)) {
        const uint32_t produced = mg_chain_count(h);
        const int nseg = (int)mg_chain_nseg(h);
        MgEPSpecSlot* const sl = &g->slot[produced % MG_EP_AHEAD];
        MgEPResume& rs = g->resume;
        MgRngR r;
        MgLaneSearch st;
        int sy, ey;
        // (nopen: a record that was never zero-initialised must not send the search outside its scratch)
        const bool resume = *reinterpret_cast<const volatile uint32_t*>(&rs.tag) == h && rs.nopen <= MG_MP_NODES;
        if (resume) {
            // this very segment's search was stopped by the end of an earlier launch: carry on where it stopped
            const volatile MgRng& inc = g->base;
            r.lo = rs.rng_lo, r.hi = rs.rng_hi, r.ilo = inc.inc_lo, r.ihi = inc.inc_hi, r.buf = rs.rng_buf, r.has = rs.rng_has;
            sy = rs.sy, ey = rs.ey;
            st.wall = rs.wall, st.open = rs.open, st.closed = rs.closed, st.nopen = rs.nopen;
            uint64_t blob[MG_LANE_BLOB_WORDS];
            MG_UNROLL
            for (int w = 0; w < MG_LANE_BLOB_WORDS; ++w) blob[w] = rs.blob[w];
            mg_lane_unpack(blob, m);
        } else {
            // the stream to continue from: the slot's own progress, else what the previous episode (or the chain's base) ends with
            const volatile MgRng& from =
                nseg > 0 ? sl->after : ((h & MG_CHAIN_FRESH) ? g->base : g->slot[(produced + MG_EP_AHEAD - 1u) % MG_EP_AHEAD].after);
            r.lo = from.state_lo, r.hi = from.state_hi, r.ilo = from.inc_lo, r.ihi = from.inc_hi, r.buf = from.buf, r.has = from.has_buf;
            if (nseg == 0) {
                volatile MgRng& b = sl->before;
                b.state_lo = r.lo, b.state_hi = r.hi, b.inc_lo = r.ilo, b.inc_hi = r.ihi, b.buf = r.buf, b.has_buf = r.has;
            }
            // EndlessMysteryPath.add_path_segment, as in mg_ep_add_segment_coop
            sy = nseg == 0 ? mg_rr_integers(r, 0, MG_MP_CELLS)
                           : (int)((*reinterpret_cast<const volatile uint64_t*>(&sl->seg[nseg - 1].path) >> 60) & 7ull);
            ey = mg_rr_integers(r, 0, MG_MP_CELLS);
        }
        uint64_t nx[2];
        bool stopped;
        const uint64_t mask = mg_mp_generate_lane(r, 0, sy, MG_MP_CELLS - 1, ey, sqrt_d2, m, nx, stopped, st, resume);
        if (stopped) {
            // the launch is ending: park the search behind the record (the next launch's helper lane resumes it)
            rs.nopen = (uint8_t)st.nopen, rs.sy = (uint8_t)sy, rs.ey = (uint8_t)ey;
            rs.wall = st.wall, rs.open = st.open, rs.closed = st.closed;
            rs.rng_lo = r.lo, rs.rng_hi = r.hi, rs.rng_buf = r.buf, rs.rng_has = r.has;
            uint64_t blob[MG_LANE_BLOB_WORDS];
            mg_lane_pack(m, blob);
            MG_UNROLL
            for (int w = 0; w < MG_LANE_BLOB_WORDS; ++w) rs.blob[w] = blob[w];
            c.fence_device();
            *reinterpret_cast<volatile uint32_t*>(&rs.tag) = h;
            return;
        }
        uint32_t next;
        if (!mask) {
            // "No valid path found": the reset will find out by itself (the reference raises there)
            next = (h & ~3u) | MG_CHAIN_FAILED;
        } else {
            volatile uint64_t* sg = reinterpret_cast<volatile uint64_t*>(&sl->seg[nseg]);
            sg[0] = mask | ((uint64_t)sy << 56) | ((uint64_t)ey << 60);
            sg[1] = 0;
            sg[2] = 0;
            sg[3] = nx[0];
            sg[4] = nx[1];
            volatile MgRng& a = sl->after;
            a.state_lo = r.lo, a.state_hi = r.hi, a.inc_lo = r.ilo, a.inc_hi = r.ihi, a.buf = r.buf, a.has_buf = r.has;
            c.fence_device();
            next = nseg < 2 ? (h & ~(3u << 2)) | ((uint32_t)(nseg + 1) << 2)
                            : (h & ~(0xFFCu | MG_CHAIN_FRESH)) | (((produced + 1u) & 255u) << 4);  // the episode is complete
        }
        if (c.cas_u32(&g->hword, h, next) != h) return;  // the team has started another chain meanwhile
        if (!mask) return;
        h = next;
    }
}

// Helper warp of a rollout launch. Lane `first` of the launch's `step` helper lanes owns envs first, first + step, ...
// (consecutive envs sit on the lanes of one warp: they are visited, and consume, at about the same time). ONE segment
// per turn: lanes whose chain got room meanwhile join at the next segment boundary, so that the walls of a segment are
// drawn by all lanes together and the searches run in step. (Turns of 16 node expansions with lanes joining at every
// turn were measured too: 204 instead of 229 M env-steps/s — every turn then pays the wall phase of the few lanes
// that begin, profiles/r02_emp_lookahead.md.)
template <class Ctx>
MG_HD void mg_ep_lane_helper_loop(Ctx& c, uint8_t* state, int n, int first, int step, const uint32_t* teams_done,
                                  int stop_at, const double* sqrt_d2, const MgLaneMem& m) {
    const int count = first < n ? (n - first + step - 1) / step : 0;
    int rot = 0;
    for (;;) {
        // the launch is over: the chains wait (m.not_before only keeps a search that is under way going a little longer)
        if (mg_vload_u32(teams_done) >= (uint32_t)stop_at) return;
        MgEPSpecRec* g = nullptr;
        uint32_t h = 0;
        for (int i = 0; i < count && !g; ++i) {
            int j = rot + i;
            if (j >= count) j -= count;
            MgEPSpecRec* cand = mg_ep_spec_of(state, first + j * step);
            h = mg_vload_u32(&cand->hword);
            if (mg_ep_chain_wants(cand, h)) {
                g = cand;
                rot = j + 1 < count ? j + 1 : 0;
            }
        }
        if (!c.reduce_or(g ? 1u : 0u)) {
            c.nanosleep(2000);
            continue;
        }
        if (g) {
            c.fence_device();  // the team's writes are visible after its words
            const uint64_t t0 = c.now_ns();
            mg_ep_lane_work(c, g, h, sqrt_d2, m, 1);
            g->work_ns = (uint32_t)(c.now_ns() - t0);
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
        MgEPState* g = reinterpret_cast<MgEPState*>(state + (size_t)job->env * MG_EP_STATE_BYTES);
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
                // the episodes after this one: a new chain from the stream this reset ends with
                mg_ep_chain_start(c, mg_ep_spec_of(state, job->env), hs.rng);
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
    if (c.lane == 0) {
        s.t = 0;
        s.ep_len = 0;
        s.ep_reward = 0.0;
        s.n_segments = 0;
    }
    c.sync();
    // 1. the episode generated ahead, if the stream still is what it was generated from
    const bool adopted = wm.ahead && mg_ep_chain_take(c, s, wm);
    // 2. single-step launches: the frame now, the path by a helper warp before the kernel ends
    // (an agent wider than a tile would show past-path tiles in its very first frame: those need the path)
    if (!adopted && wm.offload && k.radius < MG_MP_TILE && mg_ep_reset_offload(c, k, s, f, wm)) return;
    // 3. here and now
    bool ok = true;
    if (!adopted)
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
    if (wm.ahead && ok && !adopted) mg_ep_chain_restart(c, s, wm);  // (an adopted episode: its chain goes on)
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
