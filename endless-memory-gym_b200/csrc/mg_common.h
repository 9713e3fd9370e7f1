// mg_common.h — layouts shared by the host asset builder, the env logic and the CUDA kernels.
//
// Everything here is plain data: the asset blob (sparse sprite pixel lists + background frame
// templates) and the per-env state records that live in HBM as an array of fixed-size records
// (one warp owns one env, so one record = one or two fully-coalesced 128-byte lines).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MG_HD __host__ __device__ __forceinline__
// one out-of-line copy: keeps the (rarely taken, single-lane) reset / RNG paths small in the I-cache
#define MG_HD_NOINLINE __host__ __device__ __noinline__
#define MG_NOUNROLL _Pragma("unroll 1")
#define MG_UNROLL _Pragma("unroll")
#else
#define MG_HD inline
#define MG_HD_NOINLINE inline
#define MG_NOUNROLL
#define MG_UNROLL
#endif

MG_HD int mg_popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

// branch layout hints: cold blocks leave the hot instruction stream (the single-lane phases are I-cache-miss-bound)
#define MG_LIKELY(x) __builtin_expect(!!(x), 1)
#define MG_UNLIKELY(x) __builtin_expect(!!(x), 0)

#define MG_W 84                       // screen: int(336 * 0.25), mortar_mayhem.py:13,73
#define MG_COL_BYTES (MG_W * 3)       // one x-column of the [x][y][c] frame = 252 contiguous bytes
#define MG_FRAME_BYTES (MG_W * MG_W * 3)  // 21168 = 1323 * 16
#define MG_FRAME_VEC16 (MG_FRAME_BYTES / 16)

struct alignas(16) MgVec16 {
    uint32_t x, y, z, w;
};

// MortarTile colours (pygame_assets.py:323-329), packed 0x00BBGGRR = bytes R,G,B in memory order
#define MG_RGB_TILE_FILL (21u | (43u << 8) | (77u << 16))
#define MG_RGB_TILE_EDGE (29u | (60u << 8) | (107u << 16))
#define MG_RGB_TILE_FILL_ON (81u | (18u << 8) | (26u << 16))
#define MG_RGB_TILE_EDGE_ON (112u | (24u << 8) | (36u << 16))

#define MG_ASSET_MAGIC 0x4d474131u  // "MGA1"
#define MG_MAX_PALETTE 32

// palette slots (RGB) used by the sprite lists; filled by the host builder
enum {
    MG_PAL_BLACK = 0,
    MG_PAL_BODY = 1,      // (250,204,153) character_controller.py:58
    MG_PAL_HAND = 2,      // (250,250,250)
    MG_PAL_HAND_OUT = 3,  // (50,50,50)
    MG_PAL_WHITE = 4,     // command icons
    MG_PAL_RED = 5,       // fall-off cross (255,0,0)
    MG_PAL_EXIT_CLOSED = 6,
    MG_PAL_EXIT_OPEN = 7,
    MG_PAL_COIN = 8,      // (255,165,0)
    MG_PAL_BLUE = 9,      // (0,0,255) raw pixel 255
    MG_PAL_GREEN = 10,    // (0,255,0)
};

// sprite = list of opaque pixels relative to the sprite's top-left: dx | dy<<8 | palette<<16
struct MgSprite {
    uint32_t first;  // index of the first entry in the blob's entry array
    uint16_t count;
    uint8_t w, h;
};

enum {
    MG_SPR_AGENT0 = 0,   // 8 agent rotations (index = rotation // 45)
    MG_SPR_ICON0 = 8,    // 9 command icons in Command.COMMANDS order (pygame_assets.py:242-252)
    MG_SPR_CROSS = 17,   // mystery_path.py:175-182
    MG_SPR_EXIT_CLOSED = 18,
    MG_SPR_EXIT_OPEN = 19,
    MG_SPR_COIN = 20,
    MG_SPR_RING = 21,    // ground-truth render: green ring on the target tile (mortar_mayhem.py:135)
    MG_SPR_COUNT = 24
};

#define MG_MAX_DISC_R 40  // spotlight / coin disc tables (filled circle row spans)

struct MgAssetHeader {
    uint32_t magic;
    uint32_t total_bytes;
    uint32_t entries_off;  // byte offset of the uint32 entry array
    uint32_t n_entries;
    uint32_t tmpl_off[4];  // byte offsets (16-aligned) of up to 4 full-frame templates, 0 = absent
    int32_t agent_dim;     // sprite side; blit top-left = centre - (dim >> 1)
    int32_t agent_radius;  // int(25 * agent_scale)
    int32_t icon_x, icon_y;  // command blit position (mortar_mayhem.py:363)
    uint32_t disc_off;     // byte offset of the int8 filled-circle tables: radius r (1..MG_MAX_DISC_R) has 2r rows
                           // dy = -r..r-1 of (xmin, xmax) relative to the centre, starting at row r*(r-1)
    uint32_t trig_off;     // byte offset of double[360][2] = (cos, sin) of integer degrees (host libm)
    uint32_t stage_bytes;  // bytes [0, stage_bytes) = header + sprite lists + dense bitmaps + tables (no templates)
    uint32_t ring_off;     // byte offset of the 1-px circle outlines (pygame draw.circle width=1): uint16 start[MG_MAX_DISC_R + 2]
                           // (in pixel pairs), then int8 (dx, dy) pairs; 0 = absent
    MgSprite sprite[MG_SPR_COUNT];
    uint32_t dense_off[MG_SPR_COUNT];  // byte offset of a dense w*h palette-index bitmap (255 = transparent), 0 = none
    uint8_t palette[MG_MAX_PALETTE][4];
};

// ---- PCG64 stream state (numpy.random.PCG64 + the bit generator's 32-bit buffer) ----------------
struct MgRng {
    uint64_t state_lo, state_hi;
    uint64_t inc_lo, inc_hi;
    uint32_t buf;      // uinteger
    uint32_t has_buf;  // has_uint32
};  // 40 bytes

// ---- Mortar Mayhem family record (256 bytes) --------------------------------------------------
#define MG_MM_MAXCMD 152
struct MgMMState {
    MgRng rng;             // 40
    double ep_reward;      // left-to-right sum of step rewards
    int32_t ep_len;
    int32_t t;             // Endless: steps since reset
    int32_t total_completed;  // Endless
    int16_t cx, cy;        // agent.rect.center
    int16_t draw_cx, draw_cy;  // Endless: pose of the last sprite handle (stale across resets)
    uint8_t rot;           // rotation // 45
    uint8_t draw_rot;
    uint8_t has_draw;      // Endless: a sprite handle exists
    uint8_t tiles_on;
    int8_t gx, gy;         // Grid: agent.grid_position
    int8_t tx, ty;         // _target_pos
    int8_t px, py;         // normalized agent position at the last act step / reset
    uint16_t n;            // len(_commands)
    uint16_t cc;           // _current_command
    uint16_t cs;           // _command_steps
    uint16_t cv;           // _command_verify_step
    uint16_t D, E;         // _explosion_delay, _explosion_duration
    uint16_t show_pos, show_len;  // cursor into / length of the command visualisation queue
    uint16_t show_dur, show_delay;
    uint16_t show_base;    // index of the first command the queue shows
    uint8_t live;          // an episode is running (0 after make() and after an un-reset done)
    uint8_t clone_ahead;   // debug render: the cloned command queue runs one entry ahead of the real one (Endless, after
                           // the list grew: endless_mortar_mayhem.py:321 re-clones before the step pops)
    uint8_t commands[MG_MM_MAXCMD];
};

// ---- Mystery Path record (128 bytes) -----------------------------------------------------------
struct MgMPState {
    MgRng rng;             // 40
    double ep_reward;
    int32_t ep_len;
    int32_t t;
    int32_t num_fails;
    int16_t cx, cy;
    int16_t cross_x, cross_y;  // fall_off_rect top-left
    uint64_t path_mask;    // bit (i*7+j): node on the path
    uint64_t visited;      // reward_visited
    uint8_t rot;
    uint8_t sx, sy, ex, ey;
    uint8_t off_path;
    uint8_t cross_on;      // fall_off_surface alpha == 255
    uint8_t live;
    int8_t gx, gy;         // MysteryPath-Grid: agent.grid_position
    uint8_t pad0[6];
    uint64_t wall_mask;    // bit (i*7+j): wall node (ground-truth render only: MysteryPath.wall_nodes)
    uint8_t pad[16];
};

// ---- Endless Mystery Path record (2048 bytes) ---------------------------------------------------
// One path segment = a 7x7 MysteryPath + its transition tile (pygame_assets.py:559-604), 8 columns wide.
// Cells are indexed x_local * 7 + y (x_local 0..6); bit 49 of the visited masks is the transition tile.
#define MG_EP_MAXSEG 45
#define MG_EP_MAXFALL 24
struct MgEPSegment {
    uint64_t path;      // bits 0..48: cell on the path; bits 56..58 start_y, 60..62 end_y
    uint64_t reward_visited;
    uint64_t stamina_visited;
    uint64_t next[2];   // 2 bits per cell: where next_node lies (0 east, 1 north = y-1, 2 south = y+1, 3 west)
};
struct MgEPState {
    MgRng rng;             // 40
    double ep_reward;
    double camera_x;       // accumulated in double like the reference (camera_offset may be fractional)
    int32_t ep_len, t, num_fails, stamina, max_x, tiles_visited;
    int32_t cx, cy;        // agent.rect.center (x grows without bound)
    int32_t bg_scroll;     // -bg_scroll of the reference: 0 <= v < 12, a multiple of the x velocity
    int32_t n_segments;
    int32_t cur_x;         // current_node: global column, row
    uint8_t cur_y, rot, off_path, cross_on, live, dir;  // dir: target_direction 0 east, 1 north, 2 south
    uint8_t pending;  // 1 + start row while a helper warp still owes this record its path (only inside a step launch)
    uint8_t pad0[1];
    int16_t cross_x, cross_y;  // fall_off_rect top-left
    uint16_t n_fall, pad1;
    int16_t fall[MG_EP_MAXFALL][2];  // fall_off_locations (normalised x, y)
    uint8_t pad2[36];
    MgEPSegment seg[MG_EP_MAXSEG];
};

// Behind every Endless-Mystery-Path record (MG_EP_STATE_BYTES = sizeof(MgEPState) + sizeof(MgEPSpecRec)): the path
// segments of the NEXT episodes, generated ahead of their resets by the helper warps of the rollout kernels
// (mg_ep_logic.h) — a ring of MG_EP_AHEAD episodes, each generated from the stream its predecessor ends with. The env's
// team never copies this part with the record; both sides go through global memory.
#define MG_EP_AHEAD 8
struct MgEPSpecSlot {
    MgRng before;        // the stream the reset must find for this episode to be adopted
    MgRng after;         // the stream after the segments generated so far (all three: what the reset leaves behind)
    MgEPSegment seg[3];
};
struct MgEPResume {      // a segment's noisy A* stopped at the end of a launch (mg_mp_generate_lane), resumed by the next
    uint32_t tag;        // the hword the search belongs to (0: none): same chain, episode and segment, or it is dropped
    uint8_t nopen, sy, ey, pad0;
    uint64_t wall, open, closed;
    uint64_t rng_lo, rng_hi;  // the stream as it was when the search stopped (same increment as the chain)
    uint32_t rng_buf, rng_has;
    uint64_t blob[19];   // the lane's scratch, packed: uint16 info[49], then uint8 list[49] (moved as 19 independent words)
    uint8_t pad1[48];
};
struct MgEPSpecRec {
    uint32_t hword;     // chain << 13 | MG_CHAIN_FRESH | episodes produced (mod 256) << 4 | segments of the next one << 2 | MG_CHAIN_*;
                        // advanced by the helper lane with CAS only, rewritten by the team when it starts a new chain
    uint32_t tword;     // chain << 13 | episodes consumed (mod 256); written by the team only
    uint32_t adopted;   // statistics: resets of this env that adopted their path from here (team)
    uint32_t work_ns;   // statistics: duration of the helper lane's last turn for this env (device builds)
    MgRng base;         // the stream the chain's first episode starts from (team)
    uint32_t missed_empty;  // statistics: resets that found no finished episode in the chain (team)
    uint32_t missed_stale;  // statistics: resets whose stream was not the one the episode was generated from (team)
    MgEPSpecSlot slot[MG_EP_AHEAD];
    MgEPResume resume;  // the helper lane's search in progress (helper)
    uint8_t pad1[128];
};
static_assert(sizeof(MgEPSpecSlot) == 200 && sizeof(MgEPResume) == 256 && sizeof(MgEPSpecRec) == 2048,
              "MgEPSpecRec must be 2048 bytes");

#define MG_EP_MAXTILES 72
struct MgEPFrame {
    int32_t sx, sy, srot;       // sprite top-left on screen, sprite index
    int32_t cross, cross_x, cross_y;
    int32_t bar_h;              // stamina bar: red rows from the top; -1 = no bar
    int32_t bg_off;             // show_background: screen x where a column of icy tiles starts; -1000 = black
    int32_t n_tiles;            // past-path tiles that touch the screen
    int32_t pad;
    int8_t tile_x[MG_EP_MAXTILES];   // screen x of the tile's left edge (-11..83)
    uint8_t tile_y[MG_EP_MAXTILES];  // tile row 0..6
};

// ---- Searing Spotlights record (1152 bytes) -----------------------------------------------------
#define MG_SS_MAXSPOTS 32
#define MG_SS_MAXCOINS 16
struct MgSSSpot {       // Spotlight (pygame_assets.py:61-131); positions are recomputed from the angles
    double t, speed;
    uint16_t a_start, a_target, a_offset;  // degrees, normalised to [0, 360)
    uint8_t radius;
    uint8_t done;
};
struct MgSSState {
    MgRng rng;            // 40
    double ep_reward;
    double hp;            // current_agent_health
    double last_reward;
    int32_t ep_len, t;
    int32_t spawn_timer, spawns_left;
    int16_t cx, cy;
    int16_t exit_x, exit_y;
    int16_t red_w_max;    // widest health-loss rect drawn on the top bar so far
    uint8_t rot;          // agent.rotation // 45
    uint8_t alpha;        // spotlight_surface.get_alpha()
    uint8_t n_spots;
    uint8_t n_coins;      // coins still on the field
    uint8_t num_coins;    // coins spawned this episode
    uint8_t coins_collected;
    uint8_t exit_open;
    uint8_t last_a0, last_a1;  // last_action
    uint8_t bar_c0, bar_c1;    // colours of the action boxes as currently drawn
    uint8_t coin_bar;          // 0 not drawn yet, 1 yellow, 2 grey
    uint8_t bg_blue, bg_red;   // sticky state of the two background surfaces: 0 chessboard, 1 white, 2 black
    uint8_t live;
    uint8_t pad0[5];
    int16_t coins[MG_SS_MAXCOINS][2];
    MgSSSpot spots[MG_SS_MAXSPOTS];
    int32_t coin_t;            // Endless: steps since the coin (re)spawned
    uint32_t scene_ver;        // bumped whenever anything but the agent / top bar / spotlights changes on screen
    uint32_t scene_last;       // (bg kind, coin layer order, coins left, exit open) of the last frame
    uint32_t done_mask;        // bit i = spots[i].done (the removal pass of the next step reads one word, not n bytes)
    // What the env instance keeps ACROSS reset() in the reference: spotlight_surface is created once per instance and
    // only repainted by step() (searing_spotlights.py:124-126, 208-216) — a reset frame still shows the discs of the last
    // step (visible when light_dim_off_duration == 0) — and the sprite handle rotated_agent_surface / rect (:135, :462).
    uint8_t last_n;            // discs on spotlight_surface
    uint8_t hand_rot;          // sprite handle: rotation // 45 ...
    uint8_t hand_has;          // ... exists (a step has run on this instance)
    uint8_t pad1[1];
    int16_t hand_x, hand_y;    // ... and its centre
    int16_t last_x[MG_SS_MAXSPOTS], last_y[MG_SS_MAXSPOTS];
    uint8_t last_r[MG_SS_MAXSPOTS];
    uint8_t pad2[24];
};

struct MgSSFrame {
    int16_t sx, sy;
    int16_t exit_x, exit_y;
    int16_t red_w_max, red_now;
    uint8_t srot, alpha, bg_kind, n_discs;   // bg_kind: 0 blue chessboard, 1 red chessboard, 2 white, 3 black
    uint8_t n_coins, exit_open, bar_c0, bar_c1;
    uint8_t coin_bar;
    uint8_t coin_above;   // Endless: the coin layer sits above the spotlight layer this frame (layers_alt order)
    uint8_t coin_black;   // Endless without a coin: the (never cleared) coin surface is an opaque black layer
    uint8_t pad[1];
    int16_t coins[MG_SS_MAXCOINS][2];
    int16_t disc_x[MG_SS_MAXSPOTS], disc_y[MG_SS_MAXSPOTS];
    uint8_t disc_r[MG_SS_MAXSPOTS];
    uint32_t scene_ver;   // MgSSState::scene_ver when the frame was described: equal = only agent / bar / spotlights moved
};

// frame description produced by the env logic and consumed by the rasteriser (MM family)
struct MgMMFrame {
    int red;        // tiles reddened (tiles_on && visual_feedback)
    int tx, ty;     // the one tile that stays blue
    int sx, sy;     // sprite centre
    int srot;       // sprite index
    int icon;       // -2: no command layer, -1: blank, 0..8 icon
};

struct MgMPFrame {
    int sx, sy, srot;
    int cross;      // draw the cross
    int cross_x, cross_y;
};

static_assert(sizeof(MgRng) == 40, "MgRng layout");
static_assert(sizeof(MgMMState) == 256, "MgMMState must be 256 bytes (MG_MM_STATE_BYTES)");
static_assert(sizeof(MgMPState) == 128, "MgMPState must be 128 bytes (MG_MP_STATE_BYTES)");
static_assert(sizeof(MgEPSegment) == 40, "MgEPSegment layout");
static_assert(sizeof(MgEPState) + sizeof(MgEPSpecRec) == MG_EP_STATE_BYTES, "MgEPState + MgEPSpecRec must be MG_EP_STATE_BYTES");
static_assert(sizeof(MgSSSpot) == 24, "MgSSSpot layout");
static_assert(sizeof(MgSSState) == 1152, "MgSSState must be 1152 bytes (MG_SS_STATE_BYTES)");
static_assert(sizeof(MgAssetHeader) % 8 == 0, "MgAssetHeader alignment");
