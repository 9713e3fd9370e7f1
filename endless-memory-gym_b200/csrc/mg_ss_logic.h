// mg_ss_logic.h — Searing Spotlights: reset + step state machine for ONE env record.
//
// Reference paths replaced:
//   SearingSpotlightsEnv.reset/step/_step_*   memory_gym/searing_spotlights.py:187-242, 267-330, 332-450, 452-562
//   Spotlight                                  memory_gym/pygame_assets.py:61-131   (fp64 lerp, int() centre, <= hit test)
//   GridPositionSampler                        memory_gym/pygame_assets.py:7-59     (uniform over unmasked cells)
//   Coin / Exit hit tests                      memory_gym/pygame_assets.py:154-167, 207-220
//   CharacterController.step                   memory_gym/character_controller.py:89-146
// All double arithmetic is mul/add/sqrt in the reference's order (compiled with -fmad=false /
// -ffp-contract=off); sin/cos come from a host-libm table in the asset blob (pygame Vector2.rotate).
// Reference quirks kept: every spawn interval is intervals[0] (list popped from the wrong end), and removing
// a finished spotlight / collected coin while iterating skips the next element for that step (A.9#5, #6).
#pragma once
#include <math.h>

#include "../../include/memgym_b200.h"
#include "mg_common.h"
#include "mg_mm_logic.h"  // kinematics
#include "mg_pcg64.h"

enum { MG_SS_L_BG = 0, MG_SS_L_SPOT = 1, MG_SS_L_BAR = 2, MG_SS_L_COIN = 3, MG_SS_L_EXIT = 4, MG_SS_L_AGENT = 5 };

struct MgSSK {
    mg_ss_params p;
    int32_t radius;             // agent radius
    int32_t step_axis, step_diag;
    int32_t coin_radius;        // int(10 * coin_scale)
    int32_t spot_rlo, spot_rhi; // integers(spot_min_radius, spot_max_radius + 1) with int()-truncated bounds
    int32_t interval0;          // int(initial_spawn_interval + spawn_interval_threshold)
    int32_t alpha_inc;          // int(255 / light_dim_off_duration)
    int32_t cb_x, cb_w;         // last-positive-reward box on the top bar
    int32_t n_layers;
    int32_t layers[6];          // painter's order (searing_spotlights.py:526-542), MG_SS_L_*
    int32_t layers_alt[6];      // Endless: the order while the coin is fresh (coin layer above the spotlights)
    int32_t spot_li[2], bar_li[2];  // index of the spotlight layer / the top bar in layers ([0]) and layers_alt ([1])
    int32_t endless;            // Endless-SearingSpotlights-v0
    int32_t clamp_off;          // int(30 * SCALE)
    int32_t exit_d2max;         // largest integer d2 with sqrt(d2) <= exit_radius + agent radius (host-computed)
    double exit_radius;         // 20 / 2 * exit_scale
    double spawn_base;          // sqrt(84^2 + 84^2) / 2
};

// ---- GridPositionSampler: 84 rows x 84-bit occupancy masks ------------------------------------------
struct MgSSSampler {
    uint64_t lo[MG_W], hi[MG_W];  // column x of row y is blocked: bit x of (lo | hi << 64)
};

MG_HD int mg_popc64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}

// ---- spotlights -------------------------------------------------------------------------------------
// Spotlight.__init__: radius, speed, then the three angle draws
// (the option values arrive by value: an out-of-line function would read them through a pointer into parameter space)
MG_HD_NOINLINE void mg_ss_new_spot_v(int rlo, int rhi, double smin, double smax, MgRng& r, MgSSSpot& sp) {
    sp.radius = (uint8_t)mg_integers(r, rlo, rhi);
    sp.speed = mg_uniform(r, smin, smax);
    sp.t = 0.0;
    sp.done = 0;
    int start = (int)mg_integers(r, 0, 360);
    int target = start + 180 + (int)mg_integers(r, -45, 45);
    int offset = target + (int)mg_integers(r, -135, 135);
    sp.a_start = (uint16_t)mg_pymod(start, 360);
    sp.a_target = (uint16_t)mg_pymod(target, 360);
    sp.a_offset = (uint16_t)mg_pymod(offset, 360);
}

MG_HD void mg_ss_new_spot(const MgSSK& k, MgRng& r, MgSSSpot& sp) {
    mg_ss_new_spot_v(k.spot_rlo, k.spot_rhi, k.p.spot_min_speed, k.p.spot_max_speed, r, sp);
}

// centre + Vector2(spawn_radius, 0).rotate(angle)
MG_HD void mg_ss_on_circle(const double* trig, int angle, double spawn_radius, double& x, double& y) {
    x = 42.0 + trig[2 * angle] * spawn_radius;
    y = 42.0 + trig[2 * angle + 1] * spawn_radius;
}

// sqrt(v) <= R for an integer R >= 0, bit-exact without the square root in all but a sliver of cases: sqrt is
// correctly rounded and monotonic, so v <= R^2 implies sqrt(v) <= sqrt(R^2) = R, and v > R^2 (1 + 2^-50) implies
// sqrt(v) > R (1 + 2^-51 - ...) which rounds above R; only in between does the rounded root decide.
MG_HD bool mg_within(double v, int R) {
    const double r2 = (double)(R * R);
    if (v <= r2) return true;
    if (v > r2 * (1.0 + 8.8817841970012523e-16)) return false;
    return sqrt(v) <= (double)R;
}

// Spotlight.draw without the drawing + is_agent_inside: advances t, returns the circle centre and the hit flag
MG_HD bool mg_ss_advance_spot(const MgSSK& k, const double* trig, MgSSSpot& sp, int ax, int ay, int& cx, int& cy) {
    double R = k.spawn_base + (double)sp.radius;
    double sx, sy, tx, ty, ox, oy;
    mg_ss_on_circle(trig, sp.a_start, R, sx, sy);
    mg_ss_on_circle(trig, sp.a_target, R, tx, ty);
    mg_ss_on_circle(trig, sp.a_offset, R, ox, oy);
    const double t = sp.t, u = 1.0 - t;
    double lx = tx * u + ox * t, ly = ty * u + oy * t;  // target.lerp(offset, t)
    double px = sx * u + lx * t, py = sy * u + ly * t;  // spawn.lerp(lerp_target, t)
    cx = (int)px;
    cy = (int)py;
    sp.t = sp.t + sp.speed;
    if (sp.t >= 1.0) {
        sp.t = 1.0;
        sp.done = 1;
    }
    double dx = (double)ax - px, dy = (double)ay - py;
    return mg_within(dx * dx + dy * dy, (int)sp.radius + k.radius);
}

MG_HD void mg_ss_clamp_spawn(int o, int& x, int& y) {
    if (x < o)
        x = o;
    else if (x > MG_W - o)
        x = MG_W - o;
    if (y < o)
        y = o;
    else if (y > MG_W - o)
        y = MG_W - o;
}

MG_HD void mg_ss_frame_common(const MgSSState& s, int srot, MgSSFrame& f) {
    // every load first, then the stores: record and frame share field types, so the compiler may not reorder
    // them on its own and a single lane would pay one shared-memory round trip per field
    const int16_t cx = s.cx, cy = s.cy, exit_x = s.exit_x, exit_y = s.exit_y, red_w_max = s.red_w_max;
    const uint8_t exit_open = s.exit_open, alpha = s.alpha, bar_c0 = s.bar_c0, bar_c1 = s.bar_c1, coin_bar = s.coin_bar;
    const int nc = s.n_coins;
    f.coin_above = 0;
    f.coin_black = 0;
    f.sx = cx;
    f.sy = cy;
    f.srot = (uint8_t)srot;
    f.exit_x = exit_x;
    f.exit_y = exit_y;
    f.exit_open = exit_open;
    f.alpha = alpha;
    f.red_w_max = red_w_max;
    f.bar_c0 = bar_c0;
    f.bar_c1 = bar_c1;
    f.coin_bar = coin_bar;
    f.n_coins = (uint8_t)nc;
    for (int c = 0; c < nc; ++c) {
        const int16_t x = s.coins[c][0], y = s.coins[c][1];
        f.coins[c][0] = x;
        f.coins[c][1] = y;
    }
}

// Stamps the frame with the scene version: the rasteriser patches its base buffer instead of rebuilding it while
// the version stands still (coins and exit only ever change through the fields compared here, or across a reset).
MG_HD void mg_ss_scene_stamp(MgSSState& s, MgSSFrame& f, bool force) {
    const uint32_t now = (uint32_t)f.bg_kind | ((uint32_t)f.coin_above << 8) | ((uint32_t)f.n_coins << 16) | ((uint32_t)f.exit_open << 24);
    uint32_t ver = s.scene_ver;
    if (force || now != s.scene_last) ver++;
    s.scene_last = now;
    s.scene_ver = ver;
    f.scene_ver = ver;
}

MG_HD uint8_t mg_ss_bg_kind(const MgSSState& s, int red) {
    uint8_t st = red ? s.bg_red : s.bg_blue;
    return st == 0 ? (uint8_t)(red ? 1 : 0) : (st == 1 ? (uint8_t)2 : (uint8_t)3);
}

// ---- cooperative sampler: the 84 rows are spread over the lanes, one lane does the RNG draw and the scan ----
struct MgSSResetTmp {
    int32_t x, y, ok, nc;
    int32_t part[32];  // free cells of rows 3 l .. 3 l + 2
};

template <class Ctx>
MG_HD_NOINLINE void mg_ss_block_coop(Ctx c, MgSSSampler& g, int x, int y, int r) {
    int y0 = y - r + 1 < 0 ? 0 : y - r + 1, y1 = y + r - 1 > MG_W - 1 ? MG_W - 1 : y + r - 1;
    for (int yy = y0 + c.lane; yy <= y1; yy += c.nl) {
        int dy = yy - y, rem = r * r - dy * dy - 1;  // need dx^2 <= rem
        int h = (int)sqrt((double)rem);
        while (h * h > rem) --h;
        while ((h + 1) * (h + 1) <= rem) ++h;
        int xa = x - h < 0 ? 0 : x - h, xb = x + h > MG_W - 1 ? MG_W - 1 : x + h;
        if (xa > xb) continue;
        uint64_t mlo = 0, mhi = 0;
        if (xa < 64) {
            int b = xb < 63 ? xb : 63;
            mlo = (~0ull >> (63 - (b - xa))) << xa;
        }
        if (xb >= 64) {
            int a2 = xa > 64 ? xa - 64 : 0, b2 = xb - 64;
            mhi = (~0ull >> (63 - (b2 - a2))) << a2;
        }
        g.lo[yy] |= mlo;
        g.hi[yy] |= mhi;
    }
}

// the k-th (0-based) clear bit among bits 0..83 of (lo | hi << 64)
MG_HD int mg_ss_kth_free(uint64_t lo, uint64_t hi, int k) {
    uint64_t flo = ~lo, fhi = ~hi & ((1ull << (MG_W - 64)) - 1ull);
    int base = 0;
    uint64_t w = flo;
    int n = mg_popc64(flo);
    if (k >= n) {
        k -= n;
        w = fhi;
        base = 64;
    }
    for (int half = 32; half >= 1; half >>= 1) {  // binary search on popcounts
        uint64_t lowmask = (1ull << half) - 1ull;
        int cnt = mg_popc64(w & lowmask);
        if (k >= cnt) {
            k -= cnt;
            w >>= half;
            base += half;
        }
    }
    return base;
}

template <class Ctx>
MG_HD_NOINLINE void mg_ss_sample_coop(Ctx c, MgSSSampler& g, MgSSState& s, MgSSResetTmp& rt, int block_radius) {
    if (c.lane < 28) {  // 28 lanes x 3 rows
        int n = 0;
        for (int q = 0; q < 3; ++q) n += MG_W - mg_popc64(g.lo[3 * c.lane + q]) - mg_popc64(g.hi[3 * c.lane + q]);
        rt.part[c.lane] = n;
    }
    if (c.nl == 1)  // a single lane runs the whole "warp" (host harness)
        for (int l = 1; l < 28; ++l) {
            int n = 0;
            for (int q = 0; q < 3; ++q) n += MG_W - mg_popc64(g.lo[3 * l + q]) - mg_popc64(g.hi[3 * l + q]);
            rt.part[l] = n;
        }
    c.sync();
    if (c.lane == 0) {
        int total = 0;
        for (int l = 0; l < 28; ++l) total += rt.part[l];
        if (total <= 0 || !rt.ok) {
            rt.ok = 0;
            rt.x = rt.y = 0;
        } else {
            int kk = (int)mg_integers(s.rng, 0, total);  // cell_id = rng.integers(0, len(spawn_area))
            int l = 0;
            for (; l < 27 && kk >= rt.part[l]; ++l) kk -= rt.part[l];
            int row = 3 * l;
            for (;; ++row) {
                int fr = MG_W - mg_popc64(g.lo[row]) - mg_popc64(g.hi[row]);
                if (kk < fr) break;
                kk -= fr;
            }
            rt.x = mg_ss_kth_free(g.lo[row], g.hi[row], kk);
            rt.y = row;
        }
    }
    c.sync();
    if (rt.ok) mg_ss_block_coop(c, g, rt.x, rt.y, block_radius);
    c.sync();
}

// A reset frame shows spotlight_surface as the last step left it (the surface is only repainted by step()): invisible
// while its alpha is 0, i.e. whenever light_dim_off_duration > 0.
MG_HD void mg_ss_stale_discs(const MgSSState& s, MgSSFrame& f) {
    int n = 0;
    if (f.alpha != 0) {
        n = s.last_n;
        for (int j = 0; j < n; ++j) {
            f.disc_x[j] = s.last_x[j];
            f.disc_y[j] = s.last_y[j];
            f.disc_r[j] = s.last_r[j];
        }
    }
    f.n_discs = (uint8_t)n;
}

// reset(): RNG draw order — rotation; agent cell + 2 offsets; per initial spotlight radius, speed, 3 angles;
// num_coins; per coin cell + 2 offsets; exit cell + 2 offsets (searing_spotlights.py:354-414).
// Called by every lane of the warp; `g` / `rt` are per-warp scratch.
template <class Ctx>
MG_HD void mg_ss_reset_env(Ctx& c, const MgSSK& k, MgSSState& s, MgSSFrame& f, MgSSSampler& g, MgSSResetTmp& rt) {
    for (int y = c.lane; y < MG_W; y += c.nl) g.lo[y] = g.hi[y] = 0;
    if (c.lane == 0) {
        s.t = 0;
        s.ep_len = 0;
        s.ep_reward = 0.0;
        if (k.p.hide_chessboard) s.bg_blue = s.bg_red = 1;
        s.last_a0 = s.last_a1 = 0;
        s.rot = (uint8_t)mg_integers(s.rng, 0, 8);  // choice([0, 45, ..., 315])
        rt.ok = 1;
    }
    c.sync();
    if (k.p.sample_agent_position) {
        mg_ss_sample_coop(c, g, s, rt, 28);
    } else {
        if (c.lane == 0) rt.x = rt.y = MG_W / 2;
        c.sync();
        mg_ss_block_coop(c, g, MG_W / 2, MG_W / 2, 21);
        c.sync();
    }
    if (c.lane == 0) {
        int x = rt.x, y = rt.y;
        if (k.p.sample_agent_position) {
            x += (int)mg_integers(s.rng, 2, 4);
            y += (int)mg_integers(s.rng, 2, 4);
        }
        s.cx = (int16_t)x;  // not clamped (searing_spotlights.py:364-369)
        s.cy = (int16_t)y;
        s.hp = k.p.agent_health;
        s.red_w_max = 0;
        s.bar_c0 = s.bar_c1 = 0;
        s.coin_bar = 0;
        s.last_reward = 0.0;
        s.spawns_left = k.p.num_spawns;
        s.spawn_timer = 0;
        int a0 = k.p.light_threshold < 0 ? 0 : (k.p.light_threshold > 255 ? 255 : k.p.light_threshold);
        s.alpha = (uint8_t)(k.p.light_dim_off_duration > 0 ? 0 : a0);
        s.n_spots = 0;
        s.done_mask = 0;
        for (int i = 0; i < k.p.initial_spawns; ++i) mg_ss_new_spot(k, s.rng, s.spots[s.n_spots++]);
        s.coins_collected = 0;
        rt.nc = mg_choice(s.rng, k.p.num_coins);
        s.num_coins = (uint8_t)rt.nc;
        s.n_coins = 0;
    }
    c.sync();
    const int nc = rt.nc;
    for (int q = 0; q <= nc; ++q) {  // the coins, then the exit
        mg_ss_sample_coop(c, g, s, rt, 21);
        if (c.lane == 0 && rt.ok) {
            int x = rt.x + (int)mg_integers(s.rng, 2, 4);
            int y = rt.y + (int)mg_integers(s.rng, 2, 4);
            mg_ss_clamp_spawn(k.clamp_off, x, y);
            if (q < nc) {
                s.coins[s.n_coins][0] = (int16_t)x;
                s.coins[s.n_coins][1] = (int16_t)y;
                s.n_coins++;
            } else {
                s.exit_x = (int16_t)x;
                s.exit_y = (int16_t)y;
            }
        }
        c.sync();
    }
    if (c.lane == 0) {
        s.exit_open = 0;
        if (k.p.black_background) s.bg_blue = 2;
        s.live = rt.ok ? 1 : 0;  // sampler exhausted: the reference raises; here the env stays finished
        mg_ss_frame_common(s, 0, f);  // the reset frame draws get_rotated_sprite(0)
        mg_ss_stale_discs(s, f);
        f.red_now = 0;
        f.bg_kind = mg_ss_bg_kind(s, 0);
        mg_ss_scene_stamp(s, f, true);
    }
    c.sync();
}

// step(), split so that the spotlights (the only per-step work that grows with the scene) advance in
// parallel, one lane per spotlight: A (one lane) -> B (all lanes) -> C (one lane)
// C1 -> C2 of the endless variant (registers in the plain step; shared memory when the step is split in two)
struct MgESSCarry {
    double reward;
    int32_t red_now, bg_red, spot_done;
};
struct MgSSStepTmp {
    int32_t action;   // what the logic lane decided for this call: 0 idle, 1 reset, 2 step (4: + coin respawn)
    int32_t post;     // ... and after the step: 3 = auto-reset
    int32_t cx, cy;   // agent centre after the move
    int32_t n_proc;   // spotlights drawn this step
    int32_t hits;     // how many of them contain the agent (is_agent_inside)
    uint8_t proc[MG_SS_MAXSPOTS];  // their indices in s.spots
    MgESSCarry carry;
};

// A: agent move, top-bar bookkeeping, light dimming, spawn, and WHICH spotlights are drawn this step
MG_HD void mg_ss_step_a(const MgSSK& k, MgSSState& s, int a0, int a1, MgSSStepTmp& tmp) {
    int cx = s.cx, cy = s.cy, rot = s.rot;
    // walkable_rect = (0, 16 * SCALE, 84, 84 - 16 * SCALE)
    mg_move_clamped(cx, cy, rot, a0, a1, k.step_axis, k.step_diag, k.radius, 0, 4, MG_W, MG_W);
    s.cx = (int16_t)cx;
    s.cy = (int16_t)cy;
    s.rot = (uint8_t)rot;
    s.hand_x = (int16_t)cx;  // rotated_agent_surface / rect (:462)
    s.hand_y = (int16_t)cy;
    s.hand_rot = (uint8_t)rot;
    s.hand_has = 1;
    tmp.cx = cx;
    tmp.cy = cy;
    if (k.p.show_last_action) {  // the boxes show the PREVIOUS action, then last_action is updated
        s.bar_c0 = s.last_a0;
        s.bar_c1 = s.last_a1;
        s.last_a0 = (uint8_t)a0;
        s.last_a1 = (uint8_t)a1;
    }
    if ((int)s.alpha <= k.p.light_threshold) {
        int a = k.p.light_dim_off_duration > 0 ? (int)s.alpha + k.alpha_inc : k.p.light_threshold;
        s.alpha = (uint8_t)(a < 0 ? 0 : (a > 255 ? 255 : a));
    }
    s.spawn_timer++;
    if (k.endless) {  // a new spotlight every spawn_interval steps, forever (endless_searing_spotlights.py:189-194)
        if (s.spawn_timer >= k.p.spawn_interval) {
            // (a full record would drop the spawn WITH its five draws; mg_ss_check_params rules that out)
            if (s.n_spots < MG_SS_MAXSPOTS) mg_ss_new_spot(k, s.rng, s.spots[s.n_spots++]);
            s.spawn_timer = 0;
        }
    } else if (s.spawns_left > 0 && s.spawn_timer >= k.interval0) {
        if (s.n_spots < MG_SS_MAXSPOTS) mg_ss_new_spot(k, s.rng, s.spots[s.n_spots++]);
        s.spawns_left--;
        s.spawn_timer = 0;
    }
    // `for spot in self.spotlights: if spot.done: self.spotlights.remove(spot) else: draw + hit test`
    int np = 0, n = s.n_spots;
    uint32_t mask = s.done_mask;  // bit i = spots[i].done: the pass runs on registers
    for (int i = 0; i < n; ++i) {
        if ((mask >> i) & 1u) {
            for (int j = i; j + 1 < n; ++j) s.spots[j] = s.spots[j + 1];
            const uint32_t low = (1u << i) - 1u;
            mask = (mask & low) | ((mask >> 1) & ~low);
            n--;
            // the element that slid into slot i is skipped this step, exactly like list.remove inside `for`
        } else {
            tmp.proc[np++] = (uint8_t)i;
        }
    }
    s.n_spots = (uint8_t)n;
    s.done_mask = mask;
    tmp.n_proc = np;
}

// B: Spotlight.draw + is_agent_inside for every drawn spotlight (lane j takes the j-th one)
template <class Ctx>
MG_HD void mg_ss_step_b(Ctx& c, const MgSSK& k, const double* trig, MgSSState& s, MgSSStepTmp& tmp, MgSSFrame& f) {
    uint32_t done_bits = 0, hits = 0;
    for (int j = c.lane; j < tmp.n_proc; j += c.nl) {
        const int idx = tmp.proc[j];
        MgSSSpot& sp = s.spots[idx];
        int dx, dy;
        if (mg_ss_advance_spot(k, trig, sp, tmp.cx, tmp.cy, dx, dy)) hits |= 1u << j;
        if (sp.done) done_bits |= 1u << idx;
        f.disc_x[j] = (int16_t)dx;
        f.disc_y[j] = (int16_t)dy;
        f.disc_r[j] = sp.radius;
        s.last_x[j] = (int16_t)dx;  // spotlight_surface keeps them until the next step repaints it
        s.last_y[j] = (int16_t)dy;
        s.last_r[j] = sp.radius;
    }
    // at most MG_SS_MAXSPOTS = 32 spotlights: they all sit on the lanes of the team's first warp
    done_bits = c.reduce_or(done_bits);
    hits = c.reduce_or(hits);
    if (c.lane == 0) {
        s.done_mask |= done_bits;
        tmp.hits = mg_popc32(hits);
        s.last_n = (uint8_t)tmp.n_proc;
    }
}

// C: health, coins, exit, done, episode statistics, frame description
MG_HD bool mg_ss_step_c(const MgSSK& k, MgSSState& s, const MgSSStepTmp& tmp, double& reward_out, double* info,
                                 MgSSFrame& f) {
    const int cx = tmp.cx, cy = tmp.cy, rot = s.rot, nd = tmp.n_proc;
    double reward = 0.0;
    double r = 0.0;
    bool spot_done = false;
    const int hit = tmp.hits;
    int red_now = 0, bg_red = 0;
    if (hit > 0) {
        s.hp = s.hp - k.p.spot_damage;
        r += k.p.reward_inside_spotlight;
        int width = (int)((double)(MG_W / 2) * (1.0 - s.hp / k.p.agent_health));
        if (width > MG_W) width = MG_W;
        if (width > 0) {
            red_now = width;
            if (width > s.red_w_max) s.red_w_max = (int16_t)width;
        }
        bg_red = k.p.visual_feedback ? 1 : 0;
    } else {
        r += k.p.reward_outside_spotlight;
    }
    if (k.p.black_background) {
        if (bg_red)
            s.bg_red = 2;
        else
            s.bg_blue = 2;
    }
    if (s.hp <= 0.0) {
        spot_done = true;
        r += k.p.reward_death;
    }
    reward += r;
    // ---- coin task
    bool coins_done = true;
    if (s.num_coins > 0) {
        r = 0.0;
        for (int i = 0; i < (int)s.n_coins; ++i) {
            // Vector2.distance_to on integer coordinates: sqrt(d2) <= R  <=>  d2 <= R^2 exactly
            const int dx = cx - (int)s.coins[i][0], dy = cy - (int)s.coins[i][1], R = k.coin_radius + k.radius;
            if (dx * dx + dy * dy <= R * R) {
                for (int j = i; j + 1 < (int)s.n_coins; ++j) {
                    s.coins[j][0] = s.coins[j + 1][0];
                    s.coins[j][1] = s.coins[j + 1][1];
                }
                s.n_coins--;
                r += k.p.reward_coin;
                s.coins_collected++;
            }
        }
        coins_done = s.n_coins == 0;
        reward += r;
    }
    // ---- exit task
    int success = 0;
    bool exit_done = false;
    if (k.p.use_exit) {
        r = 0.0;
        if (coins_done) {
            s.exit_open = 1;
            const int dx = cx - (int)s.exit_x, dy = cy - (int)s.exit_y;
            if (dx * dx + dy * dy <= k.exit_d2max) {  // sqrt(d2) <= exit_radius + radius, threshold found on the host
                exit_done = true;
                r = k.p.reward_exit;
            }
        }
        reward += r;
    }
    bool done = false;
    if (spot_done) {
        done = true;
    } else if (coins_done) {
        if (k.p.use_exit) {
            if (exit_done) {
                done = true;
                success = 1;
            }
        } else if (s.num_coins > 0) {
            done = true;
            success = 1;
        }
    }
    s.t++;
    if (s.t == k.p.max_steps) done = true;
    if (k.p.show_last_positive_reward) {
        s.coin_bar = s.last_reward > 0.0 ? 1 : 2;
        s.last_reward = reward;
    }
    s.ep_reward = s.ep_reward + reward;
    s.ep_len++;
    reward_out = reward;
    mg_ss_frame_common(s, rot, f);
    f.n_discs = (uint8_t)nd;
    f.red_now = (int16_t)red_now;
    f.bg_kind = mg_ss_bg_kind(s, bg_red);
    mg_ss_scene_stamp(s, f, false);
    if (done) {
        s.live = 0;
        if (info) {
            info[MG_INFO_REWARD] = s.ep_reward;
            info[MG_INFO_LENGTH] = (double)s.ep_len;
            info[MG_INFO_SUCCESS] = (double)success;
            info[MG_INFO_A] = s.hp / k.p.agent_health;
            info[MG_INFO_B] = (double)s.coins_collected / (double)s.num_coins;
            info[MG_INFO_GT0] = 0.0;
            info[MG_INFO_GT1] = 0.0;
            info[7] = info[8] = info[9] = 0.0;
        }
    }
    return done;
}

// ---- Endless Searing Spotlights (endless_searing_spotlights.py) ---------------------------------------
MG_HD void mg_ess_frame(const MgSSK& k, const MgSSState& s, int srot, MgSSFrame& f) {
    mg_ss_frame_common(s, srot, f);
    // coin layer order (:380-385, 467-472): above the spotlights while visible or fresh
    f.coin_above = (k.p.coins_visible || s.coin_t < k.p.coin_show_duration) ? 1 : 0;
    f.coin_black = k.p.coin_enabled ? 0 : 1;
    f.exit_x = f.exit_y = -100;  // no exit layer in this variant
}

MG_HD void mg_ess_ground_truth(const MgSSK& k, const MgSSState& s, double* gt) {
    // np.concatenate((agent.rect.center, coin.location)) / screen_dim (:403-407, 484-501)
    gt[0] = (double)s.cx / 84.0;
    gt[1] = (double)s.cy / 84.0;
    gt[2] = k.p.coin_enabled ? (double)s.coins[0][0] / 84.0 : 0.0;
    gt[3] = k.p.coin_enabled ? (double)s.coins[0][1] / 84.0 : 0.0;
}

// _spawn_coin (:256-269): fresh sampler, only the OLD coin's surroundings blocked, sample(28) + two offsets
template <class Ctx>
MG_HD_NOINLINE void mg_ess_spawn_coin(Ctx c, int clamp_off, MgSSState& s, MgSSSampler& g, MgSSResetTmp& rt, bool have_old) {
    for (int y = c.lane; y < MG_W; y += c.nl) g.lo[y] = g.hi[y] = 0;
    if (c.lane == 0) rt.ok = 1;
    c.sync();
    if (have_old) {
        mg_ss_block_coop(c, g, s.coins[0][0], s.coins[0][1], 28);
        c.sync();
    }
    mg_ss_sample_coop(c, g, s, rt, 28);
    if (c.lane == 0) {
        int x = rt.x + (int)mg_integers(s.rng, 2, 4);
        int y = rt.y + (int)mg_integers(s.rng, 2, 4);
        mg_ss_clamp_spawn(clamp_off, x, y);
        s.coins[0][0] = (int16_t)x;
        s.coins[0][1] = (int16_t)y;
        s.n_coins = 1;
    }
    c.sync();
}

// reset(): RNG draw order — rotation; agent cell + 2 offsets; per initial spotlight radius, speed, 3 angles;
// coin cell + 2 offsets (:294-407)
template <class Ctx>
MG_HD void mg_ess_reset_env(Ctx& c, const MgSSK& k, MgSSState& s, MgSSFrame& f, MgSSSampler& g, MgSSResetTmp& rt,
                            double* gt) {
    for (int y = c.lane; y < MG_W; y += c.nl) g.lo[y] = g.hi[y] = 0;
    if (c.lane == 0) {
        s.t = 0;
        s.coin_t = 0;
        s.ep_len = 0;
        s.ep_reward = 0.0;
        if (k.p.hide_chessboard) s.bg_blue = s.bg_red = 1;
        s.last_a0 = s.last_a1 = 0;
        s.rot = (uint8_t)mg_integers(s.rng, 0, 8);
        rt.ok = 1;
    }
    c.sync();
    if (k.p.sample_agent_position) {
        mg_ss_sample_coop(c, g, s, rt, 28);
    } else {
        if (c.lane == 0) rt.x = rt.y = MG_W / 2;
        c.sync();
    }
    if (c.lane == 0) {
        int x = rt.x, y = rt.y;
        if (k.p.sample_agent_position) {
            x += (int)mg_integers(s.rng, 2, 4);
            y += (int)mg_integers(s.rng, 2, 4);
        }
        s.cx = (int16_t)x;
        s.cy = (int16_t)y;
        s.hp = k.p.agent_health;
        s.red_w_max = 0;
        s.bar_c0 = s.bar_c1 = 0;
        s.coin_bar = 0;
        s.last_reward = 0.0;
        s.spawns_left = 0;
        s.spawn_timer = 0;
        int a0 = k.p.light_threshold < 0 ? 0 : (k.p.light_threshold > 255 ? 255 : k.p.light_threshold);
        s.alpha = (uint8_t)(k.p.light_dim_off_duration > 0 ? 0 : a0);
        s.n_spots = 0;
        s.done_mask = 0;
        for (int i = 0; i < k.p.initial_spawns; ++i) mg_ss_new_spot(k, s.rng, s.spots[s.n_spots++]);
        s.coins_collected = 0;
        s.num_coins = 1;
        s.n_coins = 0;
        s.exit_open = 0;
        s.exit_x = s.exit_y = -100;
    }
    c.sync();
    if (k.p.coin_enabled) mg_ess_spawn_coin(c, k.clamp_off, s, g, rt, false);
    if (c.lane == 0) {
        if (k.p.black_background) s.bg_blue = 2;
        s.live = rt.ok ? 1 : 0;
        mg_ess_frame(k, s, 0, f);  // the reset frame draws get_rotated_sprite(0)
        mg_ss_stale_discs(s, f);
        f.red_now = 0;
        f.bg_kind = mg_ss_bg_kind(s, 0);
        mg_ss_scene_stamp(s, f, true);
        if (gt) mg_ess_ground_truth(k, s, gt);
    }
    c.sync();
}

// C1 (one lane): health / background after the spotlights moved, and the coin hit test. Returns true when the
// coin was collected: the warp then respawns it cooperatively (mg_ess_spawn_coin) before C2 runs.
MG_HD bool mg_ess_step_c1(const MgSSK& k, MgSSState& s, const MgSSStepTmp& tmp, MgESSCarry& cy) {
    const int cx = tmp.cx, cyy = tmp.cy;
    double reward = 0.0, r = 0.0;
    const int hit = tmp.hits;
    cy.red_now = 0;
    cy.bg_red = 0;
    cy.spot_done = 0;
    if (hit > 0) {
        s.hp = s.hp - k.p.spot_damage;
        r += k.p.reward_inside_spotlight;
        int width = (int)((double)(MG_W / 2) * (1.0 - s.hp / k.p.agent_health));
        if (width > MG_W) width = MG_W;
        if (width > 0) {
            cy.red_now = width;
            if (width > s.red_w_max) s.red_w_max = (int16_t)width;
        }
        cy.bg_red = k.p.visual_feedback ? 1 : 0;
    } else {
        r += k.p.reward_outside_spotlight;
    }
    if (k.p.black_background) {
        if (cy.bg_red)
            s.bg_red = 2;
        else
            s.bg_blue = 2;
    }
    if (s.hp <= 0.0) {
        cy.spot_done = 1;
        r += k.p.reward_death;
    }
    reward += r;
    bool collected = false;
    if (k.p.coin_enabled) {
        r = 0.0;
        const int dx = cx - (int)s.coins[0][0], dy = cyy - (int)s.coins[0][1], R = k.coin_radius + k.radius;
        if (dx * dx + dy * dy <= R * R) {
            r += k.p.reward_coin;
            s.coins_collected++;
            s.coin_t = 0;
            collected = true;
        }
        reward += r;
    }
    cy.reward = reward;
    return collected;
}

// C2 (one lane): time limits, top bar, episode statistics, frame description, ground truth
MG_HD bool mg_ess_step_c2(const MgSSK& k, MgSSState& s, const MgSSStepTmp& tmp, const MgESSCarry& cy,
                                   double& reward_out, double* info, MgSSFrame& f, double* gt) {
    const double reward = cy.reward;
    bool done = cy.spot_done != 0;
    s.t++;
    s.coin_t++;
    if (s.coin_t == k.p.steps_per_coin && k.p.coin_enabled) done = true;
    if (s.t == k.p.max_steps) done = true;
    if (k.p.show_last_positive_reward) {
        s.coin_bar = s.last_reward > 0.0 ? 1 : 2;
        s.last_reward = reward;
    }
    s.ep_reward = s.ep_reward + reward;
    s.ep_len++;
    reward_out = reward;
    mg_ess_frame(k, s, s.rot, f);
    f.n_discs = (uint8_t)tmp.n_proc;
    f.red_now = (int16_t)cy.red_now;
    f.bg_kind = mg_ss_bg_kind(s, cy.bg_red);
    mg_ss_scene_stamp(s, f, false);
    double g4[4];
    mg_ess_ground_truth(k, s, g4);
    if (gt)
        for (int i = 0; i < 4; ++i) gt[i] = g4[i];
    if (done) {
        s.live = 0;
        if (info) {
            info[MG_INFO_REWARD] = s.ep_reward;
            info[MG_INFO_LENGTH] = (double)s.ep_len;
            info[MG_INFO_SUCCESS] = 0.0;
            info[MG_INFO_A] = s.hp / k.p.agent_health;
            info[MG_INFO_B] = (double)s.coins_collected;
            for (int i = 0; i < 4; ++i) info[MG_INFO_GT0 + i] = g4[i];
            info[MG_INFO_C] = 0.0;
        }
    }
    return done;
}
