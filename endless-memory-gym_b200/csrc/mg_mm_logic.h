// mg_mm_logic.h — Mortar Mayhem family: reset + step state machine for ONE env record.
//
// Reference paths replaced (relative to the reference checkout):
//   MortarMayhemEnv.reset/step           memory_gym/mortar_mayhem.py:206-272, 274-369
//   GridMortarMayhemEnv.reset/step       memory_gym/mortar_mayhem_grid.py:213-278, 280-375
//   EndlessMortarMayhemEnv.reset/step    memory_gym/endless_mortar_mayhem.py:194-259, 261-373
//   CharacterController.step             memory_gym/character_controller.py:89-146
//   GridCharacterController.step         memory_gym/character_controller.py:177-210
//   ScreenWrapCharacterController.step   memory_gym/character_controller.py:226-283
// Runs on one lane of the warp that owns the env (the other 31 lanes join for the raster).
#pragma once
#include "../../include/memgym_b200.h"
#include "mg_common.h"
#include "mg_pcg64.h"

// kernel-side parameter block: the public options plus what the host derives from them once
struct MgMMK {
    mg_mm_params p;
    int32_t size;        // arena tiles per side (Endless: 6)
    int32_t origin;      // arena top-left on the 84x84 screen
    int32_t radius;      // agent radius used by the boundary clamp / wrap
    int32_t step_axis;   // int(1.0 * speed)
    int32_t step_diag;   // int((1 / sqrt(2)) * speed)
    double v_axis;       // Endless keeps the float velocity (character_controller.py:262-263)
    double v_diag;
    double wrap_off;     // radius * 0.5
    int32_t pad[2];
};

#define MG_TILE 14  // int(56 * SCALE)

MG_HD bool mg_mm_is_grid(int variant) { return variant == MG_MM_GRID || variant == MG_MM_B_GRID; }
MG_HD bool mg_mm_is_task_b(int variant) { return variant == MG_MM_B || variant == MG_MM_B_GRID; }

MG_HD int mg_floordiv(int a, int b) {
    int q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}
MG_HD int mg_pymod(int a, int b) {
    int r = a % b;
    return r < 0 ? r + b : r;
}
MG_HD void mg_cmd_delta(int k, int& dx, int& dy) {
    // Command.COMMANDS, pygame_assets.py:242-252: right, down, left, up, stay, right_down, right_up, left_down, left_up
    //   dx = { 1, 0,-1, 0, 0, 1, 1,-1,-1 }   dy = { 0, 1, 0,-1, 0, 1,-1, 1,-1 }   packed as (d + 1) in 2-bit fields
    const uint32_t DX = 2u | (1u << 2) | (0u << 4) | (1u << 6) | (1u << 8) | (2u << 10) | (2u << 12) | (0u << 14) | (0u << 16);
    const uint32_t DY = 1u | (2u << 2) | (1u << 4) | (0u << 6) | (1u << 8) | (2u << 10) | (0u << 12) | (2u << 14) | (0u << 16);
    dx = (int)((DX >> (2 * k)) & 3u) - 1;
    dy = (int)((DY >> (2 * k)) & 3u) - 1;
}
MG_HD int mg_choice(MgRng& r, const mg_choice_list& c) {
    return c.v[(int)mg_integers(r, 0, c.n)];
}

// character_controller.py:100-121: unit direction + sprite index (rotation // 45)
MG_HD void mg_dir_rot(int a0, int a1, int& vx, int& vy, int& rot) {
    vx = a0 == 1 ? -1 : (a0 == 2 ? 1 : 0);
    vy = a1 == 1 ? -1 : (a1 == 2 ? 1 : 0);
    if (a0 == 1) rot = 2;
    if (a0 == 2) rot = 6;
    if (a1 == 1) rot = 0;
    if (a1 == 2) rot = 4;
    if (vx && vy) rot = vx < 0 ? (vy < 0 ? 1 : 3) : (vy < 0 ? 7 : 5);
}

// CharacterController.step: int-truncated velocity, centre clamped to rect -+ radius
MG_HD void mg_move_clamped(int& cx, int& cy, int& rot, int a0, int a1, int step_axis, int step_diag, int radius,
                           int left, int top, int right, int bottom) {
    int vx, vy;
    mg_dir_rot(a0, a1, vx, vy, rot);
    int st = (vx && vy) ? step_diag : step_axis;
    cx += vx * st;
    cy += vy * st;
    if (cx > right - radius) cx = right - radius;
    if (cx < left + radius) cx = left + radius;
    if (cy > bottom - radius) cy = bottom - radius;
    if (cy < top + radius) cy = top + radius;
}

// ScreenWrapCharacterController.step: float velocity, Rect assignment truncates toward zero, torus wrap
MG_HD void mg_move_wrapped(int& cx, int& cy, int& rot, int a0, int a1, double v_axis, double v_diag, double off,
                           int left, int top, int right, int bottom) {
    int vx, vy;
    mg_dir_rot(a0, a1, vx, vy, rot);
    double v = (vx && vy) ? v_diag : v_axis;
    double x = (double)(int)((double)cx + (double)vx * v);
    double y = (double)(int)((double)cy + (double)vy * v);
    if (x > (double)right + off) x = (double)left - off;
    if (x < (double)left - off) x = (double)right + off;
    if (y > (double)bottom + off) y = (double)top - off;
    if (y < (double)top - off) y = (double)bottom + off;
    cx = (int)x;
    cy = (int)y;
}

MG_HD void mg_mm_frame(const MgMMK& k, const MgMMState& s, int sx, int sy, int srot, int icon, MgMMFrame& f) {
    f.red = (s.tiles_on && k.p.visual_feedback) ? 1 : 0;
    f.tx = s.tx;
    f.ty = s.ty;
    f.sx = sx;
    f.sy = sy;
    f.srot = srot;
    f.icon = icon;
}

MG_HD int mg_mm_show_entry(const MgMMState& s, int e) {
    int per = s.show_dur + s.show_delay;
    int ci = e / per, r = e % per;
    return r < s.show_dur ? (int)s.commands[s.show_base + ci] : -1;
}

// reset(): RNG draw order is part of the contract (mortar_mayhem.py:238,174,150-183,247-248,262-263)
MG_HD void mg_mm_reset_env(const MgMMK& k, MgMMState& s, MgMMFrame& f, double* gt) {
    const int size = k.size;
    const bool endless = k.p.variant == MG_MM_ENDLESS;
    MgRngR r = mg_rr_load(s.rng);  // the stream stays in registers for the whole reset
    s.ep_reward = 0.0;
    s.ep_len = 0;
    s.t = 0;
    s.total_completed = 0;
    s.live = 1;
    s.tiles_on = 0;
    int tile_id = mg_rr_integers(r, 0, size * size);
    int i = tile_id / size, j = tile_id % size;
    int cx = MG_TILE * i + k.origin + MG_TILE / 2, cy = MG_TILE * j + k.origin + MG_TILE / 2;
    if (endless || k.p.variant == MG_MM_B) {  // spawn offset, endless_mortar_mayhem.py:224, mortar_mayhem_b.py:167
        int o0 = mg_rr_integers(r, -2, 2);
        int o1 = mg_rr_integers(r, -2, 2);
        cx += o0;
        cy += o1;
    }
    int px = mg_floordiv(cx - k.origin, MG_TILE), py = mg_floordiv(cy - k.origin, MG_TILE);
    s.cx = (int16_t)cx;
    s.cy = (int16_t)cy;
    s.rot = 0;
    s.px = (int8_t)px;
    s.py = (int8_t)py;
    s.gx = (int8_t)px;
    s.gy = (int8_t)py;
    int n;
    if (endless) {
        n = k.p.initial_command_count;
        for (int c = 0; c < n; ++c) s.commands[c] = (uint8_t)mg_rr_integers(r, 0, k.p.allowed_commands);
    } else {
        n = mg_rr_choice(r, k.p.command_count);
        int simx = px, simy = py;
        for (int c = 0; c < n; ++c) {
            // available_commands: the first `allowed_commands` commands that keep the target inside the arena
            // (bit q set = command q; a command is invalid iff it steps off the arena on x or on y)
            uint32_t valid = (1u << k.p.allowed_commands) - 1u;
            if (simx >= size - 1) valid &= ~0x061u;  // dx = +1: right, right_down, right_up
            if (simx <= 0) valid &= ~0x184u;         // dx = -1: left, left_down, left_up
            if (simy >= size - 1) valid &= ~0x0A2u;  // dy = +1: down, right_down, left_down
            if (simy <= 0) valid &= ~0x148u;         // dy = -1: up, right_up, left_up
#if defined(__CUDA_ARCH__)
            const int nv = __popc(valid);
            const int q = (int)__fns(valid, 0u, mg_rr_integers(r, 0, nv) + 1);  // the pick-th set bit
#else
            int nv = 0;
            for (uint32_t m = valid; m; m &= m - 1u) nv++;
            int pick = mg_rr_integers(r, 0, nv), q = 0;
            for (;; ++q)  // the pick-th set bit
                if ((valid >> q) & 1u)
                    if (pick-- == 0) break;
#endif
            s.commands[c] = (uint8_t)q;
            int dx, dy;
            mg_cmd_delta(q, dx, dy);
            simx += dx;
            simy += dy;
        }
    }
    s.n = (uint16_t)n;
    const bool task_b = mg_mm_is_task_b(k.p.variant);
    s.show_base = 0;
    if (task_b) {  // no command visualisation: the commands are observed as a vector (mortar_mayhem_b.py:176)
        s.show_dur = s.show_delay = 0;
        s.show_len = 0;
        s.show_pos = 0;
    } else {
        s.show_dur = (uint16_t)mg_rr_choice(r, k.p.command_show_duration);
        s.show_delay = (uint16_t)mg_rr_choice(r, k.p.command_show_delay);
        s.show_len = (uint16_t)(n * (s.show_dur + s.show_delay));
        s.show_pos = 1;  // the first entry is consumed by the reset frame
    }
    int dx, dy;
    mg_cmd_delta(s.commands[0], dx, dy);
    int tx = px + dx, ty = py + dy;
    if (endless) {
        tx = mg_pymod(tx, 6);
        ty = mg_pymod(ty, 6);
    }
    s.tx = (int8_t)tx;
    s.ty = (int8_t)ty;
    s.cc = s.cs = s.cv = 0;
    s.clone_ahead = 0;  // _command_visualization_clone = copy of the full queue (mortar_mayhem.py:251)
    s.E = (uint16_t)mg_rr_choice(r, k.p.explosion_duration);
    s.D = (uint16_t)mg_rr_choice(r, k.p.explosion_delay);
    mg_rr_store(r, s.rng);
    mg_mm_frame(k, s, cx, cy, 0, task_b ? -2 : mg_mm_show_entry(s, 0), f);
    if (gt) {
        gt[0] = (double)tx / 5.0;
        gt[1] = (double)ty / 5.0;
    }
}

// step(): returns done. `info` (MG_INFO_COLS doubles) is written only when done.
MG_HD bool mg_mm_step_env(const MgMMK& k, MgMMState& s, int a0, int a1, double& reward_out, double* info,
                          MgMMFrame& f, double* gt) {
    const bool endless = k.p.variant == MG_MM_ENDLESS;
    double reward = 0.0;
    bool done = false;
    int success = 0;
    int icon = -2;
    int sx, sy, srot;
    if (s.show_pos < s.show_len) {
        icon = mg_mm_show_entry(s, s.show_pos);
        s.show_pos++;
        if (endless) {
            // the sprite handle is only refreshed by act steps and survives reset (quirk A.9#3)
            if (!s.has_draw) {
                s.draw_cx = s.cx;
                s.draw_cy = s.cy;
                s.draw_rot = 0;
                s.has_draw = 1;
            }
            sx = s.draw_cx;
            sy = s.draw_cy;
            srot = s.draw_rot;
        } else {
            sx = s.cx;
            sy = s.cy;
            srot = 0;
        }
    } else {
        int cx = s.cx, cy = s.cy, rot = s.rot;
        const int lo = k.origin, hi = k.origin + MG_TILE * k.size;
        if (mg_mm_is_grid(k.p.variant)) {
            // GridCharacterController.step: 0 noop, 1 rotate left (+90), 2 rotate right (-90), 3 forward
            int gx = s.gx, gy = s.gy;
            if (a0 == 1) rot = (rot + 2) & 7;
            if (a0 == 2) rot = (rot + 6) & 7;
            if (a0 == 3) {
                int face = rot >> 1, last = k.size - 1;
                if (face == 0 && gy > 0)
                    gy--;
                else if (face == 3 && gx < last)
                    gx++;
                else if (face == 2 && gy < last)
                    gy++;
                else if (face == 1 && gx > 0)
                    gx--;
                cx = MG_TILE * gx + k.origin + MG_TILE / 2;
                cy = MG_TILE * gy + k.origin + MG_TILE / 2;
            }
            s.gx = (int8_t)gx;
            s.gy = (int8_t)gy;
        } else if (endless) {
            mg_move_wrapped(cx, cy, rot, a0, a1, k.v_axis, k.v_diag, k.wrap_off, lo, lo, hi, hi);
            s.draw_cx = (int16_t)cx;
            s.draw_cy = (int16_t)cy;
            s.draw_rot = (uint8_t)rot;
            s.has_draw = 1;
        } else {
            mg_move_clamped(cx, cy, rot, a0, a1, k.step_axis, k.step_diag, k.radius, lo, lo, hi, hi);
        }
        s.cx = (int16_t)cx;
        s.cy = (int16_t)cy;
        s.rot = (uint8_t)rot;
        sx = cx;
        sy = cy;
        srot = rot;
        int px = mg_floordiv(cx - k.origin, MG_TILE), py = mg_floordiv(cy - k.origin, MG_TILE);
        s.px = (int8_t)px;
        s.py = (int8_t)py;
        const bool on_target = (px == s.tx && py == s.ty);
        const bool verify = (s.cs % s.D == 0) && s.cs > 0;
        if (verify && !s.tiles_on) {
            if (s.cc < s.n) {
                s.cc++;
                s.tiles_on = 1;
                if (on_target) {
                    reward += k.p.reward_command_success;
                    if (endless) {
                        s.total_completed++;
                        if (s.cc == s.n) reward += k.p.reward_episode_success;  // reward_new_command_success
                    }
                } else {
                    done = true;
                    reward += k.p.reward_command_failure;
                }
            }
            if (s.cc >= s.n) {
                if (endless) {
                    // the list grows by one command and only that command is shown (:311-322)
                    uint8_t nc = (uint8_t)mg_integers(s.rng, 0, k.p.allowed_commands);
                    if (s.n < MG_MM_MAXCMD) {
                        s.commands[s.n] = nc;
                        s.n++;
                    } else {
                        // record capacity reached: end the episode instead of corrupting the list
                        done = true;
                    }
                    s.cc = 0;
                    s.cv = 0;
                    s.show_base = (uint16_t)(s.n - 1);
                    s.show_len = (uint16_t)(s.show_dur + s.show_delay);
                    s.show_pos = 0;
                    s.clone_ahead = 1;  // re-cloned here, before any step pops the new queue (:321)
                } else {
                    done = true;
                    success = 1;  // also when the last command failed (quirk A.9#1)
                    reward += k.p.reward_episode_success;
                }
            }
            s.cs = 1;
        }
        if (s.tiles_on) {
            if ((s.cv % s.E == 0) && s.cv > 0) {
                s.tiles_on = 0;
                s.cv = 0;
                if (s.cc < s.n) {
                    int dx, dy;
                    mg_cmd_delta(s.commands[s.cc], dx, dy);
                    int tx = s.tx + dx, ty = s.ty + dy;
                    if (endless) {
                        tx = mg_pymod(tx, 6);
                        ty = mg_pymod(ty, 6);
                    }
                    s.tx = (int8_t)tx;
                    s.ty = (int8_t)ty;
                }
            } else {
                if (!on_target) {
                    done = true;
                    reward = k.p.reward_command_failure;  // assignment, not += (quirk A.9#2)
                }
                s.cv++;
            }
        } else {
            s.cs++;
        }
    }
    if (endless) {
        s.t++;
        if (s.t == k.p.max_steps) done = true;
    }
    s.ep_reward = s.ep_reward + reward;
    s.ep_len++;
    reward_out = reward;
    if (!endless) {
        // the sprite handle (rotated_agent_surface / rect, mortar_mayhem.py:291,295) = what this step drew; it is not
        // refreshed by reset(), which only the ground-truth render shows (the endless variant keeps it above)
        s.draw_cx = (int16_t)sx;
        s.draw_cy = (int16_t)sy;
        s.draw_rot = (uint8_t)srot;
        s.has_draw = 1;
    }
    mg_mm_frame(k, s, sx, sy, srot, icon, f);
    if (gt) {
        gt[0] = (double)s.tx / 5.0;
        gt[1] = (double)s.ty / 5.0;
    }
    if (done) {
        s.live = 0;
        if (info) {
            info[MG_INFO_REWARD] = s.ep_reward;
            info[MG_INFO_LENGTH] = (double)s.ep_len;
            if (endless) {
                info[MG_INFO_SUCCESS] = 0.0;
                info[MG_INFO_A] = (double)s.total_completed;
                info[MG_INFO_B] = s.n > 1 ? (double)(s.n - 1) : 0.0;
            } else {
                info[MG_INFO_SUCCESS] = (double)success;
                info[MG_INFO_A] = (double)((int)s.cc - 1 + success) / (double)s.n;
                info[MG_INFO_B] = 0.0;
            }
            info[MG_INFO_GT0] = (double)s.tx / 5.0;
            info[MG_INFO_GT1] = (double)s.ty / 5.0;
            info[7] = info[8] = info[9] = 0.0;
        }
    }
    return done;
}

// MortarMayhemB: _encode_commands_one_hot (mortar_mayhem_b.py:102-131) — slot order stay, right, left, up, down,
// right_down, right_up, left_down, left_up; written by every lane of the team
template <class Ctx>
MG_HD void mg_mm_write_vector(Ctx& c, const MgMMState& s, float* vec) {
    const uint64_t slot = 0x876503241ull;  // Command.COMMANDS index -> one-hot slot {1, 4, 2, 3, 0, 5, 6, 7, 8}, 4 bits each
    for (int i = c.lane; i < MG_MM_B_VECTOR; i += c.nl) {
        int cmd = i / 9, j = i - 9 * cmd;
        vec[i] = (cmd < (int)s.n && (int)((slot >> (4 * s.commands[cmd])) & 15u) == j) ? 1.0f : 0.0f;
    }
}
