// mg_params_host.cpp — validation of the reset options (the reference's `process_reset_params`
// asserts, e.g. memory_gym/mortar_mayhem.py:37-54, plus the limits of the fixed-size env records)
// and derivation of the per-launch constants. No CUDA here.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "mg_params_host.h"

static thread_local char g_err[256] = "";
void mg_set_err(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }
extern "C" const char* mg_last_error(void) { return g_err; }
extern "C" int mg_abi_version(void) { return MG_ABI_VERSION; }

static bool check_list(const mg_choice_list& c, int lo, int hi, const char* name) {
    if (c.n < 1 || c.n > MG_MAX_CHOICES) {
        snprintf(g_err, sizeof(g_err), "%s: list length must be 1..%d", name, MG_MAX_CHOICES);
        return false;
    }
    for (int i = 0; i < c.n; ++i)
        if (c.v[i] < lo || c.v[i] > hi) {
            snprintf(g_err, sizeof(g_err), "%s: value %d outside [%d, %d]", name, c.v[i], lo, hi);
            return false;
        }
    return true;
}
static int list_min(const mg_choice_list& c) {
    int m = c.v[0];
    for (int i = 1; i < c.n; ++i) m = c.v[i] < m ? c.v[i] : m;
    return m;
}
static int list_max(const mg_choice_list& c) {
    int m = c.v[0];
    for (int i = 1; i < c.n; ++i) m = c.v[i] > m ? c.v[i] : m;
    return m;
}

extern "C" int mg_mm_check_params(const mg_mm_params* p) {
    if (!p) {
        mg_set_err("params is NULL");
        return MG_EINVAL;
    }
    if (p->variant < MG_MM_CONTINUOUS || p->variant > MG_MM_ENDLESS) {
        mg_set_err("variant");
        return MG_EINVAL;
    }
    if (p->allowed_commands < 4 || p->allowed_commands > 9) {  // mortar_mayhem.py:52
        mg_set_err("allowed_commands must be 4..9");
        return MG_EINVAL;
    }
    int max_n;
    if (p->variant == MG_MM_ENDLESS) {
        if (p->initial_command_count < 1 || p->initial_command_count > MG_MM_MAX_COMMANDS) {
            mg_set_err("initial_command_count must be 1..MG_MM_MAX_COMMANDS");
            return MG_EINVAL;
        }
        max_n = MG_MM_MAX_COMMANDS;
    } else {
        if (p->arena_size < 2 || p->arena_size > 6) {  // mortar_mayhem.py:53
            mg_set_err("arena_size must be 2..6");
            return MG_EINVAL;
        }
        if (!check_list(p->command_count, 1, MG_MM_MAX_COMMANDS, "command_count")) return MG_EINVAL;
        max_n = list_max(p->command_count);
    }
    if (!check_list(p->command_show_duration, 0, 4096, "command_show_duration")) return MG_EINVAL;
    if (!check_list(p->command_show_delay, 0, 4096, "command_show_delay")) return MG_EINVAL;
    if (!check_list(p->explosion_duration, 1, 65535, "explosion_duration")) return MG_EINVAL;
    if (!check_list(p->explosion_delay, 1, 65535, "explosion_delay")) return MG_EINVAL;
    if (list_min(p->command_show_duration) + list_min(p->command_show_delay) < 1) {
        mg_set_err("command_show_duration + command_show_delay must be >= 1 (the reference pops an empty list otherwise)");
        return MG_EINVAL;
    }
    if ((long)max_n * (list_max(p->command_show_duration) + list_max(p->command_show_delay)) > 65535) {
        mg_set_err("command visualisation queue longer than 65535 steps");
        return MG_EINVAL;
    }
    if (!(p->agent_speed >= 0.0 && p->agent_speed < 84.0)) {
        mg_set_err("agent_speed must be in [0, 84)");
        return MG_EINVAL;
    }
    if (p->variant != MG_MM_GRID) {
        int radius = (int)(25 * p->agent_scale), hand = (int)(10 * p->agent_scale);
        if (!(p->agent_scale > 0.0) || radius * 2 + hand + 14 > 255) {
            mg_set_err("agent_scale out of range");
            return MG_EINVAL;
        }
    }
    return 0;
}

void mg_mm_derive(const mg_mm_params* p, MgMMK* k) {
    memset(k, 0, sizeof(*k));
    k->p = *p;
    k->size = p->variant == MG_MM_ENDLESS ? 6 : p->arena_size;
    k->origin = MG_W / 2 - ((MG_TILE * k->size) >> 1);
    double scale = p->variant == MG_MM_GRID ? 0.25 : p->agent_scale;
    k->radius = (int)(25 * scale);
    // velocity.normalize() * speed (character_controller.py:124): 1/sqrt(2) per diagonal component
    double inv = 1.0 / sqrt(2.0);
    k->v_axis = 1.0 * p->agent_speed;
    k->v_diag = inv * p->agent_speed;
    k->step_axis = (int)k->v_axis;
    k->step_diag = (int)k->v_diag;
    k->wrap_off = k->radius * 0.5;
}


extern "C" int mg_mp_check_params(const mg_mp_params* p) {
    if (!p) {
        mg_set_err("params is NULL");
        return MG_EINVAL;
    }
    if (!check_list(p->cardinal_origin_choice, 0, 3, "cardinal_origin_choice")) return MG_EINVAL;
    if (!(p->agent_speed >= 0.0 && p->agent_speed < 84.0)) {
        mg_set_err("agent_speed must be in [0, 84)");
        return MG_EINVAL;
    }
    int radius = (int)(25 * p->agent_scale), hand = (int)(10 * p->agent_scale);
    if (!(p->agent_scale > 0.0) || radius * 2 + hand + 14 > 255 || radius > 42) {
        mg_set_err("agent_scale out of range");
        return MG_EINVAL;
    }
    return 0;
}

void mg_mp_derive(const mg_mp_params* p, MgMPK* k) {
    memset(k, 0, sizeof(*k));
    k->p = *p;
    k->radius = (int)(25 * p->agent_scale);
    double inv = 1.0 / sqrt(2.0);
    k->step_axis = (int)(1.0 * p->agent_speed);
    k->step_diag = (int)(inv * p->agent_speed);
}

