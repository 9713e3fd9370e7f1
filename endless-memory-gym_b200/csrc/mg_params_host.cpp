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
    if (p->variant < MG_MM_CONTINUOUS || p->variant > MG_MM_B_GRID) {
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
        const bool task_b = p->variant == MG_MM_B || p->variant == MG_MM_B_GRID;  // mortar_mayhem_b.py:52
        if (!check_list(p->command_count, 1, task_b ? MG_MM_B_MAX_COMMANDS : MG_MM_MAX_COMMANDS, "command_count"))
            return MG_EINVAL;
        max_n = list_max(p->command_count);
    }
    const bool shows = !(p->variant == MG_MM_B || p->variant == MG_MM_B_GRID);
    if (shows && !check_list(p->command_show_duration, 0, 4096, "command_show_duration")) return MG_EINVAL;
    if (shows && !check_list(p->command_show_delay, 0, 4096, "command_show_delay")) return MG_EINVAL;
    if (!check_list(p->explosion_duration, 1, 65535, "explosion_duration")) return MG_EINVAL;
    if (!check_list(p->explosion_delay, 1, 65535, "explosion_delay")) return MG_EINVAL;
    if (shows && list_min(p->command_show_duration) + list_min(p->command_show_delay) < 1) {
        mg_set_err("command_show_duration + command_show_delay must be >= 1 (the reference pops an empty list otherwise)");
        return MG_EINVAL;
    }
    if (shows && (long)max_n * (list_max(p->command_show_duration) + list_max(p->command_show_delay)) > 65535) {
        mg_set_err("command visualisation queue longer than 65535 steps");
        return MG_EINVAL;
    }
    if (!(p->agent_speed >= 0.0 && p->agent_speed < 84.0)) {
        mg_set_err("agent_speed must be in [0, 84)");
        return MG_EINVAL;
    }
    if (p->variant != MG_MM_GRID && p->variant != MG_MM_B_GRID) {
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
    double scale = (p->variant == MG_MM_GRID || p->variant == MG_MM_B_GRID) ? 0.25 : p->agent_scale;
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
    if (p->variant != MG_MP_CONTINUOUS && p->variant != MG_MP_GRID) {
        mg_set_err("variant");
        return MG_EINVAL;
    }
    int radius = (int)(25 * p->agent_scale), hand = (int)(10 * p->agent_scale);
    if (p->variant == MG_MP_CONTINUOUS && (!(p->agent_scale > 0.0) || radius * 2 + hand + 14 > 255 || radius > 42)) {
        mg_set_err("agent_scale out of range");
        return MG_EINVAL;
    }
    return 0;
}

void mg_mp_derive(const mg_mp_params* p, MgMPK* k) {
    memset(k, 0, sizeof(*k));
    k->p = *p;
    // GridMysteryPathEnv places the agent with the hard-coded int(25 * SCALE) (mystery_path_grid.py:186)
    k->radius = p->variant == MG_MP_GRID ? (int)(25 * 0.25) : (int)(25 * p->agent_scale);
    double inv = 1.0 / sqrt(2.0);
    k->step_axis = (int)(1.0 * p->agent_speed);
    k->step_diag = (int)(inv * p->agent_speed);
}


extern "C" int mg_ep_check_params(const mg_ep_params* p) {
    if (!p) {
        mg_set_err("params is NULL");
        return MG_EINVAL;
    }
    if (!(p->agent_speed >= 0.0 && p->agent_speed < 84.0)) {
        mg_set_err("agent_speed must be in [0, 84)");
        return MG_EINVAL;
    }
    int radius = (int)(25 * p->agent_scale), hand = (int)(10 * p->agent_scale);
    if (!(p->agent_scale > 0.0) || radius * 2 + hand + 14 > 255 || radius > 38) {
        mg_set_err("agent_scale out of range");
        return MG_EINVAL;
    }
    if (p->stamina_level < 1 || p->max_steps == 0) {
        mg_set_err("stamina_level must be >= 1 and max_steps != 0");
        return MG_EINVAL;
    }
    // depth of the drawn past path = int(camera_offset_scale) is NOT clamped by the reference (only the camera is)
    if (!(p->camera_offset_scale > -1.0 && p->camera_offset_scale < 17.0)) {
        mg_set_err("camera_offset_scale must be in (-1, 17)");
        return MG_EINVAL;
    }
    return 0;
}

void mg_ep_derive(const mg_ep_params* p, MgEPK* k) {
    memset(k, 0, sizeof(*k));
    k->p = *p;
    k->radius = (int)(25 * p->agent_scale);
    k->step_axis = (int)(1.0 * p->agent_speed);
    double cos_ = p->camera_offset_scale;
    cos_ = cos_ < 0.0 ? 0.0 : (cos_ > 5.5 ? 5.5 : cos_);  // max(0, min(5.5, scale)) (:213)
    k->camera_offset = -12.0 * cos_;
    // sprite side = 2 * radius + int(10 * scale) + 14... taken from the asset header at run time; here only the
    // episode-constant blit x: agent.rect.topleft[0] - camera_offset with rect.centerx = radius (:219-222)
    int hand = (int)(10 * p->agent_scale);
    int dim = k->radius * 2 + hand + 14;
    k->agent_half = dim >> 1;
    k->agent_draw_x = (int)((double)(k->radius - k->agent_half) - k->camera_offset);
    k->depth = (int)p->camera_offset_scale;
}

extern "C" int mg_ss_check_params(const mg_ss_params* p) {
    if (!p) {
        mg_set_err("params is NULL");
        return MG_EINVAL;
    }
    const bool endless = p->variant == MG_SS_ENDLESS;
    if (p->variant != MG_SS_FINITE && !endless) {
        mg_set_err("variant must be MG_SS_FINITE or MG_SS_ENDLESS");
        return MG_EINVAL;
    }
    if (endless) {
        // step() paints the action boxes with self.action_colors unconditionally, which only a reset with
        // show_last_action creates (endless_searing_spotlights.py:349-354, 421-424): AttributeError otherwise
        if (!p->show_last_action) {
            mg_set_err("show_last_action=False is not supported (the reference raises AttributeError in step)");
            return MG_EINVAL;
        }
        if (p->spawn_interval < 1) {
            mg_set_err("spawn_interval must be >= 1");
            return MG_EINVAL;
        }
    }
    if (!endless && !p->use_exit) {  // the reference crashes in reset (self.exit is undefined, searing_spotlights.py:416,432)
        mg_set_err("use_exit=False is not supported (the reference raises AttributeError)");
        return MG_EINVAL;
    }
    if (!endless && !check_list(p->num_coins, 1, MG_SS_MAXCOINS, "num_coins")) return MG_EINVAL;  // 0 coins: ZeroDivisionError at done (:553)
    if (!(p->agent_health > 0.0)) {
        mg_set_err("agent_health must be > 0");
        return MG_EINVAL;
    }
    if (!(p->agent_speed >= 0.0 && p->agent_speed < 84.0)) {
        mg_set_err("agent_speed must be in [0, 84)");
        return MG_EINVAL;
    }
    int radius = (int)(25 * p->agent_scale), hand = (int)(10 * p->agent_scale);
    if (!(p->agent_scale > 0.0) || radius * 2 + hand + 14 > 255 || radius > 38) {
        mg_set_err("agent_scale out of range");
        return MG_EINVAL;
    }
    int rlo = (int)p->spot_min_radius, rhi = (int)(p->spot_max_radius + 1);
    if (rlo < 1 || rhi <= rlo || rhi - 1 > MG_MAX_DISC_R) {
        mg_set_err("spot radius must satisfy 1 <= int(min) < int(max + 1) <= MG_MAX_DISC_R + 1");
        return MG_EINVAL;
    }
    if (!(p->spot_min_speed > 0.0) || !(p->spot_max_speed >= p->spot_min_speed) || !(p->spot_max_speed <= 1.0)) {
        mg_set_err("spot speeds must satisfy 0 < min <= max <= 1");
        return MG_EINVAL;
    }
    if ((int)(10 * p->coin_scale) > MG_MAX_DISC_R || p->coin_scale < 0.0 ||
        (!endless && ((int)(20 * p->exit_scale) < 2 || (int)(20 * p->exit_scale) > 84))) {
        mg_set_err("coin_scale / exit_scale out of range");
        return MG_EINVAL;
    }
    if (p->initial_spawns < 0 || p->num_spawns < 0 || p->max_steps == 0) {
        mg_set_err("initial_spawns, num_spawns must be >= 0 and max_steps != 0");
        return MG_EINVAL;
    }
    if (p->light_dim_off_duration < 0 || p->light_dim_off_duration > 255) {
        mg_set_err("light_dim_off_duration must be 0..255");
        return MG_EINVAL;
    }
    // spotlights alive at once: the initial ones plus one per spawn interval within the longest traversal
    int interval0 = endless ? p->spawn_interval : (int)(p->initial_spawn_interval + p->spawn_interval_threshold);
    if (!endless && p->num_spawns > 0 && interval0 < 1) {
        mg_set_err("initial_spawn_interval + spawn_interval_threshold must be >= 1");
        return MG_EINVAL;
    }
    double life = 1.0 / p->spot_min_speed + 2.0;
    if (p->max_steps > 0 && (double)p->max_steps < life) life = (double)p->max_steps;
    long spawned = (endless || p->num_spawns > 0) ? (long)(life / (interval0 > 0 ? interval0 : 1)) + 1 : 0;
    if (!endless && spawned > p->num_spawns) spawned = p->num_spawns;
    if (p->initial_spawns + spawned > MG_SS_MAXSPOTS) {
        mg_set_err("too many concurrent spotlights for the env record (MG_SS_MAXSPOTS = 32)");
        return MG_EINVAL;
    }
    return 0;
}

// where the spotlight layer and the top bar sit in both painter's orders (the kernel reads them per frame)
static void ss_layer_indices(MgSSK* k) {
    for (int alt = 0; alt < 2; ++alt) {
        const int32_t* L = alt ? k->layers_alt : k->layers;
        k->spot_li[alt] = 1;
        k->bar_li[alt] = 0;
        for (int li = 0; li < k->n_layers; ++li) {
            if (L[li] == MG_SS_L_SPOT) k->spot_li[alt] = li;
            if (L[li] == MG_SS_L_BAR) k->bar_li[alt] = li;
        }
    }
}

void mg_ss_derive(const mg_ss_params* p, MgSSK* k) {
    memset(k, 0, sizeof(*k));
    k->p = *p;
    k->radius = (int)(25 * p->agent_scale);
    double inv = 1.0 / sqrt(2.0);
    k->step_axis = (int)(1.0 * p->agent_speed);
    k->step_diag = (int)(inv * p->agent_speed);
    k->coin_radius = (int)(10 * p->coin_scale);
    k->spot_rlo = (int)p->spot_min_radius;
    k->spot_rhi = (int)(p->spot_max_radius + 1);
    k->interval0 = (int)(p->initial_spawn_interval + p->spawn_interval_threshold);
    k->alpha_inc = p->light_dim_off_duration > 0 ? (int)(255.0 / p->light_dim_off_duration) : 0;
    int q = MG_W / 4;
    if (p->show_last_action) {
        k->cb_x = (int)(q * 2.75);
        k->cb_w = (int)(q * 0.5);
    } else {
        k->cb_x = (int)(q * 2);
        k->cb_w = (int)(q * 2);
    }
    k->clamp_off = (int)(30 * 0.25);
    k->exit_radius = 20 / 2 * p->exit_scale;
    {   // Exit.is_agent_inside (pygame_assets.py:207-220): distance <= radius + agent.radius on integer coordinates
        const double T = k->exit_radius + (double)k->radius;
        int d2 = -1;
        while (d2 + 1 <= 2 * 200 * 200 && sqrt((double)(d2 + 1)) <= T) ++d2;
        k->exit_d2max = d2;
    }
    k->spawn_base = sqrt(pow(84.0, 2) + pow(84.0, 2)) / 2;
    // painter's order: python list inserts of searing_spotlights.py:526-542 (an index past the end appends)
    int order[6], n = 0;
    auto insert = [&](int at, int v) {
        if (at > n) at = n;
        for (int i = n; i > at; --i) order[i] = order[i - 1];
        order[at] = v;
        n++;
    };
    k->endless = p->variant == MG_SS_ENDLESS ? 1 : 0;
    if (k->endless) {
        // endless_searing_spotlights.py:376-391 / 463-478: bg, spot, bar; coin above the spotlights while visible or
        // fresh (layers_alt) else below (layers); then the agent. No exit layer.
        for (int fresh = 0; fresh < 2; ++fresh) {
            n = 0;
            order[n++] = MG_SS_L_BG;
            order[n++] = MG_SS_L_SPOT;
            order[n++] = MG_SS_L_BAR;
            int sid = 1;
            if (p->coins_visible || fresh)
                insert(sid + 1, MG_SS_L_COIN);
            else
                insert(sid++, MG_SS_L_COIN);
            if (p->agent_visible)
                insert(sid + 3, MG_SS_L_AGENT);
            else
                insert(sid++, MG_SS_L_AGENT);
            for (int i = 0; i < n; ++i) (fresh ? k->layers_alt : k->layers)[i] = order[i];
        }
        k->n_layers = n;
        k->cb_x = (int)(q * 2.75);
        k->cb_w = (int)(q * 0.5);
        ss_layer_indices(k);
        return;
    }
    order[n++] = MG_SS_L_BG;
    order[n++] = MG_SS_L_SPOT;
    order[n++] = MG_SS_L_BAR;
    int sid = 1;
    if (p->coins_visible)
        insert(sid + 1, MG_SS_L_COIN);
    else
        insert(sid++, MG_SS_L_COIN);
    if (p->exit_visible)
        insert(sid + 2, MG_SS_L_EXIT);
    else
        insert(sid++, MG_SS_L_EXIT);
    if (p->agent_visible)
        insert(sid + 3, MG_SS_L_AGENT);
    else
        insert(sid++, MG_SS_L_AGENT);
    k->n_layers = n;
    for (int i = 0; i < n; ++i) k->layers[i] = k->layers_alt[i] = order[i];
    ss_layer_indices(k);
}
