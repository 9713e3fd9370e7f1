/* memgym_b200.h — C ABI of libmemgym_b200.so: batched Memory-Gym reset/step on NVIDIA B200 (sm_100a).
 *
 * Drop-in boundary for ONE hot path of MarcoMeter/endless-memory-gym: `env.reset()` / `env.step()`
 * of N independent environment instances, producing uint8 observations [N][84][84][3] (pygame
 * `surfarray.array3d` layout, i.e. [x][y][c]), rewards, dones and episode statistics.
 * The reference has no FFI (it is pure Python on pygame); each entry point below names the
 * reference interface it replaces (paths relative to the reference checkout):
 *
 *   mg_mm_reset / mg_mm_step   <-  MortarMayhemEnv.reset/step          memory_gym/mortar_mayhem.py:206-272, 274-369
 *                                  GridMortarMayhemEnv.reset/step      memory_gym/mortar_mayhem_grid.py:213-278, 280-375
 *                                  EndlessMortarMayhemEnv.reset/step   memory_gym/endless_mortar_mayhem.py:194-259, 261-373
 *                                  MortarMayhemTaskBEnv.reset/step     memory_gym/mortar_mayhem_b.py:133-204, 206-290
 *                                  GridMortarMayhemTaskBEnv.reset/step memory_gym/mortar_mayhem_b_grid.py:133-202, 204-288
 *                                  CharacterController.step & co.      memory_gym/character_controller.py:89-146, 177-210, 226-283
 *                                  MortarArena / Command               memory_gym/pygame_assets.py:241-418
 *   mg_mp_reset / mg_mp_step   <-  MysteryPathEnv.reset/step           memory_gym/mystery_path.py:130-200, 202-276
 *                                  GridMysteryPathEnv.reset/step       memory_gym/mystery_path_grid.py:129-200, 202-276
 *                                  MysteryPath.__init__ (noisy A*)     memory_gym/pygame_assets.py:606-736
 *   mg_ss_reset / mg_ss_step   <-  SearingSpotlightsEnv.reset/step     memory_gym/searing_spotlights.py:332-450, 452-562
 *                                  EndlessSearingSpotlightsEnv.reset/step  memory_gym/endless_searing_spotlights.py:294-407, 409-504
 *                                  Spotlight / Coin / Exit / sampler   memory_gym/pygame_assets.py:7-239
 *   mg_ep_reset / mg_ep_step   <-  EndlessMysteryPathEnv.reset/step    memory_gym/endless_mystery_path.py:192-268, 270-444
 *                                  EndlessMysteryPath (segments)       memory_gym/pygame_assets.py:495-604
 *   mg_*_build_assets          <-  CharacterController.create_character_sprites  character_controller.py:29-75,
 *                                  Command.__init__ pygame_assets.py:254-304, MortarTile :316-343, Exit.draw :191-205,
 *                                  Coin.draw :145-152, fall-off cross mystery_path.py:175-182
 *                                  (i.e. the pygame draw/transform calls the reference makes per reset / per step)
 *   mg_seed_pcg64              <-  gymnasium.utils.seeding.np_random(seed) = Generator(PCG64(SeedSequence(seed)))
 *                                  (call site: super().reset(seed=seed), mortar_mayhem.py:217)
 *
 * Conventions
 *   - every pointer marked "device" is CUDA device memory owned by the caller (e.g. a torch tensor);
 *     the library never allocates or frees device memory; its only process state are per-thread caches of
 *     launch geometry and derived option blocks (no environment variable is read).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream). Calls are asynchronous.
 *   - launches run on the device that owns `state` (looked up from the pointer), whatever the caller's
 *     current device is; `stream` and every other device pointer must belong to that device.
 *   - return value: 0 on success, a cudaError_t (> 0) from the launch, or a negative MG_E* code for
 *     bad arguments. No exceptions cross this boundary.
 *   - there is NO CPU implementation behind the reset/step symbols: without a CUDA device they fail.
 */
#ifndef MEMGYM_B200_H
#define MEMGYM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MG_ABI_VERSION 7

#define MG_SCREEN 84
#define MG_OBS_BYTES (84 * 84 * 3) /* 21168 bytes per env, [x][y][c] */
#define MG_DEBUG_DIM 336           /* debug_rgb_array: the 84x84 ground-truth surface scaled 4x */
#define MG_DEBUG_BYTES (336 * 336 * 3) /* per env, [y][x][c] like the reference's render() */
#define MG_INFO_COLS 10            /* doubles per env in the info matrix */
#define MG_GT_COLS 4               /* doubles per env in the ground-truth matrix */
#define MG_TRACE_COLS 16
#define MG_MAX_CHOICES 16          /* max length of a list-valued reset option */
#define MG_MM_MAX_COMMANDS 152     /* capacity of the per-env command list */
#define MG_MM_STATE_BYTES 256      /* one env record (Mortar Mayhem family) */
#define MG_MP_STATE_BYTES 128      /* one env record (Mystery Path) */
#define MG_SS_STATE_BYTES 1152     /* one env record (Searing Spotlights) */
#define MG_EP_STATE_BYTES 4096     /* one env record (Endless Mystery Path): 2048 + the 2048-byte look-ahead block */
#define MG_EP_MAX_SEGMENTS 45      /* capacity of the per-env path: 45 segments x 8 columns = 360 columns */

#define MG_EINVAL (-1)  /* bad argument / unsupported option value */
#define MG_ENODEV (-2)  /* no usable CUDA device */

/* info matrix columns (valid for env i only on steps where done[i] != 0) */
#define MG_INFO_REWARD 0   /* sum of step rewards, left-to-right in double */
#define MG_INFO_LENGTH 1   /* episode length in steps */
#define MG_INFO_SUCCESS 2  /* MM/MM-Grid/MP/SS: success flag                                     */
#define MG_INFO_A 3        /* MM: commands_completed (fraction); EMM: commands_completed (count); */
                           /* MP/EMP: num_fails; SS/ESS: agent_health (fraction)                  */
#define MG_INFO_B 4        /* EMM: max_command_sequence; SS: coins_collected (fraction);          */
                           /* ESS: coins_collected (count); EMP: max_x                            */
#define MG_INFO_GT0 5      /* Endless: ground_truth of the terminal step, MG_GT_COLS columns 5..8  */
#define MG_INFO_GT1 6
#define MG_INFO_GT2 7
#define MG_INFO_GT3 8
#define MG_INFO_C 9        /* EMP: tiles_visited                                                  */

/* mg_step_io.flags */
#define MG_FLAG_AUTORESET 1u /* step: envs that finish are reset in the same call (their stream continues like  */
                             /* `env.reset()` with seed=None); obs/ground_truth then hold the new episode's first */
                             /* frame and the terminal frame goes to final_obs. Without it a finished env keeps   */
                             /* its terminal frame and ignores further steps until mg_*_reset.                    */

#define MG_FLAG_STREAM_OBS 2u /* write observations with an L2 evict-first policy (they are not re-read soon) */
#define MG_FLAG_NO_HELPERS 4u /* measurement knob: Endless-MysteryPath launches without path-generation helper warps */
#define MG_FLAG_NO_PDL 8u     /* measurement knob: launch without programmatic dependent launch */

typedef struct mg_choice_list {
    int32_t n;
    int32_t v[MG_MAX_CHOICES];
} mg_choice_list;

/* ---- Mortar Mayhem family ------------------------------------------------------------------- */
enum { MG_MM_CONTINUOUS = 0, MG_MM_GRID = 1, MG_MM_ENDLESS = 2, MG_MM_B = 3, MG_MM_B_GRID = 4 };
#define MG_MM_B_MAX_COMMANDS 20   /* MortarMayhemB: commands are observed as a [20][9] one-hot vector */
#define MG_MM_B_VECTOR 180

/* Reset options, batch-uniform. Field names follow the reference's default_reset_parameters
 * (mortar_mayhem.py:21-35, mortar_mayhem_grid.py:21-34, endless_mortar_mayhem.py:21-35). */
typedef struct mg_mm_params {
    int32_t variant;               /* MG_MM_* */
    int32_t arena_size;            /* 2..6 (Endless: always 6) */
    int32_t allowed_commands;      /* 4..9 */
    int32_t visual_feedback;
    int32_t max_steps;             /* Endless only; -1 = none */
    int32_t initial_command_count; /* Endless only */
    mg_choice_list command_count;  /* MM / MM-Grid */
    mg_choice_list command_show_duration;
    mg_choice_list command_show_delay;
    mg_choice_list explosion_duration;
    mg_choice_list explosion_delay;
    double agent_scale;
    double agent_speed;
    double reward_command_failure;
    double reward_command_success;
    double reward_episode_success; /* Endless: reward_new_command_success */
} mg_mm_params;

/* ---- Mystery Path ---------------------------------------------------------------------------- */
enum { MG_MP_CONTINUOUS = 0, MG_MP_GRID = 1 };
typedef struct mg_mp_params {
    int32_t variant;   /* MG_MP_* */
    int32_t pad0;
    int32_t max_steps;
    int32_t show_origin;
    int32_t show_goal;
    int32_t visual_feedback;
    mg_choice_list cardinal_origin_choice;
    double agent_scale;
    double agent_speed;
    double reward_goal;
    double reward_fall_off;
    double reward_path_progress;
    double reward_step;
} mg_mp_params;

/* ---- Searing Spotlights ---------------------------------------------------------------------- */
enum { MG_SS_FINITE = 0, MG_SS_ENDLESS = 1 };
/* Field names follow default_reset_parameters (searing_spotlights.py:22-64, endless_searing_spotlights.py:22-58);
 * the endless variant ignores num_spawns, the three spawn-interval fields, num_coins, the exit fields,
 * reward_exit and reward_max_steps. */
typedef struct mg_ss_params {
    int32_t max_steps;
    int32_t initial_spawns;
    int32_t num_spawns;
    int32_t visual_feedback;
    int32_t black_background;
    int32_t hide_chessboard;
    int32_t light_dim_off_duration;
    int32_t light_threshold;
    int32_t coins_visible;
    int32_t use_exit;
    int32_t exit_visible;
    int32_t agent_visible;
    int32_t sample_agent_position;
    int32_t show_last_action;
    int32_t show_last_positive_reward;
    int32_t variant;               /* MG_SS_* */
    mg_choice_list num_coins;
    double initial_spawn_interval;
    double spawn_interval_threshold;
    double spawn_interval_decay;
    double spot_min_radius;
    double spot_max_radius;
    double spot_min_speed;
    double spot_max_speed;
    double spot_damage;
    double coin_scale;
    double exit_scale;
    double agent_speed;
    double agent_health;
    double agent_scale;
    double reward_inside_spotlight;
    double reward_outside_spotlight;
    double reward_death;
    double reward_exit;
    double reward_max_steps;
    double reward_coin;
    /* Endless-SearingSpotlights-v0 only (endless_searing_spotlights.py:22-58) */
    int32_t steps_per_coin;
    int32_t spawn_interval;
    int32_t coin_enabled;
    int32_t coin_show_duration;
} mg_ss_params;

/* ---- Endless Mystery Path --------------------------------------------------------------------- */
/* Field names follow default_reset_parameters (endless_mystery_path.py:22-37). Actions: Discrete(4)
 * (0 no-op, 1 right, 2 up, 3 down). show_origin has no effect in the reference (:151, the condition can
 * never hold) and is accepted for compatibility only. */
typedef struct mg_ep_params {
    int32_t max_steps;
    int32_t show_origin;
    int32_t show_past_path;
    int32_t show_background;
    int32_t show_stamina;
    int32_t visual_feedback;
    int32_t stamina_level;
    int32_t pad0;
    double agent_scale;
    double agent_speed;
    double camera_offset_scale;
    double reward_fall_off;
    double reward_path_progress;
    double reward_path_progress_dense;
    double reward_step;
} mg_ep_params;

/* ---- per-call I/O (all device pointers; nullable ones are marked) ----------------------------- */
typedef struct mg_step_io {
    const int64_t* actions; /* [n][2] (MultiDiscrete([3,3])) or [n] (Discrete(4), MM-Grid); unused by reset */
    uint8_t* obs;           /* [n][84][84][3] */
    uint8_t* final_obs;     /* nullable: [n][84][84][3], terminal frame, written only where done (auto-reset) */
    double* reward;         /* [n]; unused by reset */
    uint8_t* done;          /* [n]; unused by reset */
    double* info;           /* [n][MG_INFO_COLS], written only where done; unused by reset */
    double* ground_truth;   /* nullable: [n][MG_GT_COLS], every step and reset (Endless envs: EMM target / 5.0 in  */
                            /* columns 0..1; ESS (agent x, y, coin x, y) / 84; EMP one-hot (east, north, south)  */
                            /* of the next path tile in columns 0..2; unused columns are left alone)           */
    const uint8_t* reset_mask; /* reset only, nullable: [n], reset env i iff mask[i] != 0 */
    uint32_t flags;         /* MG_FLAG_* */
    uint32_t warps_per_cta; /* 0 = library default; tuning knob (frame buffers resident per SM) */
    float* vector_obs;      /* nullable: [n][MG_MM_B_VECTOR] (MortarMayhemB: one-hot commands of the running episode) */
    uint64_t* trace;        /* nullable: [n][MG_TRACE_COLS] uint64, per-env phase timestamps (globaltimer ns):   */
                            /* 0 kernel entry, 1 prologue done, 2 dependency wait done, 3 record in smem,      */
                            /* 4 logic done, 5 frame composed, 6 bulk store issued, 7 SM id, 8.. env-type specific  */
    uint8_t* debug_obs;     /* nullable: [n][336][336][3] ([y][x][c]): render_mode="debug_rgb_array" of the reference     */
                            /* (_build_debug_surface + transform.scale, e.g. mortar_mayhem.py:105-136, 397-399) for the   */
                            /* frame written to obs, as if render() were called once after every reset() / step().        */
                            /* Every env type whose reference class has one (the MortarMayhemB variants do not); Endless-   */
                            /* MysteryPath: launch with MG_FLAG_NO_HELPERS (the render needs the finished path).            */
} mg_step_io;

/* A whole rollout in one launch: T = n_steps consecutive steps with given actions (the reference's bench.py
 * loop with pre-sampled random actions; scripted policies). Every warp runs all T steps of its envs back to
 * back with the env record resident on chip; finished envs are always auto-reset. */
typedef struct mg_rollout_io {
    const int64_t* actions; /* [T][n][2] or [T][n] */
    uint8_t* obs;           /* [ring][n][84][84][3]: step t writes slot t % ring (a rollout ring / [T] buffer) */
    uint8_t* final_obs;     /* nullable: [n][84][84][3], terminal frame of the env's LAST finished episode */
    double* reward;         /* [T][n] */
    uint8_t* done;          /* [T][n] */
    double* info;           /* [T][n][MG_INFO_COLS] if info_per_step, else [n][MG_INFO_COLS] (last finished episode) */
    double* ground_truth;   /* nullable: [n][MG_GT_COLS], of the last step */
    int32_t n_steps;
    int32_t ring;
    int32_t info_per_step;
    uint32_t flags;
    uint32_t warps_per_cta; /* 0 = library picks the count that balances envs per warp */
    uint32_t pad0;
    uint64_t* trace;        /* nullable: [n][MG_TRACE_COLS] phase timestamps of each env's LAST step (see mg_step_io) */
    float* vector_obs;      /* nullable: [n][MG_MM_B_VECTOR], as in mg_step_io (state after the last step) */
    uint32_t* sched;        /* nullable: [2] uint32, zeroed ONCE by the caller and then owned by the library: a global   */
                            /* work counter for env -> warp scheduling across SMs (SMs that stream faster take more   */
                            /* envs). NULL = every SM works through its own share. Results never depend on it. One    */
                            /* buffer per stream of launches.                                                          */
} mg_rollout_io;

/* ---- assets & seeding (host side) -------------------------------------------------------------- */
/* Draw the static sprites / background templates for an option set with the library's own
 * restatement of pygame's draw/transform routines and pack them into `host_out`. Returns the blob
 * size in bytes (call with host_out = NULL to query), or a negative MG_E* code. The caller uploads
 * the blob to 128-byte aligned device memory and passes it as `assets`.
 * thick_circle_legacy != 0 selects the pygame <= 2.1.2 thick-circle algorithm for the "stay" icon. */
long mg_mm_build_assets(const mg_mm_params* p, int thick_circle_legacy, void* host_out, size_t capacity);
long mg_mp_build_assets(const mg_mp_params* p, void* host_out, size_t capacity);
long mg_ss_build_assets(const mg_ss_params* p, void* host_out, size_t capacity);
long mg_ep_build_assets(const mg_ep_params* p, void* host_out, size_t capacity);

/* numpy SeedSequence(seed) -> PCG64 initial state for `n` seeds (host). out: n records of
 * {state_lo, state_hi, inc_lo, inc_hi} (uint64 each). Bit-identical to
 * numpy.random.PCG64(numpy.random.SeedSequence(seed)).state for 0 <= seed < 2^64. */
int mg_seed_pcg64(const uint64_t* seeds, int n, uint64_t* out);

/* Host-side self test hook for the 128-bit PCG64 / Lemire arithmetic the kernels use (the same
 * __host__ __device__ code compiled for the CPU): starting from `state4` / {has_uint32, uinteger},
 * performs `n_draws` bounded draws integers(lo[i], hi[i]) and, if out_uniform != NULL, one
 * uniform(out_uniform[0], out_uniform[1]) whose result replaces out_uniform[0]. CPU test-suite only. */
int mg_selftest_pcg64_host(uint64_t* state4, uint32_t* has_buf2, const int64_t* lo, const int64_t* hi,
                           int n_draws, int64_t* out, double* out_uniform);

/* ---- environment entry points ------------------------------------------------------------------ */
/* state: device, n * MG_*_STATE_BYTES bytes, opaque to the caller (zero-initialise once = "make").
 *        Endless Mystery Path: the second half of a record (MG_EP_STATE_BYTES = 2048 + 2048) holds the paths of the
 *        env's next episodes, generated ahead of their resets by helper warps of the step / rollout launches; a reset
 *        adopts one only if the env's stream is the stream it was generated from, so results never depend on it.
 * rng4:  device, nullable: [n][4] uint64 from mg_seed_pcg64; where given (and selected by
 *        reset_mask) the env's PCG64 stream is re-seeded before the reset draws, exactly like
 *        `env.reset(seed=s)`; NULL continues each env's stream like `env.reset()`. */
int mg_mm_reset(const mg_mm_params* p, const void* assets, void* state, const uint64_t* rng4,
                const mg_step_io* io, int n, void* stream);
int mg_mm_step(const mg_mm_params* p, const void* assets, void* state, const mg_step_io* io, int n,
               void* stream);

/* `n_steps` consecutive steps in one call: step t uses ios[t] (its own actions / obs / reward / done
 * pointers, e.g. slices of [T][N] rollout buffers) — n_steps kernel launches enqueued back to back. For
 * scripted or pre-sampled actions (the reference's bench.py loop); warps_per_cta is taken from ios[0]. */
int mg_mm_step_many(const mg_mm_params* p, const void* assets, void* state, const mg_step_io* ios, int n_steps,
                    int n, void* stream);

int mg_mm_rollout(const mg_mm_params* p, const void* assets, void* state, const mg_rollout_io* r, int n,
                  void* stream);

int mg_mp_reset(const mg_mp_params* p, const void* assets, void* state, const uint64_t* rng4,
                const mg_step_io* io, int n, void* stream);
int mg_mp_step(const mg_mp_params* p, const void* assets, void* state, const mg_step_io* io, int n,
               void* stream);

int mg_mp_step_many(const mg_mp_params* p, const void* assets, void* state, const mg_step_io* ios, int n_steps,
                    int n, void* stream);

int mg_mp_rollout(const mg_mp_params* p, const void* assets, void* state, const mg_rollout_io* r, int n,
                  void* stream);

int mg_ss_reset(const mg_ss_params* p, const void* assets, void* state, const uint64_t* rng4,
                const mg_step_io* io, int n, void* stream);
int mg_ss_step(const mg_ss_params* p, const void* assets, void* state, const mg_step_io* io, int n,
               void* stream);

int mg_ss_step_many(const mg_ss_params* p, const void* assets, void* state, const mg_step_io* ios, int n_steps,
                    int n, void* stream);

int mg_ss_rollout(const mg_ss_params* p, const void* assets, void* state, const mg_rollout_io* r, int n,
                  void* stream);

int mg_ep_reset(const mg_ep_params* p, const void* assets, void* state, const uint64_t* rng4,
                const mg_step_io* io, int n, void* stream);
int mg_ep_step(const mg_ep_params* p, const void* assets, void* state, const mg_step_io* io, int n,
               void* stream);

int mg_ep_step_many(const mg_ep_params* p, const void* assets, void* state, const mg_step_io* ios, int n_steps,
                    int n, void* stream);

int mg_ep_rollout(const mg_ep_params* p, const void* assets, void* state, const mg_rollout_io* r, int n,
                  void* stream);

/* Option validation without a device: 0 if the reset/step kernels accept `p`, else MG_EINVAL with
 * mg_last_error() naming the offending field. */
int mg_mm_check_params(const mg_mm_params* p);
int mg_mp_check_params(const mg_mp_params* p);
int mg_ss_check_params(const mg_ss_params* p);
int mg_ep_check_params(const mg_ep_params* p);

/* The consumer step right after the env step (what every training loop on the reference does on the host:
 * `obs.astype(float32) / 255`, channels first): obs uint8 [n][84][84][3] -> out bf16 [n][3][84][84] with
 * out[i][c][x][y] = bf16_rn(float(obs[i][x][y][c]) / 255.0f), i.e. obs.permute(0, 3, 1, 2) / 255 bit for bit.
 * Both pointers 16-byte aligned device memory; asynchronous on `stream`. */
int mg_obs_to_policy_bf16(const uint8_t* obs, void* out, int n, void* stream);

/* ---- observations into HOST buffers: lossless delta transport -----------------------------------------
 * The reference returns every observation as a host array (`surfarray.array3d`, mortar_mayhem.py:270,367). A full
 * device->host copy per step is PCIe-bound (21168 B per env-step); consecutive frames of an env differ in a few
 * hundred pixels, so only 16-byte chunks that changed cross the link:
 *   mg_obs_delta_encode  device: compares `obs` [n][84][84][3] with `prev` (device, same shape: the frames the host
 *                        already holds; zero-initialise together with the host array), writes per env a change mask +
 *                        count and the changed chunks straight into `stage` (PINNED host memory addressable by the
 *                        device, 16-byte aligned, mg_obs_delta_stage_bytes(n) bytes) and updates `prev`. full != 0
 *                        sends every chunk (resynchronisation). Asynchronous on `stream`.
 *   mg_obs_delta_decode  host: after the stream reached the encode (event / synchronize), patches the changed chunks
 *                        into `host_obs` [n][84][84][3] with `threads` worker threads; `host_obs` then equals `obs`
 *                        bit for bit. bytes_out (nullable) receives the bytes that crossed the link for this step.
 * Stage layout: n headers of MG_DELTA_HDR_BYTES {uint32 count, uint32 mask[42]}, then n regions of MG_OBS_BYTES
 * holding env e's changed chunks in ascending order. A transport code only: nothing is rendered on the host. */
#define MG_DELTA_HDR_BYTES 176
size_t mg_obs_delta_stage_bytes(int n);
int mg_obs_delta_encode(const uint8_t* obs, uint8_t* prev, int n, void* stage, int full, void* stream);
int mg_obs_delta_decode(const void* stage, uint8_t* host_obs, int n, int threads, uint64_t* bytes_out);
/*   mg_obs_delta_scatter device: the same comparison, but every changed chunk is written to ITS PLACE in `host_obs`
 *                        [n][84][84][3] (PINNED host memory addressable by the device, 16-byte aligned) and in `prev`:
 *                        once the stream has reached it, `host_obs` equals `obs` bit for bit and the host has nothing
 *                        left to do. `prev` must mirror what THIS host array holds (one prev per host array when several
 *                        are used in turn). counter (nullable, DEVICE memory, 8-byte aligned): payload bytes written are
 *                        added to it. Asynchronous on `stream`. */
int mg_obs_delta_scatter(const uint8_t* obs, uint8_t* prev, int n, uint8_t* host_obs, int full, uint64_t* counter,
                         void* stream);

/* Library / device introspection. */
int mg_abi_version(void);
int mg_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin);
const char* mg_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* MEMGYM_B200_H */
