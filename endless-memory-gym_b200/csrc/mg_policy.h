// mg_policy.h — per-env-type glue between the batched I/O (mg_step_io), the env logic and the
// rasteriser. Shared by the CUDA kernel (mg_kernels.cu) and the CPU test harness (tests/hostsim),
// so the logic the GPU runs can be checked against the oracle without a GPU.
#pragma once
#include "../../include/memgym_b200.h"
#include "mg_common.h"
#include "mg_frame.h"
#include "mg_mm_logic.h"
#include "mg_mp_logic.h"
#include "mg_ep_logic.h"
#include "mg_ss_logic.h"
#include "mg_debug.h"

#if defined(__CUDACC__)
#define MG_HDN __host__ __device__
#else
#define MG_HDN inline
#endif

#define MG_MODE_RESET 0
#define MG_MODE_STEP 1

struct MgKIO {
    mg_step_io io;
    const uint64_t* rng4;
    int32_t mode;
    int32_t n;
    int32_t stage_cap; // kernel: shared-memory bytes available for staging the asset header + sprite lists
    int32_t act_dims; // int64 entries per env in io.actions (2 = MultiDiscrete([3,3]), 1 = Discrete(4))
    // rollout (n_steps > 1): step t reads actions + t * act_stride and writes obs slot t % ring_n,
    // reward/done/info rows t (strides in elements; info_stride 0 = one shared info matrix)
    int32_t n_steps, ring_n;
    int64_t act_stride, obs_stride, rew_stride, done_stride, info_stride;
    uint32_t* sched;  // rollout: global work counter (mg_rollout_io.sched) or nullptr = CTA-local scheduling
    int32_t n_helpers;  // helper warps behind the env teams of every CTA (policies with kHelperWarps > 0)
    int32_t n_lane_helpers;  // ... of which the first n_lane_helpers generate paths ahead (EPPolicy)
};

struct MgFrameJob {      // what the logic lane hands to the warp: up to two frames for this env-step
    int32_t nframes;
    int32_t to_final[2];  // frame goes to final_obs instead of obs
};

// One env, one call: reset (mode RESET) or step (+ auto-reset) — the same control flow for every env type.
// `E` supplies reset_env / step_env. a0/a1 are this env's (prefetched) action components.
template <class E>
MG_HD void mg_env_logic(const typename E::K& k, const MgKIO& q, const MgAssetView& a, int env, int t,
                         typename E::State& s, int a0, int a1, MgFrameJob& job, typename E::Frame* fr) {
    double* const reward_out = q.io.reward + (size_t)t * q.rew_stride + env;
    uint8_t* const done_out = q.io.done + (size_t)t * q.done_stride + env;
    double* const info_out = q.io.info + (size_t)t * q.info_stride + MG_INFO_COLS * (size_t)env;
    double* gt = q.io.ground_truth ? q.io.ground_truth + MG_GT_COLS * (size_t)env : nullptr;
    job.nframes = 0;
    // ONE call site for the reset (a fresh episode or the auto-reset of a finished one): the reset is inlined, so the
    // option block `k` is read through constant-bank operands instead of a pointer into parameter space
    bool reset_now = false;
    int at = 0;
    if (q.mode == MG_MODE_RESET) {
        if (q.io.reset_mask && !q.io.reset_mask[env]) return;
        if (q.rng4) {  // env.reset(seed=s): a fresh PCG64 stream
            const uint64_t* r = q.rng4 + 4 * (size_t)env;
            s.rng.state_lo = r[0];
            s.rng.state_hi = r[1];
            s.rng.inc_lo = r[2];
            s.rng.inc_hi = r[3];
            s.rng.buf = 0;
            s.rng.has_buf = 0;
        }
        reset_now = true;
    } else if (!s.live) {  // finished (or never reset): the step is ignored
        *reward_out = 0.0;
        *done_out = 1;
        return;
    } else {
        double reward;
        bool done = E::step_env(k, a, s, a0, a1, reward, info_out, fr[0], gt);
        *reward_out = reward;
        *done_out = done ? 1 : 0;
        job.nframes = 1;
        job.to_final[0] = 0;
        if (done && (q.io.flags & MG_FLAG_AUTORESET)) {
            if (q.io.final_obs) at = 1;  // terminal frame -> final_obs, first frame of the next episode -> obs
            reset_now = true;
        }
    }
    if (MG_UNLIKELY(reset_now)) {
        E::reset_env(k, a, s, fr[at], gt);
        job.to_final[0] = at;
        job.to_final[at] = 0;
        job.nframes = at + 1;
    }
}

// per-warp / per-CTA scratch beyond the frame buffer and the record slot (Searing Spotlights' base buffer, tables)
struct MgScratch {
    uint8_t* warp;
    const uint8_t* cta;
};

// The same control flow for env types whose reset is WARP-COOPERATIVE (Mystery Path family: the path generator
// runs on all lanes). Called by every lane; lane 0 runs the scalar parts and hands its decisions to the warp
// through the per-warp scratch (MgGenMem::action / flag). `E` supplies reset_coop / step_coop.
template <class E, class Ctx>
MG_HD void mg_env_logic_coop(Ctx& c, const typename E::K& k, const MgKIO& q, const MgAssetView& a, int env, int t,
                             typename E::State& s, int a0, int a1, MgFrameJob& job, typename E::Frame* fr,
                             const MgScratch& sc) {
    MgGenMem& m = *reinterpret_cast<MgGenMem*>(sc.warp);
    double* const reward_out = q.io.reward + (size_t)t * q.rew_stride + env;
    uint8_t* const done_out = q.io.done + (size_t)t * q.done_stride + env;
    double* const info_out = q.io.info + (size_t)t * q.info_stride + MG_INFO_COLS * (size_t)env;
    double* gt = q.io.ground_truth ? q.io.ground_truth + MG_GT_COLS * (size_t)env : nullptr;
    if (c.lane == 0) {
        int action = 0;
        job.nframes = 0;
        if (q.mode == MG_MODE_RESET) {
            if (!(q.io.reset_mask && !q.io.reset_mask[env])) {
                if (q.rng4) {  // env.reset(seed=s): a fresh PCG64 stream
                    const uint64_t* r = q.rng4 + 4 * (size_t)env;
                    s.rng.state_lo = r[0];
                    s.rng.state_hi = r[1];
                    s.rng.inc_lo = r[2];
                    s.rng.inc_hi = r[3];
                    s.rng.buf = 0;
                    s.rng.has_buf = 0;
                }
                action = 1;
            }
        } else if (!s.live) {  // finished (or never reset): the step is ignored
            *reward_out = 0.0;
            *done_out = 1;
        } else {
            action = 2;
        }
        m.action = action;
    }
    c.sync();
    const int action = m.action;
    if (action == 0) return;
    bool reset_now = action == 1;
    int at = 0;
    if (action == 2) {
        const bool done = E::step_coop(c, k, a, s, a0, a1, reward_out, done_out, info_out, fr[0], gt, sc);
        if (done && (q.io.flags & MG_FLAG_AUTORESET)) {
            if (q.io.final_obs) at = 1;  // terminal frame -> final_obs, first frame of the next episode -> obs
            reset_now = true;
        }
    }
    if (MG_UNLIKELY(reset_now)) E::reset_coop(c, k, a, s, fr[at], gt, sc);  // the one (inlined) call site
    if (c.lane == 0) {
        job.to_final[0] = at;
        job.to_final[at] = 0;
        job.nframes = at + 1;
    }
}

struct MMPolicy {
    static constexpr bool kPipelined = false;
    static constexpr int kHelperWarps = 0, kRolloutHelpers = 0;
    static constexpr int kWarpExtra = 0, kCtaExtra = 0;
    static constexpr int kStageReserve = 8 * 1024;  // shared memory kept free for the staged asset prefix (7.4 KB by default)
    static constexpr int kTeamWarps = 1, kMaxThreads = 320;
    static constexpr int kFrameBuf = 21248;      // per-team TMA frame buffer (MG_FRAME_BYTES rounded up to 128)
    static constexpr bool kDirectStore = false;  // frames leave through cp.async.bulk from the frame buffer
    static MG_HD void cta_init(uint8_t*, int, int) {}
    static MG_HD void warp_init(uint8_t*, int) {}
    typedef MgMMK K;
    typedef MgMMState State;
    typedef MgMMFrame Frame;
    static MG_HD bool restores_clean(const K&) { return true; }
    static MG_HD void reset_env(const K& k, const MgAssetView&, State& s, Frame& f, double* gt) {
        mg_mm_reset_env(k, s, f, gt);
    }
    static MG_HD bool step_env(const K& k, const MgAssetView&, State& s, int a0, int a1, double& reward, double* info,
                               Frame& f, double* gt) {
        return mg_mm_step_env(k, s, a0, a1, reward, info, f, gt);
    }
    template <class Ctx>
    static MG_HD void logic(Ctx& c, const K& k, const MgKIO& q, const MgAssetView& a, int env, int t, State& s, int a0,
                            int a1, MgFrameJob& job, Frame* fr, const MgScratch&) {
        if (c.lane == 0) mg_env_logic<MMPolicy>(k, q, a, env, t, s, a0, a1, job, fr);
        if (mg_mm_is_task_b(k.p.variant) && q.io.vector_obs) {
            c.sync();
            // a fresh episode (reset or auto-reset in this call): publish its command vector
            if (s.live && s.ep_len == 0) mg_mm_write_vector(c, s, q.io.vector_obs + (size_t)MG_MM_B_VECTOR * env);
        }
    }
    template <class Ctx>
    static MG_HD void compose(Ctx& c, uint8_t* buf, const MgAssetView& a, const K& k, const Frame& f, const MgScratch&) {
        mg_mm_compose(c, buf, a, f, k.origin, a.hdr->agent_dim >> 1, a.hdr->icon_x, a.hdr->icon_y);
    }
    template <class Ctx>
    static MG_HD void restore(Ctx& c, uint8_t* buf, const MgAssetView& a, const K& k, const Frame& f, const MgScratch&) {
        mg_mm_restore(c, buf, a, f, k.origin, k.size, a.hdr->agent_dim >> 1, a.hdr->icon_x, a.hdr->icon_y);
    }
    // the reference's B variants have no debug render (mortar_mayhem_b.py has no _build_debug_surface)
    template <class Ctx>
    static MG_HD void emit_debug(Ctx& c, uint8_t* out, const MgAssetView& a, const K& k, const Frame& f, const State& s,
                                 const MgScratch&) {
        if (!mg_mm_is_task_b(k.p.variant)) mg_mm_debug(c, out, a, k, f, s);
    }
};

struct MPPolicy {
    static constexpr bool kPipelined = false;
    static constexpr int kHelperWarps = 0, kRolloutHelpers = 0;
    static constexpr int kWarpExtra = (int)((sizeof(MgGenMem) + 127) / 128 * 128);
    static constexpr int kCtaExtra = (int)((sizeof(MgGenCta) + 127) / 128 * 128);
    static constexpr int kStageReserve = 8 * 1024;
    static constexpr int kTeamWarps = 1, kMaxThreads = 320;
    static constexpr int kFrameBuf = 21248;
    static constexpr bool kDirectStore = false;
    static MG_HD void cta_init(uint8_t* cta, int tid, int nthreads) {
        mg_gen_cta_init(reinterpret_cast<MgGenCta*>(cta), tid, nthreads);
    }
    static MG_HD void warp_init(uint8_t*, int) {}
    typedef MgMPK K;
    typedef MgMPState State;
    typedef MgMPDraw Frame;
    static MG_HD bool restores_clean(const K&) { return true; }
    template <class Ctx>
    static MG_HD void reset_coop(Ctx& c, const K& k, const MgAssetView&, State& s, Frame& f, double*, const MgScratch& sc) {
        mg_mp_reset_coop(c, k, s, f, *reinterpret_cast<MgGenMem*>(sc.warp), *reinterpret_cast<const MgGenCta*>(sc.cta));
    }
    template <class Ctx>
    static MG_HD bool step_coop(Ctx& c, const K& k, const MgAssetView&, State& s, int a0, int a1, double* reward_out,
                                uint8_t* done_out, double* info, Frame& f, double*, const MgScratch& sc) {
        MgGenMem& m = *reinterpret_cast<MgGenMem*>(sc.warp);
        if (c.lane == 0) {
            double reward;
            const bool done = mg_mp_step_env(k, s, a0, a1, reward, info, f);
            *reward_out = reward;
            *done_out = done ? 1 : 0;
            m.flag = done ? 1 : 0;
        }
        c.sync();
        return m.flag != 0;
    }
    template <class Ctx>
    static MG_HD void logic(Ctx& c, const K& k, const MgKIO& q, const MgAssetView& a, int env, int t, State& s, int a0,
                            int a1, MgFrameJob& job, Frame* fr, const MgScratch& sc) {
        mg_env_logic_coop<MPPolicy>(c, k, q, a, env, t, s, a0, a1, job, fr, sc);
    }
    template <class Ctx>
    static MG_HD void compose(Ctx& c, uint8_t* buf, const MgAssetView& a, const K&, const Frame& f, const MgScratch&) {
        mg_mp_compose(c, buf, a, f, a.hdr->agent_dim >> 1);
    }
    template <class Ctx>
    static MG_HD void restore(Ctx& c, uint8_t* buf, const MgAssetView& a, const K&, const Frame& f, const MgScratch&) {
        mg_mp_restore(c, buf, a, f, a.hdr->agent_dim >> 1);
    }
    template <class Ctx>
    static MG_HD void emit_debug(Ctx& c, uint8_t* out, const MgAssetView& a, const K&, const Frame& f, const State& s,
                                 const MgScratch&) {
        mg_mp_debug(c, out, a, f, s);
    }
};

struct EPPolicy {
    static constexpr bool kPipelined = false;
    static constexpr int kRolloutHelpers = 2;  // warps that generate paths ahead, one env per lane (every launch)
    static constexpr int kOffloadHelpers = 3;  // single-step launches: warps that finish the paths of auto-resets
    static constexpr int kHelperWarps = kRolloutHelpers + kOffloadHelpers;
    static constexpr int kRecordBytes = MG_EP_STATE_BYTES;  // record + the MgEPSpecRec behind it
    static constexpr int kWarpExtra = (int)((sizeof(MgEPWarpMem) + 63) / 64 * 64);
    static constexpr int kGenCtaBytes = (int)((sizeof(MgGenCta) + 127) / 128 * 128);
    static constexpr int kLaneBytes = kRolloutHelpers * MG_EP_LANE_WARP_BYTES;
    static constexpr int kCtaExtra = kGenCtaBytes + (kLaneBytes + kOffloadHelpers * MG_EP_HELPER_BYTES + 127) / 128 * 128;
    static constexpr int kStageReserve = 4608;  // (the default assets stage 4480 bytes)
    static constexpr int kTeamWarps = 1, kMaxThreads = 512;
    static constexpr int kFrameBuf = 21248;
    static constexpr bool kDirectStore = false;
    static MG_HD void cta_init(uint8_t* cta, int tid, int nthreads) {
        mg_gen_cta_init(reinterpret_cast<MgGenCta*>(cta), tid, nthreads);
    }
    static MG_HD void warp_init(uint8_t* warp, int lane) {
        if (lane == 0) {
            MgEPWarpMem* wm = reinterpret_cast<MgEPWarpMem*>(warp);
            wm->ahead = 0;
            wm->offload = 0;
            wm->cur_env = -1;
            wm->post_env = -1;
            wm->state = nullptr;
            wm->spec.state = MG_SPEC_EMPTY;
        }
    }
    // helper warps of this launch: n_lane generate ahead, n_offload finish the paths of auto-resets (single steps)
    static MG_HD void set_helpers(uint8_t* warp, int lane, int n_lane, int n_offload, uint8_t* state) {
        if (lane == 0) {
            reinterpret_cast<MgEPWarpMem*>(warp)->ahead = n_lane > 0 ? 1 : 0;
            reinterpret_cast<MgEPWarpMem*>(warp)->offload = n_offload > 0 ? 1 : 0;
            reinterpret_cast<MgEPWarpMem*>(warp)->state = state;
        }
    }
    template <class Ctx>
    static MG_HD void after_writeback(Ctx& c, uint8_t* warp) {
        mg_ep_post_offload(c, *reinterpret_cast<MgEPWarpMem*>(warp));
    }
    template <class Ctx>
    static MG_HD void helper_loop_step(Ctx& c, uint8_t* cextra, uint8_t* wextra0, int nteams, int hidx,
                                       const uint32_t* teams_done, uint8_t* state, double* ground_truth) {
        uint8_t* mine = cextra + kGenCtaBytes + kLaneBytes + (size_t)hidx * MG_EP_HELPER_BYTES;
        MgGenMem m{0, 0};
        mg_ep_helper_offload_loop(c, mine, wextra0, kWarpExtra, nteams, hidx, teams_done, m,
                                  *reinterpret_cast<const MgGenCta*>(cextra), state, ground_truth);
    }
    // rollout helper warp `hidx` of `nhelp` in CTA `cta` of `nctas`: one env per lane
    template <class Ctx>
    static MG_HD void helper_loop(Ctx& c, uint8_t* cextra, int nteams, int hidx, int nhelp, int cta, int nctas,
                                  const uint32_t* teams_done, int stop_at, uint32_t min_ns, uint8_t* state, int n) {
        uint8_t* mine = cextra + kGenCtaBytes + (size_t)hidx * MG_EP_LANE_WARP_BYTES;
        MgLaneMem m;
        m.info = reinterpret_cast<uint16_t*>(mine) + c.lane;
        m.list = mine + MG_MP_NODES * 32 * 2 + c.lane;
        m.stride = 32;
        m.stop = teams_done;
        m.stop_at = (uint32_t)stop_at;  // rollouts: every team has left; single steps: the first has (parking overlaps the rest)
        m.budget = -1;
        m.not_before = min_ns ? c.now_ns() + min_ns : 0;
        mg_ep_lane_helper_loop(c, state, n, (cta * nhelp + hidx) * 32 + c.lane, nctas * nhelp * 32, teams_done, stop_at,
                               reinterpret_cast<const MgGenCta*>(cextra)->sqrt_d2, m);
    }
    typedef MgEPK K;
    typedef MgEPState State;
    typedef MgEPFrame Frame;
    // with show_background every frame repaints the whole buffer instead of un-patching it
    static MG_HD bool restores_clean(const K& k) { return !k.p.show_background; }
    template <class Ctx>
    static MG_HD void reset_coop(Ctx& c, const K& k, const MgAssetView&, State& s, Frame& f, double* gt, const MgScratch& sc) {
        mg_ep_reset_coop(c, k, s, f, gt, *reinterpret_cast<MgEPWarpMem*>(sc.warp), *reinterpret_cast<const MgGenCta*>(sc.cta));
    }
    // the step may have to extend the path (add_path_segment, endless_mystery_path.py:300-302): the move runs on
    // one lane, the segment on the warp, the rest of the step on one lane again
    template <class Ctx>
    static MG_HD bool step_coop(Ctx& c, const K& k, const MgAssetView&, State& s, int a0, int, double* reward_out,
                                uint8_t* done_out, double* info, Frame& f, double* gt, const MgScratch& sc) {
        MgGenMem& m = *reinterpret_cast<MgGenMem*>(sc.warp);
        if (c.lane == 0) m.flag = mg_ep_step_pre(k, s, a0) ? 1 : 0;
        c.sync();
        const bool need = m.flag != 0;
        c.sync();  // every lane has read the flag before lane 0 reuses the word for "done"
        bool capped = false;
        if (need) {
            capped = !mg_ep_add_segment_coop(c, s, m, *reinterpret_cast<const MgGenCta*>(sc.cta));
            // the stream has moved: what was generated ahead for the next reset is stale, ask again
            MgEPWarpMem& wm = *reinterpret_cast<MgEPWarpMem*>(sc.warp);
            if (wm.ahead) mg_ep_chain_restart(c, s, wm);
        }
        if (c.lane == 0) {
            double reward;
            const bool done = mg_ep_step_post(k, s, capped, reward, info, f, gt);
            *reward_out = reward;
            *done_out = done ? 1 : 0;
            m.flag = done ? 1 : 0;
        }
        c.sync();
        return m.flag != 0;
    }
    template <class Ctx>
    static MG_HD void logic(Ctx& c, const K& k, const MgKIO& q, const MgAssetView& a, int env, int t, State& s, int a0,
                            int a1, MgFrameJob& job, Frame* fr, const MgScratch& sc) {
        MgEPWarpMem& wm = *reinterpret_cast<MgEPWarpMem*>(sc.warp);
        // (a record whose chain is missing or stale — first launch, re-seeded by a reset launch — is left to its next
        // auto-reset, which then generates inline and starts the chain: looking at the chain every time a team takes an
        // env over cost the team 1-5 us of dependent global loads per visit)
        if (c.lane == 0) wm.cur_env = env;  // (read by lane 0 only)
        mg_env_logic_coop<EPPolicy>(c, k, q, a, env, t, s, a0, a1, job, fr, sc);
    }
    template <class Ctx>
    static MG_HD void compose(Ctx& c, uint8_t* buf, const MgAssetView& a, const K&, const Frame& f, const MgScratch&) {
        mg_ep_compose(c, buf, a, f);
    }
    template <class Ctx>
    static MG_HD void restore(Ctx& c, uint8_t* buf, const MgAssetView& a, const K&, const Frame& f, const MgScratch&) {
        mg_ep_restore(c, buf, a, f);
    }
    template <class Ctx>
    static MG_HD void emit_debug(Ctx& c, uint8_t* out, const MgAssetView& a, const K& k, const Frame& f, const State& s,
                                 const MgScratch&) {
        mg_ep_debug(c, out, a, k, f, s);
    }
};

template <int TW, bool kEndless>
struct SSPolicyT {
    static constexpr int kHelperWarps = 0, kRolloutHelpers = 0;
    static constexpr int kWarpExtra = (int)((sizeof(MgSSWarpMem) + 127) / 128 * 128);
    static constexpr int kCtaExtra = (int)((sizeof(MgSSCtaTables) + 127) / 128 * 128);
    // instruction / latency-bound frames: TW warps share one env (more warps per base buffer hide more latency)
    static constexpr int kStageReserve = 16 * 1024;
    static constexpr int kTeamWarps = TW, kMaxThreads = TW <= 2 ? 512 : 256 * TW;
    static constexpr int kFrameBuf = 0;         // no staging buffer: the frame is a masked copy of the base buffer
    static constexpr bool kDirectStore = true;  // ... written straight to the observation with 128-bit stores
    static_assert(sizeof(MgSSStepTmp) <= sizeof(((MgSSWarpMem*)0)->step_tmp), "MgSSWarpMem::step_tmp too small");
    static MG_HD void cta_init(uint8_t* cta, int tid, int nthreads) {
        mg_ss_tables_init(reinterpret_cast<MgSSCtaTables*>(cta), tid, nthreads);
    }
    static MG_HD void warp_init(uint8_t* warp, int lane) {
        if (lane == 0) {
            MgSSWarpMem* wm = reinterpret_cast<MgSSWarpMem*>(warp);
            wm->desc.valid = 0;
            wm->desc.lut_alpha = -1;
            wm->desc.env = -1;
        }
    }
    typedef MgSSK K;
    typedef MgSSState State;
    typedef MgSSFrame Frame;
    static MG_HD bool restores_clean(const K&) { return false; }
    static MG_HD const double* trig(const MgAssetView& a) { return (const double*)(a.base + a.hdr->trig_off); }

    // Same control flow as mg_env_logic, but the step is split A (one lane) / B (one lane per spotlight) /
    // C (one lane): spotlight motion + hit tests are the only per-step work that grows with the scene.
    template <class Ctx>
    static MG_HD void logic(Ctx& c, const K& k, const MgKIO& q, const MgAssetView& a, int env, int t, State& s, int a0,
                            int a1, MgFrameJob& job, Frame* fr, const MgScratch& sc) {
        MgSSStepTmp& tmp = *reinterpret_cast<MgSSStepTmp*>(reinterpret_cast<MgSSWarpMem*>(sc.warp)->step_tmp);
        double* const reward_out = q.io.reward + (size_t)t * q.rew_stride + env;
        uint8_t* const done_out = q.io.done + (size_t)t * q.done_stride + env;
        double* const info_out = q.io.info + (size_t)t * q.info_stride + MG_INFO_COLS * (size_t)env;
        double* const gt = q.io.ground_truth ? q.io.ground_truth + MG_GT_COLS * (size_t)env : nullptr;
        // the sampler's row masks borrow the base buffer: a reset rebuilds the base anyway
        MgSSWarpMem* wm = reinterpret_cast<MgSSWarpMem*>(sc.warp);
        MgSSSampler& g = *reinterpret_cast<MgSSSampler*>(wm->base);
        MgSSResetTmp& rt = *reinterpret_cast<MgSSResetTmp*>(wm->base + sizeof(MgSSSampler));
        if (c.lane == 0) {
            job.nframes = 0;
            tmp.action = 0;
            tmp.post = 0;
            if (MG_UNLIKELY(wm->desc.env != env)) {  // the team moved on to another env: its base buffer shows a foreign scene
                wm->desc.valid = 0;
                wm->desc.env = env;
            }
            if (q.mode == MG_MODE_RESET) {
                if (!(q.io.reset_mask && !q.io.reset_mask[env])) {
                    if (q.rng4) {  // env.reset(seed=s): a fresh PCG64 stream
                        const uint64_t* r = q.rng4 + 4 * (size_t)env;
                        s.rng.state_lo = r[0];
                        s.rng.state_hi = r[1];
                        s.rng.inc_lo = r[2];
                        s.rng.inc_hi = r[3];
                        s.rng.buf = 0;
                        s.rng.has_buf = 0;
                    }
                    tmp.action = 1;
                }
            } else if (!s.live) {  // finished (or never reset): the step is ignored
                *reward_out = 0.0;
                *done_out = 1;
            } else {
                tmp.action = 2;
                mg_ss_step_a(k, s, a0, a1, tmp);
            }
        }
        c.sync();
        c.stamp(12);
        if (MG_LIKELY(tmp.action == 2)) {
            mg_ss_step_b(c, k, trig(a), s, tmp, fr[0]);
            c.sync();
            c.stamp(13);
            MgESSCarry carry;  // lane 0's registers carry C1's results to C2
            if constexpr (kEndless) {
                if (c.lane == 0 && mg_ess_step_c1(k, s, tmp, carry)) {
                    tmp.action = 4;  // coin collected: the team respawns it on the (now invalid) base buffer
                    wm->desc.valid = 0;
                }
                c.sync();
                if (MG_UNLIKELY(tmp.action == 4)) mg_ess_spawn_coin(c, k.clamp_off, s, g, rt, true);
            }
            if (c.lane == 0) {
                double reward;
                bool done;
                if constexpr (kEndless)
                    done = mg_ess_step_c2(k, s, tmp, carry, reward, info_out, fr[0], gt);
                else
                    done = mg_ss_step_c(k, s, tmp, reward, info_out, fr[0]);
                *reward_out = reward;
                *done_out = done ? 1 : 0;
                job.nframes = 1;
                job.to_final[0] = 0;
                // (a word of its own: the other warp of the team may still be reading tmp.action above)
                if (done && (q.io.flags & MG_FLAG_AUTORESET)) {
                    tmp.post = 3;  // auto-reset: every lane joins the cooperative reset below
                    if (q.io.final_obs) job.to_final[0] = 1;  // terminal frame -> final_obs, next episode's first -> obs
                }
            }
            c.sync();
            c.stamp(14);
        }
        // ONE (inlined) call site for the reset: a fresh episode (1) or the auto-reset of a finished one (3)
        if (MG_UNLIKELY(tmp.action == 1 || tmp.post == 3)) {
            // auto-reset: the terminal frame (if requested) must leave before the base buffer is borrowed by the
            // sampler: it is emitted here, out of order with respect to the generic frame loop
            const int at = (tmp.post == 3 && q.io.final_obs) ? 1 : 0;
            if (at) {
                emit_cold(c, q.io.final_obs + (size_t)env * MG_FRAME_BYTES, a, k, fr[0], sc);
                c.sync();
            }
            if (c.lane == 0) wm->desc.valid = 0;
            if constexpr (kEndless)
                mg_ess_reset_env(c, k, s, fr[at], g, rt, gt);
            else
                mg_ss_reset_env(c, k, s, fr[at], g, rt);
            if (c.lane == 0) {
                // a terminal frame has already been written: hand only the reset frame to the frame loop
                if (at) fr[0] = fr[1];
                job.to_final[0] = 0;
                job.nframes = 1;
            }
            c.sync();
        }
    }

    // ---- pipelined rollouts (mg_kernels.cu): the first warp of the team computes step t + 1 while the second one
    // patches the base buffer for frame t; both then build the lit bitmap and copy frame t out. Resets (and the
    // endless variant, whose coin respawn borrows the base buffer mid-step) stay on the whole team.
    static constexpr bool kPipelined = TW == 2;
    template <class Ctx>
    static MG_HD void pipe_begin(Ctx& c, int env, const MgScratch& sc) {
        MgSSWarpMem* wm = reinterpret_cast<MgSSWarpMem*>(sc.warp);
        if (c.lane == 0 && wm->desc.env != env) {  // the team moved on to another env: a foreign scene in the base buffer
            wm->desc.valid = 0;
            wm->desc.env = env;
        }
    }
    // step(): A / B / C on ONE warp (`c` = its 32 lanes); the frame goes to `f`. What is left for the whole team is
    // announced in tmp.post: 3 = auto-reset, 4 = (endless) the coin was collected — the respawn borrows the base
    // buffer, so the step stops after C1 and pipe_deferred finishes it.
    template <class Ctx>
    static MG_HD void pipe_step(Ctx& c, const K& k, const MgKIO& q, const MgAssetView& a, int env, int t, State& s, int a0,
                                int a1, Frame& f, const MgScratch& sc) {
        MgSSStepTmp& tmp = *reinterpret_cast<MgSSStepTmp*>(reinterpret_cast<MgSSWarpMem*>(sc.warp)->step_tmp);
        c.sync();  // every lane is done with the previous step's hand-over words
        if (c.lane == 0) {
            tmp.action = 2;
            tmp.post = 0;
            mg_ss_step_a(k, s, a0, a1, tmp);
        }
        c.sync();
        mg_ss_step_b(c, k, trig(a), s, tmp, f);
        c.sync();
        if (c.lane == 0) {
            if constexpr (kEndless) {
                if (mg_ess_step_c1(k, s, tmp, tmp.carry)) {
                    tmp.post = 4;
                    return;
                }
            }
            pipe_finish(k, q, env, t, s, tmp, f);
        }
    }
    // C / C2 on one lane: rewards, done, episode statistics, the frame description
    static MG_HD void pipe_finish(const K& k, const MgKIO& q, int env, int t, State& s, MgSSStepTmp& tmp, Frame& f) {
        double* const info_out = q.io.info + (size_t)t * q.info_stride + MG_INFO_COLS * (size_t)env;
        double reward;
        bool done;
        if constexpr (kEndless)
            done = mg_ess_step_c2(k, s, tmp, tmp.carry, reward, info_out, f,
                                  q.io.ground_truth ? q.io.ground_truth + MG_GT_COLS * (size_t)env : nullptr);
        else
            done = mg_ss_step_c(k, s, tmp, reward, info_out, f);
        q.io.reward[(size_t)t * q.rew_stride + env] = reward;
        q.io.done[(size_t)t * q.done_stride + env] = done ? 1 : 0;
        tmp.post = done ? 3 : 0;  // a rollout always auto-resets
    }
    static MG_HD bool pipe_needs_deferred(const MgScratch& sc) {
        return reinterpret_cast<const MgSSStepTmp*>(reinterpret_cast<const MgSSWarpMem*>(sc.warp)->step_tmp)->post != 0;
    }
    // what a step left for the whole team (between two frames, when nobody touches the base buffer): the coin respawn
    // of the endless variant + the rest of that step, and the auto-reset. `f` holds the step's frame and, after an
    // auto-reset, receives the first frame of the next episode.
    template <class Ctx>
    static MG_HD void pipe_deferred(Ctx& c, const K& k, const MgKIO& q, const MgAssetView& a, int env, int t, State& s,
                                    Frame& f, const MgScratch& sc) {
        MgSSWarpMem* wm = reinterpret_cast<MgSSWarpMem*>(sc.warp);
        MgSSStepTmp& tmp = *reinterpret_cast<MgSSStepTmp*>(wm->step_tmp);
        MgSSSampler& g = *reinterpret_cast<MgSSSampler*>(wm->base);
        MgSSResetTmp& rt = *reinterpret_cast<MgSSResetTmp*>(wm->base + sizeof(MgSSSampler));
        if constexpr (kEndless) {
            if (tmp.post == 4) {
                if (c.lane == 0) wm->desc.valid = 0;
                mg_ess_spawn_coin(c, k.clamp_off, s, g, rt, true);
                if (c.lane == 0) pipe_finish(k, q, env, t, s, tmp, f);
                c.sync();
            }
        }
        const bool reset = tmp.post == 3;
        c.sync();  // every thread has read the hand-over word: the next step's logic may rewrite it
        if (reset) {
            if (q.io.final_obs) {
                emit_cold(c, q.io.final_obs + (size_t)env * MG_FRAME_BYTES, a, k, f, sc);
                c.sync();
            }
            if (c.lane == 0) wm->desc.valid = 0;
            if constexpr (kEndless)
                mg_ess_reset_env(c, k, s, f, g, rt, q.io.ground_truth ? q.io.ground_truth + MG_GT_COLS * (size_t)env : nullptr);
            else
                mg_ss_reset_env(c, k, s, f, g, rt);
            c.sync();
        }
    }
    template <class Ctx>
    static MG_HD void emit_base(Ctx& c, const MgAssetView& a, const K& k, const Frame& f, const MgScratch& sc) {
        mg_ss_emit_base<kEndless>(c, a, draw_k(k, f), f, reinterpret_cast<MgSSWarpMem*>(sc.warp));
    }
    // variant of the split for the pipelined loop: the lit bitmap joins the base patch on the raster warp
    template <class Ctx>
    static MG_HD void emit_base_lit(Ctx& c, const MgAssetView& a, const K& k, const Frame& f, const MgScratch& sc) {
        const MgSSDrawK d = draw_k(k, f);
        mg_ss_emit_base<kEndless>(c, a, d, f, reinterpret_cast<MgSSWarpMem*>(sc.warp));
        mg_ss_emit_lit(c, a, d, f, reinterpret_cast<MgSSWarpMem*>(sc.warp), reinterpret_cast<const MgSSCtaTables*>(sc.cta));
    }
    template <class Ctx>
    static MG_HD void emit_copy(Ctx& c, uint8_t* dst, const MgAssetView& a, const K& k, const Frame& f, const MgScratch& sc,
                                bool stream) {
        mg_ss_emit_copy(c, dst, a, draw_k(k, f), f, reinterpret_cast<MgSSWarpMem*>(sc.warp),
                        reinterpret_cast<const MgSSCtaTables*>(sc.cta), stream);
    }
    template <class Ctx>
    static MG_HD void emit_rest(Ctx& c, uint8_t* dst, const MgAssetView& a, const K& k, const Frame& f, const MgScratch& sc,
                                bool stream) {
        mg_ss_emit_rest(c, dst, a, draw_k(k, f), f, reinterpret_cast<MgSSWarpMem*>(sc.warp),
                        reinterpret_cast<const MgSSCtaTables*>(sc.cta), stream);
    }

    static MG_HD MgSSDrawK draw_k(const K& k, const Frame& f) {
        MgSSDrawK d;
        const int alt = f.coin_above ? 1 : 0;
        d.n_layers = k.n_layers;
        d.layers = alt ? k.layers_alt : k.layers;
        d.spot_li = k.spot_li[alt];
        d.bar_li = k.bar_li[alt];
        d.cb_x = k.cb_x;
        d.cb_w = k.cb_w;
        d.show_last_action = k.p.show_last_action;
        d.has_coin_layer = 1;
        d.border = k.p.black_background;
        return d;
    }
    // the hot copy: inlined once into the kernel's frame loop
    template <class Ctx>
    static MG_HD void emit(Ctx& c, uint8_t* dst, const MgAssetView& a, const K& k, const Frame& f, const MgScratch& sc,
                           bool stream) {
        mg_ss_emit<kEndless>(c, dst, a, draw_k(k, f), f, reinterpret_cast<MgSSWarpMem*>(sc.warp),
                   reinterpret_cast<const MgSSCtaTables*>(sc.cta), stream);
    }
    // terminal frames (one per episode): an out-of-line copy keeps the hot loop's code small
    template <class Ctx>
    static MG_HD_NOINLINE void emit_cold(Ctx c, uint8_t* dst, const MgAssetView& a, const K& k, const Frame& f,
                                         const MgScratch& sc) {
        mg_ss_emit<kEndless>(c, dst, a, draw_k(k, f), f, reinterpret_cast<MgSSWarpMem*>(sc.warp),
                   reinterpret_cast<const MgSSCtaTables*>(sc.cta), false);
    }
    template <class Ctx>
    static MG_HD void emit_debug(Ctx& c, uint8_t* out, const MgAssetView& a, const K& k, const Frame& f, const State& s,
                                 const MgScratch&) {
        mg_ss_debug(c, out, a, draw_k(k, f), f, s, kEndless);
    }
    // unused with kDirectStore
    template <class Ctx>
    static MG_HD void compose(Ctx&, uint8_t*, const MgAssetView&, const K&, const Frame&, const MgScratch&) {}
    template <class Ctx>
    static MG_HD void restore(Ctx&, uint8_t*, const MgAssetView&, const K&, const Frame&, const MgScratch&) {}
};
typedef SSPolicyT<2, false> SSPolicy;
typedef SSPolicyT<2, true> ESSPolicy;
