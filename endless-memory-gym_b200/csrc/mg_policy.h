// mg_policy.h — per-env-type glue between the batched I/O (mg_step_io), the env logic and the
// rasteriser. Shared by the CUDA kernel (mg_kernels.cu) and the CPU test harness (tests/hostsim),
// so the logic the GPU runs can be checked against the oracle without a GPU.
#pragma once
#include "../../include/memgym_b200.h"
#include "mg_common.h"
#include "mg_frame.h"
#include "mg_mm_logic.h"
#include "mg_mp_logic.h"

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
};


// ---- env policies --------------------------------------------------------------------------------
struct MgFrameJob {      // what lane 0 hands to the warp: up to two frames for this env-step
    int32_t nframes;
    int32_t to_final[2];  // frame goes to final_obs instead of obs
};

struct MMPolicy {
    typedef MgMMK K;
    typedef MgMMState State;
    typedef MgMMFrame Frame;

    static MG_HDN void logic(const K& k, const MgKIO& q, int env, State& s, MgFrameJob& job, Frame* fr) {
        double* gt = q.io.ground_truth ? q.io.ground_truth + 2 * (size_t)env : nullptr;
        job.nframes = 0;
        if (q.mode == MG_MODE_RESET) {
            if (q.io.reset_mask && !q.io.reset_mask[env]) return;
            if (q.rng4) {
                const uint64_t* r = q.rng4 + 4 * (size_t)env;
                s.rng.state_lo = r[0];
                s.rng.state_hi = r[1];
                s.rng.inc_lo = r[2];
                s.rng.inc_hi = r[3];
                s.rng.buf = 0;
                s.rng.has_buf = 0;
            }
            mg_mm_reset_env(k, s, fr[0], gt);
            job.nframes = 1;
            job.to_final[0] = 0;
            return;
        }
        if (!s.live) {  // finished (or never reset): the step is ignored
            q.io.reward[env] = 0.0;
            q.io.done[env] = 1;
            return;
        }
        int a0, a1 = 0;
        if (k.p.variant == MG_MM_GRID) {
            a0 = (int)q.io.actions[env];
        } else {
            a0 = (int)q.io.actions[2 * (size_t)env];
            a1 = (int)q.io.actions[2 * (size_t)env + 1];
        }
        double reward;
        bool done = mg_mm_step_env(k, s, a0, a1, reward, q.io.info + MG_INFO_COLS * (size_t)env, fr[0], gt);
        q.io.reward[env] = reward;
        q.io.done[env] = done ? 1 : 0;
        job.nframes = 1;
        job.to_final[0] = 0;
        if (done && (q.io.flags & MG_FLAG_AUTORESET)) {
            int at = 0;
            if (q.io.final_obs) {
                job.to_final[0] = 1;
                at = 1;
            }
            mg_mm_reset_env(k, s, fr[at], gt);
            job.to_final[at] = 0;
            job.nframes = at + 1;
        }
    }
    template <class Ctx>
    static MG_HD void compose(Ctx& c, uint8_t* buf, const MgAssetView& a, const K& k,
                                                   const Frame& f) {
        mg_mm_compose(c, buf, a, f, k.origin, a.hdr->agent_dim >> 1, a.hdr->icon_x, a.hdr->icon_y);
    }
    template <class Ctx>
    static MG_HD void restore(Ctx& c, uint8_t* buf, const MgAssetView& a, const K& k,
                                                   const Frame& f) {
        mg_mm_restore(c, buf, a, f, a.hdr->agent_dim >> 1, a.hdr->icon_x, a.hdr->icon_y);
    }
};

struct MPPolicy {
    typedef MgMPK K;
    typedef MgMPState State;
    typedef MgMPDraw Frame;

    static MG_HDN void logic(const K& k, const MgKIO& q, int env, State& s, MgFrameJob& job, Frame* fr) {
        job.nframes = 0;
        if (q.mode == MG_MODE_RESET) {
            if (q.io.reset_mask && !q.io.reset_mask[env]) return;
            if (q.rng4) {
                const uint64_t* r = q.rng4 + 4 * (size_t)env;
                s.rng.state_lo = r[0];
                s.rng.state_hi = r[1];
                s.rng.inc_lo = r[2];
                s.rng.inc_hi = r[3];
                s.rng.buf = 0;
                s.rng.has_buf = 0;
            }
            mg_mp_reset_env(k, s, fr[0]);
            job.nframes = 1;
            job.to_final[0] = 0;
            return;
        }
        if (!s.live) {
            q.io.reward[env] = 0.0;
            q.io.done[env] = 1;
            return;
        }
        int a0 = (int)q.io.actions[2 * (size_t)env], a1 = (int)q.io.actions[2 * (size_t)env + 1];
        double reward;
        bool done = mg_mp_step_env(k, s, a0, a1, reward, q.io.info + MG_INFO_COLS * (size_t)env, fr[0]);
        q.io.reward[env] = reward;
        q.io.done[env] = done ? 1 : 0;
        job.nframes = 1;
        job.to_final[0] = 0;
        if (done && (q.io.flags & MG_FLAG_AUTORESET)) {
            int at = 0;
            if (q.io.final_obs) {
                job.to_final[0] = 1;
                at = 1;
            }
            mg_mp_reset_env(k, s, fr[at]);
            job.to_final[at] = 0;
            job.nframes = at + 1;
        }
    }
    template <class Ctx>
    static MG_HD void compose(Ctx& c, uint8_t* buf, const MgAssetView& a, const K&,
                                                   const Frame& f) {
        mg_mp_compose(c, buf, a, f, a.hdr->agent_dim >> 1);
    }
    template <class Ctx>
    static MG_HD void restore(Ctx& c, uint8_t* buf, const MgAssetView& a, const K&,
                                                   const Frame& f) {
        mg_mp_restore(c, buf, a, f, a.hdr->agent_dim >> 1);
    }
};

