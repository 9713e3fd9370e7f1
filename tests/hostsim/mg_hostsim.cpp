// mg_hostsim.cpp — TEST INFRASTRUCTURE: compiles the env logic + frame composition headers the
// CUDA kernels use (csrc/mg_*_logic.h, mg_frame.h, mg_policy.h) for the CPU, so that the code the
// GPU runs can be checked bit-for-bit against the oracle in the CPU-only test tier.
// It is NOT part of the product library (libmemgym_b200.so has no CPU path) and is built only by
// tests/ (see tests/hostsim_util.py). All pointers are HOST pointers here.
#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <vector>

#include "../../endless-memory-gym_b200/csrc/mg_params_host.h"
#include "../../endless-memory-gym_b200/csrc/mg_policy.h"

namespace {
struct HostCtx {
    static constexpr int kLanes = 1;
    int lane = 0, nl = 1;
    void sync() {}
    void stamp(int) {}
    void store16(MgVec16* p, const MgVec16& v, bool) { *p = v; }
    void atomic_or(uint32_t* p, uint32_t v) { *p |= v; }
    void atomic_inc(uint32_t* p) { *p += 1; }
    uint32_t reduce_min(uint32_t v) { return v; }
    uint32_t reduce_or(uint32_t v) { return v; }
    uint32_t shfl(uint32_t v, int) { return v; }
    uint32_t cas_u32(uint32_t* p, uint32_t cmp, uint32_t val) {
        uint32_t old = *p;
        if (old == cmp) *p = val;
        return old;
    }
    void fence_block() {}
    void fence_device() {}
    void nanosleep(unsigned) {}
    uint64_t now_ns() { return 0; }
    uint32_t funnel_r(uint32_t lo, uint32_t hi, unsigned sh) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31u)); }
    void copy_frame(uint8_t* dst, const uint8_t* src) { memcpy(dst, src, MG_FRAME_BYTES); }
};

template <class P, class = void>
struct RecordBytes {
    static constexpr size_t value = sizeof(typename P::State);
};
template <class P>
struct RecordBytes<P, decltype((void)P::kRecordBytes)> {
    static constexpr size_t value = (size_t)P::kRecordBytes;
};
template <class P>
constexpr size_t record_bytes() {
    return RecordBytes<P>::value;
}

int g_ep_resumed = 0;  // helper turns that resumed a parked search (test statistics)

// one turn of the "helper lane" that owns `env` (mg_ep_lane_helper_loop, minus the polling)
void ep_helper_turn(const uint8_t* cextra, uint8_t* state, int env, int max_new, int budget) {
    MgEPSpecRec* g = mg_ep_spec_of(state, env);
    const uint32_t h = g->hword;
    if (!mg_ep_chain_wants(g, h)) return;
    if (g->resume.tag == h) g_ep_resumed++;
    uint16_t info[MG_MP_NODES];
    uint8_t list[MG_MP_NODES];
    MgLaneMem m{info, list, 1, nullptr, 0u, budget, 0};  // budget >= 0: the search stops after that many expansions
    HostCtx ctx;
    mg_ep_lane_work(ctx, g, h, reinterpret_cast<const MgGenCta*>(cextra)->sqrt_d2, m, max_new);
}

// mirrors mg_env_kernel (mg_kernels.cu) with one "warp": env-major loop, all n_steps steps of an env back to
// back, a persistent frame buffer and persistent per-warp scratch (so incremental paths are exercised)
template <class P>
int run_q(const typename P::K& k, const void* assets, void* state, MgKIO q, bool rollout = false) {
    typedef typename P::State State;
    typedef typename P::Frame Frame;
    const bool kCheckClean = P::restores_clean(k);
    const MgAssetView view = mg_asset_view(assets);
    alignas(128) static thread_local uint8_t buf[MG_FRAME_BYTES];
    static thread_local std::vector<uint8_t> wextra, cextra;  // the kernel's per-warp / per-CTA scratch
    wextra.assign(P::kWarpExtra + 256, 0);
    cextra.assign(P::kCtaExtra + 256, 0);
    uint8_t* wx = wextra.data() + ((128 - ((uintptr_t)wextra.data() & 127)) & 127);  // 128-byte aligned like smem
    uint8_t* cx = cextra.data() + ((128 - ((uintptr_t)cextra.data() & 127)) & 127);
    P::cta_init(cx, 0, 1);
    P::warp_init(wx, 0);
    const MgScratch scratch{wx, cx};
    HostCtx ctx;
    ctx.copy_frame(buf, view.tmpl[0]);
    constexpr size_t kRecord = record_bytes<P>();
    // Endless Mystery Path rollouts: the kernel's helper warps generate the next episode's path ahead of the reset,
    // one env per lane (mg_ep_lane_work). Here ONE "helper lane" gets a turn after some of the steps and sometimes
    // does a single segment only or stops in the middle of a search (and resumes it at its next turn), so that adopted, resumed,
    // stale and missing episodes all occur.
    constexpr bool kLaneHelper = std::is_same<P, EPPolicy>::value;
    if constexpr (kLaneHelper) {
        if (rollout && !getenv("MG_HOSTSIM_NO_HELPER")) P::set_helpers(wx, 0, 1, false, (uint8_t*)state);
    }
    for (int env = 0; env < q.n; ++env) {
        State* st = reinterpret_cast<State*>((uint8_t*)state + (size_t)env * kRecord);
        if constexpr (P::kPipelined) {
            // mirrors the pipelined rollout loop of mg_env_kernel<P, true>: step t + 1 is computed BEFORE frame t is
            // rasterised, auto-resets (and the endless coin respawn) are deferred to the gap between two frames
            if (rollout && st->live) {
                const auto act = [&](int t, int& a0, int& a1) {
                    const int64_t* a = q.io.actions + (size_t)t * q.act_stride + (size_t)env * q.act_dims;
                    a0 = (int)a[0];
                    a1 = q.act_dims > 1 ? (int)a[1] : 0;
                };
                Frame fr[2];
                int a0, a1;
                P::pipe_begin(ctx, env, scratch);
                act(0, a0, a1);
                P::pipe_step(ctx, k, q, view, env, 0, *st, a0, a1, fr[0], scratch);
                if (P::pipe_needs_deferred(scratch)) P::pipe_deferred(ctx, k, q, view, env, 0, *st, fr[0], scratch);
                for (int t = 0; t < q.n_steps; ++t) {
                    if (t + 1 < q.n_steps) {
                        act(t + 1, a0, a1);
                        P::pipe_step(ctx, k, q, view, env, t + 1, *st, a0, a1, fr[(t + 1) & 1], scratch);
                    }
                    P::emit_base(ctx, view, k, fr[t & 1], scratch);
                    const bool deferred = t + 1 < q.n_steps && P::pipe_needs_deferred(scratch);
                    uint8_t* dst = q.io.obs + (size_t)(t % q.ring_n) * q.obs_stride + (size_t)env * MG_FRAME_BYTES;
                    P::emit_rest(ctx, dst, view, k, fr[t & 1], scratch, false);
                    if (deferred) P::pipe_deferred(ctx, k, q, view, env, t + 1, *st, fr[(t + 1) & 1], scratch);
                }
                continue;
            }
        }
        for (int t = 0; t < q.n_steps; ++t) {
            MgFrameJob job;
            Frame fr[2];
            int a0 = 0, a1 = 0;
            if (q.mode == MG_MODE_STEP) {
                const int64_t* a = q.io.actions + (size_t)t * q.act_stride + (size_t)env * q.act_dims;
                a0 = (int)a[0];
                if (q.act_dims > 1) a1 = (int)a[1];
            }
            P::logic(ctx, k, q, view, env, t, *st, a0, a1, job, fr, scratch);
            if constexpr (kLaneHelper) {
                if (rollout && (env + t) % 3 != 0)
                    ep_helper_turn(cx, (uint8_t*)state, env, (env + 2 * t) % 5 == 0 ? 1 : 2, (env * 7 + t) % 4 == 0 ? 3 + t % 9 : -1);
            }
            for (int fi = 0; fi < job.nframes; ++fi) {
                uint8_t* dst = job.to_final[fi] ? q.io.final_obs + (size_t)env * MG_FRAME_BYTES
                                                : q.io.obs + (size_t)(t % q.ring_n) * q.obs_stride + (size_t)env * MG_FRAME_BYTES;
                if constexpr (P::kDirectStore) {
                    P::emit(ctx, dst, view, k, fr[fi], scratch, false);
                    continue;
                }
                P::compose(ctx, buf, view, k, fr[fi], scratch);
                memcpy(dst, buf, MG_FRAME_BYTES);
                P::restore(ctx, buf, view, k, fr[fi], scratch);
                // the buffer must be clean again: catches any compose/restore asymmetry (envs with an un-patch path)
                if (kCheckClean && memcmp(buf, view.tmpl[0], MG_FRAME_BYTES) != 0) return -100;
            }
            if (!rollout && q.io.debug_obs && job.nframes > 0)  // like the kernel: the ground-truth render of the obs frame
                P::emit_debug(ctx, q.io.debug_obs + (size_t)env * MG_DEBUG_BYTES, view, k, fr[job.nframes - 1], *st, scratch);
        }
    }
    return 0;
}

template <class P>
int run(const typename P::K& k, const void* assets, void* state, const uint64_t* rng4, const mg_step_io* io, int n,
        int mode, int act_dims) {
    MgKIO q;
    memset(&q, 0, sizeof(q));
    q.n_steps = 1;
    q.ring_n = 1;
    q.io = *io;
    q.rng4 = rng4;
    q.mode = mode;
    q.n = n;
    q.act_dims = act_dims;
    return run_q<P>(k, assets, state, q);
}

template <class P>
int run_rollout(const typename P::K& k, const void* assets, void* state, const mg_rollout_io* r, int n, int act_dims) {
    MgKIO q;
    memset(&q, 0, sizeof(q));
    q.io.actions = r->actions;
    q.io.obs = r->obs;
    q.io.final_obs = r->final_obs;
    q.io.reward = r->reward;
    q.io.done = r->done;
    q.io.info = r->info;
    q.io.ground_truth = r->ground_truth;
    q.io.vector_obs = r->vector_obs;
    q.io.flags = r->flags | MG_FLAG_AUTORESET;
    q.mode = MG_MODE_STEP;
    q.n = n;
    q.act_dims = act_dims;
    q.n_steps = r->n_steps;
    q.ring_n = r->ring;
    q.act_stride = (int64_t)n * act_dims;
    q.obs_stride = (int64_t)n * MG_FRAME_BYTES;
    q.rew_stride = n;
    q.done_stride = n;
    q.info_stride = r->info_per_step ? (int64_t)n * MG_INFO_COLS : 0;
    return run_q<P>(k, assets, state, q, true);
}
}  // namespace

extern "C" int mgsim_mm(const mg_mm_params* p, const void* assets, void* state, const uint64_t* rng4,
                        const mg_step_io* io, int n, int mode) {
    if (mg_mm_check_params(p)) return MG_EINVAL;
    MgMMK k;
    mg_mm_derive(p, &k);
    return run<MMPolicy>(k, assets, state, rng4, io, n, mode, mg_mm_is_grid(p->variant) ? 1 : 2);
}
extern "C" int mgsim_mp(const mg_mp_params* p, const void* assets, void* state, const uint64_t* rng4,
                        const mg_step_io* io, int n, int mode) {
    if (mg_mp_check_params(p)) return MG_EINVAL;
    MgMPK k;
    mg_mp_derive(p, &k);
    return run<MPPolicy>(k, assets, state, rng4, io, n, mode, p->variant == MG_MP_GRID ? 1 : 2);
}
extern "C" int mgsim_ss(const mg_ss_params* p, const void* assets, void* state, const uint64_t* rng4,
                        const mg_step_io* io, int n, int mode) {
    if (mg_ss_check_params(p)) return MG_EINVAL;
    MgSSK k;
    mg_ss_derive(p, &k);
    if (k.endless) return run<ESSPolicy>(k, assets, state, rng4, io, n, mode, 2);
    return run<SSPolicy>(k, assets, state, rng4, io, n, mode, 2);
}
extern "C" int mgsim_ep(const mg_ep_params* p, const void* assets, void* state, const uint64_t* rng4,
                        const mg_step_io* io, int n, int mode) {
    if (mg_ep_check_params(p)) return MG_EINVAL;
    MgEPK k;
    mg_ep_derive(p, &k);
    return run<EPPolicy>(k, assets, state, rng4, io, n, mode, 1);
}
extern "C" int mgsim_ep_resumed() { return g_ep_resumed; }
extern "C" int mgsim_ep_rollout(const mg_ep_params* p, const void* assets, void* state, const mg_rollout_io* r, int n) {
    if (mg_ep_check_params(p)) return MG_EINVAL;
    MgEPK k;
    mg_ep_derive(p, &k);
    return run_rollout<EPPolicy>(k, assets, state, r, n, 1);
}
extern "C" int mgsim_mm_rollout(const mg_mm_params* p, const void* assets, void* state, const mg_rollout_io* r, int n) {
    if (mg_mm_check_params(p)) return MG_EINVAL;
    MgMMK k;
    mg_mm_derive(p, &k);
    return run_rollout<MMPolicy>(k, assets, state, r, n, mg_mm_is_grid(p->variant) ? 1 : 2);
}
extern "C" int mgsim_mp_rollout(const mg_mp_params* p, const void* assets, void* state, const mg_rollout_io* r, int n) {
    if (mg_mp_check_params(p)) return MG_EINVAL;
    MgMPK k;
    mg_mp_derive(p, &k);
    return run_rollout<MPPolicy>(k, assets, state, r, n, p->variant == MG_MP_GRID ? 1 : 2);
}
extern "C" int mgsim_ss_rollout(const mg_ss_params* p, const void* assets, void* state, const mg_rollout_io* r, int n) {
    if (mg_ss_check_params(p)) return MG_EINVAL;
    MgSSK k;
    mg_ss_derive(p, &k);
    if (k.endless) return run_rollout<ESSPolicy>(k, assets, state, r, n, 2);
    return run_rollout<SSPolicy>(k, assets, state, r, n, 2);
}
