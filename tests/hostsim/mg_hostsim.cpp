// mg_hostsim.cpp — TEST INFRASTRUCTURE: compiles the env logic + frame composition headers the
// CUDA kernels use (csrc/mg_*_logic.h, mg_frame.h, mg_policy.h) for the CPU, so that the code the
// GPU runs can be checked bit-for-bit against the oracle in the CPU-only test tier.
// It is NOT part of the product library (libmemgym_b200.so has no CPU path) and is built only by
// tests/ (see tests/hostsim_util.py). All pointers are HOST pointers here.
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "../../endless-memory-gym_b200/csrc/mg_params_host.h"
#include "../../endless-memory-gym_b200/csrc/mg_policy.h"

namespace {
struct HostCtx {
    int lane = 0, nl = 1;
    void sync() {}
    void copy_frame(uint8_t* dst, const uint8_t* src) { memcpy(dst, src, MG_FRAME_BYTES); }
};

// mirrors mg_env_kernel (mg_kernels.cu) with one "warp" and a persistent frame buffer
template <class P>
int run(const typename P::K& k, const void* assets, void* state, const uint64_t* rng4, const mg_step_io* io, int n,
        int mode, int act_dims) {
    typedef typename P::State State;
    typedef typename P::Frame Frame;
    constexpr bool kCheckClean = !std::is_same<P, SSPolicy>::value;
    MgKIO q;
    memset(&q, 0, sizeof(q));
    q.n_steps = 1;
    q.ring_n = 1;
    q.io = *io;
    q.rng4 = rng4;
    q.mode = mode;
    q.n = n;
    q.stage_cap = 0;
    q.act_dims = act_dims;
    const MgAssetView view = mg_asset_view(assets);
    static thread_local uint8_t buf[MG_FRAME_BYTES];
    HostCtx ctx;
    ctx.copy_frame(buf, view.tmpl[0]);
    for (int env = 0; env < n; ++env) {
        State* st = reinterpret_cast<State*>((uint8_t*)state + (size_t)env * sizeof(State));
        MgFrameJob job;
        Frame fr[2];
        int a0 = 0, a1 = 0;
        if (mode == MG_MODE_STEP) {
            a0 = (int)io->actions[(size_t)env * act_dims];
            if (act_dims > 1) a1 = (int)io->actions[(size_t)env * act_dims + 1];
        }
        P::logic(k, q, view, env, 0, *st, a0, a1, job, fr);
        for (int fi = 0; fi < job.nframes; ++fi) {
            uint8_t* dst = (job.to_final[fi] ? io->final_obs : io->obs) + (size_t)env * MG_FRAME_BYTES;
            P::compose(ctx, buf, view, k, fr[fi]);
            memcpy(dst, buf, MG_FRAME_BYTES);
            P::restore(ctx, buf, view, k, fr[fi]);
            // the buffer must be clean again: catches any compose/restore asymmetry (envs with an un-patch path)
            if (kCheckClean && memcmp(buf, view.tmpl[0], MG_FRAME_BYTES) != 0) return -100;
        }
    }
    return 0;
}
}  // namespace

extern "C" int mgsim_mm(const mg_mm_params* p, const void* assets, void* state, const uint64_t* rng4,
                        const mg_step_io* io, int n, int mode) {
    if (mg_mm_check_params(p)) return MG_EINVAL;
    MgMMK k;
    mg_mm_derive(p, &k);
    return run<MMPolicy>(k, assets, state, rng4, io, n, mode, p->variant == MG_MM_GRID ? 1 : 2);
}
extern "C" int mgsim_mp(const mg_mp_params* p, const void* assets, void* state, const uint64_t* rng4,
                        const mg_step_io* io, int n, int mode) {
    if (mg_mp_check_params(p)) return MG_EINVAL;
    MgMPK k;
    mg_mp_derive(p, &k);
    return run<MPPolicy>(k, assets, state, rng4, io, n, mode, 2);
}
extern "C" int mgsim_ss(const mg_ss_params* p, const void* assets, void* state, const uint64_t* rng4,
                        const mg_step_io* io, int n, int mode) {
    if (mg_ss_check_params(p)) return MG_EINVAL;
    MgSSK k;
    mg_ss_derive(p, &k);
    return run<SSPolicy>(k, assets, state, rng4, io, n, mode, 2);
}
