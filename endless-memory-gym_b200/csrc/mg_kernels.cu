// mg_kernels.cu — the sm_100a reset/step/rollout kernel and its C-ABI launchers.
//
// One launch per env type does, for every env of the batch: kinematics + tile collision, the env
// state machine (incl. in-kernel PCG64 draws for auto-reset), and the rasterisation of the
// 84x84x3 observation. ONE WARP OWNS ONE ENV:
//   * the env's state record is loaded/stored with one coalesced 8-byte access per lane into a per-warp
//     shared-memory slot; lane 0 runs the (branchy, scalar) state machine there;
//   * each warp keeps one 21168-byte frame buffer in shared memory that holds the clean background
//     template between frames. A frame = patch a few hundred sprite bytes (sprite pixel lists + palette
//     are staged in shared memory), then ONE `cp.async.bulk.global.shared::cta` (TMA bulk store, SASS
//     UBLKCP) streams the whole frame to HBM as full lines while the warp already runs the next state
//     machine; when the bulk read of the buffer is done (`cp.async.bulk.wait_group.read`) the patch is
//     undone (closed-form background, no memory traffic).
//   * background templates are brought in with `cp.async.bulk.shared::cluster.global` + mbarrier.
//   * rollout mode (n_steps > 1): a warp runs ALL steps of its envs back to back with the record
//     resident in shared memory — no kernel boundary, no inter-warp dependency, the HBM write stream
//     never drains.
// The grid is one CTA per SM (persistent over the env batch), up to 10 warps = 10 frame buffers
// (212 KB of the 227 KB shared memory) so ~200 KB of stores are in flight per SM.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/memgym_b200.h"
#include "mg_common.h"
#include "mg_frame.h"
#include "mg_mm_logic.h"
#include "mg_mp_logic.h"
#include "mg_ep_logic.h"
#include "mg_ss_logic.h"
#include "mg_params_host.h"
#include "mg_device_host.h"
#include "mg_policy.h"


// ---- PTX wrappers (TMA bulk copies, mbarrier, proxy fence) ---------------------------------------
__device__ __forceinline__ uint32_t mg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mg_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mg_bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gdst)),
                 "r"(mg_smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mg_bulk_store_hint(void* gdst, const void* ssrc, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(
                     __cvta_generic_to_global(gdst)),
                 "r"(mg_smem_u32(ssrc)), "r"(bytes), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t mg_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void mg_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void mg_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all but the newest `keep` bulk groups have fully completed (their global writes are performed)
__device__ __forceinline__ void mg_bulk_wait_done(int keep) {
    if (keep <= 0)
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    else if (keep == 1)
        asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
    else if (keep == 2)
        asm volatile("cp.async.bulk.wait_group 2;" ::: "memory");
    else
        asm volatile("cp.async.bulk.wait_group 3;" ::: "memory");
}
__device__ __forceinline__ void mg_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mg_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mg_bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mg_smem_u32(sdst)),
        "l"(__cvta_generic_to_global(gsrc)), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void mg_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ uint64_t mg_now() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// this env's action components. int64[n][2] rows are 16-byte aligned whenever the base is: one 128-bit load (the
// actions may live in pinned host memory, where every load is a PCIe read)
__device__ __forceinline__ void mg_load_actions(const int64_t* a, int dims, int64_t& a0, int64_t& a1) {
    if (dims > 1) {
        if ((reinterpret_cast<uintptr_t>(a) & 15) == 0) {
            const longlong2 v = *reinterpret_cast<const longlong2*>(a);
            a0 = v.x;
            a1 = v.y;
        } else {
            a0 = a[0];
            a1 = a[1];
        }
    } else {
        a0 = a[0];
    }
}

// per-CTA words: the asset staging barrier and the work counter of the CTA-local dynamic env -> team scheduling
#define MG_CTA_HDR_BYTES 128
struct MgCtaHdr {
    uint64_t stage_bar;
    uint32_t next;
    uint32_t teams_done;  // teams that have left the kernel (helper warps leave when all have)
};

// execution context of mg_frame.h on the device: the 32 lanes of the warp that owns the env
struct MgWarpCtx {
    static constexpr int kLanes = 32;  // lanes of a single-warp team (cooperative code sized at compile time)
    int lane, nl;
    uint32_t bar, phase;
    uint64_t policy;  // L2 evict-first cache policy
    uint64_t* trow;   // this env's trace row (mg_step_io.trace) or nullptr
    int bar_id;       // named barrier of this team (teams of more than one warp), 0 = a single warp
    __device__ __forceinline__ void stamp(int col) {
        if (MG_UNLIKELY(trow != nullptr) && lane == 0) trow[col] = mg_now();
    }
    // all threads of the team that owns the env (one warp, or kTeamWarps warps on a named barrier)
    __device__ __forceinline__ void sync() {
        if (bar_id == 0)
            __syncwarp();
        else
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nl) : "memory");
    }
    // 128-bit store to global memory; `stream` = L2 evict-first (the frame is not re-read soon)
    __device__ __forceinline__ void store16(MgVec16* p, const MgVec16& v, bool stream) {
        if (stream)
            asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(__cvta_generic_to_global(p)),
                         "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy)
                         : "memory");
        else
            *p = v;
    }
    __device__ __forceinline__ void atomic_or(uint32_t* p, uint32_t v) { atomicOr(p, v); }
    __device__ __forceinline__ void atomic_inc(uint32_t* p) { atomicAdd(p, 1u); }
    // warp reductions (single-warp teams only): every lane gets the result
    __device__ __forceinline__ uint32_t reduce_min(uint32_t v) { return __reduce_min_sync(0xffffffffu, v); }
    __device__ __forceinline__ uint32_t reduce_or(uint32_t v) { return __reduce_or_sync(0xffffffffu, v); }
    __device__ __forceinline__ uint32_t shfl(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
    // shared-memory hand-over between warps of a CTA (teams <-> helper warps)
    __device__ __forceinline__ uint32_t cas_u32(uint32_t* p, uint32_t cmp, uint32_t val) { return atomicCAS(p, cmp, val); }
    __device__ __forceinline__ void fence_block() { __threadfence_block(); }
    __device__ __forceinline__ void fence_device() { __threadfence(); }
    __device__ __forceinline__ void nanosleep(unsigned ns) { __nanosleep(ns); }
    __device__ __forceinline__ uint64_t now_ns() { return mg_now(); }
    __device__ __forceinline__ uint32_t funnel_r(uint32_t lo, uint32_t hi, unsigned sh) { return __funnelshift_r(lo, hi, sh); }
    // whole-frame template -> this warp's frame buffer, by TMA
    __device__ __forceinline__ void copy_frame(uint8_t* dst, const uint8_t* src) {
        mg_fence_proxy_async();  // earlier generic-proxy accesses to dst are ordered before the async write
        sync();
        if (lane == 0) {
            mg_mbar_expect_tx(bar, MG_FRAME_BYTES);
            mg_bulk_load(dst, src, MG_FRAME_BYTES, bar);
        }
        mg_mbar_wait(bar, phase);
        phase ^= 1u;
    }
};

// pipelined Searing-Spotlights rollouts: 1 = the raster warp also builds the lit bitmap while the logic warp computes the
// next step (measured: see profiles/r01b_summary.md), 0 = both warps build it after the barrier
#ifndef MG_SS_LIT_EARLY
#define MG_SS_LIT_EARLY 0
#endif

// distance between two env records: the record itself unless the policy keeps more behind it (P::kRecordBytes)
template <class P, class = void>
struct MgRecordBytes {
    static constexpr int value = (int)sizeof(typename P::State);
};
template <class P>
struct MgRecordBytes<P, decltype((void)P::kRecordBytes)> {
    static constexpr int value = P::kRecordBytes;
};
template <class P>
__host__ __device__ constexpr int mg_record_bytes() {
    return MgRecordBytes<P>::value;
}

// Single-step launches: a path search that is under way when the env teams are done goes on until the launch is this
// old (or the segment is finished). Without it a fast launch leaves the helper lanes ~30 us, the slowest searches fall
// behind their envs and one reset per launch finds its chain empty — the whole launch then waits ~60 us for one
// cooperative generation (env.step() loop, 8192 envs: 117 us per step without, 87 us with; profiles/r02_emp_lookahead.md)
#ifndef MG_EP_STEP_HELPER_NS
#define MG_EP_STEP_HELPER_NS 40000
#endif

// ---- the kernel ----------------------------------------------------------------------------------
// A "team" = the threads that own one env at a time: one warp, or P::kTeamWarps warps on a named barrier for
// env types whose per-frame work is instruction-bound (more warps per frame buffer hide more latency).
// kRollout: the instantiation behind mg_*_rollout (n_steps consecutive steps per env, optional global work counter);
// false = reset / single step (one pass over the batch, CTA-local scheduling only). Two instantiations keep each
// kernel's code small and let the compiler drop the other mode's paths.
template <class P, bool kRollout>
__global__ void __launch_bounds__(P::kMaxThreads, 1)
    mg_env_kernel(const __grid_constant__ typename P::K k, const __grid_constant__ MgKIO q,
                  const uint8_t* __restrict__ assets, uint8_t* __restrict__ state) {
    typedef typename P::State State;
    typedef typename P::Frame Frame;
    constexpr int kTeam = 32 * P::kTeamWarps;          // threads per team
    constexpr int kStateBytes = (int)sizeof(State);
    constexpr int kRecordBytes = mg_record_bytes<P>();  // distance between two records (>= kStateBytes)
    constexpr int kWords = kStateBytes / 8;             // 8-byte words per record
    constexpr int kPerLane = (kWords + kTeam - 1) / kTeam;  // words each thread moves
    constexpr int kFrames = P::kDirectStore ? 2 : 3;  // this step's (up to two) frames + the one being un-patched
    constexpr int kSlot = ((kStateBytes + kFrames * (int)sizeof(Frame) + (int)sizeof(MgFrameJob) + 16 + 127) / 128) * 128;
    extern __shared__ __align__(128) uint8_t smem[];

    // helper warps (rollout launches of policies that have them) sit behind the teams and own no env
    const int nhelp = P::kHelperWarps > 0 ? q.n_helpers : 0;
    const int lane = threadIdx.x % kTeam, team = threadIdx.x / kTeam, nteams = blockDim.x / kTeam - nhelp;
    const bool is_helper = team >= nteams;
    uint8_t* buf = smem + (size_t)team * P::kFrameBuf;
    uint8_t* slot = smem + (size_t)nteams * P::kFrameBuf + (size_t)team * kSlot;
    uint8_t* wextra = smem + (size_t)nteams * (P::kFrameBuf + kSlot) + (size_t)team * P::kWarpExtra;
    uint8_t* cextra = smem + (size_t)nteams * (P::kFrameBuf + kSlot + P::kWarpExtra);
    MgCtaHdr* hdr = reinterpret_cast<MgCtaHdr*>(cextra + P::kCtaExtra);
    uint8_t* stage = cextra + P::kCtaExtra + MG_CTA_HDR_BYTES;
    const MgScratch scratch{wextra, cextra};
    State* st = reinterpret_cast<State*>(slot);
    Frame* fr = reinterpret_cast<Frame*>(slot + kStateBytes);
    MgFrameJob* job = reinterpret_cast<MgFrameJob*>(slot + kStateBytes + kFrames * sizeof(Frame));
    uint64_t* bar = reinterpret_cast<uint64_t*>(slot + kSlot - 8);
    uint32_t* grab = reinterpret_cast<uint32_t*>(slot + kSlot - 16);  // next env index, broadcast to the team

    // ---- prologue: touches only the (constant) asset blob, so under programmatic dependent launch it
    // overlaps the tail of the previous kernel in the stream ---------------------------------------------
    const bool tracing = q.io.trace != nullptr;
    uint64_t tr0 = 0, tr1 = 0, tr2 = 0;
    if (tracing) tr0 = mg_now();
    MgAssetView view = mg_asset_view(assets);
    MgWarpCtx ctx{lane, kTeam, mg_smem_u32(bar), 0u, mg_policy_evict_first(), nullptr, P::kTeamWarps > 1 ? team + 1 : 0};
    if (lane == 0 && !is_helper) mg_mbar_init(ctx.bar, 1);
    // This CTA owns envs blockIdx.x + j * gridDim.x (consecutive envs on different SMs: remainders spread
    // evenly). Team i starts with j = i; every further j comes from a SHARED-MEMORY work counter, so a team
    // held up by an auto-reset ends up with fewer envs and no global atomic sits on the step path.
    const int G = gridDim.x;
    int env = (int)blockIdx.x + team * G;
    const bool active = env < q.n && !is_helper;
    if (P::kFrameBuf && active && lane == 0) {  // background template -> this team's frame buffer (TMA, L2-resident source)
        mg_mbar_expect_tx(ctx.bar, MG_FRAME_BYTES);
        mg_bulk_load(buf, view.tmpl[0], MG_FRAME_BYTES, ctx.bar);
    }
    // header, sprite lists, bitmaps, tables -> shared memory by one TMA bulk load: compose never waits on L2
    const uint32_t stage_bytes = view.hdr->stage_bytes;
    const bool staged = (int)stage_bytes <= q.stage_cap;
    if (threadIdx.x == 0) {
        hdr->next = (uint32_t)nteams;
        hdr->teams_done = 0u;
        const uint32_t sbar = mg_smem_u32(&hdr->stage_bar);
        mg_mbar_init(sbar, 1);
        if (staged) {
            mg_mbar_expect_tx(sbar, stage_bytes);
            mg_bulk_load(stage, assets, stage_bytes, sbar);
        }
    }
    P::cta_init(cextra, threadIdx.x, blockDim.x);
    if (!is_helper) {
        P::warp_init(wextra, lane);
        if constexpr (P::kHelperWarps > 0) P::set_helpers(wextra, lane, q.n_lane_helpers, nhelp - q.n_lane_helpers, state);
    }
    __syncthreads();
    if (staged) {
        mg_mbar_wait(mg_smem_u32(&hdr->stage_bar), 0u);
        view = mg_asset_view2(stage, assets);
    }
    if (tracing) tr1 = mg_now();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // previous kernel's state / obs writes are visible
    // Rollouts may pass a GLOBAL work counter instead (mg_rollout_io.sched): one atomic per env per T steps, and
    // SMs that stream faster take more envs. The last team to leave re-arms it for the next launch.
    uint32_t* const sched = kRollout ? q.sched : nullptr;
    const int total_teams = G * nteams;
    if constexpr (P::kHelperWarps > 0) {
        if (is_helper) {
            MgWarpCtx hctx{lane, kTeam, 0u, 0u, 0ull, nullptr, 0};
            uint8_t* const wextra0 = smem + (size_t)nteams * (P::kFrameBuf + kSlot);
            const int hidx = team - nteams, nlane = q.n_lane_helpers;
            if (hidx < nlane)
                P::helper_loop(hctx, cextra, nteams, hidx, nlane, (int)blockIdx.x, (int)gridDim.x, &hdr->teams_done,
                               kRollout ? nteams : 1, kRollout ? 0u : (uint32_t)MG_EP_STEP_HELPER_NS, state, q.n);
            else if constexpr (!kRollout)
                P::helper_loop_step(hctx, cextra, wextra0, nteams, hidx - nlane, &hdr->teams_done, state, q.io.ground_truth);
            return;
        }
    }
    const auto team_exit = [&]() {
        if (P::kHelperWarps > 0 && nhelp > 0 && lane == 0) atomicAdd(&hdr->teams_done, 1u);
        if (sched && lane == 0) {
            __threadfence();
            if (atomicAdd(sched + 1, 1u) == (uint32_t)total_teams - 1u) {
                sched[0] = 0u;
                sched[1] = 0u;
                __threadfence();
            }
        }
    };
    if (!active) {
        team_exit();
        return;
    }
    if (tracing) tr2 = mg_now();

    const int T = kRollout ? q.n_steps : 1;
    const bool stepping = q.mode == MG_MODE_STEP;
    const uint64_t evict_first = ctx.policy;
    // ---- first record + first actions in flight while the template lands --------------------------------
    uint64_t pre[kPerLane];
    int64_t act0 = 0, act1 = 0;
    {
        const uint64_t* g = reinterpret_cast<const uint64_t*>(state + (size_t)env * kRecordBytes);
#pragma unroll
        for (int w = 0; w < kPerLane; ++w)
            if (lane + kTeam * w < kWords) pre[w] = g[lane + kTeam * w];
        if (lane == 0 && stepping) {
            mg_load_actions(q.io.actions + (size_t)env * q.act_dims, q.act_dims, act0, act1);
        }
    }
    if (P::kFrameBuf) {
        mg_mbar_wait(ctx.bar, ctx.phase);
        ctx.phase ^= 1u;
    }

    bool pending = false;  // a bulk store is still reading `buf`; `prev` is the patch to undo afterwards
    Frame* const prev = fr + (kFrames - 1);  // kept in the team's shared slot (a per-thread copy would live in local memory)
    static_assert(sizeof(Frame) % 4 == 0, "frame descriptions are copied word-wise");
    bool have = true;
    while (have) {
        uint64_t* s64 = reinterpret_cast<uint64_t*>(st);
#pragma unroll
        for (int w = 0; w < kPerLane; ++w)
            if (lane + kTeam * w < kWords) s64[lane + kTeam * w] = pre[w];
        ctx.sync();
        int nxt = q.n;
        bool have_next = false;
        if constexpr (kRollout && P::kPipelined) {
            // ---- pipelined rollout (two-warp teams): warp 0 computes step t + 1 while warp 1 patches the base buffer
            // for frame t; then both build the lit bitmap and copy frame t out. Auto-resets run on the whole team
            // between two frames. This instantiation carries no other step loop (register pressure): no trace stamps.
            if (!st->live) {  // never reset (or finished without auto-reset): the steps are ignored
                if (lane == 0)
                    for (int t = 0; t < T; ++t) {
                        q.io.reward[(size_t)t * q.rew_stride + env] = 0.0;
                        q.io.done[(size_t)t * q.done_stride + env] = 1;
                    }
            } else {
                MgWarpCtx c32 = ctx;  // one warp of the team, warp-local barrier
                c32.lane = lane & 31;
                c32.nl = 32;
                c32.bar_id = 0;
                const int wsel = lane >> 5;
                const bool stream = (q.io.flags & MG_FLAG_STREAM_OBS) != 0;
                P::pipe_begin(ctx, env, scratch);
                if (wsel == 0) {
                    const int a0 = (int)act0, a1 = (int)act1;
                    if (lane == 0 && T > 1) {
                        mg_load_actions(q.io.actions + (size_t)q.act_stride + (size_t)env * q.act_dims, q.act_dims, act0, act1);
                    }
                    P::pipe_step(c32, k, q, view, env, 0, *st, a0, a1, fr[0], scratch);
                }
                ctx.sync();
                const bool deferred0 = P::pipe_needs_deferred(scratch);
                ctx.sync();  // (read by everyone before warp 0 starts the next step and rewrites the word)
                if (deferred0) P::pipe_deferred(ctx, k, q, view, env, 0, *st, fr[0], scratch);
                for (int t = 0; t < T; ++t) {
                    const Frame& f = fr[t & 1];
                    if (wsel == 0) {
                        if (t + 1 < T) {
                            const int a0 = (int)act0, a1 = (int)act1;
                            if (lane == 0 && t + 2 < T) {
                                mg_load_actions(q.io.actions + (size_t)(t + 2) * q.act_stride + (size_t)env * q.act_dims, q.act_dims, act0, act1);
                            }
                            P::pipe_step(c32, k, q, view, env, t + 1, *st, a0, a1, fr[(t + 1) & 1], scratch);
                        }
                    } else {
                        if (MG_SS_LIT_EARLY)
                            P::emit_base_lit(c32, view, k, f, scratch);
                        else
                            P::emit_base(c32, view, k, f, scratch);
                    }
                    ctx.sync();
                    const bool deferred = t + 1 < T && P::pipe_needs_deferred(scratch);
                    uint8_t* dst = q.io.obs + (size_t)(t % q.ring_n) * q.obs_stride + (size_t)env * MG_FRAME_BYTES;
                    if (MG_SS_LIT_EARLY)
                        P::emit_copy(ctx, dst, view, k, f, scratch, stream);
                    else
                        P::emit_rest(ctx, dst, view, k, f, scratch, stream);
                    ctx.sync();
                    if (deferred) P::pipe_deferred(ctx, k, q, view, env, t + 1, *st, fr[(t + 1) & 1], scratch);
                }
            }
            {
                // the team's next env: claimed after its last step; record + first actions fetched now
                if (lane == 0)
                    *grab = sched ? (uint32_t)total_teams + atomicAdd(sched, 1u)
                                  : (uint32_t)blockIdx.x + atomicAdd(&hdr->next, 1u) * (uint32_t)G;
                ctx.sync();
                nxt = (int)*grab;
                have_next = nxt < q.n;
                if (have_next) {
                    const uint64_t* g = reinterpret_cast<const uint64_t*>(state + (size_t)nxt * kRecordBytes);
#pragma unroll
                    for (int w = 0; w < kPerLane; ++w)
                        if (lane + kTeam * w < kWords) pre[w] = g[lane + kTeam * w];
                    if (lane == 0) {
                        mg_load_actions(q.io.actions + (size_t)nxt * q.act_dims, q.act_dims, act0, act1);
                    }
                }
            }
        }
        for (int t = 0; t < ((kRollout && P::kPipelined) ? 0 : T); ++t) {
            const int a0 = (int)act0, a1 = (int)act1;
            uint64_t* trow = tracing ? q.io.trace + (size_t)env * MG_TRACE_COLS : nullptr;
            ctx.trow = trow;
            if (MG_UNLIKELY(trow != nullptr) && lane == 0) {
                unsigned smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                trow[0] = tr0, trow[1] = tr1, trow[2] = tr2, trow[3] = mg_now(), trow[7] = smid;
            }
            if (t + 1 < T) {  // prefetch the next step's actions: the latency hides behind this step's work
                if (lane == 0) {
                    mg_load_actions(q.io.actions + (size_t)(t + 1) * q.act_stride + (size_t)env * q.act_dims, q.act_dims, act0, act1);
                }
            } else {
                // the env's last step: claim this team's next env — as LATE as possible, so that a team held up by
                // an auto-reset ends up with fewer envs — and start fetching its record and first actions
                if (lane == 0)
                    *grab = sched ? (uint32_t)total_teams + atomicAdd(sched, 1u)
                                  : (uint32_t)blockIdx.x + atomicAdd(&hdr->next, 1u) * (uint32_t)G;
                ctx.sync();
                nxt = (int)*grab;
                have_next = nxt < q.n;
                if (have_next) {
                    const uint64_t* g = reinterpret_cast<const uint64_t*>(state + (size_t)nxt * kRecordBytes);
#pragma unroll
                    for (int w = 0; w < kPerLane; ++w)
                        if (lane + kTeam * w < kWords) pre[w] = g[lane + kTeam * w];
                    if (lane == 0 && stepping) {
                        mg_load_actions(q.io.actions + (size_t)nxt * q.act_dims, q.act_dims, act0, act1);
                    }
                }
            }
            P::logic(ctx, k, q, view, env, t, *st, a0, a1, *job, fr, scratch);
            ctx.sync();
            if (MG_UNLIKELY(trow != nullptr) && lane == 0) trow[4] = mg_now();
            const int nframes = job->nframes;
            for (int fi = 0; fi < nframes; ++fi) {
                const Frame& f = fr[fi];  // lives in the team's shared slot until the next logic call
                const bool to_final = job->to_final[fi] != 0;
                uint8_t* dst = to_final ? q.io.final_obs + (size_t)env * MG_FRAME_BYTES
                                        : q.io.obs + (size_t)(t % q.ring_n) * q.obs_stride + (size_t)env * MG_FRAME_BYTES;
                if constexpr (P::kDirectStore) {
                    if (trow && lane == 0) trow[5] = mg_now();
                    P::emit(ctx, dst, view, k, f, scratch, (q.io.flags & MG_FLAG_STREAM_OBS) != 0);
                    if (trow && lane == 0) trow[6] = mg_now();
                    continue;
                }
                if (pending) {
                    if (lane == 0) mg_bulk_wait_read();
                    ctx.sync();
                    P::restore(ctx, buf, view, k, *prev, scratch);
                    ctx.sync();
                }
                P::compose(ctx, buf, view, k, f, scratch);
                mg_fence_proxy_async();  // make this thread's patch visible to the async (TMA) proxy
                ctx.sync();
                if (lane == 0) {
                    if (trow) trow[5] = mg_now();
                    // rollout: the same address is rewritten ring_n steps later (final_obs: at the next done);
                    // the earlier bulk write must have been performed before this one is issued
                    if (T > 1) mg_bulk_wait_done(to_final ? 0 : q.ring_n - 1);
                    if (q.io.flags & MG_FLAG_STREAM_OBS)
                        mg_bulk_store_hint(dst, buf, MG_FRAME_BYTES, evict_first);
                    else
                        mg_bulk_store(dst, buf, MG_FRAME_BYTES);
                    mg_bulk_commit();
                    if (trow) trow[6] = mg_now();
                }
                pending = true;
                for (int w = lane; w < (int)(sizeof(Frame) / 4); w += kTeam)
                    reinterpret_cast<uint32_t*>(prev)[w] = reinterpret_cast<const uint32_t*>(&f)[w];
            }
            if constexpr (!kRollout) {
                // ground-truth render (render_mode="debug_rgb_array"): a cold path, painted from the frame that went to obs
                if (MG_UNLIKELY(q.io.debug_obs != nullptr) && nframes > 0)
                    P::emit_debug(ctx, q.io.debug_obs + (size_t)env * MG_DEBUG_BYTES, view, k, fr[nframes - 1], *st, scratch);
            }
            ctx.sync();  // job / frame slots are reused by the next step
        }
        uint64_t* gw = reinterpret_cast<uint64_t*>(state + (size_t)env * kRecordBytes);
#pragma unroll
        for (int j = 0; j < kPerLane; ++j)
            if (lane + kTeam * j < kWords) gw[lane + kTeam * j] = s64[lane + kTeam * j];
        ctx.sync();  // record slot is reused by the next env
        if constexpr (!kRollout && P::kHelperWarps > 0) P::after_writeback(ctx, wextra);  // (a pending path: post it now)
        env = nxt;
        have = have_next;
    }
    if (pending && lane == 0) mg_bulk_wait_read();  // shared memory must outlive the bulk read
    team_exit();
}

// ---- host side -----------------------------------------------------------------------------------
extern "C" int mg_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        mg_set_err("no usable CUDA device");
        return MG_ENODEV;
    }
    int v = 0;
    if (sm_count) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        *sm_count = v;
    }
    if (cc_major) {
        cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev);
        *cc_major = v;
    }
    if (cc_minor) {
        cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev);
        *cc_minor = v;
    }
    if (smem_optin) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        *smem_optin = (size_t)v;
    }
    return 0;
}

struct MgLaunchPlan {
    int ctas, wpc, stage_cap, helpers;
    size_t smem;
};

// grid geometry: one CTA per SM, `wpc` warps = frame buffers each. For rollouts (every warp runs all steps of
// its envs) the warp count that balances envs per warp best is chosen; for single steps, the most buffers.
// The plan of a (device, batch size, requested warps, helper switch) combination is computed once per thread and the
// kernel's shared-memory opt-in is raised once per device: a step launch does no driver query.
template <class P, bool kRollout>
static int plan_launch(int dev, int n, int wpc_req, int default_wpc, bool no_helpers, MgLaunchPlan* out) {
    struct Entry {
        int dev, n, wpc_req, no_helpers;
        MgLaunchPlan plan;
    };
    static thread_local Entry cache[4] = {{-1, 0, 0, 0, {}}, {-1, 0, 0, 0, {}}, {-1, 0, 0, 0, {}}, {-1, 0, 0, 0, {}}};
    static thread_local int cache_next = 0;
    for (int i = 0; i < 4; ++i)
        if (cache[i].dev == dev && cache[i].n == n && cache[i].wpc_req == wpc_req && cache[i].no_helpers == (int)no_helpers) {
            *out = cache[i].plan;
            return 0;
        }
    const bool balance = kRollout;
    const MgDevInfo* di = mg_dev_info(dev);
    if (!di) {
        mg_set_err("no usable CUDA device");
        return MG_ENODEV;
    }
    const int sms = di->sms, smem_max = di->smem_optin;
    constexpr int kSlot =
        (((int)sizeof(typename P::State) + (P::kDirectStore ? 2 : 3) * (int)sizeof(typename P::Frame) + (int)sizeof(MgFrameJob) + 16 + 127) / 128) * 128;
    const int per_warp = P::kFrameBuf + kSlot + P::kWarpExtra;
    int wmax = 10;
    while (wmax > 1 && wmax * 32 * P::kTeamWarps > P::kMaxThreads) wmax--;
    while (wmax > 1 && wmax * per_warp + P::kCtaExtra + MG_CTA_HDR_BYTES + P::kStageReserve > smem_max) wmax--;  // leave room to stage the assets
    int wpc = wpc_req ? wpc_req : default_wpc;
    if (wpc > wmax) wpc = wmax;
    // (policies with helper warps take every frame buffer they can get: their rollouts are scheduled dynamically
    // and bound by the generation capacity of teams + helpers, not by rounds of envs per warp)
    if (!wpc_req && balance && P::kHelperWarps == 0 && n > sms * wmax) {
        double best = -1.0;
        for (int w = wmax; w >= (wmax + 1) / 2; --w) {  // at least half the buffers stay in flight
            long warps = (long)sms * w;
            double eff = ((double)n / warps) / (double)((n + warps - 1) / warps);
            if (eff > best + 0.02) {
                best = eff;
                wpc = w;
            }
        }
    }
    int need = (n + sms - 1) / sms;  // spread small batches over all SMs
    if (wpc > need) wpc = need;
    if (wpc < 1) wpc = 1;
    int ctas = (n + wpc - 1) / wpc;
    if (ctas > sms) ctas = sms;
    size_t smem = (size_t)wpc * per_warp + P::kCtaExtra + MG_CTA_HDR_BYTES;
    // what is left of the SM's shared memory (up to 48 KB) stages the asset header + sprite pixel lists
    int stage_cap = smem_max - (int)smem;
    if (stage_cap > 48 * 1024) stage_cap = 48 * 1024;
    stage_cap &= ~127;
    smem += (size_t)stage_cap;
    // the kernel may use the device's whole opt-in shared memory: set once per device (the attribute is per device)
    static std::atomic<int> optin_set[MG_MAX_DEVICES];
    if (!optin_set[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(mg_env_kernel<P, kRollout>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        if (e != cudaSuccess) {
            mg_set_err(cudaGetErrorString(e));
            cudaGetLastError();
            return (int)e;
        }
        optin_set[dev].store(1, std::memory_order_release);
    }
    // helper warps fill the thread budget behind the teams (MG_FLAG_NO_HELPERS switches them off: measurement knob)
    int helpers = 0;
    if (P::kHelperWarps > 0 && !no_helpers) {
        helpers = P::kMaxThreads / (32 * P::kTeamWarps) - wpc;
        const int most = kRollout ? P::kRolloutHelpers : P::kHelperWarps;
        if (helpers > most) helpers = most;
        if (helpers < 0) helpers = 0;
    }
    out->helpers = helpers;
    out->ctas = ctas;
    out->wpc = wpc;
    out->stage_cap = stage_cap;
    out->smem = smem;
    Entry& slot = cache[cache_next];
    cache_next = (cache_next + 1) & 3;
    slot.dev = dev, slot.n = n, slot.wpc_req = wpc_req, slot.no_helpers = (int)no_helpers, slot.plan = *out;
    return 0;
}

template <class P, bool kRollout>
static int launch_one(const typename P::K& k, const MgKIO& q, const MgLaunchPlan& plan, const void* assets, void* state,
                      void* stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan.ctas);
    cfg.blockDim = dim3((plan.wpc + plan.helpers) * 32 * P::kTeamWarps);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = (cudaStream_t)stream;
    // programmatic dependent launch: this kernel's prologue (asset staging, template loads) may overlap
    // the tail of the previous kernel in the stream; it calls griddepcontrol.wait before touching env data
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (q.io.flags & MG_FLAG_NO_PDL) ? 0 : 1;  // debugging / measurement switch
    cudaError_t e = cudaLaunchKernelEx(&cfg, mg_env_kernel<P, kRollout>, k, q, (const uint8_t*)assets, (uint8_t*)state);
    if (e != cudaSuccess) {
        mg_set_err(cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    return 0;
}

static int check_common(const void* assets, const void* state, int n) {
    if (!assets || !state || n < 0) {
        mg_set_err("NULL pointer or negative n");
        return MG_EINVAL;
    }
    if (((uintptr_t)assets & 127) || ((uintptr_t)state & 7)) {
        mg_set_err("alignment: assets 128 B, state 8 B");
        return MG_EINVAL;
    }
    return 0;
}

// reset / single steps: one launch per mg_step_io
template <class P>
static int launch_env(const typename P::K& k, const void* assets, void* state, const uint64_t* rng4,
                      const mg_step_io* ios, int n_io, int n, int mode, int act_dims, void* stream, int default_wpc) {
    if (check_common(assets, state, n)) return MG_EINVAL;
    if (!ios || n_io < 1) {
        mg_set_err("empty io list");
        return MG_EINVAL;
    }
    for (int t = 0; t < n_io; ++t) {
        const mg_step_io* io = ios + t;
        if (!io->obs || (mode == MG_MODE_STEP && (!io->actions || !io->reward || !io->done || !io->info))) {
            mg_set_err("mg_step_io: obs (and for step: actions, reward, done, info) must be non-NULL");
            return MG_EINVAL;
        }
        if (((uintptr_t)io->obs & 15) || (io->final_obs && ((uintptr_t)io->final_obs & 15))) {
            mg_set_err("alignment: obs/final_obs must be 16-byte aligned");
            return MG_EINVAL;
        }
    }
    if (n == 0) return 0;
    MgDeviceGuard guard(state);
    if (!guard.ok) {
        mg_set_err("no usable CUDA device");
        return MG_ENODEV;
    }
    MgLaunchPlan plan;
    int rc = plan_launch<P, false>(guard.dev, n, (int)ios->warps_per_cta, default_wpc, (ios->flags & MG_FLAG_NO_HELPERS) != 0, &plan);
    if (rc) return rc;
    if (mode != MG_MODE_STEP) plan.helpers = 0;  // helper warps finish AUTO-resets; a reset launch does its own
    for (int t = 0; t < n_io; ++t) {
        MgKIO q;
        memset(&q, 0, sizeof(q));
        q.io = ios[t];
        q.rng4 = rng4;
        q.mode = mode;
        q.n = n;
        q.stage_cap = plan.stage_cap;
        q.act_dims = act_dims;
        q.n_steps = 1;
        q.ring_n = 1;
        q.n_helpers = plan.helpers;
        q.n_lane_helpers = plan.helpers < P::kRolloutHelpers ? plan.helpers : P::kRolloutHelpers;
        rc = launch_one<P, false>(k, q, plan, assets, state, stream);
        if (rc) return rc;
    }
    return 0;
}

// rollout: ONE launch runs all n_steps steps
template <class P>
static int launch_rollout(const typename P::K& k, const void* assets, void* state, const mg_rollout_io* r, int n,
                          int act_dims, void* stream, int default_wpc) {
    if (check_common(assets, state, n)) return MG_EINVAL;
    if (!r || !r->actions || !r->obs || !r->reward || !r->done || !r->info || r->n_steps < 1 || r->ring < 1) {
        mg_set_err("mg_rollout_io: actions, obs, reward, done, info must be non-NULL; n_steps, ring >= 1");
        return MG_EINVAL;
    }
    if (((uintptr_t)r->obs & 15) || (r->final_obs && ((uintptr_t)r->final_obs & 15))) {
        mg_set_err("alignment: obs/final_obs must be 16-byte aligned");
        return MG_EINVAL;
    }
    if (n == 0) return 0;
    MgDeviceGuard guard(state);
    if (!guard.ok) {
        mg_set_err("no usable CUDA device");
        return MG_ENODEV;
    }
    MgLaunchPlan plan;
    int rc = plan_launch<P, true>(guard.dev, n, (int)r->warps_per_cta, default_wpc, (r->flags & MG_FLAG_NO_HELPERS) != 0, &plan);
    if (rc) return rc;
    MgKIO q;
    memset(&q, 0, sizeof(q));
    q.io.actions = r->actions;
    q.io.obs = r->obs;
    q.io.final_obs = r->final_obs;
    q.io.reward = r->reward;
    q.io.done = r->done;
    q.io.info = r->info;
    q.io.ground_truth = r->ground_truth;
    q.io.trace = r->trace;
    q.io.vector_obs = r->vector_obs;
    q.sched = r->sched;
    q.n_helpers = plan.helpers;
    q.n_lane_helpers = plan.helpers;
    q.io.flags = r->flags | MG_FLAG_AUTORESET;  // a rollout without auto-reset would only idle finished envs
    q.mode = MG_MODE_STEP;
    q.n = n;
    q.stage_cap = plan.stage_cap;
    q.act_dims = act_dims;
    q.n_steps = r->n_steps;
    q.ring_n = r->ring;
    q.act_stride = (int64_t)n * act_dims;
    q.obs_stride = (int64_t)n * MG_FRAME_BYTES;
    q.rew_stride = n;
    q.done_stride = n;
    q.info_stride = r->info_per_step ? (int64_t)n * MG_INFO_COLS : 0;
    return launch_one<P, true>(k, q, plan, assets, state, stream);
}

// Option validation + derivation of the kernel's option block happen once per option set and thread: a step call
// whose mg_*_params bytes equal the previous call's reuses the derived block.
template <class Params, class K>
static int mg_cached_k(const Params* p, int (*check)(const Params*), void (*derive)(const Params*, K*), const K** out) {
    static thread_local Params last;
    static thread_local K k;
    static thread_local bool valid = false;
    if (!p) return check(p);
    if (!valid || memcmp(&last, p, sizeof(Params)) != 0) {
        valid = false;
        if (check(p)) return MG_EINVAL;
        derive(p, &k);
        memcpy(&last, p, sizeof(Params));
        valid = true;
    }
    *out = &k;
    return 0;
}

static int mm_act_dims(const mg_mm_params* p) { return mg_mm_is_grid(p->variant) ? 1 : 2; }
static int mp_act_dims(const mg_mp_params* p) { return p->variant == MG_MP_GRID ? 1 : 2; }

extern "C" int mg_mm_reset(const mg_mm_params* p, const void* assets, void* state, const uint64_t* rng4,
                           const mg_step_io* io, int n, void* stream) {
    const MgMMK* kp;
    if (mg_cached_k(p, mg_mm_check_params, mg_mm_derive, &kp)) return MG_EINVAL;
    const MgMMK& k = *kp;
    return launch_env<MMPolicy>(k, assets, state, rng4, io, 1, n, MG_MODE_RESET, mm_act_dims(p), stream, 10);
}
extern "C" int mg_mm_step(const mg_mm_params* p, const void* assets, void* state, const mg_step_io* io, int n,
                          void* stream) {
    return mg_mm_step_many(p, assets, state, io, 1, n, stream);
}
extern "C" int mg_mm_step_many(const mg_mm_params* p, const void* assets, void* state, const mg_step_io* ios,
                               int n_steps, int n, void* stream) {
    const MgMMK* kp;
    if (mg_cached_k(p, mg_mm_check_params, mg_mm_derive, &kp)) return MG_EINVAL;
    const MgMMK& k = *kp;
    return launch_env<MMPolicy>(k, assets, state, nullptr, ios, n_steps, n, MG_MODE_STEP, mm_act_dims(p), stream, 10);
}
extern "C" int mg_mm_rollout(const mg_mm_params* p, const void* assets, void* state, const mg_rollout_io* r, int n,
                             void* stream) {
    const MgMMK* kp;
    if (mg_cached_k(p, mg_mm_check_params, mg_mm_derive, &kp)) return MG_EINVAL;
    const MgMMK& k = *kp;
    return launch_rollout<MMPolicy>(k, assets, state, r, n, mm_act_dims(p), stream, 10);
}

extern "C" int mg_mp_reset(const mg_mp_params* p, const void* assets, void* state, const uint64_t* rng4,
                           const mg_step_io* io, int n, void* stream) {
    const MgMPK* kp;
    if (mg_cached_k(p, mg_mp_check_params, mg_mp_derive, &kp)) return MG_EINVAL;
    const MgMPK& k = *kp;
    return launch_env<MPPolicy>(k, assets, state, rng4, io, 1, n, MG_MODE_RESET, mp_act_dims(p), stream, 10);
}
extern "C" int mg_mp_step(const mg_mp_params* p, const void* assets, void* state, const mg_step_io* io, int n,
                          void* stream) {
    return mg_mp_step_many(p, assets, state, io, 1, n, stream);
}
extern "C" int mg_mp_step_many(const mg_mp_params* p, const void* assets, void* state, const mg_step_io* ios,
                               int n_steps, int n, void* stream) {
    const MgMPK* kp;
    if (mg_cached_k(p, mg_mp_check_params, mg_mp_derive, &kp)) return MG_EINVAL;
    const MgMPK& k = *kp;
    return launch_env<MPPolicy>(k, assets, state, nullptr, ios, n_steps, n, MG_MODE_STEP, mp_act_dims(p), stream, 10);
}
extern "C" int mg_mp_rollout(const mg_mp_params* p, const void* assets, void* state, const mg_rollout_io* r, int n,
                             void* stream) {
    const MgMPK* kp;
    if (mg_cached_k(p, mg_mp_check_params, mg_mp_derive, &kp)) return MG_EINVAL;
    const MgMPK& k = *kp;
    return launch_rollout<MPPolicy>(k, assets, state, r, n, mp_act_dims(p), stream, 10);
}

extern "C" int mg_ep_reset(const mg_ep_params* p, const void* assets, void* state, const uint64_t* rng4,
                           const mg_step_io* io, int n, void* stream) {
    const MgEPK* kp;
    if (mg_cached_k(p, mg_ep_check_params, mg_ep_derive, &kp)) return MG_EINVAL;
    const MgEPK& k = *kp;
    return launch_env<EPPolicy>(k, assets, state, rng4, io, 1, n, MG_MODE_RESET, 1, stream, 10);
}
extern "C" int mg_ep_step(const mg_ep_params* p, const void* assets, void* state, const mg_step_io* io, int n,
                          void* stream) {
    return mg_ep_step_many(p, assets, state, io, 1, n, stream);
}
extern "C" int mg_ep_step_many(const mg_ep_params* p, const void* assets, void* state, const mg_step_io* ios,
                               int n_steps, int n, void* stream) {
    const MgEPK* kp;
    if (mg_cached_k(p, mg_ep_check_params, mg_ep_derive, &kp)) return MG_EINVAL;
    const MgEPK& k = *kp;
    return launch_env<EPPolicy>(k, assets, state, nullptr, ios, n_steps, n, MG_MODE_STEP, 1, stream, 10);
}
extern "C" int mg_ep_rollout(const mg_ep_params* p, const void* assets, void* state, const mg_rollout_io* r, int n,
                             void* stream) {
    const MgEPK* kp;
    if (mg_cached_k(p, mg_ep_check_params, mg_ep_derive, &kp)) return MG_EINVAL;
    const MgEPK& k = *kp;
    return launch_rollout<EPPolicy>(k, assets, state, r, n, 1, stream, 10);
}

// the endless variant is its own kernel instantiation: neither carries the other's code (I-cache footprint)
#define MG_SS_DISPATCH(fn, ...)                                     \
    if (k.endless) return fn<ESSPolicy>(__VA_ARGS__);               \
    return fn<SSPolicy>(__VA_ARGS__);

extern "C" int mg_ss_reset(const mg_ss_params* p, const void* assets, void* state, const uint64_t* rng4,
                           const mg_step_io* io, int n, void* stream) {
    const MgSSK* kp;
    if (mg_cached_k(p, mg_ss_check_params, mg_ss_derive, &kp)) return MG_EINVAL;
    const MgSSK& k = *kp;
    MG_SS_DISPATCH(launch_env, k, assets, state, rng4, io, 1, n, MG_MODE_RESET, 2, stream, 10);
}
extern "C" int mg_ss_step(const mg_ss_params* p, const void* assets, void* state, const mg_step_io* io, int n,
                          void* stream) {
    return mg_ss_step_many(p, assets, state, io, 1, n, stream);
}
extern "C" int mg_ss_step_many(const mg_ss_params* p, const void* assets, void* state, const mg_step_io* ios,
                               int n_steps, int n, void* stream) {
    const MgSSK* kp;
    if (mg_cached_k(p, mg_ss_check_params, mg_ss_derive, &kp)) return MG_EINVAL;
    const MgSSK& k = *kp;
    MG_SS_DISPATCH(launch_env, k, assets, state, nullptr, ios, n_steps, n, MG_MODE_STEP, 2, stream, 10);
}
extern "C" int mg_ss_rollout(const mg_ss_params* p, const void* assets, void* state, const mg_rollout_io* r, int n,
                             void* stream) {
    const MgSSK* kp;
    if (mg_cached_k(p, mg_ss_check_params, mg_ss_derive, &kp)) return MG_EINVAL;
    const MgSSK& k = *kp;
    MG_SS_DISPATCH(launch_rollout, k, assets, state, r, n, 2, stream, 10);
}
