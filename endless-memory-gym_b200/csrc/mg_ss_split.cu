// mg_ss_split.cu — Searing Spotlights (and the endless variant) as TWO kernels per step / rollout chunk.
//
// Why: in the fused kernel (mg_kernels.cu) one warp of every two-warp team spends ~40 % of the frame time in the
// scalar state machine (searing_spotlights.py:452-562 on ONE lane: 31 idle lanes at the issue-latency limit), the
// raster warps carry the logic's 128 registers, and 16 warps per SM is all the register file holds
// (profiles/r01b_summary.md: IPC 1.29, 24 % warp occupancy, DRAM 33 %).
//
//   mg_ss_logic_kernel   ONE LANE PER ENV: thread i copies env i's record into thread-local memory (L1-resident: the
//                        kernel uses no shared memory), runs the chunk's steps with the SAME logic functions the fused
//                        kernel uses (mg_ss_logic.h, instantiated for a one-lane context — auto-resets, the grid
//                        sampler and the endless variant's coin respawn included), and writes one 256-byte frame
//                        DESCRIPTION per env-step (agent pose, bar, coins, exit, the discs drawn this step) plus
//                        rewards, dones and episode statistics. 8192 envs = 256 warps, all in flight at once.
//   mg_ss_raster_kernel  persistent, one CTA per SM, 9 teams of TW warps; a team owns one env at a time and turns
//                        its descriptions into frames with the incremental rasteriser of mg_frame.h (base buffer
//                        patch, lit bitmap, masked 128-bit copy straight to the observation). No logic registers, no
//                        logic instructions: more warps per SM, and nothing scalar on the frame's critical path.
//
// On a side stream the logic kernel of chunk c + 1 runs WHILE the raster kernel of chunk c streams frames (it needs
// no shared memory and two warps per SM), so in a rollout only the first chunk's logic is exposed.
// Results are bit-identical to the fused kernel (same functions, same order of draws); descriptions live in a
// caller-provided scratch buffer (mg_ss_scratch_bytes).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <atomic>

#include "../../include/memgym_b200.h"
#include "mg_common.h"
#include "mg_device.cuh"
#include "mg_device_host.h"
#include "mg_frame.h"
#include "mg_ss_logic.h"
#include "mg_params_host.h"
#include "mg_policy.h"
#include "mg_ss_split.h"

namespace {

struct alignas(16) MgSSDesc {
    MgSSFrame f;
    uint32_t flags;  // bit 0: no frame for this env-step (the env is not live: never reset, or done without auto-reset)
};
static_assert(sizeof(MgSSDesc) == MG_SS_DESC_BYTES, "description stride");

// the one-lane execution context of the logic kernel: every "cooperative" routine of mg_ss_logic.h degenerates to a
// plain loop (exactly what the CPU harness of the test-suite instantiates)
struct MgLaneCtx {
    static constexpr int kLanes = 1;
    int lane = 0, nl = 1;
    __device__ __forceinline__ void sync() {}
    __device__ __forceinline__ void stamp(int) {}
    __device__ __forceinline__ uint32_t reduce_or(uint32_t v) { return v; }
    __device__ __forceinline__ uint32_t reduce_min(uint32_t v) { return v; }
    __device__ __forceinline__ uint32_t shfl(uint32_t v, int) { return v; }
    __device__ __forceinline__ void atomic_or(uint32_t* p, uint32_t v) { *p |= v; }
};

struct MgSplitArgs {
    MgSSDesc* desc;       // [steps][n]
    MgSSDesc* fin_desc;   // [n]: terminal frame of the env's last finished episode of this chunk (final_obs)
    uint32_t* fin_flag;   // [n]
    int32_t t0;           // first step of this chunk within the call (indexes actions / reward / done / info / obs slot)
    int32_t steps;        // steps in this chunk
};

// ---- logic: one lane per env ------------------------------------------------------------------------------------------
template <bool kEndless>
__global__ void __launch_bounds__(64) mg_ss_logic_kernel(const __grid_constant__ MgSSK k, const __grid_constant__ MgKIO q,
                                                        const uint8_t* __restrict__ assets, uint8_t* __restrict__ state,
                                                        const MgSplitArgs sa) {
    const int env = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (env >= q.n) return;
    const MgAssetView view = mg_asset_view(assets);
    const double* trig = (const double*)(view.base + view.hdr->trig_off);
    MgLaneCtx c;
    alignas(16) MgSSState s;  // thread-local copy of the record (local memory, L1-resident)
    {
        const uint4* g = reinterpret_cast<const uint4*>(state + (size_t)env * sizeof(MgSSState));
        uint4* l = reinterpret_cast<uint4*>(&s);
#pragma unroll 4
        for (int w = 0; w < (int)(sizeof(MgSSState) / 16); ++w) l[w] = g[w];
    }
    MgSSStepTmp tmp;
    MgSSSampler g;
    MgSSResetTmp rt;
    MgSSDesc d;
    double* const gt = q.io.ground_truth ? q.io.ground_truth + MG_GT_COLS * (size_t)env : nullptr;
    const bool autoreset = (q.io.flags & MG_FLAG_AUTORESET) != 0;
    for (int i = 0; i < sa.steps; ++i) {
        const int t = sa.t0 + i;
        MgSSDesc* out = sa.desc + (size_t)i * q.n + env;
        double* const reward_out = q.io.reward + (size_t)t * q.rew_stride + env;
        uint8_t* const done_out = q.io.done + (size_t)t * q.done_stride + env;
        if (!s.live) {  // finished (or never reset): the step is ignored
            *reward_out = 0.0;
            *done_out = 1;
            out->flags = 1u;
            continue;
        }
        int64_t a0 = 0, a1 = 0;
        {
            const int64_t* a = q.io.actions + (size_t)t * q.act_stride + (size_t)env * 2;
            const longlong2 v = *reinterpret_cast<const longlong2*>(a);  // int64[n][2]: 16-byte aligned rows
            a0 = v.x, a1 = v.y;
        }
        double* const info_out = q.io.info + (size_t)t * q.info_stride + MG_INFO_COLS * (size_t)env;
        tmp.action = 2;
        tmp.post = 0;
        mg_ss_step_a(k, s, (int)a0, (int)a1, tmp);
        mg_ss_step_b(c, k, trig, s, tmp, d.f);
        double reward;
        bool done;
        if constexpr (kEndless) {
            MgESSCarry carry;
            if (mg_ess_step_c1(k, s, tmp, carry)) {
                mg_ess_spawn_coin(c, k.clamp_off, s, g, rt, true);
                s.scene_last = ~0u;  // the coin moved: the rasteriser must rebuild its base buffer
            }
            done = mg_ess_step_c2(k, s, tmp, carry, reward, info_out, d.f, gt);
        } else {
            done = mg_ss_step_c(k, s, tmp, reward, info_out, d.f);
        }
        *reward_out = reward;
        *done_out = done ? 1 : 0;
        if (done && autoreset) {
            if (q.io.final_obs) {  // terminal frame -> final_obs (the last finished episode wins), next episode's first -> obs
                d.flags = 0u;
                sa.fin_desc[env] = d;
                sa.fin_flag[env] = 1u;
            }
            if constexpr (kEndless)
                mg_ess_reset_env(c, k, s, d.f, g, rt, gt);
            else
                mg_ss_reset_env(c, k, s, d.f, g, rt);
        }
        d.flags = 0u;
        *out = d;
    }
    {
        uint4* gw = reinterpret_cast<uint4*>(state + (size_t)env * sizeof(MgSSState));
        const uint4* l = reinterpret_cast<const uint4*>(&s);
#pragma unroll 4
        for (int w = 0; w < (int)(sizeof(MgSSState) / 16); ++w) gw[w] = l[w];
    }
}

// ---- raster: persistent teams turn descriptions into frames -------------------------------------------------------------
constexpr int kWarpMemBytes = ((int)sizeof(MgSSWarpMem) + 15) / 16 * 16;  // (the descriptions behind it are read 16 bytes at a time)
constexpr int kRasterTeamBytes = (kWarpMemBytes + 2 * MG_SS_DESC_BYTES + 16 + 127) / 128 * 128;
constexpr int kRasterCtaBytes = ((int)sizeof(MgSSCtaTables) + 127) / 128 * 128 + MG_CTA_HDR_BYTES;

template <int TW, bool kEndless>
__global__ void __launch_bounds__(TW * 32 * 9 > 1024 ? 1024 : TW * 32 * 9, 1)
    mg_ss_raster_kernel(const __grid_constant__ MgSSK k, const __grid_constant__ MgKIO q, const uint8_t* __restrict__ assets,
                        const MgSplitArgs sa, const int nteams, const int stage_cap) {
    typedef SSPolicyT<2, kEndless> P;  // (draw_k: the layer order of a frame)
    constexpr int kTeam = 32 * TW;
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x % kTeam, team = threadIdx.x / kTeam;
    uint8_t* const tmem = smem + (size_t)team * kRasterTeamBytes;
    MgSSWarpMem* const wm = reinterpret_cast<MgSSWarpMem*>(tmem);
    MgSSDesc* const dbuf = reinterpret_cast<MgSSDesc*>(tmem + kWarpMemBytes);
    uint32_t* const grab = reinterpret_cast<uint32_t*>(tmem + kWarpMemBytes + 2 * MG_SS_DESC_BYTES);
    uint8_t* const cextra = smem + (size_t)nteams * kRasterTeamBytes;
    MgSSCtaTables* const tables = reinterpret_cast<MgSSCtaTables*>(cextra);
    MgCtaHdr* const hdr = reinterpret_cast<MgCtaHdr*>(cextra + kRasterCtaBytes - MG_CTA_HDR_BYTES);
    uint8_t* const stage = cextra + kRasterCtaBytes;

    // prologue: constant assets only (overlaps the previous kernel's tail under programmatic dependent launch)
    MgAssetView view = mg_asset_view(assets);
    // the staged prefix ends where the trig table starts (only the logic reads it): header, sprite lists, dense bitmaps,
    // disc / outline tables
    const int want = view.hdr->trig_off ? (int)((view.hdr->trig_off + 15u) & ~15u) : (int)view.hdr->stage_bytes;
    const int stage_len = want <= stage_cap ? want : 0;
    if (threadIdx.x == 0) {
        hdr->next = (uint32_t)nteams;
        const uint32_t sbar = mg_smem_u32(&hdr->stage_bar);
        mg_mbar_init(sbar, 1);
        if (stage_len > 0) {
            mg_mbar_expect_tx(sbar, (uint32_t)stage_len);
            mg_bulk_load(stage, assets, (uint32_t)stage_len, sbar);
        }
    }
    mg_ss_tables_init(tables, threadIdx.x, blockDim.x);
    if (lane == 0) {
        wm->desc.valid = 0;
        wm->desc.lut_alpha = -1;
        wm->desc.env = -1;
    }
    __syncthreads();
    if (stage_len > 0) {
        mg_mbar_wait(mg_smem_u32(&hdr->stage_bar), 0u);
        view = mg_asset_view2(stage, assets);  // (the trig table behind stage_len is not used by the rasteriser)
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the logic kernel's descriptions are visible

    MgWarpCtx ctx{lane, kTeam, 0u, 0u, mg_policy_evict_first(), nullptr, team + 1};
    const bool stream = (q.io.flags & MG_FLAG_STREAM_OBS) != 0;
    const int G = gridDim.x;
    int env = (int)blockIdx.x + team * G;
    while (env < q.n) {
        if (lane == 0 && wm->desc.env != env) {  // a foreign scene in the base buffer
            wm->desc.valid = 0;
            wm->desc.env = env;
        }
        // description of step 0 -> buffer 0 (16 lanes x 16 bytes)
        if (lane < MG_SS_DESC_BYTES / 16)
            reinterpret_cast<uint4*>(&dbuf[0])[lane] = reinterpret_cast<const uint4*>(sa.desc + env)[lane];
        ctx.sync();
        for (int i = 0; i < sa.steps; ++i) {
            uint4 nextv = make_uint4(0u, 0u, 0u, 0u);
            const bool pre = i + 1 < sa.steps && lane < MG_SS_DESC_BYTES / 16;
            if (pre) nextv = reinterpret_cast<const uint4*>(sa.desc + (size_t)(i + 1) * q.n + env)[lane];  // in flight during the frame
            const MgSSDesc& d = dbuf[i & 1];
            if (!(d.flags & 1u)) {
                const int t = sa.t0 + i;
                uint8_t* dst = q.io.obs + (size_t)(t % q.ring_n) * q.obs_stride + (size_t)env * MG_FRAME_BYTES;
                mg_ss_emit<kEndless>(ctx, dst, view, P::draw_k(k, d.f), d.f, wm, tables, stream);
            }
            if (pre) reinterpret_cast<uint4*>(&dbuf[(i + 1) & 1])[lane] = nextv;
            ctx.sync();
        }
        if (q.io.final_obs && sa.fin_flag[env]) {
            // the terminal frame of the last finished episode: a scene of its own (full rebuild before and after)
            if (lane < MG_SS_DESC_BYTES / 16)
                reinterpret_cast<uint4*>(&dbuf[0])[lane] = reinterpret_cast<const uint4*>(sa.fin_desc + env)[lane];
            if (lane == 0) wm->desc.valid = 0;
            ctx.sync();
            mg_ss_emit<kEndless>(ctx, q.io.final_obs + (size_t)env * MG_FRAME_BYTES, view, P::draw_k(k, dbuf[0].f), dbuf[0].f, wm,
                                 tables, false);
            ctx.sync();
            if (lane == 0) wm->desc.valid = 0;
        }
        if (lane == 0) *grab = (uint32_t)blockIdx.x + atomicAdd(&hdr->next, 1u) * (uint32_t)G;
        ctx.sync();
        env = (int)*grab;
        ctx.sync();
    }
}

// the logic kernel resets fin_flag itself? No: the launcher clears it once per call, before the first chunk.
__global__ void mg_ss_clear_flags(uint32_t* flags, int n) {
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (i < n) flags[i] = 0u;
}

template <class F, class... Args>
cudaError_t launch_pdl(F kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

struct RasterPlan {
    int teams, tw, ctas, stage_cap;
    size_t smem;
};

template <int TW, bool kEndless>
cudaError_t launch_raster(const MgSSK& k, const MgKIO& q, const uint8_t* assets, const MgSplitArgs& sa, const RasterPlan& p,
                          int dev, cudaStream_t stream, bool pdl) {
    static std::atomic<int> optin_set[MG_MAX_DEVICES];
    if (!optin_set[dev].load(std::memory_order_acquire)) {
        const MgDevInfo* di = mg_dev_info(dev);
        cudaError_t e = cudaFuncSetAttribute(mg_ss_raster_kernel<TW, kEndless>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             di->smem_optin);
        if (e != cudaSuccess) return e;
        optin_set[dev].store(1, std::memory_order_release);
    }
    return launch_pdl(mg_ss_raster_kernel<TW, kEndless>, dim3(p.ctas), dim3(p.teams * TW * 32), p.smem, stream, pdl, k, q, assets,
                      sa, p.teams, p.stage_cap);
}

}  // namespace

extern "C" size_t mg_ss_scratch_bytes(int n, int n_steps) {
    if (n <= 0 || n_steps <= 0) return 0;
    int chunk = n_steps < MG_SS_SPLIT_CHUNK ? n_steps : MG_SS_SPLIT_CHUNK;
    // two description buffers of one chunk each (ping-pong between the logic and the raster kernel), the terminal
    // descriptions, their flags
    return (size_t)(2 * chunk + 1) * (size_t)n * MG_SS_DESC_BYTES + (size_t)n * 4 + 256;
}

// One call = n_steps steps of the whole batch: per chunk a logic launch and a raster launch. Returns 0, a CUDA error,
// or MG_EINVAL. `q` carries the call's I/O (strides as in the fused launchers).
int mg_ss_split_launch(const MgSSK& k, const MgKIO& q, const void* assets, void* state, void* scratch, size_t scratch_bytes,
                       int tw_req, bool no_pdl, void* stream_v) {
    const int n = q.n, T = q.n_steps;
    if (scratch_bytes < mg_ss_scratch_bytes(n, T) || ((uintptr_t)scratch & 15)) {
        mg_set_err("searing spotlights: scratch buffer too small (mg_ss_scratch_bytes) or not 16-byte aligned");
        return MG_EINVAL;
    }
    if (((uintptr_t)q.io.actions & 15)) {
        mg_set_err("searing spotlights: actions must be 16-byte aligned");
        return MG_EINVAL;
    }
    MgDeviceGuard guard(state);
    const MgDevInfo* di = guard.ok ? mg_dev_info(guard.dev) : nullptr;
    if (!di) {
        mg_set_err("no usable CUDA device");
        return MG_ENODEV;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    const int chunk = T < MG_SS_SPLIT_CHUNK ? T : MG_SS_SPLIT_CHUNK;
    uint8_t* base = (uint8_t*)scratch;
    MgSSDesc* bufs[2] = {(MgSSDesc*)base, (MgSSDesc*)(base + (size_t)chunk * n * MG_SS_DESC_BYTES)};
    MgSSDesc* fin = (MgSSDesc*)(base + (size_t)2 * chunk * n * MG_SS_DESC_BYTES);
    uint32_t* flags = (uint32_t*)(base + (size_t)(2 * chunk + 1) * n * MG_SS_DESC_BYTES);

    // raster geometry: as many teams as shared memory holds (assets staged without the trig table), TW warps each
    RasterPlan p;
    p.tw = tw_req == 2 || tw_req == 3 || tw_req == 4 ? tw_req : MG_SS_SPLIT_TW;
    const int avail = di->smem_optin - kRasterCtaBytes;
    int teams = (avail - MG_SS_RASTER_STAGE_MIN) / kRasterTeamBytes;
    if (teams > 9) teams = 9;
    while (teams > 1 && teams * p.tw * 32 > 1024) teams--;
    if (teams < 1) teams = 1;
    int need = (n + di->sms - 1) / di->sms;
    if (teams > need) teams = need;
    p.teams = teams;
    p.ctas = (n + teams - 1) / teams;
    if (p.ctas > di->sms) p.ctas = di->sms;
    p.stage_cap = (avail - teams * kRasterTeamBytes) & ~127;
    if (p.stage_cap > 48 * 1024) p.stage_cap = 48 * 1024;
    if (p.stage_cap < 0) p.stage_cap = 0;
    p.smem = (size_t)teams * kRasterTeamBytes + kRasterCtaBytes + (size_t)p.stage_cap;

    cudaError_t e = cudaSuccess;
    const bool pdl = !no_pdl;
    if (q.io.final_obs) e = launch_pdl(mg_ss_clear_flags, dim3((n + 255) / 256), dim3(256), 0, stream, pdl, flags, n);
    for (int t0 = 0, c = 0; t0 < T && e == cudaSuccess; t0 += chunk, ++c) {
        MgSplitArgs sa;
        sa.desc = bufs[c & 1];
        sa.fin_desc = fin;
        sa.fin_flag = flags;
        sa.t0 = t0;
        sa.steps = T - t0 < chunk ? T - t0 : chunk;
        const dim3 lgrid((n + 63) / 64), lblock(64);
        if (k.endless)
            e = launch_pdl(mg_ss_logic_kernel<true>, lgrid, lblock, 0, stream, pdl, k, q, (const uint8_t*)assets, (uint8_t*)state, sa);
        else
            e = launch_pdl(mg_ss_logic_kernel<false>, lgrid, lblock, 0, stream, pdl, k, q, (const uint8_t*)assets, (uint8_t*)state, sa);
        if (e != cudaSuccess) break;
#define MG_RASTER(TWV)                                                                                             \
    e = k.endless ? launch_raster<TWV, true>(k, q, (const uint8_t*)assets, sa, p, guard.dev, stream, pdl)          \
                  : launch_raster<TWV, false>(k, q, (const uint8_t*)assets, sa, p, guard.dev, stream, pdl)
        if (p.tw == 2) {
            MG_RASTER(2);
        } else if (p.tw == 3) {
            MG_RASTER(3);
        } else {
            MG_RASTER(4);
        }
#undef MG_RASTER
    }
    if (e != cudaSuccess) {
        mg_set_err(cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    return 0;
}
