// mg_kernels.cu — the sm_100a reset/step kernels and their C-ABI launchers.
//
// One launch per env type does, for every env of the batch: kinematics + tile collision, the env
// state machine (incl. in-kernel PCG64 draws for auto-reset), and the rasterisation of the
// 84x84x3 observation. ONE WARP OWNS ONE ENV:
//   * the env's state record (256 B) is loaded/stored with one coalesced 8-byte access per lane
//     into a per-warp shared-memory slot; lane 0 runs the (branchy, scalar) state machine there;
//   * each warp keeps one 21168-byte frame buffer in shared memory that holds the clean background
//     template between envs. A frame = patch a few hundred sprite bytes, then ONE
//     `cp.async.bulk.global.shared::cta` (TMA bulk store, SASS UBLKCP) streams the whole frame to
//     HBM as full 128-byte lines while the warp already runs the next env's state machine; when the
//     bulk read of the buffer is done (`cp.async.bulk.wait_group.read`) the patch is undone.
//   * background templates are brought in with `cp.async.bulk.shared::cluster.global` + mbarrier.
// The grid is one CTA per SM (persistent over the env batch), up to 10 warps = 10 frame buffers
// (212 KB of the 227 KB shared memory) so ~200 KB of stores are in flight per SM.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/memgym_b200.h"
#include "mg_common.h"
#include "mg_frame.h"
#include "mg_mm_logic.h"
#include "mg_mp_logic.h"
#include "mg_policy.h"
#include "mg_params_host.h"

#define MG_BUF_STRIDE 21248  // MG_FRAME_BYTES rounded up to 128
#define MG_CTL_BYTES 128
// ---- PTX wrappers (TMA bulk copies, mbarrier, proxy fence) ---------------------------------------
__device__ __forceinline__ uint32_t mg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mg_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mg_bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gdst)),
                 "r"(mg_smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mg_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void mg_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
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

// execution context of mg_frame.h on the device: the 32 lanes of the warp that owns the env
struct MgWarpCtx {
    int lane, nl;
    uint32_t bar, phase;
    __device__ __forceinline__ void sync() { __syncwarp(); }
    // whole-frame template -> this warp's frame buffer, by TMA
    __device__ __forceinline__ void copy_frame(uint8_t* dst, const uint8_t* src) {
        mg_fence_proxy_async();  // earlier generic-proxy accesses to dst are ordered before the async write
        __syncwarp();
        if (lane == 0) {
            mg_mbar_expect_tx(bar, MG_FRAME_BYTES);
            mg_bulk_load(dst, src, MG_FRAME_BYTES, bar);
        }
        mg_mbar_wait(bar, phase);
        phase ^= 1u;
    }
};

// ---- the kernel ----------------------------------------------------------------------------------
template <class P>
__global__ void __launch_bounds__(320, 1)
    mg_env_kernel(const __grid_constant__ typename P::K k, const __grid_constant__ MgKIO q,
                  const uint8_t* __restrict__ assets, uint8_t* __restrict__ state) {
    typedef typename P::State State;
    typedef typename P::Frame Frame;
    constexpr int kStateBytes = (int)sizeof(State);
    constexpr int kWords = kStateBytes / 8;           // 8-byte words per record
    constexpr int kPerLane = (kWords + 31) / 32;       // words each lane moves
    constexpr int kSlot = ((kStateBytes + 2 * (int)sizeof(Frame) + (int)sizeof(MgFrameJob) + 16 + 127) / 128) * 128;
    extern __shared__ __align__(128) uint8_t smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint8_t* buf = smem + (size_t)warp * MG_BUF_STRIDE;
    uint8_t* slot = smem + (size_t)nwarps * MG_BUF_STRIDE + (size_t)warp * kSlot;
    uint8_t* stage = smem + (size_t)nwarps * (MG_BUF_STRIDE + kSlot);
    State* st = reinterpret_cast<State*>(slot);
    Frame* fr = reinterpret_cast<Frame*>(slot + kStateBytes);
    MgFrameJob* job = reinterpret_cast<MgFrameJob*>(slot + kStateBytes + 2 * sizeof(Frame));
    uint64_t* bar = reinterpret_cast<uint64_t*>(slot + kSlot - 8);

    // ---- prologue: touches only the (constant) asset blob, so under programmatic dependent launch it
    // overlaps the tail of the previous step's kernel ----------------------------------------------------
    MgAssetView view = mg_asset_view(assets);
    MgWarpCtx ctx{lane, 32, mg_smem_u32(bar), 0u};
    if (lane == 0) mg_mbar_init(ctx.bar, 1);
    __syncwarp();
    const int total_warps = gridDim.x * nwarps;
    int env = blockIdx.x * nwarps + warp;
    const bool active = env < q.n;
    if (active && lane == 0) {  // background template -> this warp's frame buffer (TMA, L2-resident source)
        mg_mbar_expect_tx(ctx.bar, MG_FRAME_BYTES);
        mg_bulk_load(buf, view.tmpl[0], MG_FRAME_BYTES, ctx.bar);
    }
    const int hw = (int)(sizeof(MgAssetHeader) / 4), ne = (int)view.hdr->n_entries;
    if ((hw + ne) * 4 <= q.stage_cap) {  // sprite pixel lists + palette -> shared memory: compose/un-patch never wait on L2
        const uint32_t* src = reinterpret_cast<const uint32_t*>(assets);
        uint32_t* dst = reinterpret_cast<uint32_t*>(stage);
        const uint32_t* esrc = view.entries;
        for (int i = threadIdx.x; i < hw; i += blockDim.x) dst[i] = src[i];
        for (int i = threadIdx.x; i < ne; i += blockDim.x) dst[hw + i] = esrc[i];
        __syncthreads();
        view.hdr = reinterpret_cast<const MgAssetHeader*>(stage);
        view.entries = dst + hw;
        view.palette = view.hdr->palette;
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // previous kernel's state / obs writes are visible
    if (!active) return;

    // ---- first record + actions in flight while the template lands ------------------------------------
    uint64_t pre[kPerLane];
    int64_t act0 = 0, act1 = 0;
    {
        const uint64_t* g = reinterpret_cast<const uint64_t*>(state + (size_t)env * kStateBytes);
#pragma unroll
        for (int j = 0; j < kPerLane; ++j)
            if (lane + 32 * j < kWords) pre[j] = g[lane + 32 * j];
        if (lane == 0 && q.mode == MG_MODE_STEP) {
            act0 = q.io.actions[(size_t)env * q.act_dims];
            if (q.act_dims > 1) act1 = q.io.actions[(size_t)env * q.act_dims + 1];
        }
    }
    mg_mbar_wait(ctx.bar, ctx.phase);
    ctx.phase ^= 1u;

    bool pending = false;  // a bulk store is still reading `buf`; `prev` is the patch to undo afterwards
    Frame prev;
    for (; env < q.n; env += total_warps) {
        uint64_t* s64 = reinterpret_cast<uint64_t*>(st);
#pragma unroll
        for (int j = 0; j < kPerLane; ++j)
            if (lane + 32 * j < kWords) s64[lane + 32 * j] = pre[j];
        const int a0 = (int)act0, a1 = (int)act1;
        __syncwarp();
        // prefetch the next env's record and actions: their latency hides behind this env's work
        const int nxt = env + total_warps;
        if (nxt < q.n) {
            const uint64_t* g = reinterpret_cast<const uint64_t*>(state + (size_t)nxt * kStateBytes);
#pragma unroll
            for (int j = 0; j < kPerLane; ++j)
                if (lane + 32 * j < kWords) pre[j] = g[lane + 32 * j];
            if (lane == 0 && q.mode == MG_MODE_STEP) {
                act0 = q.io.actions[(size_t)nxt * q.act_dims];
                if (q.act_dims > 1) act1 = q.io.actions[(size_t)nxt * q.act_dims + 1];
            }
        }
        if (lane == 0) P::logic(k, q, env, *st, a0, a1, *job, fr);
        __syncwarp();
        uint64_t* gw = reinterpret_cast<uint64_t*>(state + (size_t)env * kStateBytes);
#pragma unroll
        for (int j = 0; j < kPerLane; ++j)
            if (lane + 32 * j < kWords) gw[lane + 32 * j] = s64[lane + 32 * j];
        const int nframes = job->nframes;
        for (int fi = 0; fi < nframes; ++fi) {
            const Frame f = fr[fi];
            uint8_t* dst = (job->to_final[fi] ? q.io.final_obs : q.io.obs) + (size_t)env * MG_FRAME_BYTES;
            if (pending) {
                if (lane == 0) mg_bulk_wait_read();
                __syncwarp();
                P::restore(ctx, buf, view, k, prev);
                __syncwarp();
            }
            P::compose(ctx, buf, view, k, f);
            mg_fence_proxy_async();  // make this lane's patch visible to the async (TMA) proxy
            __syncwarp();
            if (lane == 0) {
                mg_bulk_store(dst, buf, MG_FRAME_BYTES);
                mg_bulk_commit();
            }
            pending = true;
            prev = f;
        }
        __syncwarp();  // slot is reused by the next env
    }
    if (pending && lane == 0) mg_bulk_wait_read();  // shared memory must outlive the bulk read
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

template <class P>
static int launch_env(const typename P::K& k, const void* assets, void* state, const uint64_t* rng4,
                      const mg_step_io* ios, int n_io, int n, int mode, int act_dims, void* stream, int default_wpc) {
    if (!assets || !state || !ios || n < 0 || n_io < 1) {
        mg_set_err("NULL pointer, negative n or empty io list");
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
    if (((uintptr_t)assets & 127) || ((uintptr_t)state & 7)) {
        mg_set_err("alignment: assets 128 B, state 8 B");
        return MG_EINVAL;
    }
    if (n == 0) return 0;
    int dev = 0, sms = 0, smem_max = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        sms <= 0) {
        cudaGetLastError();
        mg_set_err("no usable CUDA device");
        return MG_ENODEV;
    }
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    constexpr int kSlot =
        (((int)sizeof(typename P::State) + 2 * (int)sizeof(typename P::Frame) + (int)sizeof(MgFrameJob) + 16 + 127) / 128) * 128;
    const int per_warp = MG_BUF_STRIDE + kSlot;
    int wpc = ios->warps_per_cta ? (int)ios->warps_per_cta : default_wpc;
    if (wpc > 10) wpc = 10;
    while (wpc > 1 && wpc * per_warp > smem_max) wpc--;
    int need = (n + sms - 1) / sms;  // spread small batches over all SMs
    if (wpc > need) wpc = need;
    if (wpc < 1) wpc = 1;
    int ctas = (n + wpc - 1) / wpc;
    if (ctas > sms) ctas = sms;
    size_t smem = (size_t)wpc * per_warp;
    // what is left of the SM's shared memory (up to 48 KB) stages the asset header + sprite pixel lists
    int stage_cap = smem_max - (int)smem;
    if (stage_cap > 48 * 1024) stage_cap = 48 * 1024;
    stage_cap &= ~127;
    smem += (size_t)stage_cap;
    cudaError_t e = cudaFuncSetAttribute(mg_env_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        mg_set_err(cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(wpc * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    // programmatic dependent launch: this kernel's prologue (asset staging, template loads) may overlap
    // the tail of the previous kernel in the stream; it calls griddepcontrol.wait before touching env data
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    for (int t = 0; t < n_io; ++t) {
        MgKIO q;
        q.io = ios[t];
        q.rng4 = rng4;
        q.mode = mode;
        q.n = n;
        q.stage_cap = stage_cap;
        q.act_dims = act_dims;
        e = cudaLaunchKernelEx(&cfg, mg_env_kernel<P>, k, q, (const uint8_t*)assets, (uint8_t*)state);
        if (e != cudaSuccess) {
            mg_set_err(cudaGetErrorString(e));
            cudaGetLastError();
            return (int)e;
        }
    }
    return 0;
}

static int mm_act_dims(const mg_mm_params* p) { return p->variant == MG_MM_GRID ? 1 : 2; }

extern "C" int mg_mm_reset(const mg_mm_params* p, const void* assets, void* state, const uint64_t* rng4,
                           const mg_step_io* io, int n, void* stream) {
    if (mg_mm_check_params(p)) return MG_EINVAL;
    MgMMK k;
    mg_mm_derive(p, &k);
    return launch_env<MMPolicy>(k, assets, state, rng4, io, 1, n, MG_MODE_RESET, mm_act_dims(p), stream, 10);
}
extern "C" int mg_mm_step(const mg_mm_params* p, const void* assets, void* state, const mg_step_io* io, int n,
                          void* stream) {
    return mg_mm_step_many(p, assets, state, io, 1, n, stream);
}
extern "C" int mg_mm_step_many(const mg_mm_params* p, const void* assets, void* state, const mg_step_io* ios,
                               int n_steps, int n, void* stream) {
    if (mg_mm_check_params(p)) return MG_EINVAL;
    MgMMK k;
    mg_mm_derive(p, &k);
    return launch_env<MMPolicy>(k, assets, state, nullptr, ios, n_steps, n, MG_MODE_STEP, mm_act_dims(p), stream, 10);
}

extern "C" int mg_mp_reset(const mg_mp_params* p, const void* assets, void* state, const uint64_t* rng4,
                           const mg_step_io* io, int n, void* stream) {
    if (mg_mp_check_params(p)) return MG_EINVAL;
    MgMPK k;
    mg_mp_derive(p, &k);
    return launch_env<MPPolicy>(k, assets, state, rng4, io, 1, n, MG_MODE_RESET, 2, stream, 10);
}
extern "C" int mg_mp_step(const mg_mp_params* p, const void* assets, void* state, const mg_step_io* io, int n,
                          void* stream) {
    return mg_mp_step_many(p, assets, state, io, 1, n, stream);
}
extern "C" int mg_mp_step_many(const mg_mp_params* p, const void* assets, void* state, const mg_step_io* ios,
                               int n_steps, int n, void* stream) {
    if (mg_mp_check_params(p)) return MG_EINVAL;
    MgMPK k;
    mg_mp_derive(p, &k);
    return launch_env<MPPolicy>(k, assets, state, nullptr, ios, n_steps, n, MG_MODE_STEP, 2, stream, 10);
}

// Searing Spotlights: not built yet — the symbols fail loudly.
extern "C" int mg_ss_check_params(const mg_ss_params*) {
    mg_set_err("SearingSpotlights kernels are not built yet");
    return MG_EINVAL;
}
extern "C" long mg_ss_build_assets(const mg_ss_params*, void*, size_t) {
    mg_set_err("SearingSpotlights kernels are not built yet");
    return MG_EINVAL;
}
extern "C" int mg_ss_reset(const mg_ss_params*, const void*, void*, const uint64_t*, const mg_step_io*, int, void*) {
    mg_set_err("SearingSpotlights kernels are not built yet");
    return MG_EINVAL;
}
extern "C" int mg_ss_step(const mg_ss_params*, const void*, void*, const mg_step_io*, int, void*) {
    mg_set_err("SearingSpotlights kernels are not built yet");
    return MG_EINVAL;
}
extern "C" int mg_ss_step_many(const mg_ss_params*, const void*, void*, const mg_step_io*, int, int, void*) {
    mg_set_err("SearingSpotlights kernels are not built yet");
    return MG_EINVAL;
}
