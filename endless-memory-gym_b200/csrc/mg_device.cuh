// mg_device.cuh — device-side helpers shared by the kernels: PTX wrappers (TMA bulk copies, mbarrier, proxy fence),
// the per-CTA header words and the execution context (`Ctx`) the frame / logic headers are written against.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/memgym_b200.h"
#include "mg_common.h"

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
    // warp reductions (single-warp teams only): every lane gets the result
    __device__ __forceinline__ uint32_t reduce_min(uint32_t v) { return __reduce_min_sync(0xffffffffu, v); }
    __device__ __forceinline__ uint32_t reduce_or(uint32_t v) { return __reduce_or_sync(0xffffffffu, v); }
    __device__ __forceinline__ uint32_t shfl(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
    // shared-memory hand-over between warps of a CTA (teams <-> helper warps)
    __device__ __forceinline__ uint32_t cas_u32(uint32_t* p, uint32_t cmp, uint32_t val) { return atomicCAS(p, cmp, val); }
    __device__ __forceinline__ void fence_block() { __threadfence_block(); }
    __device__ __forceinline__ void fence_device() { __threadfence(); }
    __device__ __forceinline__ void nanosleep(unsigned ns) { __nanosleep(ns); }
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

