// mg_codec.cu — observation transport to HOST buffers: a lossless delta code over PCIe.
//
// The reference hands every observation to the caller as a host array (`surfarray.array3d(...)`,
// memory_gym/mortar_mayhem.py:270,367). Copying the full batch every step is bound by the PCIe link
// (21 168 B per env-step: 54 GB/s = 2.6 M env-steps/s per GPU) although consecutive frames of an env differ
// in a few hundred pixels. This file moves only what changed:
//
//   device   mg_obs_delta_encode: one warp per env compares the new frame with the frame the host already holds
//            (`prev`, device memory) in 16-byte chunks (1323 per frame), writes a 1323-bit change mask + count and
//            the changed chunks, densely packed, straight into PINNED HOST memory (zero-copy stores over PCIe: the
//            number of bytes that cross the link is the number of bytes that changed), and brings `prev` up to date.
//   host     mg_obs_delta_decode (mg_codec_host.cpp): worker threads patch the changed chunks into the caller's
//            persistent uint8[n][84][84][3] host array, which then equals the device observation bit for bit.
//
// Stage layout (host, pinned, 16-byte aligned): n headers of MG_DELTA_HDR_BYTES = {uint32 count, uint32 mask[42],
// pad}, then n payload regions of MG_OBS_BYTES (env e's changed chunks in ascending chunk order at the start of
// region e). A pure transport code: every frame is rendered on the GPU; the host only copies bytes.
//
// mg_obs_delta_scatter goes one step further when the caller's host array itself is pinned: the changed chunks are
// written to their final place in it, and the host does nothing at all.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/memgym_b200.h"
#include "mg_common.h"
#include "mg_device_host.h"

void mg_set_err(const char* msg);  // mg_params_host.cpp

namespace {
constexpr int kChunks = MG_FRAME_VEC16;            // 1323
constexpr int kIters = (kChunks + 31) / 32;        // 42 warp passes per frame
constexpr int kWarps = 16;                         // warps per CTA
static_assert(kIters <= 64, "two mask words per lane");
static_assert(MG_DELTA_HDR_BYTES >= 4 * (1 + kIters) && MG_DELTA_HDR_BYTES % 16 == 0, "header layout");

__device__ __forceinline__ uint4 ld16(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }

__global__ void __launch_bounds__(32 * kWarps)
    mg_obs_delta_kernel(const uint8_t* __restrict__ obs, uint8_t* __restrict__ prev, int n, uint8_t* __restrict__ stage,
                        int full) {
    const int lane = threadIdx.x & 31;
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    uint8_t* const payload0 = stage + (size_t)n * MG_DELTA_HDR_BYTES;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the step kernel's frames are visible
    for (int e = warp; e < n; e += nwarps) {
        const uint8_t* cur = obs + (size_t)e * MG_FRAME_BYTES;
        uint8_t* old = prev + (size_t)e * MG_FRAME_BYTES;
        uint8_t* out = payload0 + (size_t)e * MG_FRAME_BYTES;
        uint32_t m0 = 0, m1 = 0, count = 0;
        // four passes in flight: eight independent 128-bit loads per lane before the first compare
        for (int it0 = 0; it0 < kIters; it0 += 4) {
            uint4 a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = (it0 + u) * 32 + lane;
                if (it0 + u < kIters && i < kChunks) {
                    a[u] = ld16(cur + 16 * i);
                    b[u] = ld16(old + 16 * i);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int it = it0 + u, i = it * 32 + lane;
                if (it >= kIters) break;
                const bool valid = i < kChunks;
                const bool diff = valid && (full || a[u].x != b[u].x || a[u].y != b[u].y || a[u].z != b[u].z || a[u].w != b[u].w);
                const uint32_t bal = __ballot_sync(0xffffffffu, diff);
                if (diff) {
                    const uint32_t rank = count + __popc(bal & ((1u << lane) - 1u));
                    *reinterpret_cast<uint4*>(out + 16 * rank) = a[u];  // host memory: PCIe write
                    *reinterpret_cast<uint4*>(old + 16 * i) = a[u];
                }
                count += __popc(bal);
                if ((it & 31) == lane) {
                    if (it < 32)
                        m0 = bal;
                    else
                        m1 = bal;
                }
            }
        }
        uint32_t* hdr = reinterpret_cast<uint32_t*>(stage + (size_t)e * MG_DELTA_HDR_BYTES);
        hdr[1 + lane] = m0;
        if (lane < kIters - 32) hdr[33 + lane] = m1;
        if (lane == 0) hdr[0] = count;
    }
}

// The same comparison, but the changed chunks go STRAIGHT TO THEIR PLACE in the caller's host array (pinned, device-
// addressable): no stage, no header, nothing left to do on the host. `counter` (device memory, nullable) accumulates the
// payload bytes written.
__global__ void __launch_bounds__(32 * kWarps)
    mg_obs_delta_scatter_kernel(const uint8_t* __restrict__ obs, uint8_t* __restrict__ prev, int n, uint8_t* __restrict__ host_obs,
                                int full, unsigned long long* __restrict__ counter) {
    const int lane = threadIdx.x & 31;
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the step kernel's frames are visible
    uint32_t total = 0;
    for (int e = warp; e < n; e += nwarps) {
        const uint8_t* cur = obs + (size_t)e * MG_FRAME_BYTES;
        uint8_t* old = prev + (size_t)e * MG_FRAME_BYTES;
        uint8_t* out = host_obs + (size_t)e * MG_FRAME_BYTES;
        for (int it0 = 0; it0 < kIters; it0 += 4) {
            uint4 a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = (it0 + u) * 32 + lane;
                if (it0 + u < kIters && i < kChunks) {
                    a[u] = ld16(cur + 16 * i);
                    b[u] = ld16(old + 16 * i);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int it = it0 + u, i = it * 32 + lane;
                if (it >= kIters) break;
                const bool diff = i < kChunks && (full || a[u].x != b[u].x || a[u].y != b[u].y || a[u].z != b[u].z || a[u].w != b[u].w);
                if (diff) {
                    *reinterpret_cast<uint4*>(out + 16 * i) = a[u];  // host memory: PCIe write
                    *reinterpret_cast<uint4*>(old + 16 * i) = a[u];
                    total++;
                }
            }
        }
    }
    if (counter) {
        total = __reduce_add_sync(0xffffffffu, total);
        if (lane == 0 && total) atomicAdd(counter, 16ull * total);
    }
}
}  // namespace

extern "C" int mg_obs_delta_scatter(const uint8_t* obs, uint8_t* prev, int n, uint8_t* host_obs, int full, uint64_t* counter,
                                    void* stream) {
    if (!obs || !prev || !host_obs || n < 0) {
        mg_set_err("mg_obs_delta_scatter: NULL pointer or negative n");
        return MG_EINVAL;
    }
    if (((uintptr_t)obs & 15) || ((uintptr_t)prev & 15) || ((uintptr_t)host_obs & 15) || ((uintptr_t)counter & 7)) {
        mg_set_err("mg_obs_delta_scatter: obs, prev and host_obs must be 16-byte aligned (counter: 8)");
        return MG_EINVAL;
    }
    if (n == 0) return 0;
    MgDeviceGuard guard(obs);
    const MgDevInfo* di = guard.ok ? mg_dev_info(guard.dev) : nullptr;
    if (!di) {
        mg_set_err("no usable CUDA device");
        return MG_ENODEV;
    }
    int ctas = (n + kWarps - 1) / kWarps;
    const int cap = 4 * di->sms;
    if (ctas > cap) ctas = cap;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(32 * kWarps);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, mg_obs_delta_scatter_kernel, obs, prev, n, host_obs, full,
                                       reinterpret_cast<unsigned long long*>(counter));
    if (e != cudaSuccess) {
        mg_set_err(cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    return 0;
}

extern "C" size_t mg_obs_delta_stage_bytes(int n) {
    return n < 0 ? 0 : (size_t)n * (MG_DELTA_HDR_BYTES + MG_OBS_BYTES);
}

extern "C" int mg_obs_delta_encode(const uint8_t* obs, uint8_t* prev, int n, void* stage, int full, void* stream) {
    if (!obs || !prev || !stage || n < 0) {
        mg_set_err("mg_obs_delta_encode: NULL pointer or negative n");
        return MG_EINVAL;
    }
    if (((uintptr_t)obs & 15) || ((uintptr_t)prev & 15) || ((uintptr_t)stage & 15)) {
        mg_set_err("mg_obs_delta_encode: obs, prev and stage must be 16-byte aligned");
        return MG_EINVAL;
    }
    if (n == 0) return 0;
    MgDeviceGuard guard(obs);
    const MgDevInfo* di = guard.ok ? mg_dev_info(guard.dev) : nullptr;
    if (!di) {
        mg_set_err("no usable CUDA device");
        return MG_ENODEV;
    }
    int ctas = (n + kWarps - 1) / kWarps;
    const int cap = 4 * di->sms;  // persistent: four CTAs of 16 warps per SM
    if (ctas > cap) ctas = cap;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(32 * kWarps);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, mg_obs_delta_kernel, obs, prev, n, (uint8_t*)stage, full);
    if (e != cudaSuccess) {
        mg_set_err(cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    return 0;
}
