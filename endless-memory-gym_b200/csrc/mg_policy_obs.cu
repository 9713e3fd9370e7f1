// mg_policy_obs.cu — the consumer step right after the env step: the batch of uint8
// [x][y][c] observations -> the policy's input tensor, bf16 [n][3][84][84] = obs.permute(0, 3, 1, 2) / 255.
//
// What every training loop on top of the reference does on the host (`obs.astype(float32) / 255`, channels
// first). Here: a persistent streaming kernel, one frame per CTA iteration. The frame is brought into shared
// memory by ONE TMA bulk load (double-buffered: the next frame lands while this one is converted), de-interleaved
// (both layouts enumerate the pixels in the same order, so out[c][p] = in[3 p + c]), mapped through a 256-entry
// table of round-to-nearest-even bf16(v / 255.0f) — bit-identical to torch's float32 division + cast — assembled in
// shared memory and sent off as ONE TMA bulk store per frame (double-buffered too; bulk stores reach a higher write
// bandwidth than LSU stores on this part). HBM-bound: 21 168 B read (usually still in L2 from the step kernel) +
// 42 336 B written per env.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/memgym_b200.h"
#include "mg_common.h"
#include "mg_device_host.h"

void mg_set_err(const char* msg);  // mg_params_host.cpp

namespace {
constexpr int kPix = MG_W * MG_W;            // 7056
constexpr int kChunksPerPlane = kPix / 8;    // 882 16-byte chunks of 8 bf16
constexpr int kChunks = 3 * kChunksPerPlane; // 2646
constexpr int kBuf = 21248;                  // frame bytes rounded up to 128

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kOutBytes = kChunks * 16;     // 42 336
constexpr int kOutBuf = 42368;              // rounded up to 128
constexpr int kThreads = 512;
constexpr int kLutCopies = 16;              // lanes l and l + 16 share a copy: at worst a 2-way bank conflict
constexpr int kLutBytes = 256 * kLutCopies * 4;  // 16 KB

__global__ void __launch_bounds__(kThreads, 2) mg_policy_obs_kernel(const uint8_t* __restrict__ obs, uint8_t* __restrict__ out, int n) {
    // shared memory: in[2] (TMA bulk loads) | out (the converted frame, leaves by a TMA bulk store) | table | barriers
    // = 101 KB: two CTAs per SM fill each other's pipeline bubbles
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* const outb = smem + 2 * kBuf;
    // the 256-entry table in kLutCopies interleaved copies (entry v of copy l at word kLutCopies v + l)
    uint32_t* lut = reinterpret_cast<uint32_t*>(smem + 2 * kBuf + kOutBuf);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kBuf + kOutBuf + kLutBytes);
    for (int i = threadIdx.x; i < 256 * kLutCopies; i += blockDim.x)
        lut[i] = __bfloat16_as_ushort(__float2bfloat16_rn(__fdiv_rn((float)(i / kLutCopies), 255.0f)));
    const uint32_t* const my_lut = lut + (threadIdx.x & (kLutCopies - 1));
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + b)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the step kernel's observations are visible
    const auto load = [&](int frame, int b) {  // thread 0: one TMA bulk load of the whole frame
        const uint32_t bar = smem_u32(bars + b);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(MG_FRAME_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem + b * kBuf)),
                     "l"(__cvta_generic_to_global(obs + (size_t)frame * MG_FRAME_BYTES)), "r"(MG_FRAME_BYTES), "r"(bar)
                     : "memory");
    };
    int frame = blockIdx.x;
    if (frame < n && threadIdx.x == 0) load(frame, 0);
    uint32_t phases = 0u;  // bit b = parity to wait for on buffer b (no indexed local array)
    for (int it = 0; frame < n; ++it, frame += gridDim.x) {
        const int b = it & 1, next = frame + gridDim.x;
        if (threadIdx.x == 0) {
            if (next < n) load(next, b ^ 1);  // in[b^1] was drained before the last __syncthreads
            // `out` was handed to the TMA engine one iteration ago: its read must be over before it is rewritten
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        {
            uint32_t ok = 0;
            const uint32_t bar = smem_u32(bars + b);
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok)
                             : "r"(bar), "r"((phases >> b) & 1u)
                             : "memory");
            phases ^= 1u << b;
        }
        __syncthreads();  // thread 0's wait_group.read covers everyone
        const uint8_t* src = smem + b * kBuf;
        uint4* dst = reinterpret_cast<uint4*>(outb);
        // one thread = one group of 8 pixels: 24 input bytes (three aligned 64-bit loads) -> one 16-byte chunk per plane
        for (int g = threadIdx.x; g < kChunksPerPlane; g += blockDim.x) {
            const uint2* s = reinterpret_cast<const uint2*>(src + 24 * g);
            const uint2 a0 = s[0], a1 = s[1], a2 = s[2];
            const uint32_t w[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y};
            uint32_t o[3][4];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) {  // pixels 2j, 2j+1 of channel c: bytes 6j + c, 6j + 3 + c of the group
                    const int b0 = 6 * j + c, b1 = b0 + 3;
                    const uint32_t v0 = (w[b0 >> 2] >> (8 * (b0 & 3))) & 255u, v1 = (w[b1 >> 2] >> (8 * (b1 & 3))) & 255u;
                    o[c][j] = my_lut[kLutCopies * v0] | (my_lut[kLutCopies * v1] << 16);
                }
#pragma unroll
            for (int c = 0; c < 3; ++c) dst[c * kChunksPerPlane + g] = make_uint4(o[c][0], o[c][1], o[c][2], o[c][3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of in[b] / writes of out[b] vs the TMA engine
        __syncthreads();
        if (threadIdx.x == 0) {  // the whole converted frame leaves as ONE bulk store (full lines, evict-first)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(
                             __cvta_generic_to_global(out + (size_t)frame * kOutBytes)),
                         "r"(smem_u32(dst)), "r"(kOutBytes), "l"(policy)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the reads
}
}  // namespace

extern "C" int mg_obs_to_policy_bf16(const uint8_t* obs, void* out, int n, void* stream) {
    if (!obs || !out || n < 0) {
        mg_set_err("mg_obs_to_policy_bf16: NULL pointer or negative n");
        return MG_EINVAL;
    }
    if (((uintptr_t)obs & 15) || ((uintptr_t)out & 15)) {
        mg_set_err("mg_obs_to_policy_bf16: obs and out must be 16-byte aligned");
        return MG_EINVAL;
    }
    if (n == 0) return 0;
    MgDeviceGuard guard(obs);
    const MgDevInfo* di = guard.ok ? mg_dev_info(guard.dev) : nullptr;
    if (!di) {
        mg_set_err("no usable CUDA device");
        return MG_ENODEV;
    }
    const int sms = di->sms;
    const size_t smem = 2 * kBuf + kOutBuf + kLutBytes + 16;
    static std::atomic<int> optin_set[MG_MAX_DEVICES];
    cudaError_t e = cudaSuccess;
    if (!optin_set[guard.dev].load(std::memory_order_acquire)) {
        e = cudaFuncSetAttribute(mg_policy_obs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) optin_set[guard.dev].store(1, std::memory_order_release);
    }
    if (e == cudaSuccess) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(n < 2 * sms ? n : 2 * sms);  // two persistent CTAs per SM
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, mg_policy_obs_kernel, obs, (uint8_t*)out, n);
    }
    if (e != cudaSuccess) {
        mg_set_err(cudaGetErrorString(e));
        cudaGetLastError();
        return (int)e;
    }
    return 0;
}
