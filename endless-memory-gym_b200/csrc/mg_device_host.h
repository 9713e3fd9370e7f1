// mg_device_host.h — host-only: which CUDA device a call runs on, and that device's static properties.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

// ---- device selection -------------------------------------------------------------------------------
// Every launch runs on the device that owns the env records, whatever the caller's current device is: the owner of
// `state` is looked up once per buffer (cudaPointerGetAttributes, remembered per thread) and made current for the
// duration of the call. `stream` must belong to that device (0 = its legacy default stream).
#define MG_MAX_DEVICES 64
struct MgDevInfo {
    int sms, smem_optin;
};
static inline const MgDevInfo* mg_dev_info(int dev) {
    static MgDevInfo info[MG_MAX_DEVICES];
    static std::atomic<int> known[MG_MAX_DEVICES];
    if (dev < 0 || dev >= MG_MAX_DEVICES) return nullptr;
    if (!known[dev].load(std::memory_order_acquire)) {
        int sms = 0, smem = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0 ||
            cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        info[dev].sms = sms;  // (racing initialisers write the same values)
        info[dev].smem_optin = smem;
        known[dev].store(1, std::memory_order_release);
    }
    return &info[dev];
}
struct MgDeviceGuard {
    int dev = -1, prev = -1;
    bool ok = false;
    explicit MgDeviceGuard(const void* state) {
        static thread_local const void* last_ptr = nullptr;
        static thread_local int last_dev = -1;
        if (cudaGetDevice(&prev) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        if (state == last_ptr && last_dev >= 0) {
            dev = last_dev;
        } else {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, state) != cudaSuccess || at.type != cudaMemoryTypeDevice) {
                cudaGetLastError();
                dev = prev;  // not a plain device allocation (managed memory, a sanitizer shim): stay where the caller is
            } else {
                dev = at.device;
            }
            last_ptr = state;
            last_dev = dev;
        }
        if (dev != prev && cudaSetDevice(dev) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        ok = true;
    }
    ~MgDeviceGuard() {
        if (ok && dev != prev) cudaSetDevice(prev);
    }
};

