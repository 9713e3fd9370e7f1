// mg_codec_host.cpp — host half of the observation transport (see mg_codec.cu): patches the chunks the device
// reported as changed into the caller's persistent uint8[n][84][84][3] array. No CUDA here; plain byte copies spread
// over a small pool of worker threads (the reference is single-threaded per env: its host cores are otherwise idle
// while the GPU steps). A pure transport decoder — it renders nothing.
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/memgym_b200.h"
#include "mg_common.h"

void mg_set_err(const char* msg);  // mg_params_host.cpp

namespace {
constexpr int kChunks = MG_FRAME_VEC16;
constexpr int kMaskWords = (kChunks + 31) / 32;

struct Job {
    const uint8_t* stage = nullptr;
    uint8_t* obs = nullptr;
    int n = 0;
    std::atomic<int> next{0};       // next block of envs to take
    std::atomic<int> bad{0};        // malformed headers seen
    std::atomic<uint64_t> bytes{0};  // payload bytes applied
};

// envs [lo, hi) of the job: returns payload bytes applied, or -1 for a malformed header
static int64_t decode_range(const Job& j, int lo, int hi) {
    const uint8_t* payload0 = j.stage + (size_t)j.n * MG_DELTA_HDR_BYTES;
    int64_t total = 0;
    for (int e = lo; e < hi; ++e) {
        const uint32_t* hdr = reinterpret_cast<const uint32_t*>(j.stage + (size_t)e * MG_DELTA_HDR_BYTES);
        const uint32_t count = hdr[0];
        if (count > (uint32_t)kChunks) return -1;
        if (count == 0) continue;
        const uint8_t* src = payload0 + (size_t)e * MG_FRAME_BYTES;
        uint8_t* dst = j.obs + (size_t)e * MG_FRAME_BYTES;
        if (count == (uint32_t)kChunks) {  // a whole new frame (first step, forced resync)
            memcpy(dst, src, MG_FRAME_BYTES);
            total += MG_FRAME_BYTES;
            continue;
        }
        uint32_t k = 0;
        for (int w = 0; w < kMaskWords; ++w) {
            uint32_t m = hdr[1 + w];
            while (m) {
                // a run of consecutive changed chunks is one contiguous copy on both sides
                const int b = __builtin_ctz(m);
                const uint32_t shifted = m >> b;
                const int run = (shifted == 0xFFFFFFFFu) ? 32 - b : __builtin_ctz(~shifted);
                memcpy(dst + 16 * (32 * w + b), src + 16 * (size_t)k, 16 * (size_t)run);
                k += (uint32_t)run;
                m = (run + b >= 32) ? 0u : (m & ~(((1u << run) - 1u) << b));
            }
        }
        if (k != count) return -1;
        total += 16 * (int64_t)count;
    }
    return total;
}

static void work(Job& j) {
    constexpr int kBlock = 16;  // envs per grab
    for (;;) {
        const int lo = j.next.fetch_add(kBlock, std::memory_order_relaxed);
        if (lo >= j.n) return;
        const int hi = lo + kBlock < j.n ? lo + kBlock : j.n;
        const int64_t got = decode_range(j, lo, hi);
        if (got < 0)
            j.bad.fetch_add(1, std::memory_order_relaxed);
        else
            j.bytes.fetch_add((uint64_t)got, std::memory_order_relaxed);
    }
}

static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
}

// A fixed pool of workers; the calling thread works too. A host-pipe loop hands over a job every ~150 us, so a worker
// SPINS for the next job for a short while (kSpinNs) before it parks on a condition variable: a wake-up through the
// kernel costs 30-50 us, a quarter of the job. Every worker checks in for every job, so none can touch a job after
// run() has returned.
class Pool {
   public:
    static Pool& get() {
        static Pool p;
        return p;
    }
    void run(Job& j, int threads) {
        std::lock_guard<std::mutex> serial(run_mu_);  // one decode at a time per process
        if (threads < 1) threads = 1;
        if (threads > 256) threads = 256;
        grow(threads - 1);
        job_ = &j;
        want_ = threads - 1;
        pending_.store((int)workers_.size(), std::memory_order_relaxed);
        epoch_.fetch_add(1);  // (sequentially consistent with parked_: either this thread sees the parked worker or the
        if (parked_.load() > 0) {  //  worker sees the new epoch before it sleeps)
            std::lock_guard<std::mutex> g(mu_);
            cv_.notify_all();
        }
        work(j);
        while (pending_.load(std::memory_order_acquire) != 0) cpu_relax();
        job_ = nullptr;
    }

   private:
    static constexpr long kSpinNs = 400 * 1000;
    Pool() = default;
    ~Pool() {
        stop_.store(true, std::memory_order_release);
        {
            std::lock_guard<std::mutex> g(mu_);
            cv_.notify_all();
        }
        for (auto& t : workers_) t.join();
    }
    void grow(int n) {
        while ((int)workers_.size() < n) {
            const int id = (int)workers_.size();
            const uint64_t start = epoch_.load(std::memory_order_relaxed);  // (a new worker joins from the NEXT job on)
            workers_.emplace_back([this, id, start] { loop(id, start); });
        }
    }
    void loop(int id, uint64_t seen) {
        for (;;) {
            uint64_t e;
            auto t0 = std::chrono::steady_clock::now();
            int spins = 0;
            while ((e = epoch_.load(std::memory_order_acquire)) == seen) {
                if (stop_.load(std::memory_order_acquire)) return;
                cpu_relax();
                if (++spins < 256) continue;
                spins = 0;
                if (std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count() < kSpinNs)
                    continue;
                std::unique_lock<std::mutex> g(mu_);
                parked_.fetch_add(1);
                cv_.wait_for(g, std::chrono::milliseconds(50), [&] { return stop_.load() || epoch_.load() != seen; });
                parked_.fetch_sub(1, std::memory_order_release);
                t0 = std::chrono::steady_clock::now();
            }
            seen = e;
            if (id < want_) work(*job_);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    std::mutex mu_, run_mu_;
    std::condition_variable cv_;
    std::vector<std::thread> workers_;
    Job* job_ = nullptr;
    int want_ = 0;
    std::atomic<uint64_t> epoch_{0};
    std::atomic<int> pending_{0}, parked_{0};
    std::atomic<bool> stop_{false};
};
}  // namespace

extern "C" int mg_obs_delta_decode(const void* stage, uint8_t* host_obs, int n, int threads, uint64_t* bytes_out) {
    if (!stage || !host_obs || n < 0) {
        mg_set_err("mg_obs_delta_decode: NULL pointer or negative n");
        return MG_EINVAL;
    }
    Job j;
    j.stage = (const uint8_t*)stage;
    j.obs = host_obs;
    j.n = n;
    if (threads <= 1 || n < 64) {
        work(j);
    } else {
        Pool::get().run(j, threads);
    }
    if (j.bad.load()) {
        mg_set_err("mg_obs_delta_decode: malformed stage header (was the step's event waited for?)");
        return MG_EINVAL;
    }
    if (bytes_out) *bytes_out = j.bytes.load() + (uint64_t)n * MG_DELTA_HDR_BYTES;
    return 0;
}
