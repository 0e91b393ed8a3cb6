// copy_pool.h — host threads that move PAGEABLE caller memory to and from the pinned staging buffers of the host executor
// (host_exec.cuh). Plain C++ (no CUDA): tests/copy_pool_test.cpp exercises it on the CPU.
#pragma once
#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace ib200 {

constexpr int kCopyMaxArrays = 9;  // coordinate arrays of one chunk (kMaxNd) + 1

// A copy is cut into 512 KiB pieces that the pool's threads and the calling thread pull from a shared counter, so the pool
// never deadlocks, several callers (one per GPU) share it, and a pool of zero threads still copies.
class CopyPool {
public:
    static CopyPool& get() {
        static CopyPool pool;
        return pool;
    }
    int threads() const { return static_cast<int>(workers_.size()); }

    void copy(void* dst, const void* src, size_t bytes) {
        void* d[1] = {dst};
        const void* s[1] = {src};
        copy_many(1, d, s, bytes);
    }

    // k copies of `bytes` each (the coordinate arrays of one chunk) as ONE job: one join instead of k.
    void copy_many(int k, void* const* dst, const void* const* src, size_t bytes) {
        if (k <= 0 || bytes == 0) return;
        Job job;
        job.k = k;
        for (int j = 0; j < k; ++j) {
            job.dst[j] = static_cast<char*>(dst[j]);
            job.src[j] = static_cast<const char*>(src[j]);
        }
        job.bytes = bytes;
        job.per = (bytes + kPiece - 1) / kPiece;
        job.pieces = job.per * static_cast<size_t>(k);
        if (job.pieces <= 2 || workers_.empty()) {
            work_on(job);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            jobs_.push_back(&job);
        }
        cv_.notify_all();
        work_on(job);  // the caller takes pieces as well
        {
            std::lock_guard<std::mutex> lk(mu_);  // nobody may pick the job up any more
            jobs_.erase(std::remove(jobs_.begin(), jobs_.end(), &job), jobs_.end());
        }
        while (job.done.load(std::memory_order_acquire) != job.pieces || job.users.load(std::memory_order_acquire) != 0)
            std::this_thread::yield();
    }

private:
    static constexpr size_t kPiece = size_t(512) << 10;
    struct Job {
        int k = 0;
        char* dst[kCopyMaxArrays];
        const char* src[kCopyMaxArrays];
        size_t bytes = 0, per = 0, pieces = 0;
        std::atomic<size_t> next{0}, done{0};
        std::atomic<int> users{0};  // pool threads currently inside work_on(this job)
    };
    // Large copies with non-temporal stores: the destination (a staging buffer the DMA engine reads next, or the caller's
    // output array) is not read by this core again, and a streaming store skips the read-for-ownership of every line.
    static void stream_copy(char* dst, const char* src, size_t len) {
        static const bool nt = [] {
            const char* e = getenv("INTERPN_B200_COPY_NT");
            return !(e && e[0] == '0');
        }();
        if (!nt || len < 4096) {
            memcpy(dst, src, len);
            return;
        }
        const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
        if (head) {
            memcpy(dst, src, head);
            dst += head; src += head; len -= head;
        }
        size_t i = 0;
        for (; i + 64 <= len; i += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
        }
        _mm_sfence();
        if (i < len) memcpy(dst + i, src + i, len - i);
    }
    static void work_on(Job& j) {
        for (;;) {
            const size_t p = j.next.fetch_add(1, std::memory_order_relaxed);
            if (p >= j.pieces) return;
            const size_t seg = p / j.per, lo = (p % j.per) * kPiece, len = std::min(kPiece, j.bytes - lo);
            stream_copy(j.dst[seg] + lo, j.src[seg] + lo, len);
            j.done.fetch_add(1, std::memory_order_release);
        }
    }
    CopyPool() {
        int n = 0;
        if (const char* e = getenv("INTERPN_B200_COPY_THREADS")) n = atoi(e) - 1;
        else {
            const int hw = static_cast<int>(std::thread::hardware_concurrency());
            n = std::max(2, std::min(16, hw - 2)) - 1;  // the calling thread is one of the copiers
        }
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void loop() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            Job* j = nullptr;
            cv_.wait(lk, [&] {
                if (stop_) return true;
                for (Job* k : jobs_)
                    if (k->next.load(std::memory_order_relaxed) < k->pieces) {
                        j = k;
                        return true;
                    }
                return false;
            });
            if (stop_) return;
            j->users.fetch_add(1, std::memory_order_relaxed);  // under the lock: the owner cannot have removed the job yet
            lk.unlock();
            work_on(*j);
            j->users.fetch_sub(1, std::memory_order_release);
            lk.lock();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Job*> jobs_;
    std::vector<std::thread> workers_;
    bool stop_ = false;
};

}  // namespace ib200
