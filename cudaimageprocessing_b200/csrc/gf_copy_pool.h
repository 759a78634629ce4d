// gf_copy_pool.h -- a few persistent threads that copy one block of host memory together (pure C++17, no CUDA).
// Used by gf_guided_gray_host to stage PAGEABLE caller buffers through pinned planes (gf_host.inl); tested on its own
// by tests/test_copy_pool.py.
#pragma once
#include <condition_variable>
#include <cstddef>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

class GfCopyPool {
public:
    static GfCopyPool& get() { static GfCopyPool* p = new GfCopyPool(); return *p; }     // never destroyed: workers outlive main()
    void copy(void* d, const void* s, size_t n, int threads)
    {
        if (threads < 1) threads = 1;
        if (threads > kMaxThreads) threads = kMaxThreads;
        if (n < ((size_t)1 << 20) || threads == 1) { std::memcpy(d, s, n); return; }
        std::lock_guard<std::mutex> call(call_mu_);                                 // one job at a time (pipes of several devices share the pool)
        {
            std::lock_guard<std::mutex> lk(mu_);
            while ((int)workers_.size() < threads - 1) {
                const int idx = (int)workers_.size();
                workers_.emplace_back([this, idx] { run(idx); });
                workers_.back().detach();
            }
            d_ = (char*)d; s_ = (const char*)s; n_ = n; parts_ = threads; pending_ = threads - 1;
            ++gen_;
        }
        cv_job_.notify_all();
        slice(threads - 1);                                                         // the caller takes the last part
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [this] { return pending_ == 0; });
    }

private:
    static const int kMaxThreads = 16;
    void slice(int i)
    {
        const size_t a = (n_ * (size_t)i / parts_) & ~(size_t)63, b = i == parts_ - 1 ? n_ : ((n_ * (size_t)(i + 1) / parts_) & ~(size_t)63);
        if (b > a) std::memcpy(d_ + a, s_ + a, b - a);
    }
    void run(int idx)
    {
        unsigned long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_job_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (idx >= parts_ - 1) continue;                                    // this job uses fewer threads
            }
            slice(idx);
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) cv_done_.notify_one();
        }
    }
    std::mutex mu_, call_mu_;
    std::condition_variable cv_job_, cv_done_;
    std::vector<std::thread> workers_;
    char* d_ = nullptr; const char* s_ = nullptr; size_t n_ = 0;
    int parts_ = 1, pending_ = 0;
    unsigned long gen_ = 0;
};

